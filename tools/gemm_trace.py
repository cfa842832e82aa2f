"""Where do the warp roles of the CTA-pair GEMM wait?  Per-role clock64 accounting (ops/csrc/gemm_tcgen05.cu TR_*)
for the GEMM shapes of one GPT-2 layer.  Output: one line per shape, cycles averaged over CTAs.

    python tools/gemm_trace.py            # on a B200
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.ops import functional as F, lib  # noqa: E402

B, T, C = 16, 1024, 768
M = B * T
dev = "cuda"


def r(*s):
    return (torch.randn(*s, device=dev) * 0.5).bfloat16()


def run(name, fn, iters=3):
    nsm = F.num_sms()
    tr = torch.zeros(nsm * 8, dtype=torch.int64, device=dev)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.load().aitj_gemm_set_trace(tr.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    lib.load().aitj_gemm_set_trace(None)
    t = tr.view(nsm, 8).cpu().double()
    lead = t[0::2]
    lead = lead[lead[:, 0] > 0]
    allc = t[t[:, 6] > 0]
    us = e0.elapsed_time(e1) * 1000 / iters
    tiles = lead[:, 7].mean().item()
    print(f"{name:44s} {us:7.1f} us | tiles/cluster {tiles:4.1f} | MMA total {lead[:,0].mean():8.0f} "
          f"wait_full(TMA) {lead[:,1].mean():8.0f} wait_tmem(epi) {lead[:,2].mean():8.0f} | "
          f"TMA wait_empty {t[:,3][t[:,6]>0].mean():8.0f} | EPI total {allc[:,6].mean():8.0f} "
          f"wait_full {allc[:,4].mean():8.0f} busy {allc[:,5].mean():8.0f} (per tile {allc[:,5].mean()/max(tiles,1):6.0f})",
          flush=True)


x = r(M, C)
x4 = r(M, 4 * C)
w_qkv, b_qkv = r(3 * C, C), r(3 * C)
w_proj, b_proj = r(C, C), r(C)
w_fc, b_fc = r(4 * C, C), r(4 * C)
w_fc2, b_fc2 = r(C, 4 * C), r(C)
o3 = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
o1 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
o4 = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
pre4 = r(M, 4 * C)
res = r(M, C)
d3 = r(M, 3 * C)
gq = torch.zeros(3 * C, C, device=dev)
g4 = torch.zeros(4 * C, C, device=dev)
g42 = torch.zeros(C, 4 * C, device=dev)
BN = 512

run("qkv fwd  (bias)", lambda: F.gemm(x, w_qkv, o3, bias=b_qkv, block_n=BN))
run("qkv fwd  (no epilogue math)", lambda: F.gemm(x, w_qkv, o3, block_n=BN))
run("qkv fwd  (epilogue skipped: main loop only)", lambda: F.gemm(x, w_qkv, o3, block_n=BN, _debug_skip_epilogue=True))
run("qkv fwd  (tcgen05.ld only)", lambda: F.gemm(x, w_qkv, o3, block_n=BN, _debug_skip_epilogue="ldonly"))
run("qkv fwd  (ld + staging, no TMA store)", lambda: F.gemm(x, w_qkv, o3, block_n=BN, _debug_skip_epilogue="notma"))
run("fc fwd   (epilogue skipped)", lambda: F.gemm(x, w_fc, o4, block_n=BN, _debug_skip_epilogue=True))
run("fc dgrad (epilogue skipped)", lambda: F.gemm(x4, w_fc, o1, b_mn=True, block_n=BN, _debug_skip_epilogue=True))
run("proj fwd (bias+residual)", lambda: F.gemm(x, w_proj, o1, bias=b_proj, residual=res, block_n=BN))
run("fc fwd   (bias+gelu+save_pre)", lambda: F.gemm(x, w_fc, o4, bias=b_fc, gelu=True, save_pre=True, aux=pre4, block_n=BN))
run("fc fwd   (bias only)", lambda: F.gemm(x, w_fc, o4, bias=b_fc, block_n=BN))
run("fc2 fwd  (bias+residual)", lambda: F.gemm(x4, w_fc2, o1, bias=b_fc2, residual=res, block_n=BN))
run("fc2 dgrad (dgelu)", lambda: F.gemm(x, w_fc2, o4, b_mn=True, dgelu=True, aux=pre4, block_n=BN))
run("fc2 dgrad (plain)", lambda: F.gemm(x, w_fc2, o4, b_mn=True, block_n=BN))
run("fc dgrad", lambda: F.gemm(x4, w_fc, o1, b_mn=True, block_n=BN))
run("qkv dgrad", lambda: F.gemm(d3, w_qkv, o1, b_mn=True, block_n=BN))
for sk in (1, 2, 4):
    run(f"fc wgrad sk={sk}", lambda: F.gemm(x4, x, g4, a_mn=True, b_mn=True, accumulate=True, split_k=sk, block_n=BN))
    run(f"qkv wgrad sk={sk}", lambda: F.gemm(d3, x, gq, a_mn=True, b_mn=True, accumulate=True, split_k=sk, block_n=BN))
run("fc2 wgrad sk=4", lambda: F.gemm(x, x4, g42, a_mn=True, b_mn=True, accumulate=True, split_k=4, block_n=BN))
big_a, big_b = r(8192, 8192), r(8192, 8192)
big_o = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
run("square 8192", lambda: F.gemm(big_a, big_b, big_o, block_n=BN))
run("square 8192 (epilogue skipped)", lambda: F.gemm(big_a, big_b, big_o, block_n=BN, _debug_skip_epilogue=True))
