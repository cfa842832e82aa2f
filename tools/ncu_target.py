"""Tiny single-kernel driver for `ncu` captures (one GPU, a handful of launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.ops import functional as F  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "gemm_fc"
B, T, C = 16, 1024, 768
M = B * T
if what.startswith("gemm"):
    shapes = {"gemm_fc": (M, 4 * C, C, False, False), "gemm_square": (8192, 8192, 8192, False, False),
              "gemm_fc2_dgrad": (M, 4 * C, C, False, True), "gemm_wgrad": (4 * C, C, M, True, True)}
    shapes.update({"gemm_qkv_pair": (M, 3 * C, C, False, False), "gemm_square_pair": (8192, 8192, 8192, False, False),
                   "gemm_fcwgrad_pair": (4 * C, C, M, True, True)})
    m, n, k, a_mn, b_mn = shapes[what]
    bn = 512 if what.endswith("_pair") else 0
    a = (torch.randn(k, m, device="cuda") if a_mn else torch.randn(m, k, device="cuda")).bfloat16()
    b = (torch.randn(k, n, device="cuda") if b_mn else torch.randn(n, k, device="cuda")).bfloat16()
    acc = a_mn
    out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if acc else torch.bfloat16)
    for _ in range(6):
        F.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, accumulate=acc, split_k=2 if acc else 1, block_n=bn)
elif what == "xent":
    logits = torch.randn(M, 50304, device="cuda").bfloat16()
    tgt = torch.randint(0, 50257, (M,), device="cuda")
    loss = torch.empty(M, device="cuda")
    for _ in range(4):
        F.softmax_xent(logits, tgt, loss, 50257, 1.0 / M)
elif what == "attn":
    H = 12
    qkv = torch.randn(M, 3 * C, device="cuda").bfloat16()
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda")
    for _ in range(6):
        F.attention_fwd(qkv, out, lse, B, T, H, causal=True)
elif what == "ln_bwd":
    x = torch.randn(M, C, device="cuda").bfloat16()
    dy = torch.randn(M, C, device="cuda").bfloat16()
    g = torch.ones(C, device="cuda").bfloat16()
    mean = torch.zeros(M, device="cuda"); rstd = torch.ones(M, device="cuda")
    dx = torch.empty_like(x); dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    for _ in range(6):
        F.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, dres=x)
elif what == "adamw":
    n = 124_475_904 // 256 * 256
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
    m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    mask = torch.ones(n // 256, dtype=torch.uint8, device="cuda")
    for _ in range(4):
        F.adamw(p, g, m1, v1, p16, mask, lr=1e-4, step=3)
torch.cuda.synchronize()
