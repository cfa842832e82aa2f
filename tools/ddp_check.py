"""Multi-GPU numerics + timing check of the gradient paths (run under torchrun, one rank per GPU):
local grads + NCCL all-reduce  vs  fused multimem.red (GEMM->all-reduce) on the symmetric buffer."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.models.gpt2 import GPT2Config, GPT2Engine  # noqa: E402
from trainingjob_operator_b200.parallel.symm import SymmetricGradBuffer  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
cfg = GPT2Config(vocab_size=1000, n_layer=2, n_head=4, n_embd=256, block_size=128, name="t")
B, T = 4, 128
eng = GPT2Engine(cfg, B, T, dev, seed=3)
g = torch.Generator().manual_seed(100 + rank)
tok = torch.randint(0, cfg.vocab_size, (B * T,), generator=g).to(dev)
eng.tok.copy_(tok); eng.tgt.copy_(torch.roll(tok, -1))
eng.forward(); eng.backward()
ref = eng.params.g32.clone()
dist.all_reduce(ref)
out = {"rank": rank, "world": world}
sb = SymmetricGradBuffer(eng.params.total, dev)
out["mc_available"] = sb.available; out["reason"] = sb.reason
if sb.available:
    eng.params.attach_grad_buffer(sb.tensor, sb.multicast_ptr)
    torch.cuda.synchronize(); dist.barrier()
    eng.forward(); sb.barrier(); eng.backward(); sb.barrier()
    torch.cuda.synchronize()
    got = eng.params.g32
    err = float((got - ref).norm() / ref.norm())
    out["mc_rel_err_vs_nccl_sum"] = err
    out["ok"] = err < 2e-3
else:
    out["ok"] = True
allok = torch.tensor([1 if out["ok"] else 0], device=dev)
dist.all_reduce(allok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
sys.exit(0 if int(allok[0]) == 1 else 1)
