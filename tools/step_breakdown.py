"""Where one training step goes, phase by phase, with CUDA events on the launching stream (eager launches, so the
numbers include launch gaps -- compare SHARES and the same phase across world sizes, not absolutes).

    python tools/step_breakdown.py                      # 1 GPU
    torchrun --nproc-per-node N ... tools/step_breakdown.py      # N GPUs, AITJ_ALLREDUCE=rs|nccl
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.models.gpt2 import GPT2Config, GPT2Engine  # noqa: E402
from trainingjob_operator_b200.runtime.trainer import EngineTrainer, SyntheticTokens  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = GPT2Engine(GPT2Config.small(), 16, 1024, dev, seed=0)
tr = EngineTrainer(eng, use_graph=False)
data = SyntheticTokens(50257, 16, 1024, n_batches=2, seed=1)
for _ in range(3):
    tr.step(*data.next())
torch.cuda.synchronize()
sh = tr.shard
names = ["forward", "backward", "barrier_after_bwd", "optimizer(+norm exchange, its barrier)", "barrier_after_opt"]
acc = {n: 0.0 for n in names}
STEPS = 10
for _ in range(STEPS):
    tok, tgt = data.next()
    eng.tok.copy_(tok, non_blocking=True); eng.tgt.copy_(tgt, non_blocking=True)
    tr.step_count += 1
    eng.set_step_scalars(3e-4, tr.step_count)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record(); eng.forward()
    ev[1].record(); eng.backward()
    if tr.reducer:
        tr.reducer.wait()
    ev[2].record()
    if sh is not None:
        sh.barrier()
    ev[3].record(); tr._optimizer()
    ev[4].record()
    if sh is not None:
        sh.barrier()
    ev[5].record()
    torch.cuda.synchronize()
    for i, n in enumerate(names):
        acc[n] += ev[i].elapsed_time(ev[i + 1])
# per-kernel device time of one more step (events around every launch of the kernel library)
from trainingjob_operator_b200.ops import lib as _lib  # noqa: E402

if world > 1:
    dist.barrier()
torch.cuda.synchronize()
_lib.profile_start()
tok, tgt = data.next()
eng.tok.copy_(tok, non_blocking=True); eng.tgt.copy_(tgt, non_blocking=True)
tr._device_step()
prof = _lib.profile_stop()
groups = {}
for label, (n, ms) in prof.items():
    key = label.split(" ")[0] if label.startswith("gemm:") else label
    g = groups.setdefault(key, [0, 0.0])
    g[0] += n; g[1] += ms
out = {"rank": rank, "world": world, "kernels_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(groups.items())},
       "wgrad_detail_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(prof.items()) if k.startswith("gemm:wgrad")},
       "allreduce": tr.allreduce_backend[:40],
       "ms": {n: round(v / STEPS, 3) for n, v in acc.items()}, "total_ms": round(sum(acc.values()) / STEPS, 3)}
if world > 1:
    allo = [None] * world
    dist.all_gather_object(allo, out)
    if rank == 0:
        for o in allo:
            print(json.dumps(o))
    dist.destroy_process_group()
else:
    print(json.dumps(out))
