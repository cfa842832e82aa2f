"""NVLink bytes per training step and GPU, from the driver's per-link throughput counters (`nvidia-smi nvlink -gt d`),
around exactly K steps of the GPT-2 small DDP job (run under torchrun, one rank per GPU).  The owner-sharded gradient path
should move about (N-1)/N x 498 MB of fp32 gradients out of (and into) every GPU per step plus the multicast parameter
all-gather (249 MB / N out, (N-1)/N x 249 MB in); the NCCL ring all-reduce about 2 (N-1)/N x 498 MB each way.

    torchrun --nproc-per-node N ... tools/nvlink_probe.py [--steps 20]      (AITJ_ALLREDUCE=rs|nccl)
"""
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.models.gpt2 import GPT2Config, GPT2Engine  # noqa: E402
from trainingjob_operator_b200.runtime.trainer import EngineTrainer, SyntheticTokens  # noqa: E402


def counters():
    """{gpu: (tx_kib, rx_kib)} summed over links."""
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True, timeout=60).stdout
    res, gpu = {}, None
    for line in out.splitlines():
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            res[gpu] = [0, 0]
            continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            res[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return res, out[:400]


rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 20
eng = GPT2Engine(GPT2Config.small(), 16, 1024, dev, seed=0)
tr = EngineTrainer(eng)
data = SyntheticTokens(50257, 16, 1024, n_batches=2, seed=1)
for _ in range(5):
    tr.step(*data.next())
torch.cuda.synchronize(); dist.barrier()
before, raw = counters() if rank == 0 else ({}, "")
dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    tr.step(*data.next())
ev1.record()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    after, _ = counters()
    per = {g: {"tx_mb_per_step": round((after[g][0] - before[g][0]) * 1024 / 1e6 / steps, 1),
               "rx_mb_per_step": round((after[g][1] - before[g][1]) * 1024 / 1e6 / steps, 1)} for g in sorted(after)
           if g in before}
    grad_mb = eng.params.total * 4 / 1e6
    print(json.dumps({"world": world, "allreduce": tr.allreduce_backend[:60], "steps": steps,
                      "ms_per_step": round(ev0.elapsed_time(ev1) / steps, 3), "per_gpu": per,
                      "fp32_gradient_mb": round(grad_mb, 1), "bf16_parameter_mb": round(grad_mb / 2, 1),
                      "expected_rs_tx_mb": round(grad_mb * (world - 1) / world + grad_mb / 2 / world, 1),
                      "expected_rs_rx_mb": round(grad_mb * (world - 1) / world + grad_mb / 2 * (world - 1) / world, 1),
                      "expected_ring_allreduce_each_way_mb": round(2 * grad_mb * (world - 1) / world, 1),
                      "counter_sample": raw if not per else ""}))
dist.destroy_process_group()
