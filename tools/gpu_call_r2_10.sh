#!/bin/bash
# round 2, call 10 (2 GPUs): e2e arm with peer-visible GPUs (rs inside the launched workers); elastic rescale of an rs job
set -u
O=gpurun_out/r2c10; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.jsonl 2> $O/bench_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench_n2.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['allreduce'][:30]); print(d['e2e'])" || tail -20 $O/bench_n2.err
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29702 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err
python - <<'PY'
import json
b=[json.loads(l) for l in open('gpurun_out/r2c10/breakdown_n2_rs.jsonl').read().strip().splitlines() if l.startswith('{')]
for x in b: print("phase", x["ms"], x["total_ms"]); print({k:v for k,v in x["kernels_ms"].items() if k.startswith("gemm")}); print(x["wgrad_detail_ms"])
PY
timeout 400 python tools/elastic_gpu_check.py gpt2-tiny 2 0 > $O/elastic_gpt2tiny_n2.log 2>&1; echo "elastic rc=$?"; grep '^{' $O/elastic_gpt2tiny_n2.log | tail -1 | cut -c1-1200; grep -h "allreduce\|rs unavailable\|Traceback\|Error" gpurun_out/elastic_logs_gpt2-tiny_n2_pool0/* 2>/dev/null | head -10
timeout 300 python -m pytest tests/test_gpu_runtime.py -q -m gpu -x 2>&1 | tail -3
