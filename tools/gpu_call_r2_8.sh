#!/bin/bash
# round 2, call 8 (2 GPUs): full GPU test suite; deferred peer move; bench N=2 with the e2e arm (warm pool)
set -u
O=gpurun_out/r2c8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"; tail -6 $O/pytest_gpu.txt | cut -c1-300
timeout 240 $TR --master-port 29681 tools/ddp_check.py --small > $O/ddp_check_small.json 2> $O/ddp_check_small.err; echo "ddp_check rc=$?"
python -c "
import json
d=json.loads(open('$O/ddp_check_small.json').read().strip().splitlines()[-1])
print(d['ok'], [(r['rs_grad_rel_err_owned'], r['rs_nothing_left_behind'], r['rs_p16_rel_err'], r['rs_grad_rel_err_step2']) for r in d['ranks']])" || tail -5 $O/ddp_check_small.err
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29682 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err
python - <<'PY'
import json
b=[json.loads(l) for l in open('gpurun_out/r2c8/breakdown_n2_rs.jsonl').read().strip().splitlines() if l.startswith('{')]
for x in b: print("phase", x["ms"], x["total_ms"]); print({k:v for k,v in x["kernels_ms"].items() if k.startswith("gemm")})
PY
timeout 600 $TR --master-port 29683 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.jsonl 2> $O/bench_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench_n2.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['allreduce'][:30], d['e2e'])" || tail -20 $O/bench_n2.err
for m in mnist resnet50 bert; do
  timeout 300 $TR --master-port 29684 bench.py --gpus 2 --steps 30 --warmup 5 --model $m --no-e2e > $O/bench_n2_$m.jsonl 2> $O/bench_n2_$m.err; echo "bench $m rc=$?"
  python -c "
import json
d=json.loads(open('$O/bench_n2_$m.jsonl').read().strip().splitlines()[-1])
print('$m', d['value'], d['ms_per_step'], d['config']['cuda_graph'], d['config']['graph_error'], d['config']['allreduce'][:20], d['config']['loss_first'], d['config']['loss_last'])" || tail -8 $O/bench_n2_$m.err
done
