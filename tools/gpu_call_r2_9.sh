#!/bin/bash
# round 2, call 9 (2 GPUs): bulk (TMA) move of finished gradient blocks; e2e arm detail; fault-tolerant recovery over NCCL
set -u
O=gpurun_out/r2c9; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29691 tools/ddp_check.py --small > $O/ddp_check_small.json 2> $O/ddp_check_small.err; echo "ddp_check rc=$?"
python -c "
import json
d=json.loads(open('$O/ddp_check_small.json').read().strip().splitlines()[-1])
print(d['ok'], [(r['rs_grad_rel_err_owned'], r['rs_nothing_left_behind'], r['rs_p16_rel_err'], r['rs_grad_rel_err_step2']) for r in d['ranks']])" || tail -5 $O/ddp_check_small.err
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29692 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err
python - <<'PY'
import json
b=[json.loads(l) for l in open('gpurun_out/r2c9/breakdown_n2_rs.jsonl').read().strip().splitlines() if l.startswith('{')]
for x in b: print("phase", x["ms"], x["total_ms"]); print({k:v for k,v in x["kernels_ms"].items() if k.startswith("gemm")}); print(x["wgrad_detail_ms"])
PY
timeout 600 $TR --master-port 29693 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.jsonl 2> $O/bench_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench_n2.jsonl').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['allreduce'][:30]); print(d['e2e'])" || tail -20 $O/bench_n2.err
timeout 420 python tools/fault_check.py bert 2 0 --fault-tolerant --victim 1 > $O/ft_bert_n2.log 2>&1; echo "ft rc=$?"; grep '^{' $O/ft_bert_n2.log | tail -1 | cut -c1-900; tail -5 $O/ft_bert_n2.log | cut -c1-300
