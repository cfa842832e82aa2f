#!/bin/bash
# round 2, the 8-GPU call: owner-sharded DDP at 8 ranks (numerics, NVLink bytes, bench + e2e), NCCL A/B, the stock
# comparator arms, the other models at 8 GPUs.  Everything bounded by its own timeout.
set -u
O=gpurun_out/r2n8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29801 tools/ddp_check.py --small > $O/ddp_check_small_n8.json 2> $O/ddp_check_small_n8.err; echo "ddp_check rc=$?"
python -c "
import json
d=json.loads(open('$O/ddp_check_small_n8.json').read().strip().splitlines()[-1])
print(d['ok'], [(round(r['rs_grad_rel_err_owned'],10), r['rs_nothing_left_behind'], r['rs_p16_mismatch_all_ranks_params']) for r in d['ranks']])" || tail -5 $O/ddp_check_small_n8.err
timeout 500 $TR --master-port 29802 bench.py --gpus 8 --steps 30 --warmup 5 > $O/bench_n8_rs.jsonl 2> $O/bench_n8_rs.err; echo "bench rs rc=$?"
python -c "
import json
d=json.loads(open('$O/bench_n8_rs.jsonl').read().strip().splitlines()[-1])
print('rs', d['value'], d['ms_per_step'], d['config']['allreduce'][:20], d['clocks']); print(d['e2e'])" || tail -20 $O/bench_n8_rs.err
for mode in rs nccl; do
  AITJ_ALLREDUCE=$mode timeout 200 $TR --master-port 29803 tools/nvlink_probe.py --steps 20 > $O/nvlink_$mode.json 2> $O/nvlink_$mode.err; echo "nvlink $mode rc=$?"; tail -1 $O/nvlink_$mode.json | cut -c1-1500
done
AITJ_ALLREDUCE=nccl timeout 300 $TR --master-port 29804 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e > $O/bench_n8_nccl.jsonl 2> $O/bench_n8_nccl.err; echo "bench nccl rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/bench_n8_nccl.jsonl | head -1
timeout 300 $TR --master-port 29805 bench.py --impl torch_stock_compiled --gpus 8 --steps 30 --warmup 5 > $O/bench_n8_stock_compiled.jsonl 2> $O/bench_n8_stock_compiled.err; echo "stock compiled rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/bench_n8_stock_compiled.jsonl | head -1
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29807 tools/step_breakdown.py > $O/breakdown_n8_rs.jsonl 2> $O/breakdown_n8_rs.err
python - <<'PY'
import json
try:
    b=[json.loads(l) for l in open('gpurun_out/r2n8/breakdown_n8_rs.jsonl').read().strip().splitlines() if l.startswith('{')]
    for x in b[:2]+b[-1:]: print("phase", x["rank"], x["ms"], x["total_ms"], {k:v for k,v in x["kernels_ms"].items() if k.startswith("gemm")})
except Exception as e: print("breakdown failed", e)
PY
for m in mnist resnet50 bert; do
  timeout 300 $TR --master-port 29808 bench.py --gpus 8 --steps 30 --warmup 5 --model $m --no-e2e > $O/bench_n8_$m.jsonl 2> $O/bench_n8_$m.err; echo "bench $m rc=$?"
  python -c "
import json
d=json.loads(open('$O/bench_n8_$m.jsonl').read().strip().splitlines()[-1])
print('$m', d['value'], d['ms_per_step'], d['config']['cuda_graph'], d['config']['allreduce'][:20])" || tail -8 $O/bench_n8_$m.err
done
