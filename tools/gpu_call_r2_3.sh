#!/bin/bash
# round 2, call 3 (2 GPUs): reduce-scatter through TMA bulk reduce-adds on the peer mapping vs from registers
set -u
O=gpurun_out/r2c3; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for tma in 1 0; do
  AITJ_RS_TMA=$tma timeout 240 $TR --master-port 2960$tma tools/ddp_check.py --small > $O/ddp_check_small_tma$tma.json 2> $O/ddp_check_small_tma$tma.err
  echo "ddp_check small tma=$tma rc=$?"; python -c "
import json,sys
d=json.loads(open('$O/ddp_check_small_tma$tma.json').read().strip().splitlines()[-1])
print(d['ok'], [(r['rs_grad_rel_err_owned'], r['rs_p16_rel_err'], r['rs_grad_rel_err_step2']) for r in d['ranks']])" || tail -5 $O/ddp_check_small_tma$tma.err
done
for tma in 1 0 1 0; do
  AITJ_RS_TMA=$tma timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e > $O/bench_n2_rs_tma$tma.jsonl 2> $O/bench_n2_rs_tma$tma.err
  echo "bench rs tma=$tma rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/bench_n2_rs_tma$tma.jsonl | head -1; tail -2 $O/bench_n2_rs_tma$tma.err | cut -c1-300
done
