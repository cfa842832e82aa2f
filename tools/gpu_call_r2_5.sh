#!/bin/bash
# round 2, call 5 (2 GPUs): last-arriver move for split-K wgrads; attention backward numerics; quad GEMM probe
set -u
O=gpurun_out/r2c5; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m trainingjob_operator_b200.ops.selfcheck --case attention_bwd > $O/selfcheck_attn_bwd.txt 2>&1; echo "attn_bwd rc=$?"; tail -22 $O/selfcheck_attn_bwd.txt | cut -c1-250
timeout 240 $TR --master-port 29631 tools/ddp_check.py --small > $O/ddp_check_small.json 2> $O/ddp_check_small.err; echo "ddp_check rc=$?"
python -c "
import json
d=json.loads(open('$O/ddp_check_small.json').read().strip().splitlines()[-1])
print(d['ok'], [(r['rs_grad_rel_err_owned'], r['rs_nothing_left_behind'], r['rs_p16_rel_err'], r['rs_grad_rel_err_step2']) for r in d['ranks']])" || tail -5 $O/ddp_check_small.err
for mv in 1 0; do
  AITJ_RS_MOVE=$mv timeout 300 $TR --master-port 2964$mv bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e > $O/bench_n2_rs_move$mv.jsonl 2> $O/bench_n2_rs_move$mv.err
  echo "bench rs move=$mv rc=$?"; grep -o '"ms_per_step": [0-9.]*' $O/bench_n2_rs_move$mv.jsonl | head -1; tail -2 $O/bench_n2_rs_move$mv.err | cut -c1-300
done
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29651 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err; cat $O/breakdown_n2_rs.jsonl
for c in 0 16 24 32; do
  AITJ_GEMM_QUAD_CLUSTERS=$c timeout 120 python tools/quad_probe.py 2>&1 | tail -9
done > $O/quad_probe.txt; cat $O/quad_probe.txt
timeout 200 python -m trainingjob_operator_b200.ops.selfcheck --case graph_step 2>&1 | tail -7 | cut -c1-200
