#!/bin/bash
# round 2, call 6 (2 GPUs): per-kernel profile N=1 vs N=2 (rs); quad GEMM with 2SM multicast; attention bwd
set -u
O=gpurun_out/r2c6; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 python tools/step_breakdown.py > $O/breakdown_n1.jsonl 2> $O/breakdown_n1.err; tail -2 $O/breakdown_n1.err
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29661 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err; tail -2 $O/breakdown_n2_rs.err | cut -c1-200
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2c6/breakdown_n1.jsonl').read().strip().splitlines()[-1])
b=[json.loads(l) for l in open('gpurun_out/r2c6/breakdown_n2_rs.jsonl').read().strip().splitlines() if l.startswith('{')]
print("phase", a["ms"]); [print("phase", x["ms"]) for x in b]
keys=sorted(set(a["kernels_ms"])|set(b[0]["kernels_ms"]))
for k in keys:
    print(f"{k:28s} n1={a['kernels_ms'].get(k)}  n2r0={b[0]['kernels_ms'].get(k)} n2r1={b[1]['kernels_ms'].get(k)}")
for k in sorted(set(a["wgrad_detail_ms"])|set(b[0]["wgrad_detail_ms"])):
    print(f"{k:50s} n1={a['wgrad_detail_ms'].get(k)}  n2r0={b[0]['wgrad_detail_ms'].get(k)}")
PY
timeout 300 python -m trainingjob_operator_b200.ops.selfcheck --case attention_bwd > $O/selfcheck_attn_bwd.txt 2>&1; echo "attn_bwd rc=$?"; tail -3 $O/selfcheck_attn_bwd.txt | cut -c1-250
timeout 200 python -m trainingjob_operator_b200.ops.selfcheck --case gemm_quad > $O/selfcheck_quad.txt 2>&1; echo "quad(2sm mcast) rc=$?"; tail -4 $O/selfcheck_quad.txt | cut -c1-200
timeout 120 python tools/quad_probe.py > $O/quad_probe.txt 2>&1; tail -9 $O/quad_probe.txt
