#!/bin/bash
# round 2, call 4 (2 GPUs): phase breakdown rs vs 1 GPU; quad GEMM numerics + kernel bench on GPU 0
set -u
O=gpurun_out/r2c4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 python -m trainingjob_operator_b200.ops.selfcheck --case gemm_quad > $O/selfcheck_quad.txt 2>&1; echo "quad rc=$?"; tail -25 $O/selfcheck_quad.txt | cut -c1-220
timeout 200 python -m trainingjob_operator_b200.ops.selfcheck --case graph_step > $O/selfcheck_graph.txt 2>&1; echo "graph rc=$?"; tail -8 $O/selfcheck_graph.txt | cut -c1-260
timeout 120 python tools/step_breakdown.py > $O/breakdown_n1.jsonl 2> $O/breakdown_n1.err; cat $O/breakdown_n1.jsonl; tail -2 $O/breakdown_n1.err
AITJ_ALLREDUCE=rs timeout 200 $TR --master-port 29621 tools/step_breakdown.py > $O/breakdown_n2_rs.jsonl 2> $O/breakdown_n2_rs.err; cat $O/breakdown_n2_rs.jsonl; tail -2 $O/breakdown_n2_rs.err | cut -c1-300
timeout 240 $TR --master-port 29622 tools/ddp_check.py --small > $O/ddp_check_small.json 2> $O/ddp_check_small.err; echo "ddp_check rc=$?"
timeout 400 python tools/kernel_bench.py > $O/kernel_bench.txt 2>&1; grep -v "cublas\|layernorm\|softmax\|adamw\|sumsq" $O/kernel_bench.txt | head -70
