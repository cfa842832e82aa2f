for c in gemm_2cta gemm_epilogue gpt2_engine; do timeout 150 python -m trainingjob_operator_b200.ops.selfcheck --case $c 2>&1 | grep -E "FAIL|PASS|EXC|Error|rel_err" | tail -4; done
timeout 200 python tools/gemm_trace.py > gpurun_out/gemm_trace_v5_group.txt 2>&1; echo "trace rc=$?"
cut -c1-60 gpurun_out/gemm_trace_v5_group.txt | head -18
for arm in "group AITJ_GEMM_GROUP_STORE=1" "nogroup AITJ_GEMM_GROUP_STORE=0" "group2 AITJ_GEMM_GROUP_STORE=1" "nogroup2 AITJ_GEMM_GROUP_STORE=0"; do
  set -- $arm
  env $2 timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ab_$1.log 2>&1; echo "$1 rc=$?"
  grep "^{\"metric" gpurun_out/bench_ab_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
done
