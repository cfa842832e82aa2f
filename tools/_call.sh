for c in gemm_2cta gemm_epilogue gpt2_engine; do timeout 150 python -m trainingjob_operator_b200.ops.selfcheck --case $c 2>&1 | grep -E "FAIL|PASS|EXC|Error" | head -5; done
timeout 150 python tools/gemm_trace.py > gpurun_out/gemm_trace_w16.txt 2>&1; echo "trace16 rc=$?"
AITJ_GEMM_EPI_WARPS=8 timeout 150 python tools/gemm_trace.py > gpurun_out/gemm_trace_w8.txt 2>&1; echo "trace8 rc=$?"
for arm in "w16 AITJ_GEMM_EPI_WARPS=16" "w8 AITJ_GEMM_EPI_WARPS=8" "nogather AITJ_QKV_GATHER=0" "w16b AITJ_GEMM_EPI_WARPS=16"; do
  set -- $arm
  env $2 timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ab_$1.log 2>&1; echo "$1 rc=$?"
  grep "^{\"metric" gpurun_out/bench_ab_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
done
