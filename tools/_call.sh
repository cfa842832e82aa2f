for c in gemm_2cta gemm_epilogue gpt2_engine; do AITJ_GEMM_DIRECT_STORE=1 timeout 150 python -m trainingjob_operator_b200.ops.selfcheck --case $c 2>&1 | grep -E "FAIL|PASS|EXC|Error" | head -5; done
AITJ_GEMM_DIRECT_STORE=1 timeout 200 python tools/gemm_trace.py > gpurun_out/gemm_trace_direct.txt 2>&1; echo "trace rc=$?"
cut -c1-60 gpurun_out/gemm_trace_direct.txt | head -14
for arm in "staged AITJ_GEMM_DIRECT_STORE=0" "direct AITJ_GEMM_DIRECT_STORE=1" "staged2 AITJ_GEMM_DIRECT_STORE=0" "direct2 AITJ_GEMM_DIRECT_STORE=1"; do
  set -- $arm
  env $2 timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ab_$1.log 2>&1; echo "$1 rc=$?"
  grep "^{\"metric" gpurun_out/bench_ab_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
done
