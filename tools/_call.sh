timeout 120 python -m trainingjob_operator_b200.ops.selfcheck --case fused_ops 2>&1 | grep -E "layernorm|FAIL|PASS|EXC|Error" | head -12
for t in gemm_qkv_pair gemm_square_pair gemm_fcwgrad_pair; do
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 1 -k regex:gemm_bf16 -o gpurun_out/prof_$t -f python tools/ncu_target.py $t > gpurun_out/ncu_$t.log 2>&1; echo "$t rc=$?"
done
timeout 200 python tools/kernel_bench.py --mem-only > gpurun_out/kb_ln.txt 2>&1; tail -5 gpurun_out/kb_ln.txt
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_gpt2_n1_v7.log 2>&1; echo "n1 rc=$?"; grep "^{\"metric" gpurun_out/bench_gpt2_n1_v7.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"])"
