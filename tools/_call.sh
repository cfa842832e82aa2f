for c in gemm_2cta fused_ops gpt2_engine; do timeout 150 python -m trainingjob_operator_b200.ops.selfcheck --case $c 2>&1 | grep -E "FAIL|PASS|EXC|Error|colsum|xent" | tail -6; done
timeout 200 python tools/kernel_bench.py --mem-only 2>&1 | tail -5
for arm in "new X=1" "old AITJ_XENT_CLUSTER=0,AITJ_FUSE_COLSUM=0" "new2 X=1" "old2 AITJ_XENT_CLUSTER=0,AITJ_FUSE_COLSUM=0" "xentonly AITJ_FUSE_COLSUM=0"; do
  set -- $arm
  env $(echo $2 | tr ',' ' ') timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ab_$1.log 2>&1; echo "$1 rc=$?"
  grep "^{\"metric" gpurun_out/bench_ab_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"config\"][\"loss_last\"])"
done
