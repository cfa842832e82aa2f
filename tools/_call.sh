timeout 150 python -m trainingjob_operator_b200.ops.selfcheck --case gpt2_engine 2>&1 | grep -E "FAIL|PASS|EXC|Error|loss" | tail -12
for arm in "cudnn AITJ_ATTN=cudnn" "ours AITJ_ATTN=tcgen05"; do
  set -- $arm
  env $2 timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_attn_$1.log 2>&1; echo "$1 rc=$?"
  grep "^{\"metric" gpurun_out/bench_attn_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"config\"][\"loss_last\"])"
done
