timeout 240 python tools/fault_check.py bert 2 0 > gpurun_out/fault_n2_cold.log 2>&1; echo "fault cold rc=$?"; grep "^{" gpurun_out/fault_n2_cold.log | cut -c1-600
timeout 240 python tools/fault_check.py bert 2 2 > gpurun_out/fault_n2_warm.log 2>&1; echo "fault warm rc=$?"; grep "^{" gpurun_out/fault_n2_warm.log | cut -c1-600
timeout 300 python tools/failover_check.py gpt2 2 1500 --lease 15 > gpurun_out/failover_n2.log 2>&1; echo "failover rc=$?"; grep "^{" gpurun_out/failover_n2.log | cut -c1-700
i=0
for m in bert resnet50 mnist; do i=$((i+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29590+i)) bench.py --model $m --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n2_$m.log 2>&1; echo "$m rc=$?"
  grep "^{\"metric" gpurun_out/bench_n2_$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"config\"][\"global_batch\"])"
done
