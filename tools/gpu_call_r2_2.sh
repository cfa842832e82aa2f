#!/bin/bash
# round 2, call 2 (2 GPUs): owner-sharded DDP numerics + A/B against NCCL
set -u
O=gpurun_out/r2c2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29601 tools/ddp_check.py > $O/ddp_check_tiny.json 2> $O/ddp_check_tiny.err
echo "ddp_check tiny rc=$?"; tail -c 1500 $O/ddp_check_tiny.json; tail -5 $O/ddp_check_tiny.err
timeout 240 $TR --master-port 29602 tools/ddp_check.py --small > $O/ddp_check_small.json 2> $O/ddp_check_small.err
echo "ddp_check small rc=$?"; tail -c 1500 $O/ddp_check_small.json; tail -5 $O/ddp_check_small.err
for mode in rs nccl; do
  AITJ_ALLREDUCE=$mode timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > $O/bench_n2_$mode.jsonl 2> $O/bench_n2_$mode.err
  echo "bench $mode rc=$?"; cut -c1-260 $O/bench_n2_$mode.jsonl; grep -o '"allreduce": "[^"]*"' $O/bench_n2_$mode.jsonl; tail -3 $O/bench_n2_$mode.err
done
