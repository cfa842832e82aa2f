#!/bin/bash
# round 2, call 1: GPU tests + product bench + comparator arms on one box (1 GPU)
set -u
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_gpu.txt
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > $O/bench_ours.jsonl 2> $O/bench_ours.err
timeout 300 python bench.py --impl torch_stock --gpus 1 --steps 30 --warmup 5 > $O/bench_stock.jsonl 2> $O/bench_stock.err
timeout 600 python bench.py --impl torch_stock_compiled --gpus 1 --steps 30 --warmup 5 > $O/bench_stock_compiled.jsonl 2> $O/bench_stock_compiled.err
timeout 300 python tools/kernel_bench.py > $O/kernel_bench.txt 2>&1
tail -3 $O/pytest_gpu.txt; cat $O/bench_ours.jsonl | cut -c1-400; cat $O/bench_stock.jsonl | cut -c1-300; cat $O/bench_stock_compiled.jsonl | cut -c1-300; tail -5 $O/bench_stock_compiled.err
