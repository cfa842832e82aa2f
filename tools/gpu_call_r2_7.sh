#!/bin/bash
# round 2, call 7 (1 GPU): per-GEMM kernel choice measured in the captured step
set -u
O=gpurun_out/r2c7; mkdir -p $O
timeout 900 python tools/gemm_cfg_sweep.py --steps 40 > $O/gemm_cfg_sweep.jsonl 2> $O/gemm_cfg_sweep.err
cat $O/gemm_cfg_sweep.jsonl; tail -3 $O/gemm_cfg_sweep.err
