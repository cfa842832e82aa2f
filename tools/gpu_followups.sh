#!/bin/bash
# What was written after round 1's GPU minutes were spent and has only run on CPU/gloo so far.
# Run on a box with >= 4 GPUs:   gpurun --gpus 4 --timeout 1500 -- 'bash tools/gpu_followups.sh'
# Everything is wrapped in its own timeout; results land in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== async checkpoint, CUDA path"
AITJ_GPU_FT_TEST=1 timeout 700 python -m pytest tests/test_gpu_runtime.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/followup_ckpt.txt
echo "== in-place recovery (faultTolerant), BERT-shaped, 4 ranks, rank 0 and rank 3 killed"
for v in 0 3; do
  timeout 420 python tools/fault_check.py bert 4 0 --fault-tolerant --victim $v 2>&1 | grep '^{' | tee -a gpurun_out/followup_fault_tolerant.jsonl
done
echo "== the same job with restartScope All (baseline for the comparison)"
timeout 420 python tools/fault_check.py bert 4 0 --scope All 2>&1 | grep '^{' | tee -a gpurun_out/followup_fault_tolerant.jsonl
echo "== GPU loss under a faultTolerant + Auto job, ResNet-50 workers"
timeout 420 python tools/gpu_loss_check.py 40 --gpu resnet50 2>&1 | grep -v '^[IW]09' | cut -c1-300 | tee gpurun_out/followup_gpu_loss.txt
echo "== bench e2e arm at 2 GPUs (interruptible rendezvous under NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/followup_bench_n2.jsonl
