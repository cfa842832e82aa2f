#!/bin/bash
# What round 2 left unmeasured on hardware, cheapest and most decisive first.  Every step has its own
# short timeout and every check tool leaves a post-mortem (job status + worker log tails) in gpurun_out/ when a wait runs
# out, so a step that fails costs a bounded number of GPU-seconds and still explains itself.
#
#   gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_followups.sh'          (4 GPUs x 15 min = 60 GPU-minutes at most)
#
# Lesson of round 2 (profiles/r2_fault_recovery_gpu.md): five silent 240 s waits at 4 GPUs ate the budget.  Here the
# first-step wait is 60 s and the recovery steps stop at the first failure.
set -u
O=gpurun_out/followups; mkdir -p $O
export AITJ_CHECK_FIRST_STEP_TIMEOUT=60

echo "== 1. in-place recovery (faultTolerant), BERT-base, 4 ranks, rank 3 -- ONE attempt first"
timeout 150 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 3 2>&1 | grep '^{' | tail -1 | tee $O/ft_rank3.jsonl
if grep -q '"survivors_kept_their_process": true' $O/ft_rank3.jsonl; then
  for i in 2 3 4 5; do
    timeout 150 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 3 2>&1 | grep '^{' | tail -1 | tee -a $O/ft_rank3.jsonl
  done
  timeout 150 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 0 2>&1 | grep '^{' | tail -1 | tee $O/ft_rank0.jsonl
else
  echo "   no recovery: see gpurun_out/fault_check_*_FAILED_*.json; trying once with NVLS left on and once with a 1 s breaker"
  timeout 150 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 3 NCCL_NVLS_ENABLE=1 2>&1 | grep '^{' | tail -1 | tee $O/ft_rank3_nvls1.jsonl
  timeout 150 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 3 AITJ_FT_ABORT_AFTER=1 2>&1 | grep '^{' | tail -1 | tee $O/ft_rank3_abort1.jsonl
fi

echo "== 2. restartScope Pod with the stall exit (round 1: 32.7 s through the 20 s heartbeat time-out), and scope All, warm pool"
for i in 1 2 3; do
  timeout 150 python tools/fault_check.py bert 4 4 --scope Pod --victim 3 2>&1 | grep '^{' | tail -1 | tee -a $O/scope_pod.jsonl
done
for i in 1 2 3; do
  timeout 150 python tools/fault_check.py bert 4 4 --scope All --victim 3 2>&1 | grep '^{' | tail -1 | tee -a $O/scope_all.jsonl
done

echo "== 3. weight-gradient GEMMs on a side stream (written and hazard-checked on CPU, never measured): same-box A/B"
for v in 0 1 0 1; do
  AITJ_WGRAD_STREAM=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
    --master-port 2951$v bench.py --gpus 4 --steps 30 --warmup 5 --no-e2e 2>/dev/null | grep '^{' | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'wgrad_stream': $v, 'ms_per_step': d['ms_per_step'], 'value': d['value']}))" \
    | tee -a $O/wgrad_stream_ab_n4.jsonl
done

echo "== 4. submit -> first step after the parameter init moved to the device (bench e2e arm), 1 and 4 GPUs"
timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | grep '^{' | tail -1 | tee $O/bench_n1.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 4 --steps 10 --warmup 3 2>/dev/null | grep '^{' | tail -1 | tee $O/bench_n4.jsonl

echo "== 5. leader fail-over under a GPT-2 job at 4 GPUs (BASELINE config 5)"
timeout 200 python tools/failover_check.py gpt2 4 1500 2>&1 | grep '^{' | tail -1 | tee $O/failover_n4.json

python - <<'PY'
import glob, json, statistics as st
for f in sorted(glob.glob("gpurun_out/followups/*.json*")):
    rows = [json.loads(l) for l in open(f) if l.startswith("{")]
    k = [r.get("kill_to_first_step_s") for r in rows if r.get("kill_to_first_step_s")]
    print(f, "n", len(rows), ("kill->first step p50 %.2f s" % st.median(k)) if k else "")
PY
