#!/bin/bash
# round 2, the 4-GPU call: BASELINE config 4 (SIGKILL rank 3 of a BERT-base DDP job) -- in-place recovery (faultTolerant,
# NCCL abort) x5, full restart (scope All) x2, restartScope Pod with the stall exit x2; config 5 (leader fail-over) at 4 GPUs
set -u
O=gpurun_out/r2n4; mkdir -p $O
for i in 1 2 3 4 5; do
  timeout 300 python tools/fault_check.py bert 4 0 --fault-tolerant --victim 3 2>&1 | grep '^{' | tail -1 >> $O/ft_inplace_rank3.jsonl; echo "ft $i rc=$?"
done
for i in 1 2; do
  timeout 300 python tools/fault_check.py bert 4 4 --scope All --victim 3 2>&1 | grep '^{' | tail -1 >> $O/restart_scope_all_rank3.jsonl; echo "all $i rc=$?"
done
for i in 1 2; do
  timeout 300 python tools/fault_check.py bert 4 4 --scope Pod --victim 3 2>&1 | grep '^{' | tail -1 >> $O/restart_scope_pod_rank3.jsonl; echo "pod $i rc=$?"
done
timeout 300 python tools/failover_check.py gpt2 4 1500 2>&1 | grep '^{' | tail -1 > $O/leader_failover_gpt2_n4.json; echo "failover rc=$?"
python - <<'PY'
import json, statistics as st
def load(p):
    try: return [json.loads(l) for l in open(p) if l.startswith('{')]
    except Exception: return []
for name in ("ft_inplace_rank3", "restart_scope_all_rank3", "restart_scope_pod_rank3"):
    xs = load(f"gpurun_out/r2n4/{name}.jsonl")
    ks = [x.get("kill_to_first_step_s") for x in xs if x.get("kill_to_first_step_s")]
    print(name, "n", len(xs), "kill_to_first_step_s", ks, "p50", st.median(ks) if ks else None,
          [x.get("survivors_kept_their_process") for x in xs], [x.get("restart_counts") for x in xs])
print(open("gpurun_out/r2n4/leader_failover_gpt2_n4.json").read()[:600])
PY
