"""Leader-election fail-over under a running DDP job (BASELINE config 5 shape): two operators, kill the leader while
the job trains; the standby takes the lease, the workers are never restarted, the job completes (completePolicy All).
Reports the takeover time and the job's samples/sec (which must not notice).

    python tools/failover_check.py [model] [n] [steps] [--cpu] [--lease S] [--graceful]
(default: the leader is crashed, i.e. does not hand its lease over; --graceful = clean stop with ReleaseOnCancel)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainingjob_operator_b200.cmd.local import LocalCluster  # noqa: E402
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption  # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
cpu = "--cpu" in sys.argv
lease = float(sys.argv[sys.argv.index("--lease") + 1]) if "--lease" in sys.argv else 15.0
model = argv[0] if argv else ("mlp" if cpu else "gpt2")
n = int(argv[1]) if len(argv) > 1 else 2
steps = int(argv[2]) if len(argv) > 2 else (400 if cpu else 1200)
worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", model, "--steps", str(steps),
          "--warmup", "5"] + (["--cpu", "--batch", "16", "--step-sleep", "0.02"] if cpu else [])
c = {"name": "aitj-trainer", "command": worker, "workingDir": ROOT, "env": [{"name": "PYTHONPATH", "value": ROOT}]}
if not cpu:
    c["resources"] = {"limits": {"nvidia.com/gpu": 1}}
job = {"apiVersion": "elasticdeeplearning.ai/v1", "kind": "AITrainingJob", "metadata": {"name": "ha"},
       "spec": {"frameworkType": "pytorch", "completePolicy": "All",
                "replicaSpecs": {"trainer": {"replicas": n, "completePolicy": "All",
                                             "template": {"spec": {"containers": [c]}}}}}}


def wait(fn, timeout=300):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            v = fn()
            if v:
                return v
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.02)
    raise TimeoutError


opt = TrainingJobOperatorOption(thread_num=2)
opt.leader_election.leader_elect = True
# the reference's defaults are 15 s / 5 s / 3 s (cmd/app/options/options.go:39-49); same ratios for other leases
opt.leader_election.lease_duration, opt.leader_election.renew_deadline, opt.leader_election.retry_period = \
    lease, lease / 3.0, lease / 5.0
out = {"model": model, "replicas": n, "steps": steps, "lease_s": lease, "crash": "--graceful" not in sys.argv}
with LocalCluster(num_gpus=0 if cpu else n, operators=2, option=opt, workdir="/tmp/aitj-failover") as lc:
    try:
        lock = lambda: json.loads(lc.clientset.core_v1().endpoints("kube-system").get("trainingjob-operator")  # noqa: E731
                                  ["metadata"]["annotations"]["control-plane.alpha.kubernetes.io/leader"])
        lc.apply(job)
        wait(lambda: "aitj.b200/worker-trace" in lc.jobs().get("ha").annotations)
        pids = {sid: p for sid, p in lc.agent.sup.list()}
        leader = lock()["holderIdentity"]
        t0 = time.time()
        lc.stop_operator(int(leader[-1]), crash="--graceful" not in sys.argv)
        wait(lambda: lock()["holderIdentity"] not in ("", leader), timeout=lease * 3 + 30)
        out["takeover_s"] = round(time.time() - t0, 3)
        out["workers_untouched"] = {sid: p for sid, p in lc.agent.sup.list()} == pids
        final = lc.wait_for_phase("ha", "Succeed", timeout=600)
        out["phase"] = final.status.phase
        out["restart_counts"] = final.status.restart_counts
        out["metrics"] = json.loads(final.annotations.get("aitj.b200/metrics", "{}"))
        out["new_leader"] = lock()["holderIdentity"]
    except (TimeoutError, KeyError, OSError) as err:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from _postmortem import dump

        dump(lc, "ha", out, "see the last line printed above", err, f"failover_check_{model}_n{n}")
        raise
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/failover_check_{model}_n{n}.json", "w"), indent=1)
print(json.dumps(out))
sys.exit(0 if out["phase"] == "Succeed" and out["workers_untouched"] and not any(out["restart_counts"].values()) else 1)
