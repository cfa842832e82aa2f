"""Restart-on-failure through the control plane (BASELINE config 4 shape): DDP job, SIGKILL one rank (exit 137),
``restartPolicy: OnFailure`` + ``restartScope: All`` -> every replica is re-created, resumes from rank 0's checkpoint.
Reports kill -> job Running again -> first training step after the restart.

    python tools/fault_check.py [model] [n] [warm_pool] [--cpu] [--scope Pod|All] [--hang SECONDS] [--fault-tolerant]
          [--victim RANK] [--second-victim RANK --second-after SECONDS] [--steps K]

``--scope Pod --hang S``: only the killed replica is re-created by the controller; the survivors, stuck in a collective
with a dead peer, are caught by the agent's heartbeat-based hang detection after S seconds and restarted too.

``--fault-tolerant``: the job is elastic with ``faultTolerant: true``: only the killed replica is replaced, the survivors
catch the failed collective, keep their processes (and device state) and re-rendezvous with the replacement; reports
kill -> first step of the recovered world and checks that the survivors' PIDs did not change.
"""
import json
import os
import signal
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainingjob_operator_b200.cmd.local import LocalCluster  # noqa: E402
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption  # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--") and "=" not in a]
cpu = "--cpu" in sys.argv or "--fake-gpus" in sys.argv
fake_gpus = "--fake-gpus" in sys.argv     # CPU workers that still ask for (and get bound to) one GPU each: the scheduling path
model = argv[0] if argv else ("mlp" if cpu else "bert")
n = int(argv[1]) if len(argv) > 1 else 2
pool = int(argv[2]) if len(argv) > 2 else 0
scope = sys.argv[sys.argv.index("--scope") + 1] if "--scope" in sys.argv else "All"
hang = sys.argv[sys.argv.index("--hang") + 1] if "--hang" in sys.argv else ""
ft = "--fault-tolerant" in sys.argv
victim_arg = sys.argv[sys.argv.index("--victim") + 1] if "--victim" in sys.argv else ""
victim = int(victim_arg) if victim_arg else min(3, n - 1)
second = int(sys.argv[sys.argv.index("--second-victim") + 1]) if "--second-victim" in sys.argv else None
second_after = float(sys.argv[sys.argv.index("--second-after") + 1]) if "--second-after" in sys.argv else 0.0
# --steps K: a finite job (K timed steps after 5 warm-up steps); the check then also waits for it to end Succeed after the
# restart -- a replica resumed from a checkpoint past the warm-up must still produce its throughput record and exit 0
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 0
batch = {"bert": 8, "mlp": 16, "gpt2": 4, "resnet50": 32}.get(model, 8)
worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", model, "--batch", str(batch),
          "--steps", str(steps), "--warmup", "5", "--ckpt-every", "10"] + (["--cpu", "--step-sleep", "0.02"] if cpu else []) + \
         (["--elastic"] if ft else [])
c = {"name": "aitj-trainer", "command": worker, "workingDir": ROOT, "env": [{"name": "PYTHONPATH", "value": ROOT}]}
if hang:
    c["env"].append({"name": "AITJ_HANG_TIMEOUT", "value": hang})
for kv in sys.argv[1:]:
    if "=" in kv and not kv.startswith("--"):
        c["env"].append({"name": kv.split("=", 1)[0], "value": kv.split("=", 1)[1]})
if not cpu or fake_gpus:
    c["resources"] = {"limits": {"nvidia.com/gpu": 1}}
job = {"apiVersion": "elasticdeeplearning.ai/v1", "kind": "AITrainingJob", "metadata": {"name": "ft"},
       "spec": {"frameworkType": "pytorch", "replicaSpecs": {"trainer": {
           "replicas": n, "restartPolicy": "OnFailure", "restartScope": scope, "restartLimit": 6,
           "template": {"spec": {"terminationGracePeriodSeconds": 1, "containers": [c]}}}}}}
if ft:
    job["spec"]["faultTolerant"] = True
    job["spec"]["replicaSpecs"]["trainer"].update({"minReplicas": n, "maxReplicas": n, "edlPolicy": "Manual"})


def wait(fn, timeout=240):
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


def first_step_at(lc):
    tr = json.loads(lc.jobs().get("ft").annotations.get("aitj.b200/worker-trace", "{}"))
    return tr.get("first_step_done", 0.0)


out = {"model": model, "replicas": n, "warm_pool": pool, "cpu": cpu, "victim_rank": victim, "restart_scope": scope,
       "hang_timeout_s": hang or None, "fault_tolerant": ft}
STAGE = {"at": "cluster start"}


def post_mortem(lc, err):
    from _postmortem import dump

    dump(lc, "ft", out, STAGE["at"], err, f"fault_check_{model}_n{n}_pool{pool}_{'ft' if ft else scope.lower()}")


with LocalCluster(num_gpus=n if (fake_gpus or not cpu) else 0, option=TrainingJobOperatorOption(thread_num=2),
                  health_prober=(lambda i: (True, "")) if fake_gpus else None,
                  workdir=tempfile.mkdtemp(prefix=f"aitj-fault-{pool}-"), warm_pool=pool) as lc:
    try:
      # never a stale checkpoint
        if pool:
            wait(lambda: lc.agent.warm_ready() >= pool, 120)
        t_submit = time.time()
        lc.apply(job)
        STAGE["at"] = "submit -> first training step"
        wait(lambda: first_step_at(lc) > 0, float(os.environ.get("AITJ_CHECK_FIRST_STEP_TIMEOUT", "240")))
        out["submit_to_first_step_s"] = round(time.time() - t_submit, 3)
        time.sleep(3.0 if not cpu else 1.0)               # let it train and write a checkpoint
        if pool:
            wait(lambda: lc.agent.warm_ready() >= min(pool, n), 120)
        pids = {sid.split("/")[1]: p for sid, p in lc.agent.sup.list() if "/ft-trainer-" in sid}
        pid = pids[f"ft-trainer-{victim}"]
        t_kill = time.time()
        os.kill(pid, signal.SIGKILL)
        STAGE["at"] = f"SIGKILL of rank {victim} -> recovery record / Running again"
        if ft:
            want_gen = 2
            if second is not None:             # a second rank dies while the first loss is still being repaired
                time.sleep(second_after)
                os.kill(pids[f"ft-trainer-{second}"], signal.SIGKILL)
                want_gen = 3
                out["second_victim_rank"], out["second_after_s"] = second, second_after
            rec = wait(lambda: (lambda r: r if r and r.get("generation", 0) >= want_gen and r.get("recovered_from") else None)(
                json.loads(lc.jobs().get("ft").annotations.get("aitj.b200/rescale-trace", "null"))))
            out["kill_to_first_step_s"] = round(rec["at"] - t_kill, 3)
            out["recovery"] = {k: rec.get(k) for k in ("generation", "world", "seconds", "teardown_s", "init_pg_s",
                                                        "sync_state_s", "first_step_s", "recovered_from")}
            wait(lambda: (lambda j: j.status.phase == "Running" and
                          j.status.replica_statuses["trainer"].active == n)(lc.jobs().get("ft")))
            out["kill_to_running_s"] = round(time.time() - t_kill, 3)
            now = {sid.split("/")[1]: p for sid, p in lc.agent.sup.list() if "/ft-trainer-" in sid}
            out["survivors_kept_their_process"] = all(now.get(k) == v for k, v in pids.items()
                                                      if k not in (f"ft-trainer-{victim}", f"ft-trainer-{second}"))
            j = lc.jobs().get("ft")
            out["restart_counts"] = j.status.restart_counts
            out["conditions"] = [c.type for c in j.status.conditions][-6:]
            logv = open(os.path.join(lc.workdir, "logs", f"default_ft-trainer-{victim}_aitj-trainer.log")).read()
            out["replacement_joined"] = [ln for ln in logv.splitlines() if "joined generation" in ln][-1:]
            lc.jobs().delete("ft")
            time.sleep(0.5)
            os.makedirs("gpurun_out", exist_ok=True)
            json.dump(out, open(f"gpurun_out/fault_check_{model}_n{n}_pool{pool}_ft.json", "w"), indent=1)
            print(json.dumps(out))
            sys.exit(0 if out["restart_counts"].get("trainer", 0) == (1 if second is None else 2) and
                     out["survivors_kept_their_process"] and out["recovery"]["recovered_from"] else 1)
        wait(lambda: (lambda j: j.status.phase == "Running" and j.status.restart_counts.get("trainer", 0) >= 1 and
                      j.status.replica_statuses["trainer"].active == n)(lc.jobs().get("ft")))
        out["kill_to_running_s"] = round(time.time() - t_kill, 3)
        wait(lambda: first_step_at(lc) > t_kill)
        out["kill_to_first_step_s"] = round(first_step_at(lc) - t_kill, 3)
        j = lc.jobs().get("ft")
        out["restart_counts"] = j.status.restart_counts
        out["conditions"] = [c.type for c in j.status.conditions][-6:]
        log0 = open(os.path.join(lc.workdir, "logs", "default_ft-trainer-0_aitj-trainer.log")).read()
        out["resumed"] = [ln for ln in log0.splitlines() if "resumed from checkpoint" in ln][-1:]
        if steps:
            done = wait(lambda: (lambda j: j if j.status.phase in ("Succeed", "Failed", "Timeout", "NodeFail") else None)(
                lc.jobs().get("ft")), 300)
            out["final_phase"] = done.status.phase
            out["final_restart_counts"] = done.status.restart_counts
            out["metrics"] = json.loads(done.annotations.get("aitj.b200/metrics", "null"))
        lc.jobs().delete("ft")
        time.sleep(0.5)
    except (TimeoutError, KeyError, OSError) as err:
        post_mortem(lc, err)
        raise

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/fault_check_{model}_n{n}_pool{pool}_{scope.lower()}.json", "w"), indent=1)
print(json.dumps(out))
sys.exit(0 if out["restart_counts"].get("trainer", 0) >= 1 and out["resumed"] and
         (not steps or out.get("final_phase") == "Succeed") else 1)
