"""Live elastic rescale on real GPUs through the control plane (BASELINE config 3 shape, scaled to the GPUs present):
AITrainingJob with min=1 max=N, edlPolicy Manual; replicas 1 -> N -> max(1, N/2) while training continues.
Reports per-rescale latency (spec change observed by the workers -> first step at the new world size) and
verifies that survivors were never restarted.  Usage: python tools/elastic_gpu_check.py [model] [ngpus] [warm_pool] [ENV=VALUE ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from trainingjob_operator_b200.cmd.local import LocalCluster  # noqa: E402
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
n = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
pool = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # agent warm pool size (0 = cold interpreter starts)
start = 2 if n >= 4 else 1                                  # BASELINE config 3: min=2 max=8, 2 -> 8 -> 4
batch = {"resnet50": 64, "gpt2-tiny": 4, "mnist": 256, "gpt2": 8}.get(model, 8)
worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", model, "--batch", str(batch),
          "--seq", "256", "--steps", "0", "--elastic"]
job = {"apiVersion": "elasticdeeplearning.ai/v1", "kind": "AITrainingJob", "metadata": {"name": "elastic"},
       "spec": {"frameworkType": "pytorch", "replicaSpecs": {"trainer": {
           "replicas": start, "minReplicas": start, "maxReplicas": n, "edlPolicy": "Manual",
           "template": {"spec": {"containers": [{"name": "aitj-trainer", "command": worker, "workingDir": ROOT,
                                                 "env": [{"name": "PYTHONPATH", "value": ROOT}] + [
                                                     {"name": kv.split("=", 1)[0], "value": kv.split("=", 1)[1]}
                                                     for kv in sys.argv[4:] if "=" in kv],
                                                 "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}}}}}


def wait(fn, timeout=240):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            v = fn()
            if v:
                return v
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.05)
    raise TimeoutError


def pids(lc):
    return {sid.split("/")[1]: pid for sid, pid in lc.agent.sup.list()}


out = {"model": model, "gpus": n, "warm_pool": pool, "rescales": []}
opt = TrainingJobOperatorOption(thread_num=2, scale_down_grace=60.0)
with LocalCluster(num_gpus=n, option=opt, workdir=f"/tmp/aitj-elastic-{pool}", warm_pool=pool,
                  gpu_visibility=os.environ.get("AITJ_GPU_VISIBILITY", "all")) as lc:
    try:
        if pool:
            wait(lambda: lc.agent.warm_ready() >= pool, 120)
        t_submit = time.time()
        lc.apply(job)
        wait(lambda: "aitj.b200/worker-trace" in lc.jobs().get("elastic").annotations)
        out["submit_to_first_step_s"] = round(time.time() - t_submit, 3)
        base = pids(lc)
        gen = 1
        for target in [n, max(start, n // 2)]:
            if target == lc.jobs().get("elastic").spec.replica_specs["trainer"].replicas:
                continue
            gen += 1
            if pool:
                wait(lambda: lc.agent.warm_ready() >= min(pool, max(0, target - start)), 120)
            t0 = time.time()
            lc.jobs().patch("elastic", {"spec": {"replicaSpecs": {"trainer": {"replicas": target}}}})
            rec = wait(lambda: (lambda a: json.loads(a["aitj.b200/rescale-trace"])
                                if json.loads(a.get("aitj.b200/rescale-trace", "{}")).get("generation") == gen else None)(
                lc.jobs().get("elastic").annotations))
            rec["target"] = target
            rec["patch_to_first_step_s"] = round(time.time() - t0, 3)
            wait(lambda: len([p for p in lc.pods(selector="TrainingJobName=elastic")
                              if "aitj.b200/scale-down" not in (p["metadata"].get("annotations") or {})]) == target and
                 lc.jobs().get("elastic").status.phase == "Running", 120)
            now = pids(lc)
            rec["survivors_kept_pid"] = all(now.get(k) == v for k, v in base.items() if k in now)
            out["rescales"].append(rec)
            print("rescale", rec, flush=True)
            time.sleep(2.0)
        j = lc.jobs().get("elastic")
        out["restart_counts"] = j.status.restart_counts
        out["phase"] = j.status.phase
        lc.jobs().delete("elastic")
        time.sleep(1.0)
        import shutil
        dst = f"gpurun_out/elastic_logs_{model}_n{n}_pool{pool}"
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(lc.workdir, "logs"), dst, dirs_exist_ok=True)
    except (TimeoutError, KeyError, OSError) as err:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from _postmortem import dump

        dump(lc, "elastic", out, "see the last line printed above", err, f"elastic_gpu_check_{model}_n{n}_pool{pool}")
        raise
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/elastic_gpu_check_{model}_n{n}_pool{pool}.json", "w"), indent=1)
print(json.dumps(out))
ok = out["phase"] == "Running" and all(r["survivors_kept_pid"] for r in out["rescales"]) and \
    not any(out["restart_counts"].values())
sys.exit(0 if ok else 1)
