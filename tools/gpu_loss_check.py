"""GPU loss under a ``faultTolerant`` + ``edlPolicy: Auto`` job (CPU/gloo workers on fake GPU slots, or real GPUs with
``--gpu MODEL``): inject a GPU fault under rank 1 of 4.  The controller fails the replica (NodeFail), cannot place its
replacement, shrinks the role to 3; the survivors catch the broken collective, keep their state, the surplus rank leaves,
the replacement joins on the freed slot and training continues at world 3.  Prints the job's state changes with
timestamps relative to the fault, then the workers' own log lines.

    python tools/gpu_loss_check.py [seconds_to_watch] [--gpu MODEL]     # MODEL: resnet50 | mnist | bert | gpt2 (4 GPUs)
"""
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def wait_until(fn, timeout=60.0):
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


from trainingjob_operator_b200.cli import kubectl
from trainingjob_operator_b200.cmd.local import LocalCluster
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption
opt = TrainingJobOperatorOption(thread_num=2, gc_interval=0.5, scale_down_grace=10.0)
wd = tempfile.mkdtemp()
GPU_MODEL = sys.argv[sys.argv.index("--gpu") + 1] if "--gpu" in sys.argv else ""
if GPU_MODEL:
    sys.argv = [a for a in sys.argv if a not in ("--gpu", GPU_MODEL)]
    worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", GPU_MODEL, "--batch",
              {"resnet50": "64", "mnist": "512"}.get(GPU_MODEL, "8"), "--steps", "0", "--elastic"]
else:
    worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", "mlp", "--batch", "16",
              "--steps", "0", "--cpu", "--elastic", "--step-sleep", "0.02"]
job = {"apiVersion": "elasticdeeplearning.ai/v1", "kind": "AITrainingJob", "metadata": {"name": "af"},
       "spec": {"frameworkType": "pytorch", "faultTolerant": True, "replicaSpecs": {"trainer": {
           "replicas": 4, "minReplicas": 2, "maxReplicas": 4, "edlPolicy": "Auto", "restartPolicy": "OnNodeFail", "restartLimit": 3,
           "template": {"spec": {"terminationGracePeriodSeconds": 1, "containers": [
               {"name": "aitj-trainer", "command": worker, "workingDir": ROOT, "resources": {"limits": {"nvidia.com/gpu": 1}},
                "env": [{"name": "PYTHONPATH", "value": ROOT}]}]}}}}}}
with LocalCluster(num_gpus=4, workdir=wd, option=opt, health_prober=lambda i: (True, ""), health_period=0.1) as lc:   # faults are injected, not probed
    lc.apply(job)
    wait_until(lambda: "aitj.b200/worker-trace" in lc.jobs().get("af").annotations, timeout=60)
    time.sleep(3.0 if GPU_MODEL else 1.0)
    buf = io.StringIO()
    t0 = time.time()
    kubectl.main(["inject", "gpu-fault", "gpu-1", "--message", "Xid 79"], clientset=lc.clientset, out=buf)
    last = None
    while time.time() - t0 < (float(sys.argv[1]) if len(sys.argv) > 1 else 10.0):
        j = lc.jobs().get("af")
        pods = sorted((p["metadata"]["name"][-1], (p["spec"].get("nodeName") or "-")[-1], p["status"].get("phase")[:4]) for p in lc.pods(selector="TrainingJobName=af"))
        rs = j.annotations.get("aitj.b200/rescale-trace")
        cur = (j.status.phase, j.spec.replica_specs["trainer"].replicas, dict(j.status.restart_counts), j.status.rendezvous.generation, j.status.rendezvous.world_sizes, pods, rs and json.loads(rs).get("world"))
        if cur != last:
            print(round(time.time()-t0, 2), cur, flush=True); last = cur
        time.sleep(0.05)
    for i in range(4):
        print("--- log", i)
        print("\n".join(l[:200] for l in open(os.path.join(wd, "logs", f"default_af-trainer-{i}_aitj-trainer.log")).read().splitlines() if l.startswith("[worker")))
