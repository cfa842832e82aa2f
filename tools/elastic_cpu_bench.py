"""Submit -> first step and elastic rescale latency (gloo / CPU workers) with and without the agent's warm
interpreter pool.  python tools/elastic_cpu_bench.py"""
import json, os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.cmd.local import LocalCluster
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption
def wait_until(fn, timeout=60, period=0.02):
    d = time.time() + timeout
    while time.time() < d:
        try:
            v = fn()
            if v: return v
        except Exception: pass
        time.sleep(period)
    raise TimeoutError
for pool in (0, 3):
    opt = TrainingJobOperatorOption(thread_num=2, scale_down_grace=20.0)
    with LocalCluster(num_gpus=0, workdir=tempfile.mkdtemp(), option=opt, warm_pool=pool) as lc:
        if pool: wait_until(lambda: lc.agent.warm_ready() == pool, 60, 0.1)
        worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", "mlp", "--batch", "16",
                  "--steps", "0", "--cpu", "--elastic", "--step-sleep", "0.02"]
        job = {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": "el"},
               "spec": {"frameworkType": "pytorch", "replicaSpecs": {"trainer": {
                   "replicas": 2, "minReplicas": 2, "maxReplicas": 4, "edlPolicy": "Manual",
                   "template": {"spec": {"containers": [{"name": "aitj-trainer", "command": worker, "workingDir": ROOT,
                                                         "env": [{"name": "PYTHONPATH", "value": ROOT}]}]}}}}}}
        t0 = time.time(); lc.apply(job)
        wait_until(lambda: "aitj.b200/worker-trace" in lc.jobs().get("el").annotations, 90)
        t_first = time.time() - t0
        if pool: wait_until(lambda: lc.agent.warm_ready() >= 1, 60, 0.1)
        t1 = time.time()
        lc.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 3}}}})
        rec = wait_until(lambda: (lambda a: json.loads(a["aitj.b200/rescale-trace"]) if "aitj.b200/rescale-trace" in a and json.loads(a["aitj.b200/rescale-trace"])["world"] == 3 else None)(lc.jobs().get("el").annotations), 90)
        t_up = time.time() - t1
        t2 = time.time()
        lc.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 2}}}})
        rec2 = wait_until(lambda: (lambda a: json.loads(a["aitj.b200/rescale-trace"]) if json.loads(a.get("aitj.b200/rescale-trace", "{}")).get("generation") == 3 else None)(lc.jobs().get("el").annotations), 90)
        t_down = time.time() - t2
        print(f"warm_pool={pool}: submit->first step {t_first:.2f}s | 2->3: wall {t_up:.2f}s rec {rec} | 3->2: wall {t_down:.2f}s rec {rec2}", flush=True)
