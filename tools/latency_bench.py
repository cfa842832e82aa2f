"""Control-plane latency benchmark (no GPU needed): submit -> all replicas Running p50/p95 for N-replica jobs,
scale-up / scale-down reaction time, and restart (SIGKILL -> Running again) time, measured through the
public API (LocalCluster.apply) with trivial worker processes.  Writes profiles/control_plane_latency.json."""
import json
import os
import signal
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.api import constants as C  # noqa: E402
from trainingjob_operator_b200.cmd.local import LocalCluster  # noqa: E402
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption  # noqa: E402
from trainingjob_operator_b200.utils import klog  # noqa: E402


def job(name, replicas, **role):
    c = {"name": "aitj-trainer", "command": ["/bin/sleep", "600"], "resources": {"limits": {"nvidia.com/gpu": 1}}}
    r = dict({"replicas": replicas, "template": {"spec": {"containers": [c]}}}, **role)
    return {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": name}, "spec": {"replicaSpecs": {"trainer": r}}}


def wait(fn, timeout=30):
    t0 = time.time()
    while time.time() - t0 < timeout:
        v = fn()
        if v:
            return v
        time.sleep(0.001)
    raise TimeoutError


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def main():
    """``--reference-throttle``: the same controller run under the reference operator's control-plane limits -- client-go's
    default client throttle of 5 qps / burst 10 per clientset (the reference passes its rest.Config through unchanged,
    /root/reference/cmd/app/server.go:126-144), the DefaultControllerRateLimiter work queue (10 qps / burst 100,
    controller.go:113), one synchronous create per pod / service (pod.go:186-193) and a live Node LIST per role per
    pass (pod.go:181,441).  The scheduler and the node agent stay ours (a kube-scheduler + kubelet would only add to
    it), so this is a LOWER bound of what the reference needs on the same box."""
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ref = "--reference-throttle" in sys.argv
    iters = int(args[0]) if args else (5 if ref else 20)
    out = {}
    opt = TrainingJobOperatorOption(thread_num=4, scale_down_grace=0.0)
    if ref:
        opt = TrainingJobOperatorOption(thread_num=4, scale_down_grace=0.0, queue_qps=10.0, queue_burst=100,
                                        kube_api_qps=5.0, kube_api_burst=10, live_node_list=True)
    with LocalCluster(num_gpus=8, option=opt, health_prober=lambda i: (True, "")) as lc:
        klog.set_verbosity(-1)
        import logging
        logging.getLogger("aitj").setLevel(logging.ERROR)
        for n in (1, 2, 4, 8):
            lat = []
            for i in range(iters):
                name = f"lat{n}-{i}"
                t0 = time.perf_counter()
                lc.apply(job(name, n))
                wait(lambda: lc.jobs().get(name).status.phase == "Running")
                lat.append(time.perf_counter() - t0)
                lc.jobs().delete(name)
                wait(lambda: not lc.agent.sup.list() and not lc.pods())
            out[f"submit_to_all_running_s_n{n}"] = {"p50": statistics.median(lat), "p95": pct(lat, 0.95), "n": len(lat)}
            print(n, out[f"submit_to_all_running_s_n{n}"], flush=True)
        # elastic: 2 -> 8 -> 4 (control plane part: spec change -> replicas Running / surplus gone)
        up, down = [], []
        for i in range(max(5, iters // 2)):
            name = f"el-{i}"
            lc.apply(job(name, 2, minReplicas=2, maxReplicas=8, edlPolicy="Manual"))
            wait(lambda: lc.jobs().get(name).status.phase == "Running")
            t0 = time.perf_counter()
            lc.jobs().patch(name, {"spec": {"replicaSpecs": {"trainer": {"replicas": 8}}}})
            wait(lambda: (lambda j: j.status.phase == "Running" and j.status.replica_statuses["trainer"].active == 8)(
                lc.jobs().get(name)))
            up.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            lc.jobs().patch(name, {"spec": {"replicaSpecs": {"trainer": {"replicas": 4}}}})
            wait(lambda: len(lc.pods(selector=f"TrainingJobName={name}")) == 4 and
                 lc.jobs().get(name).status.phase == "Running")
            down.append(time.perf_counter() - t0)
            lc.jobs().delete(name)
            wait(lambda: not lc.agent.sup.list() and not lc.pods())
        out["rescale_2_to_8_control_plane_s"] = {"p50": statistics.median(up), "p95": pct(up, 0.95)}
        out["rescale_8_to_4_control_plane_s"] = {"p50": statistics.median(down), "p95": pct(down, 0.95)}
        # restart: SIGKILL rank 3 of 8 -> Running again (restartPolicy OnFailure, scope Pod)
        rs = []
        for i in range(max(5, iters // 2)):
            name = f"rs-{i}"
            lc.apply(job(name, 8, restartPolicy="OnFailure", restartScope="Pod"))
            wait(lambda: lc.jobs().get(name).status.phase == "Running")
            pid = next(p for sid, p in lc.agent.sup.list() if f"/{name}-trainer-3/" in sid)
            t0 = time.perf_counter()
            os.kill(pid, signal.SIGKILL)
            wait(lambda: (lambda j: j.status.phase == "Running" and j.status.restart_counts.get("trainer") == 1)(
                lc.jobs().get(name)))
            rs.append(time.perf_counter() - t0)
            lc.jobs().delete(name)
            wait(lambda: not lc.agent.sup.list() and not lc.pods())
        out["sigkill_rank3_to_running_again_s"] = {"p50": statistics.median(rs), "p95": pct(rs, 0.95)}
    out["note"] = ("control-plane only: worker = /bin/sleep, Running = process started (kubelet semantics). The "
                   "reference's floor for N=8 is >= 2N+2 API writes behind a 5 qps / burst 10 client throttle "
                   "(BASELINE.md §2) plus scheduler + kubelet container start.")
    os.makedirs("profiles", exist_ok=True)
    if ref:
        out["mode"] = "reference-throttle: kube-api 5 qps / burst 10 per clientset, work queue 10 qps / burst 100, live node LIST per role per pass"
    json.dump(out, open("profiles/control_plane_latency_reference_throttle.json" if ref
                        else "profiles/control_plane_latency.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
