"""The I/O side of a reconcile pass (controller/executor.py) and the cluster view of the auto-scaler."""
import json

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api import register as R
from trainingjob_operator_b200.controller import elastic as E

from test_controller_unit import Harness, job_dict


def test_conflicting_status_write_never_rolls_back_worker_annotations():
    """Workers PATCH their own annotations onto the job (ready-r<rank>, metrics, rescale-trace); such a write is what
    causes the controller's conflict in the first place.  The retry must carry over the status and only the annotations
    the pass set itself -- never its stale view of the workers' keys."""
    h = Harness()
    d = job_dict()
    d["metadata"]["annotations"] = {"aitj.b200/ready-r2": "1", "aitj.b200/metrics": json.dumps({"samples_per_sec": 1.0})}
    h.add_job(d)
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    live["metadata"]["annotations"]["aitj.b200/ready-r2"] = "2"            # the rank re-joined under generation 2
    live["metadata"]["annotations"]["aitj.b200/metrics"] = json.dumps({"samples_per_sec": 6414.0})
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)          # ... which the informer cache has not seen
    h.tc.sync_handler("default/job")                                      # stale copy -> Conflict -> retry
    after = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    ann = after["metadata"]["annotations"]
    assert ann["aitj.b200/ready-r2"] == "2"
    assert json.loads(ann["aitj.b200/metrics"])["samples_per_sec"] == 6414.0
    assert C.ANN_TRACE in ann and "aitj.b200/host-ports" in ann           # the controller's own keys did land
    assert after["status"]["phase"] == "Pending"


def test_two_jobs_declaring_the_same_port_get_different_loopback_ports():
    h = Harness()
    h.add_job(job_dict(name="a"))
    h.add_job(job_dict(name="b"))
    h.sync("a"); h.sync("b")
    host_ports = [p["hostPort"] for s in h.sc.services for p in s["spec"]["ports"]]
    assert len(host_ports) == 4 and len(set(host_ports)) == 4             # (job, role, index, port) -> its own port
    assert all(p["port"] == 2222 for s in h.sc.services for p in s["spec"]["ports"])
    env = {e["name"]: e["value"] for e in h.pc.templates[0]["spec"]["containers"][0]["env"]}
    assert env["AITJ_HOST_PORTS"] in {str(p) for p in host_ports}


def _node(name, idx=None, ready=True, gpu=True):
    labels = {E.LABEL_NODE_TYPE: "gpu" if gpu else "cpu"}
    if idx is not None:
        labels[E.LABEL_GPU_INDEX] = str(idx)
    return {"kind": "Node", "metadata": {"name": name, "labels": labels},
            "status": {"conditions": [{"type": "Ready", "status": "True" if ready else "False"}]}}


def test_cluster_view_finds_gpu_nodes_by_label_not_by_name():
    """With ``--node-prefix`` the agent's nodes are called e.g. ``boxA-gpu-3``: slots are recognised by their labels."""
    from trainingjob_operator_b200.api.types import AITrainingJob

    job = AITrainingJob.from_dict(job_dict())
    nodes = [_node(f"boxA-gpu-{i}", i) for i in range(4)] + [_node("boxA-cpu-0", gpu=False),
                                                             _node("boxA-gpu-9", 9, ready=False)]
    busy_by_binding = {"metadata": {"name": "x", "namespace": "default"}, "spec": {"nodeName": "boxA-gpu-0"},
                       "status": {"phase": "Running"}}
    busy_by_annotation = {"metadata": {"name": "y", "namespace": "default", "annotations": {C.ANN_GPUS: "2,3"}},
                          "spec": {"nodeName": "boxA-gpu-2"}, "status": {"phase": "Running"}}
    assert E.observe_cluster(job, nodes, []).free_gpu_slots == 4
    assert E.observe_cluster(job, nodes, [busy_by_binding]).free_gpu_slots == 3
    assert E.observe_cluster(job, nodes, [busy_by_binding, busy_by_annotation]).free_gpu_slots == 1
