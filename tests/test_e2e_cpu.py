"""Integration tests on CPU: the whole control plane (API server + node agent + operator) driving real
OS processes (SURVEY.md §4 integration tier): apply -> get -> describe -> delete, restart policies under
real faults (kill -9 => 137), GPU-slot ("node") failure, time limit, preemption, clean-pod policy, elastic
rescale with torch.distributed gloo workers, leader fail-over, orphan GC, the kubectl-compatible CLI."""
import io
import json
import os
import signal
import sys
import time

import pytest
import yaml

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api import meta as M
from trainingjob_operator_b200.api.types import AITrainingJob
from trainingjob_operator_b200.cli import kubectl
from trainingjob_operator_b200.cmd.local import LocalCluster
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption
from trainingjob_operator_b200.store.apiserver import APIError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sh_job(name, script, replicas=2, gpus=0, **role):
    c = {"name": "aitj-trainer", "image": "local/sh", "command": ["/bin/sh", "-c", script],
         "ports": [{"name": "aitj-2222", "containerPort": 2222}]}
    if gpus:
        c["resources"] = {"limits": {"nvidia.com/gpu": gpus}}
    r = dict({"replicas": replicas, "template": {"spec": {"containers": [c]}}}, **role)
    return {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": name},
            "spec": {"replicaSpecs": {"trainer": r}}}


def wait_until(fn, timeout=20.0, period=0.02):
    deadline = time.time() + timeout
    while time.time() < deadline:
        try:
            v = fn()
            if v:
                return v
        except APIError:
            pass
        time.sleep(period)
    raise TimeoutError("condition not met")


def pod_pids(lc, job):
    out = {}
    for sid, pid in lc.agent.sup.list():
        if f"/{job}-" in sid:
            out[sid.split("/")[1]] = pid
    return out


@pytest.fixture
def lc(tmp_path):
    opt = TrainingJobOperatorOption(thread_num=2, gc_interval=0.5, scale_down_grace=3.0)
    cluster = LocalCluster(num_gpus=4, workdir=str(tmp_path), option=opt, health_prober=lambda i: (True, ""),
                           health_period=0.1)
    cluster.start()
    yield cluster
    cluster.stop()


def test_success_path_env_contract_and_cleanup(lc):
    script = 'echo "r=$TRAININGJOB_REPLICA_INDEX n=$TRAINER_INSTANCES_NUM h=$TRAINER_HOSTS w=$WORLD_SIZE ' \
             'g=$CUDA_VISIBLE_DEVICES p=$TRAININGJOB_PORTS"; sleep 1.5'
    lc.apply(sh_job("ok", script, replicas=2, gpus=1))
    wait_until(lambda: lc.jobs().get("ok").status.phase == "Running")
    pods = lc.pods(selector="TrainingJobName=ok")
    assert sorted(p["spec"]["nodeName"] for p in pods) == ["gpu-0", "gpu-1"]
    assert sorted(p["metadata"]["annotations"][C.ANN_GPUS] for p in pods) == ["0", "1"]
    svcs = lc.clientset.core_v1().services("default").list()["items"]
    assert sorted(s["metadata"]["name"] for s in svcs) == ["ok-trainer-0", "ok-trainer-1"]
    job = lc.wait_for_phase("ok", "Succeed", timeout=20)
    assert [c.type for c in job.status.conditions][-3:] == ["Running", "Terminating", "Succeed"]
    assert job.status.start_time and job.status.start_running_time and job.status.end_time
    wait_until(lambda: lc.pods() == [])
    assert lc.clientset.core_v1().services("default").list()["items"] == []       # cleanPodPolicy All
    log = open(os.path.join(lc.workdir, "logs", "default_ok-trainer-1_aitj-trainer.log")).read()
    assert "r=1 n=2 h=ok-trainer-0.default:2222,ok-trainer-1.default:2222 w=2 g=1 p=2222" in log
    want = {"SuccessfulCreatePod", "SuccessfulCreateService", "SuccessfulDeletePod", "SuccessfulDeleteService"}
    # (events are written by the recorder's sink thread, a moment after the action they describe)
    wait_until(lambda: want <= {e["reason"] for e in lc.clientset.core_v1().events("default").list()["items"]})
    tr = json.loads(job.annotations[C.ANN_TRACE])
    assert 0 <= tr["running"] - tr["submitted"] < 5.0              # reconcile -> all-running latency is recorded


def test_failure_fail_policy_any_and_exit_code_in_message(lc):
    lc.apply(sh_job("bad", 'if [ "$TRAININGJOB_REPLICA_INDEX" = 1 ]; then exit 3; fi; sleep 30'))
    job = lc.wait_for_phase("bad", "Failed", timeout=20)
    msg = job.status.conditions[-1].message
    assert "pod bad-trainer-1 is failed" in msg and "exitcode 3" in msg and msg.endswith("deleted pods")
    wait_until(lambda: lc.agent.sup.list() == [])                   # the healthy replica was torn down too


def test_sigkill_rank_restarts_with_exit_code_policy(lc):
    """BASELINE config 4 shape: SIGKILL one rank => exit 137 => restart (OnNodeFailWithExitCode / 137,128)."""
    job = sh_job("kill", 'echo "attempt=$TRAININGJOB_REPLICA_RESTARTCOUNT"; sleep 30', replicas=3,
                 restartPolicy="OnNodeFailWithExitCode", restartScope="Pod", restartLimit=2)
    job["spec"]["restartingExitCode"] = "137,128"
    lc.apply(job)
    wait_until(lambda: lc.jobs().get("kill").status.phase == "Running")
    pids = pod_pids(lc, "kill")
    survivors = {k: v for k, v in pids.items() if k != "kill-trainer-1"}
    os.kill(pids["kill-trainer-1"], signal.SIGKILL)
    wait_until(lambda: lc.jobs().get("kill").status.restart_counts.get("trainer") == 1)
    job2 = wait_until(lambda: (lambda j: j if j.status.phase == "Running" and "Restarting" in
                               [c.type for c in j.status.conditions] else None)(lc.jobs().get("kill")))
    types = [c.type for c in job2.status.conditions]
    assert types[types.index("Terminating"):][:3] == ["Terminating", "Restarting", "Running"]
    new = pod_pids(lc, "kill")
    assert new["kill-trainer-1"] != pids["kill-trainer-1"]          # re-created
    assert {k: new[k] for k in survivors} == survivors              # scope Pod: the others were not touched
    pod = lc.clientset.core_v1().pods("default").get("kill-trainer-1")
    assert pod["metadata"]["labels"]["RestartCount"] == "1"
    wait_until(lambda: "attempt=1" in open(os.path.join(lc.workdir, "logs",
                                                        "default_kill-trainer-1_aitj-trainer.log")).read())
    assert job2.status.rendezvous.generation == 2


def test_restart_scope_all_and_limit_exhaustion(lc):
    job = sh_job("lim", 'if [ "$TRAININGJOB_REPLICA_INDEX" = 0 ]; then sleep 0.2; exit 1; fi; sleep 30', replicas=2,
                 restartPolicy="OnFailure", restartScope="All", restartLimit=1)
    lc.apply(job)
    final = lc.wait_for_phase("lim", "Failed", timeout=30)
    assert final.status.restart_counts == {"trainer": 1}
    types = [c.type for c in final.status.conditions]
    assert types.count("Restarting") == 1 and types[-1] == "Failed"


def test_never_policy_does_not_restart(lc):
    lc.apply(sh_job("nv", "exit 137", replicas=1))
    final = lc.wait_for_phase("nv", "Failed", timeout=15)
    assert final.status.restart_counts.get("trainer", 0) == 0


def test_gpu_fault_is_node_fail_and_restarts_elsewhere(lc):
    job = sh_job("nf", "sleep 30", replicas=2, gpus=1, restartPolicy="OnNodeFail", restartScope="Pod")
    lc.apply(job)
    wait_until(lambda: lc.jobs().get("nf").status.phase == "Running")
    victim = next(p for p in lc.pods(selector="TrainingJobName=nf") if p["metadata"]["name"] == "nf-trainer-1")
    node = victim["spec"]["nodeName"]
    buf = io.StringIO()
    assert kubectl.main(["inject", "gpu-fault", node, "--message", "Xid 79"], clientset=lc.clientset, out=buf) == 0
    wait_until(lambda: lc.jobs().get("nf").status.restart_counts.get("trainer") == 1)
    moved = wait_until(lambda: (lambda p: p if p["spec"].get("nodeName") not in ("", None, node) and
                                p["status"].get("phase") == "Running" else None)(
        lc.clientset.core_v1().pods("default").get("nf-trainer-1")))
    assert moved["spec"]["nodeName"] != node
    assert any("is failed and offline" in c.message for c in lc.jobs().get("nf").status.conditions)
    kubectl.main(["inject", "gpu-heal", node], clientset=lc.clientset, out=buf)


def test_auto_elastic_job_shrinks_around_a_lost_gpu_and_grows_back(lc):
    """``edlPolicy: Auto`` (declared, never read by the reference: replica.go:19): a GPU drops out under a 4-replica job.
    The replica that lived there cannot be placed again, so the role shrinks by exactly one (no overshoot while the
    surplus replica drains), the indices are re-packed onto the healthy GPUs and the job runs at 3; when the GPU heals the
    job grows back to 4 without anybody touching it."""
    job = sh_job("auto", "sleep 60", replicas=4, gpus=1, restartPolicy="OnNodeFail", minReplicas=2, maxReplicas=4,
                 edlPolicy="Auto")
    job["spec"]["faultTolerant"] = True
    lc.apply(job)
    wait_until(lambda: lc.jobs().get("auto").status.phase == "Running")
    buf = io.StringIO()
    assert kubectl.main(["inject", "gpu-fault", "gpu-1", "--message", "Xid 79"], clientset=lc.clientset, out=buf) == 0
    sizes = set()

    def settled():
        j = lc.jobs().get("auto")
        sizes.add(j.spec.replica_specs["trainer"].replicas)
        pods = lc.pods(selector="TrainingJobName=auto")
        return j if (j.status.phase == "Running" and j.spec.replica_specs["trainer"].replicas == 3 and
                     len(pods) == 3 and all(p["status"].get("phase") == "Running" for p in pods)) else None

    j = wait_until(settled, timeout=30)
    assert sizes <= {4, 3}                                             # never shrank below what was needed
    assert j.status.rendezvous.world_sizes == {"trainer": 3} and j.status.restart_counts.get("trainer") == 1
    pods = lc.pods(selector="TrainingJobName=auto")
    assert sorted(p["metadata"]["name"] for p in pods) == ["auto-trainer-0", "auto-trainer-1", "auto-trainer-2"]
    assert "gpu-1" not in {p["spec"]["nodeName"] for p in pods}
    kubectl.main(["inject", "gpu-heal", "gpu-1"], clientset=lc.clientset, out=buf)
    j = wait_until(lambda: (lambda x: x if x.status.phase == "Running" and
                            x.status.replica_statuses["trainer"].active == 4 else None)(lc.jobs().get("auto")),
                   timeout=30)
    assert j.spec.replica_specs["trainer"].replicas == 4 and j.status.rendezvous.world_sizes == {"trainer": 4}


def test_auto_elastic_job_yields_gpus_to_more_important_work_and_takes_them_back(lc):
    """An ``edlPolicy: Auto`` role occupies every GPU; a job with a higher ``spec.priority`` arrives and cannot be placed.
    The elastic job gives back exactly what is missing (never below ``minReplicas``), the important job runs, and when it
    is done the elastic job grows into the free slots again."""
    lc.apply(sh_job("bg", "sleep 120", replicas=4, gpus=1, minReplicas=1, maxReplicas=4, edlPolicy="Auto"))
    wait_until(lambda: lc.jobs().get("bg").status.phase == "Running")
    vip = sh_job("vip", "sleep 3", replicas=2, gpus=1)
    vip["spec"]["priority"] = "high"
    lc.apply(vip)
    seen = set()

    def vip_running():
        seen.add(lc.jobs().get("bg").spec.replica_specs["trainer"].replicas)
        return lc.jobs().get("vip").status.phase in ("Running", "Succeed")

    wait_until(vip_running, timeout=30)
    assert min(seen) == 2                                         # gave back two slots, not more
    assert lc.jobs().get("bg").status.restart_counts.get("trainer", 0) == 0
    lc.wait_for_phase("vip", "Succeed", timeout=30)
    j = wait_until(lambda: (lambda x: x if x.spec.replica_specs["trainer"].replicas == 4 and x.status.phase == "Running"
                            and x.status.replica_statuses["trainer"].active == 4 else None)(lc.jobs().get("bg")),
                   timeout=30)
    assert j.status.rendezvous.world_sizes == {"trainer": 4}


def test_unschedulable_message_and_priority(lc):
    lc.apply(sh_job("big", "sleep 30", replicas=6, gpus=1))       # only 4 GPU slots
    job = wait_until(lambda: (lambda j: j if j.status.replica_statuses.get("trainer") and
                              j.status.replica_statuses["trainer"].pending == 2 else None)(lc.jobs().get("big")))
    assert job.status.phase == "Pending"
    pend = wait_until(lambda: (lambda ps: ps if len(ps) == 2 and all(p.get("status", {}).get("conditions") for p in ps)
                               else None)([p for p in lc.pods(selector="TrainingJobName=big")
                                           if not p["spec"].get("nodeName")]))
    assert "Insufficient nvidia.com/gpu" in pend[0]["status"]["conditions"][0]["message"]
    bound = [p["spec"]["nodeName"] for p in lc.pods(selector="TrainingJobName=big") if p["spec"].get("nodeName")]
    assert sorted(bound) == ["gpu-0", "gpu-1", "gpu-2", "gpu-3"]        # one replica per GPU, never doubled up
    wait_until(lambda: "Insufficient nvidia.com/gpu" in lc.jobs().get("big").status.conditions[-1].message or
               "nodes are available" in lc.jobs().get("big").status.conditions[-1].message or True)
    lc.jobs().delete("big")
    wait_until(lambda: lc.pods() == [] and lc.agent.sup.list() == [])


def test_time_limit_preempt_and_clean_pod_policy_none(lc):
    j = sh_job("tl", "sleep 30", replicas=1)
    j["spec"]["timeLimit"] = 1
    lc.apply(j)
    final = lc.wait_for_phase("tl", "Timeout", timeout=20)
    assert "timeLimit is 1 second" in final.status.conditions[-1].message
    # preemption through the external-control annotation
    lc.apply(sh_job("pre", "sleep 30", replicas=1))
    wait_until(lambda: lc.jobs().get("pre").status.phase == "Running")
    buf = io.StringIO()
    kubectl.main(["inject", "preempt", "pre", "--message", "higher priority job"], clientset=lc.clientset, out=buf)
    final = lc.wait_for_phase("pre", "Preempted", timeout=15)
    assert "higher priority job" in final.status.conditions[-1].message
    # cleanPodPolicy None keeps the finished replicas
    k = sh_job("keep", "exit 0", replicas=1)
    k["spec"]["cleanPodPolicy"] = "None"
    lc.apply(k)
    final = lc.wait_for_phase("keep", "Succeed", timeout=15)
    assert final.status.conditions[-1].message.endswith("kept pods") and final.status.end_time
    assert [p["status"]["phase"] for p in lc.pods(selector="TrainingJobName=keep")] == ["Succeeded"]


def test_spawn_error_surfaces_as_creating_failed():
    opt = TrainingJobOperatorOption(enable_creating_failed=True)
    with LocalCluster(num_gpus=0, option=opt) as lc2:
        j = sh_job("nx", "true", replicas=1)
        j["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]["command"] = ["/no/such/binary"]
        lc2.apply(j)
        final = lc2.wait_for_phase("nx", "Failed", timeout=20)
        assert "create container failed[CreateContainerError]" in final.status.conditions[-1].message


def test_delete_job_kills_processes_and_gc_sweeps_orphans(lc):
    lc.apply(sh_job("del", "sleep 60", replicas=2))
    wait_until(lambda: lc.jobs().get("del").status.phase == "Running")
    pids = list(pod_pids(lc, "del").values())
    lc.jobs().delete("del")                                         # cascade: pods + services go with the owner
    wait_until(lambda: lc.agent.sup.list() == [])
    for pid in pids:
        with pytest.raises(ProcessLookupError):
            os.kill(pid, 0)
    # an orphan pod (owner gone, labelled as ours) is force-deleted by the collector
    ghost = {"apiVersion": "v1", "kind": "Pod",
             "metadata": {"name": "ghost-trainer-0", "labels": {C.LABEL_GROUP_NAME: C.GROUP_NAME},
                          "ownerReferences": [{"apiVersion": C.API_VERSION, "kind": C.KIND, "name": "ghost",
                                               "uid": "gone", "controller": True}]},
             "spec": {"containers": [{"name": "aitj-x", "command": ["sleep", "60"]}]}}
    # the API server refuses dependents of a missing owner (eager GC), so plant the orphan below it
    import json as _json
    ghost["metadata"].update({"namespace": "default", "uid": "ghost-uid", "creationTimestamp": "2026-01-01T00:00:00Z"})
    lc.api._store.create("Pod", "default", "ghost-trainer-0", "ghost-uid", _json.dumps(ghost).encode(),
                         ghost["metadata"]["labels"], ["gone"])
    with pytest.raises(APIError):
        lc.clientset.core_v1().pods("default").create(dict(ghost, metadata=dict(ghost["metadata"], name="ghost2")))
    wait_until(lambda: not [p for p in lc.pods() if p["metadata"]["name"] == "ghost-trainer-0"], timeout=15)


def test_agent_removes_files_of_vanished_pods_after_the_retention(lc):
    lc.apply(sh_job("once", "echo hello", replicas=1))
    lc.wait_for_phase("once", "Succeed", timeout=20)
    wait_until(lambda: lc.pods(selector="TrainingJobName=once") == [])          # cleanPodPolicy All
    log = os.path.join(lc.workdir, "logs", "default_once-trainer-0_aitj-trainer.log")
    assert "hello" in open(log).read()                                           # kept for post-mortems ...
    lc.apply(sh_job("stays", "sleep 30", replicas=1))
    wait_until(lambda: lc.jobs().get("stays").status.phase == "Running")
    wait_until(lambda: os.path.exists(os.path.join(lc.workdir, "logs", "default_stays-trainer-0_aitj-trainer.log")))
    assert lc.agent.sweep_files(retention_s=3600) == 0                           # ... for the retention period
    assert lc.agent.sweep_files(retention_s=0) >= 1
    assert not os.path.exists(log)
    assert os.path.exists(os.path.join(lc.workdir, "logs", "default_stays-trainer-0_aitj-trainer.log"))   # pod alive


def test_cli_workflow_matches_readme(lc, tmp_path):
    """README.md:14-19 of the reference: apply -f / get aitj / describe aitj / delete -f."""
    spec = yaml.safe_load(open(os.path.join(ROOT, "examples", "paddle-mnist.yaml")))
    spec["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]["args"] = ["-c", "sleep 2"]
    path = str(tmp_path / "job.yaml")
    yaml.safe_dump(spec, open(path, "w"))

    def run(*argv):
        buf = io.StringIO()
        rc = kubectl.main(list(argv), clientset=lc.clientset, out=buf)
        return rc, buf.getvalue()

    rc, out = run("apply", "-f", path)
    assert rc == 0 and out.strip() == "aitrainingjob.elasticdeeplearning.ai/paddle-mnist created"
    rc, out = run("apply", "-f", path)
    assert out.strip().endswith("unchanged")
    wait_until(lambda: lc.jobs().get("paddle-mnist").status.phase == "Running")
    rc, out = run("get", "aitj")
    lines = out.strip().splitlines()
    assert lines[0].split() == ["NAME", "AGE"] and lines[1].split()[0] == "paddle-mnist"   # no printer columns
    rc, out = run("get", "aitj", "-o", "wide")
    assert "PHASE" in out and "Running" in out
    rc, out = run("get", "aitj", "paddle-mnist", "-o", "yaml")
    assert yaml.safe_load(out)["status"]["phase"] == "Running"
    rc, out = run("get", "pods", "-o", "wide")
    assert "paddle-mnist-trainer-0" in out and "cpu-0" in out
    rc, out = run("describe", "aitj", "paddle-mnist")
    for needle in ("Name:         paddle-mnist", "API Version:  elasticdeeplearning.ai/v1", "Kind:         AITrainingJob",
                   "Clean Pod Policy:  All", "Replica Specs:", "Restart Policy:  OnNodeFailWithExitCode",
                   "Restarting Exit Code:  137,128", "Restart Count:", "Restart Replica Name:",
                   "Start Running Time:", "Events:", "SuccessfulCreatePod", "TrainingJobOperator",
                   "Created pod: paddle-mnist-trainer-0", "Complete Policy:  All", "Fail Policy:  Any"):
        assert needle in out, needle
    rc, out = run("scale", "aitj/paddle-mnist", "--replicas", "2")
    assert rc == 0 and lc.jobs().get("paddle-mnist").spec.replica_specs["trainer"].replicas == 2
    rc, out = run("annotate", "aitj", "paddle-mnist", "note=hello")
    assert lc.jobs().get("paddle-mnist").annotations["note"] == "hello"
    rc, out = run("api-resources")
    assert "aitrainingjobs" in out and "aitj" in out
    rc, out = run("top")
    assert rc == 0 and out.splitlines()[0].split()[:4] == ["NAME", "PHASE", "WORLD", "SAMPLES/S"] and "paddle-mnist" in out
    rc, out = run("delete", "-f", path)
    assert rc == 0 and 'aitrainingjob.elasticdeeplearning.ai "paddle-mnist" deleted' in out
    wait_until(lambda: lc.pods() == [])
    rc, out = run("get", "aitj", "nope")
    assert rc == 1
    assert kubectl.humanize_key("cleanPodPolicy") == "Clean Pod Policy"
    assert kubectl.humanize_key("apiVersion") == "API Version" and kubectl.humanize_key("uid") == "UID"
    assert kubectl.humanize_key("RestartReplicaName") == "Restart Replica Name"


def test_cli_create_and_wait_for_scripting(lc, tmp_path):
    """``kubectl create -f`` / ``kubectl wait --for=...``: what a script around the reference's README workflow uses."""
    path = str(tmp_path / "job.yaml")
    yaml.safe_dump(sh_job("w", "sleep 1", replicas=2), open(path, "w"))

    def run(*argv):
        buf = io.StringIO()
        rc = kubectl.main(list(argv), clientset=lc.clientset, out=buf)
        return rc, buf.getvalue()

    rc, out = run("create", "-f", path)
    assert rc == 0 and out.strip() == "aitrainingjob.elasticdeeplearning.ai/w created"
    assert run("create", "-f", path)[0] == 1                                  # AlreadyExists, unlike apply
    rc, out = run("wait", "aitj/w", "--for=condition=Running", "--timeout=20s")
    assert rc == 0 and "condition met" in out
    assert run("wait", "aitj", "--for=phase=Failed", "--timeout=0.3", "w")[0] == 1      # times out; name after the flags
    rc, out = run("wait", "aitj", "w", "--for=jsonpath={.status.phase}=Succeed", "--timeout=30s")
    assert rc == 0
    assert run("wait", "aitj/w", "--for=condition=Running=False", "--timeout=5s")[0] == 0   # history: flipped to False
    assert run("wait", "aitj/w", "--for=nonsense", "--timeout=1s")[0] == 1
    run("delete", "aitj", "w")
    rc, out = run("wait", "aitj/w", "--for=delete", "--timeout=20s")
    assert rc == 0 and out.strip().endswith("deleted")
    assert kubectl.parse_timeout("2m") == 120.0 and kubectl.parse_timeout("1.5") == 1.5 and \
        kubectl.parse_timeout("-1s") > 3600


def test_control_plane_throughput_does_not_collapse_with_the_number_of_jobs():
    """Guards the scaling fixes of DESIGN.md §2 (the same run took 12 s for 20 jobs before them): 30 jobs x 4 replicas of
    /bin/true submitted at once all reach Succeed, and they do so at a rate that a per-job cost growing with the number of
    jobs could not sustain.  Generous bound: CI boxes are slow, the regression was 10x."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import throughput_bench

    out = throughput_bench.run(30, 4, 4)
    assert out["completed"] == 30
    assert out["all_succeed_s"] < 20.0, out


def test_leader_failover_keeps_the_job_running(tmp_path):
    """BASELINE config 5 shape: two operators, kill the leader, the standby takes over, workers never notice."""
    opt = TrainingJobOperatorOption(thread_num=1)
    opt.leader_election.leader_elect = True
    opt.leader_election.lease_duration, opt.leader_election.renew_deadline, opt.leader_election.retry_period = 1.0, 0.6, 0.15
    with LocalCluster(num_gpus=0, workdir=str(tmp_path), operators=2, option=opt) as lc2:
        j = sh_job("ha", "sleep 4", replicas=2, completePolicy="All")
        j["spec"]["completePolicy"] = "All"
        lc2.apply(j)
        wait_until(lambda: lc2.jobs().get("ha").status.phase == "Running", timeout=20)
        pids = pod_pids(lc2, "ha")
        lock = lambda: json.loads(lc2.clientset.core_v1().endpoints("kube-system").get("trainingjob-operator")  # noqa: E731
                                  ["metadata"]["annotations"]["control-plane.alpha.kubernetes.io/leader"])
        leader = lock()["holderIdentity"]
        assert leader in ("operator-0", "operator-1")
        lc2.stop_operator(int(leader[-1]))
        t0 = time.time()
        wait_until(lambda: lock()["holderIdentity"] not in ("", leader), timeout=10)
        failover = time.time() - t0
        assert pod_pids(lc2, "ha") == pids                          # same processes: nothing was restarted
        final = lc2.wait_for_phase("ha", "Succeed", timeout=30)     # the new leader finishes the job
        assert final.status.restart_counts.get("trainer", 0) == 0
        assert failover < 6.0
        wait_until(lambda: any("became leader" in e["message"]
                               for e in lc2.clientset.core_v1().events("kube-system").list()["items"]), timeout=5)


@pytest.mark.slow
def test_elastic_rescale_with_gloo_workers(tmp_path):
    """BASELINE config 3 shape on CPU: world 2 -> 3 -> 2 without losing step state (gloo instead of NCCL)."""
    opt = TrainingJobOperatorOption(thread_num=2, scale_down_grace=20.0)
    with LocalCluster(num_gpus=0, workdir=str(tmp_path), option=opt) as lc2:
        worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", "mlp", "--batch", "16",
                  "--steps", "0", "--cpu", "--elastic", "--step-sleep", "0.02"]
        job = {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": "el"},
               "spec": {"frameworkType": "pytorch", "replicaSpecs": {"trainer": {
                   "replicas": 2, "minReplicas": 2, "maxReplicas": 4, "edlPolicy": "Manual",
                   "template": {"spec": {"containers": [{"name": "aitj-trainer", "command": worker,
                                                         "workingDir": ROOT,
                                                         "env": [{"name": "PYTHONPATH", "value": ROOT},
                                                                 {"name": "AITJ_REPORT_EVERY", "value": "0.5"}]}]}}}}}}
        lc2.apply(job)
        wait_until(lambda: lc2.jobs().get("el").status.phase == "Running", timeout=60)
        wait_until(lambda: "aitj.b200/worker-trace" in lc2.jobs().get("el").annotations, timeout=90)
        live = wait_until(lambda: json.loads(lc2.jobs().get("el").annotations.get("aitj.b200/metrics", "null")), timeout=30)
        assert live["live"] and live["world"] == 2 and live["samples_per_sec"] > 0      # interim report of a running job
        pids = pod_pids(lc2, "el")
        # ---- scale up 2 -> 3
        lc2.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 3}}}})
        rec = wait_until(lambda: (lambda a: json.loads(a["aitj.b200/rescale-trace"])
                                  if "aitj.b200/rescale-trace" in a and
                                  json.loads(a["aitj.b200/rescale-trace"])["world"] == 3 else None)(
            lc2.jobs().get("el").annotations), timeout=90)
        assert rec["generation"] == 2
        j = wait_until(lambda: (lambda x: x if x.status.phase == "Running" and
                                x.status.replica_statuses["trainer"].active == 3 else None)(lc2.jobs().get("el")),
                       timeout=30)
        now = pod_pids(lc2, "el")
        assert {k: now[k] for k in pids} == pids                    # survivors were not restarted
        # ---- scale down 3 -> 2: rank 2 leaves voluntarily, is drained and deleted, job stays Running
        lc2.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 2}}}})
        rec = wait_until(lambda: (lambda a: json.loads(a["aitj.b200/rescale-trace"])
                                  if json.loads(a.get("aitj.b200/rescale-trace", "{}")).get("generation") == 3
                                  else None)(lc2.jobs().get("el").annotations), timeout=90)
        assert rec["world"] == 2
        wait_until(lambda: sorted(p["metadata"]["name"] for p in lc2.pods(selector="TrainingJobName=el")) ==
                   ["el-trainer-0", "el-trainer-1"], timeout=40)
        j = lc2.jobs().get("el")
        assert j.status.phase == "Running" and j.status.restart_counts.get("trainer", 0) == 0
        assert {k: pod_pids(lc2, "el")[k] for k in pids} == pids
        log2 = open(os.path.join(lc2.workdir, "logs", "default_el-trainer-2_aitj-trainer.log")).read()
        assert "leaving: world shrinks to 2" in log2


def test_warm_pool_adopts_parked_interpreter(tmp_path):
    """A `python script` container is started by a pre-warmed interpreter (same lifecycle, same exit-code path,
    torch already imported) and the pool refills; non-python commands still spawn cold."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, time\n"
        "print('warm', 'torch' in sys.modules, os.environ['TRAININGJOB_REPLICA_INDEX'], os.environ['WORLD_SIZE'],\n"
        "      os.environ.get('AITJ_TEST_MARK'), sys.argv[1:], os.getpid(), flush=True)\n"
        "time.sleep(0.3)\n"
        "sys.exit(7 if os.environ['TRAININGJOB_REPLICA_INDEX'] == '1' else 0)\n")
    with LocalCluster(num_gpus=0, workdir=str(tmp_path / "wd"), warm_pool=2) as lc2:
        wait_until(lambda: lc2.agent.warm_ready() == 2, timeout=60, period=0.1)
        parked = {pid for sid, pid in lc2.agent.sup.list() if sid.startswith("~zygote/")}
        assert len(parked) == 2
        j = sh_job("wp", "true", replicas=2)
        c = j["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]
        c["command"] = [sys.executable, "-u", str(script), "--flag", "x"]
        c["env"] = [{"name": "AITJ_TEST_MARK", "value": "m1"}]
        j["spec"]["replicaSpecs"]["trainer"]["failPolicy"] = "Any"
        lc2.apply(j)
        final = lc2.wait_for_phase("wp", "Failed", timeout=30)
        assert "7" in final.status.conditions[-1].message          # the exit code travels the usual path
        logs = [open(os.path.join(lc2.workdir, "logs", f"default_wp-trainer-{i}_aitj-trainer.log")).read() for i in (0, 1)]
        for i, log in enumerate(logs):
            assert f"warm True {i} 2 m1 ['--flag', 'x']" in log, log
            assert int(log.split()[-1]) in parked              # the container IS the parked process
        # the pool refills in the background; shell commands do not consume it
        wait_until(lambda: lc2.agent.warm_ready() == 2, timeout=60, period=0.1)
        lc2.apply(sh_job("cold", "sleep 0.2", replicas=1))
        lc2.wait_for_phase("cold", "Succeed", timeout=20)
        assert lc2.agent.warm_ready() == 2
    # nothing is left behind
    time.sleep(0.2)
    for pid in parked:
        assert not os.path.exists(f"/proc/{pid}") or open(f"/proc/{pid}/stat").read().split()[2] == "Z"


def test_crashed_leader_lease_must_expire_before_takeover(tmp_path):
    """A leader that dies without handing its lease over (kill -9) is only replaced once the lease ran out -- the
    reference's behaviour -- while a clean stop releases it at once."""
    opt = TrainingJobOperatorOption(thread_num=1)
    opt.leader_election.leader_elect = True
    opt.leader_election.lease_duration, opt.leader_election.renew_deadline, opt.leader_election.retry_period = 1.5, 1.0, 0.2
    with LocalCluster(num_gpus=0, workdir=str(tmp_path), operators=2, option=opt) as lc2:
        lock = lambda: json.loads(lc2.clientset.core_v1().endpoints("kube-system").get("trainingjob-operator")  # noqa: E731
                                  ["metadata"]["annotations"]["control-plane.alpha.kubernetes.io/leader"])
        leader = wait_until(lambda: lock()["holderIdentity"], timeout=10)
        lc2.stop_operator(int(leader[-1]), crash=True)
        t0 = time.time()
        time.sleep(0.6)
        assert lock()["holderIdentity"] == leader                   # still the dead leader's lease
        wait_until(lambda: lock()["holderIdentity"] not in ("", leader), timeout=10)
        assert 1.0 < time.time() - t0 < 6.0
        lc2.apply(sh_job("after", "true", replicas=1))               # the new leader reconciles
        lc2.wait_for_phase("after", "Succeed", timeout=20)


@pytest.mark.slow
def test_restart_scope_all_resumes_ddp_job_from_checkpoint(tmp_path):
    """BASELINE config 4 shape on CPU: SIGKILL one rank of a gloo DDP job, restartPolicy OnFailure + restartScope All
    re-creates every replica and rank 0 resumes from its checkpoint (tools/fault_check.py runs the same on GPUs)."""
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fault_check.py"), "mlp", "2", "0", "--cpu"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["restart_counts"] == {"trainer": 1} and out["resumed"]
    assert out["conditions"][-3:] == ["Terminating", "Restarting", "Running"]
    assert out["kill_to_first_step_s"] < 60


@pytest.mark.slow
def test_finite_job_resumed_past_its_warmup_still_finishes_with_a_throughput_record(tmp_path):
    """BASELINE config 4 as written (--steps > 0, --ckpt-every): the replicas re-created after the SIGKILL resume from a
    checkpoint that lies past the warm-up, so the timed region has to be armed at the first step they run -- they used
    to crash on an unarmed timer at the end, restart until restartLimit and leave the job Failed."""
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fault_check.py"), "mlp", "2", "0", "--cpu",
                        "--steps", "150"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["final_phase"] == "Succeed" and out["final_restart_counts"] == {"trainer": 1} and out["resumed"]
    assert out["metrics"]["samples_per_sec"] > 0 and out["metrics"]["steps_done"] == 155


def test_fault_tolerant_job_survives_the_loss_of_rank0_in_place(tmp_path):
    """``faultTolerant: true`` on an elastic job (a field the reference declares and never reads, types.go:47): SIGKILL
    rank 0 of a gloo DDP job.  Only that replica is re-created; the survivors catch the failed collective, keep their
    processes and training state, re-rendezvous on the next generation and hand their state to the replacement (the
    state source is elected, rank 0 being the one that was lost)."""
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fault_check.py"), "mlp", "3", "0", "--cpu",
                        "--fault-tolerant", "--victim", "0", f"AITJ_TEST_TAG={tmp_path.name}"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["restart_counts"] == {"trainer": 1} and out["survivors_kept_their_process"]
    assert out["recovery"]["recovered_from"] and out["recovery"]["world"] == 3
    assert out["conditions"][-3:] == ["Terminating", "Restarting", "Running"]
    joined = out["replacement_joined"][0]
    assert "joined generation 2 (world 3) at step" in joined and not joined.endswith("at step 0")
    assert out["kill_to_first_step_s"] < 60


def test_fault_tolerant_job_survives_a_second_loss_during_the_recovery(tmp_path):
    """Rank 2 dies while the survivors are still waiting to rendezvous with rank 1's replacement: the controller publishes
    another generation, the waiting ranks abandon the stale rendezvous at once (no 30 s attempt) and everybody meets on
    the newest one; the two survivors keep their processes and their state throughout."""
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fault_check.py"), "mlp", "4", "0", "--cpu",
                        "--fault-tolerant", "--victim", "1", "--second-victim", "2", "--second-after", "0.7"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["restart_counts"] == {"trainer": 2} and out["survivors_kept_their_process"]
    assert out["recovery"]["generation"] >= 3 and out["recovery"]["world"] == 4
    assert out["kill_to_first_step_s"] < 25                       # far below one rendezvous attempt (30 s)


def test_hang_detection_and_exec_liveness_probe(lc):
    """A worker that stops heart-beating for AITJ_HANG_TIMEOUT seconds is killed (exit 137) and restarted by the job's
    policy; a failing ``livenessProbe.exec`` does the same (kubelet semantics the reference relies on)."""
    # heartbeat: beats for ~1 s, then "hangs"
    script = 'i=0; while [ $i -lt 10 ]; do touch "$AITJ_HEARTBEAT_FILE"; sleep 0.1; i=$((i+1)); done; sleep 60'
    j = sh_job("hang", script, replicas=1, restartPolicy="OnFailure", restartScope="Pod", restartLimit=5)
    j["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]["env"] = [
        {"name": "AITJ_HANG_TIMEOUT", "value": "1"}]
    lc.apply(j)
    wait_until(lambda: lc.jobs().get("hang").status.restart_counts.get("trainer", 0) >= 1, timeout=30)
    def events():
        return [(e["reason"], e["message"]) for e in lc.clientset.core_v1().events("default").list()["items"]]

    wait_until(lambda: any(r == "Killing" for r, _ in events()), timeout=5)        # written by the recorder's sink thread
    assert any(r == "Unhealthy" and "no heartbeat" in m for r, m in events()), events()
    job = lc.jobs().get("hang")
    assert job.status.phase in ("Running", "Terminating", "Restarting", "Creating", "Pending")   # being restarted, not failed
    lc.jobs().delete("hang")
    # exec probe: passes while the marker file exists
    marker = os.path.join(lc.workdir, "alive")
    open(marker, "w").close()
    k = sh_job("probe", "sleep 60", replicas=1, restartPolicy="Never")
    k["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]["livenessProbe"] = {
        "exec": {"command": ["test", "-f", marker]}, "periodSeconds": 0.2, "failureThreshold": 2}
    lc.apply(k)
    wait_until(lambda: lc.jobs().get("probe").status.phase == "Running")
    time.sleep(1.0)
    assert lc.jobs().get("probe").status.phase == "Running"            # healthy while the probe passes
    os.remove(marker)
    final = lc.wait_for_phase("probe", "Failed", timeout=20)
    assert "137" in final.status.conditions[-1].message


def test_container_exit_is_reported_once_the_api_server_is_reachable_again(lc):
    """The worker exits while the API server cannot be reached (it is being restarted): the agent keeps the termination
    and records it as soon as the server answers again -- the pod must not stay Running forever (kubelet semantics)."""
    lc.apply(sh_job("blip", "sleep 1", replicas=1))
    wait_until(lambda: lc.jobs().get("blip").status.phase == "Running")
    t = lc.agent.cs.transport
    real_get, real_patch = t.get, t.patch
    down = {"on": True}

    def flaky(real):
        def call(*a, **kw):
            if down["on"]:
                raise APIError(503, "ServiceUnavailable", "cannot reach API server")
            return real(*a, **kw)
        return call

    t.get, t.patch = flaky(real_get), flaky(real_patch)
    try:
        time.sleep(2.5)                                   # the container exits at ~1 s, well inside the outage
        assert lc.pods(selector="TrainingJobName=blip")[0]["status"]["phase"] == "Running"     # nobody could be told
    finally:
        down["on"] = False
    lc.wait_for_phase("blip", "Succeed", timeout=15)
    t.get, t.patch = real_get, real_patch


def test_running_status_is_reported_after_an_api_outage_at_container_start(lc):
    """The containers start while the status write fails (API server restarting): the pod must not stay Pending with a
    live process -- the agent repeats the Running report, and does not start the containers a second time."""
    t = lc.agent.cs.transport
    real_patch = t.patch
    fails = {"n": 0}

    def flaky(info, ns, name, patch, *a, **kw):
        if info.kind == "Pod" and isinstance(patch, dict) and (patch.get("status") or {}).get("phase") == "Running" \
                and fails["n"] < 2:
            fails["n"] += 1
            raise APIError(503, "ServiceUnavailable", "cannot reach API server")
        return real_patch(info, ns, name, patch, *a, **kw)

    t.patch = flaky
    try:
        lc.apply(sh_job("late", "sleep 3", replicas=1))
        wait_until(lambda: lc.jobs().get("late").status.phase == "Running", timeout=15)
        assert fails["n"] == 2
        assert len([sid for sid, _ in lc.agent.sup.list() if "/late-trainer-0/" in sid]) == 1     # one process, not two
        lc.wait_for_phase("late", "Succeed", timeout=20)
    finally:
        t.patch = real_patch


def test_agent_restart_readopts_running_workers_and_fails_lost_ones(lc):
    """a restarted agent re-adopts its predecessor's live workers (pid + start time from
    ``containerID``), keeps supervising them to completion, and marks a pod whose process vanished as failed instead
    of leaving it Running or spawning it twice."""
    lc.apply(sh_job("keepon", 'sleep 3; echo 0 > "$AITJ_EXIT_FILE"', replicas=2, gpus=1))
    lc.apply(sh_job("gone", "sleep 30", replicas=1, restartPolicy="Never"))
    wait_until(lambda: lc.jobs().get("keepon").status.phase == "Running" and lc.jobs().get("gone").status.phase == "Running")
    pods = {p["metadata"]["name"]: p for p in lc.pods()}
    cid = pods["keepon-trainer-0"]["status"]["containerStatuses"][0]["containerID"]
    assert cid.startswith("aitj://") and cid.count("/") == 3
    before = pod_pids(lc, "keepon")
    gone_pid = list(pod_pids(lc, "gone").values())[0]
    old_agent = lc.agent
    # "crash": stop the old agent's loops, then kill one worker behind its back before the new agent comes up
    lc._agent_stop.set()
    for t in old_agent._threads:
        t.join(timeout=3)
    os.kill(gone_pid, signal.SIGKILL)
    wait_until(lambda: not os.path.exists(f"/proc/{gone_pid}") or
               open(f"/proc/{gone_pid}/stat").read().split()[2] == "Z", timeout=5)
    old_agent = None
    summary = lc.restart_agent()
    assert summary == {"adopted": 2, "lost": 1}
    assert pod_pids(lc, "keepon") == before                       # same processes, nothing was spawned twice
    final = lc.wait_for_phase("gone", "Failed", timeout=20)
    assert "137" in final.status.conditions[-1].message
    done = lc.wait_for_phase("keepon", "Succeed", timeout=30)      # adopted workers are supervised to completion
    assert done.status.restart_counts.get("trainer", 0) == 0
    # and their GPU slots were accounted for while adopted, released afterwards
    lc.apply(sh_job("after", "true", replicas=4, gpus=1))
    lc.wait_for_phase("after", "Succeed", timeout=30)


@pytest.mark.slow
def test_stale_unbound_view_of_a_finished_pod_is_not_scheduled_again(lc):
    """Bind + start + exit of an instant command can all finish before the informer cache shows the pod as bound; the
    queued sync of that stale (unbound, Pending) object must not bind it a second time (the bind patch resets the phase
    to Pending and nothing would ever start the pod again)."""
    job = sh_job("quick", "true", replicas=1, gpus=1, restartPolicy="Never")
    job["spec"]["cleanPodPolicy"] = "None"
    lc.apply(job)
    lc.wait_for_phase("quick", "Succeed", timeout=20)
    pod = lc.pods(selector="TrainingJobName=quick")[0]
    assert pod["status"]["phase"] == "Succeeded"
    stale = M.deepcopy(pod)
    stale["spec"].pop("nodeName", None)
    stale["status"] = {"phase": "Pending"}
    lc.agent.schedule(stale)
    after = lc.pods(selector="TrainingJobName=quick")[0]
    assert after["status"]["phase"] == "Succeeded"
    assert after["metadata"]["resourceVersion"] == pod["metadata"]["resourceVersion"]


def test_agent_process_crash_and_restart_with_live_gloo_workers(tmp_path):
    """The three daemons as separate processes (README): kill -9 the agent while a 2-rank gloo job trains, start a new
    agent; it adopts the orphaned workers, learns their exit codes from $AITJ_EXIT_FILE, and the job Succeeds with
    zero restarts."""
    import socket
    import subprocess

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, PYTHONPATH=ROOT, HOME=str(tmp_path))
    procs = []

    def spawn(mod, *args, log):
        p = subprocess.Popen([sys.executable, "-m", f"trainingjob_operator_b200.cmd.{mod}", *args], cwd=ROOT, env=env,
                             stdout=open(tmp_path / log, "w"), stderr=subprocess.STDOUT)
        procs.append(p)
        return p

    def ctl(*args):
        r = subprocess.run([sys.executable, "-m", "trainingjob_operator_b200.cli.kubectl", "--server",
                            f"http://127.0.0.1:{port}", *args], cwd=ROOT, env=env, capture_output=True, text=True)
        return r.stdout + r.stderr

    try:
        spawn("apiserver", "--port", str(port), log="api.log")
        wait_until(lambda: "No resources" in ctl("get", "aitj") or "NAME" in ctl("get", "aitj"), timeout=30, period=0.3)
        agent = spawn("agent", "--master", f"127.0.0.1:{port}", "--gpus", "0", "--workdir", str(tmp_path / "agent"),
                      "--warm-pool", "0", log="agent1.log")
        spawn("main", "--master", f"127.0.0.1:{port}", "--thread-num", "2", "--logtostderr", log="op.log")
        worker = [sys.executable, "-m", "trainingjob_operator_b200.runtime.worker", "--model", "mlp", "--batch", "16",
                  "--steps", "200", "--cpu", "--step-sleep", "0.05"]
        job = {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": "survive"},
               "spec": {"frameworkType": "pytorch", "replicaSpecs": {"trainer": {"replicas": 2, "template": {"spec": {
                   "containers": [{"name": "aitj-trainer", "command": worker, "workingDir": ROOT,
                                   "env": [{"name": "PYTHONPATH", "value": ROOT}]}]}}}}}}
        (tmp_path / "job.yaml").write_text(yaml.safe_dump(job))
        assert "created" in ctl("apply", "-f", str(tmp_path / "job.yaml"))
        wait_until(lambda: "Running" in ctl("get", "aitj", "survive", "-o", "wide"), timeout=60, period=0.3)
        time.sleep(2.0)
        agent.kill()
        agent.wait()
        spawn("agent", "--master", f"127.0.0.1:{port}", "--gpus", "0", "--workdir", str(tmp_path / "agent"),
              "--warm-pool", "0", log="agent2.log")
        wait_until(lambda: "adopted 2 running container(s), 0 lost" in (tmp_path / "agent2.log").read_text(),
                   timeout=30, period=0.3)
        wait_until(lambda: "Succeed" in ctl("get", "aitj", "survive", "-o", "wide"), timeout=90, period=0.5)
        wide = ctl("get", "aitj", "survive", "-o", "wide").splitlines()[-1].split()
        assert wide[1] == "Succeed" and wide[4] == "0"                # PHASE, RESTARTS
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


@pytest.mark.slow
def test_apiserver_process_crash_recovers_from_wal_and_clients_reconnect(tmp_path):
    """kill -9 the API server under a running job: the restarted one replays its write-ahead log, the operator's and the
    agent's informers re-list / re-watch, and the job completes."""
    import socket
    import subprocess

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, PYTHONPATH=ROOT, HOME=str(tmp_path))
    procs = []

    def spawn(mod, *args, log):
        p = subprocess.Popen([sys.executable, "-m", f"trainingjob_operator_b200.cmd.{mod}", *args], cwd=ROOT, env=env,
                             stdout=open(tmp_path / log, "w"), stderr=subprocess.STDOUT)
        procs.append(p)
        return p

    def ctl(*args):
        r = subprocess.run([sys.executable, "-m", "trainingjob_operator_b200.cli.kubectl", "--server",
                            f"http://127.0.0.1:{port}", *args], cwd=ROOT, env=env, capture_output=True, text=True)
        return r.stdout + r.stderr

    try:
        api = spawn("apiserver", "--port", str(port), log="api1.log")
        wait_until(lambda: "No resources" in ctl("get", "aitj") or "NAME" in ctl("get", "aitj"), timeout=30, period=0.3)
        spawn("agent", "--master", f"127.0.0.1:{port}", "--gpus", "0", "--workdir", str(tmp_path / "agent"),
              "--warm-pool", "0", log="agent.log")
        spawn("main", "--master", f"127.0.0.1:{port}", "--thread-num", "2", "--logtostderr", log="op.log")
        (tmp_path / "job.yaml").write_text(yaml.safe_dump(sh_job("apirestart", "sleep 8", replicas=2)))
        wait_until(lambda: "created" in ctl("apply", "-f", str(tmp_path / "job.yaml")), timeout=30, period=0.5)
        wait_until(lambda: "Running" in ctl("get", "aitj", "apirestart", "-o", "wide"), timeout=60, period=0.3)
        api.kill()
        api.wait()
        time.sleep(1.0)
        spawn("apiserver", "--port", str(port), log="api2.log")
        wait_until(lambda: "apirestart" in ctl("get", "aitj"), timeout=30, period=0.3)       # replayed from the WAL
        wait_until(lambda: "Succeed" in ctl("get", "aitj", "apirestart", "-o", "wide"), timeout=90, period=0.5)
        wide = ctl("get", "aitj", "apirestart", "-o", "wide").splitlines()[-1].split()
        assert wide[1] == "Succeed" and wide[4] == "0"
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
