"""The reconcile engine as a pure function: recorded (Observation -> Decision) pairs replay bit for bit, and the facts
that matter in each scenario are asserted by hand (table-driven state-machine tests; the reference has no
tests at all).  Regenerate the recording with ``python tools/record_engine_golden.py`` and review the diff."""
import copy
import json
import os

import pytest

from trainingjob_operator_b200.controller import engine, replay
from trainingjob_operator_b200.controller.pod import RESTART_MATRIX, Health

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "engine_cases.json")))


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_recorded_pass_replays_to_the_recorded_decision(name):
    case = GOLDEN[name]
    assert replay.replay(copy.deepcopy(case)) == case["decision"]


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_engine_is_deterministic_and_touches_nothing_but_its_private_copies(name):
    case = GOLDEN[name]
    a = replay.replay(copy.deepcopy(case))
    b = replay.replay(copy.deepcopy(case))
    assert a == b
    frozen = json.dumps(case["observation"], sort_keys=True)
    replay.replay(case)                       # from_json deep-copies: the recorded observation is never edited
    assert json.dumps(case["observation"], sort_keys=True) == frozen


def dec(name):
    return GOLDEN[name]["decision"]


def test_first_pass_is_two_phase_ports_then_everything():
    assert dec("first_pass_asks_for_ports")["ports_wanted"] == 3          # MASTER_PORT + one host port per replica
    assert dec("first_pass_asks_for_ports")["pod_creates"] == [] and not dec("first_pass_asks_for_ports")["write_status"]
    d = dec("first_pass_creates_everything")
    assert [c["template"]["metadata"]["name"] for c in d["pod_creates"]] == ["job-trainer-0", "job-trainer-1"]
    assert d["status"]["rendezvous"] == {"generation": 1, "worldSizes": {"trainer": 2}, "masterPort": 40001,
                                         "changedAt": "2026-09-21T12:00:00Z"}
    ports = json.loads(d["job_annotations"]["aitj.b200/host-ports"])
    assert ports == {"trainer/0/2222": 41000, "trainer/1/2222": 41001}
    svc = {s["metadata"]["name"]: s["spec"]["ports"][0] for _, s in d["service_creates"]}
    assert svc["job-trainer-1"] == {"name": "aitj-2222", "port": 2222, "hostPort": 41001}   # recorded in the Service
    env = {e["name"]: e["value"] for e in d["pod_creates"][1]["template"]["spec"]["containers"][0]["env"]}
    assert env["TRAINER_ADDRS"] == "127.0.0.1:41000,127.0.0.1:41001" and env["AITJ_HOST_PORTS"] == "41001"
    assert env["MASTER_PORT"] == "40001" and d["status"]["phase"] == "Pending"


def test_restart_decisions_follow_scope_and_cause():
    d = dec("crash_restarts_scope_all")
    assert sorted(x["name"] for x in d["pod_deletes"]) == ["job-ps-0", "job-trainer-0", "job-trainer-1"]
    assert d["status"]["RestartCount"] == {"trainer": 1, "ps": 1} and d["status"]["RestartReplicaName"] == "trainer"
    assert d["status"]["rendezvous"]["generation"] == 2 and d["status"]["rendezvous"]["masterPort"] == 40002
    assert d["status"]["phase"] == "Terminating" and d["counters"][0] == ["aitj_restarts_total", {"scope": "All"}]
    n = dec("node_lost_force_deletes_one_pod")
    assert [(x["name"], x["grace"]) for x in n["pod_deletes"]] == [("job-trainer-1", 0)]     # grace 0 on a dead node
    f = dec("exit_code_not_listed_fails_job")
    assert f["job_annotations"]["Failed"].startswith("pod job-trainer-1 is failed, container aitj-trainer on node gpu-1")
    assert f["status"]["RestartCount"] == {"trainer": 0}


def test_barrier_and_termination_paths():
    assert dec("restart_barrier_holds_while_victims_exist")["status"]["RestartReplicaName"] == "trainer"
    lifted = dec("restart_barrier_lifts_when_victims_are_gone")["status"]
    assert lifted["phase"] == "Restarting" and lifted["RestartReplicaName"] == ""
    assert lifted["conditions"][-1]["message"] == "All pods are restarting now"
    t = dec("complete_all_terminates_and_deletes")
    assert t["job_annotations"]["Succeed"] == "job job completed" and t["service_deletes"] == [["default", "job-trainer-0"]]
    assert t["pod_creates"] == [] and t["service_creates"] == []          # nothing is born into a dying job
    fin = dec("parked_verdict_finalises_when_pods_are_gone")["status"]
    assert fin["phase"] == "Succeed" and fin["endTime"] and fin["conditions"][-1]["message"] == "job job completed; deleted pods"
    keep = dec("clean_pod_policy_none_keeps_pods")
    assert keep["pod_deletes"] == [] and keep["status"]["conditions"][-1]["message"] == "job job completed; kept pods"
    assert "timeLimit is 60 second" in dec("time_limit_expired")["job_annotations"]["Timeout"]
    assert dec("preempted_from_outside")["status"]["conditions"][-1]["message"] == "capacity needed elsewhere; deleting pods"
    assert "create container failed[ErrImagePull]" in dec("start_error_outlives_window")["job_annotations"]["Failed"]


def test_elastic_decisions():
    m = dec("scale_down_marks_surplus_draining")
    assert [p[1] for p in m["pod_patches"]] == ["job-trainer-2", "job-trainer-3"] and m["pod_deletes"] == []
    assert m["status"]["replicaStatuses"]["trainer"] == {"active": 2} and m["status"]["rendezvous"]["generation"] == 2
    dl = dec("scale_down_deletes_drained_surplus")
    assert [x["name"] for x in dl["pod_deletes"]] == ["job-trainer-2"]      # exited; trainer-3 still has grace left
    up = dec("scale_up_new_generation")
    env = {e["name"]: e["value"] for e in up["pod_creates"][0]["template"]["spec"]["containers"][0]["env"]}
    assert env["WORLD_SIZE"] == "4" and env["AITJ_RENDEZVOUS_GENERATION"] == "2" and env["MASTER_PORT"] == "40002"
    assert dec("auto_grows_into_free_gpus")["spec_patch"] == {"spec": {"replicaSpecs": {"trainer": {"replicas": 4}}}}
    y = dec("auto_yields_to_more_important_work")
    assert y["spec_patch"] == {"spec": {"replicaSpecs": {"trainer": {"replicas": 3}}}} and not y["write_status"]


def test_restart_matrix_is_the_whole_policy():
    """restartPolicy x what happened -> restart?  (pod.go:385-419 as a table; ``Always`` does not restart a success.)"""
    crashed = {Health.CRASHED, Health.CRASHED_LISTED}
    assert RESTART_MATRIX["Always"] == crashed | {Health.NODE_LOST}
    assert RESTART_MATRIX["OnFailure"] == crashed
    assert RESTART_MATRIX["OnNodeFail"] == {Health.NODE_LOST}
    assert RESTART_MATRIX["ExitCode"] == {Health.CRASHED_LISTED}
    assert RESTART_MATRIX["OnNodeFailWithExitCode"] == {Health.CRASHED_LISTED, Health.NODE_LOST}
    assert RESTART_MATRIX["Never"] == frozenset()
    assert all(Health.COMPLETED not in v and Health.RUNNING not in v for v in RESTART_MATRIX.values())
    assert [r.__name__ for r in engine.ROLE_VERDICTS] == ["_verdict_any_ok", "_verdict_any_bad", "_verdict_rank0_ok",
                                                          "_verdict_rank0_bad"]


def test_every_pass_of_a_live_cluster_replays_to_the_decision_it_took(tmp_path, monkeypatch):
    """AITJ_RECORD_DIR on a real (in-process) cluster: jobs that complete, a SIGKILLed replica that is restarted, a live
    rescale and a deletion -- every pass the operator made is written down, and every one of them, replayed offline
    from its JSON alone, yields exactly the decision (creates, deletes, status, annotations, requeues) that was applied.
    This is what makes a recorded decision debuggable after the fact: the engine has no hidden inputs."""
    import glob
    import os
    import signal
    import sys
    import time

    from trainingjob_operator_b200.api import constants as C
    from trainingjob_operator_b200.cmd.local import LocalCluster
    from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption
    from trainingjob_operator_b200.controller import replay as R

    rec_dir = tmp_path / "passes"
    monkeypatch.setenv("AITJ_RECORD_DIR", str(rec_dir))

    def job(name, command, replicas, **role):
        c = {"name": "aitj-trainer", "command": command}
        r = dict({"replicas": replicas, "template": {"spec": {"terminationGracePeriodSeconds": 0, "containers": [c]}}}, **role)
        return {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": name},
                "spec": {"cleanPodPolicy": "All", "replicaSpecs": {"trainer": r}}}

    def wait(fn, timeout=30.0):
        t0 = time.time()
        while time.time() - t0 < timeout:
            v = fn()
            if v:
                return v
            time.sleep(0.01)
        raise TimeoutError

    opt = TrainingJobOperatorOption(thread_num=2, scale_down_grace=1.0)       # (default drain grace: 30 s)
    with LocalCluster(num_gpus=0, workdir=str(tmp_path / "wd"), option=opt) as lc:
        lc.apply(job("ok", ["/bin/sh", "-c", "sleep 0.2"], 2))
        lc.apply(job("bad", ["/bin/sh", "-c", "sleep 0.1; exit 3"], 2))
        lc.apply(job("el", ["/bin/sleep", "600"], 2, minReplicas=1, maxReplicas=4, edlPolicy="Manual",
                     restartPolicy="OnFailure", restartScope="Pod"))
        wait(lambda: lc.jobs().get("ok").status.phase == "Succeed")
        wait(lambda: lc.jobs().get("bad").status.phase == "Failed")
        wait(lambda: lc.jobs().get("el").status.phase == "Running")
        pid = next(p for sid, p in lc.agent.sup.list() if "/el-trainer-1/" in sid)
        os.kill(pid, signal.SIGKILL)
        wait(lambda: (lambda j: j.status.phase == "Running" and j.status.restart_counts.get("trainer") == 1)(
            lc.jobs().get("el")))
        lc.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 3}}}})
        wait(lambda: (lambda j: j.status.phase == "Running" and j.status.replica_statuses["trainer"].active == 3)(
            lc.jobs().get("el")))
        lc.jobs().patch("el", {"spec": {"replicaSpecs": {"trainer": {"replicas": 1}}}})
        wait(lambda: len(lc.pods(selector="TrainingJobName=el")) == 1)
        lc.jobs().delete("el")
        wait(lambda: not lc.pods(selector="TrainingJobName=el"))
    files = sorted(glob.glob(str(rec_dir / "*.json")))
    assert len(files) >= 20, len(files)
    kinds = set()
    for f in files:
        case = json.load(open(f))
        assert R.replay(case) == case["decision"], f
        d = case["decision"]
        kinds.update(k for k in ("pod_creates", "pod_deletes", "service_creates", "service_deletes") if d.get(k))
        if d.get("ports_wanted"):
            kinds.add("ports_wanted")
        kinds.add("phase:" + (d["status"].get("phase") or ""))
    # the recording really covers the interesting decisions, not only idle passes
    assert {"pod_creates", "pod_deletes", "service_creates", "ports_wanted"} <= kinds, kinds
    assert {"phase:Running", "phase:Succeed", "phase:Failed"} <= kinds, kinds
