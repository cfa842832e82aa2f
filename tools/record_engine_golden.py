"""Writes tests/golden/engine_cases.json: scenario Observations and the Decisions the engine takes on them.

    python tools/record_engine_golden.py            # regenerate (review the diff before committing!)

Every case is built from plain data with a fixed clock and fixed ports, so the file only changes when the engine's
behaviour does.  tests/test_engine_golden.py replays the file and, separately, asserts the interesting facts of each
scenario by hand.
"""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from trainingjob_operator_b200.api import constants as C  # noqa: E402
from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob  # noqa: E402
from trainingjob_operator_b200.api.types import AITrainingJob  # noqa: E402
from trainingjob_operator_b200.controller import replay  # noqa: E402

NOW = "2026-09-21T12:00:00Z"
EPOCH = 1789992000.0     # the same instant as NOW (a replay rebuilds `now` from this float)
UID = "11111111-2222-3333-4444-555555555555"


def job(roles=None, status=None, annotations=None, **spec):
    roles = roles or {"trainer": {}}
    rs = {}
    for role, over in roles.items():
        r = {"replicas": 2, "template": {"spec": {"containers": [
            {"name": f"aitj-{role.lower()}", "image": "img", "command": ["run"],
             "ports": [{"name": "aitj-port", "containerPort": 2222}]}]}}}
        r.update(over)
        rs[role] = r
    d = {"apiVersion": C.API_VERSION, "kind": C.KIND,
         "metadata": {"name": "job", "namespace": "default", "uid": UID, "resourceVersion": "7",
                      "annotations": dict(annotations or {})},
         "spec": dict({"replicaSpecs": rs}, **spec), "status": status or {}}
    j = AITrainingJob.from_dict(d)
    set_defaults_aitrainingjob(j)
    return j.to_dict()


def pod(role, index, phase="Running", node="gpu-0", exit_codes=None, waiting=None, annotations=None,
        unschedulable=None, deleting=False):
    rt = role.lower()
    name = f"job-{rt}-{index}"
    css = []
    if exit_codes is not None:
        css = [{"name": f"aitj-{rt}", "state": {"terminated": {"exitCode": c, "reason": "Error" if c else "Completed"}}}
               for c in exit_codes]
    elif waiting is not None:
        css = [{"name": f"aitj-{rt}", "state": {"waiting": {"reason": waiting, "message": "boom"}}}]
    elif phase == "Running":
        css = [{"name": f"aitj-{rt}", "state": {"running": {}}}]
    p = {"apiVersion": "v1", "kind": "Pod",
         "metadata": {"name": name, "namespace": "default", "uid": f"uid-{name}", "creationTimestamp": NOW,
                      "labels": {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: "job", C.LABEL_REPLICA_NAME: rt,
                                 C.LABEL_REPLICA_INDEX: str(index)},
                      "ownerReferences": [{"apiVersion": C.API_VERSION, "kind": C.KIND, "name": "job", "uid": UID,
                                           "controller": True}]},
         "spec": {"nodeName": node} if node else {}, "status": {"phase": phase, "containerStatuses": css}}
    if annotations:
        p["metadata"]["annotations"] = annotations
    if unschedulable:
        p["status"]["conditions"] = [{"type": "PodScheduled", "status": "False", "message": unschedulable}]
    if deleting:
        p["metadata"]["deletionTimestamp"] = NOW
    return p


def obs(j, pods=(), services=(), ready=("gpu-0", "gpu-1", "gpu-2", "gpu-3", "cpu-0"), spare=(), cluster=None,
        window=None, grace=30.0):
    return {"job": j, "pods": list(pods), "services": list(services), "ready_nodes": sorted(ready), "now": NOW,
            "now_epoch": EPOCH,
            "options": {"window": window or {"restart_period": 0.0, "duration_period": 900.0,
                                             "fail_after_window": False},
                        "scale_down_grace": grace, "master_url": ""},
            "cluster": cluster, "spare_ports": list(spare)}


RUNNING = {"phase": "Running", "conditions": [{"type": "Running", "status": "True", "reason": "TrainingJobRunning",
                                               "message": "all pods are running", "lastProbeTime": NOW,
                                               "lastTransitionTime": NOW}],
           "replicaStatuses": {}, "RestartReplicaName": "", "startTime": NOW, "startRunningTime": NOW,
           "rendezvous": {"generation": 1, "worldSizes": {"trainer": 2}, "masterPort": 40001}}
PORTS = {"aitj.b200/host-ports": json.dumps({"trainer/0/2222": 41000, "trainer/1/2222": 41001}, sort_keys=True)}


def cases():
    out = {}
    out["first_pass_asks_for_ports"] = obs(job())
    out["first_pass_creates_everything"] = obs(job(), spare=(40001, 41000, 41001))
    out["all_running"] = obs(job(status={"phase": "Creating", "conditions": [], "replicaStatuses": {},
                                         "RestartReplicaName": "", "startTime": NOW,
                                         "rendezvous": RUNNING["rendezvous"]}, annotations=PORTS),
                             [pod("trainer", 0), pod("trainer", 1, node="gpu-1")])
    out["crash_restarts_scope_all"] = obs(
        job(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": "All"}, "ps": {"replicas": 1}},
            status=dict(RUNNING, rendezvous={"generation": 1, "worldSizes": {"trainer": 2, "ps": 1},
                                             "masterPort": 40001}),
            annotations={"aitj.b200/host-ports": json.dumps({"trainer/0/2222": 41000, "trainer/1/2222": 41001,
                                                             "ps/0/2222": 41002}, sort_keys=True)}),
        [pod("trainer", 0), pod("trainer", 1, phase="Failed", exit_codes=[137], node="gpu-1"),
         pod("ps", 0, node="cpu-0")], spare=(40002,))
    out["exit_code_not_listed_fails_job"] = obs(
        job(roles={"trainer": {"restartPolicy": "ExitCode"}}, restartingExitCode="137,128", status=RUNNING,
            annotations=PORTS),
        [pod("trainer", 0), pod("trainer", 1, phase="Failed", exit_codes=[1], node="gpu-1")])
    out["node_lost_force_deletes_one_pod"] = obs(
        job(roles={"trainer": {"restartPolicy": "OnNodeFail", "restartScope": "Pod"}}, status=RUNNING,
            annotations=PORTS),
        [pod("trainer", 0), pod("trainer", 1, node="gpu-1")], ready=("gpu-0", "cpu-0"), spare=(40002,))
    barrier = dict(RUNNING, phase="Terminating", RestartReplicaName="trainer", RestartCount={"trainer": 1})
    barrier["conditions"] = [dict(RUNNING["conditions"][0], status="False"),
                             {"type": "Terminating", "status": "True", "reason": "TrainingJobTerminating",
                              "message": "restart times is 1, x ", "lastProbeTime": NOW, "lastTransitionTime": NOW}]
    out["restart_barrier_holds_while_victims_exist"] = obs(
        job(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": "All"}}, status=barrier,
            annotations=PORTS), [pod("trainer", 0, deleting=True)])
    out["restart_barrier_lifts_when_victims_are_gone"] = obs(
        job(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": "All"}}, status=barrier,
            annotations=PORTS), [])
    out["complete_all_terminates_and_deletes"] = obs(
        job(status=RUNNING, annotations=PORTS),
        [pod("trainer", 0, phase="Succeeded", exit_codes=[0]),
         pod("trainer", 1, phase="Succeeded", exit_codes=[0], node="gpu-1")],
        services=[{"kind": "Service", "metadata": {"name": "job-trainer-0", "namespace": "default",
                                                    "labels": {C.LABEL_REPLICA_NAME: "trainer",
                                                               C.LABEL_REPLICA_INDEX: "0"}}}])
    out["parked_verdict_finalises_when_pods_are_gone"] = obs(
        job(status=dict(RUNNING, phase="Terminating"), annotations=dict(PORTS, Succeed="job job completed")), [])
    out["clean_pod_policy_none_keeps_pods"] = obs(
        job(cleanPodPolicy="None", status=RUNNING, annotations=PORTS),
        [pod("trainer", 0, phase="Succeeded", exit_codes=[0]),
         pod("trainer", 1, phase="Succeeded", exit_codes=[0], node="gpu-1")])
    out["time_limit_expired"] = obs(
        job(timeLimit=60, status=dict(RUNNING, startRunningTime="2026-09-21T11:00:00Z"), annotations=PORTS),
        [pod("trainer", 0), pod("trainer", 1, node="gpu-1")])
    out["preempted_from_outside"] = obs(
        job(status=RUNNING, annotations=dict(PORTS, Preempted="capacity needed elsewhere")),
        [pod("trainer", 0), pod("trainer", 1, node="gpu-1")])
    four = dict(RUNNING, rendezvous={"generation": 1, "worldSizes": {"trainer": 4}, "masterPort": 40001})
    ports4 = {"aitj.b200/host-ports": json.dumps({f"trainer/{i}/2222": 41000 + i for i in range(4)}, sort_keys=True)}
    out["scale_down_marks_surplus_draining"] = obs(
        job(roles={"trainer": {"replicas": 2, "minReplicas": 1, "maxReplicas": 4, "edlPolicy": "Manual"}},
            status=four, annotations=ports4),
        [pod("trainer", i, node=f"gpu-{i}") for i in range(4)], spare=(40002,))
    out["scale_down_deletes_drained_surplus"] = obs(
        job(roles={"trainer": {"replicas": 2, "minReplicas": 1, "maxReplicas": 4, "edlPolicy": "Manual"}},
            status=dict(RUNNING, rendezvous={"generation": 2, "worldSizes": {"trainer": 2}, "masterPort": 40002}),
            annotations=ports4),
        [pod("trainer", 0), pod("trainer", 1, node="gpu-1"),
         pod("trainer", 2, phase="Succeeded", exit_codes=[0], node="gpu-2", annotations={C.ANN_SCALE_DOWN: NOW}),
         pod("trainer", 3, node="gpu-3", annotations={C.ANN_SCALE_DOWN: "2026-09-21T11:59:50Z"})])
    out["scale_up_new_generation"] = obs(
        job(roles={"trainer": {"replicas": 4, "maxReplicas": 8, "edlPolicy": "Manual"}}, status=RUNNING,
            annotations=PORTS),
        [pod("trainer", 0), pod("trainer", 1, node="gpu-1")], spare=(40002, 41002, 41003))
    auto = {"trainer": {"replicas": 2, "minReplicas": 1, "maxReplicas": 4, "edlPolicy": "Auto"}}
    out["auto_grows_into_free_gpus"] = obs(
        job(roles=auto, status=RUNNING, annotations=PORTS), [pod("trainer", 0), pod("trainer", 1, node="gpu-1")],
        cluster={"free_gpu_slots": 2, "waiting_higher": 0, "waiting_at_least": 0})
    out["auto_yields_to_more_important_work"] = obs(
        job(roles={"trainer": dict(auto["trainer"], replicas=4, minReplicas=3)},
            status=dict(RUNNING, rendezvous={"generation": 1, "worldSizes": {"trainer": 4}, "masterPort": 40001}),
            annotations=ports4),
        [pod("trainer", i, node=f"gpu-{i}") for i in range(4)],
        cluster={"free_gpu_slots": 0, "waiting_higher": 2, "waiting_at_least": 2})
    out["start_error_outlives_window"] = obs(
        job(roles={"trainer": {"replicas": 1}},
            status={"phase": "Creating", "conditions": [{"type": "Creating", "status": "True",
                                                         "reason": "TrainingJobCreating", "message": "",
                                                         "lastProbeTime": "2026-09-21T11:00:00Z",
                                                         "lastTransitionTime": "2026-09-21T11:00:00Z"}],
                    "replicaStatuses": {}, "RestartReplicaName": "", "startTime": NOW,
                    "rendezvous": {"generation": 1, "worldSizes": {"trainer": 1}, "masterPort": 40001}},
            annotations={"aitj.b200/host-ports": json.dumps({"trainer/0/2222": 41000})}),
        [pod("trainer", 0, phase="Pending", waiting="ErrImagePull")],
        window={"restart_period": 0.0, "duration_period": 900.0, "fail_after_window": True})
    return out


def main():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "engine_cases.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    doc = {}
    for name, o in cases().items():
        o = json.loads(json.dumps(o, sort_keys=True))     # the order the file will be read back in (role iteration)
        doc[name] = {"observation": o, "decision": replay.replay({"observation": copy.deepcopy(o)})}
    with open(path, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print(f"wrote {len(doc)} cases to {path}")


if __name__ == "__main__":
    main()
