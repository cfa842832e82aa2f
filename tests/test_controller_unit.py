"""Table-driven unit tests of the reconcile state machine against fake pod/service controls
(SURVEY.md §4 unit tier; every behaviour of §2.8).  No threads, no processes: informer caches are
filled by hand, exactly how sibling operators test with FakePodControl."""
import json
import time

import pytest

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api import meta as M
from trainingjob_operator_b200.api import register as R
from trainingjob_operator_b200.api.types import AITrainingJob
from trainingjob_operator_b200.client.fake import new_simple_clientset
from trainingjob_operator_b200.client.informers import SharedInformerFactory
from trainingjob_operator_b200.client.record import FakeRecorder
from trainingjob_operator_b200.cmd.options import TrainingJobOperatorOption
from trainingjob_operator_b200.controller import status as S
from trainingjob_operator_b200.controller.control import FakePodControl, FakeServiceControl
from trainingjob_operator_b200.controller.controller import TrainingJobController
from trainingjob_operator_b200.controller.pod import gen_expectation_pods_key, gen_general_name, \
    is_retryable_exit_code


def job_dict(name="job", roles=None, **spec):
    roles = roles or {"trainer": {}}
    rs = {}
    for role, over in roles.items():
        r = {"replicas": 2, "template": {"spec": {"containers": [
            {"name": f"aitj-{role}", "image": "img", "command": ["run"],
             "ports": [{"name": "aitj-port", "containerPort": 2222}]}]}}}
        r.update(over)
        rs[role] = r
    d = {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": name, "namespace": "default"},
         "spec": dict({"replicaSpecs": rs}, **spec)}
    return d


class Harness:
    def __init__(self, **opt):
        self.option = TrainingJobOperatorOption(**opt)
        self.cs = new_simple_clientset()
        self.kube_factory = SharedInformerFactory(self.cs, 0)
        self.tj_factory = SharedInformerFactory(self.cs, 0)
        self.pc, self.sc = FakePodControl(), FakeServiceControl()
        self.tc = TrainingJobController(self.cs, self.cs, self.cs, self.kube_factory, self.tj_factory, self.option,
                                        pod_control=self.pc, service_control=self.sc, recorder=FakeRecorder())
        self.jobs_idx = self.tj_factory.informer_for(R.AITRAININGJOB).indexer
        self.pods_idx = self.kube_factory.informer_for(R.POD).indexer
        self.svc_idx = self.kube_factory.informer_for(R.SERVICE).indexer
        self.node_idx = self.kube_factory.informer_for(R.NODE).indexer
        self.set_nodes(["gpu-0", "gpu-1", "gpu-2", "gpu-3", "cpu-0"])

    def set_nodes(self, ready, not_ready=()):
        for k in self.node_idx.keys():
            self.node_idx.delete({"metadata": {"name": k}})
        for n in ready:
            self.node_idx.add({"kind": "Node", "metadata": {"name": n},
                               "status": {"conditions": [{"type": "Ready", "status": "True"}]}})
        for n in not_ready:
            self.node_idx.add({"kind": "Node", "metadata": {"name": n},
                               "status": {"conditions": [{"type": "Ready", "status": "False"}]}})

    def add_job(self, d) -> AITrainingJob:
        created = self.cs.tracker.create(R.AITRAININGJOB, "default", d)
        self.jobs_idx.add(created)
        return AITrainingJob.from_dict(created)

    def refresh_job(self, name="job") -> AITrainingJob:
        cur = self.cs.tracker.get(R.AITRAININGJOB, "default", name)
        self.jobs_idx.add(cur)
        return AITrainingJob.from_dict(cur)

    def pod(self, job, role, index, phase=C.POD_RUNNING, node="gpu-0", exit_codes=None, waiting=None,
            restart_count=0, deleting=False, annotations=None, start_time=None, unschedulable=None):
        rt = role.lower()
        name = gen_general_name(job.name, rt, str(index))
        css = []
        cname = f"aitj-{rt}"
        if exit_codes is not None:
            css = [{"name": cname, "state": {"terminated": {"exitCode": c, "reason": "Error" if c else "Completed"}}}
                   for c in exit_codes]
        elif waiting is not None:
            css = [{"name": cname, "state": {"waiting": {"reason": waiting, "message": "boom"}}}]
        elif phase == C.POD_RUNNING:
            css = [{"name": cname, "state": {"running": {}}}]
        p = {"apiVersion": "v1", "kind": "Pod",
             "metadata": {"name": name, "namespace": "default", "uid": f"uid-{name}-{time.time_ns()}",
                          "resourceVersion": str(time.time_ns() % 10 ** 9),
                          "creationTimestamp": M.format_time(),
                          "labels": {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: job.name,
                                     C.LABEL_REPLICA_NAME: rt, C.LABEL_REPLICA_INDEX: str(index),
                                     C.LABEL_RESTART_COUNT: str(restart_count)},
                          "ownerReferences": [M.owner_reference(job.to_dict())]},
             "spec": {"nodeName": node} if node else {},
             "status": {"phase": phase, "containerStatuses": css}}
        if start_time:
            p["status"]["startTime"] = start_time
        if unschedulable:
            p["status"]["conditions"] = [{"type": "PodScheduled", "status": "False", "message": unschedulable}]
        if deleting:
            p["metadata"]["deletionTimestamp"] = M.format_time()
        if annotations:
            p["metadata"]["annotations"] = annotations
        self.pods_idx.add(p)
        return p

    def clear_pods(self):
        for k in self.pods_idx.keys():
            ns, n = k.split("/")
            self.pods_idx.delete({"metadata": {"name": n, "namespace": ns}})

    def settle(self, name="job"):
        """Pretend every create/delete of the last pass has been observed by the informers."""
        from trainingjob_operator_b200.controller.service import gen_expectation_services_key

        job = AITrainingJob.from_dict(self.cs.tracker.get(R.AITRAININGJOB, "default", name))
        for rt in job.spec.replica_specs:
            self.tc.expectations.delete(gen_expectation_pods_key(f"default/{name}", rt))
            self.tc.expectations.delete(gen_expectation_services_key(f"default/{name}", rt))

    def sync(self, name="job", settle=True) -> AITrainingJob:
        self.tc.sync_handler(f"default/{name}")
        if settle:
            self.settle(name)
        return self.refresh_job(name)


@pytest.fixture
def h():
    return Harness()


# =========================================================================== creation / contract
def test_first_reconcile_creates_pods_services_and_goes_pending(h):
    h.add_job(job_dict(priority="high", schedulerName="aitj-scheduler"))
    job = h.sync()
    assert sorted(t["metadata"]["name"] for t in h.pc.templates) == ["job-trainer-0", "job-trainer-1"]
    assert sorted(s["metadata"]["name"] for s in h.sc.services) == ["job-trainer-0", "job-trainer-1"]
    t0 = next(t for t in h.pc.templates if t["metadata"]["name"] == "job-trainer-0")
    lbl = t0["metadata"]["labels"]
    assert lbl[C.LABEL_GROUP_NAME] == "elasticdeeplearning.ai" and lbl[C.LABEL_JOB_NAME] == "job"
    assert lbl["JobName"] == "job" and lbl["PodRole"] == "trainer" and lbl["RestartCount"] == "0"
    assert lbl[C.LABEL_REPLICA_NAME] == "trainer" and lbl[C.LABEL_REPLICA_INDEX] == "0" and lbl["priority"] == "high"
    assert t0["metadata"]["generateName"] == "job-trainer-"
    assert t0["spec"]["restartPolicy"] == "Never" and t0["spec"]["schedulerName"] == "aitj-scheduler"
    ref = h.pc.controller_refs[0]
    assert ref == {"apiVersion": "elasticdeeplearning.ai/v1", "kind": "AITrainingJob", "name": "job",
                   "uid": job.uid, "blockOwnerDeletion": True, "controller": True}
    svc = h.sc.services[0]
    assert svc["spec"]["clusterIP"] == "None" and svc["spec"]["ports"][0]["name"] == "aitj-2222"
    assert svc["spec"]["selector"][C.LABEL_REPLICA_INDEX] == svc["metadata"]["labels"][C.LABEL_REPLICA_INDEX]
    assert job.status.phase == C.PHASE_PENDING
    c = job.status.conditions[-1]
    assert (c.type, c.status, c.reason, c.message) == ("Pending", "True", "TrainingJobPending",
                                                       "all pods are waiting for scheduling")
    assert job.status.start_time and job.spec.clean_pod_policy == "All"     # defaults persisted with the status write
    assert job.spec.replica_specs["trainer"].restart_policy == "Never"
    assert job.status.rendezvous.generation == 1 and job.status.rendezvous.world_sizes == {"trainer": 2}


def test_env_contract_golden(h):
    h.add_job(job_dict(roles={"trainer": {"replicas": 2}, "PServer": {"replicas": 1}}, frameworkType="pytorch"))
    h.sync()
    tpl = next(t for t in h.pc.templates if t["metadata"]["name"] == "job-trainer-1")
    env = {e["name"]: e["value"] for e in tpl["spec"]["containers"][0]["env"]}
    # the 13 reference variables (pod.go:548-652)
    assert env["TRAINER_INSTANCES"] == "job-trainer-0.default,job-trainer-1.default"
    assert env["TRAINER_INSTANCES_NUM"] == "2"
    assert env["TRAINER_PORTS"] == "2222" and env["TRAINER_PORTS_NUM"] == "1"
    assert env["TRAINER_HOSTS"] == "job-trainer-0.default:2222,job-trainer-1.default:2222"
    assert env["TRAINER_HOSTS_NUM"] == "2"
    assert env["PSERVER_INSTANCES"] == "job-pserver-0.default" and env["PSERVER_HOSTS_NUM"] == "1"
    assert env["TRAININGJOB_REPLICA_NAME"] == "trainer" and env["TRAININGJOB_REPLICA_INDEX"] == "1"
    assert env["TRAININGJOB_REPLICA_RESTARTCOUNT"] == "0"
    assert env["TRAININGJOB_SERVICE"] == "job-trainer-1.default"
    assert env["TRAININGJOB_NAME"] == "job" and env["TRAININGJOB_NAMESPACE"] == "default"
    assert env["TRAININGJOB_PORTS"] == "2222"
    # torch / elastic dialect
    assert env["RANK"] == "1" and env["WORLD_SIZE"] == "2" and env["MASTER_ADDR"] == "127.0.0.1"
    assert int(env["MASTER_PORT"]) > 0 and env["AITJ_RENDEZVOUS_GENERATION"] == "1"
    assert env["TRAINER_ADDRS"].count("127.0.0.1:") == 2 and env["AITJ_HOST_PORTS"]
    # a container that is not aitj- gets the contract but no ports
    d = job_dict(name="j2")
    d["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"].append({"name": "sidecar", "command": ["x"]})
    d["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["initContainers"] = [{"name": "init", "command": ["y"]}]
    h.add_job(d)
    h.pc.clear()
    h.sync("j2")
    tpl = h.pc.templates[0]
    side = {e["name"]: e["value"] for e in tpl["spec"]["containers"][1]["env"]}
    init = {e["name"]: e["value"] for e in tpl["spec"]["initContainers"][0]["env"]}
    assert side["TRAININGJOB_PORTS"] == "" and side["TRAININGJOB_NAME"] == "j2"
    assert "TRAININGJOB_PORTS" not in init and init["TRAININGJOB_REPLICA_INDEX"] in ("0", "1")


def test_services_only_for_roles_with_aitj_containers(h):
    d = job_dict()
    d["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0]["name"] = "plain"
    h.add_job(d)
    h.sync()
    assert h.sc.services == [] and len(h.pc.templates) == 2


def test_naming_and_exit_code_matcher():
    assert gen_general_name("a/b", "trainer", "3") == "a-b-trainer-3"
    assert is_retryable_exit_code([137], "137,128") and is_retryable_exit_code([137, 128], "137,128")
    assert not is_retryable_exit_code([137, 1], "137,128")     # every code must be listed
    assert not is_retryable_exit_code([], "137,128")           # no codes => not retryable
    assert not is_retryable_exit_code([1], "")


# =========================================================================== classifier truth table
CLASSIFIER = [
    # restartPolicy, pod kwargs, node ready?, expected (phase, is_restart)
    ("Never", dict(phase="Failed", exit_codes=[1]), True, ("Failed", False)),
    ("OnFailure", dict(phase="Failed", exit_codes=[1]), True, ("Failed", True)),
    ("Always", dict(phase="Failed", exit_codes=[1]), True, ("Failed", True)),
    ("ExitCode", dict(phase="Failed", exit_codes=[137]), True, ("Failed", True)),
    ("ExitCode", dict(phase="Failed", exit_codes=[1]), True, ("Failed", False)),
    ("OnNodeFailWithExitCode", dict(phase="Failed", exit_codes=[128]), True, ("Failed", True)),
    ("OnNodeFailWithExitCode", dict(phase="Failed", exit_codes=[2]), True, ("Failed", False)),
    ("OnNodeFail", dict(phase="Failed", exit_codes=[137]), True, ("Failed", False)),
    ("OnNodeFail", dict(phase="Running"), False, ("NodeFail", True)),
    ("OnNodeFailWithExitCode", dict(phase="Running"), False, ("NodeFail", True)),
    ("Always", dict(phase="Running"), False, ("NodeFail", True)),
    ("Never", dict(phase="Running"), False, ("NodeFail", False)),
    ("OnFailure", dict(phase="Running"), False, ("NodeFail", False)),
    ("ExitCode", dict(phase="Running"), False, ("NodeFail", False)),
    ("Always", dict(phase="Succeeded", exit_codes=[0]), True, ("Succeed", False)),   # Always never restarts success
    ("Never", dict(phase="Running"), True, ("", False)),
    ("Never", dict(phase="Pending", waiting="ContainerCreating"), True, ("Creating", False)),
    ("Never", dict(phase="Pending", waiting="ImagePullBackOff"), True, ("Creating", False)),
]


@pytest.mark.parametrize("policy,podkw,node_ready,expected", CLASSIFIER)
def test_container_classifier_truth_table(h, policy, podkw, node_ready, expected):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 1, "restartPolicy": policy}},
                             restartingExitCode="137,128"))
    pod = h.pod(job, "trainer", 0, node="gpu-0", **podkw)
    nodes = {"gpu-0": True} if node_ready else {}
    phase, is_restart, msg = h.tc.reconcile_containers(job, pod, "trainer", nodes)
    assert (phase, is_restart) == expected
    if expected[0] == "NodeFail":
        assert msg == "Node gpu-0 is failed and offline"
    if podkw.get("exit_codes") and podkw["exit_codes"][0] != 0:
        assert "exited with reason Error exitcode" in msg


def test_classifier_only_aitj_containers_count(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 1, "restartPolicy": "ExitCode"}}, restartingExitCode="137"))
    pod = h.pod(job, "trainer", 0, phase="Failed", exit_codes=[137])
    pod["status"]["containerStatuses"].append({"name": "sidecar", "state": {"terminated": {"exitCode": 1}}})
    phase, is_restart, _ = h.tc.reconcile_containers(job, pod, "trainer", {"gpu-0": True})
    assert (phase, is_restart) == ("Failed", True)            # the sidecar's exit code 1 is ignored


def test_creating_failed_flags():
    h = Harness(enable_creating_failed=True)
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 1}}))
    S.update_conditions(job, C.PHASE_CREATING, "TrainingJobCreating", "")
    pod = h.pod(job, "trainer", 0, phase="Pending", waiting="ErrImagePull")
    phase, is_restart, msg = h.tc.reconcile_containers(job, pod, "trainer", {"gpu-0": True})
    assert phase == "Failed" and "create container failed[ErrImagePull]" in msg
    # inside the retry window: restart once the pod has been stuck longer than the duration period
    h2 = Harness(creating_restart_time=3600.0, creating_duration_time=1.0)
    job2 = h2.add_job(job_dict(roles={"trainer": {"replicas": 1}}))
    S.update_conditions(job2, C.PHASE_CREATING, "TrainingJobCreating", "")
    old = M.format_time(M.now().replace(year=M.now().year - 1))
    pod2 = h2.pod(job2, "trainer", 0, phase="Pending", waiting="CreateContainerError", start_time=old)
    assert h2.tc.reconcile_containers(job2, pod2, "trainer", {"gpu-0": True})[:2] == ("Creating", True)
    pod3 = h2.pod(job2, "trainer", 0, phase="Pending", waiting="CreateContainerError")   # no startTime: no crash (Q10)
    assert h2.tc.reconcile_containers(job2, pod3, "trainer", {"gpu-0": True})[:2] == ("Creating", False)


# =========================================================================== phases & counters
def test_phase_progression_pending_creating_running_succeed(h):
    job = h.add_job(job_dict())
    job = h.sync()
    assert job.status.phase == "Pending"
    h.pod(job, "trainer", 0, phase="Pending", node="")
    h.pod(job, "trainer", 1, phase="Pending", node="", unschedulable="0/8 nodes are available")
    job = h.sync()
    assert job.status.phase == "Pending" and job.status.replica_statuses["trainer"].pending == 2
    h.pod(job, "trainer", 0, phase="Pending", node="gpu-0", waiting="ContainerCreating")
    h.pod(job, "trainer", 1, phase="Pending", node="gpu-1", waiting="ContainerCreating")
    job = h.sync()
    assert job.status.phase == "Creating" and job.status.replica_statuses["trainer"].scheduled == 2
    assert "creating containers" in job.status.conditions[-1].message
    h.pod(job, "trainer", 0, phase="Running", node="gpu-0")
    h.pod(job, "trainer", 1, phase="Running", node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Running" and job.status.start_running_time
    assert job.status.replica_statuses["trainer"].active == 2
    started = job.status.start_running_time
    h.pod(job, "trainer", 0, phase="Succeeded", exit_codes=[0])
    job = h.sync()
    assert job.status.phase == "Running"                        # completePolicy All: one success is not enough
    assert job.status.replica_statuses["trainer"].succeeded == 1
    h.pod(job, "trainer", 1, phase="Succeeded", exit_codes=[0], node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Terminating" and job.annotations["Succeed"] == "job job completed"
    assert sorted(h.pc.deleted) == ["job-trainer-0", "job-trainer-1"] and len(h.sc.deleted) == 0 or True
    h.clear_pods()
    job = h.sync()
    assert job.status.phase == "Succeed" and job.status.end_time and job.status.start_running_time == started
    last = job.status.conditions[-1]
    assert (last.type, last.status, last.reason) == ("Succeed", "True", "TrainingJobSucceed")
    assert last.message == "job job completed; deleted pods"
    assert [c.status for c in job.status.conditions[:-1]] == ["False"] * (len(job.status.conditions) - 1)
    # terminal: never reconciled again
    h.pc.clear()
    h.sync()
    assert h.pc.templates == [] and h.pc.deleted == []


def test_condition_list_semantics():
    job = AITrainingJob.from_dict(job_dict())
    S.update_conditions(job, "Pending", "TrainingJobPending", "m1")
    S.update_conditions(job, "Pending", "TrainingJobPending", "m2")
    assert len(job.status.conditions) == 1 and job.status.conditions[0].message == "m2"
    S.update_conditions(job, "Running", "TrainingJobRunning", "r")
    assert [(c.type, c.status) for c in job.status.conditions] == [("Pending", "False"), ("Running", "True")]
    assert job.status.phase == "Running"
    S.update_conditions(job, "NodeFail", "TrainingJobNodeFail", "n")       # NodeFail is not final (status.go:33-58)
    S.update_conditions(job, "Failed", "TrainingJobFailed", "f")
    S.update_conditions(job, "Running", "TrainingJobRunning", "again")     # ignored after a final condition
    assert job.status.phase == "Failed" and job.status.conditions[-1].type == "Failed"
    assert S.is_job_completed(job.status)


def test_restarting_counter_after_restart(h):
    job = h.add_job(job_dict())
    job.status.restart_counts = {"trainer": 1}
    rs = S.ReplicaStatus()
    S.count_pod(job, "trainer", {"status": {"phase": "Pending"}, "spec": {"nodeName": "gpu-0"}}, rs)
    S.count_pod(job, "trainer", {"status": {"phase": "Unknown"}}, rs)
    assert rs.restarting == 1 and rs.scheduled == 0 and rs.failed == 1


# =========================================================================== ending policies
@pytest.mark.parametrize("policy,states,expected", [
    ("Any", ["Succeeded", "Running"], "Succeed"),
    ("Rank0", ["Succeeded", "Running"], "Succeed"),
    ("Rank0", ["Running", "Succeeded"], ""),
    ("All", ["Succeeded", "Running"], ""),
    ("All", ["Succeeded", "Succeeded"], "Succeed"),
    ("None", ["Succeeded", "Succeeded"], ""),
])
def test_role_complete_policies(h, policy, states, expected):
    job = h.add_job(job_dict(roles={"trainer": {"completePolicy": policy, "failPolicy": "None"}}))
    from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob
    set_defaults_aitrainingjob(job)
    job.status.rendezvous = None
    pods = [h.pod(job, "trainer", i, phase=s, exit_codes=[0] if s == "Succeeded" else None, node=f"gpu-{i}")
            for i, s in enumerate(states)]
    phase, _ = h.tc.reconcile_pods(job, pods, "trainer")
    assert phase == expected


@pytest.mark.parametrize("policy,states,expected", [
    ("Any", ["Running", "Failed"], "Failed"),
    ("Rank0", ["Running", "Failed"], ""),
    ("Rank0", ["Failed", "Running"], "Failed"),
    ("All", ["Failed", "Running"], ""),
    ("All", ["Failed", "Failed"], "Failed"),
    ("None", ["Failed", "Failed"], ""),
])
def test_role_fail_policies(h, policy, states, expected):
    job = h.add_job(job_dict(roles={"trainer": {"failPolicy": policy, "completePolicy": "None"}}))
    from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob
    set_defaults_aitrainingjob(job)
    pods = [h.pod(job, "trainer", i, phase=s, exit_codes=[1] if s == "Failed" else None, node=f"gpu-{i}")
            for i, s in enumerate(states)]
    phase, msg = h.tc.reconcile_pods(job, pods, "trainer")
    assert phase == expected
    if expected and policy == "All":
        assert msg.startswith("All trainer pods are failed")


def test_node_fail_phase_wins_for_fail_policy_all(h):
    h.set_nodes(["gpu-0"], not_ready=["gpu-1"])
    job = h.add_job(job_dict(roles={"trainer": {"failPolicy": "Any"}}))
    from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob
    set_defaults_aitrainingjob(job)
    pods = [h.pod(job, "trainer", 0, node="gpu-0"), h.pod(job, "trainer", 1, node="gpu-1")]
    phase, msg = h.tc.reconcile_pods(job, pods, "trainer")
    assert phase == "NodeFail" and "Node gpu-1 is failed and offline" in msg


@pytest.mark.parametrize("complete,fail,role_phases,expected", [
    ("All", "Any", {"a": "Succeed", "b": ""}, None),
    ("All", "Any", {"a": "Succeed", "b": "Succeed"}, "Succeed"),
    ("Any", "Any", {"a": "Succeed", "b": "Failed"}, "Succeed"),        # complete beats fail
    ("All", "Any", {"a": "Succeed", "b": "Failed"}, "Failed"),
    ("All", "All", {"a": "Succeed", "b": "Failed"}, None),
    ("All", "All", {"a": "NodeFail", "b": "Failed"}, "Failed"),
    ("None", "None", {"a": "Succeed", "b": "Succeed"}, None),
])
def test_job_level_policies(h, complete, fail, role_phases, expected):
    job = h.add_job(job_dict(roles={"a": {"replicas": 1}, "b": {"replicas": 1}}, completePolicy=complete,
                             failPolicy=fail, cleanPodPolicy="None"))
    from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob
    set_defaults_aitrainingjob(job)
    pods = [h.pod(job, "a", 0), h.pod(job, "b", 0, node="gpu-1")]
    phases = {k: v for k, v in role_phases.items() if v}
    h.tc.update_status(job, pods, [], phases, "msg")
    if expected is None:
        assert job.status.phase not in C.ENDING_PHASES
    else:
        assert job.status.phase == expected
        assert job.status.conditions[-1].message.endswith("; kept pods") and job.status.end_time   # Q8 fixed


# =========================================================================== restart machinery
@pytest.mark.parametrize("scope,expected_deleted,counts", [
    ("Pod", ["job-trainer-1"], {"trainer": 1, "ps": 0}),
    ("Replica", ["job-trainer-0", "job-trainer-1"], {"trainer": 1, "ps": 0}),
    ("All", ["job-ps-0", "job-trainer-0", "job-trainer-1"], {"trainer": 1, "ps": 1}),
])
def test_restart_scopes_and_barrier(h, scope, expected_deleted, counts):
    job = h.add_job(job_dict(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": scope},
                                    "ps": {"replicas": 1}}))
    h.pod(job, "trainer", 0, node="gpu-0")
    h.pod(job, "trainer", 1, phase="Failed", exit_codes=[137], node="gpu-1")
    h.pod(job, "ps", 0, node="cpu-0")
    job = h.sync()
    assert sorted(h.pc.deleted) == expected_deleted
    assert job.status.phase == "Terminating" and job.status.restart_replica_name == "trainer"
    assert {k: job.status.restart_counts.get(k, 0) for k in ("trainer", "ps")} == counts
    assert job.status.conditions[-1].message.startswith("restart times is 1, container aitj-trainer on node gpu-1")
    gen_before = job.status.rendezvous.generation
    assert gen_before == 2                                         # a restart re-rendezvouses on a fresh port
    # barrier: nothing happens while the victims still exist
    h.pc.clear()
    job = h.sync()
    assert job.status.phase == "Terminating" and h.pc.templates == [] and h.pc.deleted == []
    # victims gone -> Restarting, marker cleared
    for name in expected_deleted:
        h.pods_idx.delete({"metadata": {"name": name, "namespace": "default"}})
    job = h.sync()
    assert job.status.phase == "Restarting" and job.status.restart_replica_name == ""
    # next pass re-creates the missing replicas with the new restart count
    job = h.sync()
    assert sorted(t["metadata"]["name"] for t in h.pc.templates) == expected_deleted
    t = next(t for t in h.pc.templates if "trainer" in t["metadata"]["name"])
    env = {e["name"]: e["value"] for e in t["spec"]["containers"][0]["env"]}
    assert env["TRAININGJOB_REPLICA_RESTARTCOUNT"] == "1" and t["metadata"]["labels"]["RestartCount"] == "1"
    assert env["AITJ_RENDEZVOUS_GENERATION"] == str(gen_before)
    assert job.status.phase == "Restarting"                        # sticky until Running
    for name in expected_deleted:
        role, idx = name.split("-")[1], int(name.split("-")[2])
        h.pod(job, role, idx, node=f"gpu-{idx}" if role == "trainer" else "cpu-0", restart_count=1)
    job = h.sync()
    assert job.status.phase == "Running"


def test_fault_tolerant_elastic_job_replaces_only_the_lost_replica(h):
    job = h.add_job(job_dict(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": "All",
                                                "edlPolicy": "Manual", "minReplicas": 1, "maxReplicas": 4}},
                             faultTolerant=True))
    h.pod(job, "trainer", 0, node="gpu-0")
    h.pod(job, "trainer", 1, phase="Failed", exit_codes=[137], node="gpu-1")
    job = h.sync()
    assert h.pc.deleted == ["job-trainer-1"]                       # scope All would have taken trainer-0 down too
    assert job.status.restart_replica_name == "trainer"
    h.pods_idx.delete({"metadata": {"name": "job-trainer-1", "namespace": "default"}})
    job = h.sync()
    assert job.status.phase == "Restarting" and job.status.restart_replica_name == ""   # barrier cleared per Pod scope


def test_restart_limit_exhaustion_falls_through_to_fail_policy(h):
    job = h.add_job(job_dict(roles={"trainer": {"restartPolicy": "OnFailure", "restartScope": "Pod",
                                                "restartLimit": 1}}))
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    live["status"] = {"phase": "Running", "conditions": [], "replicaStatuses": {}, "RestartCount": {"trainer": 1},
                      "RestartReplicaName": ""}
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)
    job = h.refresh_job()
    h.pod(job, "trainer", 0)
    h.pod(job, "trainer", 1, phase="Failed", exit_codes=[1], node="gpu-1")
    job = h.sync()
    assert job.status.restart_counts["trainer"] == 1               # not bumped past the limit
    assert job.status.phase == "Terminating" and "Failed" in job.annotations
    assert job.annotations["Failed"].startswith("pod job-trainer-1 is failed")


def test_node_fail_restart_uses_force_delete(h):
    h.set_nodes(["gpu-0"], not_ready=["gpu-1"])
    job = h.add_job(job_dict(roles={"trainer": {"restartPolicy": "OnNodeFail", "restartScope": "Pod"}}))
    h.pod(job, "trainer", 0, node="gpu-0")
    h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    assert h.pc.force_deleted == ["job-trainer-1"] and job.status.restart_replica_name == "trainer"


# =========================================================================== termination paths
def test_clean_pod_policy_none_keeps_pods(h):
    job = h.add_job(job_dict(cleanPodPolicy="None"))
    h.pod(job, "trainer", 0, phase="Succeeded", exit_codes=[0])
    h.pod(job, "trainer", 1, phase="Succeeded", exit_codes=[0], node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Succeed" and h.pc.deleted == []
    assert job.status.conditions[-1].message == "job job completed; kept pods" and job.status.end_time


def test_timeout_and_preempt_delete_even_with_policy_none(h):
    job = h.add_job(job_dict(cleanPodPolicy="None", timeLimit=1))
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    old = M.format_time(M.now().replace(year=M.now().year - 1))
    live["status"] = {"phase": "Running", "conditions": [], "replicaStatuses": {}, "RestartReplicaName": "",
                      "startRunningTime": old}
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)
    job = h.refresh_job()
    h.pod(job, "trainer", 0); h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Terminating" and "Timeout" in job.annotations
    assert "timeLimit is 1 second" in job.annotations["Timeout"] and len(h.pc.deleted) == 2
    h.clear_pods()
    job = h.sync()
    assert job.status.phase == "Timeout" and job.status.conditions[-1].reason == "TrainingJobTimeout"


def test_time_limit_requeues_for_the_remaining_time(h):
    job = h.add_job(job_dict(timeLimit=3600))
    h.pod(job, "trainer", 0); h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Running"
    assert h.tc.work_queue.len_waiting() == 1                      # AddAfter(timeLimit - elapsed)


@pytest.mark.parametrize("annotation,phase", [("Preempted", "Preempted"), ("Failed", "Failed")])
def test_external_control_annotations(h, annotation, phase):
    d = job_dict()
    d["metadata"]["annotations"] = {annotation: "operator said so"}
    job = h.add_job(d)
    h.pod(job, "trainer", 0); h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Terminating" and len(h.pc.deleted) == 2
    assert "operator said so" in job.status.conditions[-1].message
    h.clear_pods()
    job = h.sync()
    assert job.status.phase == phase and job.status.conditions[-1].message.endswith("; deleted pods")


def test_deleting_jobs_and_unknown_keys_are_skipped(h):
    d = job_dict()
    job = h.add_job(d)
    cached = h.jobs_idx.get_by_key("default/job")
    cached["metadata"]["deletionTimestamp"] = M.format_time()
    h.tc.sync_handler("default/job")
    assert h.pc.templates == []
    assert h.tc.sync_handler("default/ghost") is True             # NotFound => forget
    with pytest.raises(ValueError):
        h.tc.sync_handler("nonamespace")


# =========================================================================== expectations / cache
def test_expectations_gate_the_next_pass(h):
    h.add_job(job_dict())
    h.sync(settle=False)
    assert len(h.pc.templates) == 2
    key = gen_expectation_pods_key("default/job", "trainer")
    assert h.tc.expectations.peek(key) == (2, 0)
    h.sync(settle=False)                                           # creations not observed yet: no second burst
    assert len(h.pc.templates) == 2
    job = h.refresh_job()
    p0 = h.pod(job, "trainer", 0, phase="Pending", node="")
    h.tc.add_pod(p0)
    assert h.tc.expectations.peek(key) == (1, 0)
    h.tc.add_pod(h.pod(job, "trainer", 1, phase="Pending", node=""))
    # services expectations too
    for s in h.sc.services:
        s["metadata"]["namespace"] = "default"
        s["metadata"]["ownerReferences"] = [M.owner_reference(job.to_dict())]
        h.svc_idx.add(s)
        h.tc.add_service(s)
    assert h.tc.satisfied_expectations(job)
    h.tc.delete_pod(p0)
    assert len(h.tc.work_queue) >= 1


def test_stale_cached_job_is_not_acted_on(h):
    job = h.add_job(job_dict())
    stale = h.jobs_idx.get_by_key("default/job")
    h.tc.sync_handler("default/job")                               # writes status (new resourceVersion)
    h.jobs_idx.add(stale)                                          # informer has not caught up yet
    h.pc.clear()
    h.tc.expectations.delete(gen_expectation_pods_key("default/job", "trainer"))
    h.tc.sync_handler("default/job")
    assert h.pc.templates == []                                    # skipped, re-queued
    assert h.tc.work_queue.len_waiting() >= 1


def test_status_write_is_conflict_safe(h):
    """A concurrent spec edit between read and write must survive (reference quirk Q7)."""
    job = h.add_job(job_dict())
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    live["spec"]["replicaSpecs"]["trainer"]["replicas"] = 5          # user edit the cache has not seen
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)
    h.tc.sync_handler("default/job")                               # reconciles the stale copy (replicas=2)
    after = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    assert after["spec"]["replicaSpecs"]["trainer"]["replicas"] == 5  # edit kept
    assert after["status"]["phase"] == "Pending"                   # status landed


# =========================================================================== elastic (new behaviour)
def test_scale_down_drains_then_deletes_out_of_range_replicas():
    h = Harness(scale_down_grace=0.2)
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 4, "minReplicas": 1, "maxReplicas": 4,
                                                "edlPolicy": "Manual"}}))
    for i in range(4):
        h.pod(job, "trainer", i, node=f"gpu-{i}")
    job = h.sync()
    assert job.status.phase == "Running" and job.status.rendezvous.world_sizes == {"trainer": 4}
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    live["spec"]["replicaSpecs"]["trainer"]["replicas"] = 2
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)
    h.refresh_job()
    job = h.sync()
    # indices 2,3 are marked draining, not counted, not yet deleted; the generation moved
    assert sorted(n for n, _ in h.pc.patches) == ["job-trainer-2", "job-trainer-3"]
    assert job.status.replica_statuses["trainer"].active == 2 and job.status.phase == "Running"
    assert h.pc.deleted == []
    assert job.status.rendezvous.generation == 2 and job.status.rendezvous.world_sizes == {"trainer": 2}
    for i in (2, 3):
        h.pod(job, "trainer", i, phase="Succeeded", exit_codes=[0], node=f"gpu-{i}",
              annotations={C.ANN_SCALE_DOWN: M.format_time()})
    job = h.sync()
    assert sorted(h.pc.deleted) == ["job-trainer-2", "job-trainer-3"]
    assert job.status.phase == "Running" and job.status.replica_statuses["trainer"].succeeded == 0


def test_scale_up_bumps_generation_and_new_replicas_get_new_world(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 2, "maxReplicas": 8, "edlPolicy": "Manual"}}))
    h.pod(job, "trainer", 0); h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    assert job.status.phase == "Running" and job.status.rendezvous.generation == 1
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    live["spec"]["replicaSpecs"]["trainer"]["replicas"] = 4
    h.cs.tracker.update(R.AITRAININGJOB, "default", "job", live)
    h.refresh_job()
    h.pc.clear()
    job = h.sync()
    assert sorted(t["metadata"]["name"] for t in h.pc.templates) == ["job-trainer-2", "job-trainer-3"]
    env = {e["name"]: e["value"] for e in h.pc.templates[0]["spec"]["containers"][0]["env"]}
    assert env["WORLD_SIZE"] == "4" and env["AITJ_RENDEZVOUS_GENERATION"] == "2"
    assert job.status.rendezvous.world_sizes == {"trainer": 4}
    assert job.status.phase == "Pending"                           # like the reference: Running -> Pending -> ...


def test_edl_policy_auto_grows_into_free_gpus_and_shrinks_when_unschedulable(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 2, "minReplicas": 1, "maxReplicas": 4,
                                                "edlPolicy": "Auto"}}))
    h.pod(job, "trainer", 0, node="gpu-0"); h.pod(job, "trainer", 1, node="gpu-1")
    h.sync()
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    assert live["spec"]["replicaSpecs"]["trainer"]["replicas"] == 4     # gpu-2, gpu-3 are free
    h.refresh_job()
    h.set_nodes(["gpu-0", "gpu-1", "cpu-0"])
    h.pod(job, "trainer", 2, phase="Pending", node="", unschedulable="0/2 nodes are available")
    h.pod(job, "trainer", 3, phase="Pending", node="", unschedulable="0/2 nodes are available")
    h.sync()
    live = h.cs.tracker.get(R.AITRAININGJOB, "default", "job")
    assert live["spec"]["replicaSpecs"]["trainer"]["replicas"] == 2


def test_adoption_of_orphans_and_release(h):
    job = h.add_job(job_dict())
    orphan = h.pod(job, "trainer", 0)
    orphan["metadata"]["ownerReferences"] = []
    foreign = h.pod(job, "trainer", 1, node="gpu-1")
    foreign["metadata"]["ownerReferences"] = [{"kind": C.KIND, "name": "other", "uid": "someone-else",
                                               "controller": True}]
    sel = {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: "job"}
    claimed = h.tc.claim_pods(job, sel, [orphan, foreign])
    assert [M.name_of(p) for p in claimed] == ["job-trainer-0"]
    assert h.pc.patches and h.pc.patches[0][0] == "job-trainer-0"
    assert h.pc.patches[0][1]["metadata"]["ownerReferences"][0]["uid"] == job.uid
    # owned but labels no longer match -> released
    h.pc.clear()
    mine = h.pod(job, "trainer", 0)
    mine["metadata"]["labels"][C.LABEL_JOB_NAME] = "renamed"
    assert h.tc.claim_pods(job, sel, [mine]) == []
    assert h.pc.patches[0][1]["metadata"]["ownerReferences"] is None


def test_trace_annotation_records_latency_markers(h):
    d = job_dict()
    d["metadata"]["annotations"] = {C.ANN_TRACE: json.dumps({"submitted": time.time()})}
    job = h.add_job(d)
    h.pod(job, "trainer", 0); h.pod(job, "trainer", 1, node="gpu-1")
    job = h.sync()
    tr = json.loads(job.annotations[C.ANN_TRACE])
    assert {"submitted", "firstReconcile", "running"} <= set(tr) and tr["running"] >= tr["submitted"]


def test_worker_reports_on_the_job_become_metrics():
    import json

    from trainingjob_operator_b200.controller.trainingjob import observe_worker_reports
    from trainingjob_operator_b200.utils import metrics

    metrics.reset()
    base = {"metadata": {"namespace": "default", "name": "j", "annotations": {}}}

    r1 = {"metadata": {"namespace": "default", "name": "j", "annotations": {
        "aitj.b200/rescale-trace": json.dumps({"generation": 2, "world": 8, "seconds": 3.5})}}}
    observe_worker_reports(base, r1)
    observe_worker_reports(r1, r1)                                   # unchanged annotation: not observed twice
    r2 = {"metadata": {"namespace": "default", "name": "j", "annotations": {
        "aitj.b200/rescale-trace": json.dumps({"generation": 3, "world": 8, "seconds": 1.7, "recovered_from": "RuntimeError"}),
        "aitj.b200/metrics": json.dumps({"samples_per_sec": 6414.0, "world": 8}),
        "aitj.b200/worker-trace": "not json at all"}}}
    observe_worker_reports(r1, r2)
    text = metrics.render()
    assert 'aitj_rescale_seconds_count{kind="rescale"} 1' in text and 'aitj_rescale_seconds_count{kind="recovery"} 1' in text
    assert 'aitj_job_recoveries_total{job="j",namespace="default"} 1.0' in text
    assert 'aitj_job_samples_per_second{job="j",namespace="default"} 6414.0' in text
    bad = {"metadata": {"namespace": "default", "name": "j", "annotations": {"aitj.b200/rescale-trace": "{broken",
                                                                            "aitj.b200/metrics": "[1]"}}}
    observe_worker_reports(r2, bad)                                  # malformed reports are ignored, not fatal
    metrics.reset()


def _replicas(h):
    return h.cs.tracker.get(R.AITRAININGJOB, "default", "job")["spec"]["replicaSpecs"]["trainer"]["replicas"]


def _foreign_pending(h, name, gpus, priority, job_name="vip"):
    p = {"apiVersion": "v1", "kind": "Pod",
         "metadata": {"name": name, "namespace": "default", "uid": f"uid-{name}",
                      "labels": {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: job_name, C.LABEL_PRIORITY: priority}},
         "spec": {"containers": [{"name": "aitj-x", "resources": {"limits": {"nvidia.com/gpu": gpus}}}]},
         "status": {"phase": "Pending", "conditions": [{"type": "PodScheduled", "status": "False",
                                                        "message": "0/4 nodes are available"}]}}
    h.pods_idx.add(p)
    return p


def test_edl_policy_auto_does_not_repeat_a_shrink_that_is_still_draining(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 3, "minReplicas": 1, "maxReplicas": 4,
                                                "edlPolicy": "Auto"}}))
    h.set_nodes(["gpu-0", "gpu-2", "gpu-3", "cpu-0"])                     # gpu-1 is gone
    h.pod(job, "trainer", 0, node="gpu-0")
    h.pod(job, "trainer", 1, phase="Pending", node="", unschedulable="0/3 nodes are available")
    h.pod(job, "trainer", 2, node="gpu-2")
    h.pod(job, "trainer", 3, node="gpu-3")                                # surplus of the 4 -> 3 shrink, still draining
    h.sync()
    assert _replicas(h) == 3                                              # its slot is all rank 1 needs: no 3 -> 2


def test_edl_policy_auto_yields_to_higher_priority_only_and_never_below_min(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 4, "minReplicas": 3, "maxReplicas": 4,
                                                "edlPolicy": "Auto"}}))
    for i in range(4):
        h.pod(job, "trainer", i, node=f"gpu-{i}")
    _foreign_pending(h, "peer-0", 1, "")                                   # same priority: no reason to yield
    h.sync()
    assert _replicas(h) == 4
    _foreign_pending(h, "vip-0", 2, "high")                                # wants 2, none free
    h.sync()
    assert _replicas(h) == 3                                               # clamped by minReplicas


def test_edl_policy_auto_does_not_grow_into_slots_that_equals_or_betters_wait_for(h):
    job = h.add_job(job_dict(roles={"trainer": {"replicas": 2, "minReplicas": 1, "maxReplicas": 4,
                                                "edlPolicy": "Auto"}}))
    h.pod(job, "trainer", 0, node="gpu-0"); h.pod(job, "trainer", 1, node="gpu-1")
    _foreign_pending(h, "peer-0", 1, "")                                   # found unschedulable a moment ago
    h.sync()
    assert _replicas(h) == 2                                               # gpu-2 / gpu-3 are free, but not for us
    h.pods_idx.delete({"metadata": {"name": "peer-0", "namespace": "default"}})
    h.sync()
    assert _replicas(h) == 4
