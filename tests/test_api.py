"""API layer: types, JSON spellings, defaults, validation."""
import json
import os

import pytest
import yaml

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api import meta as M
from trainingjob_operator_b200.api import register as R
from trainingjob_operator_b200.api.defaults import set_defaults_aitrainingjob
from trainingjob_operator_b200.api.types import AITrainingJob, AITrainingJobList, ReplicaStatus, TrainingJobStatus
from trainingjob_operator_b200.api.validation import parse_exit_codes, validate_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "examples", "paddle-mnist.yaml")


def example():
    return yaml.safe_load(open(EXAMPLE))


def test_reference_example_parses_unchanged_and_round_trips():
    d = example()
    job = AITrainingJob.from_dict(d)
    assert job.name == "paddle-mnist"
    assert job.spec.restarting_exit_code == "137,128"
    t = job.spec.replica_specs["trainer"]
    assert (t.replicas, t.restart_limit, t.restart_policy, t.fail_policy, t.complete_policy) == \
        (1, 1, "OnNodeFailWithExitCode", "Rank0", "All")
    out = job.to_dict()
    assert out["spec"]["replicaSpecs"]["trainer"]["template"] == d["spec"]["replicaSpecs"]["trainer"]["template"]
    assert out["spec"]["cleanPodPolicy"] == "All"
    assert validate_dict(d) == []


def test_group_version_kind_and_crd_name():
    assert (C.GROUP_NAME, C.GROUP_VERSION, C.KIND, C.KIND_PLURAL, C.SHORT_NAME) == \
        ("elasticdeeplearning.ai", "v1", "AITrainingJob", "aitrainingjobs", "aitj")
    assert C.crd_name() == "aitrainingjobs.elasticdeeplearning.ai"
    crd = R.crd_object()
    assert crd["spec"]["scope"] == "Namespaced" and crd["spec"]["names"]["shortNames"] == ["aitj"]
    assert R.lookup("aitj") is R.AITRAININGJOB and R.lookup("AITrainingJob") is R.AITRAININGJOB
    assert R.AITRAININGJOB.path("ns1", "j") == "/apis/elasticdeeplearning.ai/v1/namespaces/ns1/aitrainingjobs/j"
    assert R.POD.path("default") == "/api/v1/namespaces/default/pods"


def test_status_json_spellings_are_kept():
    st = TrainingJobStatus(phase=C.PHASE_SUCCEEDED)
    st.restart_counts["trainer"] = 2
    st.replica_statuses["trainer"] = ReplicaStatus(active=3)
    d = st.to_dict()
    assert d["phase"] == "Succeed"                      # sic (types.go:111)
    assert d["RestartCount"] == {"trainer": 2}          # capitalised key (types.go:84)
    assert "RestartReplicaName" in d                    # always emitted (types.go:86)
    assert d["replicaStatuses"]["trainer"] == {"active": 3}   # zero counters vanish (omitempty)
    assert "startTime" not in d and "endTime" not in d
    back = TrainingJobStatus.from_dict(d)
    assert back.restart_counts == {"trainer": 2} and back.replica_statuses["trainer"].active == 3


def test_phase_reason_tables():
    assert C.TRAINING_JOB_REASON[C.PHASE_SUCCEEDED] == "TrainingJobSucceed"
    assert C.TRAINING_JOB_REASON[C.PHASE_NODE_FAIL] == "TrainingJobNodeFail"
    assert set(C.ENDING_PHASES) == {"Succeed", "Failed", "Timeout", "Preempted", "NodeFail"}
    assert "ImagePullBackOff" in C.ERROR_CONTAINER_STATUS and len(C.ERROR_CONTAINER_STATUS) == 8
    assert C.DEFAULT_CONTAINER_PREFIX == "aitj-" and C.DEFAULT_PORT_PREFIX == "aitj-"


def test_defaults_match_reference():
    job = AITrainingJob.from_dict({"metadata": {"name": "j"}, "spec": {"replicaSpecs": {"w": {
        "template": {"spec": {"containers": [{"name": "aitj-w", "command": ["true"]}]}}}}}})
    set_defaults_aitrainingjob(job)
    assert job.spec.clean_pod_policy == "All" and job.spec.fail_policy == "Any" and job.spec.complete_policy == "All"
    w = job.spec.replica_specs["w"]
    assert (w.replicas, w.restart_policy, w.restart_scope, w.fail_policy, w.complete_policy) == \
        (1, "Never", "All", "Any", "All")
    assert w.restart_limit is None  # nil = unlimited (pod.go:215-216)
    # explicit values survive
    job2 = AITrainingJob.from_dict({"metadata": {"name": "j"}, "spec": {"cleanPodPolicy": "None", "replicaSpecs": {
        "w": {"replicas": 0, "restartScope": "Pod", "template": {}}}}})
    set_defaults_aitrainingjob(job2)
    assert job2.spec.clean_pod_policy == "None" and job2.spec.replica_specs["w"].replicas == 0
    assert job2.spec.replica_specs["w"].restart_scope == "Pod"


@pytest.mark.parametrize("mutate,needle", [
    (lambda d: d["spec"].__setitem__("replicaSpecs", {}), "replicaSpecs must not be empty"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"]["template"]["spec"].__setitem__("containers", []),
     "at least one container"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"]["template"]["spec"]["containers"][0].update(
        {"image": "", "command": [], "args": []}), "needs an image or a command"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].__setitem__("restartPolicy", "Sometimes"), "restartPolicy"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].__setitem__("restartScope", "Node"), "restartScope"),
    (lambda d: d["spec"].__setitem__("failPolicy", "Most"), "failPolicy"),
    (lambda d: d["spec"].__setitem__("cleanPodPolicy", "Some"), "cleanPodPolicy"),
    (lambda d: d["spec"].__setitem__("restartingExitCode", "137,abc"), "restartingExitCode"),
    (lambda d: d["spec"].__setitem__("timeLimit", -5), "timeLimit"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].update({"minReplicas": 2, "replicas": 1}), "minReplicas"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].update({"maxReplicas": 2, "replicas": 3}), "maxReplicas"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].update({"minReplicas": 4, "maxReplicas": 2, "replicas": 3}),
     "minReplicas"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].__setitem__("edlPolicy", "Maybe"), "edlPolicy"),
    (lambda d: d["metadata"].__setitem__("name", "Bad_Name"), "metadata.name"),
    (lambda d: d["spec"]["replicaSpecs"]["trainer"].__setitem__("replicas", -1), "non-negative"),
])
def test_validation_rejects(mutate, needle):
    d = example()
    mutate(d)
    errs = validate_dict(d)
    assert errs and any(needle in e for e in errs), errs


def test_parse_exit_codes():
    assert parse_exit_codes("137,128") == [137, 128]
    assert parse_exit_codes("") == [] and parse_exit_codes(" 1 , 2 ") == [1, 2]
    with pytest.raises(ValueError):
        parse_exit_codes("x")


def test_deepcopy_is_independent():
    job = AITrainingJob.from_dict(example())
    cp = job.deepcopy()
    cp.spec.replica_specs["trainer"].template["spec"]["containers"][0]["name"] = "changed"
    cp.status.phase = "Running"
    assert job.spec.replica_specs["trainer"].containers()[0]["name"] == "aitj-trainer" and job.status.phase == ""


def test_list_type_and_unknown_fields_survive():
    d = example()
    d["spec"]["futureField"] = {"x": 1}
    job = AITrainingJob.from_dict(d)
    assert job.to_dict()["spec"]["futureField"] == {"x": 1}
    lst = AITrainingJobList.from_dict({"items": [d], "metadata": {"resourceVersion": "5"}})
    assert lst.kind == "AITrainingJobList" and lst.items[0].name == "paddle-mnist"


def test_meta_helpers():
    assert M.parse_selector("a=b, c==d") == {"a": "b", "c": "d"}
    assert M.selector_matches({"a": "b"}, {"a": "b", "x": "y"}) and not M.selector_matches({"a": "b"}, {})
    t = M.format_time()
    assert M.parse_time(t) is not None and t.endswith("Z")
    assert M.split_key("ns/name") == ("ns", "name") and M.split_key("name") == ("", "name")
    job = {"apiVersion": C.API_VERSION, "kind": C.KIND, "metadata": {"name": "j", "uid": "u1"}}
    ref = M.owner_reference(job)
    assert ref["controller"] and ref["blockOwnerDeletion"] and ref["uid"] == "u1" and ref["kind"] == C.KIND
    assert M.get_controller_of({"metadata": {"ownerReferences": [ref]}})["name"] == "j"


@pytest.mark.parametrize("path", sorted(os.listdir(os.path.join(ROOT, "examples"))))
def test_every_example_manifest_validates_and_round_trips(path):
    d = yaml.safe_load(open(os.path.join(ROOT, "examples", path)))
    assert validate_dict(d) == []
    job = AITrainingJob.from_dict(d)
    set_defaults_aitrainingjob(job)
    again = AITrainingJob.from_dict(json.loads(json.dumps(job.to_dict())))
    assert again.to_dict() == job.to_dict()
    for rt, spec in job.spec.replica_specs.items():
        names = [c["name"] for c in spec.template["spec"]["containers"]]
        assert any(n.startswith(C.DEFAULT_CONTAINER_PREFIX) for n in names), (path, rt)   # exit codes only count for aitj-*
