"""API server semantics, in process and over HTTP (SURVEY.md §4 component tier, C8 contract)."""
import json
import threading
import time
import urllib.request

import pytest
import yaml

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api import register as R
from trainingjob_operator_b200.api.types import AITrainingJob
from trainingjob_operator_b200.client.clientset import new_for_config
from trainingjob_operator_b200.store.apiserver import APIError, APIServer, json_patch, merge_patch
from trainingjob_operator_b200.store.http import APIHTTPServer

from test_api import example


@pytest.fixture(params=["local", "http"])
def cs(request):
    api = APIServer()
    if request.param == "local":
        yield new_for_config(server=api)
        return
    srv = APIHTTPServer(api).start()
    try:
        yield new_for_config(master=srv.url)
    finally:
        srv.stop()


def pod(name, labels=None, owner=None, node="", phase=None):
    p = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": name, "labels": labels or {}},
         "spec": {"containers": [{"name": "aitj-c", "command": ["true"]}]}}
    if node:
        p["spec"]["nodeName"] = node
    if phase:
        p["status"] = {"phase": phase}
    if owner:
        p["metadata"]["ownerReferences"] = [owner]
    return p


def test_crud_resource_version_and_optimistic_concurrency(cs):
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    created = jobs.create(AITrainingJob.from_dict(example()))
    assert created.uid and created.resource_version and created.metadata["creationTimestamp"]
    assert created.namespace == "default" and created.metadata["generation"] == 1
    got = jobs.get("paddle-mnist")
    got.status.phase = "Pending"
    updated = jobs.update(got)
    assert int(updated.resource_version) > int(created.resource_version)
    assert updated.metadata["generation"] == 1            # status-only change keeps generation
    with pytest.raises(APIError) as ei:
        jobs.update(got)                                   # stale resourceVersion
    assert ei.value.reason == "Conflict" and ei.value.code == 409
    updated.spec.replica_specs["trainer"].replicas = 3
    assert jobs.update(updated).metadata["generation"] == 2
    with pytest.raises(APIError) as ei:
        jobs.create(AITrainingJob.from_dict(example()))
    assert ei.value.reason == "AlreadyExists"
    lst = jobs.list()
    assert [j.name for j in lst.items] == ["paddle-mnist"] and lst.metadata["resourceVersion"]
    jobs.delete("paddle-mnist")
    with pytest.raises(APIError) as ei:
        jobs.get("paddle-mnist")
    assert ei.value.reason == "NotFound" and ei.value.code == 404


def test_update_status_only_touches_status(cs):
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    j = jobs.create(AITrainingJob.from_dict(example()))
    j.status.phase = "Running"
    j.spec.replica_specs["trainer"].replicas = 7        # must be ignored by the status subresource
    out = jobs.update_status(j)
    assert out.status.phase == "Running" and out.spec.replica_specs["trainer"].replicas == 1


def test_admission_rejects_invalid_jobs(cs):
    bad = example()
    bad["spec"]["replicaSpecs"]["trainer"]["restartPolicy"] = "Bogus"
    with pytest.raises(APIError) as ei:
        cs.elasticdeeplearning_v1().aitrainingjobs("default").create(AITrainingJob.from_dict(bad))
    assert ei.value.code == 422 and ei.value.reason == "Invalid" and "restartPolicy" in ei.value.message


def test_patch_merge_and_json(cs):
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    jobs.create(AITrainingJob.from_dict(example()))
    p = jobs.patch("paddle-mnist", {"metadata": {"annotations": {"Preempted": "by test"}},
                                    "spec": {"replicaSpecs": {"trainer": {"replicas": 4}}}})
    assert p.annotations["Preempted"] == "by test" and p.spec.replica_specs["trainer"].replicas == 4
    assert p.spec.replica_specs["trainer"].restart_policy == "OnNodeFailWithExitCode"   # untouched siblings survive
    p = jobs.patch("paddle-mnist", {"metadata": {"annotations": {"Preempted": None}}})
    assert "Preempted" not in p.annotations
    p = jobs.patch("paddle-mnist", [{"op": "replace", "path": "/spec/replicaSpecs/trainer/replicas", "value": 2}],
                   "application/json-patch+json")
    assert p.spec.replica_specs["trainer"].replicas == 2
    assert merge_patch({"a": {"b": 1, "c": 2}}, {"a": {"b": None, "d": 3}}) == {"a": {"c": 2, "d": 3}}
    assert json_patch({"l": [1, 2]}, [{"op": "add", "path": "/l/-", "value": 3}, {"op": "remove", "path": "/l/0"}]) \
        == {"l": [2, 3]}


def test_label_and_field_selectors(cs):
    pods = cs.core_v1().pods("default")
    pods.create(pod("a", {"TrainingJobName": "j1", "role": "x"}, node="gpu-0"))
    pods.create(pod("b", {"TrainingJobName": "j1"}))
    pods.create(pod("c", {"TrainingJobName": "j2"}))
    assert [p["metadata"]["name"] for p in pods.list("TrainingJobName=j1")["items"]] == ["a", "b"]
    assert [p["metadata"]["name"] for p in pods.list("TrainingJobName=j1,role=x")["items"]] == ["a"]
    assert [p["metadata"]["name"] for p in pods.list(field_selector="spec.nodeName=gpu-0")["items"]] == ["a"]
    assert len(cs.core_v1().pods("").list()["items"]) == 3
    pods.delete_collection("TrainingJobName=j1")
    left = {p["metadata"]["name"]: p for p in pods.list()["items"]}
    assert sorted(left) == ["a", "c"]                     # "a" is bound + live: graceful, waits for the agent
    assert left["a"]["metadata"]["deletionTimestamp"] and "deletionTimestamp" not in left["c"]["metadata"]


def test_owner_reference_cascade(cs):
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    j = jobs.create(AITrainingJob.from_dict(example()))
    ref = {"apiVersion": C.API_VERSION, "kind": C.KIND, "name": j.name, "uid": j.uid, "controller": True}
    cs.core_v1().pods("default").create(pod("p0", owner=ref))
    cs.core_v1().services("default").create({"kind": "Service", "metadata": {"name": "s0", "ownerReferences": [ref]},
                                             "spec": {"clusterIP": "None"}})
    cs.core_v1().pods("default").create(pod("stranger"))
    jobs.delete(j.name)
    assert [p["metadata"]["name"] for p in cs.core_v1().pods("default").list()["items"]] == ["stranger"]
    assert cs.core_v1().services("default").list()["items"] == []


def test_graceful_pod_delete_needs_agent_confirmation(cs):
    pods = cs.core_v1().pods("default")
    pods.create(pod("run", node="gpu-1", phase="Running"))
    d = pods.delete("run")
    assert d["metadata"]["deletionTimestamp"] and d["metadata"]["deletionGracePeriodSeconds"] == 30
    assert pods.get("run")["metadata"]["deletionTimestamp"]      # still there until the agent confirms
    pods.delete("run", grace_period_seconds=0)
    with pytest.raises(APIError):
        pods.get("run")
    pods.create(pod("done", node="gpu-1", phase="Succeeded"))
    pods.delete("done")                                         # finished pods go at once
    with pytest.raises(APIError):
        pods.get("done")
    pods.create(pod("unbound"))
    pods.delete("unbound")
    with pytest.raises(APIError):
        pods.get("unbound")


def test_watch_stream_from_resource_version(cs):
    pods = cs.core_v1().pods("default")
    pods.create(pod("a"))
    rv = pods.list()["metadata"]["resourceVersion"]
    events = []
    w = pods.watch(resource_version=rv, timeout=5)

    def consume():
        for ev in w:
            events.append((ev["type"], ev["object"]["metadata"]["name"]))
            if len(events) == 3:
                break
        w.close()

    t = threading.Thread(target=consume)
    t.start()
    time.sleep(0.1)
    b = pods.create(pod("b"))
    b["metadata"]["labels"] = {"x": "y"}
    pods.update(b)
    pods.delete("b")
    t.join(5)
    assert events == [("ADDED", "b"), ("MODIFIED", "b"), ("DELETED", "b")]


def test_http_discovery_health_metrics_and_yaml_body():
    api = APIServer()
    srv = APIHTTPServer(api).start()
    try:
        get = lambda p: urllib.request.urlopen(srv.url + p, timeout=5).read().decode()  # noqa: E731
        assert get("/healthz") == "ok"
        assert "v1.13.5" in get("/version")
        groups = json.loads(get("/apis"))
        assert any(g["name"] == C.GROUP_NAME for g in groups["groups"])
        res = json.loads(get(f"/apis/{C.GROUP_NAME}/v1"))
        assert res["resources"][0]["name"] == "aitrainingjobs" and res["resources"][0]["shortNames"] == ["aitj"]
        req = urllib.request.Request(srv.url + R.AITRAININGJOB.path("default"), data=yaml.safe_dump(example()).encode(),
                                     headers={"Content-Type": "application/yaml"}, method="POST")
        created = json.loads(urllib.request.urlopen(req, timeout=5).read())
        assert created["metadata"]["name"] == "paddle-mnist"
        assert "aitj_apiserver_objects{kind=\"AITrainingJob\"} 1" in get("/metrics")
        with pytest.raises(urllib.error.HTTPError) as ei:
            urllib.request.urlopen(srv.url + "/apis/nope/v1/things", timeout=5)
        assert ei.value.code == 404
    finally:
        srv.stop()


def test_wal_survives_restart(tmp_path):
    wal = str(tmp_path / "s.wal")
    api = APIServer(wal)
    cs1 = new_for_config(server=api)
    j = cs1.elasticdeeplearning_v1().aitrainingjobs("default").create(AITrainingJob.from_dict(example()))
    del api, cs1
    cs2 = new_for_config(server=APIServer(wal))
    again = cs2.elasticdeeplearning_v1().aitrainingjobs("default").get("paddle-mnist")
    assert again.uid == j.uid and again.resource_version == j.resource_version


def test_housekeeping_expires_old_events_and_compacts_the_wal(tmp_path):
    """Events have a TTL (kube-apiserver --event-ttl) and the write-ahead log is rewritten as a snapshot once it has
    outgrown its budget; both survive a restart of the server."""
    import os

    from trainingjob_operator_b200.api import meta as M
    from trainingjob_operator_b200.api import register as R
    from trainingjob_operator_b200.store.apiserver import APIServer

    wal = str(tmp_path / "store.wal")
    api = APIServer(wal)
    ev = R.by_kind("Event")
    pod = R.by_kind("Pod")
    old = M.format_time(M.now() - __import__("datetime").timedelta(hours=2))
    api.create(ev, "default", {"metadata": {"name": "old"}, "reason": "R", "lastTimestamp": old})
    api.create(ev, "default", {"metadata": {"name": "fresh"}, "reason": "R", "lastTimestamp": M.format_time()})
    p = api.create(pod, "default", {"metadata": {"name": "p"}, "spec": {}})
    for i in range(300):                                         # churn: every update is one more WAL record
        p["metadata"]["annotations"] = {"i": str(i), "pad": "x" * 200}
        p = api.update(pod, "default", "p", p)
    before = os.path.getsize(wal)
    done = api.housekeeping_once(event_ttl_s=3600, wal_max_bytes=16 << 10)
    assert done == {"events_expired": 1, "wal_compacted": 1}
    assert [e["metadata"]["name"] for e in api.list(ev)["items"]] == ["fresh"]
    assert os.path.getsize(wal) < before / 10
    assert api.housekeeping_once(event_ttl_s=3600, wal_max_bytes=16 << 10) == {"events_expired": 0, "wal_compacted": 0}
    again = APIServer(wal)                                        # the compacted log replays to the same state
    assert again.get(pod, "default", "p")["metadata"]["annotations"]["i"] == "299"
    assert [e["metadata"]["name"] for e in again.list(ev)["items"]] == ["fresh"]
    assert int(again.get(pod, "default", "p")["metadata"]["resourceVersion"]) == int(p["metadata"]["resourceVersion"])


def test_lean_header_parser_keeps_http_semantics():
    """store/fasthttp.py replaces the e-mail-package header parsing of http.server / http.client: names are
    case-insensitive, repeats are joined, folded lines continue a value, limits raise the stdlib's exceptions."""
    import http.client
    import io

    from trainingjob_operator_b200.store.fasthttp import read_headers

    raw = (b"Content-Type: application/json\r\ncontent-length:  12 \r\nX-Many: a\r\nx-many: b\r\n"
           b"X-Fold: first\r\n\tsecond part\r\nnot a header line\r\nConnection: Keep-Alive\r\n\r\nBODY")
    fp = io.BytesIO(raw)
    h = read_headers(fp)
    assert fp.read() == b"BODY"                                   # stops right after the blank line
    assert h.get("Content-Length") == "12" and h["CONTENT-TYPE"] == "application/json" and "connection" in h
    assert h.get("x-many") == "a, b" and h.get_all("X-Many") == ["a, b"]
    assert h.get("x-fold") == "first second part"
    assert h.get("missing") is None and h.get("missing", "d") == "d" and h.get_all("missing") is None
    with pytest.raises(http.client.LineTooLong):
        read_headers(io.BytesIO(b"X: " + b"a" * 70000 + b"\r\n\r\n"))
    with pytest.raises(http.client.HTTPException):
        read_headers(io.BytesIO(b"".join(b"H%d: v\r\n" % i for i in range(150)) + b"\r\n"))


def test_http_server_request_line_keep_alive_and_errors_over_a_raw_socket():
    """The handler's own parse_request: keep-alive on one connection (HTTP/1.1 default), `Connection: close` honoured,
    HTTP/1.0 closes by default, a malformed request line gets 400, an over-long header line 431."""
    import socket

    api = APIServer()
    srv = APIHTTPServer(api).start()

    def read_response(f):
        status = f.readline().decode()
        headers = {}
        while True:
            line = f.readline()
            if line in (b"\r\n", b""):
                break
            k, _, v = line.decode().partition(":")
            headers[k.strip().lower()] = v.strip()
        body = f.read(int(headers.get("content-length", 0)))
        return status, headers, body

    try:
        s = socket.create_connection(("127.0.0.1", srv.port), timeout=5)
        f = s.makefile("rb")
        for _ in range(2):                                        # two requests on the same connection
            s.sendall(b"GET /healthz HTTP/1.1\r\nHost: x\r\nACCEPT: */*\r\n\r\n")
            status, _h, body = read_response(f)
            assert status.startswith("HTTP/1.1 200") and body == b"ok"
        s.sendall(b"GET /version HTTP/1.1\r\nHost: x\r\nconnection: CLOSE\r\n\r\n")
        status, _h, _b = read_response(f)
        assert status.startswith("HTTP/1.1 200")
        assert f.read() == b""                                    # the server closed the connection
        s.close()
        s = socket.create_connection(("127.0.0.1", srv.port), timeout=5)
        f = s.makefile("rb")
        s.sendall(b"GET /healthz HTTP/1.0\r\n\r\n")
        status, _h, body = read_response(f)
        assert " 200" in status and body == b"ok" and f.read() == b""      # 1.0 without keep-alive: closed
        s.close()
        for bad, code in ((b"GARBAGE\r\n\r\n", b" 400 "), (b"GET /healthz HTTP/9.9\r\n\r\n", b" 400 "),
                          (b"GET /healthz HTTP/1.1\r\nX: " + b"a" * 70000 + b"\r\n\r\n", b" 431 ")):
            s = socket.create_connection(("127.0.0.1", srv.port), timeout=5)
            s.sendall(bad)
            assert code in s.makefile("rb").readline()
            s.close()
        # a body with a mixed-case Content-Length header is read in full
        s = socket.create_connection(("127.0.0.1", srv.port), timeout=5)
        f = s.makefile("rb")
        body = json.dumps({"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "raw"}, "spec": {}}).encode()
        s.sendall(b"POST /api/v1/namespaces/default/pods HTTP/1.1\r\nHost: x\r\ncOnTeNt-LeNgTh: %d\r\n"
                  b"Content-Type: application/json; charset=utf-8\r\n\r\n" % len(body) + body)
        status, _h, out = read_response(f)
        assert " 201 " in status and json.loads(out)["metadata"]["name"] == "raw"
        s.close()
    finally:
        srv.stop()


def test_lean_header_parser_agrees_with_the_stdlib_parser_on_generated_blocks():
    """Property test: for header blocks made of token names and printable values (the shapes HTTP clients send), the
    lean parser returns what ``http.client.parse_headers`` returns for every name, in any letter case."""
    import http.client
    import io

    from hypothesis import given, settings
    from hypothesis import strategies as st

    from trainingjob_operator_b200.store.fasthttp import read_headers

    name = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-", min_size=1, max_size=12)
    value = st.text(alphabet=st.characters(min_codepoint=0x21, max_codepoint=0x7e), min_size=0, max_size=40)
    pad = st.sampled_from(["", " ", "  ", "\t"])

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.tuples(name, pad, value, pad), min_size=0, max_size=12))
    def check(items):
        raw = "".join(f"{n}:{p1}{v}{p2}\r\n" for n, p1, v, p2 in items).encode("iso-8859-1") + b"\r\nrest"
        ours = read_headers(io.BytesIO(raw))
        ref = http.client.parse_headers(io.BytesIO(raw))
        for n, _p1, _v, _p2 in items:
            vals = ref.get_all(n)
            assert vals is not None
            for probe in (n, n.lower(), n.upper()):
                assert ours.get(probe) == ", ".join(x.strip() for x in vals), (raw, n)
        assert ours.get("x-not-there") is None and len(ours) == len({n.lower() for n, *_ in items})

    check()
