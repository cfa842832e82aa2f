"""Client machinery: fake clientset, informers/listers, event recorder, leader election (C8f, C9, C10, upstream)."""
import threading
import time

import pytest

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.api.types import AITrainingJob
from trainingjob_operator_b200.client.clientset import new_for_config
from trainingjob_operator_b200.client.fake import new_simple_clientset
from trainingjob_operator_b200.client.informers import SharedInformerFactory
from trainingjob_operator_b200.client.leaderelection import LEADER_ANNOTATION, LeaderElectionConfig, LeaderElector
from trainingjob_operator_b200.client.record import EventRecorder, FakeRecorder, events_for
from trainingjob_operator_b200.store.apiserver import APIError, APIServer

from test_api import example


def wait_until(fn, timeout=5.0):
    deadline = time.time() + timeout
    while time.time() < deadline:
        if fn():
            return True
        time.sleep(0.01)
    return False


def test_fake_clientset_tracks_actions_and_reactors():
    seed = example()
    seed["metadata"]["namespace"] = "default"
    cs = new_simple_clientset(seed)
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    assert jobs.get("paddle-mnist").name == "paddle-mnist"
    j = jobs.get("paddle-mnist")
    j.status.phase = "Running"
    jobs.update(j)
    verbs = [(a.verb, a.resource) for a in cs.actions()]
    assert verbs == [("get", "aitrainingjobs"), ("get", "aitrainingjobs"), ("update", "aitrainingjobs")]
    cs.prepend_reactor("update", "aitrainingjobs", lambda a: (True, APIError(500, "InternalError", "boom")))
    with pytest.raises(APIError):
        jobs.update(jobs.get("paddle-mnist"))
    cs.prepend_reactor("create", "pods", lambda a: (True, {"metadata": {"name": "intercepted"}}))
    assert cs.core_v1().pods("default").create({"metadata": {"name": "p"}})["metadata"]["name"] == "intercepted"
    assert cs.core_v1().pods("default").list()["items"] == []     # the reactor swallowed the create


def test_informer_cache_handlers_lister_and_resync():
    api = APIServer()
    cs = new_for_config(server=api)
    jobs = cs.elasticdeeplearning_v1().aitrainingjobs("default")
    jobs.create(AITrainingJob.from_dict(example()))           # exists before the informer starts
    stop = threading.Event()
    factory = SharedInformerFactory(cs, default_resync=0.2)
    typed = factory.elasticdeeplearning().v1().aitrainingjobs()
    events = []
    typed.informer().add_event_handler(add=lambda o: events.append(("add", o["metadata"]["name"])),
                                       update=lambda a, b: events.append(("update", a["metadata"]["resourceVersion"],
                                                                          b["metadata"]["resourceVersion"])),
                                       delete=lambda o: events.append(("delete",)))
    assert factory.elasticdeeplearning().v1().aitrainingjobs().informer() is typed.informer()   # shared
    factory.start(stop)
    assert factory.wait_for_cache_sync(stop) == {"AITrainingJob": True}
    lister = typed.lister()
    assert lister.aitrainingjobs("default").get("paddle-mnist").name == "paddle-mnist"
    with pytest.raises(APIError):
        lister.aitrainingjobs("default").get("nope")
    # listers hand out copies: mutating one must not touch the cache (fixes reference quirk Q6)
    a = lister.aitrainingjobs("default").get("paddle-mnist")
    a.status.phase = "Hacked"
    assert lister.aitrainingjobs("default").get("paddle-mnist").status.phase == ""
    j = jobs.get("paddle-mnist")
    j.status.phase = "Pending"
    jobs.update(j)
    assert wait_until(lambda: lister.aitrainingjobs("default").get("paddle-mnist").status.phase == "Pending")
    assert wait_until(lambda: any(e[0] == "update" and e[1] == e[2] for e in events))   # resync: same rv twice
    jobs.delete("paddle-mnist")
    assert wait_until(lambda: ("delete",) in events)
    assert events[0] == ("add", "paddle-mnist")
    assert lister.list() == []
    stop.set()


def test_informer_namespace_scope_and_selector():
    api = APIServer()
    cs = new_for_config(server=api)
    for ns, name, lbl in (("a", "p1", {"k": "1"}), ("a", "p2", {"k": "2"}), ("b", "p3", {"k": "1"})):
        cs.core_v1().pods(ns).create({"metadata": {"name": name, "labels": lbl}, "spec": {}})
    stop = threading.Event()
    f = SharedInformerFactory(cs, 0, namespace="a")
    pods = f.core().v1().pods()
    pods.informer()
    f.start(stop)
    f.wait_for_cache_sync(stop)
    assert sorted(p["metadata"]["name"] for p in pods.lister().list()) == ["p1", "p2"]
    assert [p["metadata"]["name"] for p in pods.lister().namespaced("a").list({"k": "2"})] == ["p2"]
    stop.set()


def test_event_recorder_writes_and_aggregates_events():
    api = APIServer()
    cs = new_for_config(server=api)
    job = cs.elasticdeeplearning_v1().aitrainingjobs("default").create(AITrainingJob.from_dict(example()))
    rec = EventRecorder(cs, C.CONTROLLER_NAME, log=False)
    rec.eventf(job, "Normal", "SuccessfulCreatePod", "Created pod: %s", "paddle-mnist-trainer-0")
    rec.eventf(job, "Warning", "FailedCreatePod", "Error creating: %s", "x")
    rec.eventf(job, "Warning", "FailedCreatePod", "Error creating: %s", "x")
    assert rec.flush(5.0)                 # recording is asynchronous (a sink thread writes the Events API)
    evs = events_for(cs, job)
    assert [(e["reason"], e["count"]) for e in evs] == [("SuccessfulCreatePod", 1), ("FailedCreatePod", 2)]
    assert evs[0]["source"]["component"] == "TrainingJobOperator" and evs[0]["involvedObject"]["uid"] == job.uid
    fr = FakeRecorder()
    fr.eventf(job, "Normal", "R", "m %d", 1)
    assert fr.events == ["Normal R m 1"]


def _elector(cs, ident, started, stopped, lease=0.6, renew=0.4, retry=0.1):
    cfg = LeaderElectionConfig(identity=ident, lease_duration=lease, renew_deadline=renew, retry_period=retry)
    return LeaderElector(cs, cfg, lambda stop: (started.append(ident), stop.wait()), lambda: stopped.append(ident))


def test_leader_election_single_leader_then_failover():
    api = APIServer()
    cs = new_for_config(server=api)
    started, stopped = [], []
    stop_a, stop_b = threading.Event(), threading.Event()
    a = _elector(cs, "a", started, stopped)
    b = _elector(cs, "b", started, stopped)
    ta = threading.Thread(target=a.run, args=(stop_a,), daemon=True)
    ta.start()
    assert wait_until(lambda: started == ["a"])
    tb = threading.Thread(target=b.run, args=(stop_b,), daemon=True)
    tb.start()
    time.sleep(0.5)
    assert started == ["a"] and b.get_leader() == "a"           # standby observes the holder, does not lead
    lock = cs.core_v1().endpoints("kube-system").get("trainingjob-operator")
    assert '"holderIdentity": "a"' in lock["metadata"]["annotations"][LEADER_ANNOTATION]
    # kill the leader without releasing (crash): b must take over after the lease expires
    a._client = lambda: (_ for _ in ()).throw(APIError(503, "ServiceUnavailable", "partitioned"))
    assert wait_until(lambda: "a" in stopped, 5)
    assert wait_until(lambda: started == ["a", "b"], 5)
    rec = cs.core_v1().endpoints("kube-system").get("trainingjob-operator")["metadata"]["annotations"][LEADER_ANNOTATION]
    assert '"holderIdentity": "b"' in rec and '"leaderTransitions": 1' in rec
    stop_a.set(); stop_b.set()
    tb.join(3)
    assert "b" in stopped


def test_leader_election_release_on_clean_stop_and_config_validation():
    api = APIServer()
    cs = new_for_config(server=api)
    started, stopped = [], []
    stop = threading.Event()
    a = _elector(cs, "a", started, stopped)
    t = threading.Thread(target=a.run, args=(stop,), daemon=True)
    t.start()
    assert wait_until(lambda: started == ["a"])
    stop.set()
    t.join(3)
    b = _elector(cs, "b", started, stopped, lease=30, renew=20, retry=0.05)
    assert b.try_acquire_or_renew()                              # released lock is taken at once
    with pytest.raises(ValueError):
        LeaderElector(cs, LeaderElectionConfig(lease_duration=1, renew_deadline=2, retry_period=0.1), None, None)


def test_leases_lock_type_also_works():
    cs = new_for_config(server=APIServer())
    cfg = LeaderElectionConfig(lock_type="leases", identity="x", lease_duration=1, renew_deadline=0.5, retry_period=0.1)
    e = LeaderElector(cs, cfg, lambda s: None, lambda: None)
    assert e.try_acquire_or_renew() and e.try_acquire_or_renew()
    assert cs.coordination_v1().leases("kube-system").get("trainingjob-operator")


def test_indexer_secondary_indices_and_updates_never_hide_an_object_from_lock_free_readers():
    from trainingjob_operator_b200.client.informers import Indexer

    def mk(ns, name, job, rv):
        return {"metadata": {"namespace": ns, "name": name, "labels": {"j": job}, "resourceVersion": str(rv)}}

    ix = Indexer({"job": lambda o: [f'{o["metadata"]["namespace"]}/{o["metadata"]["labels"]["j"]}']})
    for o in (mk("a", "p1", "x", 1), mk("a", "p2", "x", 1), mk("b", "p3", "x", 1)):
        ix.add(o)
    assert sorted(o["metadata"]["name"] for o in ix.by_index("job", "a/x")) == ["p1", "p2"]
    assert [o["metadata"]["name"] for o in ix.list("b")] == ["p3"] and len(ix.list()) == 3
    ix.add(mk("a", "p1", "y", 2))                                   # label changed: moves between index buckets
    assert [o["metadata"]["name"] for o in ix.by_index("job", "a/x")] == ["p2"]
    assert ix.by_index("job", "a/y")[0]["metadata"]["resourceVersion"] == "2"
    ix.delete(mk("a", "p2", "x", 1))
    assert ix.by_index("job", "a/x") == [] and ix.get_by_key("a/p2") is None
    ix.add_indexers({"ns": lambda o: [o["metadata"]["namespace"]]})  # added late: existing objects are indexed
    assert len(ix.by_index("ns", "a")) == 1 and len(ix.by_index("ns", "b")) == 1
    ix.replace([mk("c", "q", "z", 1)])
    assert ix.list("a") == [] and len(ix) == 1 and len(ix.by_index("job", "c/z")) == 1

    # readers take no lock: while a writer keeps updating one object, a reader must see it in every view, every time
    ix.replace([mk("a", "hot", "x", 0)])
    stop, missing = threading.Event(), []

    def writer():
        rv = 0
        while not stop.is_set():
            rv += 1
            ix.add(mk("a", "hot", "x", rv))

    t = threading.Thread(target=writer, daemon=True)
    t.start()
    deadline = time.time() + 0.5
    while time.time() < deadline:
        if not ix.by_index("job", "a/x") or not ix.list("a") or ix.get_by_key("a/hot") is None:
            missing.append(1)
    stop.set()
    t.join(2)
    assert not missing


def test_http_transport_does_not_repeat_a_write_that_may_have_been_applied():
    """A read time-out after the request was sent: GET is retried, POST / PUT / PATCH / DELETE surface the failure (a
    create that landed would otherwise come back as AlreadyExists, a guarded PUT as Conflict)."""
    import http.server
    import threading
    import time as _time

    import pytest as _pytest

    from trainingjob_operator_b200.api import register as R
    from trainingjob_operator_b200.store.apiserver import APIError
    from trainingjob_operator_b200.store.transport import HTTPTransport

    seen = []

    class Slow(http.server.BaseHTTPRequestHandler):
        def _serve(self):
            n = int(self.headers.get("Content-Length") or 0)
            if n:
                self.rfile.read(n)
            seen.append(self.command)
            _time.sleep(0.6)                       # "applied", but the answer comes too late

        do_GET = do_POST = do_PUT = do_DELETE = do_PATCH = _serve

        def log_message(self, *a):
            pass

    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), Slow)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        tr = HTTPTransport(f"http://127.0.0.1:{srv.server_address[1]}", timeout=0.2)
        with _pytest.raises(APIError) as ei:
            tr.create(R.POD, "default", {"metadata": {"name": "p"}})
        assert ei.value.reason == "Timeout" and seen == ["POST"]           # sent once, not twice
        with _pytest.raises(APIError) as ei:
            tr.get(R.POD, "default", "p")
        assert ei.value.reason == "ServiceUnavailable" and seen.count("GET") == 2
        with _pytest.raises(APIError):
            tr.delete(R.POD, "default", "p")
        assert seen.count("DELETE") == 1
    finally:
        srv.shutdown()


def test_client_side_throttle_is_a_token_bucket_and_never_touches_watches():
    """--kube-api-qps / --kube-api-burst (client-go's default 5 / 10 is what the reference runs with): the first `burst`
    requests pass immediately, the rest are spaced 1/qps apart; watches are long-lived streams and are not throttled."""
    import time as _time

    from trainingjob_operator_b200.api import register as R
    from trainingjob_operator_b200.store.transport import ThrottledTransport, Transport

    calls = []

    class Null(Transport):
        def get(self, info, namespace, name):
            calls.append(("get", _time.monotonic()))
            return {}

        def watch(self, *a, **kw):
            calls.append(("watch", _time.monotonic()))
            return iter(())

    t = ThrottledTransport(Null(), qps=50.0, burst=5)
    t0 = _time.monotonic()
    for _ in range(10):
        t.get(R.POD, "default", "x")
    spent = _time.monotonic() - t0
    assert 0.08 <= spent < 0.5                      # 5 free, then 5 x 20 ms
    assert t.requests == 10 and t.waited_s > 0.05
    t1 = _time.monotonic()
    for _ in range(20):
        t.watch(R.POD)
    assert _time.monotonic() - t1 < 0.05 and t.requests == 10


def test_reference_throttle_flags_parse_and_reach_the_clientsets():
    import argparse

    from trainingjob_operator_b200.cmd import options as O
    from trainingjob_operator_b200.cmd.server import create_client_sets
    from trainingjob_operator_b200.store.apiserver import APIServer
    from trainingjob_operator_b200.store.transport import ThrottledTransport

    p = argparse.ArgumentParser()
    O.add_flags(p)
    ns = p.parse_args(["--kube-api-qps", "5", "--kube-api-burst", "10", "--live-node-list", "--queue-qps", "10",
                       "--queue-burst", "100"])
    assert (ns.kube_api_qps, ns.kube_api_burst, ns.live_node_list) == (5.0, 10, True)
    opt = O.TrainingJobOperatorOption(kube_api_qps=5.0, kube_api_burst=10)
    kube, *_ = create_client_sets(opt, server=APIServer(""))
    assert isinstance(kube.transport, ThrottledTransport) and kube.transport.burst == 10
    plain, *_ = create_client_sets(O.TrainingJobOperatorOption(), server=APIServer(""))
    assert not isinstance(plain.transport, ThrottledTransport)
