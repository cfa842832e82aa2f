"""Native core (C++/pybind11): work queue, expectations, store, supervisor (SURVEY.md §2.2, §4 component tier)."""
import os
import signal
import threading
import time

import pytest

from trainingjob_operator_b200.core import _aitj_core as core


# ---------------------------------------------------------------------------------- work queue
def test_workqueue_dedup_and_requeue_after_done():
    q = core.WorkQueue("t")
    q.add("a"); q.add("a"); q.add("b")
    assert len(q) == 2
    assert q.get(0.1) == "a"
    q.add("a")                      # re-added while processing: must come back after done()
    assert q.get(0.1) == "b"
    assert q.get(0.02) is None
    q.done("a")
    assert q.get(0.1) == "a"
    q.done("a"); q.done("b")
    assert q.get(0.02) is None


def test_workqueue_rate_limit_backoff_and_forget():
    q = core.WorkQueue("t", 0.005, 1000.0, 1000.0, 1000)
    delays = [q.add_rate_limited("x") for _ in range(6)]
    assert delays[0] == pytest.approx(0.005) and delays[5] == pytest.approx(0.005 * 32)
    assert q.num_requeues("x") == 6
    q.forget("x")
    assert q.num_requeues("x") == 0 and q.add_rate_limited("x") == pytest.approx(0.005)


def test_workqueue_token_bucket_limits_overall_rate():
    q = core.WorkQueue("t", 0.0001, 1000.0, 10.0, 5)   # 10 qps, burst 5
    delays = [q.add_rate_limited(f"k{i}") for i in range(8)]
    assert max(delays[:5]) < 0.01 and delays[7] > 0.15


def test_workqueue_add_after_and_blocking_get():
    q = core.WorkQueue("t")
    t0 = time.time()
    q.add_after("late", 0.15)
    assert q.get(0.05) is None
    assert q.get(1.0) == "late" and 0.1 < time.time() - t0 < 0.6
    got = []
    th = threading.Thread(target=lambda: got.append(q.get(-1.0)))
    th.start()
    time.sleep(0.05)
    q.add("z")
    th.join(2)
    assert got == ["z"]
    q.shutdown()
    assert q.get(-1.0) is None and q.shutting_down()


def test_expectations_accumulate_and_expire():
    e = core.Expectations(0.2)
    assert e.satisfied("k")
    e.raise_expectations("k", 2, 0)
    e.raise_expectations("k", 1, 1)
    assert not e.satisfied("k") and e.peek("k") == (3, 1)
    for _ in range(3):
        e.creation_observed("k")
    assert not e.satisfied("k")
    e.deletion_observed("k")
    assert e.satisfied("k")
    e.expect_creations("j", 5)
    assert not e.satisfied("j")
    time.sleep(0.25)
    assert e.satisfied("j")          # TTL expiry (5 min upstream, 0.2 s here)


# ---------------------------------------------------------------------------------- store
def test_store_versions_conflicts_watch_and_cascade(tmp_path):
    s = core.Store()
    a = s.create("Pod", "ns", "a", "ua", b'{"x":1}', {"app": "t"}, [])
    assert a["rv"] == 1
    with pytest.raises(core.StoreError) as ei:
        s.create("Pod", "ns", "a", "ua2", b"{}", {}, [])
    assert ei.value.reason == "AlreadyExists"
    w = s.watch_open("Pod", "", 0)
    b = s.update("Pod", "ns", "a", "ua", b'{"x":2}', {"app": "t"}, [], a["rv"])
    with pytest.raises(core.StoreError) as ei:
        s.update("Pod", "ns", "a", "ua", b'{"x":3}', {"app": "t"}, [], a["rv"])
    assert ei.value.reason == "Conflict"
    s.create("Pod", "ns", "child", "uc", b"{}", {"app": "u"}, ["ua"])
    s.create("Service", "ns", "svc", "us", b"{}", {}, ["uc"])     # grandchild through another kind
    items, rv = s.list("Pod", "ns", {"app": "t"})
    assert [i["name"] for i in items] == ["a"] and rv == 4
    removed = [r["name"] for r in s.remove("Pod", "ns", "a")]
    assert removed == ["a", "child", "svc"]
    types = []
    while True:
        ev = s.watch_next(w, 0.05)
        if ev is None:
            break
        types.append((ev[0], ev[1]["name"]))
    assert types == [("MODIFIED", "a"), ("ADDED", "child"), ("DELETED", "a"), ("DELETED", "child")]
    with pytest.raises(core.StoreError) as ei:
        s.get("Pod", "ns", "a")
    assert ei.value.reason == "NotFound"


def test_store_cascade_follows_the_owner_index_through_updates_and_a_restart(tmp_path):
    """Dependents are found through an owner-uid index (not a sweep over every object): it has to follow adoption
    (owner added by an update), release (owner removed), deletion of a dependent, and a replay of the write-ahead log."""
    wal = str(tmp_path / "s.wal")
    s = core.Store(wal)
    s.create("AITrainingJob", "ns", "j1", "uj1", b"{}", {}, [])
    s.create("AITrainingJob", "ns", "j2", "uj2", b"{}", {}, [])
    s.create("Pod", "ns", "orphan", "u0", b"{}", {}, [])
    s.create("Pod", "ns", "p1", "u1", b"{}", {}, ["uj1"])
    s.create("Pod", "ns", "p2", "u2", b"{}", {}, ["uj1"])
    s.create("Pod", "ns", "gone", "u3", b"{}", {}, ["uj1"])
    s.create("Event", "ns", "e1", "ue", b"{}", {}, [])
    s.update("Pod", "ns", "orphan", "u0", b"{}", {}, ["uj1"], 0)      # adopted by j1
    s.update("Pod", "ns", "p2", "u2", b"{}", {}, ["uj2"], 0)          # handed from j1 to j2
    s.remove("Pod", "ns", "gone")                                      # a dependent that is deleted on its own
    del s
    s = core.Store(wal)                                                # the index is rebuilt from the log
    removed = sorted(r["name"] for r in s.remove("AITrainingJob", "ns", "j1"))
    assert removed == ["j1", "orphan", "p1"]
    assert s.count("Pod") == 1 and s.count("Event") == 1
    s.update("Pod", "ns", "p2", "u2", b"{}", {}, [], 0)               # released: no owner any more
    assert [r["name"] for r in s.remove("AITrainingJob", "ns", "j2")] == ["j2"]
    assert s.get("Pod", "ns", "p2")["owner_uids"] == []
    # the same name re-created with another owner is not tied to the old one
    s.create("AITrainingJob", "ns", "j1", "uj1b", b"{}", {}, [])
    s.create("Pod", "ns", "p1", "u1b", b"{}", {}, ["uj1b"])
    assert sorted(r["name"] for r in s.remove("AITrainingJob", "ns", "j1")) == ["j1", "p1"]


def test_store_watch_replays_from_resource_version():
    s = core.Store()
    s.create("Pod", "ns", "a", "u1", b"{}", {}, [])
    s.create("Pod", "ns", "b", "u2", b"{}", {}, [])
    s.create("Pod", "other", "c", "u3", b"{}", {}, [])
    w = s.watch_open("Pod", "ns", 1)
    ev = s.watch_next(w, 0.1)
    assert ev[0] == "ADDED" and ev[1]["name"] == "b"
    assert s.watch_next(w, 0.05) is None     # namespace filter drops "c"
    s.watch_close(w)
    assert s.watch_next(w, 0.01) is None


def test_store_wal_replay_and_compaction(tmp_path):
    wal = str(tmp_path / "store.wal")
    s = core.Store(wal)
    s.create("AITrainingJob", "default", "j", "uj", b'{"spec": {"a": "multi word\\nvalue"}}', {"k": "v w"}, [])
    s.create("Pod", "default", "p", "up", b"{}", {}, ["uj"])
    s.update("AITrainingJob", "default", "j", "uj", b'{"spec": 2}', {"k": "v"}, [], 0)
    s.remove("Pod", "default", "p")
    del s
    s2 = core.Store(wal)
    assert s2.get("AITrainingJob", "default", "j")["data"] == b'{"spec": 2}'
    assert s2.current_rv() == 4 and s2.count("Pod") == 0
    s2.compact()
    s2.create("Pod", "default", "q", "uq", b"{}", {}, [])
    del s2
    s3 = core.Store(wal)
    assert s3.count("Pod") == 1 and s3.get("AITrainingJob", "default", "j")["labels"] == {"k": "v"}


# ---------------------------------------------------------------------------------- supervisor
ENV = {"PATH": os.environ.get("PATH", "/usr/bin:/bin")}


def _pgid(pid):
    try:
        return os.getpgid(pid)
    except OSError:
        return -1


def _wait_exit(sup, timeout=5.0):
    deadline = time.time() + timeout
    while time.time() < deadline:
        evs = sup.poll_exits(0.2)
        if evs:
            return evs
    return []


def test_supervisor_exit_codes_env_and_logs(tmp_path):
    sup = core.Supervisor()
    log = str(tmp_path / "out.log")
    env = dict(ENV, FOO="bar")
    sup.spawn("ok", ["/bin/sh", "-c", "echo hello $FOO; exit 0"], env, "", log, "", [])
    evs = _wait_exit(sup)
    assert evs[0]["id"] == "ok" and evs[0]["exit_code"] == 0 and evs[0]["signal"] == 0
    assert open(log).read().strip() == "hello bar"
    sup.spawn("bad", ["/bin/sh", "-c", "exit 3"], ENV)
    assert _wait_exit(sup)[0]["exit_code"] == 3
    assert not sup.alive("bad")


def test_supervisor_sigkill_maps_to_137_and_kills_the_group():
    sup = core.Supervisor()
    sup.spawn("grp", ["/bin/sh", "-c", "sleep 30 & sleep 30"], ENV)
    pid = sup.pid_of("grp")
    time.sleep(0.1)
    assert sup.alive("grp") and sup.kill("grp", signal.SIGKILL, True)
    ev = _wait_exit(sup)[0]
    assert ev["exit_code"] == 137 and ev["signal"] == 9       # matches restartingExitCode: 137,128
    time.sleep(0.1)
    import psutil

    live = [p for p in psutil.process_iter(["pid", "status"])
            if p.info["status"] != psutil.STATUS_ZOMBIE and _pgid(p.info["pid"]) == pid]
    assert not live                                            # the whole process group is gone


def test_supervisor_spawn_errors_are_synchronous(tmp_path):
    sup = core.Supervisor()
    with pytest.raises(OSError) as ei:
        sup.spawn("nx", ["/nonexistent/binary"], ENV)
    assert ei.value.errno == 2
    with pytest.raises(OSError):
        sup.spawn("cwd", ["/bin/true"], ENV, "/nonexistent/dir")
    sup.spawn("dup", ["/bin/sleep", "5"], ENV)
    with pytest.raises(OSError):
        sup.spawn("dup", ["/bin/sleep", "5"], ENV)
    sup.kill("dup", signal.SIGKILL, True)


def test_store_wal_survives_a_torn_tail_and_a_second_crash(tmp_path):
    """kill -9 in the middle of an append leaves half a record at the end of the log.  Replay stops there -- and cuts the
    file back, so that what is written afterwards is still found by the *next* replay."""
    from trainingjob_operator_b200.core import _aitj_core as core

    wal = str(tmp_path / "store.wal")
    s = core.Store(wal)
    s.create("Pod", "default", "a", "uid-a", b'{"metadata":{"name":"a"}}', {"k": "v"}, [])
    s.create("Pod", "default", "b", "uid-b", b'{"metadata":{"name":"b"}}', {}, [])
    del s
    size = os.path.getsize(wal)
    with open(wal, "ab") as f:
        f.write(b'P 3 3 Pod7 default1 c5 uid-c400 {"metadata":{"na')           # torn: 400 bytes promised, 17 written
    s2 = core.Store(wal)
    assert s2.count("Pod") == 2 and s2.current_rv() == 2
    assert os.path.getsize(wal) == size                                         # torn bytes removed
    s2.create("Pod", "default", "c", "uid-c", b'{"metadata":{"name":"c"}}', {}, [])
    del s2
    s3 = core.Store(wal)                                                        # second start: nothing hidden
    assert s3.count("Pod") == 3 and s3.get("Pod", "default", "c")["uid"] == "uid-c"
    assert s3.get("Pod", "default", "a")["labels"] == {"k": "v"}


def test_store_restart_makes_older_watches_gone_and_keeps_the_version_counter(tmp_path):
    """After a restart the event history is empty: a watch from a resourceVersion older than the log's end must be told
    to re-list (Gone) instead of silently missing what happened in between; and a compaction whose newest operations were
    deletes must not let resourceVersions be handed out twice."""
    wal = str(tmp_path / "store.wal")
    s = core.Store(wal)
    s.create("Pod", "ns", "a", "u1", b"{}", {}, [])          # rv 1
    s.create("Pod", "ns", "b", "u2", b"{}", {}, [])          # rv 2
    s.remove("Pod", "ns", "b")                               # rv 3
    del s
    s2 = core.Store(wal)
    assert s2.current_rv() == 3
    with pytest.raises(core.StoreError) as ei:
        s2.watch_open("Pod", "ns", 1)                        # events 2, 3 are not replayable any more
    assert ei.value.reason == "Gone"
    w = s2.watch_open("Pod", "ns", 3)                        # from "now": fine
    s2.create("Pod", "ns", "c", "u3", b"{}", {}, [])         # rv 4
    assert s2.watch_next(w, 0.2)[1]["name"] == "c"
    w2 = s2.watch_open("Pod", "ns", 3)                       # history of this incarnation replays
    assert s2.watch_next(w2, 0.2)[1]["name"] == "c"
    s2.remove("Pod", "ns", "c")                              # rv 5: the newest operation is a delete
    s2.compact()
    del s2
    s3 = core.Store(wal)
    assert s3.current_rv() == 5                              # not 1 (the highest rv among the live objects)
    assert s3.create("Pod", "ns", "d", "u4", b"{}", {}, [])["rv"] == 6


def test_store_wal_cut_right_after_a_records_last_digit(tmp_path):
    """A crash can end the log on the final count digit of a record (its separator and newline missing): the record is
    complete, and whatever is appended next must not be glued onto it."""
    wal = str(tmp_path / "store.wal")
    s = core.Store(wal)
    s.create("Pod", "default", "a", "uid-a", b"{}", {}, [])
    del s
    raw = open(wal, "rb").read()
    assert raw.endswith(b" 0 \n")
    open(wal, "wb").write(raw[:-2])                          # "... 0" -- cut after the owner count
    s2 = core.Store(wal)
    assert s2.count("Pod") == 1
    s2.create("Pod", "default", "b", "uid-b", b"{}", {}, [])
    del s2
    s3 = core.Store(wal)
    assert s3.count("Pod") == 2 and s3.current_rv() == 2


def test_native_tree_copy_equals_deepcopy_on_random_json_trees():
    """``core.jcopy`` (behind ``meta.deepcopy`` and the typed conversions): equal to ``copy.deepcopy`` on arbitrary JSON-shaped
    trees, shares no container with its input, leaves non-JSON members to the fallback, and refuses cyclic input."""
    import copy

    from hypothesis import given, settings
    from hypothesis import strategies as st

    leaves = st.one_of(st.none(), st.booleans(), st.integers(), st.floats(allow_nan=False), st.text(max_size=8),
                       st.binary(max_size=4))
    trees = st.recursive(leaves, lambda kids: st.one_of(st.lists(kids, max_size=4),
                                                        st.dictionaries(st.text(max_size=5), kids, max_size=4)), max_leaves=25)

    def containers(x, out):
        if isinstance(x, (dict, list)):
            out.append(id(x))
            for v in (x.values() if isinstance(x, dict) else x):
                containers(v, out)
        return out

    @settings(max_examples=300, deadline=None)
    @given(trees)
    def check(tree):
        c = core.jcopy(tree, copy.deepcopy)
        assert c == tree and type(c) is type(tree)
        assert not (set(containers(tree, [])) & set(containers(c, [])))

    check()
    odd = {"t": (1, [2]), "s": {3}, "d": {"k": [1, {"z": None}]}}
    c = core.jcopy(odd, copy.deepcopy)
    assert c == odd and c["t"] is not odd["t"] and c["s"] is not odd["s"] and c["d"]["k"] is not odd["d"]["k"]
    loop = []
    loop.append(loop)
    with pytest.raises(RecursionError):
        core.jcopy(loop, copy.deepcopy)
