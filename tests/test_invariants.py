"""Property tests (hypothesis) of the status engine: whatever pod states are thrown at it, counters never
exceed what exists, terminal phases are absorbing, and restart counts respect the limit."""

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from trainingjob_operator_b200.api import constants as C
from trainingjob_operator_b200.controller import status as S

from test_controller_unit import Harness, job_dict

POD_STATES = st.sampled_from(["absent", "pending", "scheduled", "running", "succeeded", "failed137", "failed1",
                              "creating_err"])


def apply_state(h, job, role, idx, state):
    name = f"job-{role}-{idx}"
    h.pods_idx.delete({"metadata": {"name": name, "namespace": "default"}})
    if state == "absent":
        return
    kw = {"pending": dict(phase="Pending", node=""), "scheduled": dict(phase="Pending", node=f"gpu-{idx % 4}"),
          "running": dict(phase="Running", node=f"gpu-{idx % 4}"),
          "succeeded": dict(phase="Succeeded", exit_codes=[0], node=f"gpu-{idx % 4}"),
          "failed137": dict(phase="Failed", exit_codes=[137], node=f"gpu-{idx % 4}"),
          "failed1": dict(phase="Failed", exit_codes=[1], node=f"gpu-{idx % 4}"),
          "creating_err": dict(phase="Pending", waiting="ErrImagePull", node=f"gpu-{idx % 4}")}[state]
    h.pod(job, role, idx, **kw)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(policy=st.sampled_from(list(C.RESTART_POLICIES)), scope=st.sampled_from(list(C.RESTART_SCOPES)),
       limit=st.one_of(st.none(), st.integers(0, 2)), fail=st.sampled_from(list(C.ENDING_POLICIES)),
       complete=st.sampled_from(list(C.ENDING_POLICIES)),
       rounds=st.lists(st.lists(POD_STATES, min_size=3, max_size=3), min_size=1, max_size=6))
def test_state_machine_invariants(policy, scope, limit, fail, complete, rounds):
    h = Harness()
    role = {"replicas": 3, "restartPolicy": policy, "restartScope": scope, "failPolicy": fail,
            "completePolicy": complete}
    if limit is not None:
        role["restartLimit"] = limit
    job = h.add_job(job_dict(roles={"trainer": role}, restartingExitCode="137"))
    terminal_seen = None
    for states in rounds:
        for i, stt in enumerate(states):
            apply_state(h, job, "trainer", i, stt)
        job = h.sync()
        st_ = job.status
        rs = st_.replica_statuses.get("trainer")
        if rs is not None:
            assert rs.total() <= 3                                  # never counts more than exist
            assert min(rs.pending, rs.scheduled, rs.active, rs.succeeded, rs.restarting, rs.failed) >= 0
        if limit is not None:
            assert st_.restart_counts.get("trainer", 0) <= limit    # restartCount <= restartLimit
        if st_.conditions:
            assert st_.phase == st_.conditions[-1].type             # phase mirrors the newest condition
            assert [c.status for c in st_.conditions].count("True") == 1
        if terminal_seen is not None:
            assert st_.phase == terminal_seen                       # terminal phases are absorbing
        elif st_.phase in (C.PHASE_SUCCEEDED, C.PHASE_FAILED, C.PHASE_TIMEOUT, C.PHASE_PREEMPTED):
            terminal_seen = st_.phase
        # whatever the controller deleted disappears from the cache before the next round
        for name in list(h.pc.deleted):
            h.pods_idx.delete({"metadata": {"name": name, "namespace": "default"}})
        h.pc.clear()
