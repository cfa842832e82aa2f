"""Status primitives: the condition history and the per-role replica counters.

These are the parts of the status that are API (they are what ``kubectl describe aitj`` shows), specified by
/root/reference/pkg/controller/status.go:13-99 and :307-380:

* conditions are an append-only history: repeating the newest entry's type + status + reason only refreshes its
  message, anything else flips the newest entry to ``False`` and appends a new ``True`` one; ``status.phase`` mirrors
  the newest entry; once a True ``Succeed`` / ``Failed`` / ``Preempted`` / ``Timeout`` condition exists the history is
  closed (``NodeFail`` is not in that list);
* counters: a Pending pod counts as ``restarting`` once its role has restarted, else ``scheduled`` when bound to a
  node, else ``pending``; Running -> ``active``; Succeeded; Failed / Unknown -> ``failed``.  Replicas draining after
  a scale-down are not part of the job any more and are not counted (quirk Q1).

How a pass derives the next status from these lives in ``controller.engine``; how it is written back (optimistic
concurrency, only controller-owned annotations carried over a conflict) in ``controller.executor``.
"""
from __future__ import annotations

from typing import List

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaStatus, TrainingJobCondition, TrainingJobStatus  # noqa: F401
from .pod import of_role, pod_node, pod_phase

CLOSING_CONDITIONS = (C.PHASE_SUCCEEDED, C.PHASE_FAILED, C.PHASE_PREEMPTED, C.PHASE_TIMEOUT)


def is_job_completed(status: TrainingJobStatus) -> bool:
    return any(c.status == "True" and c.type in CLOSING_CONDITIONS for c in status.conditions)


def update_conditions(job: AITrainingJob, ctype: str, reason: str, message: str, now=None) -> None:
    """Record that the job is in phase ``ctype`` now (no-op once the history is closed)."""
    st = job.status
    if is_job_completed(st):
        return
    st.phase = ctype
    newest = st.conditions[-1] if st.conditions else None
    if newest is not None and (newest.type, newest.status, newest.reason) == (ctype, "True", reason):
        newest.message = message
        return
    if newest is not None:
        newest.status = "False"
    stamp = M.format_time(now)
    st.conditions.append(TrainingJobCondition(type=ctype, status="True", reason=reason, message=message,
                                              last_probe_time=stamp, last_transition_time=stamp))


def enter_phase(job: AITrainingJob, phase: str, message: str, now=None) -> None:
    update_conditions(job, phase, C.TRAINING_JOB_REASON[phase], message, now)


def is_failed_phase(phase: str) -> bool:
    return phase != C.PHASE_SUCCEEDED and phase in C.ENDING_PHASES


def is_draining(pod: dict) -> bool:
    return C.ANN_SCALE_DOWN in M.annotations_of(pod)


# pod phase -> counter, for pods that are not Pending
_COUNTER_OF_POD_PHASE = {C.POD_RUNNING: "active", C.POD_SUCCEEDED: "succeeded", C.POD_FAILED: "failed",
                         C.POD_UNKNOWN: "failed"}


def count_pod(job: AITrainingJob, rtype: str, pod: dict, rs: ReplicaStatus) -> None:
    phase = pod_phase(pod)
    if phase == C.POD_PENDING:
        name = "restarting" if job.status.restart_counts.get(rtype, 0) > 0 else \
            ("scheduled" if pod_node(pod) else "pending")
    else:
        name = _COUNTER_OF_POD_PHASE.get(phase)
    if name:
        setattr(rs, name, getattr(rs, name) + 1)


def count_role(job: AITrainingJob, rtype: str, pods: List[dict]) -> ReplicaStatus:
    rs = ReplicaStatus()
    for pod in of_role(pods, rtype.lower()):
        if not is_draining(pod):
            count_pod(job, rtype, pod, rs)
    return rs


def effective_restart_scope(job: AITrainingJob, rtype: str) -> str:
    """``restartScope`` of a role; an elastic ``faultTolerant`` job replaces only the lost replica (scope Pod) -- the
    survivors keep their state and re-rendezvous with the replacement (types.go:47 is never read by the reference)."""
    spec = job.spec.replica_specs[rtype]
    if job.spec.fault_tolerant and spec.edl_policy in (C.EDL_POLICY_AUTO, C.EDL_POLICY_MANUAL):
        return C.RESTART_SCOPE_POD
    return spec.restart_scope
