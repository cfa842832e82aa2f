"""Status engine: condition history, replica counters, phase derivation, termination, write-back.

Behavioural parity with /root/reference/pkg/controller/status.go:13-380:

* condition list (status.go:60-87): same type+status+reason as the newest entry only refreshes the
  message, otherwise the newest entry flips to ``False`` and a new ``True`` entry is appended; phase
  mirrors the newest entry; once a True Succeed/Failed/Preempted/Timeout condition exists nothing
  else is accepted (status.go:33-58 -- ``NodeFail`` is not in that list);
* counters (status.go:332-359): Pending pod -> ``restarting`` if the role has restarted, else
  ``scheduled`` if bound to a node, else ``pending``; Running -> active; Succeeded; Failed/Unknown;
* restart barrier (status.go:113-143), job-level ending policies with "complete beats fail"
  (status.go:146-174), annotation-driven finalisation (status.go:176-187), time limit
  (status.go:189-198, 246-252), phase derivation by sequential ifs (status.go:200-244),
  terminate (status.go:256-283), status write-back with 5 attempts (status.go:285-305).

Deliberate fixes: Q7 -- the write-back re-reads the *live* object on conflict and
carries only status + annotations + defaulted spec fields over, under optimistic concurrency, instead
of blindly re-PUTting; Q8 -- ``endTime`` is set on the keep-pods path too; replicas that are draining
after a scale-down are not counted (Q1); ``lastReconcileTime`` is written.
"""
from __future__ import annotations

from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaStatus, TrainingJobCondition, TrainingJobStatus
from ..store.apiserver import APIError
from ..utils import klog

TERMINAL_CONDITIONS = (C.PHASE_SUCCEEDED, C.PHASE_FAILED, C.PHASE_PREEMPTED, C.PHASE_TIMEOUT)


# ------------------------------------------------------------------------------ conditions
def new_condition(ctype: str, reason: str, message: str) -> TrainingJobCondition:
    now = M.format_time()
    return TrainingJobCondition(type=ctype, status="True", reason=reason, message=message, last_probe_time=now,
                                last_transition_time=now)


def is_job_completed(status: TrainingJobStatus) -> bool:
    for t in TERMINAL_CONDITIONS:
        c = status.get_condition(t)
        if c is not None and c.status == "True":
            return True
    return False


def set_condition(status: TrainingJobStatus, cond: TrainingJobCondition) -> None:
    if status.conditions:
        cur = status.conditions[-1]
        if cur.type == cond.type and cur.status == cond.status and cur.reason == cond.reason:
            cur.message = cond.message
            return
        cur.status = "False"
    status.conditions.append(cond)


def update_conditions(job: AITrainingJob, ctype: str, reason: str, message: str) -> None:
    if is_job_completed(job.status):
        return
    set_condition(job.status, new_condition(ctype, reason, message))
    job.status.phase = ctype


def is_failed_phase(phase: str) -> bool:
    return phase != C.PHASE_SUCCEEDED and phase in C.ENDING_PHASES


# ------------------------------------------------------------------------------ counters
def pod_phase(pod: dict) -> str:
    return pod.get("status", {}).get("phase") or C.POD_PENDING


def pod_node(pod: dict) -> str:
    return pod.get("spec", {}).get("nodeName") or ""


def is_draining(pod: dict) -> bool:
    """Replica being removed by a scale-down (index >= replicas): not part of the job any more."""
    return C.ANN_SCALE_DOWN in M.annotations_of(pod)


def initialize_replica_statuses(job: AITrainingJob, rtype: str) -> None:
    job.status.replica_statuses[rtype] = ReplicaStatus()


def initialize_restart_counts(job: AITrainingJob, rtype: str) -> None:
    job.status.restart_counts.setdefault(rtype, 0)


def update_restart_count(job: AITrainingJob, rtype: str) -> None:
    """status.go:322-330: scope All bumps every role."""
    if effective_restart_scope(job, rtype) == C.RESTART_SCOPE_ALL:
        for r in job.spec.replica_specs:
            job.status.restart_counts[r] = job.status.restart_counts.get(r, 0) + 1
    else:
        job.status.restart_counts[rtype] = job.status.restart_counts.get(rtype, 0) + 1


def count_pod(job: AITrainingJob, rtype: str, pod: dict, rs: ReplicaStatus) -> None:
    phase = pod_phase(pod)
    if phase == C.POD_PENDING:
        if job.status.restart_counts.get(rtype, 0) > 0:
            rs.restarting += 1
        elif pod_node(pod):
            rs.scheduled += 1
        else:
            rs.pending += 1
    elif phase == C.POD_RUNNING:
        rs.active += 1
    elif phase == C.POD_SUCCEEDED:
        rs.succeeded += 1
    elif phase in (C.POD_FAILED, C.POD_UNKNOWN):
        rs.failed += 1


def update_replica_statuses(job: AITrainingJob, rtype: str, pods: List[dict]) -> None:
    rs = ReplicaStatus()
    for pod in pods:
        if is_draining(pod):
            continue
        count_pod(job, rtype, pod, rs)
    job.status.replica_statuses[rtype] = rs


def effective_restart_scope(job: AITrainingJob, rtype: str) -> str:
    """``restartScope`` of a role; an elastic ``faultTolerant`` job replaces only the lost replica (scope Pod)."""
    spec = job.spec.replica_specs[rtype]
    if job.spec.fault_tolerant and spec.edl_policy in (C.EDL_POLICY_AUTO, C.EDL_POLICY_MANUAL):
        return C.RESTART_SCOPE_POD
    return spec.restart_scope


def filter_pods_for_replica_type(pods: List[dict], rt_lower: str) -> List[dict]:
    return [p for p in pods if M.labels_of(p).get(C.LABEL_REPLICA_NAME) == rt_lower]


# ------------------------------------------------------------------------------ the engine
class StatusEngine:
    """Mixed into ``TrainingJobController``; needs ``self.trainingjob_client``, ``self.enqueue_job``,
    ``self.delete_pods_and_services``."""

    def update_status(self, job: AITrainingJob, pods: List[dict], services: List[dict],
                      job_phases: Dict[str, str], message: str) -> None:
        live_pods = [p for p in pods if not is_draining(p)]
        for rtype in job.spec.replica_specs:
            update_replica_statuses(job, rtype, filter_pods_for_replica_type(pods, rtype.lower()))

        # -- restart barrier: wait until the deleted pods are really gone --------------------------
        if job.status.restart_replica_name:
            rname = job.status.restart_replica_name
            spec = job.spec.replica_specs.get(rname)
            if spec is None:
                job.status.restart_replica_name = ""
                return
            reason = C.TRAINING_JOB_REASON[C.PHASE_RESTARTING]
            replica_pods = filter_pods_for_replica_type(live_pods, rname.lower())
            scope = effective_restart_scope(job, rname)
            if scope == C.RESTART_SCOPE_ALL and not live_pods:
                update_conditions(job, C.PHASE_RESTARTING, reason, "All pods are restarting now")
                job.status.restart_replica_name = ""
            elif scope == C.RESTART_SCOPE_REPLICA and not replica_pods:
                update_conditions(job, C.PHASE_RESTARTING, reason, f"{rname.lower()} pods are restarting now")
                job.status.restart_replica_name = ""
            elif scope == C.RESTART_SCOPE_POD and len(replica_pods) < int(spec.replicas or 0):
                update_conditions(job, C.PHASE_RESTARTING, reason, "pod is restarting now")
                job.status.restart_replica_name = ""
            return

        now = M.now()
        spec = job.spec
        if not job.status.start_time:
            job.status.start_time = M.format_time(now)  # "acknowledged by the job controller" (types.go:87-88)
        completed = sum(1 for ph in job_phases.values() if ph == C.PHASE_SUCCEEDED)
        failed_roles = [ph for ph in job_phases.values() if is_failed_phase(ph)]
        ending_phase = failed_roles[-1] if failed_roles else ""
        replica_count = len(spec.replica_specs)

        # complete policy has priority over fail policy (status.go:159-174)
        if spec.complete_policy == C.ENDING_POLICY_ANY and completed > 0:
            return self.terminate_training_job(job, pods, services, C.PHASE_SUCCEEDED, f"job {job.name} completed")
        if spec.complete_policy == C.ENDING_POLICY_ALL and completed == replica_count:
            return self.terminate_training_job(job, pods, services, C.PHASE_SUCCEEDED, f"job {job.name} completed")
        if spec.fail_policy == C.ENDING_POLICY_ANY and failed_roles:
            return self.terminate_training_job(job, pods, services, ending_phase, message)
        if spec.fail_policy == C.ENDING_POLICY_ALL and len(failed_roles) == replica_count:
            return self.terminate_training_job(job, pods, services, ending_phase, message)

        # finalisation through an ending-phase annotation (status.go:176-187)
        for phase in C.ENDING_PHASES:
            if phase in job.annotations:
                msg = job.annotations[phase]
                if not pods:
                    job.status.end_time = M.format_time(now)
                    update_conditions(job, phase, C.TRAINING_JOB_REASON[phase], f"{msg}; deleted pods")
                else:
                    self.enqueue_job(job, True, 0)
                return

        # time limit (status.go:189-198)
        if spec.time_limit is not None and job.status.start_running_time:
            elapsed = M.seconds_since(job.status.start_running_time, now)
            if int(elapsed) >= spec.time_limit:
                started = M.parse_time(job.status.start_running_time)
                msg = (f"started at {started.strftime('%Y-%m-%d %H:%M:%S')},current time is "
                       f"{now.strftime('%Y-%m-%d %H:%M:%S')}, timeLimit is {spec.time_limit} second")
                klog.info("job %s: %s", job.name, msg)
                return self.terminate_training_job(job, pods, services, C.PHASE_TIMEOUT, msg)

        is_scheduled, is_creating, is_running, is_restarting = True, False, True, False
        for rtype, rspec in spec.replica_specs.items():
            replicas = int(rspec.replicas or 0)
            rs = job.status.replica_statuses[rtype]
            is_scheduled = is_scheduled and (rs.scheduled + rs.active + rs.succeeded + rs.failed + rs.restarting
                                             == replicas)
            is_creating = is_creating or rs.scheduled > 0
            is_restarting = is_restarting or rs.restarting > 0
            is_running = is_running and replicas == rs.active
        klog.V(4).info("state => %s %s %s %s", is_scheduled, is_creating, is_restarting, is_running)

        if job.status.phase != C.PHASE_RUNNING and is_running:
            if not job.status.start_running_time:
                job.status.start_running_time = M.format_time(now)
            update_conditions(job, C.PHASE_RUNNING, C.TRAINING_JOB_REASON[C.PHASE_RUNNING], "all pods are running")
        if is_creating and is_scheduled and job.status.phase != C.PHASE_RESTARTING:
            update_conditions(job, C.PHASE_CREATING, C.TRAINING_JOB_REASON[C.PHASE_CREATING], message)
        if is_restarting and job.status.phase != C.PHASE_RESTARTING:
            update_conditions(job, C.PHASE_RESTARTING, C.TRAINING_JOB_REASON[C.PHASE_RESTARTING], message)
        if not is_scheduled and not is_restarting and job.status.phase != C.PHASE_RESTARTING:
            if not job.status.start_time:
                job.status.start_time = M.format_time(now)
            update_conditions(job, C.PHASE_PENDING, C.TRAINING_JOB_REASON[C.PHASE_PENDING],
                              "all pods are waiting for scheduling")

        if spec.time_limit is not None and job.status.start_running_time:
            remaining = spec.time_limit - int(M.seconds_since(job.status.start_running_time))
            klog.V(2).info("Job with TimeLimit will sync after %d seconds", remaining)
            self.enqueue_job(job, False, max(remaining, 0) + 0.05)

    # -------------------------------------------------------------------------------------------
    def terminate_training_job(self, job: AITrainingJob, pods: List[dict], services: List[dict], ending_phase: str,
                               message: str) -> None:
        """status.go:256-283."""
        keep = job.spec.clean_pod_policy in (None, C.CLEAN_POD_POLICY_NONE) and \
            ending_phase in (C.PHASE_SUCCEEDED, C.PHASE_FAILED)
        if keep:
            job.status.end_time = M.format_time()  # Q8: the reference forgets endTime on this path
            update_conditions(job, ending_phase, C.TRAINING_JOB_REASON[ending_phase], f"{message}; kept pods")
            return
        job.set_annotation(ending_phase, message)
        self.delete_pods_and_services(job, pods, services)
        update_conditions(job, C.PHASE_TERMINATING, C.TRAINING_JOB_REASON[C.PHASE_TERMINATING],
                          f"{message}; deleting pods")

    # -------------------------------------------------------------------------------------------
    def update_training_job_phase(self, job: AITrainingJob) -> AITrainingJob:
        """Persist status (+ annotations, + defaulted spec) with optimistic concurrency, 5 attempts."""
        client = self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace)
        last: Optional[Exception] = None
        cur = job
        for attempt in range(5):
            klog.V(4).info("try %d time update job phase %s", attempt, job.status.phase)
            try:
                return client.update(cur)
            except APIError as e:
                last = e
                if e.reason == "NotFound":
                    raise
                klog.V(2).info("update job phase %s failed: %s", job.status.phase, e.message)
            try:
                fresh = client.get(job.name)
            except APIError as e:
                last = e
                if e.reason == "NotFound":
                    raise
                continue
            if fresh.uid != job.uid:
                raise APIError(409, "Conflict", f"job {job.key()} was recreated (uid changed)")
            fresh.status = job.status
            ann = dict(fresh.annotations)
            ann.update(job.annotations)
            if ann:
                fresh.metadata["annotations"] = ann
            from ..api.defaults import set_defaults_aitrainingjob

            set_defaults_aitrainingjob(fresh)
            cur = fresh
        raise last if last else RuntimeError("status update failed")
