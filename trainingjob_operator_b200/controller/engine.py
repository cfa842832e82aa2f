"""The reconcile engine: one pure function from what was observed to what must happen.

    reconcile(Observation) -> Decision

``Observation`` is everything a pass may look at -- the job (a private copy, defaults applied), the pods and services
it owns, the Ready nodes, the clock, the operator options, a view of the rest of the box for ``edlPolicy: Auto`` and a
few free loopback ports.  ``Decision`` is everything the pass wants done -- pods / services to create, delete or
patch, a spec patch (auto-scale), the next status, controller-owned annotations, re-queue requests, metric bumps, log
lines.  The engine holds no client, queue, expectations cache or clock, so a recorded Observation replays to the same
Decision (``tests/test_engine_golden.py``), and ``controller.executor`` is the only place that performs I/O.

Behaviour is the reference's state machine (/root/reference/pkg/controller/controller.go:314-388,
pod.go:152-437, status.go:101-305) with the deliberate fixes of SURVEY.md §2.9 (Q1 scale-down drains and deletes,
Q3 no create burst while creations are in flight -- the executor's job --, Q8 ``endTime`` on the keep-pods path, Q9
services are deleted with the job's pods, Q10 no crash without ``startTime``), but it is organised as data:

* ``pod.RESTART_MATRIX``      restart policy x replica health -> restart?
* ``ROLE_VERDICTS``           per-role complete / fail policy x (index, replica outcome) -> ending phase
* ``RESTART_VICTIMS``         restart scope -> which pods go down together
* ``BARRIER_LIFTED``          restart scope -> "the victims are really gone"
* ``JOB_VERDICTS``            job-level complete / fail policy (complete beats fail)
* ``PHASE_RULES``             replica counters -> Running / Creating / Restarting / Pending

Condition messages and Event reasons are API (they show up in ``kubectl describe``) and are kept verbatim.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaSpec, ReplicaStatus
from . import elastic as E
from . import status as S
from .pod import (ANN_HOST_PORTS, ReplicaView, StartWindow, bucket_by_index, build_pod_template, classify_replica,
                  get_ports_from_job, host_port_map, of_role, pod_node, pod_phase, scheduling_message,
                  wanted_host_port_keys)
from .service import build_service, has_contract_container


# ============================================================================================ inputs / outputs
@dataclass(frozen=True)
class EngineOptions:
    window: StartWindow = StartWindow()
    scale_down_grace: float = 30.0
    master_url: str = ""

    @staticmethod
    def from_operator_option(opt) -> "EngineOptions":
        return EngineOptions(StartWindow(float(getattr(opt, "creating_restart_time", 0.0)),
                                         float(getattr(opt, "creating_duration_time", 900.0)),
                                         bool(getattr(opt, "enable_creating_failed", False))),
                             float(getattr(opt, "scale_down_grace", 30.0)), getattr(opt, "master_url", "") or "")


@dataclass
class Observation:
    job: AITrainingJob                    # private copy with defaults applied; the engine edits its status in place
    pods: List[dict]                      # pods claimed for the job (private copies)
    services: List[dict]
    ready_nodes: frozenset
    now: object                           # datetime (UTC)
    now_epoch: float                      # the same instant as a float, for the sub-second trace annotation
    options: EngineOptions = EngineOptions()
    cluster: Optional[E.ClusterView] = None   # only needed for edlPolicy: Auto roles
    spare_ports: Tuple[int, ...] = ()     # free loopback ports the engine may hand out (see Decision.ports_wanted)


@dataclass
class PodCreate:
    role: str          # lower-case role
    index: int
    template: dict


@dataclass
class PodDelete:
    namespace: str
    name: str
    role: str          # lower-case role label of the pod (expectation key)
    grace: Optional[int] = None   # 0 = force delete (replica on a failed node); None = the pod's own grace period
    why: str = ""


@dataclass
class Decision:
    ports_wanted: int = 0                 # > 0: nothing was decided; call again with this many spare_ports
    spec_patch: Optional[dict] = None     # auto-scale: patch the job spec and stop (the update event re-queues)
    pod_patches: List[Tuple[str, str, dict]] = field(default_factory=list)    # (namespace, name, patch)
    pod_deletes: List[PodDelete] = field(default_factory=list)
    pod_creates: List[PodCreate] = field(default_factory=list)
    service_creates: List[Tuple[str, dict]] = field(default_factory=list)     # (lower-case role, service)
    service_deletes: List[Tuple[str, str]] = field(default_factory=list)      # (namespace, name)
    annotations: Dict[str, str] = field(default_factory=dict)                 # controller-owned keys set this pass
    write_status: bool = False
    requeue_after: List[float] = field(default_factory=list)
    requeue_rate_limited: bool = False
    counters: List[Tuple[str, Dict[str, str]]] = field(default_factory=list)  # metrics to bump
    observations: List[Tuple[str, float]] = field(default_factory=list)       # histogram samples
    log: List[Tuple[str, str]] = field(default_factory=list)                  # (level, line)
    role_outcomes: Dict[str, Tuple[str, str]] = field(default_factory=dict)   # role -> (ending phase, message)

    def note(self, line: str, level: str = "info") -> None:
        self.log.append((level, line))


# ============================================================================================ role level
@dataclass
class RoleOutcome:
    phase: str = C.PHASE_NONE     # "" | Restarting | Terminating | an ending phase
    message: str = ""


def _verdict_any_ok(spec: ReplicaSpec, idx: int, v: ReplicaView):
    if spec.complete_policy == C.ENDING_POLICY_ANY and v.finished_ok:
        return C.PHASE_SUCCEEDED, f"pod {M.name_of(v.pod)} have completed"


def _verdict_any_bad(spec: ReplicaSpec, idx: int, v: ReplicaView):
    if spec.fail_policy == C.ENDING_POLICY_ANY and v.finished_bad:
        return v.phase, f"pod {M.name_of(v.pod)} is failed, {v.message}"


def _verdict_rank0_ok(spec: ReplicaSpec, idx: int, v: ReplicaView):
    if idx == 0 and spec.complete_policy == C.ENDING_POLICY_RANK0 and v.finished_ok:
        return C.PHASE_SUCCEEDED, f"rank0 pod {M.name_of(v.pod)} have completed"


def _verdict_rank0_bad(spec: ReplicaSpec, idx: int, v: ReplicaView):
    if idx == 0 and spec.fail_policy == C.ENDING_POLICY_RANK0 and v.finished_bad:
        return v.phase, f"rank0 pod {M.name_of(v.pod)} is failed, {v.message}"


# evaluated for every replica in index order, first hit ends the role; the ``All`` policies are counted afterwards
ROLE_VERDICTS: Tuple[Callable, ...] = (_verdict_any_ok, _verdict_any_bad, _verdict_rank0_ok, _verdict_rank0_bad)

# restart scope -> the pods that go down together with the failed replica ``pod``
RESTART_VICTIMS: Dict[str, Callable] = {
    C.RESTART_SCOPE_POD: lambda pod, role_pods, job_pods: [pod],
    C.RESTART_SCOPE_REPLICA: lambda pod, role_pods, job_pods: list(role_pods),
    C.RESTART_SCOPE_ALL: lambda pod, role_pods, job_pods: list(job_pods),
}

# restart scope -> (are the victims gone?, condition message); arguments: live pods of the job, of the role, replicas
BARRIER_LIFTED: Dict[str, Tuple[Callable, Callable]] = {
    C.RESTART_SCOPE_ALL: (lambda job_pods, role_pods, replicas: not job_pods,
                          lambda rt: "All pods are restarting now"),
    C.RESTART_SCOPE_REPLICA: (lambda job_pods, role_pods, replicas: not role_pods,
                              lambda rt: f"{rt} pods are restarting now"),
    C.RESTART_SCOPE_POD: (lambda job_pods, role_pods, replicas: len(role_pods) < replicas,
                          lambda rt: "pod is restarting now"),
}


def drain_surplus(obs: Observation, d: Decision, surplus: List[dict]) -> None:
    """Scale-down (quirk Q1): replicas whose index fell out of range are marked draining -- from then on they are not
    counted --, leave at a step boundary, and are deleted once they exited or ``--scale-down-grace`` ran out."""
    grace = obs.options.scale_down_grace
    for pod in surplus:
        ann = M.annotations_of(pod)
        if C.ANN_SCALE_DOWN in ann:
            drained_for = M.seconds_since(ann[C.ANN_SCALE_DOWN], obs.now)
        else:
            stamp = M.format_time(obs.now)
            d.pod_patches.append((M.namespace_of(pod), M.name_of(pod),
                                  {"metadata": {"annotations": {C.ANN_SCALE_DOWN: stamp}}}))
            pod.setdefault("metadata", {}).setdefault("annotations", {})[C.ANN_SCALE_DOWN] = stamp
            drained_for = 0.0
        gone = pod_phase(pod) in (C.POD_SUCCEEDED, C.POD_FAILED) or not pod_node(pod)
        if gone or drained_for >= grace:
            if not pod.get("metadata", {}).get("deletionTimestamp"):
                d.note(f"scale-down: deleting replica {M.name_of(pod)} (index out of range)")
                d.pod_deletes.append(PodDelete(M.namespace_of(pod), M.name_of(pod),
                                               M.labels_of(pod).get(C.LABEL_REPLICA_NAME, ""), None, "scale-down"))
        else:
            d.requeue_after.append(max(0.2, min(1.0, grace - drained_for)))


def plan_role(obs: Observation, d: Decision, rtype: str) -> RoleOutcome:
    """One role's replicas: what is missing, what each existing replica is up to, whether that restarts something or
    ends the role.  Missing replicas are only created when the role neither restarts nor ends in this pass."""
    job = obs.job
    if job.status.phase == C.PHASE_TERMINATING:
        return RoleOutcome(C.PHASE_TERMINATING)
    for external in (C.PHASE_PREEMPTED, C.PHASE_FAILED):      # annotations set from outside: preempt / fail the job
        if external in job.annotations:
            return RoleOutcome(external, job.annotations[external])

    rt = rtype.lower()
    spec = job.spec.replica_specs[rtype]
    replicas = int(spec.replicas or 0)
    job.status.restart_counts.setdefault(rtype, 0)
    slots, surplus = bucket_by_index(of_role(obs.pods, rt), replicas)
    drain_surplus(obs, d, surplus)

    waiting_note = ""
    missing: List[int] = []
    starting: Dict[str, List[str]] = {}
    complaints: List[str] = []
    worst = C.PHASE_FAILED
    for idx, slot in enumerate(slots):
        if not slot:
            d.note(f"Need to create new pod: {job.namespace}/{job.name} {rt}-{idx}")
            missing.append(idx)
            continue
        pod = slot[0]
        sched = scheduling_message(pod)
        if sched:
            waiting_note = f"{rt}: {sched} "
        view = classify_replica(job, pod, obs.ready_nodes, obs.options.window, obs.now)
        if view.message:
            complaints.append(view.message)
        if view.wants_restart(spec.restart_policy) and \
                (spec.restart_limit is None or job.status.restart_counts.get(rtype, 0) < spec.restart_limit):
            return _restart(obs, d, rtype, view, [p for sl in slots for p in sl])
        if view.phase == C.PHASE_CREATING:
            starting.setdefault(view.message, []).append(M.name_of(pod))
        for rule in ROLE_VERDICTS:
            hit = rule(spec, idx, view)
            if hit:
                return RoleOutcome(*hit)
        if view.phase == C.PHASE_NODE_FAIL:
            worst = C.PHASE_NODE_FAIL

    if missing:
        ports = host_port_map(job)
        count = str(job.status.restart_counts.get(rtype, 0))
        d.pod_creates += [PodCreate(rt, i, build_pod_template(job, rt, str(i), count, spec, ports,
                                                              obs.options.master_url)) for i in missing]
    counted = ReplicaStatus()
    for slot in slots:
        if slot and not S.is_draining(slot[0]):
            S.count_pod(job, rtype, slot[0], counted)
    if replicas > 0 and spec.complete_policy == C.ENDING_POLICY_ALL and counted.succeeded == replicas:
        return RoleOutcome(C.PHASE_SUCCEEDED, f"All {rtype} pods have completed")
    if replicas > 0 and spec.fail_policy == C.ENDING_POLICY_ALL and counted.failed == replicas:
        return RoleOutcome(worst, f"All {rtype} pods are failed, {', '.join(complaints) or waiting_note}")
    if starting:
        return RoleOutcome(C.PHASE_NONE, ", ".join(f"pods {names} {m}" for m, names in starting.items()))
    return RoleOutcome(C.PHASE_NONE, waiting_note)


def _restart(obs: Observation, d: Decision, rtype: str, view: ReplicaView, role_pods: List[dict]) -> RoleOutcome:
    job = obs.job
    spec = job.spec.replica_specs[rtype]
    scope = S.effective_restart_scope(job, rtype)
    bumped = list(job.spec.replica_specs) if scope == C.RESTART_SCOPE_ALL else [rtype]
    for r in bumped:
        job.status.restart_counts[r] = job.status.restart_counts.get(r, 0) + 1
    victims = RESTART_VICTIMS.get(scope, RESTART_VICTIMS[C.RESTART_SCOPE_ALL])(view.pod, role_pods, obs.pods)
    grace = 0 if view.phase == C.PHASE_NODE_FAIL else None      # nobody is left on a failed node to honour a grace period
    d.note({C.RESTART_SCOPE_POD: f"According to restartscope, need to restart the pod: "
                                 f"{M.namespace_of(view.pod)}.{M.name_of(view.pod)}",
            C.RESTART_SCOPE_REPLICA: f"According to restartscope, need to restart all pods of the replica: {rtype}"}
           .get(scope, "According to restartscope, need to restart all pods"), "warning")
    d.pod_deletes += [PodDelete(M.namespace_of(p), M.name_of(p), M.labels_of(p).get(C.LABEL_REPLICA_NAME, ""), grace,
                                "restart") for p in victims]
    d.counters.append(("aitj_restarts_total", {"scope": spec.restart_scope}))
    return RoleOutcome(C.PHASE_RESTARTING, f"restart times is {job.status.restart_counts[rtype]}, {view.message} ")


def plan_services(obs: Observation, d: Decision, rtype: str) -> None:
    """One Service per (role, index) for roles that have an ``aitj-`` container; the address of a replica removed by a
    scale-down is dropped."""
    job = obs.job
    replicas = int(job.spec.replica_specs[rtype].replicas or 0)
    slots, surplus = bucket_by_index(of_role(obs.services, rtype.lower()), replicas)
    d.service_deletes += [(M.namespace_of(s), M.name_of(s)) for s in surplus]
    if not has_contract_container(job, rtype):
        return
    declared = get_ports_from_job(job, rtype)
    ports = host_port_map(job)
    for idx, slot in enumerate(slots):
        if not slot:
            d.note(f"need to create new service: {rtype.lower()}-{idx}")
            d.service_creates.append((rtype.lower(), build_service(job, rtype, idx, declared, ports)))


# ============================================================================================ job level
def _job_verdict(job: AITrainingJob, outcomes: Dict[str, str], message: str) -> Optional[Tuple[str, str]]:
    """Job-level ending policies over the roles' ending phases; completion is looked at before failure."""
    spec = job.spec
    done = sum(1 for ph in outcomes.values() if ph == C.PHASE_SUCCEEDED)
    bad = [ph for ph in outcomes.values() if S.is_failed_phase(ph)]
    roles = len(spec.replica_specs)
    verdicts = (
        (spec.complete_policy, done, C.PHASE_SUCCEEDED, f"job {job.name} completed"),
        (spec.fail_policy, len(bad), bad[-1] if bad else "", message),
    )
    for policy, hits, phase, msg in verdicts:
        if (policy == C.ENDING_POLICY_ANY and hits > 0) or (policy == C.ENDING_POLICY_ALL and hits == roles):
            return phase, msg
    return None


JOB_VERDICTS = _job_verdict


@dataclass
class _Placement:
    scheduled: bool = True       # every role has all of its replicas placed (or beyond)
    creating: bool = False       # somebody is placed but not running yet
    running: bool = True         # every replica of every role is active
    restarting: bool = False


# (applies?, phase, message or None for "the roles' aggregated message"); evaluated in order, each sees the phase the
# previous rule may have set -- once Restarting, only Running or an ending phase gets the job out
PHASE_RULES = (
    (lambda p, phase: p.running and phase != C.PHASE_RUNNING, C.PHASE_RUNNING, "all pods are running"),
    (lambda p, phase: p.creating and p.scheduled and phase != C.PHASE_RESTARTING, C.PHASE_CREATING, None),
    (lambda p, phase: p.restarting and phase != C.PHASE_RESTARTING, C.PHASE_RESTARTING, None),
    (lambda p, phase: not p.scheduled and not p.restarting and phase != C.PHASE_RESTARTING, C.PHASE_PENDING,
     "all pods are waiting for scheduling"),
)


def terminate(obs: Observation, d: Decision, ending_phase: str, message: str) -> None:
    """End the job.  ``cleanPodPolicy: None`` keeps the pods of a Succeed / Failed job (final condition right away);
    otherwise the verdict is parked in an annotation, everything is deleted and a later pass -- when the pods are gone --
    turns the annotation into the final condition."""
    job = obs.job
    keep = job.spec.clean_pod_policy in (None, C.CLEAN_POD_POLICY_NONE) and \
        ending_phase in (C.PHASE_SUCCEEDED, C.PHASE_FAILED)
    if keep:
        job.status.end_time = M.format_time(obs.now)
        S.enter_phase(job, ending_phase, f"{message}; kept pods", obs.now)
        return
    job.set_annotation(ending_phase, message)
    d.annotations[ending_phase] = message
    # nothing is born into a job that is being torn down (a replica created now would never be deleted again: the
    # passes that follow only wait for the pods to disappear)
    d.pod_creates.clear()
    d.service_creates.clear()
    d.pod_deletes += [PodDelete(M.namespace_of(p), M.name_of(p), M.labels_of(p).get(C.LABEL_REPLICA_NAME, ""), None,
                                "terminate") for p in obs.pods if not p.get("metadata", {}).get("deletionTimestamp")]
    d.service_deletes += [(M.namespace_of(s), M.name_of(s)) for s in obs.services]
    S.enter_phase(job, C.PHASE_TERMINATING, f"{message}; deleting pods", obs.now)


def derive_status(obs: Observation, d: Decision, outcomes: Dict[str, str], message: str) -> None:
    """Counters, restart barrier, job-level verdicts, annotation-driven finalisation, time limit, phase."""
    job, now = obs.job, obs.now
    st = job.status
    for rtype in job.spec.replica_specs:
        st.replica_statuses[rtype] = S.count_role(job, rtype, obs.pods)

    if st.restart_replica_name:
        # restart barrier: nothing else happens until the deleted replicas are really gone
        rname = st.restart_replica_name
        spec = job.spec.replica_specs.get(rname)
        if spec is None:
            st.restart_replica_name = ""
            return
        live = [p for p in obs.pods if not S.is_draining(p)]
        lifted, text = BARRIER_LIFTED.get(S.effective_restart_scope(job, rname), BARRIER_LIFTED[C.RESTART_SCOPE_ALL])
        if lifted(live, of_role(live, rname.lower()), int(spec.replicas or 0)):
            S.enter_phase(job, C.PHASE_RESTARTING, text(rname.lower()), now)
            st.restart_replica_name = ""
        return

    if not st.start_time:
        st.start_time = M.format_time(now)
    verdict = JOB_VERDICTS(job, outcomes, message)
    if verdict is not None:
        return terminate(obs, d, *verdict)

    for phase in C.ENDING_PHASES:          # a parked verdict (see ``terminate``), or one written from outside
        if phase in job.annotations:
            if obs.pods:
                d.requeue_rate_limited = True
            else:
                st.end_time = M.format_time(now)
                S.enter_phase(job, phase, f"{job.annotations[phase]}; deleted pods", now)
            return

    limit = job.spec.time_limit
    if limit is not None and st.start_running_time:
        elapsed = M.seconds_since(st.start_running_time, now)
        if int(elapsed) >= limit:
            started = M.parse_time(st.start_running_time)
            msg = (f"started at {started.strftime('%Y-%m-%d %H:%M:%S')},current time is "
                   f"{now.strftime('%Y-%m-%d %H:%M:%S')}, timeLimit is {limit} second")
            d.note(f"job {job.name}: {msg}")
            return terminate(obs, d, C.PHASE_TIMEOUT, msg)

    p = _Placement()
    for rtype, rspec in job.spec.replica_specs.items():
        n, rs = int(rspec.replicas or 0), st.replica_statuses[rtype]
        p.scheduled &= rs.scheduled + rs.active + rs.succeeded + rs.failed + rs.restarting == n
        p.creating |= rs.scheduled > 0
        p.restarting |= rs.restarting > 0
        p.running &= rs.active == n
    for applies, phase, text in PHASE_RULES:
        if applies(p, st.phase):
            if phase == C.PHASE_RUNNING and not st.start_running_time:
                st.start_running_time = M.format_time(now)
            S.enter_phase(job, phase, message if text is None else text, now)

    if limit is not None and st.start_running_time:
        d.requeue_after.append(max(limit - int(M.seconds_since(st.start_running_time, now)), 0) + 0.05)


# ============================================================================================ rendezvous / ports
def _trace(obs: Observation, d: Decision, event: str, extra: Optional[dict] = None) -> None:
    """Sub-second lifecycle timestamps in the ``aitj.b200/trace`` annotation (metav1.Time has 1 s resolution)."""
    job = obs.job
    try:
        tr = json.loads(job.annotations.get(C.ANN_TRACE) or "{}")
    except ValueError:
        tr = {}
    if extra is not None:
        tr.setdefault(event, []).append(extra)
        tr[event] = tr[event][-16:]
    elif event in tr and event != "ended":
        return
    else:
        tr[event] = round(obs.now_epoch, 4)
    raw = json.dumps(tr, sort_keys=True)
    job.set_annotation(C.ANN_TRACE, raw)
    d.annotations[C.ANN_TRACE] = raw


class _Ports:
    """Hands out the observation's spare loopback ports and counts how many more would have been needed."""

    def __init__(self, spare: Tuple[int, ...]):
        self.spare = list(spare)
        self.short = 0

    def take(self) -> int:
        if self.spare:
            return self.spare.pop(0)
        self.short += 1
        return 0


def _bump_generation(obs: Observation, d: Decision, ports: _Ports, why: str,
                     world_sizes: Optional[Dict[str, int]] = None) -> None:
    rdv = E.next_generation(obs.job, ports.take(), obs.now, world_sizes)
    d.note(f"job {obs.job.key()}: rendezvous generation {rdv.generation} ({why}): world {rdv.world_sizes} "
           f"port {rdv.master_port}")
    d.counters.append(("aitj_rendezvous_generations_total", {"reason": why}))


def _settle_rendezvous(obs: Observation, d: Decision, ports: _Ports) -> None:
    job = obs.job
    target = E.rendezvous_target(job)
    rdv = job.status.rendezvous
    if rdv is None:
        _bump_generation(obs, d, ports, "start", target)
    elif target != rdv.world_sizes:
        if obs.pods:      # a running / starting job needs a new generation; before any pod exists adopt the sizes
            _bump_generation(obs, d, ports, "rescale", target)
            _trace(obs, d, "rescales", {"generation": rdv.generation, "at": round(obs.now_epoch, 4),
                                        "world": dict(rdv.world_sizes)})
        else:
            rdv.world_sizes = target


def _settle_host_ports(obs: Observation, d: Decision, ports: _Ports) -> None:
    """Every (role, index, declared port) gets its own loopback port, remembered on the job (ANN_HOST_PORTS)."""
    job = obs.job
    have = host_port_map(job)
    missing = [k for k in wanted_host_port_keys(job) if k not in have]
    if not missing:
        return
    for k in missing:
        have[k] = ports.take()
    raw = json.dumps(have, sort_keys=True)
    job.set_annotation(ANN_HOST_PORTS, raw)
    d.annotations[ANN_HOST_PORTS] = raw


# ============================================================================================ the pass
def reconcile(obs: Observation) -> Decision:
    job = obs.job
    d = Decision()
    before = (job.status.to_dict(), dict(job.annotations), job.spec.to_dict())
    ports = _Ports(obs.spare_ports)
    _trace(obs, d, "firstReconcile")

    # edlPolicy: Auto -- not while replicas are being torn down for a restart
    if E.auto_roles(job) and job.status.phase != C.PHASE_TERMINATING and not job.status.restart_replica_name:
        patch, notes = E.plan_autoscale(job, obs.pods, obs.cluster or E.ClusterView())
        for n in notes:
            d.note(n)
        if patch:
            d.spec_patch = {"spec": {"replicaSpecs": patch}}
            d.counters.append(("aitj_autoscale_total", {}))
            return d
        # slots freed by other jobs, or more important pods that cannot be placed, raise no event on this job
        d.requeue_after.append(E.AUTO_RECHECK_SECONDS)

    outcomes: Dict[str, str] = {}
    notes: List[str] = []
    if not job.status.restart_replica_name:
        _settle_rendezvous(obs, d, ports)
        _settle_host_ports(obs, d, ports)
        for rtype in list(job.spec.replica_specs):
            out = plan_role(obs, d, rtype)
            d.role_outcomes[rtype] = (out.phase, out.message)
            if out.message and out.message not in notes:
                notes.append(out.message)
            if out.phase == C.PHASE_RESTARTING:
                # the whole pass turns into "tear down, then come back": replicas planned for creation stay unborn
                d.pod_creates.clear()
                d.service_creates.clear()
                S.enter_phase(job, C.PHASE_TERMINATING, out.message, obs.now)
                job.status.restart_replica_name = rtype
                _bump_generation(obs, d, ports, "restart")
                break
            if out.phase:
                outcomes[rtype] = out.phase
            else:
                plan_services(obs, d, rtype)
    if ports.short:
        return Decision(ports_wanted=len(obs.spare_ports) + ports.short)

    was = job.status.phase
    derive_status(obs, d, outcomes, "; ".join(notes))
    if job.status.phase == C.PHASE_RUNNING and was != C.PHASE_RUNNING:
        _trace(obs, d, "running")
        try:
            tr = json.loads(job.annotations.get(C.ANN_TRACE, "{}"))
            t0 = tr.get("submitted") or tr.get("firstReconcile")
            if t0 and "running" in tr:
                d.observations.append(("aitj_job_startup_seconds", tr["running"] - t0))
        except ValueError:
            pass
    if job.status.phase in C.ENDING_PHASES:
        _trace(obs, d, "ended")
    if (job.status.to_dict(), dict(job.annotations), job.spec.to_dict()) != before:
        job.status.last_reconcile_time = M.format_time(obs.now)
        d.write_status = True
    return d
