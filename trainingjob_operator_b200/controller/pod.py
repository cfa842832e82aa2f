"""Replica-level building blocks of the reconcile engine: naming, health classification, pod construction.

Everything here is a pure function of its arguments (no client, queue or clock), so ``controller.engine`` can be
replayed from recorded inputs.  What the reference specifies and where:

* names ``<job>-<role>-<index>`` (/root/reference/pkg/controller/trainingjob.go:12-15) and the expectation key
  ``<ns>/<job>/<role>/pods`` (upstream ``GenExpectationPodsKey``);
* which exit codes are retryable (/root/reference/pkg/controller/controller.go:442-462);
* what a replica's container states mean (/root/reference/pkg/controller/pod.go:328-437) -- expressed here as a
  ``Health`` value plus ``RESTART_MATRIX`` (restart policy x health -> restart?) instead of an ``if`` ladder;
* labels, owner-facing metadata and the 13-variable environment contract of a replica
  (/root/reference/pkg/controller/pod.go:483-652), extended by the torch / paddle / TF_CONFIG dialects and the
  rendezvous generation.

Loopback ports: on one box every replica shares 127.0.0.1, so each (role, index, declared ``aitj-`` port) is mapped to
its own free host port.  The mapping is allocated per job, kept in the ``aitj.b200/host-ports`` annotation and
recorded in the replica's Service (``spec.ports[].hostPort``); two jobs that declare the same port never collide.
"""
from __future__ import annotations

import enum
import json
from dataclasses import dataclass, field
from typing import Dict, FrozenSet, List, Optional, Tuple

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaSpec
from ..api.validation import parse_exit_codes

ANN_HOST_PORTS = "aitj.b200/host-ports"     # on a job: {"<role>/<index>/<declared port>": <host port>}


# ------------------------------------------------------------------------------------------ names / keys
def gen_expectation_pods_key(job_key: str, rt: str) -> str:
    return f"{job_key}/{rt.lower()}/pods"


def gen_general_name(job_name: str, rtype: str, index: str) -> str:
    return f"{job_name}-{rtype}-{index}".replace("/", "-")


def is_retryable_exit_code(exit_codes: List[int], restarting_exit_code: str) -> bool:
    """Every collected code must be listed in ``spec.restartingExitCode``; no codes at all is not retryable."""
    if not exit_codes:
        return False
    try:
        allowed = set(parse_exit_codes(restarting_exit_code))
    except ValueError:
        return False
    return all(c in allowed for c in exit_codes)


def get_ports_from_container(container: dict) -> List[int]:
    """Declared ports that take part in the contract: ``aitj-`` port names of ``aitj-`` containers (service.go:33-43)."""
    if not str(container.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX):
        return []
    return [p.get("containerPort") for p in container.get("ports") or []
            if str(p.get("name", "")).startswith(C.DEFAULT_PORT_PREFIX)]


def get_ports_from_job(job: AITrainingJob, rtype: str) -> List[int]:
    return [p for c in job.spec.replica_specs[rtype].containers() for p in get_ports_from_container(c)]


# ------------------------------------------------------------------------------------------ loopback port map
def host_port_key(rtype: str, index: int, port: int) -> str:
    return f"{rtype.lower()}/{index}/{port}"


def host_port_map(job: AITrainingJob) -> Dict[str, int]:
    raw = job.annotations.get(ANN_HOST_PORTS)
    if not raw:
        return {}
    try:
        d = json.loads(raw)
        return {str(k): int(v) for k, v in d.items()} if isinstance(d, dict) else {}
    except (ValueError, TypeError):
        return {}


def wanted_host_port_keys(job: AITrainingJob) -> List[str]:
    """Every (role, index < replicas, declared port) of the job, in a stable order."""
    keys = []
    for rt in sorted(job.spec.replica_specs):
        ports = get_ports_from_job(job, rt)
        for i in range(int(job.spec.replica_specs[rt].replicas or 0)):
            keys += [host_port_key(rt, i, p) for p in ports]
    return keys


def host_port(ports: Dict[str, int], rtype: str, index: int, port: int) -> int:
    """The loopback port (role, index, declared port) is served on; the declared port itself until one is assigned."""
    return int(ports.get(host_port_key(rtype, index, port), port))


# ------------------------------------------------------------------------------------------ replica health
class Health(enum.Enum):
    STARTING = "starting"            # a container is still waiting to start
    START_FAILED = "start-failed"    # start error outlived the retry window (--enable-creating-failed)
    RUNNING = "running"
    COMPLETED = "completed"          # every aitj- container exited 0
    CRASHED_LISTED = "crashed-listed"  # pod Failed and every aitj- exit code is in restartingExitCode
    CRASHED = "crashed"              # pod Failed otherwise
    NODE_LOST = "node-lost"          # bound to a node (GPU slot) that is not Ready


# the phase vocabulary the job-level machinery speaks (types.go:98-124)
PHASE_OF_HEALTH = {
    Health.STARTING: C.PHASE_CREATING, Health.START_FAILED: C.PHASE_FAILED, Health.RUNNING: C.PHASE_NONE,
    Health.COMPLETED: C.PHASE_SUCCEEDED, Health.CRASHED_LISTED: C.PHASE_FAILED, Health.CRASHED: C.PHASE_FAILED,
    Health.NODE_LOST: C.PHASE_NODE_FAIL,
}

# restartPolicy -> the kinds of trouble it answers with a restart (``Always`` does not restart a success)
RESTART_MATRIX: Dict[str, FrozenSet[Health]] = {
    C.RESTART_POLICY_ALWAYS: frozenset({Health.CRASHED, Health.CRASHED_LISTED, Health.NODE_LOST}),
    C.RESTART_POLICY_ON_FAILURE: frozenset({Health.CRASHED, Health.CRASHED_LISTED}),
    C.RESTART_POLICY_ON_NODE_FAIL: frozenset({Health.NODE_LOST}),
    C.RESTART_POLICY_EXIT_CODE: frozenset({Health.CRASHED_LISTED}),
    C.RESTART_POLICY_ON_NODE_FAIL_WITH_EXIT_CODE: frozenset({Health.CRASHED_LISTED, Health.NODE_LOST}),
    C.RESTART_POLICY_NEVER: frozenset(),
}


@dataclass(frozen=True)
class StartWindow:
    """Spawn-failure retry window (``--creating-restart-period`` / ``--creating-duration-period`` /
    ``--enable-creating-failed``, options.go:55-70)."""
    restart_period: float = 0.0
    duration_period: float = 15 * 60.0
    fail_after_window: bool = False


@dataclass
class ReplicaView:
    """What the engine knows about one replica after looking at its pod."""
    pod: dict
    health: Health
    message: str = ""
    start_stuck: bool = False        # start error inside the retry window, stuck longer than the duration period
    exit_codes: List[int] = field(default_factory=list)

    @property
    def phase(self) -> str:
        return PHASE_OF_HEALTH[self.health]

    @property
    def pod_phase(self) -> str:
        return pod_phase(self.pod)

    def wants_restart(self, policy: str) -> bool:
        return self.start_stuck or self.health in RESTART_MATRIX.get(policy, frozenset())

    @property
    def finished_ok(self) -> bool:
        return self.health is Health.COMPLETED and self.pod_phase == C.POD_SUCCEEDED

    @property
    def finished_bad(self) -> bool:
        return self.phase in (C.PHASE_FAILED, C.PHASE_NODE_FAIL)


def pod_phase(pod: dict) -> str:
    return pod.get("status", {}).get("phase") or C.POD_PENDING


def pod_node(pod: dict) -> str:
    return pod.get("spec", {}).get("nodeName") or ""


def scheduling_message(pod: dict) -> str:
    """Why the scheduler could not place a pending replica (``PodScheduled=False``), if it said so."""
    if pod_phase(pod) == C.POD_PENDING and not pod_node(pod):
        for cond in pod.get("status", {}).get("conditions") or []:
            if cond.get("type") == "PodScheduled" and cond.get("status") == "False":
                return cond.get("message", "")
    return ""


def classify_replica(job: AITrainingJob, pod: dict, ready_nodes, window: StartWindow, now) -> ReplicaView:
    """Container states + pod phase + node readiness -> ``ReplicaView``.  Only ``aitj-`` containers contribute exit
    codes and success; any waiting container makes the replica STARTING."""
    pstatus = pod.get("status", {})
    node = pod_node(pod)
    codes: List[int] = []
    complaints: List[str] = []
    waiting_seen = False
    all_done_ok = True
    stuck = False
    for cs in pstatus.get("containerStatuses") or []:
        state = cs.get("state") or {}
        term, waiting = state.get("terminated"), state.get("waiting")
        if str(cs.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX):
            if term is None:
                all_done_ok = False
            else:
                code = int(term.get("exitCode", 0))
                codes.append(code)
                if code != 0:
                    all_done_ok = False
                    complaints.append(f"container {cs.get('name')} on node {node} exited with reason "
                                      f"{term.get('reason', '')} exitcode {code}")
        if waiting is None:
            continue
        waiting_seen = True
        reason = waiting.get("reason", "")
        if reason not in C.ERROR_CONTAINER_STATUS:
            continue
        creating = job.status.get_condition(C.PHASE_CREATING)
        if creating is not None and creating.status == "True":
            if M.seconds_since(creating.last_transition_time, now) < window.restart_period:
                started = pstatus.get("startTime")
                if started and M.seconds_since(started, now) > window.duration_period:
                    stuck = True
            elif window.fail_after_window:
                return ReplicaView(pod, Health.START_FAILED,
                                   f"pod {M.name_of(pod)} create container failed[{reason}] and has been retrying "
                                   f"for {window.restart_period:g} seconds", stuck, codes)
        complaints.append(reason)

    if pod_phase(pod) == C.POD_FAILED:
        listed = is_retryable_exit_code(codes, job.spec.restarting_exit_code)
        if complaints:
            msg = "; ".join(complaints)
        elif pstatus.get("reason"):
            msg = f"{pstatus['reason']}, {pstatus['message']}" if pstatus.get("message") else pstatus["reason"]
        else:
            msg = ""
        return ReplicaView(pod, Health.CRASHED_LISTED if listed else Health.CRASHED, msg, stuck, codes)
    if node and node not in ready_nodes:
        return ReplicaView(pod, Health.NODE_LOST, f"Node {node} is failed and offline", stuck, codes)
    if waiting_seen:
        return ReplicaView(pod, Health.STARTING, "; ".join(complaints) or "creating containers", stuck, codes)
    if all_done_ok:
        return ReplicaView(pod, Health.COMPLETED, "", stuck, codes)
    return ReplicaView(pod, Health.RUNNING, "", stuck, codes)


# ------------------------------------------------------------------------------------------ bucketing
def replica_index(obj: dict) -> Optional[int]:
    raw = M.labels_of(obj).get(C.LABEL_REPLICA_INDEX)
    try:
        return int(raw) if raw is not None else None
    except ValueError:
        return None


def bucket_by_index(objs: List[dict], replicas: int) -> Tuple[List[List[dict]], List[dict]]:
    """(slots[0..replicas), surplus): objects of one role by their index label; indices >= replicas are the surplus a
    scale-down leaves behind (the reference only logs them, pod.go:688-689).  Inside a slot the live, oldest object
    comes first."""
    slots: List[List[dict]] = [[] for _ in range(replicas)]
    surplus: List[dict] = []
    for o in objs:
        idx = replica_index(o)
        if idx is None or idx < 0:
            continue
        (surplus if idx >= replicas else slots[idx]).append(o)
    for sl in slots:
        sl.sort(key=lambda p: (p.get("metadata", {}).get("deletionTimestamp") is not None,
                               p.get("metadata", {}).get("creationTimestamp", "")))
    return slots, surplus


def of_role(objs: List[dict], rt_lower: str) -> List[dict]:
    return [o for o in objs if M.labels_of(o).get(C.LABEL_REPLICA_NAME) == rt_lower]


# ------------------------------------------------------------------------------------------ construction
def job_labels(job_name: str) -> Dict[str, str]:
    return {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: job_name.replace("/", "-")}


def owner_reference_of(job: AITrainingJob) -> dict:
    return {"apiVersion": C.API_VERSION, "kind": C.KIND, "name": job.name, "uid": job.uid,
            "blockOwnerDeletion": True, "controller": True}


def _contract_env(job: AITrainingJob, spec: ReplicaSpec, rt: str, index: str, restart_count: str,
                  ports: Dict[str, int], master_url: str) -> List[dict]:
    env: List[dict] = []

    def add(name: str, value) -> None:
        env.append({"name": name, "value": str(value)})

    ns = job.namespace
    for role, rspec in job.spec.replica_specs.items():
        declared = get_ports_from_job(job, role)
        n = int(rspec.replicas or 0)
        instances = [f"{gen_general_name(job.name, role.lower(), str(i))}.{ns}" for i in range(n)]
        hosts = [f"{inst}:{p}" for inst in instances for p in declared]
        up = role.upper()
        for suffix, items in (("INSTANCES", instances), ("PORTS", [str(p) for p in declared]), ("HOSTS", hosts)):
            add(f"{up}_{suffix}", ",".join(items))
            add(f"{up}_{suffix}_NUM", len(items))
        # loopback-resolvable form of <ROLE>_HOSTS (new)
        add(f"{up}_ADDRS", ",".join(f"127.0.0.1:{host_port(ports, role, i, p)}" for i in range(n) for p in declared))
    for name, value in ((C.ENV_REPLICA_NAME, rt), (C.ENV_REPLICA_INDEX, index),
                        (C.ENV_REPLICA_RESTARTCOUNT, restart_count),
                        (C.ENV_SERVICE, f"{gen_general_name(job.name, rt, index)}.{ns}"),
                        (C.ENV_JOB_NAME, job.name), (C.ENV_JOB_NAMESPACE, ns)):
        add(name, value)

    # rendezvous dialect selected by frameworkType (a dead field in the reference)
    role_key = next((r for r in job.spec.replica_specs if r.lower() == rt), rt)
    rdv = job.status.rendezvous
    world = int(spec.replicas or 0)
    if rdv is not None and role_key in rdv.world_sizes:
        world = rdv.world_sizes[role_key]
    fw = (job.spec.framework_type or "pytorch").lower()
    add("AITJ_JOB_UID", job.uid)
    add("AITJ_FRAMEWORK", fw)
    add("AITJ_RENDEZVOUS_GENERATION", rdv.generation if rdv else 0)
    add("AITJ_FAULT_TOLERANT", "1" if job.spec.fault_tolerant else "0")
    add("AITJ_EDL_POLICY", spec.edl_policy or C.EDL_POLICY_NEVER)
    if spec.min_replicas is not None:
        add("AITJ_MIN_REPLICAS", spec.min_replicas)
    if spec.max_replicas is not None:
        add("AITJ_MAX_REPLICAS", spec.max_replicas)
    if master_url:
        add("AITJ_MASTER", master_url)
    if fw in ("pytorch", "torch", ""):
        for name, value in (("RANK", index), ("WORLD_SIZE", world), ("LOCAL_RANK", 0), ("LOCAL_WORLD_SIZE", 1),
                            ("MASTER_ADDR", "127.0.0.1"),
                            ("MASTER_PORT", rdv.master_port if rdv and rdv.master_port else 29500)):
            add(name, value)
    elif fw in ("paddle", "paddlepaddle"):
        add("PADDLE_TRAINER_ID", index)
        add("PADDLE_TRAINERS_NUM", world)
    elif fw in ("tensorflow", "tf"):
        cluster = {role.lower(): [f"127.0.0.1:{host_port(ports, role, i, p)}" for i in range(int(rs.replicas or 0))
                                  for p in get_ports_from_job(job, role)[:1]]
                   for role, rs in job.spec.replica_specs.items()}
        add("TF_CONFIG", json.dumps({"cluster": cluster, "task": {"type": rt, "index": int(index)}}))
    return env


def build_pod_template(job: AITrainingJob, rt: str, index: str, restart_count: str, spec: ReplicaSpec,
                       ports: Optional[Dict[str, int]] = None, master_url: str = "") -> dict:
    """The pod template of replica (role ``rt``, ``index``): user template + identity labels + contract."""
    ports = ports if ports is not None else host_port_map(job)
    rdv = job.status.rendezvous
    labels = job_labels(job.name)
    labels.update({C.LABEL_JOBNAME_COMPAT: job.name, C.LABEL_POD_ROLE: rt, C.LABEL_RESTART_COUNT: restart_count,
                   C.LABEL_REPLICA_NAME: rt, C.LABEL_REPLICA_INDEX: index})
    if job.spec.priority:
        labels[C.LABEL_PRIORITY] = job.spec.priority
    if rdv is not None:
        labels[C.LABEL_GENERATION] = str(rdv.generation)
    tpl = M.deepcopy(spec.template)
    md = tpl.setdefault("metadata", {})
    md["name"] = gen_general_name(job.name, rt, index)
    md["generateName"] = gen_general_name(job.name, rt, "")
    merged = md.setdefault("labels", {})
    merged.update(labels)
    for k, v in job.labels.items():
        merged.setdefault(k, v)
    if rdv is not None and rdv.master_port:
        md.setdefault("annotations", {})[C.ANN_RENDEZVOUS_PORT] = str(rdv.master_port)
    pspec = tpl.setdefault("spec", {})
    if job.spec.scheduler_name:
        pspec["schedulerName"] = job.spec.scheduler_name
    if spec.restart_policy:
        pspec["restartPolicy"] = "Never"     # restarts are delete + re-create by the controller, never in place
    env = _contract_env(job, spec, rt, index, restart_count, ports, master_url)
    role_key = next((r for r in job.spec.replica_specs if r.lower() == rt), rt)
    for c in pspec.get("initContainers") or []:
        c["env"] = list(c.get("env") or []) + M.deepcopy(env)
    for c in pspec.get("containers") or []:
        declared = get_ports_from_container(c)
        c["env"] = list(c.get("env") or []) + M.deepcopy(env) + [
            {"name": C.ENV_PORTS, "value": ",".join(str(p) for p in declared)},
            {"name": "AITJ_HOST_PORTS",
             "value": ",".join(str(host_port(ports, role_key, int(index), p)) for p in declared)}]
    return tpl
