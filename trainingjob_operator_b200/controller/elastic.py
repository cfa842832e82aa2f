"""Elastic rescale: real semantics for ``minReplicas`` / ``maxReplicas`` / ``edlPolicy``.

In the reference these fields are API surface only (/root/reference/pkg/apis/aitrainingjob/v1/replica.go:10-11,19 --
never read by controller code): scale-up just creates pods whose environment disagrees
with the running ones (pod.go:186-193, env fixed at creation pod.go:528) and scale-down is unimplemented
(pod.go:688-689).  Here a change of the desired world size bumps a *rendezvous generation* kept in
``status.rendezvous``:

* every replica is created for one generation and gets ``AITJ_RENDEZVOUS_GENERATION``, ``WORLD_SIZE`` and a
  generation-specific ``MASTER_PORT``;
* running workers watch ``status.rendezvous`` (``runtime.elastic``) and re-rendezvous at a step boundary -- survivors
  keep their step state on the device and broadcast it to joiners over NVLink;
* ranks whose index falls out of range leave voluntarily and are then deleted;
* ``edlPolicy: Auto`` lets the controller pick ``replicas`` within [min, max] from the healthy free GPU slots;
  ``Manual`` honours user edits within the bounds; ``Never`` freezes the size the job started with.
A restart also bumps the generation so re-created replicas rendezvous on a fresh port.

The functions here are pure: what they need to know about the rest of the box arrives as a ``ClusterView`` the
controller computes from its informer caches, and a new ``MASTER_PORT`` is taken from the ports the caller hands in.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from ..api import constants as C
from ..core import _aitj_core as _core
from ..api import meta as M
from ..api.types import AITrainingJob, Rendezvous
from ..utils import metrics
from .pod import of_role, pod_node, replica_index, scheduling_message

metrics.describe("aitj_rendezvous_generations_total", "rendezvous generation bumps (scale up/down, restart)")

AUTO_RECHECK_SECONDS = 2.0
LABEL_NODE_TYPE = "aitj.b200/type"
LABEL_GPU_INDEX = "aitj.b200/gpu-index"

_PORT_LOCK = threading.Lock()
_RECENT_PORTS: List[int] = []
_RECENT_SET: set = set()


def allocate_ports(n: int) -> List[int]:
    """``n`` distinct free loopback TCP ports.  The probe is one native call (``core.free_loopback_ports``: all probe
    sockets are bound to port 0 before any is closed, so the kernel cannot hand a port out twice within a call); the
    last few hundred results are remembered so that two jobs do not race to one port between the probe and the worker's
    own bind.  One call per reconcile pass: a job's MASTER_PORT and all of its per-replica host ports at once."""
    out: List[int] = []
    with _PORT_LOCK:
        for _ in range(4):
            for port in _core.free_loopback_ports(n - len(out)):
                if port not in _RECENT_SET:
                    out.append(port)
                    _RECENT_SET.add(port)
                    _RECENT_PORTS.append(port)
            if len(out) >= n:
                break
        while len(_RECENT_PORTS) > 512:
            _RECENT_SET.discard(_RECENT_PORTS.pop(0))
        if len(out) < n:          # every probe landed on a remembered port four times over: take what the kernel offers
            out += _core.free_loopback_ports(n - len(out))
    if len(out) < n:
        raise OSError(f"could not find {n} free loopback ports")
    return out


def allocate_port() -> int:
    """A free loopback TCP port (see ``allocate_ports``)."""
    return allocate_ports(1)[0]


@dataclass(frozen=True)
class ClusterView:
    """The part of the box an ``edlPolicy: Auto`` decision depends on."""
    free_gpu_slots: int = 0
    waiting_higher: int = 0       # GPUs wanted by unschedulable pods of strictly more important jobs
    waiting_at_least: int = 0     # ... of jobs at least as important as this one


def observe_cluster(job: AITrainingJob, nodes: List[dict], all_pods: List[dict]) -> ClusterView:
    """Free healthy GPU slots and the unplaced demand of other jobs, from informer-cache snapshots.  GPU nodes are
    recognised by their ``aitj.b200/type=gpu`` label and pods' ``aitj.b200/gpus`` indices are mapped through the
    nodes' ``aitj.b200/gpu-index`` label (node names carry the agent's ``--node-prefix``)."""
    ready_gpu: Dict[str, str] = {}       # node name -> gpu index
    for n in nodes:
        labels = M.labels_of(n)
        is_gpu = labels.get(LABEL_NODE_TYPE) == "gpu" or \
            (LABEL_NODE_TYPE not in labels and "gpu-" in M.name_of(n))
        if is_gpu and any(c.get("type") == "Ready" and c.get("status") == "True"
                          for c in n.get("status", {}).get("conditions") or []):
            ready_gpu[M.name_of(n)] = labels.get(LABEL_GPU_INDEX, M.name_of(n).rsplit("-", 1)[-1])
    by_index = {idx: name for name, idx in ready_gpu.items()}
    busy = set()
    mine = M.priority_value(job.spec.priority)
    higher = at_least = 0
    for pod in all_pods:
        node = pod_node(pod)
        if node:
            if pod.get("status", {}).get("phase") in (C.POD_PENDING, C.POD_RUNNING, None):
                busy.add(node)
                for g in (M.annotations_of(pod).get(C.ANN_GPUS) or "").split(","):
                    if g.strip() in by_index:
                        busy.add(by_index[g.strip()])
            continue
        if pod.get("metadata", {}).get("deletionTimestamp"):
            continue
        if M.labels_of(pod).get(C.LABEL_JOB_NAME) == job.name and M.namespace_of(pod) == job.namespace:
            continue
        if not scheduling_message(pod):
            continue                      # not (yet) found unschedulable
        want = M.pod_gpu_request(pod)
        if want <= 0:
            continue
        prio = M.priority_value(M.labels_of(pod).get(C.LABEL_PRIORITY, ""))
        higher += want if prio > mine else 0
        at_least += want if prio >= mine else 0
    return ClusterView(len(set(ready_gpu) - busy), higher, at_least)


def desired_world_sizes(job: AITrainingJob) -> Dict[str, int]:
    return {rt: int(spec.replicas or 0) for rt, spec in job.spec.replica_specs.items()}


def clamp_replicas(spec, value: int) -> int:
    lo = spec.min_replicas if spec.min_replicas is not None else 0
    hi = spec.max_replicas if spec.max_replicas is not None else max(value, lo)
    return max(lo, min(hi, value))


def auto_roles(job: AITrainingJob) -> List[str]:
    return [rt for rt, s in job.spec.replica_specs.items()
            if s.edl_policy == C.EDL_POLICY_AUTO and (s.min_replicas is not None or s.max_replicas is not None)]


def plan_autoscale(job: AITrainingJob, pods: List[dict], view: ClusterView) -> Tuple[Dict[str, dict], List[str]]:
    """``edlPolicy: Auto``: (spec patch per role, log lines).  A role shrinks to what fits when replicas of its own are
    unschedulable, yields slots to strictly more important work that cannot be placed, and grows into free slots
    nobody at its priority or above is waiting for -- always within [minReplicas, maxReplicas], and never while a
    previous shrink is still draining (its slots may be all the waiting replicas need)."""
    patch: Dict[str, dict] = {}
    notes: List[str] = []
    free = view.free_gpu_slots
    for rt in auto_roles(job):
        spec = job.spec.replica_specs[rt]
        cur = int(spec.replicas or 0)
        mine = of_role(pods, rt.lower())
        stuck = [p for p in mine if not pod_node(p) and scheduling_message(p)]
        draining = [p for p in mine if (replica_index(p) or 0) >= cur]
        if stuck and draining:
            continue
        if stuck:
            target = cur - len(stuck)
        elif view.waiting_higher > free and not draining:
            target = cur - (view.waiting_higher - free)
        elif free > 0 and view.waiting_at_least == 0 and len(mine) == cur and \
                all(p.get("status", {}).get("phase") == C.POD_RUNNING for p in mine):
            target = cur + free
        else:
            target = cur
        target = clamp_replicas(spec, target)
        if target != cur:
            notes.append(f"job {job.key()} role {rt}: edlPolicy Auto: replicas {cur} -> {target} "
                         f"(free slots {free}, unschedulable {len(stuck)})")
            patch[rt] = {"replicas": target}
            free = max(0, free - max(0, target - cur))
    return patch, notes


def frozen_roles(job: AITrainingJob) -> set:
    """Roles with ``edlPolicy: Never`` keep the world size they started running with."""
    rdv = job.status.rendezvous
    if rdv is None or not job.status.start_running_time:
        return set()
    return {rt for rt, spec in job.spec.replica_specs.items()
            if spec.edl_policy == C.EDL_POLICY_NEVER and rt in rdv.world_sizes}


def rendezvous_target(job: AITrainingJob) -> Dict[str, int]:
    """World size per role the current spec asks for, frozen roles excepted."""
    want = desired_world_sizes(job)
    rdv = job.status.rendezvous
    if rdv is None:
        return want
    frozen = frozen_roles(job)
    return {rt: (rdv.world_sizes.get(rt, n) if rt in frozen else n) for rt, n in want.items()}


def next_generation(job: AITrainingJob, port: int, now, world_sizes: Optional[Dict[str, int]] = None) -> Rendezvous:
    """Advance ``status.rendezvous`` in place: generation + 1 on ``port``."""
    rdv = job.status.rendezvous
    if rdv is None:
        rdv = job.status.rendezvous = Rendezvous(generation=0, world_sizes=desired_world_sizes(job))
    rdv.generation += 1
    rdv.master_port = int(port)
    rdv.changed_at = M.format_time(now)
    if world_sizes is not None:
        rdv.world_sizes = dict(world_sizes)
    return rdv
