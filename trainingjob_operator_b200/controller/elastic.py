"""Elastic rescale: real semantics for ``minReplicas`` / ``maxReplicas`` / ``edlPolicy``.

In the reference these fields are API surface only (/root/reference/pkg/apis/aitrainingjob/v1/
replica.go:10-11,19 -- never read by any controller code): scale-up just
creates pods whose environment disagrees with the already running ones (pod.go:186-193, env fixed at
creation pod.go:528) and scale-down is unimplemented (pod.go:688-689).  Here a change of the desired
world size bumps a *rendezvous generation* kept in ``status.rendezvous``:

* every replica is created for one generation and gets ``AITJ_RENDEZVOUS_GENERATION``,
  ``WORLD_SIZE`` and a generation-specific ``MASTER_PORT``;
* running workers watch ``status.rendezvous`` (``runtime.elastic``) and, at a step boundary,
  tear their process group down and re-initialise with the new world size -- survivors keep their
  step state on the device and broadcast it to joiners over NVLink;
* ranks whose index falls out of range leave voluntarily and are then deleted (``pod.py``);
* ``edlPolicy: Auto`` lets the controller pick ``replicas`` within [min, max] from the number of
  healthy, free GPU slots; ``Manual`` honours user edits within the bounds; ``Never`` freezes the size
  the job started with.
A restart also bumps the generation so re-created replicas rendezvous on a fresh port.
"""
from __future__ import annotations

import socket
import threading
from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, Rendezvous
from ..store.apiserver import APIError
from ..utils import klog, metrics

metrics.describe("aitj_rendezvous_generations_total", "rendezvous generation bumps (scale up/down, restart)")

AUTO_RECHECK_SECONDS = 2.0

_PORT_LOCK = threading.Lock()
_RECENT_PORTS: List[int] = []


def allocate_port() -> int:
    """A free loopback TCP port (bind to 0, remember the last few so two jobs do not race to one)."""
    with _PORT_LOCK:
        for _ in range(32):
            s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            try:
                s.bind(("127.0.0.1", 0))
                port = s.getsockname()[1]
            finally:
                s.close()
            if port not in _RECENT_PORTS:
                _RECENT_PORTS.append(port)
                del _RECENT_PORTS[:-64]
                return port
        return port


def _index_of(pod: dict) -> int:
    try:
        return int(M.labels_of(pod).get(C.LABEL_REPLICA_INDEX, "0"))
    except ValueError:
        return 0


def desired_world_sizes(job: AITrainingJob) -> Dict[str, int]:
    return {rt: int(spec.replicas or 0) for rt, spec in job.spec.replica_specs.items()}


def clamp_replicas(spec, value: int) -> int:
    lo = spec.min_replicas if spec.min_replicas is not None else 0
    hi = spec.max_replicas if spec.max_replicas is not None else max(value, lo)
    return max(lo, min(hi, value))


class ElasticMixin:
    """Mixed into ``TrainingJobController``."""

    def reconcile_rendezvous(self, job: AITrainingJob, pods: List[dict]) -> None:
        want = desired_world_sizes(job)
        rdv = job.status.rendezvous
        if rdv is None:
            job.status.rendezvous = Rendezvous(generation=1, world_sizes=want, master_port=allocate_port(),
                                               changed_at=M.format_time())
            return
        frozen = self._frozen_roles(job)
        effective = {rt: (rdv.world_sizes.get(rt, n) if rt in frozen else n) for rt, n in want.items()}
        if effective != rdv.world_sizes:
            # only a running/starting job needs a new generation; before any pod exists just adopt the sizes
            if pods:
                self.bump_rendezvous(job, "rescale", effective)
                self.trace_rescale(job)
            else:
                rdv.world_sizes = effective

    def _frozen_roles(self, job: AITrainingJob) -> set:
        """Roles with ``edlPolicy: Never`` keep the world size they started with."""
        rdv = job.status.rendezvous
        if rdv is None:
            return set()
        return {rt for rt, spec in job.spec.replica_specs.items()
                if spec.edl_policy == C.EDL_POLICY_NEVER and rt in rdv.world_sizes and job.status.start_running_time}

    def bump_rendezvous(self, job: AITrainingJob, why: str, world_sizes: Optional[Dict[str, int]] = None) -> None:
        rdv = job.status.rendezvous
        if rdv is None:
            rdv = job.status.rendezvous = Rendezvous(generation=0, world_sizes=desired_world_sizes(job))
        rdv.generation += 1
        rdv.master_port = allocate_port()
        rdv.changed_at = M.format_time()
        if world_sizes is not None:
            rdv.world_sizes = dict(world_sizes)
        klog.info("job %s: rendezvous generation %d (%s): world %s port %d", job.key(), rdv.generation, why,
                  rdv.world_sizes, rdv.master_port)
        metrics.inc("aitj_rendezvous_generations_total", labels={"reason": why})

    def trace_rescale(self, job: AITrainingJob) -> None:
        import json
        import time

        raw = job.annotations.get(C.ANN_TRACE)
        try:
            tr = json.loads(raw) if raw else {}
        except ValueError:
            tr = {}
        tr.setdefault("rescales", []).append({"generation": job.status.rendezvous.generation,
                                              "at": round(time.time(), 4),
                                              "world": dict(job.status.rendezvous.world_sizes)})
        tr["rescales"] = tr["rescales"][-16:]
        job.set_annotation(C.ANN_TRACE, json.dumps(tr, sort_keys=True))

    # ------------------------------------------------------------------ edlPolicy: Auto
    def reconcile_elastic(self, job: AITrainingJob, pods: List[dict]) -> bool:
        """For ``edlPolicy: Auto`` roles choose replicas in [min,max] from free healthy GPU slots.
        Returns True when the job spec was patched (the caller stops; the update event re-queues)."""
        # not while replicas are being torn down for a restart (barrier pending).  Once they are re-created the job sits in
        # Restarting until all of them run again -- if one cannot be placed because its GPU is gone, shrinking to what
        # fits is exactly what gets the job out of that state (growth needs every replica Running, so it cannot fire)
        if job.status.phase == C.PHASE_TERMINATING or job.status.restart_replica_name:
            return False
        patch: Dict[str, dict] = {}
        ready = None
        for rt, spec in job.spec.replica_specs.items():
            if spec.edl_policy != C.EDL_POLICY_AUTO:
                continue
            if spec.min_replicas is None and spec.max_replicas is None:
                continue
            if ready is None:
                ready = self._free_gpu_slots(job)
            cur = int(spec.replicas or 0)
            mine = [p for p in pods if M.labels_of(p).get(C.LABEL_REPLICA_NAME) == rt.lower()]
            unschedulable = [p for p in mine if not p.get("spec", {}).get("nodeName")
                             and self.get_pod_scheduling_message(p)]
            target = cur
            surplus = [p for p in mine if _index_of(p) >= cur]
            if unschedulable and surplus:
                continue      # a previous shrink is still draining: its slots may be all the waiting replicas need
            higher, at_least = self._waiting_gpu_demand(job)
            if unschedulable:
                target = cur - len(unschedulable)          # shrink to what fits
            elif higher > ready and not surplus:
                target = cur - (higher - ready)            # yield to more important work that cannot be placed
            elif ready > 0 and at_least == 0 and len(mine) == cur and \
                    all(p.get("status", {}).get("phase") == C.POD_RUNNING for p in mine):
                target = cur + ready                       # grow into free slots nobody at our priority or above waits for
            target = clamp_replicas(spec, target)
            if target != cur:
                klog.info("job %s role %s: edlPolicy Auto: replicas %d -> %d (free slots %d, unschedulable %d)",
                          job.key(), rt, cur, target, ready, len(unschedulable))
                patch[rt] = {"replicas": target}
                ready = max(0, ready - max(0, target - cur))
        if not patch:
            if any(s.edl_policy == C.EDL_POLICY_AUTO and (s.min_replicas is not None or s.max_replicas is not None)
                   for s in job.spec.replica_specs.values()):
                # slots freed by other jobs, or more important pods that cannot be placed, raise no event on this job:
                # look again in a while
                self.work_queue.add_after(job.key(), AUTO_RECHECK_SECONDS)
            return False
        try:
            self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace).patch(
                job.name, {"spec": {"replicaSpecs": patch}})
        except APIError as e:
            klog.warning("auto-scale patch of %s failed: %s", job.key(), e.message)
            return False
        metrics.inc("aitj_autoscale_total")
        return True

    def _waiting_gpu_demand(self, job: AITrainingJob) -> tuple:
        """GPUs asked for by other jobs' pods that the scheduler could not place: (by strictly more important pods, by pods
        at least as important as this job).  ``spec.priority`` is copied onto every pod as the ``priority`` label
        (pod.go:503-505) and the node agent's scheduler already orders by it; an ``edlPolicy: Auto`` role additionally
        gives slots back (down to ``minReplicas``) when that ordering alone cannot help."""
        mine = M.priority_value(job.spec.priority)
        higher = at_least = 0
        for pod in self.pod_lister.peek():
            if pod.get("spec", {}).get("nodeName") or pod.get("metadata", {}).get("deletionTimestamp"):
                continue
            if M.labels_of(pod).get(C.LABEL_JOB_NAME) == job.name and M.namespace_of(pod) == job.namespace:
                continue
            if not self.get_pod_scheduling_message(pod):
                continue                    # not (yet) found unschedulable
            want = M.pod_gpu_request(pod)
            if want <= 0:
                continue
            prio = M.priority_value(M.labels_of(pod).get(C.LABEL_PRIORITY, ""))
            if prio > mine:
                higher += want
            if prio >= mine:
                at_least += want
        return higher, at_least

    def _free_gpu_slots(self, job: AITrainingJob) -> int:
        ready_nodes = {n for n in self.get_node_status() if n.startswith("gpu-")}
        busy = set()
        for pod in self.pod_lister.list():
            node = pod.get("spec", {}).get("nodeName")
            if node and pod.get("status", {}).get("phase") in (C.POD_PENDING, C.POD_RUNNING, None):
                busy.add(node)
                for g in (M.annotations_of(pod).get(C.ANN_GPUS) or "").split(","):
                    if g.strip():
                        busy.add(f"gpu-{g.strip()}")
        return len(ready_nodes - busy)
