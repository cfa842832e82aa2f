"""Scheduler half of the node agent: binds Pending pods to free healthy GPU slots, highest ``priority`` label first
(/root/reference/pkg/controller/pod.go:503-505), honours ``schedulerName`` (pod.go:524-526) by ignoring pods addressed to a
foreign scheduler, and writes the ``PodScheduled=False`` condition whose message the controller surfaces (pod.go:457-467)
when nothing fits.  The reference leaves all of this to kube-scheduler.  Mixed into ``NodeAgent``.
"""
from __future__ import annotations

import time
from typing import List

from ..api import constants as C
from ..api import meta as M
from ..store.apiserver import APIError
from ..utils import klog

GPU_RESOURCE = M.GPU_RESOURCE
OWN_SCHEDULERS = ("", "default-scheduler", "aitj-scheduler")
pod_gpu_request = M.pod_gpu_request


def pod_priority(pod: dict) -> int:
    return M.priority_value(M.labels_of(pod).get(C.LABEL_PRIORITY, ""))


class SchedulerMixin:
    def _free_gpus(self) -> List[int]:
        ready = set()
        for n in self.node_lister.peek():
            nm = M.name_of(n)
            if not nm.startswith(f"{self.prefix}gpu-"):
                continue
            if any(c.get("type") == "Ready" and c.get("status") == "True"
                   for c in n.get("status", {}).get("conditions") or []):
                ready.add(int(nm.rsplit("-", 1)[1]))
        busy = set()
        live_uids = set()
        for p in self.pod_lister.peek():
            done = (p.get("status", {}).get("phase") or C.POD_PENDING) in (C.POD_SUCCEEDED, C.POD_FAILED)
            if not done:
                live_uids.add(M.uid_of(p))
            if not p.get("spec", {}).get("nodeName") or done:
                continue
            for g in (M.annotations_of(p).get(C.ANN_GPUS) or "").split(","):
                if g.strip():
                    busy.add(int(g))
        with self._lock:
            # drop allocations whose pod is finished or gone (seen by the cache); keep binds the cache has not
            # caught up with yet (younger than 1 s)
            for g, (uid, at) in list(self._gpu_owner.items()):
                if uid not in live_uids and time.monotonic() - at > 1.0:
                    del self._gpu_owner[g]
            busy |= set(self._gpu_owner)
        return sorted(ready - busy)

    def schedule(self, pod: dict) -> None:
        if (pod.get("spec", {}).get("schedulerName") or "") not in OWN_SCHEDULERS:
            return
        want = pod_gpu_request(pod)
        ns, name = M.namespace_of(pod), M.name_of(pod)
        with self._lock:
            if M.uid_of(pod) in self._bound:
                return                      # already bound by us; the informer just has not caught up
            if len(self._bound) > 4096:
                cutoff = time.monotonic() - 60.0
                self._bound = {u: t for u, t in self._bound.items() if t > cutoff}
        if want == 0:
            self._bind(pod, self.cpu_node, [])
            return
        # higher-priority pending pods go first: yield if someone more important is waiting
        mine = (pod_priority(pod), )
        for other in self.pod_lister.peek():
            if other.get("spec", {}).get("nodeName") or M.uid_of(other) == M.uid_of(pod):
                continue
            if other.get("metadata", {}).get("deletionTimestamp") or pod_gpu_request(other) == 0:
                continue
            if (pod_priority(other),) > mine:
                self.queue.add_after(M.key_of(pod), 0.05)
                free = self._free_gpus()
                if len(free) < pod_gpu_request(other) + want:
                    self._mark_unschedulable(pod, f"0/{self.num_gpus} nodes are available: waiting for "
                                             f"higher-priority pod {M.name_of(other)}.")
                    return
        with self._lock:
            free = self._free_gpus()
            if len(free) < want:
                total = self.num_gpus
                self._mark_unschedulable(pod, f"0/{total} nodes are available: {total - len(free)} Insufficient "
                                         f"{GPU_RESOURCE}, {len(free)} free but {want} requested.")
                return
            gpus = free[:want]
            if want == 1:
                # rank i prefers GPU slot i when it is free: replicas are created in parallel, so arrival order is not
                # rank order, and a stable rank -> GPU mapping is what the pinned warm pool and a human reading
                # nvidia-smi both expect
                try:
                    pref = int(M.labels_of(pod).get(C.LABEL_REPLICA_INDEX, "")) % max(1, self.num_gpus)
                    if pref in free:
                        gpus = [pref]
                except ValueError:
                    pass
            for g in gpus:
                self._gpu_owner[g] = (M.uid_of(pod), time.monotonic())
            try:
                self._bind(pod, self.gpu_node(gpus[0]), gpus)
            except APIError:
                for g in gpus:
                    self._gpu_owner.pop(g, None)
                raise

    def _release_gpus(self, uid: str, kick: bool = True) -> None:
        with self._lock:
            for g in [g for g, (u, _t) in self._gpu_owner.items() if u == uid]:
                del self._gpu_owner[g]
            # `_bound` keeps the uid (pruned by age in `schedule`): a pod that ran to completion before the informer cache
            # even showed it as bound must not be bound a second time by a stale queue entry
        if kick:
            self._kick_pending()

    def _mark_unschedulable(self, pod: dict, message: str) -> None:
        self.queue.add_after(M.key_of(pod), 1.0)   # safety net: retry even if no event announces a free GPU
        with self._lock:
            self._unschedulable.add(M.key_of(pod))
        conds = pod.get("status", {}).get("conditions") or []
        cur = M.condition(conds, "PodScheduled")
        if cur is not None and cur.get("status") == "False" and cur.get("message") == message:
            return
        patch = {"status": {"phase": C.POD_PENDING, "conditions": [
            {"type": "PodScheduled", "status": "False", "reason": "Unschedulable", "message": message,
             "lastTransitionTime": M.format_time()}]}}
        self.cs.core_v1().pods(M.namespace_of(pod)).patch(M.name_of(pod), patch, subresource="status")

    def _bind(self, pod: dict, node: str, gpus: List[int]) -> None:
        patch = {"spec": {"nodeName": node},
                 "metadata": {"annotations": {C.ANN_GPUS: ",".join(str(g) for g in gpus)}},
                 "status": {"phase": C.POD_PENDING, "hostIP": "127.0.0.1", "podIP": "127.0.0.1",
                            "conditions": [{"type": "PodScheduled", "status": "True",
                                            "lastTransitionTime": M.format_time()}]}}
        with self._lock:
            self._bound[M.uid_of(pod)] = time.monotonic()
        try:
            bound = self.cs.core_v1().pods(M.namespace_of(pod)).patch(M.name_of(pod), patch)
        except APIError:
            with self._lock:
                self._bound.pop(M.uid_of(pod), None)
            raise
        klog.V(2).info("scheduled %s -> %s gpus=%s", M.key_of(pod), node, gpus)
        # scheduler and kubelet are one process here: start the containers from the object the bind returned instead of
        # waiting for it to come back through the informer (one watch round trip per replica on the submit -> Running path)
        if isinstance(bound, dict) and bound.get("spec", {}).get("nodeName") == node and \
                not bound.get("metadata", {}).get("deletionTimestamp"):
            try:
                self._start_pod(bound, M.key_of(pod))
                return
            except APIError as e:
                klog.V(2).info("agent: direct start of %s failed (%s), retrying through the queue", M.key_of(pod), e.message)
        self.queue.add(M.key_of(pod))

    def _kick_pending(self) -> None:
        """A slot was freed: retry the pods that were found unschedulable (they also retry on their own once a second;
        pods that were never looked at yet are in the queue already)."""
        with self._lock:
            keys, self._unschedulable = list(self._unschedulable), set()
        for key in keys:
            self.queue.add(key)
