"""Orphan sweep: replicas whose job is gone, or whose termination is overdue, are force-deleted.

What the reference specifies (/root/reference/pkg/controller/garbage_collection.go:28-106, every 10 minutes from
controller.go:204; ``--gc-interval`` here): among all pods labelled ``GroupName=elasticdeeplearning.ai``, a pod whose
``deletionTimestamp`` + grace period has passed is overdue, and a pod whose controlling ``AITrainingJob`` (by name *and*
uid) no longer exists is an orphan -- unless it is already terminating within its grace period on a node that is still
Ready, in which case the node agent is trusted to finish the job.

``sweep`` is a pure function over snapshots; ``GarbageCollector`` feeds it from the API and performs the grace-0
deletes.  On the single box a pod is a process group, so a force delete makes the agent SIGKILL it; the agent
additionally sweeps processes whose pod record vanished (``agent.agent``).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Callable, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..store.apiserver import APIError
from ..utils import klog, metrics

OVERDUE = "terminated expired"
ORPHAN = "owner job does not exist"


@dataclass(frozen=True)
class Garbage:
    namespace: str
    name: str
    reason: str


def sweep(pods: List[dict], owner_uid: Callable[[str, str], Optional[str]], node_ready: Callable[[str], bool],
          now) -> List[Garbage]:
    """``owner_uid(namespace, job_name)`` -> uid of the live job, ``""`` if there is none, ``None`` if unknown (lookup
    failed: leave the pod alone); ``node_ready(name)`` -> whether the pod's node can still be trusted to terminate it."""
    out: List[Garbage] = []
    for pod in pods:
        if M.labels_of(pod).get(C.LABEL_GROUP_NAME) != C.GROUP_NAME:
            continue
        md = pod.get("metadata", {})
        dying_since = M.parse_time(md.get("deletionTimestamp"))
        if dying_since is not None and \
                (now - dying_since).total_seconds() >= (md.get("deletionGracePeriodSeconds") or 0):
            out.append(Garbage(M.namespace_of(pod), M.name_of(pod), OVERDUE))
            continue
        ref = M.get_controller_of(pod)
        if ref is None or ref.get("kind") != C.KIND or ref.get("apiVersion") != C.API_VERSION:
            continue
        live = owner_uid(M.namespace_of(pod), ref.get("name", ""))
        if live is None or live == ref.get("uid"):
            continue
        node = pod.get("spec", {}).get("nodeName")
        if dying_since is not None and (not node or node_ready(node)):
            continue          # terminating within its grace period where somebody is looking after it
        out.append(Garbage(M.namespace_of(pod), M.name_of(pod), ORPHAN))
    return out


class GarbageCollector:
    def __init__(self, kube_client, trainingjob_lister):
        self.kube_cli = kube_client
        self.trainingjob_lister = trainingjob_lister
        self.deleted = 0

    def clean_orphans(self, period: float, stop: threading.Event) -> None:
        while not stop.wait(period):
            try:
                self.clean_garbage_pods()
            except Exception as e:  # noqa: BLE001
                klog.error("garbage collection pass failed: %r", e)

    def _owner_uid(self, namespace: str, name: str) -> Optional[str]:
        try:
            return self.trainingjob_lister.aitrainingjobs(namespace).get(name).uid
        except APIError as e:
            return "" if e.reason == "NotFound" else None

    def _node_ready(self, name: str) -> bool:
        try:
            node = self.kube_cli.core_v1().nodes().get(name)
        except APIError:
            return True       # cannot tell: do not escalate
        return any(c.get("type") == "Ready" and c.get("status") == "True"
                   for c in node.get("status", {}).get("conditions") or [])

    def clean_garbage_pods(self) -> int:
        try:
            pods = self.kube_cli.core_v1().pods("").list().get("items", [])
        except APIError as e:
            klog.error("garbage collection: cannot list pods: %s", e.message)
            return 0
        n = 0
        for g in sweep(pods, self._owner_uid, self._node_ready, M.now()):
            klog.info("garbage collection: force-deleting pod %s/%s (%s)", g.namespace, g.name, g.reason)
            try:
                self.kube_cli.core_v1().pods(g.namespace).delete(g.name, grace_period_seconds=0)
                n += 1
            except APIError as e:
                if e.reason != "NotFound":
                    klog.error("garbage collection: deleting %s/%s failed: %s", g.namespace, g.name, e.message)
        self.deleted += n
        if n:
            metrics.inc("aitj_gc_deleted_pods_total", n)
        return n
