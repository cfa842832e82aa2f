"""Orphan garbage collector.

Parity: /root/reference/pkg/controller/garbage_collection.go:20-106 -- on a timer
(controller.go:204: 10 minutes, configurable here with ``--gc-interval``) list every pod, and for those
labelled ``GroupName=elasticdeeplearning.ai``: force-delete when ``deletionTimestamp`` has already
passed (:48-52), and force-delete when the owning ``AITrainingJob`` is no longer in the lister (:54-72)
unless it is still within its grace period on a Ready node (:62-65, ``checkNode`` :91-106).
On the single box a pod is a process group, so a force delete makes the agent SIGKILL it; the agent
additionally sweeps processes whose pod record vanished (``agent.kubelet``).
"""
from __future__ import annotations

import threading

from ..api import constants as C
from ..api import meta as M
from ..store.apiserver import APIError
from ..utils import klog, metrics


class GarbageCollector:
    def __init__(self, kube_client, trainingjob_lister):
        self.kube_cli = kube_client
        self.trainingjob_lister = trainingjob_lister
        self.deleted = 0

    def clean_orphans(self, period: float, stop: threading.Event) -> None:
        while not stop.wait(period):
            klog.V(4).info("Garbage collector working now ...")
            try:
                self.clean_garbage_pods()
            except Exception as e:  # noqa: BLE001
                klog.error("garbage collection pass failed: %r", e)

    def clean_garbage_pods(self) -> int:
        try:
            pods = self.kube_cli.core_v1().pods("").list().get("items", [])
        except APIError:
            klog.error("List garbage pod failed")
            return 0
        n = 0
        now = M.now()
        for pod in pods:
            if M.labels_of(pod).get(C.LABEL_GROUP_NAME) != C.GROUP_NAME:
                continue
            md = pod.get("metadata", {})
            dts = M.parse_time(md.get("deletionTimestamp"))
            grace = md.get("deletionGracePeriodSeconds") or 0
            expired = dts is not None and (now - dts).total_seconds() >= grace
            if expired:
                klog.error("Find garbage pod %s, reason: terminated expired", M.name_of(pod))
                n += self._delete(pod)
                continue
            ref = M.get_controller_of(pod)
            if ref is None or ref.get("kind") != C.KIND or ref.get("apiVersion") != C.API_VERSION:
                continue
            try:
                owner = self.trainingjob_lister.aitrainingjobs(M.namespace_of(pod)).get(ref.get("name", ""))
                if owner.uid == ref.get("uid"):
                    continue
            except APIError as e:
                if e.reason != "NotFound":
                    continue
            if dts is not None and not expired and self.check_node(pod):
                klog.V(4).info("Find pod %s to delete but deletion timestamp is %s, waiting", M.name_of(pod),
                               md.get("deletionTimestamp"))
                continue
            klog.info("Find pod %s, whose owner Job %s not existed", M.name_of(pod), ref.get("name"))
            n += self._delete(pod)
        self.deleted += n
        if n:
            metrics.inc("aitj_gc_deleted_pods_total", n)
        return n

    def _delete(self, pod: dict) -> int:
        try:
            self.kube_cli.core_v1().pods(M.namespace_of(pod)).delete(M.name_of(pod), grace_period_seconds=0)
            return 1
        except APIError as e:
            if e.reason != "NotFound":
                klog.error("Delete pod %s/%s failed, reason: %s", M.namespace_of(pod), M.name_of(pod), e.message)
            return 0

    def check_node(self, pod: dict) -> bool:
        node_name = pod.get("spec", {}).get("nodeName")
        if not node_name:
            return True
        try:
            node = self.kube_cli.core_v1().nodes().get(node_name)
        except APIError:
            klog.error("check node %s with pod %s failed!!", node_name, M.name_of(pod))
            return True
        return any(c.get("type") == "Ready" and c.get("status") == "True"
                   for c in node.get("status", {}).get("conditions") or [])
