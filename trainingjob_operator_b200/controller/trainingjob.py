"""Job event handlers, naming and teardown.

Parity: /root/reference/pkg/controller/trainingjob.go:12-73 -- ``<job>-<role>-<index>`` naming
(:12-15, implemented in ``pod.gen_general_name``), add -> enqueue now (:17-24), update with a new
resourceVersion -> enqueue rate-limited + delayed enqueue when ``timeLimit`` changed (:26-47), delete
-> enqueue (:49-51), delete pods then services (:53-73).  The reference's two
``// FIXME: need to validate trainingjob`` (:21,:33) are resolved: invalid objects are rejected at
admission and, defensively, skipped here with a warning.  Quirk Q9 (services leak when there are no
pods) is fixed: services are deleted regardless.
"""
from __future__ import annotations

from typing import List

from ..api import meta as M
from ..api.types import AITrainingJob
from ..api.validation import validate_dict
from ..store.apiserver import APIError
from ..utils import klog


class TrainingJobHandlers:
    """Mixed into ``TrainingJobController``."""

    def add_training_job(self, obj: dict) -> None:
        klog.V(2).info("Informer: Add TrainingJob %s/%s.", M.namespace_of(obj), M.name_of(obj))
        errs = validate_dict(obj)
        if errs:
            klog.warning("TrainingJob %s/%s is invalid, not reconciling: %s", M.namespace_of(obj), M.name_of(obj),
                         "; ".join(errs))
            return
        self.enqueue_job(obj, False, 0)

    def update_training_job(self, old: dict, cur: dict) -> None:
        if M.resource_version(old) == M.resource_version(cur):
            klog.V(4).info("Same Resourceversion for training job %s/%s, skipped", M.namespace_of(old),
                           M.name_of(old))
            return
        errs = validate_dict(cur)
        if errs:
            klog.warning("TrainingJob %s/%s is invalid, not reconciling: %s", M.namespace_of(cur), M.name_of(cur),
                         "; ".join(errs))
            return
        klog.V(2).info("Informer: Update TrainingJob %s/%s.", M.namespace_of(old), M.name_of(old))
        spec_changed = old.get("spec") != cur.get("spec") or \
            M.annotations_of(old) != M.annotations_of(cur)
        # a spec/annotation edit by the user is acted on immediately; our own status writes are rate limited
        self.enqueue_job(cur, not spec_changed, 0)
        new_limit = cur.get("spec", {}).get("timeLimit")
        old_limit = old.get("spec", {}).get("timeLimit")
        started = cur.get("status", {}).get("startRunningTime")
        if started and new_limit is not None and new_limit != old_limit:
            remaining = int(new_limit) - int(M.seconds_since(started))
            self.enqueue_job(cur, False, max(remaining, 0) + 0.05)
            klog.info("job TimeLimit updated, will rsync after %d seconds", remaining)

    def delete_training_job(self, obj) -> None:
        self.enqueue_job(obj, False, 0)

    def delete_pods_and_services(self, job: AITrainingJob, pods: List[dict], services: List[dict]) -> None:
        live = [p for p in pods if not p.get("metadata", {}).get("deletionTimestamp")]
        if live:
            self.delete_pods_expecting(job, live, None)
        for svc in services:
            try:
                self.service_control.delete_service(M.namespace_of(svc), M.name_of(svc), job)
            except APIError as e:
                klog.warning("delete service %s failed: %s", M.name_of(svc), e.message)
