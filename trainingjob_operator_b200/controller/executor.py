"""Applies an engine ``Decision``: the only part of the reconcile path that talks to the API server.

Owns what the reference spreads over its reconcilers:

* the expectations cache (upstream ``ControllerExpectations``): every create / delete is announced
  before it is issued and taken back when the call fails, so the next pass of the job waits until the informers have
  seen what this pass did (quirk Q3: accumulate instead of overwrite, deletions are expected too);
* pod / service creation through ``PodControl`` / ``ServiceControl`` (Events ``SuccessfulCreatePod`` ... as upstream);
  missing replicas of one pass are created concurrently when a create is a network round trip (the reference issues
  one synchronous POST per pod behind a 5 qps client throttle, /root/reference/pkg/controller/pod.go:186-193);
* the status write-back (/root/reference/pkg/controller/status.go:285-305 re-PUTs the whole object five times): here a
  conflict re-reads the *live* object and carries over the status plus only the annotations this pass set itself --
  workers patch their own annotations onto the job (ready-r<rank>, metrics, rescale-trace) and those writes are
  exactly what causes the conflict, so they must never be rolled back to the controller's stale view (quirk Q7).
"""
from __future__ import annotations

import concurrent.futures as cf
import time
from typing import Dict, List, Optional

from ..api import meta as M
from ..api.defaults import set_defaults_aitrainingjob
from ..api.types import AITrainingJob
from ..store.apiserver import APIError
from ..utils import klog, metrics
from .engine import Decision, PodCreate, PodDelete
from .pod import gen_expectation_pods_key, owner_reference_of
from .service import gen_expectation_services_key

_LOG = {"info": klog.info, "warning": klog.warning, "error": klog.error}


class Executor:
    def __init__(self, kube_client, trainingjob_client, pod_control, service_control, expectations, work_queue):
        self.kube_client = kube_client
        self.trainingjob_client = trainingjob_client
        self.pod_control = pod_control
        self.service_control = service_control
        self.expectations = expectations
        self.work_queue = work_queue
        self._pool: Optional[cf.ThreadPoolExecutor] = None
        # concurrency only pays when a create is a network round trip (separate API server process); against the
        # in-process store creates are sub-millisecond pure CPU and threads would just queue on the GIL
        self.remote = getattr(getattr(kube_client, "transport", None), "master", None) is not None

    # ------------------------------------------------------------------------------------ whole decision
    def apply(self, job: AITrainingJob, d: Decision) -> Optional[AITrainingJob]:
        """Returns the job as written when the status changed.  An API error of a create / delete propagates after
        the remaining actions of its kind were attempted, and then nothing is written: the pass is retried."""
        for level, line in d.log:
            _LOG.get(level, klog.info)("%s", line)
        for name, labels in d.counters:
            metrics.inc(name, labels=labels or None)
        for name, value in d.observations:
            metrics.observe(name, value)
        if d.spec_patch is not None:
            try:
                self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace).patch(job.name,
                                                                                                    d.spec_patch)
            except APIError as e:
                klog.warning("auto-scale patch of %s failed: %s", job.key(), e.message)
            return None
        for ns, name, patch in d.pod_patches:
            try:
                self.pod_control.patch_pod(ns, name, patch)
            except APIError as e:
                klog.warning("cannot patch %s: %s", name, e.message)
        self.delete_pods(job, d.pod_deletes)
        self.create_pods(job, d.pod_creates)
        for ns, name in d.service_deletes:
            try:
                self.service_control.delete_service(ns, name, job)
            except APIError as e:
                klog.warning("delete service %s failed: %s", name, e.message)
        for rt, svc in d.service_creates:
            self.create_service(job, rt, svc)
        written = self.write_status(job, set(d.annotations)) if d.write_status else None
        key = job.key()
        if d.requeue_rate_limited:
            self.work_queue.add_rate_limited(key)
        for delay in d.requeue_after:
            self.work_queue.add_after(key, float(delay))
        return written

    # ------------------------------------------------------------------------------------ pods
    def delete_pods(self, job: AITrainingJob, deletes: List[PodDelete]) -> None:
        if not deletes:
            return
        per_role: Dict[str, int] = {}
        for x in deletes:
            per_role[x.role] = per_role.get(x.role, 0) + 1
        for rt, n in per_role.items():
            self.expectations.raise_expectations(gen_expectation_pods_key(job.key(), rt), 0, n)
        for x in deletes:
            try:
                self.pod_control.delete_pod(x.namespace, x.name, job, grace_period_seconds=x.grace)
            except APIError as e:
                self.expectations.deletion_observed(gen_expectation_pods_key(job.key(), x.role))
                klog.error("delete pod %s failed: %s", x.name, e.message)

    def create_pods(self, job: AITrainingJob, creates: List[PodCreate]) -> None:
        if not creates:
            return
        per_role: Dict[str, int] = {}
        for c in creates:
            per_role[c.role] = per_role.get(c.role, 0) + 1
        for rt, n in per_role.items():
            self.expectations.raise_expectations(gen_expectation_pods_key(job.key(), rt), n, 0)
        ref = owner_reference_of(job)
        errors: List[Exception] = []

        def one(c: PodCreate) -> None:
            t0 = time.perf_counter()
            try:
                self.pod_control.create_pods_with_controller_ref(job.namespace, c.template, job, ref)
            except APIError as e:
                self.expectations.creation_observed(gen_expectation_pods_key(job.key(), c.role))
                # AlreadyExists: the informer is behind our own create.  An owner that no longer exists: the job was
                # deleted while this pass ran on a cached copy (the API server refuses to create the orphan) -- nothing
                # to retry, the deletion event ends the job's reconciliation.
                gone = e.reason == "NotFound" and "refusing to create an orphan" in (e.message or "")
                if e.reason != "AlreadyExists" and not gone:
                    errors.append(e)
                return
            metrics.observe("aitj_pod_create_seconds", time.perf_counter() - t0)

        if len(creates) == 1 or not self.remote:
            for c in creates:
                one(c)
        else:
            if self._pool is None:        # one long-lived pool per controller, not one per pass
                self._pool = cf.ThreadPoolExecutor(max_workers=16, thread_name_prefix="pod-create")
            list(self._pool.map(one, creates))
        if errors:
            raise errors[0]

    def create_service(self, job: AITrainingJob, rt: str, service: dict) -> None:
        key = gen_expectation_services_key(job.key(), rt)
        self.expectations.raise_expectations(key, 1, 0)
        try:
            self.service_control.create_services_with_controller_ref(job.namespace, service, job,
                                                                     owner_reference_of(job))
        except APIError as e:
            self.expectations.creation_observed(key)
            if e.reason not in ("AlreadyExists", "Timeout"):
                raise

    # ------------------------------------------------------------------------------------ status
    def write_status(self, job: AITrainingJob, owned_annotations: set, attempts: int = 5) -> AITrainingJob:
        """Persist status (+ the annotations in ``owned_annotations``, + defaulted spec) under optimistic concurrency."""
        client = self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace)
        last: Optional[Exception] = None
        cur = job
        for _ in range(attempts):
            try:
                return client.update(cur)
            except APIError as e:
                if e.reason == "NotFound":
                    raise
                last = e
                klog.V(2).info("status write of %s (%s) failed: %s", job.key(), job.status.phase, e.message)
            try:
                fresh = client.get(job.name)
            except APIError as e:
                if e.reason == "NotFound":
                    raise
                last = e
                continue
            if fresh.uid != job.uid:
                raise APIError(409, "Conflict", f"job {job.key()} was recreated (uid changed)")
            fresh.status = job.status
            mine = {k: v for k, v in job.annotations.items() if k in owned_annotations}
            if mine:
                fresh.metadata["annotations"] = dict(fresh.annotations, **mine)
            set_defaults_aitrainingjob(fresh)
            cur = fresh
        raise last if last else RuntimeError("status update failed")
