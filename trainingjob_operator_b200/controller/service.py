"""Service reconciler: one headless Service per (role, index) = a stable address + port reservation.

Parity: /root/reference/pkg/controller/service.go:19-240 -- port discovery from ``aitj-`` containers /
``aitj-`` ports (service.go:19-52), add handler with expectations (service.go:54-81; update/delete are
empty stubs there, service.go:83-88), claim (service.go:90-115), create missing services only for
roles that have an ``aitj-`` container (service.go:117-146), headless ``clusterIP: None`` with the
job/replica/index selector (service.go:148-196), filters and index bucketing (service.go:198-240).
On a single box the Service also records the loopback ``hostPort`` each declared port is remapped to
("loopback addresses + reserved ports").
"""
from __future__ import annotations

from typing import Dict, List

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaSpec
from ..store.apiserver import APIError
from ..utils import klog
from .control import ControllerRefManager, recheck_deletion_timestamp
from .pod import gen_general_name, get_ports_from_job, host_port


def gen_expectation_services_key(job_key: str, rt: str) -> str:
    return f"{job_key}/{rt.lower()}/services"


def has_container_port(job: AITrainingJob, rtype: str) -> bool:
    """service.go:45-52 (name notwithstanding it checks for an ``aitj-`` *container*)."""
    return any(str(c.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX)
               for c in job.spec.replica_specs[rtype].containers())


class ServiceReconciler:
    """Mixed into ``TrainingJobController``."""

    def add_service(self, svc: dict) -> None:
        if svc.get("metadata", {}).get("deletionTimestamp"):
            return
        ref = M.get_controller_of(svc)
        if ref is None:
            return
        job = self.resolve_controller_ref(M.namespace_of(svc), ref)
        if job is None:
            return
        rt = M.labels_of(svc).get(C.LABEL_REPLICA_NAME)
        if rt is None:
            klog.info("This service may not created by %s", C.CONTROLLER_NAME)
            return
        self.expectations.creation_observed(gen_expectation_services_key(job.key(), rt))
        self.work_queue.add(job.key())

    def update_service(self, old: dict, new: dict) -> None:  # service.go:83-85: no-op
        return

    def delete_service(self, obj) -> None:  # service.go:86-88: no-op
        return

    def get_services_by_job_and_selector(self, job: AITrainingJob, selector: Dict[str, str]) -> List[dict]:
        from .controller import claim_candidates

        return self.claim_services(job, selector, claim_candidates(self.service_lister, job, selector))

    def claim_services(self, job: AITrainingJob, selector: Dict[str, str], services: List[dict]) -> List[dict]:
        def fresh():
            f = self.trainingjob_lister.aitrainingjobs(job.namespace).get(job.name)
            if f.uid != job.uid:
                raise RuntimeError(f"original {C.KIND} {job.namespace}/{job.name} is gone")
            return f

        mgr = ControllerRefManager(self.service_control.patch_service, job, selector,
                                   recheck_deletion_timestamp(fresh))
        return mgr.claim(services)

    def reconcile_services(self, job: AITrainingJob, services: List[dict], rtype: str) -> None:
        ports = get_ports_from_job(job, rtype)
        rt = rtype.lower()
        spec = job.spec.replica_specs[rtype]
        replicas = int(spec.replicas or 0)
        mine = self.filter_services_for_replica_type(services, rt)
        slices, surplus = self.get_service_slices(mine, replicas)
        klog.V(4).info("job %s type %s ports: %s", job.name, rtype, ports)
        for svc in surplus:  # scale-down: drop the address of a removed replica
            try:
                self.service_control.delete_service(M.namespace_of(svc), M.name_of(svc), job)
            except APIError as e:
                klog.warning("delete surplus service %s: %s", M.name_of(svc), e.message)
        if not has_container_port(job, rtype):
            return
        for index, sl in enumerate(slices):
            if not sl:
                klog.info("need to create new service: %s-%d", rt, index)
                self.create_new_service(job, rtype, str(index), spec, ports)

    def create_new_service(self, job: AITrainingJob, rtype: str, index: str, spec: ReplicaSpec,
                           ports: List[int]) -> None:
        rt = rtype.lower()
        key = gen_expectation_services_key(job.key(), rt)
        self.expectations.raise_expectations(key, 1, 0)
        labels = self.gen_labels(job.name)
        labels[C.LABEL_REPLICA_NAME] = rt
        labels[C.LABEL_REPLICA_INDEX] = index
        service = {
            "apiVersion": "v1", "kind": "Service",
            "metadata": {"name": gen_general_name(job.name, rt, index), "labels": dict(labels)},
            "spec": {
                "clusterIP": "None",
                "selector": dict(labels),
                "ports": [{"name": f"{C.DEFAULT_PORT_PREFIX}{p}", "port": p,
                           "hostPort": host_port(job, rtype, int(index), p)} for p in ports],
            },
        }
        try:
            self.service_control.create_services_with_controller_ref(job.namespace, service, job,
                                                                     self.gen_owner_reference(job))
        except APIError as e:
            self.expectations.creation_observed(key)
            if e.reason in ("AlreadyExists", "Timeout"):
                return
            raise

    @staticmethod
    def filter_services_for_replica_type(services: List[dict], rt: str) -> List[dict]:
        return [s for s in services if M.labels_of(s).get(C.LABEL_REPLICA_NAME) == rt]

    @staticmethod
    def get_service_slices(services: List[dict], replicas: int):
        slices: List[List[dict]] = [[] for _ in range(replicas)]
        surplus: List[dict] = []
        for svc in services:
            raw = M.labels_of(svc).get(C.LABEL_REPLICA_INDEX)
            if raw is None:
                klog.warning("The service do not have the index label.")
                continue
            try:
                idx = int(raw)
            except ValueError:
                continue
            if idx < 0:
                continue
            if idx >= replicas:
                surplus.append(svc)
            else:
                slices[idx].append(svc)
        return slices, surplus
