"""Per-replica Service objects: a stable address plus the record of the loopback ports reserved for the replica.

The reference creates one headless Service (``clusterIP: None``) per (role, index) whose selector is the replica's
identity labels, only for roles that have an ``aitj-`` container (/root/reference/pkg/controller/service.go:117-196;
port discovery :19-52).  On one box the Service additionally carries ``hostPort`` for every declared port: the free
loopback port the job was given for that (role, index, port) -- see ``controller.pod.host_port`` -- so anything that
resolves ``<job>-<role>-<index>.<ns>:<port>`` can find ``127.0.0.1:<hostPort>``.

Pure helpers only; which Services to create or drop is decided in ``controller.engine``.
"""
from __future__ import annotations

from typing import Dict, List

from ..api import constants as C
from ..api.types import AITrainingJob
from .pod import gen_general_name, host_port, job_labels


def gen_expectation_services_key(job_key: str, rt: str) -> str:
    return f"{job_key}/{rt.lower()}/services"


def has_contract_container(job: AITrainingJob, rtype: str) -> bool:
    return any(str(c.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX)
               for c in job.spec.replica_specs[rtype].containers())


def build_service(job: AITrainingJob, rtype: str, index: int, declared_ports: List[int],
                  ports: Dict[str, int]) -> dict:
    rt = rtype.lower()
    identity = job_labels(job.name)
    identity[C.LABEL_REPLICA_NAME] = rt
    identity[C.LABEL_REPLICA_INDEX] = str(index)
    return {
        "apiVersion": "v1", "kind": "Service",
        "metadata": {"name": gen_general_name(job.name, rt, str(index)), "labels": dict(identity)},
        "spec": {"clusterIP": "None", "selector": dict(identity),
                 "ports": [{"name": f"{C.DEFAULT_PORT_PREFIX}{p}", "port": p,
                            "hostPort": host_port(ports, rtype, index, p)} for p in declared_ports]},
    }
