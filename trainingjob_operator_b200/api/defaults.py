"""Defaulting of ``AITrainingJob`` objects.

Parity: /root/reference/pkg/apis/aitrainingjob/v1/defaults.go:15-53 (replicas=1,
restartPolicy=Never, restartScope=All, replica failPolicy=Any / completePolicy=All; job
cleanPodPolicy=All, failPolicy=Any, completePolicy=All) and zz_generated.defaults.go:29-44.
Like the reference (controller.go:297) defaults are applied at reconcile time and persisted
with the first status write -- but on a copy, never on the informer cache (fixes quirk Q6).
"""
from __future__ import annotations

from . import constants as C
from .types import AITrainingJob, ReplicaSpec


def set_defaults_replica_spec(spec: ReplicaSpec) -> None:
    if spec.replicas is None:
        spec.replicas = 1
    if not spec.restart_policy:
        spec.restart_policy = C.RESTART_POLICY_NEVER
    if not spec.restart_scope:
        spec.restart_scope = C.RESTART_SCOPE_ALL
    if not spec.fail_policy:
        spec.fail_policy = C.ENDING_POLICY_ANY
    if not spec.complete_policy:
        spec.complete_policy = C.ENDING_POLICY_ALL


def set_defaults_aitrainingjob(job: AITrainingJob) -> AITrainingJob:
    if job.spec.clean_pod_policy is None:
        job.spec.clean_pod_policy = C.CLEAN_POD_POLICY_ALL
    if not job.spec.fail_policy:
        job.spec.fail_policy = C.ENDING_POLICY_ANY
    if not job.spec.complete_policy:
        job.spec.complete_policy = C.ENDING_POLICY_ALL
    for spec in job.spec.replica_specs.values():
        set_defaults_replica_spec(spec)
    return job
