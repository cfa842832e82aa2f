"""``AITrainingJob`` API types (group ``elasticdeeplearning.ai/v1``).

Parity: /root/reference/pkg/apis/aitrainingjob/v1/types.go:29-152 (AITrainingJob, spec, status,
conditions, list) and replica.go:9-63 (ReplicaSpec, ReplicaStatus, policy enums).  JSON key
spellings are byte-compatible, including the odd ones (``RestartCount`` / ``RestartReplicaName``
capitalised, types.go:84-86; phase ``Succeed``).  ``deepcopy`` stands in for the generated
zz_generated.deepcopy.go:27-258.

New (behind reference fields that were dead code there): ``status.rendezvous``
carries the elastic re-rendezvous generation, and ``lastReconcileTime`` is actually written.
"""
from __future__ import annotations

import copy
import dataclasses
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from ..core import _aitj_core as _core
from . import constants as C


def _j(name: str, omitempty: bool = True, **kw):
    return field(metadata={"json": name, "omitempty": omitempty}, **kw)


def _jcopy(v):
    """Private copy of a JSON-shaped value (the native tree copy; copy.deepcopy for anything else)."""
    return _core.jcopy(v, copy.deepcopy)


_FIELD_TABLES: Dict[type, tuple] = {}


def _table(cls) -> tuple:
    """(fields, known json names) of a dataclass, computed once: (attribute, json key, converter, omitempty, omitzero)."""
    t = _FIELD_TABLES.get(cls)
    if t is None:
        rows = tuple((f.name, f.metadata.get("json", f.name), f.metadata.get("conv"), f.metadata.get("omitempty", True),
                      f.metadata.get("omitzero", True)) for f in dataclasses.fields(cls))
        t = _FIELD_TABLES[cls] = (rows, frozenset(r[1] for r in rows))
    return t


class _Serde:
    """dataclass <-> JSON dict with Go-style ``omitempty``."""

    def to_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for attr, name, _conv, omitempty, omitzero in _table(type(self))[0]:
            v = getattr(self, attr)
            if omitempty and (v is None or v == "" or v == {} or v == [] or v is False
                              or (isinstance(v, int) and not isinstance(v, bool) and v == 0 and omitzero)):
                continue
            out[name] = _dump(v)
        return out

    @classmethod
    def from_dict(cls, d: Optional[Dict[str, Any]]):
        d = d or {}
        rows, known = _table(cls)
        kwargs = {}
        for attr, name, conv, _oe, _oz in rows:
            v = d.get(name)
            if v is None:
                continue
            kwargs[attr] = conv(v) if conv else _jcopy(v)
        obj = cls(**kwargs)
        if len(d) > len(kwargs):
            extra = {k: _jcopy(v) for k, v in d.items() if k not in known}
            if extra:
                object.__setattr__(obj, "_extra", extra)
        return obj

    def deepcopy(self):
        return copy.deepcopy(self)


def _dump(v):
    if isinstance(v, _Serde):
        d = v.to_dict()
        extra = getattr(v, "_extra", None)
        if extra:
            for k, x in extra.items():
                d.setdefault(k, _jcopy(x))
        return d
    if isinstance(v, dict):
        return {k: _dump(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_dump(x) for x in v]
    return v


@dataclass
class ReplicaSpec(_Serde):
    """replica.go:9-20."""
    min_replicas: Optional[int] = field(default=None, metadata={"json": "minReplicas", "omitempty": True, "omitzero": False})
    max_replicas: Optional[int] = field(default=None, metadata={"json": "maxReplicas", "omitempty": True, "omitzero": False})
    replicas: Optional[int] = field(default=None, metadata={"json": "replicas", "omitempty": True, "omitzero": False})
    restart_limit: Optional[int] = field(default=None, metadata={"json": "restartLimit", "omitempty": True, "omitzero": False})
    template: Dict[str, Any] = field(default_factory=dict, metadata={"json": "template", "omitempty": False})
    restart_policy: str = _j("restartPolicy", default="")
    restart_scope: str = _j("restartScope", default="")
    fail_policy: str = _j("failPolicy", default="")
    complete_policy: str = _j("completePolicy", default="")
    edl_policy: str = _j("edlPolicy", default="")

    # -- template helpers -------------------------------------------------------------------
    def pod_spec(self) -> Dict[str, Any]:
        return self.template.setdefault("spec", {})

    def containers(self) -> List[Dict[str, Any]]:
        return self.pod_spec().get("containers") or []

    def init_containers(self) -> List[Dict[str, Any]]:
        return self.pod_spec().get("initContainers") or []


@dataclass
class ReplicaStatus(_Serde):
    """replica.go:36-49 (all counters ``omitempty``: zeros vanish from the JSON)."""
    pending: int = _j("pending", default=0)
    scheduled: int = _j("scheduled", default=0)
    active: int = _j("active", default=0)
    succeeded: int = _j("succeeded", default=0)
    restarting: int = _j("restarting", default=0)
    failed: int = _j("failed", default=0)

    def total(self) -> int:
        return self.pending + self.scheduled + self.active + self.succeeded + self.restarting + self.failed


@dataclass
class TrainingJobCondition(_Serde):
    """types.go:128-142."""
    type: str = _j("type", omitempty=False, default="")
    status: str = _j("status", omitempty=False, default="True")
    reason: str = _j("reason", default="")
    message: str = _j("message", default="")
    last_probe_time: str = _j("lastProbeTime", default="")
    last_transition_time: str = _j("lastTransitionTime", default="")


@dataclass
class Rendezvous(_Serde):
    """Elastic re-rendezvous record (new; gives minReplicas/maxReplicas/edlPolicy real semantics)."""
    generation: int = field(default=0, metadata={"json": "generation", "omitempty": False})
    world_sizes: Dict[str, int] = field(default_factory=dict, metadata={"json": "worldSizes", "omitempty": False})
    master_port: int = field(default=0, metadata={"json": "masterPort", "omitempty": False})
    changed_at: str = _j("changedAt", default="")


def _conv_map(cls):
    return lambda d: {k: cls.from_dict(v) for k, v in (d or {}).items()}


def _conv_list(cls):
    return lambda xs: [cls.from_dict(x) for x in (xs or [])]


@dataclass
class TrainingJobStatus(_Serde):
    """types.go:76-95."""
    phase: str = _j("phase", omitempty=False, default="")
    conditions: List[TrainingJobCondition] = field(
        default_factory=list, metadata={"json": "conditions", "omitempty": False, "conv": _conv_list(TrainingJobCondition)})
    replica_statuses: Dict[str, ReplicaStatus] = field(
        default_factory=dict, metadata={"json": "replicaStatuses", "omitempty": False, "conv": _conv_map(ReplicaStatus)})
    restart_counts: Dict[str, int] = _j("RestartCount", default_factory=dict)
    restart_replica_name: str = _j("RestartReplicaName", omitempty=False, default="")
    start_time: Optional[str] = _j("startTime", default=None)
    start_running_time: Optional[str] = _j("startRunningTime", default=None)
    end_time: Optional[str] = _j("endTime", default=None)
    last_reconcile_time: Optional[str] = _j("lastReconcileTime", default=None)
    rendezvous: Optional[Rendezvous] = field(
        default=None, metadata={"json": "rendezvous", "omitempty": True, "conv": Rendezvous.from_dict})

    def get_condition(self, ctype: str) -> Optional[TrainingJobCondition]:
        for c in self.conditions:
            if c.type == ctype:
                return c
        return None


@dataclass
class TrainingJobSpec(_Serde):
    """types.go:41-62."""
    restarting_exit_code: str = _j("restartingExitCode", default="")
    framework_type: str = _j("frameworkType", default="")
    fault_tolerant: bool = _j("faultTolerant", default=False)
    priority: str = _j("priority", default="")
    scheduler_name: str = _j("schedulerName", default="")
    time_limit: Optional[int] = field(default=None, metadata={"json": "timeLimit", "omitempty": True, "omitzero": False})
    clean_pod_policy: Optional[str] = _j("cleanPodPolicy", default=None)
    fail_policy: str = _j("failPolicy", default="")
    complete_policy: str = _j("completePolicy", default="")
    replica_specs: Dict[str, ReplicaSpec] = field(
        default_factory=dict, metadata={"json": "replicaSpecs", "omitempty": False, "conv": _conv_map(ReplicaSpec)})


@dataclass
class AITrainingJob(_Serde):
    """types.go:29-38."""
    api_version: str = _j("apiVersion", omitempty=False, default=C.API_VERSION)
    kind: str = _j("kind", omitempty=False, default=C.KIND)
    metadata: Dict[str, Any] = field(default_factory=dict, metadata={"json": "metadata", "omitempty": False})
    spec: TrainingJobSpec = field(default_factory=TrainingJobSpec,
                                  metadata={"json": "spec", "omitempty": False, "conv": TrainingJobSpec.from_dict})
    status: TrainingJobStatus = field(default_factory=TrainingJobStatus,
                                      metadata={"json": "status", "omitempty": False, "conv": TrainingJobStatus.from_dict})

    # -- metadata accessors --------------------------------------------------------------------
    @property
    def name(self) -> str:
        return self.metadata.get("name", "")

    @property
    def namespace(self) -> str:
        return self.metadata.get("namespace", "")

    @property
    def uid(self) -> str:
        return self.metadata.get("uid", "")

    @property
    def resource_version(self) -> str:
        return str(self.metadata.get("resourceVersion", ""))

    @property
    def labels(self) -> Dict[str, str]:
        return self.metadata.get("labels") or {}

    @property
    def annotations(self) -> Dict[str, str]:
        return self.metadata.get("annotations") or {}

    def set_annotation(self, k: str, v: str) -> None:
        self.metadata.setdefault("annotations", {})[k] = v

    @property
    def deletion_timestamp(self) -> Optional[str]:
        return self.metadata.get("deletionTimestamp")

    def key(self) -> str:
        return f"{self.namespace}/{self.name}"


@dataclass
class AITrainingJobList(_Serde):
    """types.go:147-152."""
    api_version: str = _j("apiVersion", omitempty=False, default=C.API_VERSION)
    kind: str = _j("kind", omitempty=False, default=C.KIND_LIST)
    metadata: Dict[str, Any] = field(default_factory=dict, metadata={"json": "metadata", "omitempty": False})
    items: List[AITrainingJob] = field(default_factory=list,
                                       metadata={"json": "items", "omitempty": False, "conv": _conv_list(AITrainingJob)})
