"""Scheme: the kinds the store serves and how they map onto REST paths.

Parity: /root/reference/pkg/apis/aitrainingjob/v1/register.go:27-67 (group, version, kind, plural,
short name, AddToScheme), pkg/apis/aitrainingjob/register.go:1-22 and the CRD object the controller
registers for itself at start-up (pkg/controller/controller.go:210-234: namespaced, no schema,
no status subresource, no printer columns; AlreadyExists tolerated).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

from . import constants as C


@dataclass(frozen=True)
class ResourceInfo:
    group: str          # "" = core
    version: str
    kind: str
    plural: str
    namespaced: bool = True
    short_names: tuple = ()

    @property
    def api_version(self) -> str:
        return f"{self.group}/{self.version}" if self.group else self.version

    def path(self, namespace: str = "", name: str = "") -> str:
        root = f"/apis/{self.group}/{self.version}" if self.group else f"/api/{self.version}"
        p = root
        if self.namespaced and namespace:
            p += f"/namespaces/{namespace}"
        p += f"/{self.plural}"
        if name:
            p += f"/{name}"
        return p


AITRAININGJOB = ResourceInfo(C.GROUP_NAME, C.GROUP_VERSION, C.KIND, C.KIND_PLURAL, True, (C.SHORT_NAME,))
POD = ResourceInfo("", "v1", "Pod", "pods", True, ("po",))
SERVICE = ResourceInfo("", "v1", "Service", "services", True, ("svc",))
EVENT = ResourceInfo("", "v1", "Event", "events", True, ("ev",))
NODE = ResourceInfo("", "v1", "Node", "nodes", False, ("no",))
ENDPOINTS = ResourceInfo("", "v1", "Endpoints", "endpoints", True, ("ep",))
NAMESPACE = ResourceInfo("", "v1", "Namespace", "namespaces", False, ("ns",))
LEASE = ResourceInfo("coordination.k8s.io", "v1", "Lease", "leases", True, ())
CRD = ResourceInfo("apiextensions.k8s.io", "v1beta1", "CustomResourceDefinition", "customresourcedefinitions",
                   False, ("crd", "crds"))

_BY_KIND: Dict[str, ResourceInfo] = {}
_BY_PLURAL: Dict[str, ResourceInfo] = {}


def add_to_scheme(info: ResourceInfo) -> None:
    _BY_KIND[info.kind] = info
    _BY_PLURAL[info.plural] = info
    _BY_PLURAL[info.kind.lower()] = info
    for s in info.short_names:
        _BY_PLURAL[s] = info


for _r in (AITRAININGJOB, POD, SERVICE, EVENT, NODE, ENDPOINTS, NAMESPACE, LEASE, CRD):
    add_to_scheme(_r)


def by_kind(kind: str) -> ResourceInfo:
    return _BY_KIND[kind]


def lookup(name: str) -> Optional[ResourceInfo]:
    """Resolve a kubectl-style resource argument (plural, singular, kind or short name)."""
    n = name.lower()
    if n in _BY_PLURAL:
        return _BY_PLURAL[n]
    if n.endswith("s") and n[:-1] in _BY_PLURAL:
        return _BY_PLURAL[n[:-1]]
    return _BY_KIND.get(name)


def all_resources():
    return list(_BY_KIND.values())


def crd_object() -> dict:
    """The CustomResourceDefinition the controller self-registers (controller.go:210-234)."""
    return {
        "apiVersion": CRD.api_version,
        "kind": CRD.kind,
        "metadata": {"name": C.crd_name()},
        "spec": {
            "group": C.GROUP_NAME,
            "version": C.GROUP_VERSION,
            "scope": "Namespaced",
            "names": {"kind": C.KIND, "plural": C.KIND_PLURAL, "shortNames": [C.SHORT_NAME]},
        },
    }
