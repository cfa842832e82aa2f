"""Typed clientset: the hand-written counterpart of the reference's generated client.

Parity: /root/reference/pkg/client/clientset/versioned/clientset.go:27-97 (``Interface``,
``NewForConfig``), typed/aitrainingjob/v1/aitrainingjob_client.go:27-89 (group client) and
typed/aitrainingjob/v1/aitrainingjob.go:38-190 (Create / Update / UpdateStatus / Delete /
DeleteCollection / Get / List / Watch / Patch).  The same generic ``ResourceClient`` serves the
core kinds (pods, services, events, nodes, endpoints) and leases that the reference reaches through
``kubernetes.Interface`` (cmd/app/server.go:111-151).  ``AITrainingJobs`` returns typed
``api.types.AITrainingJob`` objects; the core kinds stay JSON dicts ("unstructured").
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from ..api import register as R
from ..api.types import AITrainingJob, AITrainingJobList
from ..store.transport import HTTPTransport, LocalTransport, Transport


class ResourceClient:
    """CRUD + watch for one kind in one namespace (``namespace=""`` = all namespaces for list/watch)."""

    def __init__(self, transport: Transport, info: R.ResourceInfo, namespace: str = ""):
        self._t = transport
        self.info = info
        self.namespace = namespace if info.namespaced else ""

    def _ns(self, obj: Optional[Dict[str, Any]] = None) -> str:
        if not self.info.namespaced:
            return ""
        if self.namespace:
            return self.namespace
        if obj is not None:
            return obj.get("metadata", {}).get("namespace", "") or "default"
        return "default"

    def create(self, obj: Dict[str, Any]) -> Dict[str, Any]:
        return self._t.create(self.info, self._ns(obj), obj)

    def get(self, name: str) -> Dict[str, Any]:
        return self._t.get(self.info, self._ns(), name)

    def list(self, label_selector: str = "", field_selector: str = "") -> Dict[str, Any]:
        return self._t.list(self.info, self.namespace, label_selector, field_selector)

    def update(self, obj: Dict[str, Any]) -> Dict[str, Any]:
        return self._t.update(self.info, self._ns(obj), obj["metadata"]["name"], obj)

    def update_status(self, obj: Dict[str, Any]) -> Dict[str, Any]:
        return self._t.update(self.info, self._ns(obj), obj["metadata"]["name"], obj, "status")

    def patch(self, name: str, patch: Any, patch_type: str = "application/merge-patch+json",
              subresource: str = "") -> Dict[str, Any]:
        return self._t.patch(self.info, self._ns(), name, patch, patch_type, subresource)

    def delete(self, name: str, grace_period_seconds: Optional[int] = None, uid: str = "") -> Dict[str, Any]:
        return self._t.delete(self.info, self._ns(), name, grace_period_seconds, uid)

    def delete_collection(self, label_selector: str = "", grace_period_seconds: Optional[int] = None):
        return self._t.delete_collection(self.info, self.namespace, label_selector, grace_period_seconds)

    def watch(self, resource_version: str = "", label_selector: str = "", timeout: Optional[float] = None):
        return self._t.watch(self.info, self.namespace, resource_version, label_selector, timeout)


class AITrainingJobInterface:
    """typed/aitrainingjob/v1/aitrainingjob.go:38-49 -- typed verbs over ``AITrainingJob``."""

    def __init__(self, transport: Transport, namespace: str):
        self._rc = ResourceClient(transport, R.AITRAININGJOB, namespace)

    @property
    def raw(self) -> ResourceClient:
        return self._rc

    def create(self, job: AITrainingJob) -> AITrainingJob:
        return AITrainingJob.from_dict(self._rc.create(job.to_dict()))

    def update(self, job: AITrainingJob) -> AITrainingJob:
        return AITrainingJob.from_dict(self._rc.update(job.to_dict()))

    def update_status(self, job: AITrainingJob) -> AITrainingJob:
        return AITrainingJob.from_dict(self._rc.update_status(job.to_dict()))

    def delete(self, name: str, grace_period_seconds: Optional[int] = None) -> None:
        self._rc.delete(name, grace_period_seconds)

    def delete_collection(self, label_selector: str = "") -> None:
        self._rc.delete_collection(label_selector)

    def get(self, name: str) -> AITrainingJob:
        return AITrainingJob.from_dict(self._rc.get(name))

    def list(self, label_selector: str = "") -> AITrainingJobList:
        return AITrainingJobList.from_dict(self._rc.list(label_selector))

    def watch(self, resource_version: str = "", label_selector: str = "", timeout: Optional[float] = None):
        for ev in self._rc.watch(resource_version, label_selector, timeout):
            yield {"type": ev["type"], "object": AITrainingJob.from_dict(ev["object"])}

    def patch(self, name: str, patch: Any, patch_type: str = "application/merge-patch+json",
              subresource: str = "") -> AITrainingJob:
        return AITrainingJob.from_dict(self._rc.patch(name, patch, patch_type, subresource))


class ElasticdeeplearningV1Client:
    """aitrainingjob_client.go:27-43."""

    def __init__(self, transport: Transport):
        self._t = transport

    def aitrainingjobs(self, namespace: str = "") -> AITrainingJobInterface:
        return AITrainingJobInterface(self._t, namespace)


class CoreV1Client:
    def __init__(self, transport: Transport):
        self._t = transport

    def pods(self, namespace: str = "") -> ResourceClient:
        return ResourceClient(self._t, R.POD, namespace)

    def services(self, namespace: str = "") -> ResourceClient:
        return ResourceClient(self._t, R.SERVICE, namespace)

    def events(self, namespace: str = "") -> ResourceClient:
        return ResourceClient(self._t, R.EVENT, namespace)

    def endpoints(self, namespace: str = "") -> ResourceClient:
        return ResourceClient(self._t, R.ENDPOINTS, namespace)

    def nodes(self) -> ResourceClient:
        return ResourceClient(self._t, R.NODE, "")


class CoordinationV1Client:
    def __init__(self, transport: Transport):
        self._t = transport

    def leases(self, namespace: str = "") -> ResourceClient:
        return ResourceClient(self._t, R.LEASE, namespace)


class ApiextensionsClient:
    def __init__(self, transport: Transport):
        self._t = transport

    def customresourcedefinitions(self) -> ResourceClient:
        return ResourceClient(self._t, R.CRD, "")


class Clientset:
    """clientset.go:27-32 ``Interface`` + the kube / apiextensions clients of server.go:111-151 in one."""

    def __init__(self, transport: Transport):
        self.transport = transport

    def elasticdeeplearning_v1(self) -> ElasticdeeplearningV1Client:
        return ElasticdeeplearningV1Client(self.transport)

    def core_v1(self) -> CoreV1Client:
        return CoreV1Client(self.transport)

    def coordination_v1(self) -> CoordinationV1Client:
        return CoordinationV1Client(self.transport)

    def apiextensions_v1beta1(self) -> ApiextensionsClient:
        return ApiextensionsClient(self.transport)

    def resource(self, info: R.ResourceInfo, namespace: str = "") -> ResourceClient:
        return ResourceClient(self.transport, info, namespace)


def new_for_config(master: str = "", server=None) -> Clientset:
    """clientset.go:61-78 ``NewForConfig``: an HTTP endpoint (``--master``) or an in-process server."""
    if server is not None:
        return Clientset(LocalTransport(server))
    return Clientset(HTTPTransport(master))
