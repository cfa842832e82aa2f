"""Pod / Service control (create, delete, patch + Events) and controller-ref managers.

The reference delegates these to upstream helpers that are *not* in its tree
(``job_controller.RealPodControl`` / ``RealServiceControl`` at
/root/reference/pkg/controller/controller.go:94-102, ``NewPodControllerRefManager(...).ClaimPods``
at pod.go:148-149, ``NewServiceControllerRefManager`` at service.go:113-114); SURVEY.md §2.2
re-specifies them and this module implements that specification:

* ``create_*_with_controller_ref``: object from template + ownerRef, Event
  ``Normal SuccessfulCreatePod "Created pod: <name>"`` / ``Warning FailedCreatePod`` on the owner;
* ``delete_*``: Events ``SuccessfulDeletePod`` / ``FailedDeletePod``;
* claim: owned & selector matches -> keep; owned but labels no longer match -> release; orphan &
  matches & owner not being deleted (live re-check) -> adopt; owned by someone else -> ignore.

``FakePodControl`` / ``FakeServiceControl`` record what would have been done (the analogue of
upstream's fakes that the reference's interfaces were designed for).
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, List, Optional

from ..api import meta as M
from ..client.record import EVENT_NORMAL, EVENT_WARNING, EventRecorder
from ..store.apiserver import APIError
from ..utils import klog


def _with_owner(obj: Dict[str, Any], controller_ref: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    o = M.deepcopy(obj)
    if controller_ref is not None:
        o.setdefault("metadata", {}).setdefault("ownerReferences", []).append(M.deepcopy(controller_ref))
    return o


class RealPodControl:
    def __init__(self, clientset, recorder: EventRecorder):
        self._cs = clientset
        self._rec = recorder

    def create_pods_with_controller_ref(self, namespace: str, template: Dict[str, Any], owner,
                                        controller_ref: Dict[str, Any]) -> Dict[str, Any]:
        pod = {"apiVersion": "v1", "kind": "Pod",
               "metadata": M.deepcopy(template.get("metadata") or {}),
               "spec": M.deepcopy(template.get("spec") or {})}
        pod = _with_owner(pod, controller_ref)
        pod["metadata"]["namespace"] = namespace
        try:
            created = self._cs.core_v1().pods(namespace).create(pod)
        except APIError as e:
            self._rec.eventf(owner, EVENT_WARNING, "FailedCreatePod", "Error creating: %s", e.message)
            raise
        self._rec.eventf(owner, EVENT_NORMAL, "SuccessfulCreatePod", "Created pod: %s", M.name_of(created))
        return created

    def delete_pod(self, namespace: str, name: str, owner, grace_period_seconds: Optional[int] = None) -> None:
        try:
            self._cs.core_v1().pods(namespace).delete(name, grace_period_seconds)
        except APIError as e:
            if e.reason == "NotFound":
                return
            self._rec.eventf(owner, EVENT_WARNING, "FailedDeletePod", "Error deleting: %s", e.message)
            raise
        self._rec.eventf(owner, EVENT_NORMAL, "SuccessfulDeletePod", "Deleted pod: %s", name)

    def patch_pod(self, namespace: str, name: str, patch: Dict[str, Any]) -> Dict[str, Any]:
        return self._cs.core_v1().pods(namespace).patch(name, patch)


class RealServiceControl:
    def __init__(self, clientset, recorder: EventRecorder):
        self._cs = clientset
        self._rec = recorder

    def create_services_with_controller_ref(self, namespace: str, service: Dict[str, Any], owner,
                                            controller_ref: Dict[str, Any]) -> Dict[str, Any]:
        svc = _with_owner(service, controller_ref)
        svc.setdefault("apiVersion", "v1")
        svc.setdefault("kind", "Service")
        svc["metadata"]["namespace"] = namespace
        try:
            created = self._cs.core_v1().services(namespace).create(svc)
        except APIError as e:
            self._rec.eventf(owner, EVENT_WARNING, "FailedCreateService", "Error creating: %s", e.message)
            raise
        self._rec.eventf(owner, EVENT_NORMAL, "SuccessfulCreateService", "Created service: %s", M.name_of(created))
        return created

    def delete_service(self, namespace: str, name: str, owner) -> None:
        try:
            self._cs.core_v1().services(namespace).delete(name)
        except APIError as e:
            if e.reason == "NotFound":
                return
            self._rec.eventf(owner, EVENT_WARNING, "FailedDeleteService", "Error deleting: %s", e.message)
            raise
        self._rec.eventf(owner, EVENT_NORMAL, "SuccessfulDeleteService", "Deleted service: %s", name)

    def patch_service(self, namespace: str, name: str, patch: Dict[str, Any]) -> Dict[str, Any]:
        return self._cs.core_v1().services(namespace).patch(name, patch)


class FakePodControl:
    """Records templates / deleted names instead of touching a store."""

    def __init__(self):
        self.lock = threading.Lock()
        self.templates: List[Dict[str, Any]] = []
        self.controller_refs: List[Dict[str, Any]] = []
        self.deleted: List[str] = []
        self.force_deleted: List[str] = []
        self.patches: List[Any] = []
        self.create_error: Optional[Exception] = None

    def create_pods_with_controller_ref(self, namespace, template, owner, controller_ref):
        with self.lock:
            if self.create_error is not None:
                raise self.create_error
            self.templates.append(M.deepcopy(template))
            self.controller_refs.append(controller_ref)
        return {"metadata": dict(template.get("metadata") or {}, namespace=namespace)}

    def delete_pod(self, namespace, name, owner, grace_period_seconds=None):
        with self.lock:
            self.deleted.append(name)
            if grace_period_seconds == 0:
                self.force_deleted.append(name)

    def patch_pod(self, namespace, name, patch):
        with self.lock:
            self.patches.append((name, patch))
        return {}

    def clear(self):
        with self.lock:
            self.templates.clear()
            self.controller_refs.clear()
            self.deleted.clear()
            self.force_deleted.clear()
            self.patches.clear()


class FakeServiceControl:
    def __init__(self):
        self.lock = threading.Lock()
        self.services: List[Dict[str, Any]] = []
        self.deleted: List[str] = []
        self.patches: List[Any] = []

    def create_services_with_controller_ref(self, namespace, service, owner, controller_ref):
        with self.lock:
            self.services.append(M.deepcopy(service))
        return service

    def delete_service(self, namespace, name, owner):
        with self.lock:
            self.deleted.append(name)

    def patch_service(self, namespace, name, patch):
        with self.lock:
            self.patches.append((name, patch))
        return {}


class ControllerRefManager:
    """Adopt / release logic shared by pods and services (SURVEY.md §2.2 ``ClaimPods`` row)."""

    def __init__(self, patch_fn: Callable[[str, str, Dict[str, Any]], Any], owner, selector: Dict[str, str],
                 can_adopt: Callable[[], None]):
        self._patch = patch_fn
        self._owner = owner.to_dict() if hasattr(owner, "to_dict") else owner
        self._selector = selector
        self._can_adopt = can_adopt
        self._can_adopt_err: Optional[Exception] = None
        self._checked = False

    def _check_adopt(self) -> bool:
        if not self._checked:
            self._checked = True
            try:
                self._can_adopt()
            except Exception as e:  # noqa: BLE001
                self._can_adopt_err = e
        return self._can_adopt_err is None

    def claim(self, objs: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
        claimed = []
        owner_uid = M.uid_of(self._owner)
        for o in objs:
            ref = M.get_controller_of(o)
            matches = M.selector_matches(self._selector, M.labels_of(o))
            if ref is not None:
                if ref.get("uid") != owner_uid:
                    continue  # owned by someone else
                if matches:
                    claimed.append(o)
                    continue
                if self._owner.get("metadata", {}).get("deletionTimestamp"):
                    continue
                self._release(o)
                continue
            # orphan
            if not matches or self._owner.get("metadata", {}).get("deletionTimestamp"):
                continue
            if o.get("metadata", {}).get("deletionTimestamp"):
                continue
            if not self._check_adopt():
                continue
            if self._adopt(o):
                claimed.append(o)
        return claimed

    def _adopt(self, o) -> bool:
        refs = list(o.get("metadata", {}).get("ownerReferences") or []) + [M.owner_reference(self._owner)]
        try:
            self._patch(M.namespace_of(o), M.name_of(o), {"metadata": {"ownerReferences": refs, "uid": M.uid_of(o)}})
            o.setdefault("metadata", {})["ownerReferences"] = refs
            klog.V(2).info("adopted %s %s/%s", o.get("kind"), M.namespace_of(o), M.name_of(o))
            return True
        except APIError as e:
            if e.reason not in ("NotFound", "Invalid"):
                raise
            return False

    def _release(self, o) -> None:
        owner_uid = M.uid_of(self._owner)
        refs = [r for r in o.get("metadata", {}).get("ownerReferences") or [] if r.get("uid") != owner_uid]
        try:
            self._patch(M.namespace_of(o), M.name_of(o), {"metadata": {"ownerReferences": refs or None}})
            klog.V(2).info("released %s %s/%s", o.get("kind"), M.namespace_of(o), M.name_of(o))
        except APIError as e:
            if e.reason not in ("NotFound", "Invalid"):
                raise


def recheck_deletion_timestamp(get_fresh: Callable[[], Any]) -> Callable[[], None]:
    """``controller.RecheckDeletionTimestamp`` (pod.go:138-147): a live GET guards adoption."""

    def check() -> None:
        fresh = get_fresh()
        ts = fresh.deletion_timestamp if hasattr(fresh, "deletion_timestamp") else \
            fresh.get("metadata", {}).get("deletionTimestamp")
        if ts:
            name = fresh.name if hasattr(fresh, "name") else M.name_of(fresh)
            raise RuntimeError(f"{name} has just been deleted at {ts}")

    return check
