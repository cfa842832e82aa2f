"""Fake clientset for unit tests: in-memory object tracker + reactors + action log.

Parity: /root/reference/pkg/client/clientset/versioned/fake/clientset_generated.go:35-81
(``NewSimpleClientset(objects...)`` with an ObjectTracker, a ``*`` reactor and a watch reactor) and
typed/aitrainingjob/v1/fake/fake_aitrainingjob.go:41-139 (each verb records an Action and defers
to the tracker).  The reference never uses its fake (it has no tests); here the
controller's unit tests run against it.

The tracker is a real ``APIServer`` (in-memory, admission off) so resourceVersion / watch /
cascade semantics match production; reactors can intercept any (verb, resource) first.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Tuple

from ..api import meta as M
from ..api import register as R
from ..store.apiserver import APIServer
from ..store.transport import LocalTransport, Transport
from .clientset import Clientset


@dataclass
class Action:
    verb: str
    resource: str
    namespace: str = ""
    name: str = ""
    subresource: str = ""
    obj: Any = None

    def matches(self, verb: str, resource: str) -> bool:
        return (verb == "*" or verb == self.verb) and (resource == "*" or resource == self.resource)


Reactor = Callable[[Action], Tuple[bool, Any]]


class FakeTransport(Transport):
    def __init__(self, tracker: APIServer):
        self.tracker = tracker
        self._local = LocalTransport(tracker)
        self.actions: List[Action] = []
        self._reactors: List[Tuple[str, str, Reactor]] = []
        self._lock = threading.Lock()

    # -- reactor chain ------------------------------------------------------------------------
    def prepend_reactor(self, verb: str, resource: str, fn: Reactor) -> None:
        self._reactors.insert(0, (verb, resource, fn))

    def add_reactor(self, verb: str, resource: str, fn: Reactor) -> None:
        self._reactors.append((verb, resource, fn))

    def _invoke(self, action: Action, default: Callable[[], Any]) -> Any:
        with self._lock:
            self.actions.append(action)
            reactors = list(self._reactors)
        for verb, resource, fn in reactors:
            if action.matches(verb, resource):
                handled, ret = fn(action)
                if handled:
                    if isinstance(ret, Exception):
                        raise ret
                    return ret
        return default()

    def clear_actions(self) -> None:
        with self._lock:
            self.actions.clear()

    def actions_for(self, verb: str, resource: str) -> List[Action]:
        with self._lock:
            return [a for a in self.actions if a.matches(verb, resource)]

    # -- Transport ------------------------------------------------------------------------------
    def create(self, info, namespace, obj):
        a = Action("create", info.plural, namespace, M.name_of(obj), obj=M.deepcopy(obj))
        return self._invoke(a, lambda: self._local.create(info, namespace, obj))

    def get(self, info, namespace, name):
        return self._invoke(Action("get", info.plural, namespace, name), lambda: self._local.get(info, namespace, name))

    def list(self, info, namespace="", label_selector="", field_selector=""):
        return self._invoke(Action("list", info.plural, namespace),
                            lambda: self._local.list(info, namespace, label_selector, field_selector))

    def update(self, info, namespace, name, obj, subresource=""):
        a = Action("update", info.plural, namespace, name, subresource, M.deepcopy(obj))
        return self._invoke(a, lambda: self._local.update(info, namespace, name, obj, subresource))

    def patch(self, info, namespace, name, patch, patch_type="application/merge-patch+json", subresource=""):
        a = Action("patch", info.plural, namespace, name, subresource, M.deepcopy(patch))
        return self._invoke(a, lambda: self._local.patch(info, namespace, name, patch, patch_type, subresource))

    def delete(self, info, namespace, name, grace_period_seconds=None, uid=""):
        a = Action("delete", info.plural, namespace, name, obj={"gracePeriodSeconds": grace_period_seconds})
        return self._invoke(a, lambda: self._local.delete(info, namespace, name, grace_period_seconds, uid))

    def delete_collection(self, info, namespace, label_selector="", grace_period_seconds=None):
        a = Action("delete-collection", info.plural, namespace)
        return self._invoke(a, lambda: self._local.delete_collection(info, namespace, label_selector,
                                                                     grace_period_seconds))

    def watch(self, info, namespace="", resource_version="", label_selector="", timeout=None):
        return self._invoke(Action("watch", info.plural, namespace),
                            lambda: self._local.watch(info, namespace, resource_version, label_selector, timeout))


class FakeClientset(Clientset):
    """``NewSimpleClientset``: seed objects are created in the tracker before use."""

    def __init__(self, *objects: Dict[str, Any]):
        tracker = APIServer(admission=False)
        super().__init__(FakeTransport(tracker))
        self.tracker = tracker
        for o in objects:
            info = R.by_kind(o["kind"])
            tracker.create(info, M.namespace_of(o), o)

    @property
    def fake(self) -> FakeTransport:
        return self.transport  # type: ignore[return-value]

    def actions(self) -> List[Action]:
        return list(self.fake.actions)

    def clear_actions(self) -> None:
        self.fake.clear_actions()

    def prepend_reactor(self, verb: str, resource: str, fn: Reactor) -> None:
        self.fake.prepend_reactor(verb, resource, fn)


def new_simple_clientset(*objects: Dict[str, Any]) -> FakeClientset:
    return FakeClientset(*objects)
