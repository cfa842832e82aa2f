"""Event recording: ``record.NewBroadcaster`` + recorder + sink, and the fake recorder for tests.

Parity: the reference builds a broadcaster that logs every event and writes it to the Events API
with component ``TrainingJobOperator`` (/root/reference/pkg/controller/controller.go:89-102); the
events themselves are emitted by the upstream pod/service controls (``SuccessfulCreatePod`` etc.,
SURVEY.md §2.2).  ``kubectl describe aitj`` lists them (Appendix B).  Events with the same
(object, type, reason, message) are aggregated by bumping ``count`` like client-go's correlator.
"""
from __future__ import annotations

import threading
from typing import Any, Dict, List, Optional, Tuple

from ..api import constants as C
from ..api import meta as M
from ..store.apiserver import APIError
from ..utils import klog

EVENT_NORMAL = "Normal"
EVENT_WARNING = "Warning"


def _object_reference(obj) -> Dict[str, Any]:
    d = obj.to_dict() if hasattr(obj, "to_dict") else obj
    md = d.get("metadata", {})
    return {"apiVersion": d.get("apiVersion", ""), "kind": d.get("kind", ""), "name": md.get("name", ""),
            "namespace": md.get("namespace", ""), "uid": md.get("uid", ""),
            "resourceVersion": str(md.get("resourceVersion", ""))}


class EventRecorder:
    def __init__(self, clientset=None, component: str = C.CONTROLLER_NAME, log: bool = True):
        self._cs = clientset
        self._component = component
        self._log = log
        self._lock = threading.Lock()
        self._seen: Dict[Tuple, str] = {}

    def event(self, obj, etype: str, reason: str, message: str) -> None:
        ref = _object_reference(obj)
        if self._log:
            klog.info("Event(%s/%s %s): type: '%s' reason: '%s' %s", ref["namespace"], ref["name"], ref["kind"], etype,
                      reason, message)
        if self._cs is None:
            return
        ns = ref["namespace"] or "default"
        key = (ref["uid"], etype, reason, message)
        now = M.format_time()
        try:
            with self._lock:
                existing = self._seen.get(key)
            if existing:
                try:
                    ev = self._cs.core_v1().events(ns).get(existing)
                    ev["count"] = int(ev.get("count", 1)) + 1
                    ev["lastTimestamp"] = now
                    self._cs.core_v1().events(ns).update(ev)
                    return
                except APIError:
                    pass
            ev = {
                "apiVersion": "v1", "kind": "Event",
                "metadata": {"generateName": f"{ref['name']}.", "namespace": ns},
                "involvedObject": ref, "reason": reason, "message": message, "type": etype,
                "source": {"component": self._component}, "firstTimestamp": now, "lastTimestamp": now, "count": 1,
            }
            created = self._cs.core_v1().events(ns).create(ev)
            with self._lock:
                self._seen[key] = created["metadata"]["name"]
                if len(self._seen) > 4096:
                    self._seen.clear()
        except Exception as e:  # noqa: BLE001 - events are best effort
            klog.warning("failed to record event %s/%s: %r", reason, message, e)

    def eventf(self, obj, etype: str, reason: str, fmt: str, *args) -> None:
        self.event(obj, etype, reason, fmt % args if args else fmt)


class FakeRecorder(EventRecorder):
    """Records ``"<type> <reason> <message>"`` strings (client-go ``record.FakeRecorder``)."""

    def __init__(self):
        super().__init__(None, log=False)
        self.events: List[str] = []

    def event(self, obj, etype: str, reason: str, message: str) -> None:
        self.events.append(f"{etype} {reason} {message}")


def events_for(clientset, obj) -> List[Dict[str, Any]]:
    """Events whose involvedObject is ``obj`` (what ``describe`` prints), oldest first."""
    ref = _object_reference(obj)
    lst = clientset.core_v1().events(ref["namespace"] or "default").list()
    out = [e for e in lst.get("items", []) if e.get("involvedObject", {}).get("uid") == ref["uid"]
           or (not ref["uid"] and e.get("involvedObject", {}).get("name") == ref["name"])]
    out.sort(key=lambda e: (e.get("firstTimestamp", ""), e.get("metadata", {}).get("resourceVersion", "0").zfill(12)))
    return out
