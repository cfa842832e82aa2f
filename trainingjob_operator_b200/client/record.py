"""Event recording: ``record.NewBroadcaster`` + recorder + sink, and the fake recorder for tests.

Parity: the reference builds a broadcaster that logs every event and writes it to the Events API
with component ``TrainingJobOperator`` (/root/reference/pkg/controller/controller.go:89-102); the
events themselves are emitted by the upstream pod/service controls (``SuccessfulCreatePod`` etc.,
SURVEY.md §2.2).  ``kubectl describe aitj`` lists them (Appendix B).  Events with the same
(object, type, reason, message) are aggregated by bumping ``count`` like client-go's correlator.
"""
from __future__ import annotations

import itertools
import queue
import threading
import time
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
    """``event()`` never touches the API: like client-go's broadcaster (a buffered channel drained by a sink goroutine,
    controller.go:89-92) it logs, queues the event and returns; one daemon thread per recorder writes the queue to the
    Events API and does the (object, type, reason, message) aggregation.  A reconcile pass that creates or deletes N
    replicas used to pay 2N synchronous event writes inside the pass."""

    def __init__(self, clientset=None, component: str = C.CONTROLLER_NAME, log: bool = True):
        self._cs = clientset
        self._component = component
        self._log = log
        self._seen: Dict[Tuple, str] = {}            # only touched by the sink thread
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        # no lock on the recording path (a contended lock costs a GIL hand-over per event under load): the number of
        # recorded events comes from an itertools counter (atomic in CPython), the number written is owned by the sink
        self._ticket = itertools.count(1)
        self._recorded = 0
        self._written = 0
        self._start_lock = threading.Lock()
        self._sink: Optional[threading.Thread] = None

    def event(self, obj, etype: str, reason: str, message: str) -> None:
        ref = _object_reference(obj)
        if self._log:
            klog.info("Event(%s/%s %s): type: '%s' reason: '%s' %s", ref["namespace"], ref["name"], ref["kind"], etype,
                      reason, message)
        if self._cs is None:
            return
        n = next(self._ticket)
        if n > self._recorded:
            self._recorded = n                   # monotonic high-water mark; a lost race only under-reports briefly
        self._q.put((n, ref, etype, reason, message, M.format_time()))
        if self._sink is None:
            with self._start_lock:
                if self._sink is None:
                    self._sink = threading.Thread(target=self._drain, name="event-sink", daemon=True)
                    self._sink.start()

    def flush(self, timeout: float = 5.0) -> bool:
        """Wait until everything recorded so far has been written (tests, orderly shutdown)."""
        deadline = time.monotonic() + timeout
        target = self._recorded
        while self._written < target:
            if time.monotonic() > deadline:
                return False
            time.sleep(0.002)
        return True

    def _drain(self) -> None:
        while True:
            _n, *item = self._q.get()
            try:
                self._write(*item)
            finally:
                self._written += 1               # tickets are dense: written == recorded means the queue is drained

    def _write(self, ref: Dict[str, Any], etype: str, reason: str, message: str, now: str) -> None:
        ns = ref["namespace"] or "default"
        key = (ref["uid"], etype, reason, message)
        try:
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
