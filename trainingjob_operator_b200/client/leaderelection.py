"""Leader election over a lock object in the store (``leaderelection.RunOrDie``).

Parity: /root/reference/cmd/app/server.go:74-106 -- an ``endpoints`` resource lock
``kube-system/trainingjob-operator`` with identity ``<hostname>_<uuid>``, lease 15 s / renew
deadline 5 s / retry 3 s (cmd/app/options/options.go:39-52), ``OnStartedLeading`` runs the
controller, ``OnStoppedLeading`` is fatal (process exits; a supervisor restarts it).  The lock
record (holderIdentity, leaseDurationSeconds, acquireTime, renewTime, leaderTransitions) lives in
the ``control-plane.alpha.kubernetes.io/leader`` annotation of an Endpoints object, exactly like
client-go's EndpointsLock; ``leases`` is accepted as lock type as well.  Optimistic concurrency
(resourceVersion) on the lock object makes acquisition race-free.
"""
from __future__ import annotations

import json
import os
import socket
import threading
import time
from dataclasses import dataclass
from typing import Callable, Optional

from ..api import meta as M
from ..store.apiserver import APIError
from ..utils import klog
from .record import EVENT_NORMAL, EventRecorder

LEADER_ANNOTATION = "control-plane.alpha.kubernetes.io/leader"


def default_identity() -> str:
    return f"{socket.gethostname()}_{M.new_uid()}"


@dataclass
class LeaderElectionConfig:
    lock_type: str = "endpoints"
    lock_namespace: str = "kube-system"
    lock_name: str = "trainingjob-operator"
    identity: str = ""
    lease_duration: float = 15.0
    renew_deadline: float = 5.0
    retry_period: float = 3.0
    # hand the lease over on a clean stop (client-go ReleaseOnCancel); a callable lets a harness simulate a crash,
    # after which the standby has to wait for the lease to expire -- the reference's behaviour on a killed leader
    release_on_cancel: object = True


class LeaderElector:
    def __init__(self, clientset, config: LeaderElectionConfig, on_started_leading: Callable[[threading.Event], None],
                 on_stopped_leading: Callable[[], None], on_new_leader: Optional[Callable[[str], None]] = None,
                 recorder: Optional[EventRecorder] = None, clock: Callable[[], float] = time.time):
        if config.lease_duration <= config.renew_deadline:
            raise ValueError("leaseDuration must be greater than renewDeadline")
        if config.renew_deadline <= config.retry_period:
            raise ValueError("renewDeadline must be greater than retryPeriod")
        self.cfg = config
        if not self.cfg.identity:
            self.cfg.identity = default_identity()
        self._cs = clientset
        self._on_start = on_started_leading
        self._on_stop = on_stopped_leading
        self._on_new = on_new_leader
        self._recorder = recorder
        self._clock = clock
        self._observed_holder = ""
        self._observed_record: Optional[dict] = None
        self._observed_at = 0.0
        self._leading = threading.Event()

    # -- lock object ----------------------------------------------------------------------------
    def _client(self):
        if self.cfg.lock_type in ("leases", "lease"):
            return self._cs.coordination_v1().leases(self.cfg.lock_namespace)
        return self._cs.core_v1().endpoints(self.cfg.lock_namespace)

    def _read(self):
        obj = self._client().get(self.cfg.lock_name)
        raw = M.annotations_of(obj).get(LEADER_ANNOTATION)
        rec = json.loads(raw) if raw else None
        return obj, rec

    def _record(self, acquire_time: str, transitions: int) -> dict:
        return {"holderIdentity": self.cfg.identity, "leaseDurationSeconds": int(round(self.cfg.lease_duration)),
                "leaseDurationMillis": int(self.cfg.lease_duration * 1000),
                "acquireTime": acquire_time, "renewTime": M.format_time(), "renewTimeUnix": self._clock(),
                "leaderTransitions": transitions}

    def try_acquire_or_renew(self) -> bool:
        now = self._clock()
        try:
            obj, rec = self._read()
        except APIError as e:
            if e.reason != "NotFound":
                klog.warning("error retrieving resource lock %s/%s: %s", self.cfg.lock_namespace, self.cfg.lock_name,
                             e.message)
                return False
            new = {"metadata": {"name": self.cfg.lock_name, "namespace": self.cfg.lock_namespace,
                                "annotations": {LEADER_ANNOTATION: json.dumps(self._record(M.format_time(), 0))}}}
            try:
                self._client().create(new)
            except APIError as e2:
                klog.V(4).info("error initially creating leader election record: %s", e2.message)
                return False
            self._set_observed(self._record(M.format_time(), 0), now)
            return True
        if rec is None:
            rec = {"holderIdentity": "", "leaderTransitions": 0}
        if self._observed_record is None or rec.get("renewTimeUnix") != self._observed_record.get("renewTimeUnix") \
                or rec.get("holderIdentity") != self._observed_record.get("holderIdentity"):
            self._set_observed(rec, now)
        holder = rec.get("holderIdentity", "")
        lease = float(rec.get("leaseDurationMillis", rec.get("leaseDurationSeconds", 0) * 1000)) / 1000.0
        if holder and holder != self.cfg.identity and self._observed_at + lease > now:
            klog.V(4).info("lock is held by %s and has not yet expired", holder)
            return False
        if holder == self.cfg.identity:
            new_rec = self._record(rec.get("acquireTime", M.format_time()), int(rec.get("leaderTransitions", 0)))
        else:
            new_rec = self._record(M.format_time(), int(rec.get("leaderTransitions", 0)) + (1 if holder else 0))
        obj.setdefault("metadata", {}).setdefault("annotations", {})[LEADER_ANNOTATION] = json.dumps(new_rec)
        try:
            self._client().update(obj)  # carries resourceVersion -> optimistic concurrency
        except APIError as e:
            klog.V(4).info("failed to update lock: %s", e.message)
            return False
        self._set_observed(new_rec, now)
        return True

    def _set_observed(self, rec: dict, now: float) -> None:
        self._observed_record = rec
        self._observed_at = now
        holder = rec.get("holderIdentity", "")
        if holder != self._observed_holder:
            self._observed_holder = holder
            if self._on_new and holder:
                try:
                    self._on_new(holder)
                except Exception:  # noqa: BLE001
                    pass

    def is_leader(self) -> bool:
        return self._observed_holder == self.cfg.identity and self._leading.is_set()

    def get_leader(self) -> str:
        return self._observed_holder

    # -- loops ----------------------------------------------------------------------------------
    def _acquire(self, stop: threading.Event) -> bool:
        klog.info("attempting to acquire leader lease %s/%s...", self.cfg.lock_namespace, self.cfg.lock_name)
        while not stop.is_set():
            if self.try_acquire_or_renew():
                klog.info("successfully acquired lease %s/%s", self.cfg.lock_namespace, self.cfg.lock_name)
                if self._recorder is not None:
                    try:
                        lock_obj = self._client().get(self.cfg.lock_name)
                        self._recorder.event(lock_obj, EVENT_NORMAL, "LeaderElection",
                                             f"{self.cfg.identity} became leader")
                    except APIError:
                        pass
                return True
            stop.wait(self.cfg.retry_period * (1.0 + 0.2 * (os.getpid() % 5) / 5.0))
        return False

    def _renew(self, stop: threading.Event) -> None:
        while not stop.is_set():
            deadline = time.monotonic() + self.cfg.renew_deadline
            ok = False
            while time.monotonic() < deadline and not stop.is_set():
                if self.try_acquire_or_renew():
                    ok = True
                    break
                stop.wait(min(self.cfg.retry_period, max(0.0, deadline - time.monotonic())))
            if not ok:
                klog.info("failed to renew lease %s/%s: timed out waiting for the condition", self.cfg.lock_namespace,
                          self.cfg.lock_name)
                return
            stop.wait(self.cfg.retry_period)

    def release(self) -> None:
        """Best-effort hand-over on clean shutdown (newer client-go ``ReleaseOnCancel``)."""
        try:
            obj, rec = self._read()
            if rec and rec.get("holderIdentity") == self.cfg.identity:
                rec["holderIdentity"] = ""
                rec["leaseDurationMillis"] = 1
                rec["leaseDurationSeconds"] = 0
                obj["metadata"]["annotations"][LEADER_ANNOTATION] = json.dumps(rec)
                self._client().update(obj)
        except Exception:  # noqa: BLE001
            pass

    def run(self, stop: threading.Event) -> None:
        """Blocks: acquire, run ``on_started_leading`` in a thread, renew until lost, then ``on_stopped_leading``."""
        if not self._acquire(stop):
            return
        self._leading.set()
        lead_stop = threading.Event()
        t = threading.Thread(target=self._on_start, args=(lead_stop,), name="leader-run", daemon=True)
        t.start()
        try:
            self._renew(stop)
        finally:
            self._leading.clear()
            lead_stop.set()
            roc = self.cfg.release_on_cancel
            if stop.is_set() and (roc() if callable(roc) else roc):
                self.release()
            self._on_stop()


def run_or_die(clientset, config: LeaderElectionConfig, on_started_leading, on_stopped_leading, stop: threading.Event,
               recorder: Optional[EventRecorder] = None, on_new_leader=None) -> None:
    LeaderElector(clientset, config, on_started_leading, on_stopped_leading, on_new_leader, recorder).run(stop)
