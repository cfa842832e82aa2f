"""The API-server analogue: Kubernetes REST semantics over the native versioned store.

The reference never ships an API server -- it is a client of one (cmd/app/server.go:111-151 builds
four clientsets against kube-apiserver).  A single-box deployment has no cluster, so this module
provides the same contract the reference's generated client expects
(pkg/client/clientset/versioned/typed/aitrainingjob/v1/aitrainingjob.go:66-190): create / get /
list / watch / update / updateStatus / patch / delete / deleteCollection with ``resourceVersion``
optimistic concurrency, label selectors, owner-reference cascade and watch streams, for the kinds
in ``api.register``.  Storage, versioning, watch fan-out and the WAL live in the C++ core
(``core/csrc/store.h``); this layer adds metadata management, admission (validation)
and graceful pod deletion (the kubelet analogue confirms the kill, then removes the object).
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Any, Dict, Iterator, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api import register as R
from ..api.validation import validate_dict
from ..core import _aitj_core as core

DEFAULT_NAMESPACE = "default"
DEFAULT_GRACE_SECONDS = 30


class APIError(Exception):
    """A Kubernetes ``Status`` failure."""

    def __init__(self, code: int, reason: str, message: str):
        super().__init__(message)
        self.code = code
        self.reason = reason
        self.message = message

    def status(self) -> Dict[str, Any]:
        return {"kind": "Status", "apiVersion": "v1", "metadata": {}, "status": "Failure", "message": self.message,
                "reason": self.reason, "code": self.code}

    @staticmethod
    def from_status(st: Dict[str, Any]) -> "APIError":
        return APIError(int(st.get("code", 500)), st.get("reason", "InternalError"), st.get("message", ""))


def is_not_found(e: Exception) -> bool:
    return isinstance(e, APIError) and e.reason == "NotFound"


def is_already_exists(e: Exception) -> bool:
    return isinstance(e, APIError) and e.reason == "AlreadyExists"


def is_conflict(e: Exception) -> bool:
    return isinstance(e, APIError) and e.reason == "Conflict"


def is_gone(e: Exception) -> bool:
    return isinstance(e, APIError) and e.reason in ("Gone", "Expired")


_CODE = {"NotFound": 404, "AlreadyExists": 409, "Conflict": 409, "Gone": 410}


def _wrap(e: Exception) -> APIError:
    reason = getattr(e, "reason", "InternalError")
    return APIError(_CODE.get(reason, 500), reason, str(e))


def merge_patch(target: Any, patch: Any) -> Any:
    """RFC 7386 JSON merge patch."""
    if not isinstance(patch, dict):
        return M.deepcopy(patch)
    if not isinstance(target, dict):
        target = {}
    out = dict(target)
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = merge_patch(out.get(k), v)
    return out


def json_patch(target: Any, ops: List[Dict[str, Any]]) -> Any:
    """RFC 6902 subset: add / replace / remove on object members and list indices."""
    doc = M.deepcopy(target)
    for op in ops:
        parts = [p.replace("~1", "/").replace("~0", "~") for p in op["path"].split("/")[1:]]
        parent = doc
        for p in parts[:-1]:
            parent = parent[int(p)] if isinstance(parent, list) else parent.setdefault(p, {})
        last = parts[-1]
        kind = op["op"]
        if isinstance(parent, list):
            idx = len(parent) if last == "-" else int(last)
            if kind == "add":
                parent.insert(idx, op["value"])
            elif kind == "replace":
                parent[idx] = op["value"]
            elif kind == "remove":
                parent.pop(idx)
        else:
            if kind in ("add", "replace"):
                parent[last] = op["value"]
            elif kind == "remove":
                parent.pop(last, None)
    return doc


def _field_matches(obj: Dict[str, Any], fields: Dict[str, str]) -> bool:
    for path, want in fields.items():
        cur: Any = obj
        for p in path.split("."):
            cur = cur.get(p) if isinstance(cur, dict) else None
        if (cur or "") != want:
            return False
    return True


class WatchStream:
    """Iterator over watch events ``{"type": ..., "object": ...}``; ``close()`` ends it."""

    def __init__(self, server: "APIServer", wid: int, info: R.ResourceInfo, selector: Dict[str, str],
                 timeout: Optional[float]):
        self._server = server
        self._wid = wid
        self._info = info
        self._selector = selector
        self._deadline = None if timeout is None else time.monotonic() + timeout
        self._closed = threading.Event()

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, Any]:
        while not self._closed.is_set():
            if self._deadline is not None and time.monotonic() > self._deadline:
                break
            ev = self._server._store.watch_next(self._wid, 0.2)
            if ev is None:
                continue
            etype, rec = ev
            obj = self._server._decode(rec)
            if self._selector and not M.selector_matches(self._selector, M.labels_of(obj)):
                continue
            return {"type": etype, "object": obj}
        self.close()
        raise StopIteration

    def poll(self, timeout: float = 0.2) -> Optional[Dict[str, Any]]:
        """One event or None after ``timeout`` (non-raising variant used by the HTTP façade)."""
        if self._closed.is_set():
            return None
        ev = self._server._store.watch_next(self._wid, timeout)
        if ev is None:
            return None
        etype, rec = ev
        obj = self._server._decode(rec)
        if self._selector and not M.selector_matches(self._selector, M.labels_of(obj)):
            return None
        return {"type": etype, "object": obj}

    def poll_raw(self, timeout: float = 0.2) -> Optional[bytes]:
        """One event already serialised as a JSON line, or None.  The store keeps objects as JSON bytes: the HTTP watch
        forwards them with the resourceVersion spliced in instead of decoding and re-encoding every event once per
        watcher (the label selector is evaluated on the labels the store keeps next to the bytes)."""
        if self._closed.is_set():
            return None
        ev = self._server._store.watch_next(self._wid, timeout)
        if ev is None:
            return None
        etype, rec = ev
        if self._selector and not M.selector_matches(self._selector, rec.get("labels") or {}):
            return None
        return b'{"type":"' + etype.encode() + b'","object":' + self._server._raw_with_rv(rec) + b'}\n'

    def poll_raw_many(self, timeout: float = 0.2, limit: int = 64) -> bytes:
        """Every event that is queued right now (at most ``limit``; waits up to ``timeout`` for the first one) as JSON
        lines in one buffer: a burst of writes costs a watcher one wake-up, one chunk and one flush, not one per event."""
        if self._closed.is_set():
            return b""
        out = []
        for etype, rec in self._server._store.watch_next_many(self._wid, timeout, limit):
            if self._selector and not M.selector_matches(self._selector, rec.get("labels") or {}):
                continue
            out.append(b'{"type":"' + etype.encode() + b'","object":' + self._server._raw_with_rv(rec) + b'}\n')
        return b"".join(out)

    @property
    def expired(self) -> bool:
        return self._closed.is_set() or (self._deadline is not None and time.monotonic() > self._deadline)

    def close(self) -> None:
        if not self._closed.is_set():
            self._closed.set()
            self._server._store.watch_close(self._wid)


class APIServer:
    def __init__(self, wal_path: str = "", admission: bool = True, history: int = 16384):
        self._wal_path = wal_path
        self._store = core.Store(wal_path, history)
        self._admission = admission
        self.started_at = time.time()
        self.request_count = 0

    # ------------------------------------------------------------------ encoding helpers
    _MD_PREFIX = b'{"metadata":{'

    @staticmethod
    def _encode(obj: Dict[str, Any]) -> bytes:
        md = dict(obj.get("metadata") or {})
        md.pop("resourceVersion", None)
        o = {"metadata": md}                 # metadata first: `_raw_with_rv` splices the resourceVersion in after it
        for k, v in obj.items():
            if k != "metadata":
                o[k] = v
        return json.dumps(o, separators=(",", ":")).encode()

    @classmethod
    def _raw_with_rv(cls, rec: Dict[str, Any]) -> bytes:
        """The stored JSON with ``metadata.resourceVersion`` set, without a decode / encode round trip."""
        data: bytes = rec["data"]
        n = len(cls._MD_PREFIX)
        if data.startswith(cls._MD_PREFIX) and data[n:n + 1] == b'"':
            return cls._MD_PREFIX + b'"resourceVersion":"' + str(rec["rv"]).encode() + b'",' + data[n:]
        return json.dumps(cls._decode(rec), separators=(",", ":")).encode()     # written by an older layout (WAL)

    @staticmethod
    def _decode(rec: Dict[str, Any]) -> Dict[str, Any]:
        obj = json.loads(rec["data"])
        obj.setdefault("metadata", {})["resourceVersion"] = str(rec["rv"])
        return obj

    def _ns(self, info: R.ResourceInfo, namespace: str, for_write: bool = False) -> str:
        if not info.namespaced:
            return ""
        return namespace or (DEFAULT_NAMESPACE if for_write else "")

    def _admit(self, info: R.ResourceInfo, obj: Dict[str, Any]) -> None:
        if self._admission and info.kind == C.KIND:
            errs = validate_dict(obj)
            if errs:
                raise APIError(422, "Invalid", f"{C.KIND}.{C.GROUP_NAME} \"{M.name_of(obj)}\" is invalid: "
                               + "; ".join(errs))

    def _check_owners(self, info: R.ResourceInfo, ns: str, obj: Dict[str, Any]) -> None:
        """Garbage-collector semantics applied eagerly: a dependent whose controller owner no longer exists (same
        kind/name/uid) would be collected by the Kubernetes GC right after creation; here it is refused, so a
        reconcile pass racing with the deletion of its job cannot leave orphan replicas behind."""
        for ref in obj.get("metadata", {}).get("ownerReferences") or []:
            if not ref.get("controller") or not ref.get("uid"):
                continue
            try:
                kind_info = R.by_kind(ref.get("kind", ""))
            except KeyError:
                continue
            try:
                owner = self._store.get(kind_info.kind, ns if kind_info.namespaced else "", ref.get("name", ""))
                if owner["uid"] == ref["uid"]:
                    continue
            except core.StoreError:
                pass
            raise APIError(404, "NotFound", f"owner {ref.get('kind')} \"{ref.get('name')}\" (uid {ref['uid']}) of "
                           f"{info.kind} \"{M.name_of(obj)}\" not found: refusing to create an orphan")

    # ------------------------------------------------------------------ verbs
    def create(self, info: R.ResourceInfo, namespace: str, obj: Dict[str, Any], as_bytes: bool = False):
        """``as_bytes`` (the HTTP façade): the answer is the stored JSON with the resourceVersion spliced in -- the
        bytes that were just produced for the store -- instead of a dict the caller would serialise a second time."""
        self.request_count += 1
        obj = M.deepcopy(obj)
        obj.setdefault("apiVersion", info.api_version)
        obj.setdefault("kind", info.kind)
        md = obj.setdefault("metadata", {})
        ns = self._ns(info, namespace or md.get("namespace", ""), for_write=True)
        if info.namespaced:
            if md.get("namespace") and namespace and md["namespace"] != namespace:
                raise APIError(400, "BadRequest", "the namespace of the provided object does not match the namespace "
                               "sent on the request")
            md["namespace"] = ns
        if not md.get("name"):
            if md.get("generateName"):
                md["name"] = md["generateName"] + M.new_uid()[:5]
            else:
                raise APIError(422, "Invalid", "metadata.name: Required value: name or generateName is required")
        self._admit(info, obj)
        self._check_owners(info, ns, obj)
        md["uid"] = M.new_uid()
        md["creationTimestamp"] = M.format_time()
        md.setdefault("generation", 1)
        md.pop("resourceVersion", None)
        md.pop("deletionTimestamp", None)
        try:
            rec = self._store.create(info.kind, ns, md["name"], md["uid"], self._encode(obj), M.labels_of(obj),
                                     M.owner_uids(obj))
        except core.StoreError as e:
            raise _wrap(e) from None
        if as_bytes:
            return self._raw_with_rv(rec)
        md["resourceVersion"] = str(rec["rv"])      # `obj` is our own copy and exactly what was stored: no re-parse
        return obj

    def get(self, info: R.ResourceInfo, namespace: str, name: str) -> Dict[str, Any]:
        return self._decode(self._get_record(info, namespace, name))

    def get_raw(self, info: R.ResourceInfo, namespace: str, name: str) -> bytes:
        """``get`` as the JSON bytes the HTTP façade sends (no decode / encode round trip)."""
        return self._raw_with_rv(self._get_record(info, namespace, name))

    def _get_record(self, info: R.ResourceInfo, namespace: str, name: str) -> Dict[str, Any]:
        self.request_count += 1
        try:
            return self._store.get(info.kind, self._ns(info, namespace, True), name)
        except core.StoreError as e:
            err = _wrap(e)
            if err.reason == "NotFound":
                err.message = f"{info.plural}.{info.group} \"{name}\" not found" if info.group else \
                    f"{info.plural} \"{name}\" not found"
            raise err from None

    def list(self, info: R.ResourceInfo, namespace: str = "", label_selector: str = "",
             field_selector: str = "") -> Dict[str, Any]:
        self.request_count += 1
        sel = M.parse_selector(label_selector)
        recs, rv = self._store.list(info.kind, self._ns(info, namespace), sel)
        items = [self._decode(r) for r in recs]
        if field_selector:
            fs = M.parse_selector(field_selector)
            items = [o for o in items if _field_matches(o, fs)]
        items.sort(key=lambda o: (M.namespace_of(o), M.name_of(o)))
        return {"apiVersion": info.api_version, "kind": info.kind + "List",
                "metadata": {"resourceVersion": str(rv)}, "items": items}

    def update(self, info: R.ResourceInfo, namespace: str, name: str, obj: Dict[str, Any],
               subresource: str = "", _cur: Optional[Dict[str, Any]] = None, as_bytes: bool = False):
        """``_cur``: the live object the caller has just read and built ``obj`` from (``patch``); saves reading and
        parsing it a second time -- if it is stale the store's optimistic-concurrency check says so (Conflict)."""
        self.request_count += 1
        ns = self._ns(info, namespace or M.namespace_of(obj), True)
        if _cur is not None and obj.get("metadata", {}).get("resourceVersion"):
            cur = _cur
        else:
            try:
                cur = self._decode(self._store.get(info.kind, ns, name))
            except core.StoreError as e:
                raise _wrap(e) from None
        new = M.deepcopy(obj)
        md = new.setdefault("metadata", {})
        if md.get("name") and md["name"] != name:
            raise APIError(400, "BadRequest", "the name of the object does not match the name on the URL")
        if md.get("uid") and md["uid"] != M.uid_of(cur):
            raise APIError(409, "Conflict", f"Precondition failed: UID in precondition: {md['uid']}, UID in object "
                           f"meta: {M.uid_of(cur)}")
        expected = int(md["resourceVersion"]) if md.get("resourceVersion") else 0
        if subresource == "status":
            merged = M.deepcopy(cur)
            merged["status"] = new.get("status", {})
            new = merged
            md = new["metadata"]
        else:
            self._admit(info, new)
            if "status" not in new and "status" in cur:
                new["status"] = cur["status"]
        # immutable / server-owned metadata
        cmd = cur["metadata"]
        md["name"] = name
        if info.namespaced:
            md["namespace"] = ns
        md["uid"] = cmd.get("uid")
        md["creationTimestamp"] = cmd.get("creationTimestamp")
        if cmd.get("deletionTimestamp"):
            md["deletionTimestamp"] = cmd["deletionTimestamp"]
            md["deletionGracePeriodSeconds"] = cmd.get("deletionGracePeriodSeconds")
        gen = int(cmd.get("generation", 1))
        if subresource != "status" and new.get("spec") != cur.get("spec"):
            gen += 1
        md["generation"] = gen
        new.setdefault("apiVersion", info.api_version)
        new.setdefault("kind", info.kind)
        try:
            rec = self._store.update(info.kind, ns, name, md["uid"], self._encode(new), M.labels_of(new),
                                     M.owner_uids(new), expected)
        except core.StoreError as e:
            raise _wrap(e) from None
        if as_bytes:
            return self._raw_with_rv(rec)
        md["resourceVersion"] = str(rec["rv"])
        return new

    def patch(self, info: R.ResourceInfo, namespace: str, name: str, patch: Any,
              patch_type: str = "application/merge-patch+json", subresource: str = "", as_bytes: bool = False):
        """Read-modify-write with retry on conflict (merge, strategic-as-merge, or JSON patch)."""
        for _ in range(16):
            cur = self.get(info, namespace, name)
            if patch_type == "application/json-patch+json":
                new = json_patch(cur, patch)
            else:
                new = merge_patch(cur, patch)
            new["metadata"]["resourceVersion"] = cur["metadata"]["resourceVersion"]
            try:
                return self.update(info, namespace, name, new, subresource=subresource, _cur=cur, as_bytes=as_bytes)
            except APIError as e:
                if e.reason != "Conflict":
                    raise
        raise APIError(409, "Conflict", f"patch of {info.kind} {name} kept conflicting")

    def delete(self, info: R.ResourceInfo, namespace: str, name: str, grace_period_seconds: Optional[int] = None,
               uid: str = "") -> Dict[str, Any]:
        """Delete ``name``.  Running pods are deleted gracefully (deletionTimestamp now, removal when the
        agent confirms with grace 0); everything else is removed at once, with owner-reference cascade."""
        self.request_count += 1
        ns = self._ns(info, namespace, True)
        try:
            cur = self._decode(self._store.get(info.kind, ns, name))
        except core.StoreError as e:
            raise _wrap(e) from None
        if uid and M.uid_of(cur) != uid:
            raise APIError(409, "Conflict", "uid precondition failed")
        if info.kind == "Pod":
            phase = cur.get("status", {}).get("phase", C.POD_PENDING)
            bound = bool(cur.get("spec", {}).get("nodeName"))
            live = bound and phase in (C.POD_PENDING, C.POD_RUNNING, C.POD_UNKNOWN)
            if live and grace_period_seconds != 0:
                if cur["metadata"].get("deletionTimestamp"):
                    return cur
                grace = DEFAULT_GRACE_SECONDS if grace_period_seconds is None else int(grace_period_seconds)
                tgrace = cur.get("spec", {}).get("terminationGracePeriodSeconds")
                if grace_period_seconds is None and isinstance(tgrace, int):
                    grace = tgrace
                cur["metadata"]["deletionTimestamp"] = M.format_time()
                cur["metadata"]["deletionGracePeriodSeconds"] = grace
                try:
                    rec = self._store.update(info.kind, ns, name, M.uid_of(cur), self._encode(cur), M.labels_of(cur),
                                             M.owner_uids(cur), 0)
                except core.StoreError as e:
                    raise _wrap(e) from None
                return self._decode(rec)
        try:
            removed = self._store.remove(info.kind, ns, name, uid)
        except core.StoreError as e:
            raise _wrap(e) from None
        return self._decode(removed[0])

    def delete_collection(self, info: R.ResourceInfo, namespace: str, label_selector: str = "",
                          grace_period_seconds: Optional[int] = None) -> Dict[str, Any]:
        lst = self.list(info, namespace, label_selector)
        for o in lst["items"]:
            try:
                self.delete(info, M.namespace_of(o), M.name_of(o), grace_period_seconds)
            except APIError as e:
                if e.reason != "NotFound":
                    raise
        return lst

    def watch(self, info: R.ResourceInfo, namespace: str = "", resource_version: str = "",
              label_selector: str = "", timeout: Optional[float] = None) -> WatchStream:
        self.request_count += 1
        since = int(resource_version) if resource_version not in ("", None) else 0
        try:
            wid = self._store.watch_open(info.kind, self._ns(info, namespace), since)
        except core.StoreError as e:
            err = _wrap(e)
            if err.reason == "Gone":
                err.reason = "Expired"
            raise err from None
        return WatchStream(self, wid, info, M.parse_selector(label_selector), timeout)

    # ------------------------------------------------------------------ misc
    def current_resource_version(self) -> int:
        return self._store.current_rv()

    def compact(self) -> None:
        self._store.compact()

    # ------------------------------------------------------------------ housekeeping
    def housekeeping_once(self, event_ttl_s: float = 3600.0, wal_max_bytes: int = 64 << 20) -> Dict[str, int]:
        """What etcd / kube-apiserver do in the background: Events expire (``--event-ttl``, 1 h upstream) and the
        write-ahead log -- which records every write, ~2 KB each -- is rewritten as a snapshot of the live objects once it
        has outgrown ``wal_max_bytes``.  Returns what was done."""
        done = {"events_expired": 0, "wal_compacted": 0}
        if event_ttl_s > 0:
            ev = R.by_kind("Event")
            for rec in self.list(ev).get("items", []):
                stamp = rec.get("lastTimestamp") or rec.get("metadata", {}).get("creationTimestamp")
                if stamp and M.seconds_since(stamp) > event_ttl_s:
                    try:
                        self.delete(ev, M.namespace_of(rec), M.name_of(rec), grace_period_seconds=0)
                        done["events_expired"] += 1
                    except APIError:
                        pass
        if self._wal_path and wal_max_bytes > 0:
            try:
                if os.path.getsize(self._wal_path) > wal_max_bytes:
                    self.compact()
                    done["wal_compacted"] = 1
            except OSError:
                pass
        return done

    def start_housekeeping(self, stop: threading.Event, period_s: float = 60.0, event_ttl_s: float = 3600.0,
                           wal_max_bytes: int = 64 << 20) -> threading.Thread:
        def loop():
            while not stop.wait(period_s):
                try:
                    self.housekeeping_once(event_ttl_s, wal_max_bytes)
                except Exception:  # noqa: BLE001 - never take the API server down for housekeeping
                    pass

        t = threading.Thread(target=loop, name="apiserver-housekeeping", daemon=True)
        t.start()
        return t

    def stats(self) -> Dict[str, Any]:
        return {"resourceVersion": self._store.current_rv(), "watchers": self._store.num_watchers(),
                "requests": self.request_count,
                "objects": {r.kind: self._store.count(r.kind) for r in R.all_resources()}}
