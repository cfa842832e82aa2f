"""Shared informers (LIST + WATCH -> indexed local cache -> event handlers) and listers.

Parity: /root/reference/pkg/client/informers/externalversions/factory.go:91-175
(``NewSharedInformerFactoryWithOptions``, ``Start``, ``WaitForCacheSync``, ``InformerFor``: one
shared informer per type, namespace scoping, default resync), generic.go:1-61,
aitrainingjob/v1/aitrainingjob.go:56-88 (ListWatch funcs, namespace index) and the listers of
pkg/client/listers/aitrainingjob/v1/aitrainingjob.go:28-93 (``List(selector)``,
``AITrainingJobs(ns).Get(name)`` by key ``ns/name``).  The upstream reflector/DeltaFIFO the reference
gets from client-go is re-specified here: relist on ``Expired``, tombstones for deletes missed
during a disconnect, periodic resync re-delivering updates.

Listers hand out *copies*: the reference mutates cached objects in place (controller.go:297,
status.go:85) -- that cannot happen here.
"""
from __future__ import annotations

import threading
import time
from typing import Any, Callable, Dict, List, Optional

from ..api import meta as M
from ..api import register as R
from ..api.types import AITrainingJob
from ..store.apiserver import APIError
from ..utils import klog, lifecycle
from .clientset import Clientset, ResourceClient


class DeletedFinalStateUnknown:
    """Tombstone delivered when a delete was missed and only noticed on relist (client-go cache pkg)."""

    def __init__(self, key: str, obj: Dict[str, Any]):
        self.key = key
        self.obj = obj


def deletion_handling_key(obj) -> str:
    """``cache.DeletionHandlingMetaNamespaceKeyFunc`` (controller.go:407)."""
    if isinstance(obj, DeletedFinalStateUnknown):
        return obj.key
    if isinstance(obj, AITrainingJob):
        return obj.key()
    return M.key_of(obj)


class Indexer:
    """``ns/name`` -> object cache with a namespace index (aitrainingjob.go:79) and optional secondary indices
    (client-go ``cache.Indexers``: name -> function returning the index values of an object).

    Writers (the informer's reflector thread, tests) serialise on a lock.  Readers take none: every read is a single
    C-level dict operation (``get``, ``list(d.values())``) on a dict that writers only mutate with single operations or
    replace wholesale, and cached objects are never modified in place -- an update stores a new object.  Under load the
    read lock was the hottest lock of the process (each blocked acquisition also costs a GIL hand-over)."""

    def __init__(self, indexers: Optional[Dict[str, Callable[[Dict[str, Any]], List[str]]]] = None):
        self._lock = threading.RLock()
        self._items: Dict[str, Dict[str, Any]] = {}
        self._by_ns: Dict[str, Dict[str, Dict[str, Any]]] = {}
        self._indexers: Dict[str, Callable[[Dict[str, Any]], List[str]]] = dict(indexers or {})
        self._indices: Dict[str, Dict[str, Dict[str, Dict[str, Any]]]] = {n: {} for n in self._indexers}

    def add_indexers(self, indexers: Dict[str, Callable[[Dict[str, Any]], List[str]]]) -> None:
        with self._lock:
            for name, fn in indexers.items():
                if name in self._indexers:
                    continue
                self._indexers[name] = fn
                idx: Dict[str, Dict[str, Dict[str, Any]]] = {}
                for key, o in self._items.items():
                    for v in fn(o):
                        idx.setdefault(v, {})[key] = o
                self._indices[name] = idx

    def _unindex(self, key: str, old: Dict[str, Any]) -> None:
        ns = self._by_ns.get(M.namespace_of(old))
        if ns is not None:
            ns.pop(key, None)
        for name, fn in self._indexers.items():
            idx = self._indices[name]
            for v in fn(old):
                bucket = idx.get(v)
                if bucket is not None:
                    bucket.pop(key, None)
                    if not bucket:
                        idx.pop(v, None)

    def replace(self, items: List[Dict[str, Any]]) -> None:
        with self._lock:
            new_items = {M.key_of(o): o for o in items}
            by_ns: Dict[str, Dict[str, Dict[str, Any]]] = {}
            indices: Dict[str, Dict[str, Dict[str, Dict[str, Any]]]] = {n: {} for n in self._indexers}
            for key, o in new_items.items():
                by_ns.setdefault(M.namespace_of(o), {})[key] = o
                for name, fn in self._indexers.items():
                    for v in fn(o):
                        indices[name].setdefault(v, {})[key] = o
            self._items, self._by_ns, self._indices = new_items, by_ns, indices

    def add(self, obj) -> None:
        key = M.key_of(obj)
        with self._lock:
            old = self._items.get(key)
            self._items[key] = obj
            # an update must never hide the object from a concurrent (lock-free) reader, not even for a moment: entries
            # are overwritten in place and only the index values the new version no longer has are removed
            ns_new = M.namespace_of(obj)
            self._by_ns.setdefault(ns_new, {})[key] = obj
            if old is not None and M.namespace_of(old) != ns_new:
                self._by_ns.get(M.namespace_of(old), {}).pop(key, None)
            for name, fn in self._indexers.items():
                idx = self._indices[name]
                new_vals = fn(obj)
                for v in new_vals:
                    idx.setdefault(v, {})[key] = obj
                if old is not None:
                    for v in fn(old):
                        if v not in new_vals:
                            bucket = idx.get(v)
                            if bucket is not None:
                                bucket.pop(key, None)
                                if not bucket:
                                    idx.pop(v, None)

    def delete(self, obj) -> None:
        key = M.key_of(obj)
        with self._lock:
            old = self._items.pop(key, None)
            if old is not None:
                self._unindex(key, old)

    def get_by_key(self, key: str) -> Optional[Dict[str, Any]]:
        return self._items.get(key)

    def list(self, namespace: str = "") -> List[Dict[str, Any]]:
        if not namespace:
            return list(self._items.values())
        return list(self._by_ns.get(namespace, _EMPTY).values())

    def by_index(self, name: str, value: str) -> List[Dict[str, Any]]:
        """Objects whose indexer ``name`` yields ``value`` (uncopied)."""
        return list(self._indices[name].get(value, _EMPTY).values())

    def keys(self) -> List[str]:
        return list(self._items.keys())

    def __len__(self) -> int:
        return len(self._items)


_EMPTY: Dict[str, Any] = {}


class SharedIndexInformer:
    def __init__(self, client: ResourceClient, resync_period: float = 0.0, name: str = ""):
        self._client = client
        self._resync = resync_period
        self.name = name or client.info.plural
        self.indexer = Indexer()
        self._handlers: List[Dict[str, Optional[Callable]]] = []
        self._synced = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._resync_thread: Optional[threading.Thread] = None
        self._stop = threading.Event()
        self._watch = None
        self._dispatch_lock = threading.Lock()
        self.last_sync_resource_version = ""

    # -- registration ---------------------------------------------------------------------------
    def add_event_handler(self, add: Optional[Callable] = None, update: Optional[Callable] = None,
                          delete: Optional[Callable] = None, filter_func: Optional[Callable] = None) -> None:
        """``cache.ResourceEventHandlerFuncs`` (+ ``FilteringResourceEventHandler`` via ``filter_func``)."""
        self._handlers.append({"add": add, "update": update, "delete": delete, "filter": filter_func})
        if self._synced.is_set() and add is not None:
            for o in self.indexer.list():
                if filter_func is None or filter_func(o):
                    add(o)

    def has_synced(self) -> bool:
        return self._synced.is_set()

    # -- dispatch -------------------------------------------------------------------------------
    def _fire(self, kind: str, *args) -> None:
        with self._dispatch_lock:
            for h in self._handlers:
                fn = h.get(kind)
                if fn is None:
                    continue
                flt = h.get("filter")
                probe = args[-1]
                if flt is not None and not flt(probe.obj if isinstance(probe, DeletedFinalStateUnknown) else probe):
                    continue
                try:
                    fn(*args)
                except Exception as e:  # noqa: BLE001 - a handler must not kill the reflector
                    klog.error("informer %s: %s handler raised: %r", self.name, kind, e)

    # -- reflector ------------------------------------------------------------------------------
    def _list_and_replace(self) -> str:
        lst = self._client.list()
        items = lst.get("items", [])
        old = {k: self.indexer.get_by_key(k) for k in self.indexer.keys()}
        new_keys = set()
        for o in items:
            k = M.key_of(o)
            new_keys.add(k)
            prev = old.get(k)
            self.indexer.add(o)
            if prev is None:
                self._fire("add", o)
            elif M.resource_version(prev) != M.resource_version(o):
                self._fire("update", prev, o)
        for k, prev in old.items():
            if k not in new_keys and prev is not None:
                self.indexer.delete(prev)
                self._fire("delete", DeletedFinalStateUnknown(k, prev))
        rv = str(lst.get("metadata", {}).get("resourceVersion", ""))
        self.last_sync_resource_version = rv
        return rv

    def _run(self) -> None:
        backoff = 0.05
        rv = ""
        need_list = True
        while not self._stop.is_set():
            try:
                if need_list:
                    rv = self._list_and_replace()
                    self._synced.set()
                    need_list = False
                backoff = 0.05
                self._watch = self._client.watch(resource_version=rv, timeout=300.0)
                for ev in self._watch:
                    if self._stop.is_set():
                        break
                    self._handle(ev)
                    rv = self.last_sync_resource_version or rv
            except APIError as e:
                if self._stop.is_set():
                    break
                need_list = True
                if e.reason not in ("Expired", "Gone"):
                    klog.warning("informer %s: list/watch failed: %s; retrying in %.2fs", self.name, e.message, backoff)
                    self._stop.wait(backoff)
                    backoff = min(backoff * 2, 5.0)
            except Exception as e:  # noqa: BLE001
                if self._stop.is_set():
                    break
                need_list = True
                klog.warning("informer %s: reflector error %r; retrying in %.2fs", self.name, e, backoff)
                self._stop.wait(backoff)
                backoff = min(backoff * 2, 5.0)
            finally:
                w, self._watch = self._watch, None
                if w is not None:
                    try:
                        w.close()
                    except Exception:  # noqa: BLE001
                        pass

    def _handle(self, ev: Dict[str, Any]) -> None:
        obj = ev["object"]
        et = ev["type"]
        key = M.key_of(obj)
        self.last_sync_resource_version = M.resource_version(obj)
        if et == "ADDED":
            prev = self.indexer.get_by_key(key)
            self.indexer.add(obj)
            if prev is None:
                self._fire("add", obj)
            else:
                self._fire("update", prev, obj)
        elif et == "MODIFIED":
            prev = self.indexer.get_by_key(key)
            self.indexer.add(obj)
            if prev is None:
                self._fire("add", obj)
            else:
                self._fire("update", prev, obj)
        elif et == "DELETED":
            self.indexer.delete(obj)
            self._fire("delete", obj)

    def _resync_loop(self) -> None:
        while not self._stop.wait(self._resync):
            if not self._synced.is_set():
                continue
            for o in self.indexer.list():
                self._fire("update", o, o)

    def run(self, stop: Optional[threading.Event] = None) -> None:
        if self._thread is not None:
            return
        if stop is not None:
            self._stop = stop
        self._thread = lifecycle.spawn(self._run, f"informer-{self.name}")
        lifecycle.register_stop(self.stop)
        threading.Thread(target=lambda: (self._stop.wait(), self.stop()), name=f"informer-stop-{self.name}",
                         daemon=True).start()
        if self._resync and self._resync > 0:
            self._resync_thread = lifecycle.spawn(self._resync_loop, f"resync-{self.name}")

    def stop(self) -> None:
        self._stop.set()
        w = self._watch
        if w is not None:
            try:
                w.close()
            except Exception:  # noqa: BLE001
                pass


# ------------------------------------------------------------------------------------ listers
class NamespaceLister:
    def __init__(self, indexer: Indexer, namespace: str, conv: Callable[[Dict[str, Any]], Any], what: str):
        self._indexer = indexer
        self._ns = namespace
        self._conv = conv
        self._what = what

    def list(self, selector: Optional[Dict[str, str]] = None) -> List[Any]:
        sel = selector or {}
        return [self._conv(o) for o in self._indexer.list(self._ns) if M.selector_matches(sel, M.labels_of(o))]

    def get(self, name: str) -> Any:
        o = self._indexer.get_by_key(f"{self._ns}/{name}" if self._ns else name)
        if o is None:
            raise APIError(404, "NotFound", f"{self._what} \"{name}\" not found")
        return self._conv(o)


class GenericLister:
    """Read-only cache access; objects are deep-copied on the way out."""

    def __init__(self, indexer: Indexer, what: str, conv: Optional[Callable] = None):
        self._indexer = indexer
        self._what = what
        self._conv = conv or M.deepcopy

    def list(self, selector: Optional[Dict[str, str]] = None) -> List[Any]:
        sel = selector or {}
        return [self._conv(o) for o in self._indexer.list() if M.selector_matches(sel, M.labels_of(o))]

    def by_index(self, name: str, value: str) -> List[Dict[str, Any]]:
        """The cache's own objects (uncopied, read-only) filed under ``value`` by the informer's indexer ``name``."""
        return self._indexer.by_index(name, value)

    def copy_of(self, obj: Dict[str, Any]) -> Any:
        return self._conv(obj)

    def peek_key(self, key: str) -> Optional[Dict[str, Any]]:
        """The cached object under ``ns/name`` itself (uncopied, read-only), or None."""
        return self._indexer.get_by_key(key)

    def peek(self) -> List[Dict[str, Any]]:
        """The cache's own objects, uncopied, for read-only aggregation (free-slot accounting).  Never mutate them."""
        return list(self._indexer.list())

    def namespaced(self, namespace: str) -> NamespaceLister:
        return NamespaceLister(self._indexer, namespace, self._conv, self._what)

    def get(self, name: str) -> Any:
        return NamespaceLister(self._indexer, "", self._conv, self._what).get(name)


class AITrainingJobLister(GenericLister):
    """pkg/client/listers/aitrainingjob/v1/aitrainingjob.go:28-34."""

    def __init__(self, indexer: Indexer):
        super().__init__(indexer, f"{R.AITRAININGJOB.plural}.{R.AITRAININGJOB.group}", AITrainingJob.from_dict)

    def aitrainingjobs(self, namespace: str) -> NamespaceLister:
        return self.namespaced(namespace)


# ------------------------------------------------------------------------------------ factory
class _TypedInformer:
    def __init__(self, factory: "SharedInformerFactory", info: R.ResourceInfo):
        self._factory = factory
        self._info = info

    def informer(self) -> SharedIndexInformer:
        return self._factory.informer_for(self._info)

    def lister(self):
        idx = self.informer().indexer
        if self._info.kind == R.AITRAININGJOB.kind:
            return AITrainingJobLister(idx)
        return GenericLister(idx, self._info.plural)


class _V1Group:
    def __init__(self, factory, infos: Dict[str, R.ResourceInfo]):
        self._factory = factory
        self._infos = infos

    def __getattr__(self, item):
        if item in self._infos:
            return lambda: _TypedInformer(self._factory, self._infos[item])
        raise AttributeError(item)


class _Group:
    def __init__(self, factory, infos):
        self._v1 = _V1Group(factory, infos)

    def v1(self) -> _V1Group:
        return self._v1


class SharedInformerFactory:
    """factory.go:91-175: one shared informer per resource, optional namespace scope, default resync."""

    def __init__(self, clientset: Clientset, default_resync: float = 0.0, namespace: str = ""):
        self._cs = clientset
        self._resync = default_resync
        self._namespace = namespace
        self._informers: Dict[str, SharedIndexInformer] = {}
        self._started: set = set()
        self._lock = threading.Lock()

    def informer_for(self, info: R.ResourceInfo) -> SharedIndexInformer:
        with self._lock:
            inf = self._informers.get(info.kind)
            if inf is None:
                inf = SharedIndexInformer(self._cs.resource(info, self._namespace), self._resync, info.plural)
                self._informers[info.kind] = inf
            return inf

    # typed accessors mirroring factory.Elasticdeeplearning().V1().AITrainingJobs() / Core().V1().Pods()
    def elasticdeeplearning(self) -> _Group:
        return _Group(self, {"aitrainingjobs": R.AITRAININGJOB})

    def core(self) -> _Group:
        return _Group(self, {"pods": R.POD, "services": R.SERVICE, "nodes": R.NODE, "events": R.EVENT})

    def start(self, stop: threading.Event) -> None:
        with self._lock:
            for kind, inf in self._informers.items():
                if kind not in self._started:
                    self._started.add(kind)
                    inf.run(stop)

    def wait_for_cache_sync(self, stop: threading.Event, timeout: float = 30.0) -> Dict[str, bool]:
        deadline = time.monotonic() + timeout
        out: Dict[str, bool] = {}
        with self._lock:
            infs = dict(self._informers)
        for kind, inf in infs.items():
            while not inf.has_synced() and not stop.is_set() and time.monotonic() < deadline:
                time.sleep(0.005)
            out[kind] = inf.has_synced()
        return out


def wait_for_cache_sync(stop: threading.Event, *synced: Callable[[], bool], timeout: float = 30.0) -> bool:
    """``controller.WaitForCacheSync`` (controller.go:195)."""
    deadline = time.monotonic() + timeout
    while not stop.is_set() and time.monotonic() < deadline:
        if all(fn() for fn in synced):
            return True
        time.sleep(0.005)
    return all(fn() for fn in synced)
