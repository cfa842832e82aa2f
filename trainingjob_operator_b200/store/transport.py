"""Transports between the typed clients and the API server.

``LocalTransport`` calls an in-process ``APIServer`` directly (all-in-one daemon, tests);
``HTTPTransport`` speaks the REST/watch wire protocol to ``store.http.APIHTTPServer`` -- the
process boundary the reference crosses with client-go's ``rest.Interface``
(pkg/client/clientset/versioned/typed/aitrainingjob/v1/aitrainingjob_client.go:69-89, APIPath
``/apis``).  Both expose the same five calls so clientsets, informers and the CLI are agnostic.
"""
from __future__ import annotations

import http.client
import json
import socket
import threading
import urllib.parse
from typing import Any, Dict, Optional

from ..api import register as R
from .apiserver import APIError, APIServer
from .fasthttp import FastResponse


class Transport:
    def create(self, info, namespace, obj): raise NotImplementedError
    def get(self, info, namespace, name): raise NotImplementedError
    def list(self, info, namespace="", label_selector="", field_selector=""): raise NotImplementedError
    def update(self, info, namespace, name, obj, subresource=""): raise NotImplementedError
    def patch(self, info, namespace, name, patch, patch_type="application/merge-patch+json", subresource=""):
        raise NotImplementedError
    def delete(self, info, namespace, name, grace_period_seconds=None, uid=""): raise NotImplementedError
    def delete_collection(self, info, namespace, label_selector="", grace_period_seconds=None):
        raise NotImplementedError
    def watch(self, info, namespace="", resource_version="", label_selector="", timeout=None):
        raise NotImplementedError


class LocalTransport(Transport):
    def __init__(self, server: APIServer):
        self.server = server

    def create(self, info, namespace, obj):
        return self.server.create(info, namespace, obj)

    def get(self, info, namespace, name):
        return self.server.get(info, namespace, name)

    def list(self, info, namespace="", label_selector="", field_selector=""):
        return self.server.list(info, namespace, label_selector, field_selector)

    def update(self, info, namespace, name, obj, subresource=""):
        return self.server.update(info, namespace, name, obj, subresource)

    def patch(self, info, namespace, name, patch, patch_type="application/merge-patch+json", subresource=""):
        return self.server.patch(info, namespace, name, patch, patch_type, subresource)

    def delete(self, info, namespace, name, grace_period_seconds=None, uid=""):
        return self.server.delete(info, namespace, name, grace_period_seconds, uid)

    def delete_collection(self, info, namespace, label_selector="", grace_period_seconds=None):
        return self.server.delete_collection(info, namespace, label_selector, grace_period_seconds)

    def watch(self, info, namespace="", resource_version="", label_selector="", timeout=None):
        return self.server.watch(info, namespace, resource_version, label_selector, timeout)


class ThrottledTransport(Transport):
    """Client-side request throttle, client-go's ``flowcontrol.NewTokenBucketRateLimiter(qps, burst)``: every request
    except a watch takes a token and waits when the bucket is empty.  The reference passes its ``rest.Config`` through
    unmodified (/root/reference/cmd/app/server.go:126-144), so each of its clientsets runs at client-go's defaults --
    5 qps, burst 10 -- which is what bounds how fast it can POST N pods + N services (``--kube-api-qps`` /
    ``--kube-api-burst`` here; 0 = unthrottled, and ``tools/latency_bench.py --reference-throttle`` measures the effect)."""

    def __init__(self, inner: Transport, qps: float, burst: int):
        import time as _time

        self.inner = inner
        self.qps = float(qps)
        self.burst = max(1, int(burst))
        self._tokens = float(self.burst)
        self._last = _time.monotonic()
        self._lock = threading.Lock()
        self.waited_s = 0.0
        self.requests = 0
        for name in ("master", "server"):
            if hasattr(inner, name):
                setattr(self, name, getattr(inner, name))

    def _take(self) -> None:
        import time as _time

        with self._lock:
            now = _time.monotonic()
            self._tokens = min(self.burst, self._tokens + (now - self._last) * self.qps)
            self._last = now
            self._tokens -= 1.0
            wait = -self._tokens / self.qps if self._tokens < 0 else 0.0
            self.requests += 1
            self.waited_s += wait
        if wait > 0:
            _time.sleep(wait)

    def create(self, *a, **kw):
        self._take(); return self.inner.create(*a, **kw)

    def get(self, *a, **kw):
        self._take(); return self.inner.get(*a, **kw)

    def list(self, *a, **kw):
        self._take(); return self.inner.list(*a, **kw)

    def update(self, *a, **kw):
        self._take(); return self.inner.update(*a, **kw)

    def patch(self, *a, **kw):
        self._take(); return self.inner.patch(*a, **kw)

    def delete(self, *a, **kw):
        self._take(); return self.inner.delete(*a, **kw)

    def delete_collection(self, *a, **kw):
        self._take(); return self.inner.delete_collection(*a, **kw)

    def watch(self, *a, **kw):
        return self.inner.watch(*a, **kw)


class _HTTPWatch:
    """Streaming watch: newline-delimited JSON events over a chunked response."""

    def __init__(self, conn: http.client.HTTPConnection, resp: http.client.HTTPResponse):
        self._conn = conn
        self._resp = resp
        self._closed = False

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, Any]:
        while not self._closed:
            try:
                line = self._resp.readline()
            except (socket.timeout, TimeoutError):
                continue
            except (OSError, http.client.HTTPException, ValueError, AttributeError):
                break
            if not line:
                break
            line = line.strip()
            if not line:
                continue
            ev = json.loads(line)
            if ev.get("type") == "ERROR":
                self.close()
                raise APIError.from_status(ev.get("object", {}))
            return ev
        self.close()
        raise StopIteration

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            try:
                self._conn.close()
            except Exception:  # noqa: BLE001
                pass


class _NoDelayConnection(http.client.HTTPConnection):
    """http.client sends the request head and the body as two writes; without TCP_NODELAY the second one waits for the
    server's delayed ACK (40 ms per request on loopback).  Responses are parsed by ``fasthttp.FastResponse`` (no e-mail
    package on the header block)."""

    response_class = FastResponse

    def connect(self):
        super().connect()
        try:
            self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        except OSError:
            pass


class HTTPTransport(Transport):
    """``--master http://127.0.0.1:8001`` style endpoint (or ``unix:///path.sock``)."""

    def __init__(self, master: str, timeout: float = 30.0, user_agent: str = "trainingjob-operator"):
        if "://" not in master:
            master = "http://" + master
        self.master = master.rstrip("/")
        u = urllib.parse.urlparse(self.master)
        self._host = u.hostname or "127.0.0.1"
        self._port = u.port or 8001
        self._timeout = timeout
        self._ua = user_agent
        self._local = threading.local()

    def _conn(self) -> http.client.HTTPConnection:
        c = getattr(self._local, "conn", None)
        if c is None:
            c = _NoDelayConnection(self._host, self._port, timeout=self._timeout)
            self._local.conn = c
        return c

    def _request(self, method: str, path: str, params: Optional[Dict[str, Any]] = None, body: Any = None,
                 content_type: str = "application/json") -> Dict[str, Any]:
        q = {k: v for k, v in (params or {}).items() if v not in (None, "")}
        url = path + ("?" + urllib.parse.urlencode(q) if q else "")
        data = None if body is None else json.dumps(body).encode()
        headers = {"Accept": "application/json", "User-Agent": self._ua}
        if data is not None:
            headers["Content-Type"] = content_type
        # One retry -- but never of a write the server may already have applied: a create reported as AlreadyExists or a
        # guarded PUT reported as Conflict would make callers (leader election, expectation bookkeeping) see a failure
        # for a write that landed.  Safe to repeat: any GET; a failure while the request was still being written; and a
        # kept-alive connection the server had already closed (it never saw the request).
        last: Optional[Exception] = None
        for attempt in range(2):
            conn = self._conn()
            reused = conn.sock is not None
            sent = False
            try:
                conn.request(method, url, body=data, headers=headers)
                sent = True
                resp = conn.getresponse()
                raw = resp.read()
                break
            except (OSError, http.client.HTTPException) as e:
                last = e
                conn.close()
                self._local.conn = None
                stale_keepalive = reused and isinstance(e, (http.client.RemoteDisconnected, ConnectionResetError,
                                                            BrokenPipeError))
                if method != "GET" and sent and not stale_keepalive:
                    timed_out = isinstance(e, (socket.timeout, TimeoutError))
                    raise APIError(504 if timed_out else 503, "Timeout" if timed_out else "ServiceUnavailable",
                                   f"{method} {path}: no answer from API server {self.master} ({e}); the request may "
                                   f"have been applied") from None
        else:
            raise APIError(503, "ServiceUnavailable", f"cannot reach API server {self.master}: {last}")
        out = json.loads(raw) if raw else {}
        if resp.status >= 400:
            if isinstance(out, dict) and out.get("kind") == "Status":
                raise APIError.from_status(out)
            raise APIError(resp.status, "InternalError", raw.decode(errors="replace")[:500])
        return out

    def create(self, info, namespace, obj):
        ns = namespace or (obj.get("metadata", {}).get("namespace", "") if info.namespaced else "")
        return self._request("POST", info.path(ns or ("default" if info.namespaced else "")), body=obj)

    def get(self, info, namespace, name):
        return self._request("GET", info.path(namespace or ("default" if info.namespaced else ""), name))

    def list(self, info, namespace="", label_selector="", field_selector=""):
        return self._request("GET", info.path(namespace), {"labelSelector": label_selector,
                                                           "fieldSelector": field_selector})

    def update(self, info, namespace, name, obj, subresource=""):
        p = info.path(namespace or ("default" if info.namespaced else ""), name)
        if subresource:
            p += "/" + subresource
        return self._request("PUT", p, body=obj)

    def patch(self, info, namespace, name, patch, patch_type="application/merge-patch+json", subresource=""):
        p = info.path(namespace or ("default" if info.namespaced else ""), name)
        if subresource:
            p += "/" + subresource
        return self._request("PATCH", p, body=patch, content_type=patch_type)

    def delete(self, info, namespace, name, grace_period_seconds=None, uid=""):
        body = {}
        if grace_period_seconds is not None:
            body["gracePeriodSeconds"] = grace_period_seconds
        if uid:
            body["preconditions"] = {"uid": uid}
        return self._request("DELETE", info.path(namespace or ("default" if info.namespaced else ""), name),
                             body=body or None)

    def delete_collection(self, info, namespace, label_selector="", grace_period_seconds=None):
        return self._request("DELETE", info.path(namespace), {"labelSelector": label_selector,
                                                              "gracePeriodSeconds": grace_period_seconds})

    def watch(self, info, namespace="", resource_version="", label_selector="", timeout=None):
        q = {"watch": "true", "resourceVersion": resource_version, "labelSelector": label_selector}
        if timeout is not None:
            q["timeoutSeconds"] = int(max(1, timeout))
        url = info.path(namespace) + "?" + urllib.parse.urlencode({k: v for k, v in q.items() if v not in (None, "")})
        conn = _NoDelayConnection(self._host, self._port, timeout=5.0)
        try:
            conn.request("GET", url, headers={"Accept": "application/json", "User-Agent": self._ua})
            resp = conn.getresponse()
        except (OSError, http.client.HTTPException) as e:
            conn.close()
            raise APIError(503, "ServiceUnavailable", f"cannot reach API server {self.master}: {e}") from None
        if resp.status >= 400:
            raw = resp.read()
            conn.close()
            try:
                raise APIError.from_status(json.loads(raw))
            except ValueError:
                raise APIError(resp.status, "InternalError", raw.decode(errors="replace")[:300]) from None
        return _HTTPWatch(conn, resp)
