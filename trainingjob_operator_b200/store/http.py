"""HTTP façade of the API server: a wire-compatible subset of the Kubernetes REST API.

Serves ``/api/v1/...`` (pods, services, events, nodes, endpoints) and
``/apis/elasticdeeplearning.ai/v1/namespaces/{ns}/aitrainingjobs[/name[/status]]`` with
``?watch=true&resourceVersion=&labelSelector=&timeoutSeconds=`` streams, discovery documents,
``/healthz``, ``/version`` and Prometheus ``/metrics`` (the reference has no metrics endpoint,
SURVEY.md §5.5; this one exports queue / reconcile / spawn latency series registered by the
controller and agent).  The reference's generated REST client targets exactly these paths
(pkg/client/clientset/versioned/typed/aitrainingjob/v1/aitrainingjob.go:66-190).
"""
from __future__ import annotations

import http.client
import json
import re
import socket
import threading
import urllib.parse
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Optional, Tuple

from ..api import register as R
from ..utils import metrics as metrics_mod
from .apiserver import APIError, APIServer
from .fasthttp import read_headers

_PATH = re.compile(
    r"^/(?:api/(?P<corev>v1)|apis/(?P<group>[^/]+)/(?P<version>[^/]+))"
    r"(?:/namespaces/(?P<ns>[^/]+))?/(?P<plural>[^/]+)(?:/(?P<name>[^/]+))?(?:/(?P<sub>status))?$")


def _resolve(path: str) -> Optional[Tuple[R.ResourceInfo, str, str, str]]:
    m = _PATH.match(path)
    if not m:
        return None
    info = R.lookup(m.group("plural"))
    if info is None:
        return None
    group = m.group("group") or ""
    if info.group != group:
        return None
    return info, m.group("ns") or "", m.group("name") or "", m.group("sub") or ""


class _Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"
    server_version = "aitj-apiserver/1.0"
    # One segment per response: with the default unbuffered wfile the status line + headers and the body left as two
    # small writes, and Nagle held the second back until the client's delayed ACK (40 ms) -- every API call of the
    # separately running agent / operator / CLI paid it.  Buffered writes (flushed once per response, and after every
    # watch event) plus TCP_NODELAY on both ends.
    wbufsize = 64 * 1024

    def setup(self):
        super().setup()
        try:
            self.request.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        except OSError:
            pass

    def log_message(self, fmt, *args):  # silence stderr access log
        pass

    def parse_request(self) -> bool:
        """``BaseHTTPRequestHandler.parse_request`` with the lean header parser (``store/fasthttp.py``): the stock one
        runs the e-mail package over every request's header block, a third of this process' CPU time under load.
        HTTP/1.0 and 1.1 request lines only (no HTTP/0.9 simple requests)."""
        self.command = None
        self.request_version = self.default_request_version
        self.close_connection = True
        requestline = str(self.raw_requestline, "iso-8859-1").rstrip("\r\n")
        self.requestline = requestline
        words = requestline.split()
        if not words:
            return False
        if len(words) != 3 or words[2] not in ("HTTP/1.1", "HTTP/1.0"):
            self.request_version = "HTTP/1.1"         # answer with a status line (an HTTP/0.9 reply has none)
            self.send_error(400, f"Bad request syntax ({requestline!r})")
            return False
        self.command, self.path, self.request_version = words
        if self.path.startswith("//"):
            self.path = "/" + self.path.lstrip("/")
        try:
            self.headers = read_headers(self.rfile)
        except http.client.LineTooLong as err:
            self.send_error(431, "Line too long", str(err))
            return False
        except http.client.HTTPException as err:
            self.send_error(431, "Too many headers", str(err))
            return False
        conn = self.headers.get("connection", "").lower()
        if conn == "close":
            self.close_connection = True
        elif self.request_version == "HTTP/1.1" or conn == "keep-alive":
            self.close_connection = False
        if self.headers.get("expect", "").lower() == "100-continue" and self.request_version == "HTTP/1.1":
            if not self.handle_expect_100():
                return False
        return True

    # ---------------------------------------------------------------- helpers
    @property
    def api(self) -> APIServer:
        return self.server.api  # type: ignore[attr-defined]

    def _send_json(self, code: int, obj: Any) -> None:
        raw = json.dumps(obj).encode()
        self.send_response(code)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(raw)))
        self.end_headers()
        self.wfile.write(raw)

    def _send_raw_json(self, code: int, raw: bytes) -> None:
        self.send_response(code)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(raw)))
        self.end_headers()
        self.wfile.write(raw)

    def _send_text(self, code: int, text: str, ctype: str = "text/plain; charset=utf-8") -> None:
        raw = text.encode()
        self.send_response(code)
        self.send_header("Content-Type", ctype)
        self.send_header("Content-Length", str(len(raw)))
        self.end_headers()
        self.wfile.write(raw)

    def _error(self, e: APIError) -> None:
        self._send_json(e.code, e.status())

    def _body(self) -> Any:
        n = int(self.headers.get("Content-Length") or 0)
        if n <= 0:
            return None
        raw = self.rfile.read(n)
        ctype = (self.headers.get("Content-Type") or "").split(";")[0].strip()
        if ctype in ("application/yaml", "application/x-yaml", "text/yaml"):
            import yaml

            return yaml.safe_load(raw)
        return json.loads(raw)

    # ---------------------------------------------------------------- discovery & misc
    def _discovery(self, path: str) -> bool:
        if path in ("/healthz", "/readyz", "/livez"):
            self._send_text(200, "ok")
            return True
        if path == "/version":
            self._send_json(200, {"major": "1", "minor": "13", "gitVersion": "v1.13.5-aitj-b200", "platform": "linux/amd64"})
            return True
        if path == "/metrics":
            self._send_text(200, metrics_mod.render(extra=self.api.stats()), "text/plain; version=0.0.4")
            return True
        if path == "/api":
            self._send_json(200, {"kind": "APIVersions", "versions": ["v1"]})
            return True
        if path == "/apis":
            groups = {}
            for r in R.all_resources():
                if r.group:
                    groups.setdefault(r.group, set()).add(r.version)
            self._send_json(200, {"kind": "APIGroupList", "apiVersion": "v1", "groups": [
                {"name": g, "versions": [{"groupVersion": f"{g}/{v}", "version": v} for v in sorted(vs)],
                 "preferredVersion": {"groupVersion": f"{g}/{sorted(vs)[0]}", "version": sorted(vs)[0]}}
                for g, vs in sorted(groups.items())]})
            return True
        m = re.match(r"^/(?:api/(v1)|apis/([^/]+)/([^/]+))$", path)
        if m:
            group = m.group(2) or ""
            version = m.group(1) or m.group(3)
            res = [r for r in R.all_resources() if r.group == group and r.version == version]
            self._send_json(200, {"kind": "APIResourceList", "apiVersion": "v1",
                                  "groupVersion": f"{group}/{version}" if group else version,
                                  "resources": [{"name": r.plural, "singularName": r.kind.lower(), "namespaced": r.namespaced,
                                                 "kind": r.kind, "shortNames": list(r.short_names),
                                                 "verbs": ["create", "delete", "deletecollection", "get", "list", "patch",
                                                           "update", "watch"]} for r in res]})
            return True
        return False

    # ---------------------------------------------------------------- verbs
    def do_GET(self):  # noqa: N802
        u = urllib.parse.urlparse(self.path)
        q = {k: v[-1] for k, v in urllib.parse.parse_qs(u.query).items()}
        try:
            if self._discovery(u.path):
                return
            r = _resolve(u.path)
            if r is None:
                raise APIError(404, "NotFound", f"the server could not find the requested resource ({u.path})")
            info, ns, name, _sub = r
            if name:
                self._send_raw_json(200, self.api.get_raw(info, ns, name))
            elif q.get("watch") in ("true", "1"):
                self._watch(info, ns, q)
            else:
                self._send_json(200, self.api.list(info, ns, q.get("labelSelector", ""), q.get("fieldSelector", "")))
        except APIError as e:
            self._error(e)
        except (BrokenPipeError, ConnectionResetError):
            pass

    def _watch(self, info, ns, q) -> None:
        timeout = float(q["timeoutSeconds"]) if q.get("timeoutSeconds") else None
        stream = self.api.watch(info, ns, q.get("resourceVersion", ""), q.get("labelSelector", ""), timeout)
        self.send_response(200)
        self.send_header("Content-Type", "application/json")
        self.send_header("Transfer-Encoding", "chunked")
        self.end_headers()
        self.wfile.flush()
        stopping = self.server.stopping  # type: ignore[attr-defined]
        try:
            while not stream.expired and not stopping.is_set():
                raw = stream.poll_raw_many(0.25)
                if not raw:
                    continue
                self.wfile.write(f"{len(raw):x}\r\n".encode() + raw + b"\r\n")
                self.wfile.flush()
            self.wfile.write(b"0\r\n\r\n")
            self.wfile.flush()
        except (BrokenPipeError, ConnectionResetError, OSError):
            pass
        finally:
            stream.close()
            self.close_connection = True

    def do_POST(self):  # noqa: N802
        u = urllib.parse.urlparse(self.path)
        try:
            r = _resolve(u.path)
            if r is None or r[2]:
                raise APIError(404, "NotFound", f"the server could not find the requested resource ({u.path})")
            info, ns, _name, _sub = r
            self._send_raw_json(201, self.api.create(info, ns, self._body() or {}, as_bytes=True))
        except APIError as e:
            self._error(e)
        except (ValueError, KeyError) as e:
            self._error(APIError(400, "BadRequest", f"malformed body: {e}"))

    def do_PUT(self):  # noqa: N802
        u = urllib.parse.urlparse(self.path)
        try:
            r = _resolve(u.path)
            if r is None or not r[2]:
                raise APIError(404, "NotFound", f"the server could not find the requested resource ({u.path})")
            info, ns, name, sub = r
            self._send_raw_json(200, self.api.update(info, ns, name, self._body() or {}, sub, as_bytes=True))
        except APIError as e:
            self._error(e)
        except (ValueError, KeyError) as e:
            self._error(APIError(400, "BadRequest", f"malformed body: {e}"))

    def do_PATCH(self):  # noqa: N802
        u = urllib.parse.urlparse(self.path)
        try:
            r = _resolve(u.path)
            if r is None or not r[2]:
                raise APIError(404, "NotFound", f"the server could not find the requested resource ({u.path})")
            info, ns, name, sub = r
            ctype = (self.headers.get("Content-Type") or "application/merge-patch+json").split(";")[0].strip()
            if ctype == "application/strategic-merge-patch+json":
                ctype = "application/merge-patch+json"
            self._send_raw_json(200, self.api.patch(info, ns, name, self._body(), ctype, sub, as_bytes=True))
        except APIError as e:
            self._error(e)
        except (ValueError, KeyError, IndexError) as e:
            self._error(APIError(400, "BadRequest", f"malformed patch: {e}"))

    def do_DELETE(self):  # noqa: N802
        u = urllib.parse.urlparse(self.path)
        q = {k: v[-1] for k, v in urllib.parse.parse_qs(u.query).items()}
        try:
            r = _resolve(u.path)
            if r is None:
                raise APIError(404, "NotFound", f"the server could not find the requested resource ({u.path})")
            info, ns, name, _sub = r
            body = self._body() or {}
            grace = body.get("gracePeriodSeconds", q.get("gracePeriodSeconds"))
            grace = int(grace) if grace not in (None, "") else None
            if name:
                uid = (body.get("preconditions") or {}).get("uid", "")
                self._send_json(200, self.api.delete(info, ns, name, grace, uid))
            else:
                self._send_json(200, self.api.delete_collection(info, ns, q.get("labelSelector", ""), grace))
        except APIError as e:
            self._error(e)


class _QuietServer(ThreadingHTTPServer):
    request_queue_size = 128          # listen backlog: --thread-num 1000 operators connect at once

    def handle_error(self, request, client_address):
        import sys

        exc = sys.exc_info()[1]
        if isinstance(exc, (ConnectionResetError, BrokenPipeError, TimeoutError)):
            return                    # a client went away mid-request (a killed worker): not worth a traceback
        super().handle_error(request, client_address)


class APIHTTPServer:
    """Threaded HTTP server bound to loopback; ``start()`` returns once it is listening."""

    def __init__(self, api: APIServer, host: str = "127.0.0.1", port: int = 0):
        self.api = api
        self._httpd = _QuietServer((host, port), _Handler)
        self._httpd.daemon_threads = True
        self._httpd.api = api  # type: ignore[attr-defined]
        self._httpd.stopping = threading.Event()  # type: ignore[attr-defined]
        self._thread: Optional[threading.Thread] = None

    @property
    def port(self) -> int:
        return self._httpd.server_address[1]

    @property
    def url(self) -> str:
        return f"http://{self._httpd.server_address[0]}:{self.port}"

    def start(self) -> "APIHTTPServer":
        self._thread = threading.Thread(target=self._httpd.serve_forever, kwargs={"poll_interval": 0.1},
                                        name="aitj-apiserver", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._httpd.stopping.set()  # type: ignore[attr-defined]
        self._httpd.shutdown()
        self._httpd.server_close()
        if self._thread:
            self._thread.join(timeout=5)
