"""Stand-alone API server daemon (``python -m trainingjob_operator_b200.cmd.apiserver --port 8001``)."""
from __future__ import annotations

import argparse
import os
import sys

from ..signals import setup_signal_handler
from ..store.apiserver import APIServer
from ..store.http import APIHTTPServer
from ..utils import klog


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="aitj-apiserver")
    ap.add_argument("--port", type=int, default=8001)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--data-dir", default=os.path.expanduser("~/.aitj"))
    ap.add_argument("--no-wal", action="store_true")
    ap.add_argument("--event-ttl", type=float, default=3600.0, help="seconds an Event is kept (kube-apiserver: 1h)")
    ap.add_argument("--wal-compact-mb", type=float, default=64.0,
                    help="rewrite the write-ahead log as a snapshot once it is larger than this")
    ap.add_argument("--v", type=int, default=0)
    args = ap.parse_args(argv)
    klog.configure(args.v, True)
    os.makedirs(args.data_dir, exist_ok=True)
    api = APIServer("" if args.no_wal else os.path.join(args.data_dir, "store.wal"))
    srv = APIHTTPServer(api, args.host, args.port).start()
    print(f"aitj-apiserver listening on {srv.url}", flush=True)
    stop = setup_signal_handler()
    api.start_housekeeping(stop, event_ttl_s=args.event_ttl, wal_max_bytes=int(args.wal_compact_mb * (1 << 20)))
    stop.wait()
    srv.stop()
    return 0


if __name__ == "__main__":
    sys.exit(main())
