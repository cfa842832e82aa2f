"""Signal handling for the operator binaries.

Parity: /root/reference/pkg/signals/signal.go:29-43 (``SetupSignalHandler``: the first SIGINT/SIGTERM
closes the stop channel, the second exits with code 1; installing it twice panics) and
signal_posix.go:26 / signal_windows.go:23 (the signal set: SIGINT+SIGTERM on POSIX, SIGINT only on
Windows).
"""
from __future__ import annotations

import os
import signal
import sys
import threading

SHUTDOWN_SIGNALS = (signal.SIGINT,) if sys.platform.startswith("win") else (signal.SIGINT, signal.SIGTERM)

_installed = False


def setup_signal_handler() -> threading.Event:
    """Returns a stop event set on the first shutdown signal; a second signal exits(1)."""
    global _installed
    if _installed:
        raise RuntimeError("close of closed channel: SetupSignalHandler called twice")
    _installed = True
    stop = threading.Event()

    def handler(signum, frame):  # noqa: ARG001
        if stop.is_set():
            os._exit(1)  # second signal: exit directly
        stop.set()

    for s in SHUTDOWN_SIGNALS:
        signal.signal(s, handler)
    return stop


def _reset_for_tests() -> None:
    global _installed
    _installed = False
    for s in SHUTDOWN_SIGNALS:
        signal.signal(s, signal.SIG_DFL)
