"""Process-lifetime bookkeeping for background threads that block inside the native core.

Daemon threads parked in a GIL-released C++ call (work-queue ``get``, store ``watch_next``, supervisor
``poll_exits``) must be gone before the interpreter finalises -- CPython unwinds such threads with
``pthread_exit`` when they try to re-take the GIL, and a forced unwind through pybind11's dispatcher
aborts the process.  Components register a stop callable and their threads here; an ``atexit`` hook
stops and joins them.
"""
from __future__ import annotations

import atexit
import threading
from typing import Callable, List

_LOCK = threading.Lock()
_STOPPERS: List[Callable[[], None]] = []
_THREADS: List[threading.Thread] = []


def register_stop(fn: Callable[[], None]) -> None:
    with _LOCK:
        _STOPPERS.append(fn)


def track(t: threading.Thread) -> threading.Thread:
    with _LOCK:
        _THREADS.append(t)
        if len(_THREADS) > 512:
            _THREADS[:] = [x for x in _THREADS if x.is_alive()]
    return t


def spawn(target, name: str, args=()) -> threading.Thread:
    t = threading.Thread(target=target, args=args, name=name, daemon=True)
    t.start()
    return track(t)


def shutdown(timeout: float = 2.0) -> None:
    with _LOCK:
        stoppers, threads = list(_STOPPERS), list(_THREADS)
        _STOPPERS.clear()
    for fn in stoppers:
        try:
            fn()
        except Exception:  # noqa: BLE001
            pass
    for t in threads:
        if t.is_alive() and t is not threading.current_thread():
            t.join(timeout)


atexit.register(shutdown)
