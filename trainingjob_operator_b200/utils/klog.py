"""klog-style levelled logging (``--v N --logtostderr``), as used by the reference binary
(README.md:11; 3 call sites at V(2), 17 at V(4): SURVEY.md §5.5)."""
from __future__ import annotations

import logging
import sys

_VERBOSITY = 0
_LOGGER = logging.getLogger("aitj")
_CONFIGURED = False


def configure(v: int = 0, logtostderr: bool = True, log_file: str = "") -> None:
    global _VERBOSITY, _CONFIGURED
    _VERBOSITY = int(v)
    if _CONFIGURED:
        return
    handler: logging.Handler
    if log_file:
        handler = logging.FileHandler(log_file)
    else:
        handler = logging.StreamHandler(sys.stderr if logtostderr else sys.stdout)
    handler.setFormatter(logging.Formatter("%(levelname).1s%(asctime)s.%(msecs)03d %(threadName)s] %(message)s",
                                           datefmt="%m%d %H:%M:%S"))
    _LOGGER.addHandler(handler)
    _LOGGER.setLevel(logging.INFO)
    _LOGGER.propagate = False
    _CONFIGURED = True


def set_verbosity(v: int) -> None:
    global _VERBOSITY
    _VERBOSITY = int(v)


def verbosity() -> int:
    return _VERBOSITY


def info(msg, *args):
    _LOGGER.info(msg, *args)


def warning(msg, *args):
    _LOGGER.warning(msg, *args)


def error(msg, *args):
    _LOGGER.error(msg, *args)


def fatal(msg, *args):
    _LOGGER.critical(msg, *args)
    logging.shutdown()
    raise SystemExit(255)


class _V:
    def __init__(self, level: int):
        self.enabled = level <= _VERBOSITY

    def info(self, msg, *args):
        if self.enabled:
            _LOGGER.info(msg, *args)

    def __bool__(self):
        return self.enabled


def V(level: int) -> _V:  # noqa: N802 - mirrors klog.V
    return _V(level)
