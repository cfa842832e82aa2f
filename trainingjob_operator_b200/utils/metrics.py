"""Process-wide metrics registry rendered in the Prometheus text format at ``/metrics``.

The reference exports no metrics at all (no HTTP server; prometheus is only an indirect
dependency: go.mod:50).  Reconcile latency, queue depth, spawn->running and
rescale latency are first-class measured quantities here because BASELINE.json names them.
"""
from __future__ import annotations

import bisect
import threading
from typing import Dict, List, Optional, Tuple

# Writers never take a lock: every thread updates its own shard (counters and histograms are merged when somebody reads
# them), gauges are single dict assignments.  A process-wide lock here was a convoy point under load -- eight reconcile
# workers bumping a histogram per sync each paid a GIL hand-over per contended acquisition.
_LOCK = threading.Lock()                      # guards the shard registry and readers
_Key = Tuple[str, Tuple[Tuple[str, str], ...]]
_GAUGES: Dict[_Key, float] = {}
_HELP: Dict[str, str] = {}
_SHARDS: List["_Shard"] = []
_LOCAL = threading.local()

DEFAULT_BUCKETS = (0.0005, 0.001, 0.0025, 0.005, 0.01, 0.025, 0.05, 0.1, 0.25, 0.5, 1, 2.5, 5, 10, 30, 60)


class _Hist:
    def __init__(self, buckets):
        self.buckets = list(buckets)
        self.counts = [0] * (len(self.buckets) + 1)
        self.sum = 0.0
        self.n = 0
        self.samples: List[float] = []

    def observe(self, v: float) -> None:
        self.counts[bisect.bisect_left(self.buckets, v)] += 1
        self.sum += v
        self.n += 1
        if len(self.samples) < 4096:
            self.samples.append(v)

    def merge(self, other: "_Hist") -> None:
        for i, c in enumerate(list(other.counts)):
            self.counts[i] += c
        self.sum += other.sum
        self.n += other.n
        self.samples.extend(other.samples[:max(0, 4096 - len(self.samples))])


class _Shard:
    def __init__(self):
        self.counters: Dict[_Key, float] = {}
        self.hists: Dict[_Key, _Hist] = {}


def _shard() -> _Shard:
    sh = getattr(_LOCAL, "shard", None)
    if sh is None:
        sh = _LOCAL.shard = _Shard()
        with _LOCK:
            _SHARDS.append(sh)                # kept after the thread ends: its counts still belong to the totals
    return sh


def _key(name: str, labels: Optional[Dict[str, str]]):
    return name, tuple(sorted((labels or {}).items()))


def describe(name: str, help_text: str) -> None:
    _HELP[name] = help_text


def inc(name: str, value: float = 1.0, labels: Optional[Dict[str, str]] = None) -> None:
    c = _shard().counters
    k = _key(name, labels)
    c[k] = c.get(k, 0.0) + value


def set_gauge(name: str, value: float, labels: Optional[Dict[str, str]] = None) -> None:
    _GAUGES[_key(name, labels)] = float(value)


def observe(name: str, value: float, labels: Optional[Dict[str, str]] = None, buckets=DEFAULT_BUCKETS) -> None:
    hs = _shard().hists
    k = _key(name, labels)
    h = hs.get(k)
    if h is None:
        h = hs[k] = _Hist(buckets)
    h.observe(value)


def _merged_counters() -> Dict[_Key, float]:
    out: Dict[_Key, float] = {}
    with _LOCK:
        shards = list(_SHARDS)
    for sh in shards:
        for k, v in list(sh.counters.items()):
            out[k] = out.get(k, 0.0) + v
    return out


def _merged_hists() -> Dict[_Key, _Hist]:
    out: Dict[_Key, _Hist] = {}
    with _LOCK:
        shards = list(_SHARDS)
    for sh in shards:
        for k, h in list(sh.hists.items()):
            m = out.get(k)
            if m is None:
                m = out[k] = _Hist(h.buckets)
            m.merge(h)
    return out


def quantile(name: str, q: float, labels: Optional[Dict[str, str]] = None) -> Optional[float]:
    h = _merged_hists().get(_key(name, labels))
    if not h or not h.samples:
        return None
    s = sorted(h.samples)
    return s[min(len(s) - 1, int(q * len(s)))]


def get_counter(name: str, labels: Optional[Dict[str, str]] = None) -> float:
    return _merged_counters().get(_key(name, labels), 0.0)


def reset() -> None:
    with _LOCK:
        for sh in _SHARDS:
            sh.counters.clear()
            sh.hists.clear()
    _GAUGES.clear()


def _fmt_labels(labels) -> str:
    if not labels:
        return ""
    return "{" + ",".join(f'{k}="{v}"' for k, v in labels) + "}"


def render(extra: Optional[dict] = None) -> str:
    lines: List[str] = []
    counters, hists = _merged_counters(), _merged_hists()
    seen = set()
    for (name, labels), v in sorted(counters.items()):
        if name not in seen:
            seen.add(name)
            lines.append(f"# HELP {name} {_HELP.get(name, name)}")
            lines.append(f"# TYPE {name} counter")
        lines.append(f"{name}{_fmt_labels(labels)} {v}")
    for (name, labels), v in sorted(list(_GAUGES.items())):
        if name not in seen:
            seen.add(name)
            lines.append(f"# HELP {name} {_HELP.get(name, name)}")
            lines.append(f"# TYPE {name} gauge")
        lines.append(f"{name}{_fmt_labels(labels)} {v}")
    for (name, labels), h in sorted(hists.items(), key=lambda kv: kv[0]):
        if name not in seen:
            seen.add(name)
            lines.append(f"# HELP {name} {_HELP.get(name, name)}")
            lines.append(f"# TYPE {name} histogram")
        acc = 0
        for b, c in zip(h.buckets, h.counts):
            acc += c
            lines.append(f"{name}_bucket{_fmt_labels(labels + (('le', str(b)),))} {acc}")
        lines.append(f"{name}_bucket{_fmt_labels(labels + (('le', '+Inf'),))} {h.n}")
        lines.append(f"{name}_sum{_fmt_labels(labels)} {h.sum}")
        lines.append(f"{name}_count{_fmt_labels(labels)} {h.n}")
    if extra:
        lines.append("# TYPE aitj_apiserver_resource_version gauge")
        lines.append(f"aitj_apiserver_resource_version {extra.get('resourceVersion', 0)}")
        lines.append(f"aitj_apiserver_watchers {extra.get('watchers', 0)}")
        lines.append(f"aitj_apiserver_requests_total {extra.get('requests', 0)}")
        for kind, n in (extra.get("objects") or {}).items():
            lines.append(f'aitj_apiserver_objects{{kind="{kind}"}} {n}')
    return "\n".join(lines) + "\n"
