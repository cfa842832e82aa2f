"""Operator options and flags.

Parity: /root/reference/cmd/app/options/options.go:12-72 -- same flag names and defaults
(``--master``, ``--kubeconfig``, ``--run-in-cluster``, ``--thread-num`` 1, ``--namespace`` all,
``--resync-period`` 10s, ``--creating-restart-period`` 0, ``--creating-duration-period`` 15m,
``--enable-creating-failed`` false) plus the upstream ``leaderelectionconfig.BindFlags`` set
(``--leader-elect``, ``--leader-elect-lease-duration`` 15s, ``--leader-elect-renew-deadline`` 5s,
``--leader-elect-retry-period`` 3s, ``--leader-elect-resource-lock`` endpoints) and klog's
``--v`` / ``--logtostderr`` (README.md:11).  ``--master`` / ``--kubeconfig`` point at the local
API server instead of a kube-apiserver.  New flags are grouped at the end.
"""
from __future__ import annotations

import argparse
import os
import re
from dataclasses import dataclass, field
from typing import Optional

_DUR = re.compile(r"(\d+(?:\.\d+)?)(ns|us|µs|ms|s|m|h)")
_UNIT = {"ns": 1e-9, "us": 1e-6, "µs": 1e-6, "ms": 1e-3, "s": 1.0, "m": 60.0, "h": 3600.0}


def parse_duration(s) -> float:
    """Go ``time.ParseDuration`` subset: ``10s``, ``15m``, ``1h30m``, ``500ms``; bare numbers = seconds."""
    if isinstance(s, (int, float)):
        return float(s)
    s = str(s).strip()
    if not s:
        return 0.0
    try:
        return float(s)
    except ValueError:
        pass
    pos, total = 0, 0.0
    for m in _DUR.finditer(s):
        if m.start() != pos:
            raise ValueError(f"invalid duration {s!r}")
        total += float(m.group(1)) * _UNIT[m.group(2)]
        pos = m.end()
    if pos != len(s):
        raise ValueError(f"invalid duration {s!r}")
    return total


def _bool(v) -> bool:
    if isinstance(v, bool):
        return v
    return str(v).lower() in ("1", "true", "t", "yes", "y")


@dataclass
class LeaderElectionConfiguration:
    leader_elect: bool = False
    lease_duration: float = 15.0
    renew_deadline: float = 5.0
    retry_period: float = 3.0
    resource_lock: str = "endpoints"


@dataclass
class TrainingJobOperatorOption:
    master_url: str = ""
    kubeconfig: str = ""
    run_in_cluster: bool = False
    thread_num: int = 1
    creating_restart_time: float = 0.0
    creating_duration_time: float = 15 * 60.0
    enable_creating_failed: bool = False
    namespace: str = ""              # v1.NamespaceAll
    resync_period: float = 10.0
    leader_election: LeaderElectionConfiguration = field(default_factory=LeaderElectionConfiguration)
    # klog
    v: int = 0
    logtostderr: bool = True
    # --- new in this framework ---------------------------------------------------------------
    gc_interval: float = 600.0       # orphan sweep period (controller.go:204 hard-codes 10 min)
    scale_down_grace: float = 30.0   # how long an out-of-range replica may drain before deletion
    identity: str = ""               # leader-election identity override (tests)
    metrics_port: int = 0
    # overall token bucket of the reconcile queue.  The reference inherits client-go's DefaultControllerRateLimiter
    # (10 qps / burst 100), and because every status write of the controller re-enqueues its job rate-limited, a busy
    # operator degrades to one reconcile per 100 ms; here the store is in-process, so the bucket is sized for it.
    queue_qps: float = 2000.0
    queue_burst: int = 2000
    # client-side API throttle per clientset (client-go: 5 qps / burst 10, which the reference runs with); 0 = none
    kube_api_qps: float = 0.0
    kube_api_burst: int = 0
    # reproduce the reference's per-pass API behaviour: one live Node LIST per role per reconcile (pod.go:181,441)
    # instead of the node informer
    live_node_list: bool = False


def new_training_job_operator_option() -> TrainingJobOperatorOption:
    return TrainingJobOperatorOption()


def add_flags(parser: argparse.ArgumentParser, opt: Optional[TrainingJobOperatorOption] = None) -> None:
    o = opt or TrainingJobOperatorOption()
    a = parser.add_argument
    a("--master", dest="master_url", default=o.master_url,
      help="The address of the API server (http://127.0.0.1:PORT). Overrides any value in kubeconfig.")
    a("--kubeconfig", default=o.kubeconfig, help="Path to a kubeconfig (YAML with clusters[0].cluster.server).")
    a("--run-in-cluster", type=_bool, nargs="?", const=True, default=o.run_in_cluster,
      help="TrainingJob Operator run in cluster or out of cluster (in-cluster: AITJ_MASTER from the environment).")
    a("--thread-num", type=int, default=o.thread_num, help="The num of worker thread")
    a("--namespace", default=o.namespace, help="The namespace to monitor trainingjobs. Default all namespaces.")
    a("--resync-period", type=parse_duration, default=o.resync_period, help="Resync interval of the trainingjob operator.")
    a("--creating-restart-period", dest="creating_restart_time", type=parse_duration, default=o.creating_restart_time,
      help="The period time of retrying to create container")
    a("--creating-duration-period", dest="creating_duration_time", type=parse_duration,
      default=o.creating_duration_time, help="The period time of creating container")
    a("--enable-creating-failed", type=_bool, nargs="?", const=True, default=o.enable_creating_failed,
      help="set job failed if containers have been creating exceed creating-restart-period.")
    le = o.leader_election
    a("--leader-elect", type=_bool, nargs="?", const=True, default=le.leader_elect,
      help="Start a leader election client and gain leadership before executing the main loop.")
    a("--leader-elect-lease-duration", type=parse_duration, default=le.lease_duration)
    a("--leader-elect-renew-deadline", type=parse_duration, default=le.renew_deadline)
    a("--leader-elect-retry-period", type=parse_duration, default=le.retry_period)
    a("--leader-elect-resource-lock", default=le.resource_lock, help="endpoints | leases")
    a("--v", "-v", type=int, default=o.v, help="log level for V logs")
    a("--logtostderr", type=_bool, nargs="?", const=True, default=o.logtostderr)
    a("--gc-interval", type=parse_duration, default=o.gc_interval, help="orphan garbage-collection period")
    a("--scale-down-grace", type=parse_duration, default=o.scale_down_grace)
    a("--identity", default=o.identity, help="leader election identity (default <hostname>_<uuid>)")
    a("--metrics-port", type=int, default=o.metrics_port)
    a("--queue-qps", type=float, default=o.queue_qps,
      help="overall rate limit of the reconcile work queue (client-go default would be 10)")
    a("--queue-burst", type=int, default=o.queue_burst, help="burst of that limiter (client-go default would be 100)")
    a("--kube-api-qps", type=float, default=o.kube_api_qps,
      help="client-side request throttle per clientset (client-go / the reference: 5); 0 = unthrottled")
    a("--kube-api-burst", type=int, default=o.kube_api_burst, help="burst of that throttle (client-go / the reference: 10)")
    a("--live-node-list", type=_bool, nargs="?", const=True, default=o.live_node_list,
      help="LIST nodes from the API server once per role per reconcile, as the reference does, instead of using the informer")


def from_args(ns: argparse.Namespace) -> TrainingJobOperatorOption:
    o = TrainingJobOperatorOption()
    for f in ("master_url", "kubeconfig", "run_in_cluster", "thread_num", "creating_restart_time",
              "creating_duration_time", "enable_creating_failed", "namespace", "resync_period", "v", "logtostderr",
              "gc_interval", "scale_down_grace", "identity", "metrics_port", "queue_qps", "queue_burst",
              "kube_api_qps", "kube_api_burst", "live_node_list"):
        setattr(o, f, getattr(ns, f))
    o.leader_election = LeaderElectionConfiguration(
        leader_elect=ns.leader_elect, lease_duration=ns.leader_elect_lease_duration,
        renew_deadline=ns.leader_elect_renew_deadline, retry_period=ns.leader_elect_retry_period,
        resource_lock=ns.leader_elect_resource_lock)
    return o


def resolve_master(opt: TrainingJobOperatorOption) -> str:
    """``clientcmd.BuildConfigFromFlags(master, kubeconfig)`` / ``InClusterConfig`` analogue."""
    if opt.run_in_cluster:
        m = os.environ.get("AITJ_MASTER") or os.environ.get("KUBERNETES_SERVICE_HOST")
        if not m:
            raise RuntimeError("--run-in-cluster set but AITJ_MASTER is not in the environment")
        return m
    if opt.master_url:
        return opt.master_url
    path = opt.kubeconfig or os.environ.get("KUBECONFIG") or os.path.expanduser("~/.aitj/config")
    if os.path.exists(path):
        import yaml

        cfg = yaml.safe_load(open(path)) or {}
        clusters = cfg.get("clusters") or []
        if clusters:
            return clusters[0].get("cluster", {}).get("server", "")
        if cfg.get("server"):
            return cfg["server"]
    if os.environ.get("AITJ_MASTER"):
        return os.environ["AITJ_MASTER"]
    raise RuntimeError("no API server configured: pass --master or --kubeconfig")
