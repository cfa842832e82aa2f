"""Operator bootstrap: clientsets -> informer factories -> controller -> (leader-elected) run.

Parity: /root/reference/cmd/app/server.go:26-157 -- signal handler, client sets (:111-151), two
informer factories with the resync period and namespace scope (:43-44), controller construction (:47),
informers started *before* leadership so a standby has warm caches (:50-51 precede :94), run directly
when ``--leader-elect=false`` (:74-77) or under ``leaderelection.RunOrDie`` with an Endpoints lock
``kube-system/trainingjob-operator`` and identity ``<hostname>_<uuid>`` (:79-106); losing the lease is
fatal (:101-103).  Worker processes are children of the *agent*, never of this process, so the
operator's suicide on a lost lease does not touch running replicas.
"""
from __future__ import annotations

import threading
from typing import Optional

from ..client.clientset import new_for_config
from ..client.informers import SharedInformerFactory
from ..client.leaderelection import LeaderElectionConfig, LeaderElector, default_identity
from ..client.record import EventRecorder
from ..controller.controller import new_training_job_controller
from ..utils import klog
from .options import TrainingJobOperatorOption, resolve_master


class LeaseLost(RuntimeError):
    pass


def create_client_sets(opt: TrainingJobOperatorOption, server=None):
    """server.go:111-151: kube, leader-election, trainingjob and apiextensions clients."""
    if server is not None:
        base = lambda: new_for_config(server=server)  # noqa: E731
    else:
        master = resolve_master(opt)
        base = lambda: new_for_config(master=master)  # noqa: E731

    def mk():
        cs = base()
        if getattr(opt, "kube_api_qps", 0) and opt.kube_api_qps > 0:
            from ..store.transport import ThrottledTransport

            cs.transport = ThrottledTransport(cs.transport, opt.kube_api_qps, opt.kube_api_burst or 2 * int(opt.kube_api_qps))
        return cs

    return mk(), mk(), mk(), mk()


def build_controller(opt: TrainingJobOperatorOption, server=None, stop: Optional[threading.Event] = None):
    kube_client, le_client, tj_client, ext_client = create_client_sets(opt, server)
    kube_factory = SharedInformerFactory(kube_client, opt.resync_period, opt.namespace)
    tj_factory = SharedInformerFactory(tj_client, opt.resync_period, opt.namespace)
    tc = new_training_job_controller(kube_client, tj_client, ext_client, kube_factory, tj_factory, opt)
    return tc, kube_factory, tj_factory, le_client


def run(opt: TrainingJobOperatorOption, stop: Optional[threading.Event] = None, server=None,
        fatal_on_lost_lease: bool = True) -> None:
    """``app.Run`` (server.go:26-109).  Blocks until ``stop``; raises ``LeaseLost`` (or exits) on a lost lease."""
    if opt.namespace == "":
        klog.info("List-watch resources across all namespaces")
    else:
        klog.info("List-watch resources to namespace %s", opt.namespace)
    if stop is None:
        from ..signals import setup_signal_handler

        stop = setup_signal_handler()
    tc, kube_factory, tj_factory, le_client = build_controller(opt, server, stop)
    kube_factory.start(stop)
    tj_factory.start(stop)

    def run_controller(lead_stop: threading.Event) -> None:
        klog.V(4).info("I won the leader election")
        both = threading.Event()
        threading.Thread(target=lambda: (lead_stop.wait(), both.set()), daemon=True).start()
        threading.Thread(target=lambda: (stop.wait(), both.set()), daemon=True).start()
        try:
            tc.run(opt.thread_num, both)
        except Exception as e:  # noqa: BLE001
            klog.error("Failed to run the controller: %r", e)
            stop.set()

    le = opt.leader_election
    if not le.leader_elect:
        run_controller(stop)
        return

    cfg = LeaderElectionConfig(lock_type=le.resource_lock, lock_namespace="kube-system",
                               lock_name="trainingjob-operator", identity=opt.identity or default_identity(),
                               lease_duration=le.lease_duration, renew_deadline=le.renew_deadline,
                               retry_period=le.retry_period,
                               release_on_cancel=lambda: not getattr(opt, "crash_on_stop", False))
    lost = threading.Event()

    def on_stopped() -> None:
        if not stop.is_set():
            lost.set()

    recorder = EventRecorder(le_client, "trainingjob-operator")
    LeaderElector(le_client, cfg, run_controller, on_stopped, recorder=recorder).run(stop)
    if lost.is_set():
        if fatal_on_lost_lease:
            klog.fatal("leaderelection lost")
        raise LeaseLost("leaderelection lost")
