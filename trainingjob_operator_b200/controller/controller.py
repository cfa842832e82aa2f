"""Controller core: wiring, work queue, sync handler, reconcile loop.

Parity: /root/reference/pkg/controller/controller.go:37-462:

* struct + constructor (controller.go:37-159): three informers (jobs, pods, services) with event
  handlers, event recorder, pod/service control, expectations, a named rate-limited work queue;
* ``Run`` (controller.go:182-208): self-register the CRD, wait for cache sync, N workers, GC, block;
* ``createCRD`` (controller.go:210-234; AlreadyExists tolerated);
* worker / processNextWorkItem (controller.go:236-268): Forget on success, AddRateLimited on error;
* ``syncHandler`` (controller.go:270-312): lister Get, NotFound => done, expectations gate, defaults,
  reconcile only non-deleted jobs in a reconcilable phase;
* ``reconcileTrainingJobs`` (controller.go:314-388): claim pods + services, per-role pods then services,
  ``Restarting`` => condition Terminating + RestartReplicaName + break, ending phases collected,
  aggregate messages, ``updateStatus``, write back only if the status changed;
* expectations check (controller.go:390-404), ``enqueueJob`` (controller.go:406-421), owner-ref
  resolution (controller.go:424-440), labels / owner reference (controller.go:161-180).

New behind the reference's unused fields: the rendezvous generation (elastic rescale, ``elastic.py``)
and the lifecycle trace annotation used to measure reconcile -> all-ranks-running latency.
"""
from __future__ import annotations

import json
import threading
import time
from typing import Dict, List, Optional

from ..api import constants as C
from ..api import meta as M
from ..api import register as R
from ..api.defaults import set_defaults_aitrainingjob
from ..api.types import AITrainingJob
from ..client.informers import DeletedFinalStateUnknown, SharedInformerFactory, deletion_handling_key, \
    wait_for_cache_sync
from ..client.record import EventRecorder
from ..cmd.options import TrainingJobOperatorOption
from ..core import _aitj_core as core
from ..store.apiserver import APIError
from ..utils import klog, lifecycle, metrics
from .control import RealPodControl, RealServiceControl
from .elastic import ElasticMixin
from .garbage_collection import GarbageCollector
from .pod import PodReconciler, gen_expectation_pods_key
from .service import ServiceReconciler, gen_expectation_services_key
from .status import StatusEngine, update_conditions
from .trainingjob import TrainingJobHandlers

metrics.describe("aitj_reconcile_seconds", "duration of one syncHandler pass")
metrics.describe("aitj_workqueue_depth", "keys waiting in the AITrainingJob work queue")
metrics.describe("aitj_job_startup_seconds", "job created -> all replicas Running")


INDEX_JOB_LABEL = "jobLabel"              # "<namespace>/<TrainingJobName label>"
INDEX_CONTROLLER_UID = "controllerUID"    # uid of the controlling owner reference


def index_by_job_label(obj: dict) -> list:
    v = M.labels_of(obj).get(C.LABEL_JOB_NAME)
    return [f"{M.namespace_of(obj)}/{v}"] if v else []


def index_by_controller_uid(obj: dict) -> list:
    ref = M.get_controller_of(obj)
    return [ref["uid"]] if ref is not None and ref.get("uid") else []


def claim_candidates(lister, job, selector) -> list:
    """Copies of the objects a ControllerRefManager pass of ``job`` can act on (pod.go:125-150, service.go:99-115 hand it
    the whole namespace): the ones matching the selector (keep / adopt) -- all of which carry the job-name label -- and
    the ones the job controls (release when the labels stopped matching)."""
    seen = {}
    for o in lister.by_index(INDEX_JOB_LABEL, f"{job.namespace}/{selector.get(C.LABEL_JOB_NAME, '')}"):
        if M.selector_matches(selector, M.labels_of(o)):
            seen[M.key_of(o)] = o
    for o in lister.by_index(INDEX_CONTROLLER_UID, job.uid):
        if M.namespace_of(o) == job.namespace:
            seen.setdefault(M.key_of(o), o)
    return [lister.copy_of(o) for o in seen.values()]


class TrainingJobController(PodReconciler, ServiceReconciler, StatusEngine, TrainingJobHandlers, ElasticMixin):
    kind = C.KIND
    group = C.GROUP_NAME

    def __init__(self, kube_client, trainingjob_client, ext_api_client, kube_informer_factory: SharedInformerFactory,
                 trainingjob_informer_factory: SharedInformerFactory, option: TrainingJobOperatorOption,
                 pod_control=None, service_control=None, recorder: Optional[EventRecorder] = None):
        self.kube_client = kube_client
        self.trainingjob_client = trainingjob_client
        self.api_extensions_client = ext_api_client
        self.option = option
        self.master_url = getattr(option, "master_url", "")

        job_informer = trainingjob_informer_factory.elasticdeeplearning().v1().aitrainingjobs()
        pod_informer = kube_informer_factory.core().v1().pods()
        service_informer = kube_informer_factory.core().v1().services()
        node_informer = kube_informer_factory.core().v1().nodes()

        klog.V(2).info("Creating event broadcaster")
        self.recorder = recorder or EventRecorder(kube_client, C.CONTROLLER_NAME)
        self.pod_control = pod_control or RealPodControl(kube_client, self.recorder)
        self.service_control = service_control or RealServiceControl(kube_client, self.recorder)
        self.expectations = core.Expectations(300.0)
        # per-item back-off as upstream (5 ms * 2^n <= 1000 s); overall bucket from the options (see cmd/options.py)
        self.work_queue = core.WorkQueue(C.KIND, 0.005, 1000.0, float(getattr(option, "queue_qps", 10.0)),
                                         int(getattr(option, "queue_burst", 100)))

        job_informer.informer().add_event_handler(
            add=self.add_training_job, update=self.update_training_job, delete=self.delete_training_job,
            filter_func=lambda o: o.get("kind", C.KIND) == C.KIND)
        self.trainingjob_lister = job_informer.lister()
        self.trainingjob_informer_synced = job_informer.informer().has_synced

        # client-go style secondary indices: a reconcile looks at the pods / services that carry its job label or that it
        # controls -- not at every object of the namespace
        for inf in (pod_informer.informer(), service_informer.informer()):
            inf.indexer.add_indexers({INDEX_JOB_LABEL: index_by_job_label, INDEX_CONTROLLER_UID: index_by_controller_uid})
        pod_informer.informer().add_event_handler(add=self.add_pod, update=self.update_pod, delete=self.delete_pod)
        self.pod_lister = pod_informer.lister()
        self.pod_informer_synced = pod_informer.informer().has_synced

        service_informer.informer().add_event_handler(add=self.add_service, update=self.update_service,
                                                      delete=self.delete_service)
        self.service_lister = service_informer.lister()
        self.service_informer_synced = service_informer.informer().has_synced

        # a node going NotReady must wake the jobs that have replicas on it (the reference polls the node
        # list on every pass instead, pod.go:441)
        node_informer.informer().add_event_handler(update=self._node_changed, delete=self._node_changed_del)
        self.node_lister = node_informer.lister()
        self.node_informer_synced = node_informer.informer().has_synced

        self._workers: List[threading.Thread] = []
        self.gc: Optional[GarbageCollector] = None
        self.sync_count = 0
        # resourceVersion of our own last write per job: a cached copy older than that must not be acted on
        self._written_rv: Dict[str, int] = {}
        self._written_lock = threading.Lock()

    # ------------------------------------------------------------------ identity helpers
    def gen_owner_reference(self, job: AITrainingJob) -> dict:
        """controller.go:161-173."""
        return {"apiVersion": C.API_VERSION, "kind": C.KIND, "name": job.name, "uid": job.uid,
                "blockOwnerDeletion": True, "controller": True}

    def gen_labels(self, job_name: str) -> Dict[str, str]:
        """controller.go:175-180."""
        return {C.LABEL_GROUP_NAME: C.GROUP_NAME, C.LABEL_JOB_NAME: job_name.replace("/", "-")}

    # ------------------------------------------------------------------ run
    def run(self, workers: int, stop: threading.Event) -> None:
        klog.info("Starting training-job controller")
        try:
            klog.info("Starting to create TrainingJob CRD")
            self.create_crd()
            klog.info("Waiting for informer caches to sync")
            if not wait_for_cache_sync(stop, self.trainingjob_informer_synced, self.pod_informer_synced,
                                       self.service_informer_synced, self.node_informer_synced):
                raise RuntimeError("failed to wait for caches for sync")
            klog.info("Starting workers")
            lifecycle.register_stop(lambda: (stop.set(), self.work_queue.shutdown()))
            for i in range(max(1, workers)):
                self._workers.append(lifecycle.spawn(self._worker_loop, f"aitj-worker-{i}", (stop,)))
            self.gc = GarbageCollector(self.kube_client, self.trainingjob_lister)
            threading.Thread(target=self.gc.clean_orphans, args=(self.option.gc_interval, stop), name="aitj-gc",
                             daemon=True).start()
            stop.wait()
        finally:
            self.work_queue.shutdown()
            flush = getattr(self.recorder, "flush", None)
            if flush is not None:
                flush(2.0)                    # events are written by a sink thread: do not drop the tail on a clean stop
            klog.info("Shutting down training-job controller")

    def create_crd(self) -> None:
        try:
            self.api_extensions_client.apiextensions_v1beta1().customresourcedefinitions().create(R.crd_object())
        except APIError as e:
            if e.reason != "AlreadyExists":
                klog.error("Failed to create crd, error: %s", e.message)
                raise

    def _worker_loop(self, stop: threading.Event) -> None:
        # wait.Until(tc.worker, time.Second, stopCh): restart the worker 1 s after it returns
        while not stop.is_set():
            self.worker(stop)
            if self.work_queue.shutting_down():
                return
            stop.wait(1.0)

    def worker(self, stop: Optional[threading.Event] = None) -> None:
        while self.process_next_work_item():
            if stop is not None and stop.is_set():
                return

    def process_next_work_item(self, timeout: float = -1.0) -> bool:
        key = self.work_queue.get(timeout)
        if key is None:
            return not self.work_queue.shutting_down() and timeout >= 0
        metrics.set_gauge("aitj_workqueue_depth", len(self.work_queue))
        t0 = time.perf_counter()
        try:
            forget = self.sync_handler(key)
            if forget:
                self.work_queue.forget(key)
        except APIError as e:
            if e.reason == "NotFound":
                # the job (or the owner of something we tried to create) vanished mid-pass: nothing left to do
                klog.V(2).info("Sync %r: %s", key, e.message)
                self.work_queue.forget(key)
            else:
                klog.error("Sync %r failed with %r", key, e)
                metrics.inc("aitj_reconcile_errors_total")
                self.work_queue.add_rate_limited(key)
        except Exception as e:  # noqa: BLE001 - utilruntime.HandleError + AddRateLimited
            klog.error("Sync %r failed with %r", key, e)
            metrics.inc("aitj_reconcile_errors_total")
            self.work_queue.add_rate_limited(key)
        finally:
            self.work_queue.done(key)
            metrics.observe("aitj_reconcile_seconds", time.perf_counter() - t0)
        return True

    # ------------------------------------------------------------------ sync
    def sync_handler(self, key: str) -> bool:
        t0 = time.perf_counter()
        namespace, name = M.split_key(key)
        if not namespace or not name:
            raise ValueError(f"invalid trainingjob key {key!r}")
        try:
            job = self.trainingjob_lister.aitrainingjobs(namespace).get(name)
        except APIError as e:
            if e.reason == "NotFound":
                klog.V(4).info("%s %s has been deleted", self.kind, key)
                self._forget_job(key)
                return True
            raise
        if self._cache_is_stale(key, job):
            # our own status write has not reached the informer cache yet: acting on the old phase could e.g.
            # re-create pods of a job we just terminated (pods vanish faster here than under a kubelet)
            self.work_queue.add_after(key, 0.005)
            return True
        need_sync = self.satisfied_expectations(job)
        set_defaults_aitrainingjob(job)  # on our private copy (lister returns copies)
        self.sync_count += 1
        if need_sync and job.deletion_timestamp is None and job.status.phase in C.RECONCILABLE_PHASES:
            self.reconcile_training_jobs(job)
        klog.V(4).info("Finished syncing %s %r (%.3f ms)", self.kind, key, (time.perf_counter() - t0) * 1e3)
        return True

    def _forget_job(self, key: str) -> None:
        with self._written_lock:
            self._written_rv.pop(key, None)

    def _cache_is_stale(self, key: str, job: AITrainingJob) -> bool:
        with self._written_lock:
            rec = self._written_rv.get(key)
        if rec is None:
            return False
        uid, rv = rec
        if uid != job.uid:
            return False
        try:
            return int(job.resource_version or 0) < rv
        except ValueError:
            return False

    def _remember_write(self, key: str, updated: AITrainingJob) -> None:
        try:
            rv = int(updated.resource_version or 0)
        except ValueError:
            return
        with self._written_lock:
            self._written_rv[key] = (updated.uid, rv)

    def reconcile_training_jobs(self, job: AITrainingJob) -> None:
        klog.V(4).info("Reconcile training job: %s/%s", job.namespace, job.name)
        old_status = job.status.to_dict()
        old_annotations = dict(job.annotations)
        old_spec = job.spec.to_dict()
        selector = {C.LABEL_GROUP_NAME: self.group, C.LABEL_JOB_NAME: job.name}
        pods = self.get_pods_by_job_and_selector(job, selector)
        services = self.get_services_by_job_and_selector(job, selector)

        self.trace(job, "firstReconcile")
        if self.reconcile_elastic(job, pods):
            return  # spec was patched (auto-scale); the update event re-queues the job

        ending_phases: Dict[str, str] = {}
        aggregation: List[str] = []
        if not job.status.restart_replica_name:
            self.reconcile_rendezvous(job, pods)
            for rtype in list(job.spec.replica_specs):
                ending_phase, msg = self.reconcile_pods(job, pods, rtype)
                if msg and msg not in aggregation:
                    aggregation.append(msg)
                if ending_phase == C.PHASE_RESTARTING:
                    update_conditions(job, C.PHASE_TERMINATING, C.TRAINING_JOB_REASON[C.PHASE_TERMINATING], msg)
                    job.status.restart_replica_name = rtype
                    self.bump_rendezvous(job, "restart")
                    break
                if ending_phase:
                    ending_phases[rtype] = ending_phase
                    continue
                self.reconcile_services(job, services, rtype)
        message = "; ".join(aggregation)
        prev_phase = job.status.phase
        self.update_status(job, pods, services, ending_phases, message)
        if job.status.phase == C.PHASE_RUNNING and prev_phase != C.PHASE_RUNNING:
            self.trace(job, "running")
            self._observe_startup(job)
        if job.status.phase in C.ENDING_PHASES:
            self.trace(job, "ended")

        changed = job.status.to_dict() != old_status or dict(job.annotations) != old_annotations \
            or job.spec.to_dict() != old_spec
        if changed:
            job.status.last_reconcile_time = M.format_time()
            updated = self.update_training_job_phase(job)
            if updated is not None:
                self._remember_write(job.key(), updated)

    def satisfied_expectations(self, job: AITrainingJob) -> bool:
        """Expectations gate (controller.go:390-404).  The reference ORs over every role's pod and service
        keys, so one satisfied key lets a pass run while another create is still in flight (it then trips over
        AlreadyExists); here every key must be satisfied before the next pass, which is what the expectations
        cache is for."""
        key = job.key()
        for rtype in job.spec.replica_specs:
            if not self.expectations.satisfied(gen_expectation_pods_key(key, rtype)):
                return False
            if not self.expectations.satisfied(gen_expectation_services_key(key, rtype)):
                return False
        return True

    def enqueue_job(self, job, is_limited: bool, delay: float) -> None:
        """controller.go:406-421."""
        key = deletion_handling_key(job)
        klog.V(4).info("Enqueue key: %s", key)
        if is_limited:
            self.work_queue.add_rate_limited(key)
        elif delay and delay > 0:
            self.work_queue.add_after(key, float(delay))
        else:
            self.work_queue.add(key)

    def resolve_controller_ref(self, namespace: str, ref: dict) -> Optional[AITrainingJob]:
        """controller.go:424-440: kind must match, uid must match."""
        if ref.get("kind") != self.kind:
            return None
        try:
            job = self.trainingjob_lister.aitrainingjobs(namespace).get(ref.get("name", ""))
        except APIError:
            return None
        if job.uid != ref.get("uid"):
            return None
        return job

    # ------------------------------------------------------------------ node events
    def _node_changed(self, old: dict, cur: dict) -> None:
        def ready(n):
            return any(c.get("type") == "Ready" and c.get("status") == "True"
                       for c in n.get("status", {}).get("conditions") or [])

        if ready(old) == ready(cur):
            return
        self._enqueue_jobs_on_node(M.name_of(cur))
        if ready(cur):
            self._enqueue_autoscaled_jobs()      # a slot came back: edlPolicy Auto roles may grow into it

    def _node_changed_del(self, obj) -> None:
        node = obj.obj if isinstance(obj, DeletedFinalStateUnknown) else obj
        self._enqueue_jobs_on_node(M.name_of(node))

    def _enqueue_jobs_on_node(self, node_name: str) -> None:
        for pod in self.pod_lister.list():
            if pod.get("spec", {}).get("nodeName") != node_name:
                continue
            ref = M.get_controller_of(pod)
            if ref is None:
                continue
            job = self.resolve_controller_ref(M.namespace_of(pod), ref)
            if job is not None:
                self.enqueue_job(job, False, 0)

    def _enqueue_autoscaled_jobs(self) -> None:
        for job in self.trainingjob_lister.list():
            if job.status.phase in C.ENDING_PHASES:
                continue
            if any(r.edl_policy == C.EDL_POLICY_AUTO for r in job.spec.replica_specs.values()):
                self.enqueue_job(job, False, 0)

    # ------------------------------------------------------------------ lifecycle trace
    def trace(self, job: AITrainingJob, event: str) -> None:
        """Sub-second lifecycle timestamps kept in an annotation (metav1.Time has 1 s resolution)."""
        raw = job.annotations.get(C.ANN_TRACE)
        try:
            tr = json.loads(raw) if raw else {}
        except ValueError:
            tr = {}
        if event in tr and event != "ended":
            return
        tr[event] = round(time.time(), 4)
        job.set_annotation(C.ANN_TRACE, json.dumps(tr, sort_keys=True))

    def _observe_startup(self, job: AITrainingJob) -> None:
        try:
            tr = json.loads(job.annotations.get(C.ANN_TRACE, "{}"))
            t0 = tr.get("submitted") or tr.get("firstReconcile")
            if t0 and "running" in tr:
                metrics.observe("aitj_job_startup_seconds", tr["running"] - t0)
        except ValueError:
            pass


def new_training_job_controller(kube_client, trainingjob_client, ext_api_client, kube_informer_factory,
                                trainingjob_informer_factory, option, **kw) -> TrainingJobController:
    """``NewTrainingJobController`` (controller.go:73-159)."""
    return TrainingJobController(kube_client, trainingjob_client, ext_api_client, kube_informer_factory,
                                 trainingjob_informer_factory, option, **kw)
