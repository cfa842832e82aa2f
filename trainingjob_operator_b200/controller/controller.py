"""The controller shell: informers -> rate-limited work queue -> observe -> ``engine.reconcile`` -> ``Executor``.

The capability set is the reference controller's (/root/reference/pkg/controller/controller.go:37-462, SURVEY.md
§2.8, §3.1-3.2): three informers (jobs, pods, services; here also nodes, instead of a live node LIST per pass) feed a
de-duplicating rate-limited queue of ``namespace/name`` keys, N workers pop keys, terminal / deleting jobs are skipped,
the ``AITrainingJob`` kind registers itself on start-up (``AlreadyExists`` tolerated), errors re-queue with per-item
back-off, orphans are swept on a timer.  The shape is not the reference's: a pass is

    observe(key)  -- read-only snapshot of the job and what it owns (claiming adopts / releases by selector + owner uid)
    engine.reconcile(observation) -> decision          -- pure (``controller.engine``)
    executor.apply(decision)                           -- all writes, expectations bookkeeping (``controller.executor``)

plus the two gates that protect a pass from its own past: the expectations gate (creations / deletions of the previous
pass must have been seen by the informers) and the staleness gate (a cached job older than this controller's own last
status write is not acted on -- pods vanish within milliseconds on one box, so a stale ``Running`` phase could re-create
the replicas of a job that just finished).
"""
from __future__ import annotations

import threading
import datetime as _dt
import time
from typing import Dict, List, Optional, Tuple

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
from . import elastic as E
from . import engine
from .control import ControllerRefManager, RealPodControl, RealServiceControl, recheck_deletion_timestamp
from .executor import Executor
from .garbage_collection import GarbageCollector
from .pod import gen_expectation_pods_key, job_labels, owner_reference_of
from .service import gen_expectation_services_key
from .trainingjob import TrainingJobHandlers

metrics.describe("aitj_reconcile_seconds", "duration of one pass over a job key")
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
    """Private copies of the objects a claim pass of ``job`` can act on: the ones matching the selector (keep / adopt)
    -- found through the job-label index -- and the ones the job controls (release when their labels stopped
    matching) -- found through the controller-uid index.  The reference hands the whole namespace to the ref manager."""
    seen = {}
    for o in lister.by_index(INDEX_JOB_LABEL, f"{job.namespace}/{selector.get(C.LABEL_JOB_NAME, '')}"):
        if M.selector_matches(selector, M.labels_of(o)):
            seen[M.key_of(o)] = o
    for o in lister.by_index(INDEX_CONTROLLER_UID, job.uid):
        if M.namespace_of(o) == job.namespace:
            seen.setdefault(M.key_of(o), o)
    return [lister.copy_of(o) for o in seen.values()]


def _node_is_ready(node: dict) -> bool:
    return any(c.get("type") == "Ready" and c.get("status") == "True"
               for c in node.get("status", {}).get("conditions") or [])


class TrainingJobController(TrainingJobHandlers):
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

        self.recorder = recorder or EventRecorder(kube_client, C.CONTROLLER_NAME)
        self.pod_control = pod_control or RealPodControl(kube_client, self.recorder)
        self.service_control = service_control or RealServiceControl(kube_client, self.recorder)
        self.expectations = core.Expectations(300.0)
        # per-item back-off 5 ms * 2^n <= 1000 s as upstream; the overall bucket comes from the options
        self.work_queue = core.WorkQueue(C.KIND, 0.005, 1000.0, float(getattr(option, "queue_qps", 10.0)),
                                         int(getattr(option, "queue_burst", 100)))
        self.executor = Executor(kube_client, trainingjob_client, self.pod_control, self.service_control,
                                 self.expectations, self.work_queue)

        jobs = trainingjob_informer_factory.elasticdeeplearning().v1().aitrainingjobs()
        pods = kube_informer_factory.core().v1().pods()
        services = kube_informer_factory.core().v1().services()
        nodes = kube_informer_factory.core().v1().nodes()
        for inf in (pods.informer(), services.informer()):
            inf.indexer.add_indexers({INDEX_JOB_LABEL: index_by_job_label, INDEX_CONTROLLER_UID: index_by_controller_uid})
        jobs.informer().add_event_handler(add=self.add_training_job, update=self.update_training_job,
                                          delete=self.delete_training_job,
                                          filter_func=lambda o: o.get("kind", C.KIND) == C.KIND)
        pods.informer().add_event_handler(add=self.add_pod, update=self.update_pod, delete=self.delete_pod)
        services.informer().add_event_handler(add=self.add_service)
        # a GPU dropping out must wake the jobs with replicas on it (the reference polls the node list every pass)
        nodes.informer().add_event_handler(update=self._node_changed, delete=self._node_removed)
        self.trainingjob_lister, self.pod_lister = jobs.lister(), pods.lister()
        self.service_lister, self.node_lister = services.lister(), nodes.lister()
        self._synced = (jobs.informer().has_synced, pods.informer().has_synced, services.informer().has_synced,
                        nodes.informer().has_synced)

        self._workers: List[threading.Thread] = []
        self.gc: Optional[GarbageCollector] = None
        self.sync_count = 0
        self._written_rv: Dict[str, Tuple[str, int]] = {}   # job key -> (uid, resourceVersion of our own last write)
        self._written_lock = threading.Lock()
        # AITJ_RECORD_DIR: every pass is written down as (observation, decision) JSON and can be replayed offline
        import os

        self._recorder = None
        if os.environ.get("AITJ_RECORD_DIR"):
            from .replay import Recorder

            self._recorder = Recorder(os.environ["AITJ_RECORD_DIR"])

    # ------------------------------------------------------------------ identity (used by tools / tests)
    gen_owner_reference = staticmethod(owner_reference_of)
    gen_labels = staticmethod(job_labels)

    # ------------------------------------------------------------------ run
    def run(self, workers: int, stop: threading.Event) -> None:
        klog.info("Starting training-job controller")
        try:
            self.create_crd()
            klog.info("Waiting for informer caches to sync")
            if not wait_for_cache_sync(stop, *self._synced):
                raise RuntimeError("failed to wait for caches for sync")
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
                flush(2.0)        # events are written by a sink thread: do not drop the tail on a clean stop
            klog.info("Shutting down training-job controller")

    def create_crd(self) -> None:
        """The kind registers itself (controller.go:210-234); somebody else having done so is fine."""
        try:
            self.api_extensions_client.apiextensions_v1beta1().customresourcedefinitions().create(R.crd_object())
        except APIError as e:
            if e.reason != "AlreadyExists":
                klog.error("Failed to create crd, error: %s", e.message)
                raise

    def _worker_loop(self, stop: threading.Event) -> None:
        while not stop.is_set():       # a worker that returns is restarted after a second, as wait.Until does
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
            if self.sync_handler(key):
                self.work_queue.forget(key)
        except Exception as e:  # noqa: BLE001
            if isinstance(e, APIError) and e.reason == "NotFound":
                # the job (or the owner of something a pass tried to create) vanished mid-pass: nothing left to do
                klog.V(2).info("Sync %r: %s", key, e.message)
                self.work_queue.forget(key)
            else:
                klog.error("Sync %r failed with %r", key, e)
                metrics.inc("aitj_reconcile_errors_total")
                self.work_queue.add_rate_limited(key)
        finally:
            self.work_queue.done(key)
            metrics.observe("aitj_reconcile_seconds", time.perf_counter() - t0)
        return True

    # ------------------------------------------------------------------ one pass
    def sync_handler(self, key: str) -> bool:
        namespace, name = M.split_key(key)
        if not namespace or not name:
            raise ValueError(f"invalid trainingjob key {key!r}")
        try:
            job = self.trainingjob_lister.aitrainingjobs(namespace).get(name)      # a private copy
        except APIError as e:
            if e.reason != "NotFound":
                raise
            with self._written_lock:
                self._written_rv.pop(key, None)
            return True
        if self._cache_is_stale(key, job):
            self.work_queue.add_after(key, 0.005)
            return True
        gate_open = self.satisfied_expectations(job)
        set_defaults_aitrainingjob(job)
        self.sync_count += 1
        if gate_open and job.deletion_timestamp is None and job.status.phase in C.RECONCILABLE_PHASES:
            self.reconcile_training_jobs(job)
        return True

    def reconcile_training_jobs(self, job: AITrainingJob) -> None:
        spare: Tuple[int, ...] = ()
        for _ in range(3):
            obs = self.observe(job, spare)
            before = self._recorder.snapshot(obs) if self._recorder is not None else None
            decision = engine.reconcile(obs)
            if before is not None:
                self._recorder.write(before, obs, decision)
            if not decision.ports_wanted:
                break
            # the pass needs free loopback ports (a new MASTER_PORT, per-replica host ports): allocate, decide again
            spare = tuple(E.allocate_ports(decision.ports_wanted))
            job = self.trainingjob_lister.aitrainingjobs(job.namespace).get(job.name)
            set_defaults_aitrainingjob(job)
        else:
            raise RuntimeError(f"job {job.key()}: the engine keeps asking for ports")
        written = self.executor.apply(obs.job, decision)
        if written is not None:
            self._remember_write(obs.job.key(), written)

    def observe(self, job: AITrainingJob, spare_ports: Tuple[int, ...] = ()) -> engine.Observation:
        selector = {C.LABEL_GROUP_NAME: self.group, C.LABEL_JOB_NAME: job.name}
        pods = self.claim_pods(job, selector, claim_candidates(self.pod_lister, job, selector))
        services = self.claim_services(job, selector, claim_candidates(self.service_lister, job, selector))
        nodes = self._nodes()
        if getattr(self.option, "live_node_list", False):
            # the reference's behaviour, for the control-plane comparison: one live (throttled) LIST per role per pass
            for _ in job.spec.replica_specs:
                try:
                    nodes = self.kube_client.core_v1().nodes().list().get("items", [])
                except APIError as e:
                    klog.error("cannot list nodes: %s", e.message)
        cluster = E.observe_cluster(job, nodes, self.pod_lister.peek()) if E.auto_roles(job) else None
        t = time.time()       # ONE clock read: `now` and `now_epoch` are the same instant, so a recorded pass replays exactly
        return engine.Observation(job=job, pods=pods, services=services,
                                  ready_nodes=frozenset(M.name_of(n) for n in nodes if _node_is_ready(n)),
                                  now=_dt.datetime.fromtimestamp(t, _dt.timezone.utc), now_epoch=t,
                                  options=engine.EngineOptions.from_operator_option(self.option),
                                  cluster=cluster, spare_ports=spare_ports)

    def _nodes(self) -> List[dict]:
        try:
            return self.node_lister.list() if self.node_lister is not None else \
                self.kube_client.core_v1().nodes().list().get("items", [])
        except APIError as e:
            klog.error("cannot list nodes: %s", e.message)
            return []

    def get_node_status(self) -> Dict[str, bool]:
        return {M.name_of(n): True for n in self._nodes() if _node_is_ready(n)}

    # ------------------------------------------------------------------ claiming (adopt / release)
    def _claim(self, patch_fn, job: AITrainingJob, selector: Dict[str, str], objs: List[dict]) -> List[dict]:
        def live_owner():
            f = self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace).get(job.name)
            if f.uid != job.uid:
                raise RuntimeError(f"original {C.KIND} {job.namespace}/{job.name} is gone: got uid {f.uid}, "
                                   f"wanted {job.uid}")
            return f

        return ControllerRefManager(patch_fn, job, selector, recheck_deletion_timestamp(live_owner)).claim(objs)

    def claim_pods(self, job, selector, pods):
        return self._claim(self.pod_control.patch_pod, job, selector, pods)

    def claim_services(self, job, selector, services):
        return self._claim(self.service_control.patch_service, job, selector, services)

    # ------------------------------------------------------------------ gates
    def satisfied_expectations(self, job: AITrainingJob) -> bool:
        """Every pod / service key of the job must be satisfied (the reference ORs them, controller.go:390-404, lets a
        pass run while another create is still in flight and then trips over AlreadyExists)."""
        key = job.key()
        return all(self.expectations.satisfied(gen_expectation_pods_key(key, rt)) and
                   self.expectations.satisfied(gen_expectation_services_key(key, rt))
                   for rt in job.spec.replica_specs)

    def _cache_is_stale(self, key: str, job: AITrainingJob) -> bool:
        with self._written_lock:
            rec = self._written_rv.get(key)
        if rec is None or rec[0] != job.uid:
            return False
        try:
            return int(job.resource_version or 0) < rec[1]
        except ValueError:
            return False

    def _remember_write(self, key: str, updated: AITrainingJob) -> None:
        try:
            rv = int(updated.resource_version or 0)
        except ValueError:
            return
        with self._written_lock:
            self._written_rv[key] = (updated.uid, rv)

    # ------------------------------------------------------------------ queue feeding
    def enqueue_job(self, job, is_limited: bool, delay: float) -> None:
        key = deletion_handling_key(job)
        if is_limited:
            self.work_queue.add_rate_limited(key)
        elif delay and delay > 0:
            self.work_queue.add_after(key, float(delay))
        else:
            self.work_queue.add(key)

    def resolve_controller_ref(self, namespace: str, ref: dict) -> Optional[AITrainingJob]:
        """The job a controller reference points at -- kind, name and uid must all match."""
        if ref.get("kind") != self.kind:
            return None
        try:
            job = self.trainingjob_lister.aitrainingjobs(namespace).get(ref.get("name", ""))
        except APIError:
            return None
        return job if job.uid == ref.get("uid") else None

    def _owner_of(self, obj: dict) -> Tuple[Optional[AITrainingJob], Optional[str]]:
        ref = M.get_controller_of(obj)
        if ref is None:
            return None, None
        job = self.resolve_controller_ref(M.namespace_of(obj), ref)
        return job, (M.labels_of(obj).get(C.LABEL_REPLICA_NAME) if job is not None else None)

    def add_pod(self, pod: dict) -> None:
        if pod.get("metadata", {}).get("deletionTimestamp"):
            return
        job, rt = self._owner_of(pod)
        if job is None or rt is None:
            return
        self.expectations.creation_observed(gen_expectation_pods_key(job.key(), rt))
        self.work_queue.add(job.key())

    def update_pod(self, old: dict, cur: dict) -> None:
        if M.resource_version(cur) == M.resource_version(old):
            return            # a resync re-delivery, nothing changed
        cur_ref, old_ref = M.get_controller_of(cur), M.get_controller_of(old)
        owners = [(M.namespace_of(old), old_ref)] if (old_ref is not None and old_ref != cur_ref) else []
        if cur_ref is not None:
            owners.append((M.namespace_of(cur), cur_ref))
        for ns, ref in owners:
            job = self.resolve_controller_ref(ns, ref)
            if job is not None:
                self.enqueue_job(job, False, 0)

    def delete_pod(self, obj) -> None:
        pod = obj.obj if isinstance(obj, DeletedFinalStateUnknown) else obj
        job, rt = self._owner_of(pod)
        if job is None or rt is None:
            return
        self.expectations.deletion_observed(gen_expectation_pods_key(job.key(), rt))
        self.work_queue.add(job.key())

    def add_service(self, svc: dict) -> None:
        if svc.get("metadata", {}).get("deletionTimestamp"):
            return
        job, rt = self._owner_of(svc)
        if job is None or rt is None:
            return
        self.expectations.creation_observed(gen_expectation_services_key(job.key(), rt))
        self.work_queue.add(job.key())

    def _node_changed(self, old: dict, cur: dict) -> None:
        if _node_is_ready(old) == _node_is_ready(cur):
            return
        self._enqueue_jobs_on_node(M.name_of(cur))
        if _node_is_ready(cur):           # a slot came back: edlPolicy Auto roles may grow into it
            for job in self.trainingjob_lister.list():
                if job.status.phase not in C.ENDING_PHASES and E.auto_roles(job):
                    self.enqueue_job(job, False, 0)

    def _node_removed(self, obj) -> None:
        node = obj.obj if isinstance(obj, DeletedFinalStateUnknown) else obj
        self._enqueue_jobs_on_node(M.name_of(node))

    def _enqueue_jobs_on_node(self, node_name: str) -> None:
        for pod in self.pod_lister.list():
            if pod.get("spec", {}).get("nodeName") == node_name:
                job, _ = self._owner_of(pod)
                if job is not None:
                    self.enqueue_job(job, False, 0)

    # ------------------------------------------------------------------ unit-test seams into the pure engine
    def _obs_for(self, job: AITrainingJob, pods: List[dict], services: List[dict], ready=None) -> engine.Observation:
        ready = frozenset(ready) if ready is not None else frozenset(self.get_node_status())
        return engine.Observation(job=job, pods=pods, services=services, ready_nodes=ready, now=M.now(),
                                  now_epoch=time.time(), options=engine.EngineOptions.from_operator_option(self.option))

    def reconcile_containers(self, job, pod, rtype, node_status):
        """(phase, restart?, message) of one replica, as ``pod.classify_replica`` + ``RESTART_MATRIX`` see it."""
        from .pod import classify_replica

        v = classify_replica(job, pod, node_status, engine.EngineOptions.from_operator_option(self.option).window,
                             M.now())
        return v.phase, v.wants_restart(job.spec.replica_specs[rtype].restart_policy), v.message

    def reconcile_pods(self, job, pods, rtype):
        """(ending phase, message) of one role; the role's actions are applied."""
        d = engine.Decision()
        out = engine.plan_role(self._obs_for(job, pods, []), d, rtype)
        self.executor.apply(job, d)
        return out.phase, out.message

    def update_status(self, job, pods, services, role_phases, message):
        d = engine.Decision()
        engine.derive_status(self._obs_for(job, pods, services), d, role_phases, message)
        self.executor.apply(job, d)


def new_training_job_controller(kube_client, trainingjob_client, ext_api_client, kube_informer_factory,
                                trainingjob_informer_factory, option, **kw) -> TrainingJobController:
    return TrainingJobController(kube_client, trainingjob_client, ext_api_client, kube_informer_factory,
                                 trainingjob_informer_factory, option, **kw)
