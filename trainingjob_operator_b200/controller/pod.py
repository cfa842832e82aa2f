"""Replica (pod) reconciler: the heart of the controller.

Behavioural parity with /root/reference/pkg/controller/pod.go:23-696:

* event handlers with expectations bookkeeping (pod.go:23-123);
* claim/adopt by selector + owner uid (pod.go:125-150);
* per-role reconcile (pod.go:152-326): short-circuits for Terminating / ``Preempted`` / ``Failed``
  annotations, index bucketing, create missing indices, classify existing pods, restart policy x
  scope x limit x exit-code list, per-role complete/fail policy (Any / Rank0 / All / None),
  scheduling and creating messages;
* container-state classifier (pod.go:328-437) incl. the creating-failed window flags;
* node readiness (pod.go:439-455), scheduling message (pod.go:457-467), grace-0 force delete
  (pod.go:469-481);
* pod construction (pod.go:483-546): labels, ``<job>-<role>-<index>`` names, schedulerName, priority
  label, restartPolicy forced to Never, ownerRef; env contract (pod.go:548-652).

Where this differs on purpose: Q1 out-of-range indices are drained then deleted on
scale-down instead of being ignored-but-counted; Q3 expectations accumulate (raise/lower) and
deletions are expected too; Q4 role keys are lower-cased consistently in names; Q10
``pod.status.startTime`` may be absent without a crash; missing pods of one pass are created in
parallel rather than sequentially.  New behind the reference's dead fields (Q2): replicas are created
for the current *rendezvous generation* and get torch / NCCL rendezvous variables.
"""
from __future__ import annotations

import concurrent.futures as cf
import time
from typing import Dict, List, Optional, Tuple

from ..api import constants as C
from ..api import meta as M
from ..api.types import AITrainingJob, ReplicaSpec
from ..api.validation import parse_exit_codes
from ..client.informers import DeletedFinalStateUnknown
from ..store.apiserver import APIError
from ..utils import klog, metrics
from . import status as S
from .control import ControllerRefManager, recheck_deletion_timestamp

ROLE_PORT_STRIDE = 64  # host-port remap on the single box: P + role_idx * 64 + replica_index


def gen_expectation_pods_key(job_key: str, rt: str) -> str:
    """``job_controller.GenExpectationPodsKey``: ``<ns>/<job>/<rt lower>/pods``."""
    return f"{job_key}/{rt.lower()}/pods"


def gen_general_name(job_name: str, rtype: str, index: str) -> str:
    """trainingjob.go:12-15."""
    return f"{job_name}-{rtype}-{index}".replace("/", "-")


def is_retryable_exit_code(exit_codes: List[int], restarting_exit_code: str) -> bool:
    """controller.go:442-462: every collected code must be listed; an empty list is not retryable."""
    if not exit_codes:
        return False
    try:
        allowed = set(parse_exit_codes(restarting_exit_code))
    except ValueError:
        allowed = set()
    return all(c in allowed for c in exit_codes)


def get_ports_from_container(container: dict) -> List[int]:
    """service.go:33-43: ``aitj-`` container, ``aitj-`` port names."""
    if not str(container.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX):
        return []
    return [p.get("containerPort") for p in container.get("ports") or []
            if str(p.get("name", "")).startswith(C.DEFAULT_PORT_PREFIX)]


def get_ports_from_job(job: AITrainingJob, rtype: str) -> List[int]:
    """service.go:19-31."""
    ports: List[int] = []
    for c in job.spec.replica_specs[rtype].containers():
        ports += get_ports_from_container(c)
    return ports


def host_port(job: AITrainingJob, rtype: str, index: int, port: int) -> int:
    """Unique loopback port for (role, index, declared port) on the single box."""
    roles = sorted(job.spec.replica_specs)
    return int(port) + roles.index(rtype) * ROLE_PORT_STRIDE + index


class PodReconciler:
    """Mixed into ``TrainingJobController``."""

    # ---------------------------------------------------------------------------- handlers
    def add_pod(self, pod: dict) -> None:
        if pod.get("metadata", {}).get("deletionTimestamp"):
            return
        ref = M.get_controller_of(pod)
        if ref is None:
            return
        job = self.resolve_controller_ref(M.namespace_of(pod), ref)
        if job is None:
            return
        rt = M.labels_of(pod).get(C.LABEL_REPLICA_NAME)
        if rt is None:
            klog.info("This pod may not created by %s", C.CONTROLLER_NAME)
            return
        klog.V(4).info("Pod %s created", M.name_of(pod))
        self.expectations.creation_observed(gen_expectation_pods_key(job.key(), rt))
        self.work_queue.add(job.key())

    def update_pod(self, old: dict, cur: dict) -> None:
        if M.resource_version(cur) == M.resource_version(old):
            return
        cur_ref, old_ref = M.get_controller_of(cur), M.get_controller_of(old)
        if cur_ref != old_ref and old_ref is not None:
            job = self.resolve_controller_ref(M.namespace_of(old), old_ref)
            if job is not None:
                self.enqueue_job(job, False, 0)
        if cur_ref is not None:
            job = self.resolve_controller_ref(M.namespace_of(cur), cur_ref)
            if job is None:
                return
            klog.V(4).info("Pod %s updated", M.name_of(cur))
            self.enqueue_job(job, False, 0)

    def delete_pod(self, obj) -> None:
        pod = obj.obj if isinstance(obj, DeletedFinalStateUnknown) else obj
        ref = M.get_controller_of(pod)
        if ref is None:
            return
        job = self.resolve_controller_ref(M.namespace_of(pod), ref)
        if job is None:
            return
        rt = M.labels_of(pod).get(C.LABEL_REPLICA_NAME)
        if rt is None:
            return
        klog.V(4).info("Pod %s/%s deleted", M.namespace_of(pod), M.name_of(pod))
        self.expectations.deletion_observed(gen_expectation_pods_key(job.key(), rt))
        self.work_queue.add(job.key())

    # ---------------------------------------------------------------------------- claim
    def get_pods_by_job_and_selector(self, job: AITrainingJob, selector: Dict[str, str]) -> List[dict]:
        from .controller import claim_candidates

        return self.claim_pods(job, selector, claim_candidates(self.pod_lister, job, selector))

    def claim_pods(self, job: AITrainingJob, selector: Dict[str, str], pods: List[dict]) -> List[dict]:
        def fresh():
            f = self.trainingjob_client.elasticdeeplearning_v1().aitrainingjobs(job.namespace).get(job.name)
            if f.uid != job.uid:
                raise RuntimeError(f"original {C.KIND} {job.namespace}/{job.name} is gone: got uid {f.uid}, "
                                   f"wanted {job.uid}")
            return f

        mgr = ControllerRefManager(self.pod_control.patch_pod, job, selector, recheck_deletion_timestamp(fresh))
        return mgr.claim(pods)

    # ---------------------------------------------------------------------------- reconcile
    def reconcile_pods(self, job: AITrainingJob, pods: List[dict], rtype: str) -> Tuple[str, str]:
        """Returns (ending phase or "" / Restarting / Terminating, message)."""
        if job.status.phase == C.PHASE_TERMINATING:
            return C.PHASE_TERMINATING, ""
        if C.PHASE_PREEMPTED in job.annotations:
            return C.PHASE_PREEMPTED, job.annotations[C.PHASE_PREEMPTED]
        if C.PHASE_FAILED in job.annotations:
            return C.PHASE_FAILED, job.annotations[C.PHASE_FAILED]

        rt = rtype.lower()
        spec = job.spec.replica_specs[rtype]
        replica_pods = S.filter_pods_for_replica_type(pods, rt)
        replicas = int(spec.replicas or 0)
        S.initialize_replica_statuses(job, rtype)
        S.initialize_restart_counts(job, rtype)

        pod_slices, surplus = self.get_pod_slices(replica_pods, replicas)
        self.reconcile_surplus_pods(job, rtype, surplus)
        node_status = self.get_node_status()
        message = ""
        failed_reason: List[str] = []
        failed_phase = C.PHASE_FAILED
        creating_msg: Dict[str, List[str]] = {}
        to_create: List[int] = []
        rs = job.status.replica_statuses[rtype]

        for index, pod_slice in enumerate(pod_slices):
            if not pod_slice:
                klog.info("Need to create new pod: %s/%s %s-%d", job.namespace, job.name, rt, index)
                to_create.append(index)
                continue
            pod = pod_slice[0]
            msg = self.get_pod_scheduling_message(pod)
            if msg:
                klog.V(2).info("pod %s is scheduling:%s", M.name_of(pod), msg)
                message = f"{rt}: {msg} "
            phase, is_restart, msg = self.reconcile_containers(job, pod, rtype, node_status)
            klog.V(2).info("reconcileContainers %s => %r %s %r", M.name_of(pod), phase, is_restart, msg)
            if msg:
                failed_reason.append(msg)

            if is_restart:
                grace = 0 if phase == C.PHASE_NODE_FAIL else None
                limit = spec.restart_limit
                if limit is None or job.status.restart_counts.get(rtype, 0) < limit:
                    S.update_restart_count(job, rtype)
                    msg = f"restart times is {job.status.restart_counts[rtype]}, {msg} "
                    # faultTolerant (unused in the reference, types.go:47): an elastic job replaces only the lost
                    # replica; survivors keep their state and re-rendezvous with the replacement
                    scope = S.effective_restart_scope(job, rtype)
                    if scope == C.RESTART_SCOPE_POD:
                        klog.warning("According to restartscope, need to restart the pod: %s.%s",
                                     M.namespace_of(pod), M.name_of(pod))
                        victims = [pod]
                    elif scope == C.RESTART_SCOPE_REPLICA:
                        klog.warning("According to restartscope, need to restart all pods of the replica: %s", rtype)
                        victims = [p for sl in pod_slices for p in sl]
                    else:
                        klog.warning("According to restartscope, need to restart all pods")
                        victims = list(pods)
                    self.delete_pods_expecting(job, victims, grace)
                    for r in job.spec.replica_specs:
                        S.update_replica_statuses(job, r, S.filter_pods_for_replica_type(pods, r.lower()))
                    metrics.inc("aitj_restarts_total", labels={"scope": spec.restart_scope})
                    return C.PHASE_RESTARTING, msg

            if phase == C.PHASE_CREATING:
                creating_msg.setdefault(msg, []).append(M.name_of(pod))

            pphase = S.pod_phase(pod)
            if phase == C.PHASE_SUCCEEDED and pphase == C.POD_SUCCEEDED and \
                    spec.complete_policy == C.ENDING_POLICY_ANY:
                return phase, f"pod {M.name_of(pod)} have completed"
            if phase in (C.PHASE_FAILED, C.PHASE_NODE_FAIL) and spec.fail_policy == C.ENDING_POLICY_ANY:
                return phase, f"pod {M.name_of(pod)} is failed, {msg}"
            if index == 0:
                if phase == C.PHASE_SUCCEEDED and pphase == C.POD_SUCCEEDED and \
                        spec.complete_policy == C.ENDING_POLICY_RANK0:
                    return C.PHASE_SUCCEEDED, f"rank0 pod {M.name_of(pod)} have completed"
                if phase in (C.PHASE_FAILED, C.PHASE_NODE_FAIL) and spec.fail_policy == C.ENDING_POLICY_RANK0:
                    return phase, f"rank0 pod {M.name_of(pod)} is failed, {msg}"
            if phase == C.PHASE_NODE_FAIL:
                failed_phase = C.PHASE_NODE_FAIL
            S.count_pod(job, rtype, pod, rs)

        if to_create:
            self.create_new_pods(job, rt, to_create, job.status.restart_counts.get(rtype, 0), spec)

        in_range = [p for sl in pod_slices for p in sl[:1]]
        S.update_replica_statuses(job, rtype, in_range)
        rs = job.status.replica_statuses[rtype]
        klog.V(4).info("%s status %s", rtype, rs)

        if spec.complete_policy == C.ENDING_POLICY_ALL and replicas > 0 and rs.succeeded == replicas:
            return C.PHASE_SUCCEEDED, f"All {rtype} pods have completed"
        if spec.fail_policy == C.ENDING_POLICY_ALL and replicas > 0 and rs.failed == replicas:
            if failed_reason:
                message = ", ".join(failed_reason)
            return failed_phase, f"All {rtype} pods are failed, {message}"
        if creating_msg:
            return C.PHASE_NONE, ", ".join(f"pods {names} {m}" for m, names in creating_msg.items())
        return C.PHASE_NONE, message

    # ---------------------------------------------------------------------------- classifier
    def reconcile_containers(self, job: AITrainingJob, pod: dict, rtype: str,
                             node_status: Dict[str, bool]) -> Tuple[str, bool, str]:
        """pod.go:328-437 -> (phase, is_restart, message)."""
        spec = job.spec.replica_specs[rtype]
        exit_codes: List[int] = []
        failed_reason: List[str] = []
        is_restart = False
        is_succeeded = True
        is_creating = False
        pstatus = pod.get("status", {})
        node_name = S.pod_node(pod)
        for cs in pstatus.get("containerStatuses") or []:
            state = cs.get("state") or {}
            term, waiting = state.get("terminated"), state.get("waiting")
            if str(cs.get("name", "")).startswith(C.DEFAULT_CONTAINER_PREFIX):
                is_succeeded = is_succeeded and term is not None
                if term is not None:
                    code = int(term.get("exitCode", 0))
                    is_succeeded = is_succeeded and code == 0
                    exit_codes.append(code)
                    msg = (f"container {cs.get('name')} on node {node_name} exited with reason "
                           f"{term.get('reason', '')} exitcode {code}")
                    klog.V(2).info(msg)
                    if code != 0:
                        failed_reason.append(msg)
            if waiting is not None:
                is_creating = True
                reason = waiting.get("reason", "")
                if reason in C.ERROR_CONTAINER_STATUS:
                    creating = job.status.get_condition(C.PHASE_CREATING)
                    if creating is not None and creating.status == "True":
                        since_creating = M.seconds_since(creating.last_transition_time)
                        if since_creating < self.option.creating_restart_time:
                            started = pstatus.get("startTime")
                            if started and M.seconds_since(started) > self.option.creating_duration_time:
                                klog.warning("pod %s create container failed: %s", M.name_of(pod),
                                             waiting.get("message", ""))
                                is_restart = True
                        elif self.option.enable_creating_failed:
                            msg = (f"pod {M.name_of(pod)} create container failed[{reason}] and has been retrying "
                                   f"for {self.option.creating_restart_time:g} seconds")
                            klog.warning(msg)
                            return C.PHASE_FAILED, is_restart, msg
                    failed_reason.append(reason)

        pphase = S.pod_phase(pod)
        if pphase == C.POD_FAILED:
            rp = spec.restart_policy
            if (rp in (C.RESTART_POLICY_EXIT_CODE, C.RESTART_POLICY_ON_NODE_FAIL_WITH_EXIT_CODE)
                    and is_retryable_exit_code(exit_codes, job.spec.restarting_exit_code)) \
                    or rp in (C.RESTART_POLICY_ON_FAILURE, C.RESTART_POLICY_ALWAYS):
                is_restart = True
            message = ""
            if failed_reason:
                message = "; ".join(failed_reason)
            elif pstatus.get("reason"):
                message = pstatus["reason"]
                if pstatus.get("message"):
                    message = f"{pstatus['reason']}, {pstatus['message']}"
            return C.PHASE_FAILED, is_restart, message

        if node_name and node_name not in node_status:
            if spec.restart_policy in (C.RESTART_POLICY_ON_NODE_FAIL_WITH_EXIT_CODE, C.RESTART_POLICY_ON_NODE_FAIL,
                                       C.RESTART_POLICY_ALWAYS):
                is_restart = True
            return C.PHASE_NODE_FAIL, is_restart, f"Node {node_name} is failed and offline"

        if is_creating:
            if failed_reason:
                return C.PHASE_CREATING, is_restart, "; ".join(failed_reason)
            return C.PHASE_CREATING, is_restart, "creating containers"
        if is_succeeded:
            return C.PHASE_SUCCEEDED, is_restart, ""
        return C.PHASE_NONE, is_restart, ""

    def get_node_status(self) -> Dict[str, bool]:
        """pod.go:439-455: set of Ready nodes (here: healthy GPU slots + the CPU slot)."""
        ready: Dict[str, bool] = {}
        try:
            nodes = self.node_lister.list() if self.node_lister is not None else \
                self.kube_client.core_v1().nodes().list().get("items", [])
        except APIError as e:
            klog.error("getNodeStatus failed %s", e.message)
            return ready
        for node in nodes:
            for cond in node.get("status", {}).get("conditions") or []:
                if cond.get("type") == "Ready" and cond.get("status") == "True":
                    ready[M.name_of(node)] = True
                    break
        return ready

    @staticmethod
    def get_pod_scheduling_message(pod: dict) -> str:
        """pod.go:457-467."""
        if S.pod_phase(pod) == C.POD_PENDING and not S.pod_node(pod):
            for cond in pod.get("status", {}).get("conditions") or []:
                if cond.get("type") == "PodScheduled" and cond.get("status") == "False":
                    return cond.get("message", "")
        return ""

    def delete_pods_expecting(self, job: AITrainingJob, victims: List[dict], grace: Optional[int]) -> None:
        """Delete pods and record the expected deletions (the reference observes deletions without ever
        expecting them).  ``grace=0`` is the reference's ``forceDeletePod`` (pod.go:469-481), used for
        replicas on a failed node; ``None`` is the pod's own termination grace period."""
        per_role: Dict[str, int] = {}
        for p in victims:
            rt = M.labels_of(p).get(C.LABEL_REPLICA_NAME, "")
            per_role[rt] = per_role.get(rt, 0) + 1
        for rt, n in per_role.items():
            self.expectations.raise_expectations(gen_expectation_pods_key(job.key(), rt), 0, n)
        for p in victims:
            try:
                self.pod_control.delete_pod(M.namespace_of(p), M.name_of(p), job, grace_period_seconds=grace)
            except APIError as e:
                self.expectations.deletion_observed(
                    gen_expectation_pods_key(job.key(), M.labels_of(p).get(C.LABEL_REPLICA_NAME, "")))
                klog.error("delete pod %s failed: %s", M.name_of(p), e.message)

    # ---------------------------------------------------------------------------- scale-down (Q1)
    def reconcile_surplus_pods(self, job: AITrainingJob, rtype: str, surplus: List[dict]) -> None:
        """Replicas whose index >= spec.replicas: mark draining (so they are no longer counted), let the
        worker leave at a step boundary, delete once exited or after ``--scale-down-grace``."""
        for pod in surplus:
            ann = M.annotations_of(pod)
            phase = S.pod_phase(pod)
            if C.ANN_SCALE_DOWN not in ann:
                try:
                    self.pod_control.patch_pod(M.namespace_of(pod), M.name_of(pod),
                                               {"metadata": {"annotations": {C.ANN_SCALE_DOWN: M.format_time()}}})
                except APIError as e:
                    klog.warning("cannot mark %s draining: %s", M.name_of(pod), e.message)
                pod.setdefault("metadata", {}).setdefault("annotations", {})[C.ANN_SCALE_DOWN] = M.format_time()
                drained_for = 0.0
            else:
                drained_for = M.seconds_since(ann[C.ANN_SCALE_DOWN])
            exited = phase in (C.POD_SUCCEEDED, C.POD_FAILED)
            if exited or not S.pod_node(pod) or drained_for >= self.option.scale_down_grace:
                if not pod.get("metadata", {}).get("deletionTimestamp"):
                    klog.info("scale-down: deleting replica %s (index out of range)", M.name_of(pod))
                    self.delete_pods_expecting(job, [pod], None)
            else:
                self.enqueue_job(job, False, max(0.2, min(1.0, self.option.scale_down_grace - drained_for)))

    # ---------------------------------------------------------------------------- construction
    def create_new_pods(self, job: AITrainingJob, rt: str, indices: List[int], restart_count: int,
                        spec: ReplicaSpec) -> None:
        """Create all missing replicas of one pass concurrently (the reference loops sequentially with one
        synchronous POST each, pod.go:186-193)."""
        key = gen_expectation_pods_key(job.key(), rt)
        self.expectations.raise_expectations(key, len(indices), 0)
        templates = [(i, self.build_pod_template(job, rt, str(i), str(restart_count), spec)) for i in indices]
        ref = self.gen_owner_reference(job)
        errors: List[Exception] = []

        def one(item):
            i, tpl = item
            t0 = time.perf_counter()
            try:
                self.pod_control.create_pods_with_controller_ref(job.namespace, tpl, job, ref)
            except APIError as e:
                self.expectations.creation_observed(key)
                if e.reason == "AlreadyExists":
                    return
                errors.append(e)
            metrics.observe("aitj_pod_create_seconds", time.perf_counter() - t0)

        # Concurrency only pays when a create is a network round trip (separate API server process); against the
        # in-process store the creates are sub-millisecond and pure CPU, so threads would just queue on the GIL.
        remote = getattr(getattr(self.kube_client, "transport", None), "master", None) is not None
        if len(templates) == 1 or not remote:
            for item in templates:
                one(item)
        else:
            list(self._create_pool().map(one, templates))
        if errors:
            raise errors[0]

    def _create_pool(self) -> cf.ThreadPoolExecutor:
        pool = getattr(self, "_pod_create_pool", None)
        if pool is None:                    # one long-lived pool per controller, not one per reconcile pass
            pool = self._pod_create_pool = cf.ThreadPoolExecutor(max_workers=16, thread_name_prefix="pod-create")
        return pool

    def build_pod_template(self, job: AITrainingJob, rt: str, index: str, restart_count: str,
                           spec: ReplicaSpec) -> dict:
        """pod.go:483-546 (everything except the API call)."""
        labels = self.gen_labels(job.name)
        labels[C.LABEL_JOBNAME_COMPAT] = job.name
        labels[C.LABEL_POD_ROLE] = rt
        labels[C.LABEL_RESTART_COUNT] = restart_count
        labels[C.LABEL_REPLICA_NAME] = rt
        labels[C.LABEL_REPLICA_INDEX] = index
        if job.spec.priority:
            labels[C.LABEL_PRIORITY] = job.spec.priority
        rdv = job.status.rendezvous
        if rdv is not None:
            labels[C.LABEL_GENERATION] = str(rdv.generation)
        tpl = M.deepcopy(spec.template)
        md = tpl.setdefault("metadata", {})
        md["name"] = gen_general_name(job.name, rt, index)
        md["generateName"] = gen_general_name(job.name, rt, "")
        tl = md.setdefault("labels", {})
        tl.update(labels)
        for k, v in job.labels.items():
            tl.setdefault(k, v)
        pspec = tpl.setdefault("spec", {})
        if job.spec.scheduler_name:
            pspec["schedulerName"] = job.spec.scheduler_name
        self.set_env(tpl, job, spec, rt, index, restart_count)
        if spec.restart_policy:
            pspec["restartPolicy"] = "Never"
        if rdv is not None and rdv.master_port:
            md.setdefault("annotations", {})[C.ANN_RENDEZVOUS_PORT] = str(rdv.master_port)
        return tpl

    def set_env(self, tpl: dict, job: AITrainingJob, spec: ReplicaSpec, rtype: str, index: str,
                restart_count: str) -> None:
        """The rendezvous environment contract (pod.go:548-652), appended to every init container and
        container, plus ``TRAININGJOB_PORTS`` per container.  Extra variables follow the 13 reference ones."""
        env: List[dict] = []

        def add(name: str, value) -> None:
            env.append({"name": name, "value": str(value)})

        role_key = next((r for r in job.spec.replica_specs if r.lower() == rtype), rtype)
        for rt, rspec in job.spec.replica_specs.items():
            ports = get_ports_from_job(job, rt)
            n = int(rspec.replicas or 0)
            instances = [f"{gen_general_name(job.name, rt.lower(), str(i))}.{job.namespace}" for i in range(n)]
            hosts = [f"{inst}:{p}" for inst in instances for p in ports]
            addrs = [f"127.0.0.1:{host_port(job, rt, i, p)}" for i in range(n) for p in ports]
            up = rt.upper()
            add(f"{up}_INSTANCES", ",".join(instances))
            add(f"{up}_INSTANCES_NUM", len(instances))
            add(f"{up}_PORTS", ",".join(str(p) for p in ports))
            add(f"{up}_PORTS_NUM", len(ports))
            add(f"{up}_HOSTS", ",".join(hosts))
            add(f"{up}_HOSTS_NUM", len(hosts))
            add(f"{up}_ADDRS", ",".join(addrs))  # new: loopback-resolvable form of <ROLE>_HOSTS
        add(C.ENV_REPLICA_NAME, rtype)
        add(C.ENV_REPLICA_INDEX, index)
        add(C.ENV_REPLICA_RESTARTCOUNT, restart_count)
        add(C.ENV_SERVICE, f"{gen_general_name(job.name, rtype, index)}.{job.namespace}")
        add(C.ENV_JOB_NAME, job.name)
        add(C.ENV_JOB_NAMESPACE, job.namespace)

        # --- new: torch.distributed / elastic rendezvous dialect (frameworkType selects it) ------------
        fw = (job.spec.framework_type or "pytorch").lower()
        rdv = job.status.rendezvous
        world = int(spec.replicas or 0)
        if rdv is not None and role_key in rdv.world_sizes:
            world = rdv.world_sizes[role_key]
        add("AITJ_JOB_UID", job.uid)
        add("AITJ_FRAMEWORK", fw)
        add("AITJ_RENDEZVOUS_GENERATION", rdv.generation if rdv else 0)
        add("AITJ_FAULT_TOLERANT", "1" if job.spec.fault_tolerant else "0")
        add("AITJ_EDL_POLICY", spec.edl_policy or C.EDL_POLICY_NEVER)
        if spec.min_replicas is not None:
            add("AITJ_MIN_REPLICAS", spec.min_replicas)
        if spec.max_replicas is not None:
            add("AITJ_MAX_REPLICAS", spec.max_replicas)
        if getattr(self, "master_url", ""):
            add("AITJ_MASTER", self.master_url)
        if fw in ("pytorch", "torch", ""):
            add("RANK", index)
            add("WORLD_SIZE", world)
            add("LOCAL_RANK", 0)
            add("LOCAL_WORLD_SIZE", 1)
            add("MASTER_ADDR", "127.0.0.1")
            add("MASTER_PORT", rdv.master_port if rdv and rdv.master_port else 29500)
        elif fw in ("paddle", "paddlepaddle"):
            add("PADDLE_TRAINER_ID", index)
            add("PADDLE_TRAINERS_NUM", world)
        elif fw in ("tensorflow", "tf"):
            import json as _json

            cluster = {rt.lower(): [f"127.0.0.1:{host_port(job, rt, i, p)}" for i in range(int(rs.replicas or 0))
                                    for p in get_ports_from_job(job, rt)[:1]]
                       for rt, rs in job.spec.replica_specs.items()}
            add("TF_CONFIG", _json.dumps({"cluster": cluster, "task": {"type": rtype, "index": int(index)}}))

        pspec = tpl.setdefault("spec", {})
        for c in pspec.get("initContainers") or []:
            c["env"] = list(c.get("env") or []) + M.deepcopy(env)
        for c in pspec.get("containers") or []:
            ports = get_ports_from_container(c)
            c["env"] = list(c.get("env") or []) + M.deepcopy(env) + [
                {"name": C.ENV_PORTS, "value": ",".join(str(p) for p in ports)},
                {"name": "AITJ_HOST_PORTS",
                 "value": ",".join(str(host_port(job, role_key, int(index), p)) for p in ports)},
            ]

    # ---------------------------------------------------------------------------- bucketing
    @staticmethod
    def get_pod_slices(pods: List[dict], replicas: int) -> Tuple[List[List[dict]], List[dict]]:
        """pod.go:676-696, plus the out-of-range pods the reference only logs (pod.go:688-689)."""
        slices: List[List[dict]] = [[] for _ in range(replicas)]
        surplus: List[dict] = []
        for pod in pods:
            raw = M.labels_of(pod).get(C.LABEL_REPLICA_INDEX)
            if raw is None:
                klog.warning("The pod do not have the index label.")
                continue
            try:
                idx = int(raw)
            except ValueError:
                klog.warning("Error when parsing index label %r", raw)
                continue
            if idx < 0:
                klog.warning("The label index is not expected: %d", idx)
            elif idx >= replicas:
                surplus.append(pod)
            else:
                slices[idx].append(pod)
        for sl in slices:
            sl.sort(key=lambda p: (p.get("metadata", {}).get("deletionTimestamp") is not None,
                                   p.get("metadata", {}).get("creationTimestamp", "")))
        return slices, surplus
