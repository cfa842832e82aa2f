"""Liveness half of the node agent's kubelet role: ``containers[].livenessProbe.exec`` (period / failureThreshold /
initialDelay / timeout as in Kubernetes) and the built-in heartbeat contract -- a worker that stops touching
``$AITJ_HEARTBEAT_FILE`` for ``AITJ_HANG_TIMEOUT`` seconds is killed (exit 137) and handed to the job's restart policy.
This is the hang detection the reference leaves to kubelet probes.  Mixed into ``NodeAgent``.
"""
from __future__ import annotations

import os
import signal
import subprocess
import threading
import time

from ..api import constants as C
from ..api import meta as M
from ..utils import klog, metrics

metrics.describe("aitj_liveness_kills_total", "containers killed by a failed liveness probe / missing heartbeat")


class LivenessMixin:
    def _probe_loop(self, stop: threading.Event) -> None:
        while not stop.wait(0.25):
            try:
                self.check_liveness_once()
            except Exception as e:  # noqa: BLE001
                klog.V(2).info("agent: liveness pass failed: %r", e)

    def check_liveness_once(self) -> None:
        now = time.monotonic()
        with self._lock:
            states = [st for st in self._states.values() if st.started and st.containers]
        for st in states:
            pod = self.pod_lister.peek_key(st.key)          # read-only: no copy per pod per pass
            if pod is None:
                continue
            if M.uid_of(pod) != st.uid or pod.get("metadata", {}).get("deletionTimestamp") or \
                    pod.get("status", {}).get("phase") != C.POD_RUNNING:
                continue
            for c in pod.get("spec", {}).get("containers") or []:
                sid = st.containers.get(c["name"])
                if not sid or not self.sup.alive(sid):
                    continue
                verdict = self._probe_container(pod, st, c, now)
                if verdict:
                    self.recorder.event(pod, "Warning", "Unhealthy", f"Liveness probe failed: {verdict}")
                    self.recorder.event(pod, "Normal", "Killing",
                                        f"Container {c['name']} failed liveness probe, will be restarted")
                    klog.warning("pod %s container %s: liveness failed (%s), killing", st.key, c["name"], verdict)
                    metrics.inc("aitj_liveness_kills_total")
                    st.probe_failures.pop(c["name"], None)
                    self.sup.kill(sid, signal.SIGKILL, True)

    def _probe_container(self, pod: dict, st: _PodState, c: dict, now: float) -> str:
        """'' while the container is considered alive, else the reason to kill it."""
        cname = c["name"]
        env = {str(e.get("name")): str(e.get("value", "")) for e in c.get("env") or [] if "name" in e}
        # built-in hang detection: the worker touches its heartbeat file every step
        try:
            hang = float(env.get("AITJ_HANG_TIMEOUT", "0") or 0)
        except ValueError:
            hang = 0.0
        if hang > 0:
            try:
                age = time.time() - os.stat(self.heartbeat_path(pod, cname)).st_mtime
            except OSError:
                age = now - st.started_at            # never written: count from container start
            if age > hang:
                return f"no heartbeat for {age:.1f}s (AITJ_HANG_TIMEOUT={hang:g}s)"
        probe = (c.get("livenessProbe") or {})
        cmd = (probe.get("exec") or {}).get("command")
        if not cmd:
            return ""
        period = float(probe.get("periodSeconds", 10))
        if now - st.started_at < float(probe.get("initialDelaySeconds", 0)) or now < st.probe_next.get(cname, 0.0):
            return ""
        st.probe_next[cname] = now + period
        import subprocess

        gpus = [int(g) for g in (M.annotations_of(pod).get(C.ANN_GPUS) or "").split(",") if g.strip()]
        try:
            r = subprocess.run([str(x) for x in cmd], env=self._container_env(pod, c, gpus), capture_output=True,
                               timeout=float(probe.get("timeoutSeconds", 1)), cwd=c.get("workingDir") or None)
            ok, why = r.returncode == 0, f"exit code {r.returncode}"
        except subprocess.TimeoutExpired:
            ok, why = False, "timed out"
        except OSError as e:
            ok, why = False, str(e)
        if ok:
            st.probe_failures[cname] = 0
            return ""
        st.probe_failures[cname] = st.probe_failures.get(cname, 0) + 1
        if st.probe_failures[cname] >= int(probe.get("failureThreshold", 3)):
            return f"exec {cmd!r}: {why} ({st.probe_failures[cname]} consecutive failures)"
        return ""
