"""Warm pool of the node agent (``warm_pool=N``): parked interpreters with torch already imported (``runtime/zygote.py``),
on a GPU box pinned to one GPU slot each with a live CUDA context.  A container whose command is ``python -m mod`` /
``python script`` adopts one -- the C++ supervisor re-keys the process, the assignment travels over a FIFO -- instead of
paying interpreter start + ``import torch`` (seconds) on the spawn -> Running path.  The reference
has no counterpart (kubelet always starts a fresh container).  Mixed into ``NodeAgent``.
"""
from __future__ import annotations

import json
import os
import shutil
import signal
import sys
import threading
import time
from typing import Any, Dict, List

from ..api import constants as C
from ..api import meta as M
from ..utils import klog, metrics

metrics.describe("aitj_warm_adoptions_total", "containers started by adopting a pre-warmed interpreter")
# environment of the agent that every container (and every parked interpreter) inherits
PASS_ENV = ("PATH", "HOME", "USER", "LANG", "LC_ALL", "LD_LIBRARY_PATH", "VIRTUAL_ENV", "PYTHONPATH", "TMPDIR",
             "CUDA_HOME", "NCCL_DEBUG", "NCCL_SOCKET_IFNAME", "NCCL_IB_DISABLE", "TORCH_NCCL_ASYNC_ERROR_HANDLING",
             "GRAFT_REPO_ROOT", "HF_HOME", "TORCH_HOME", "XDG_CACHE_HOME")
ZYGOTE_PREFIX = "~zygote/"     # supervisor ids of parked interpreters ('~' cannot start a namespace name)


class WarmPoolMixin:
    def _zygote_env(self) -> Dict[str, str]:
        env = {k: os.environ[k] for k in PASS_ENV if k in os.environ}
        env["PYTHONUNBUFFERED"] = "1"
        return env

    def _ensure_pool(self) -> None:
        """Top the pool up to ``warm_pool`` parked interpreters (no-op when disabled or shutting down)."""
        if self.warm_pool <= 0 or self._stopping or self._zy_failures >= 3:
            return
        os.makedirs(self._zy_dir, exist_ok=True)
        with self._lock:
            while len(self._zygotes) < self.warm_pool:
                # GPU box: one parked interpreter per GPU slot, CUDA context included; CPU-only box: generic ones
                gpu = None
                if self.num_gpus > 0:
                    taken = {z["gpu"] for z in self._zygotes.values()}
                    gpu = next((g for g in range(min(self.num_gpus, self.warm_pool)) if g not in taken), None)
                    if gpu is None:
                        return
                self._zy_seq += 1
                zid = f"{ZYGOTE_PREFIX}{os.getpid()}-{self._zy_seq}"
                fifo = os.path.join(self._zy_dir, f"z{os.getpid()}-{self._zy_seq}.fifo")
                for path in (fifo, fifo + ".ready"):
                    try:
                        os.unlink(path)
                    except OSError:
                        pass
                try:
                    os.mkfifo(fifo, 0o600)
                    env = self._zygote_env()
                    cmd = [sys.executable, "-m", "trainingjob_operator_b200.runtime.zygote", fifo]
                    if gpu is not None:
                        share = self.gpu_visibility == "all" and self.num_gpus > 1
                        env["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in range(self.num_gpus)) if share else str(gpu)
                        env["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
                        env["AITJ_ZYGOTE_DEVICE"] = str(gpu) if share else "0"
                        cmd.append("--cuda")
                    self.sup.spawn(zid, cmd, env, "", os.path.join(self.log_dir, "zygotes.log"), "",
                                   self._cpus_for([gpu]) if gpu is not None else [])
                except OSError as e:
                    klog.warning("warm pool: cannot start an interpreter: %s", e)
                    self._zy_failures += 1
                    return
                self._zygotes[zid] = {"fifo": fifo, "spawned": time.monotonic(), "gpu": gpu}

    def warm_ready(self) -> int:
        """Number of parked interpreters that finished their imports."""
        with self._lock:
            return sum(1 for z in self._zygotes.values() if os.path.exists(z["fifo"] + ".ready"))

    def _on_zygote_exit(self, ev: Dict[str, Any]) -> None:
        with self._lock:
            z = self._zygotes.pop(ev["id"], None)
        if z is None:
            return
        for path in (z["fifo"], z["fifo"] + ".ready"):
            try:
                os.unlink(path)
            except OSError:
                pass
        if self._stopping:
            return
        if time.monotonic() - z["spawned"] < 5.0:
            self._zy_failures += 1
            klog.warning("warm pool: parked interpreter exited early (code %s)", ev.get("exit_code"))
        self._ensure_pool()

    def _kill_zygotes(self) -> None:
        self._stopping = True
        with self._lock:
            ids = list(self._zygotes)
        for zid in ids:
            self.sup.kill(zid, signal.SIGKILL, True)

    def _adopt_zygote(self, sid: str, argv: List[str], env: Dict[str, str], cwd: str, log: str,
                      cpus: List[int]) -> bool:
        """Start the container by handing it to a parked interpreter.  False => caller spawns it cold."""
        if self.warm_pool <= 0 or self._stopping:
            return False
        from ..runtime.zygote import split_python_command

        if split_python_command(argv) is None:
            return False
        exe = shutil.which(argv[0], path=env.get("PATH")) or argv[0]
        try:
            if os.path.realpath(exe) != os.path.realpath(sys.executable):
                return False
        except OSError:
            return False
        want_gpu: Optional[int] = None
        if self.num_gpus > 0:
            vis = env.get("AITJ_PINNED_GPU", env.get("CUDA_VISIBLE_DEVICES", ""))
            if not vis.isdigit():
                return False          # CPU-only or multi-GPU container: parked interpreters are pinned to one slot each
            want_gpu = int(vis)
        with self._lock:
            zid = next((z for z, info in self._zygotes.items()
                        if info["gpu"] == want_gpu and os.path.exists(info["fifo"] + ".ready")), None)
            info = self._zygotes.pop(zid) if zid else None
        if info is None:
            return False
        fd = -1
        deadline = time.monotonic() + 0.25
        while fd < 0:
            try:
                fd = os.open(info["fifo"], os.O_WRONLY | os.O_NONBLOCK)
            except OSError:          # ENXIO: the reader has not reached open() yet
                if time.monotonic() > deadline:
                    break
                time.sleep(0.002)
        ok = fd >= 0 and self.sup.rename(zid, sid)
        if ok:
            msg = json.dumps({"argv": argv, "env": env, "cwd": cwd, "log": log, "cpus": cpus}) + "\n"
            try:
                os.write(fd, msg.encode())
            except OSError:
                ok = False
                self.sup.kill(sid, signal.SIGKILL, True)
        if fd >= 0:
            os.close(fd)
        if not ok:
            self.sup.kill(zid, signal.SIGKILL, True)
            for path in (info["fifo"], info["fifo"] + ".ready"):
                try:
                    os.unlink(path)
                except OSError:
                    pass
        else:
            metrics.inc("aitj_warm_adoptions_total")
            klog.V(2).info("container %s adopted parked interpreter %s", sid, zid)
        threading.Thread(target=self._ensure_pool, daemon=True).start()
        return ok
