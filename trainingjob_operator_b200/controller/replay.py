"""Record / replay of reconcile passes.

``engine.reconcile`` is a pure function, so a pass can be written down as JSON -- the Observation it saw and the
Decision it took -- and replayed anywhere: in the golden tests (``tests/golden/engine_cases.json``), or from a live
operator started with ``AITJ_RECORD_DIR=<dir>`` (one ``<job>-<n>.json`` per pass) to debug a decision after the fact.
The reference has nothing comparable (its reconcile interleaves API calls with decisions).
"""
from __future__ import annotations

import dataclasses
import datetime as _dt
import json
from typing import Any, Dict

from ..api import meta as M
from ..api.types import AITrainingJob
from . import elastic as E
from . import engine
from .pod import StartWindow


def observation_to_json(obs: engine.Observation) -> Dict[str, Any]:
    return {
        "job": obs.job.to_dict(), "pods": M.deepcopy(obs.pods), "services": M.deepcopy(obs.services),
        "ready_nodes": sorted(obs.ready_nodes), "now": M.format_time(obs.now), "now_epoch": obs.now_epoch,
        "options": {"window": dataclasses.asdict(obs.options.window), "scale_down_grace": obs.options.scale_down_grace,
                    "master_url": obs.options.master_url},
        "cluster": dataclasses.asdict(obs.cluster) if obs.cluster is not None else None,
        "spare_ports": list(obs.spare_ports),
    }


def observation_from_json(d: Dict[str, Any]) -> engine.Observation:
    o = d["options"]
    # the float carries the sub-second part (drain / time-limit arithmetic uses it); "now" is the readable form
    now = _dt.datetime.fromtimestamp(d["now_epoch"], _dt.timezone.utc) if d.get("now_epoch") else M.parse_time(d["now"])
    return engine.Observation(
        job=AITrainingJob.from_dict(d["job"]), pods=M.deepcopy(d["pods"]), services=M.deepcopy(d["services"]),
        ready_nodes=frozenset(d["ready_nodes"]), now=now, now_epoch=float(d["now_epoch"]),
        options=engine.EngineOptions(StartWindow(**o["window"]), float(o["scale_down_grace"]), o.get("master_url", "")),
        cluster=E.ClusterView(**d["cluster"]) if d.get("cluster") else None,
        spare_ports=tuple(d.get("spare_ports") or ()))


def decision_to_json(obs: engine.Observation, dec: engine.Decision) -> Dict[str, Any]:
    """The decision plus the status / annotations the pass left on its private copy of the job."""
    out = dataclasses.asdict(dec)
    out["pod_patches"] = [list(x) for x in dec.pod_patches]
    out["service_creates"] = [list(x) for x in dec.service_creates]
    out["service_deletes"] = [list(x) for x in dec.service_deletes]
    out["counters"] = [list(x) for x in dec.counters]
    out["observations"] = [list(x) for x in dec.observations]
    out["log"] = [list(x) for x in dec.log]
    out["role_outcomes"] = {k: list(v) for k, v in dec.role_outcomes.items()}
    out["status"] = obs.job.status.to_dict()
    out["job_annotations"] = dict(obs.job.annotations)
    return json.loads(json.dumps(out))     # tuples -> lists, keys -> str


def replay(case: Dict[str, Any]) -> Dict[str, Any]:
    obs = observation_from_json(case["observation"])
    return decision_to_json(obs, engine.reconcile(obs))


class Recorder:
    """Writes one JSON file per pass into ``directory`` (``AITJ_RECORD_DIR``)."""

    def __init__(self, directory: str):
        import os

        self.dir = directory
        os.makedirs(directory, exist_ok=True)
        self.n = 0

    def snapshot(self, obs: engine.Observation) -> Dict[str, Any]:
        return observation_to_json(obs)          # before the engine edits the job in place

    def write(self, before: Dict[str, Any], obs: engine.Observation, dec: engine.Decision) -> None:
        import os

        self.n += 1
        path = os.path.join(self.dir, f"{obs.job.namespace}-{obs.job.name}-{self.n:05d}.json")
        with open(path + ".tmp", "w") as f:
            json.dump({"observation": before, "decision": decision_to_json(obs, dec)}, f, indent=1, sort_keys=True)
        os.replace(path + ".tmp", path)
