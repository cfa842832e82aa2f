"""Shared by the GPU check tools (fault_check / elastic_gpu_check / failover_check): when a wait runs out, leave what is
needed to see why -- the stage, the job's status / annotations, the supervised processes and the tail of every worker log
-- under ``gpurun_out/`` (a GPU box is gone after the call; round 2 lost a 4-GPU call to five silent time-outs)."""
import json
import os
import time


def dump(lc, job_name: str, context: dict, stage: str, err: BaseException, tag: str) -> str:
    rec = dict(context, failed_at=stage, error=f"{type(err).__name__}: {err}")
    try:
        j = lc.jobs().get(job_name)
        rec["phase"] = j.status.phase
        rec["conditions"] = [f"{c.type}: {c.message}" for c in j.status.conditions][-8:]
        rec["restart_counts"] = j.status.restart_counts
        rec["annotations"] = {k: v[:400] for k, v in (j.annotations or {}).items() if k.startswith("aitj.b200/")}
        rec["rendezvous"] = getattr(j.status, "rendezvous", None) and j.status.rendezvous.__dict__
    except Exception as e:  # noqa: BLE001
        rec["job_read_error"] = repr(e)
    try:
        rec["processes"] = [sid for sid, _ in lc.agent.sup.list()]
        logs = os.path.join(lc.workdir, "logs")
        rec["logs"] = {fn: open(os.path.join(logs, fn), errors="replace").read()[-3000:]
                       for fn in sorted(os.listdir(logs)) if fn.endswith(".log")}
    except Exception as e:  # noqa: BLE001
        rec["log_read_error"] = repr(e)
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/{tag}_FAILED_{int(time.time())}.json"
    json.dump(rec, open(path, "w"), indent=1, default=str)
    print(json.dumps({"failed_at": stage, "error": rec["error"], "post_mortem": path}), flush=True)
    return path
