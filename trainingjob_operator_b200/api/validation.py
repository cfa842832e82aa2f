"""Admission validation for ``AITrainingJob``.

The reference only *intended* validation: pkg/apis/aitrainingjob/validation/validation.go:14-32
is dead code that does not compile and the controller carries
``// FIXME: need to validate trainingjob`` (pkg/controller/trainingjob.go:21,33).  This module
implements that intent (non-nil replicaSpecs, >=1 container per role, non-empty image) for real
and adds what the live-elastic semantics need: known enum values, parseable
``restartingExitCode``, ``minReplicas <= replicas <= maxReplicas``, sane integers.  A local
process needs a ``command``/``args``; ``image`` alone is accepted when the launcher has an image
map (the Paddle example keeps both).
"""
from __future__ import annotations

import re
from typing import Any, Dict, List

from . import constants as C
from .types import AITrainingJob

_DNS1123 = re.compile(r"^[a-z0-9]([-a-z0-9.]*[a-z0-9])?$")


class ValidationError(ValueError):
    def __init__(self, errors: List[str]):
        super().__init__("; ".join(errors))
        self.errors = errors


def parse_exit_codes(s: str) -> List[int]:
    """``"137,128"`` -> [137, 128]; raises ValueError on junk."""
    out = []
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        out.append(int(part))
    return out


def validate_replica_specs(specs, path: str = "spec.replicaSpecs") -> List[str]:
    errs: List[str] = []
    if not specs:
        return [f"{path}: Required value: AITrainingJob spec is not valid: replicaSpecs must not be empty"]
    for rtype, spec in specs.items():
        p = f"{path}[{rtype}]"
        if not rtype or not _DNS1123.match(rtype.lower()):
            errs.append(f"{p}: Invalid value: role name must be a DNS-1123 label")
        if spec is None:
            errs.append(f"{p}: Required value")
            continue
        containers = spec.containers()
        if not containers:
            errs.append(f"{p}.template.spec.containers: Required value: replica spec must have at least one container")
        names = set()
        for i, c in enumerate(containers + spec.init_containers()):
            cp = f"{p}.template.spec.containers[{i}]"
            if not c.get("name"):
                errs.append(f"{cp}.name: Required value")
            elif c["name"] in names:
                errs.append(f"{cp}.name: Duplicate value: {c['name']!r}")
            names.add(c.get("name"))
            if not c.get("image") and not c.get("command") and not c.get("args"):
                errs.append(f"{cp}.image: Required value: container needs an image or a command")
            for j, port in enumerate(c.get("ports") or []):
                cport = port.get("containerPort")
                if not isinstance(cport, int) or not (0 < cport < 65536):
                    errs.append(f"{cp}.ports[{j}].containerPort: Invalid value: {cport!r}")
        for fname, jname in (("replicas", "replicas"), ("min_replicas", "minReplicas"),
                             ("max_replicas", "maxReplicas"), ("restart_limit", "restartLimit")):
            v = getattr(spec, fname)
            if v is not None and (not isinstance(v, int) or isinstance(v, bool) or v < 0):
                errs.append(f"{p}.{jname}: Invalid value: {v!r}: must be a non-negative integer")
        r = spec.replicas if spec.replicas is not None else 1
        if isinstance(r, int):
            if isinstance(spec.min_replicas, int) and r < spec.min_replicas:
                errs.append(f"{p}.replicas: Invalid value: {r}: must be >= minReplicas ({spec.min_replicas})")
            if isinstance(spec.max_replicas, int) and r > spec.max_replicas:
                errs.append(f"{p}.replicas: Invalid value: {r}: must be <= maxReplicas ({spec.max_replicas})")
        if isinstance(spec.min_replicas, int) and isinstance(spec.max_replicas, int) and \
                spec.min_replicas > spec.max_replicas:
            errs.append(f"{p}.minReplicas: Invalid value: {spec.min_replicas}: must be <= maxReplicas")
        for fname, jname, allowed in (("restart_policy", "restartPolicy", C.RESTART_POLICIES),
                                      ("restart_scope", "restartScope", C.RESTART_SCOPES),
                                      ("fail_policy", "failPolicy", C.ENDING_POLICIES),
                                      ("complete_policy", "completePolicy", C.ENDING_POLICIES),
                                      ("edl_policy", "edlPolicy", C.EDL_POLICIES)):
            v = getattr(spec, fname)
            if v and v not in allowed:
                errs.append(f"{p}.{jname}: Unsupported value: {v!r}: supported values: {', '.join(allowed)}")
    return errs


def validate_aitrainingjob(job: AITrainingJob) -> List[str]:
    errs: List[str] = []
    if job.api_version != C.API_VERSION:
        errs.append(f"apiVersion: Invalid value: {job.api_version!r}: expected {C.API_VERSION}")
    if job.kind != C.KIND:
        errs.append(f"kind: Invalid value: {job.kind!r}: expected {C.KIND}")
    name = job.name
    if not name:
        errs.append("metadata.name: Required value")
    elif not _DNS1123.match(name) or len(name) > 253:
        errs.append(f"metadata.name: Invalid value: {name!r}: must be a lowercase RFC 1123 subdomain")
    spec = job.spec
    errs += validate_replica_specs(spec.replica_specs)
    try:
        parse_exit_codes(spec.restarting_exit_code)
    except ValueError:
        errs.append(f"spec.restartingExitCode: Invalid value: {spec.restarting_exit_code!r}: "
                    "must be a comma separated list of integers")
    if spec.time_limit is not None and (not isinstance(spec.time_limit, int) or spec.time_limit < 0):
        errs.append(f"spec.timeLimit: Invalid value: {spec.time_limit!r}: must be a non-negative integer (seconds)")
    if spec.clean_pod_policy is not None and spec.clean_pod_policy not in C.CLEAN_POD_POLICIES:
        errs.append(f"spec.cleanPodPolicy: Unsupported value: {spec.clean_pod_policy!r}: supported values: All, None")
    for fname, jname in (("fail_policy", "failPolicy"), ("complete_policy", "completePolicy")):
        v = getattr(spec, fname)
        if v and v not in C.ENDING_POLICIES:
            errs.append(f"spec.{jname}: Unsupported value: {v!r}: supported values: {', '.join(C.ENDING_POLICIES)}")
    return errs


def validate_dict(obj: Dict[str, Any]) -> List[str]:
    try:
        job = AITrainingJob.from_dict(obj)
    except Exception as e:  # noqa: BLE001 - malformed structure
        return [f"malformed object: {e}"]
    if not isinstance(obj.get("spec", {}).get("replicaSpecs", {}), dict):
        return ["spec.replicaSpecs: Invalid value: must be a map of role name to replica spec"]
    return validate_aitrainingjob(job)


def validate_or_raise(obj: Dict[str, Any]) -> None:
    errs = validate_dict(obj)
    if errs:
        raise ValidationError(errs)
