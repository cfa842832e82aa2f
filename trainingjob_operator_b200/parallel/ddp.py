"""Data-parallel gradient synchronisation over NVLink 5 / NVSwitch for flat gradient buffers.

The reference wires N identical ``trainer`` replicas together (env contract, pod.go:553-628) and
leaves data parallelism to the framework inside the containers; here the launched
workers' DDP is part of the product.  Gradients live in one flat fp32 buffer
(``models.flat_params``), so a bucket is a slice: no flatten/unflatten copies.  Buckets are reduced
on a side stream as soon as backward finishes them (event fork/join; with CUDA graphs the collectives
are launched between graph segments, ``runtime.trainer``) and the optimizer sweep divides by the world
size while it reads the sum.

Backends: ``nccl`` (torch.distributed: ring/tree/NVLS chosen by NCCL); ``gloo`` keeps the same code path
testable on CPU.  The alternative without any collective kernel -- gradients reduced through the NVSwitch
multicast alias by the GEMM epilogues that produce them -- lives in ``parallel.symm`` (``AITJ_ALLREDUCE=mc``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


class BucketAllReducer:
    def __init__(self, flat_grad: torch.Tensor, buckets: List[Tuple[str, int, int]], group=None,
                 backend: str = "nccl", min_bucket_bytes: int = 8 << 20):
        self.flat = flat_grad
        self.group = group
        self.backend = backend
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        # merge small neighbouring buckets (they are contiguous and complete in order)
        merged: List[List] = []
        for name, a, b in buckets:
            if merged and (merged[-1][2] - merged[-1][1]) * 4 < min_bucket_bytes and \
                    (a == merged[-1][2] or b == merged[-1][1]):
                merged[-1][0] += "+" + name
                merged[-1][1] = min(merged[-1][1], a)
                merged[-1][2] = max(merged[-1][2], b)
                merged[-1][3] = name
            else:
                merged.append([name, a, b, name])
        self.buckets = [(m[0], m[1], m[2]) for m in merged]
        self.trigger: Dict[str, Tuple[int, int]] = {m[3]: (m[1], m[2]) for m in merged}
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self._works: List = []

    def hook(self, name: str) -> None:
        """Called by the engine when bucket ``name`` is final: reduce its slice asynchronously."""
        if self.world <= 1 or name not in self.trigger:
            return
        a, b = self.trigger[name]
        view = self.flat[a:b]
        if not self.cuda:
            self._works.append(dist.all_reduce(view, group=self.group, async_op=True))
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self._works.append(dist.all_reduce(view, group=self.group, async_op=True))

    def wait(self) -> None:
        """Join: the current stream waits for every outstanding bucket reduction."""
        for w in self._works:
            w.wait()
        self._works.clear()
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def all_reduce_all(self) -> None:
        if self.world <= 1:
            return
        dist.all_reduce(self.flat, group=self.group)


def broadcast_state(tensors: List[torch.Tensor], src: int = 0, group=None) -> None:
    """State hand-off after a (re-)rendezvous: survivors -> joiners over NVLink."""
    if not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
