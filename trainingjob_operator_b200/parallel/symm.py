"""Symmetric (peer-mapped + NVSwitch-multicast) gradient buffer for the fused GEMM -> all-reduce path.

The DDP hot op is "weight-gradient GEMM followed by a gradient all-reduce".  Instead of calling a
collective after the GEMM, the gradient buffer is allocated as CUDA symmetric memory with an NVLS
multicast alias (``torch.distributed._symmetric_memory`` provides the allocation/rendezvous plumbing), and
every gradient-producing kernel of the backward pass -- the tcgen05 wgrad GEMM epilogue, LayerNorm-backward,
the bias column-reduce, the embedding scatter -- issues ``multimem.red.add`` on that alias: the NVSwitch
applies each contribution to all peers' buffers while the math of the next tile proceeds.  When backward
ends, every rank already holds the summed gradient; two device-side barriers per step replace all
gradient collectives.  (The reference has no collective code at all)

Falls back (``available == False``) when the platform has no multicast support or the world size is 1.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class SymmetricGradBuffer:
    def __init__(self, numel: int, device: torch.device, group=None):
        self.available = False
        self.reason = ""
        self.tensor: Optional[torch.Tensor] = None
        self.handle = None
        self.multicast_ptr = 0
        self._chan = 0
        if not dist.is_initialized() or dist.get_world_size(group) <= 1:
            self.reason = "world size 1"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem

            pg = group if group is not None else dist.group.WORLD
            try:
                symm_mem.enable_symm_mem_for_group(pg.group_name)
            except Exception:  # noqa: BLE001 - newer torch enables lazily
                pass
            t = symm_mem.empty(numel, dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, pg.group_name)
            mc = int(getattr(h, "multicast_ptr", 0) or 0)
            if mc == 0:
                self.reason = "no NVLS multicast support on this platform"
                return
            t.zero_()
            self.tensor, self.handle, self.multicast_ptr = t, h, mc
            self.available = True
        except Exception as e:  # noqa: BLE001
            self.reason = f"{type(e).__name__}: {e}"

    def barrier(self) -> None:
        """Device-side barrier across ranks on the current stream (release/acquire at system scope)."""
        self.handle.barrier(channel=self._chan)
        self._chan = (self._chan + 1) % 2
