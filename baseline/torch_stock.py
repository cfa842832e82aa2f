"""Comparator arm: what a user container under the reference operator would run.

The reference operator contains no training code: its example job runs the framework's own stock
``train.py`` (/root/reference/example/paddle-mnist.yaml:20-21) and the operator only injects the rank /
peer environment (/root/reference/pkg/controller/pod.go:548-652).  The honest stand-in for "the
reference driving the same job on the same box" is therefore the *same* GPT-2 124M configuration written
the way a stock PyTorch user writes it:

  ``nn.Module`` (nanoGPT-shaped GPT-2 small, tied embeddings, tanh-GELU) + ``DistributedDataParallel``
  + ``torch.optim.AdamW(fused=True)`` + ``F.scaled_dot_product_attention`` + bf16 autocast
  + ``clip_grad_norm_(1.0)``, eager or ``torch.compile``-d, launched by ``torchrun``.

None of this repo's kernels, engine or runtime is imported here -- ``run`` asserts that neither
``libaitj_kernels.so`` nor ``_aitj_core`` is mapped into the process.  Timing rules are the same as for the
product arm (bench.py): W warm-up steps, K steps bracketed by barrier + synchronize, CUDA events, max over
ranks; every step copies its tokens from pinned host memory and reads the loss back.
"""
from __future__ import annotations

import math
import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, C: int, H: int):
        super().__init__()
        self.H = H
        self.ln1 = nn.LayerNorm(C)
        self.qkv = nn.Linear(C, 3 * C)
        self.proj = nn.Linear(C, C)
        self.ln2 = nn.LayerNorm(C)
        self.fc = nn.Linear(C, 4 * C)
        self.fc2 = nn.Linear(4 * C, C)

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.qkv(self.ln1(x)).split(C, dim=2)
        q, k, v = (t.view(B, T, self.H, C // self.H).transpose(1, 2) for t in (q, k, v))
        y = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        x = x + self.proj(y.transpose(1, 2).reshape(B, T, C))
        return x + self.fc2(F.gelu(self.fc(self.ln2(x)), approximate="tanh"))


class GPT2(nn.Module):
    """GPT-2 small, vocabulary padded to 50304 (the usual nanoGPT choice; the product arm pads the same way)."""

    def __init__(self, vocab=50304, n_layer=12, n_head=12, n_embd=768, block_size=1024):
        super().__init__()
        self.wte = nn.Embedding(vocab, n_embd)
        self.wpe = nn.Embedding(block_size, n_embd)
        self.h = nn.ModuleList(Block(n_embd, n_head) for _ in range(n_layer))
        self.lnf = nn.LayerNorm(n_embd)
        self.lm_head = nn.Linear(n_embd, vocab, bias=False)
        self.lm_head.weight = self.wte.weight
        self.apply(self._init)
        for n, p in self.named_parameters():
            if n.endswith("proj.weight") or n.endswith("fc2.weight"):
                nn.init.normal_(p, mean=0.0, std=0.02 / math.sqrt(2 * n_layer))

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)

    def forward(self, tok, tgt):
        B, T = tok.shape
        x = self.wte(tok) + self.wpe(torch.arange(T, device=tok.device))
        for blk in self.h:
            x = blk(x)
        logits = self.lm_head(self.lnf(x))
        return F.cross_entropy(logits.view(B * T, -1).float(), tgt.view(-1))


def _native_maps():
    try:
        return [ln.split()[-1] for ln in open("/proc/self/maps") if "libaitj_kernels" in ln or "_aitj_core" in ln]
    except OSError:
        return []


def run(batch: int, seq: int, steps: int, warmup: int, compiled: bool, lr: float = 3e-4, seed: int = 0) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(seed)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    model = GPT2().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    decay = [p for p in model.parameters() if p.dim() >= 2]
    no_decay = [p for p in model.parameters() if p.dim() < 2]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}],
                            lr=lr, betas=(0.9, 0.95), eps=1e-8, fused=True)
    fwd = torch.compile(model) if compiled else model
    ddp = nn.parallel.DistributedDataParallel(fwd, device_ids=[local], gradient_as_bucket_view=True) if world > 1 \
        else fwd
    g = torch.Generator().manual_seed(seed + 1 + rank)
    host = []
    for _ in range(4):
        tok = torch.randint(0, 50257, (batch, seq), generator=g, dtype=torch.int64)
        host.append((tok.pin_memory(), torch.roll(tok, -1, dims=1).pin_memory()))
    tok_d = torch.empty(batch, seq, dtype=torch.int64, device=dev)
    tgt_d = torch.empty(batch, seq, dtype=torch.int64, device=dev)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def step(i: int) -> float:
        tok, tgt = host[i % len(host)]
        tok_d.copy_(tok, non_blocking=True)
        tgt_d.copy_(tgt, non_blocking=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = ddp(tok_d, tgt_d)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[0])

    t_c0 = time.time()
    losses = [step(i) for i in range(warmup)]
    compile_s = time.time() - t_c0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        losses.append(step(warmup + i))
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    maps = _native_maps()
    assert not maps, f"comparator arm must not load this repo's native code: {maps}"
    return {"ms_per_step": ms, "samples_per_sec": batch * world / (ms / 1e3), "global_batch": batch * world,
            "batch_per_gpu": batch, "seq_len": seq, "params": n_params, "loss_first": losses[0],
            "loss_last": losses[-1], "warmup_wall_s": round(compile_s, 2), "compiled": compiled,
            "h2d_bytes_per_step": 2 * batch * seq * 8 * world, "d2h_bytes_per_step": 4 * world,
            "torch": torch.__version__, "repo_native_code_mapped": maps}
