"""Per-phase clock64 accounting of the attention forward kernel's softmax warp (GPT-2 small shape)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.ops import functional as F, lib  # noqa: E402

B, T, H = 16, 1024, 12
C = H * 64
qkv = (torch.randn(B * T, 3 * C, device="cuda")).bfloat16()
out = torch.empty(B * T, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
n_cta = min((T // 128) * B * H, 2 * F.num_sms())
tr = torch.zeros(n_cta * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    F.attention_fwd(qkv, out, lse, B, T, H)
torch.cuda.synchronize()
lib.load().aitj_attn_set_trace(tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
F.attention_fwd(qkv, out, lse, B, T, H)
e1.record()
torch.cuda.synchronize()
lib.load().aitj_attn_set_trace(None)
t = tr.view(n_cta, 8).cpu().double()
nb = t[:, 6]
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us; blocks total {nb.sum():.0f}")
names = ["total", "wait S", "pass1 max", "pass2 exp+P", "wait O", "accumulate"]
for i, n in enumerate(names):
    print(f"  {n:12s} per block: {(t[:, i].sum() / nb.sum()):8.0f} clk")
print(f"  per CTA: blocks {nb.min():.0f}..{nb.max():.0f}, total clk {t[:, 0].min():.0f}..{t[:, 0].max():.0f}")
