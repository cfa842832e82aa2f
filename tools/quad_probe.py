"""Probe of the 4-CTA cluster GEMM: how many clusters fit, and time vs the CTA-pair kernel on the step's shapes."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trainingjob_operator_b200.ops import functional as F, lib  # noqa: E402
from tools.kernel_bench import timeit  # noqa: E402

M = 16384
for name, n, k in (("qkv fwd", 2304, 768), ("fc fwd", 3072, 768), ("fc2 fwd", 768, 3072), ("sq4096", 4096, 4096)):
    m = 4096 if name == "sq4096" else M
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for bn in (1024, 512):
        med, best = timeit(lambda: F.gemm(a, b, out, block_n=bn))
        print(f"{name:10s} bn={bn:5d} {med * 1e3:8.1f} us  ({2.0 * m * n * k / med / 1e9:7.1f} TF)", flush=True)
L = lib.load()
L.aitj_gemm_quad_clusters.restype = ctypes.c_int
print("cudaOccupancyMaxActiveClusters(4-CTA cluster):", L.aitj_gemm_quad_clusters(), "override:",
      os.environ.get("AITJ_GEMM_QUAD_CLUSTERS"))
