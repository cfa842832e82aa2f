"""GPU numerics tests: each hand-written sm_100a kernel vs. a plain fp32 PyTorch reference.

Cases run in their own process (``ops.selfcheck``) so one trapping kernel cannot poison the
CUDA context of the others.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_case(name: str, timeout: int = 420):
    r = subprocess.run([sys.executable, "-m", "trainingjob_operator_b200.ops.selfcheck", "--case", name], cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0, f"selfcheck {name} failed:\n{r.stdout[-3000:]}\n{r.stderr[-1500:]}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["gemm_tn", "gemm_nn", "gemm_tt", "gemm_epilogue", "gemm_2cta", "fused_ops", "gpt2_engine"])
def test_kernel_numerics(case):
    _run_case(case)


@pytest.mark.gpu
def test_kernel_library_is_loaded_not_a_fallback():
    import torch

    from trainingjob_operator_b200.ops import functional as F
    from trainingjob_operator_b200.ops import lib

    lib.load(build_if_missing=False)
    a = torch.randn(256, 128, device="cuda").bfloat16()
    b = torch.randn(256, 128, device="cuda").bfloat16()
    out = torch.empty(256, 256, device="cuda", dtype=torch.bfloat16)
    before = lib.LAUNCHES
    F.gemm(a, b, out)
    torch.cuda.synchronize()
    assert lib.LAUNCHES == before + 1
    maps = open("/proc/self/maps").read()
    assert "libaitj_kernels.so" in maps
