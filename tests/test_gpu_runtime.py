"""GPU checks of the worker runtime pieces around the kernels (run in a child process with a timeout, like the kernel
cases: a wedged CUDA context must not take the test session with it)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_async_checkpoint_snapshots_in_stream_order_on_the_device(tmp_path):
    script = textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r)
        from trainingjob_operator_b200.runtime import checkpoint as ck
        dev = torch.device("cuda", 0)
        state = [torch.arange(1 << 22, dtype=torch.float32, device=dev), torch.ones(1 << 20, device=dev)]
        w = ck.AsyncCheckpointer(%r)
        for step in (1, 2, 3):
            for t in state:
                t.add_(1.0)                       # "optimizer step" on the training stream
            while not w.save(step, state):       # snapshot is ordered after it, the next add_ after the snapshot
                w.wait(60)
        for t in state:
            t.add_(100.0)                         # must not leak into the checkpoint that is still draining
        assert w.wait(60)
        out = [torch.zeros_like(t) for t in state]
        meta = ck.load_into(%r, out)
        torch.cuda.synchronize()
        assert meta["step"] == 3, meta
        assert torch.equal(out[0].cpu(), torch.arange(1 << 22, dtype=torch.float32) + 3.0)
        assert float(out[1].min()) == 4.0 and float(out[1].max()) == 4.0
        print("OK", w.stats)
    """ % (ROOT, str(tmp_path / "s.pt"), str(tmp_path / "s.pt")))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.gpu2
def test_fault_tolerant_recovery_keeps_the_survivor_on_its_gpu(tmp_path):
    """In-place recovery over NCCL (needs two GPUs): SIGKILL rank 1 of a faultTolerant BERT-shaped job; rank 0 keeps its
    process, CUDA context and state (the StallBreaker aborts the communicator it is stuck in), the replacement joins."""
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fault_check.py"), "bert", "2", "0",
                        "--fault-tolerant", "--victim", "1"], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=420)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["survivors_kept_their_process"] and out["restart_counts"] == {"trainer": 1}
    assert out["recovery"]["world"] == 2 and out["recovery"]["recovered_from"]
