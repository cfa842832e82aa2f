"""Which GEMM kernel (CTA pair 256x256 vs 1-CTA 128x256 / 128x128, split-K) each GEMM of the GPT-2 step should use,
measured where it matters: in the captured training step, one change at a time against the same-box baseline.

    python tools/gemm_cfg_sweep.py [--steps 40] > gpurun_out/gemm_cfg_sweep.jsonl
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = sys.argv[sys.argv.index("--steps") + 1] if "--steps" in sys.argv else "40"
CANDIDATES = [
    "", "fwd:2304x768=256", "fwd:3072x768=256", "fwd:768x3072=256", "fwd:768x768=256", "fwd:50304x768=256",
    "dgrad:768x2304=256", "dgrad:768x3072=256", "dgrad:3072x768=256", "dgrad:768x768=256", "dgrad:768x50304=256",
    "wgrad:2304x768=256/8", "wgrad:3072x768=256/2", "wgrad:768x3072=256/2", "wgrad:768x768=256/8",
    "wgrad:50304x768=256/1", "wgrad:2304x768=128/4", "wgrad:768x768=128/8", "",
]


def run(cfg):
    env = dict(os.environ, AITJ_GEMM_CFG=cfg)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", steps, "--warmup", "5",
                        "--no-e2e"], env=env, capture_output=True, text=True, timeout=300)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])["ms_per_step"]
    except Exception:  # noqa: BLE001
        return None


base = None
for cfg in CANDIDATES:
    ms = run(cfg)
    if cfg == "" and base is None:
        base = ms
    print(json.dumps({"cfg": cfg or "(baseline)", "ms_per_step": ms,
                      "delta_ms": None if ms is None or base is None else round(ms - base, 4)}), flush=True)
