timeout 120 python -m trainingjob_operator_b200.ops.selfcheck --case attention 2>&1 | tail -15
timeout 120 python tools/attn_trace.py 2>&1 | tail -11
