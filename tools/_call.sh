timeout 120 python -m trainingjob_operator_b200.ops.selfcheck --case attention 2>&1 | grep -v " ok$" | tail -5
AITJ_ATTN_STAGGER=0 timeout 120 python tools/attn_trace.py 2>&1 | tail -8
