OBMAN_CONV64_CFG=24 timeout 120 python scripts/trace_conv64.py
OBMAN_CONV64_CFG=18 timeout 120 python scripts/trace_conv64.py | head -16
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
