timeout 120 python scripts/trace_conv64.py | tail -18
