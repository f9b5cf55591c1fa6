timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|bias\+relu\+add|rev\+mask\+add"
timeout 120 python scripts/trace_conv64.py | tail -36
