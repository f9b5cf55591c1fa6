timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "conv or dgrad" 2>&1 | tail -3
AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|bias\+relu\+add|rev\+mask\+add"
timeout 600 python -m pytest tests/test_gpu_bn_train.py -m gpu -q -s 2>&1 | grep -v Warning | grep -E "rel err|worst|passed|failed|Error|assert" | head -30
