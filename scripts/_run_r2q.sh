timeout 120 python scripts/trace_conv64.py | tail -14
timeout 600 python -m pytest tests/test_gpu_bn_train.py -m gpu -q -s 2>&1 | grep -v Warning | grep -E "rel err|worst|passed|failed|Error|assert|total" | head -30
