timeout 120 python scripts/trace_conv64.py
timeout 600 python -m pytest tests/test_gpu_bn_train.py tests/test_gpu_encoder.py -m gpu -x -q -s 2>&1 | grep -v Warning | tail -15
