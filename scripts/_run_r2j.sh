mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "conv or dgrad" 2>&1 | tail -12
for gen in 1 2; do
  echo "== OBMAN_CONV64_GEN=$gen"
  OBMAN_CONV64_GEN=$gen AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|bias\+relu\+add|rev\+mask\+add"
done
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_handnet.py -m gpu -x -q 2>&1 | tail -4
