for dbg in 0 1 2 4 8 3 7 15; do
  echo "== OBMAN_CONV64_DEBUG=$dbg"
  OBMAN_CONV64_CFG=24 OBMAN_CONV64_DEBUG=$dbg AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|rev\+mask\+add"
done
timeout 300 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_handnet.py -m gpu -x -q 2>&1 | tail -3
