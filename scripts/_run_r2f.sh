timeout 300 python -m pytest tests/test_gpu_handnet.py -m gpu -x -q -k "training_driver" 2>&1 | grep -E "Error|error|assert|^E |line " | head -30
for ring in 2 3; do
  echo "== ring $ring cfg 24"
  OBMAN_CONV64_RING=$ring OBMAN_CONV64_CFG=24 AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|rev\+mask\+add"
done
echo "== ring 3 cfg 24, no MMAs"
OBMAN_CONV64_DEBUG=1 OBMAN_CONV64_CFG=24 AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|rev\+mask\+add"
echo "== ring 3 cfg 34"
OBMAN_CONV64_CFG=34 AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|rev\+mask\+add"
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "conv or dgrad" 2>&1 | tail -2
