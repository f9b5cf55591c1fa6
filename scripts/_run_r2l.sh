mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "conv or dgrad" 2>&1 | tail -3
echo "== OBMAN_CONV64_GEN=2 (ring 3, 16-col rounds)"
AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|bias\+relu\+add|rev\+mask\+add"
echo "== ring capped at 2"
OBMAN_CONV64_RING=2 AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|rev\+mask\+add"
