# Next-round bring-up of the stacked-N variants (gemm_tc_kernel<64,3,0,3,2>: OBMAN_GEMM_STACK64=1;
# wgrad_bf16_kernel<64,2,2>: OBMAN_WGRAD_STACK64=1): numerics first, then speed.
mkdir -p gpurun_out
OBMAN_GEMM_STACK64=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -5
OBMAN_WGRAD_STACK64=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -5
OBMAN_WGRAD_STACK64=1 PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=20 timeout 120 python scripts/prof_kernels.py 2>&1 | grep "wgrad3x3 64x64"
for knob in "OBMAN_GEMM_STACK64=0" "OBMAN_GEMM_STACK64=1" "OBMAN_GEMM_STACK64=1 OBMAN_WGRAD_STACK64=1"; do
  echo "== $knob"
  env $knob python scripts/ab_conv.py 2>&1 | grep -E "c64->64"
  env $knob timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value'],1), 'gemm', round(d['roofline']['gemm_ms_per_step'],3))"
done
