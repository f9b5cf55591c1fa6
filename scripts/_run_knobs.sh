mkdir -p gpurun_out
for knob in "OBMAN_GEMM_DEEP_SMALL=0" "OBMAN_GEMM_DEEP_SMALL=1"; do
  echo "== $knob"
  env $knob python scripts/ab_conv.py 2>&1 | grep -E "512 plain|512 rev\+mask\+add"
  env $knob timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value'],1), 'gemm', round(d['roofline']['gemm_ms_per_step'],3))"
done
OBMAN_GEMM_DEEP_SMALL=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py tests/test_gpu_handnet.py -m gpu -x -q 2>&1 | tail -2
