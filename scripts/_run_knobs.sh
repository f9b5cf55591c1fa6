mkdir -p gpurun_out
for knob in "OBMAN_CONV_HALO=1"; do
  echo "== $knob"
  env $knob python scripts/ab_conv.py 2>&1 | grep -E "plain|rev\+mask\+add"
  env $knob timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
  env $knob timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value'],1))"
done
