mkdir -p gpurun_out
for knob in "X=0" "OBMAN_GEMM_CLUSTER=2" "OBMAN_GEMM_CLUSTER=4" "OBMAN_GEMM_SMALLGRID_BN=64"; do
  echo "== $knob"
  env $knob python scripts/ab_conv.py 2>&1 | grep -E "plain|rev\+mask\+add"
  env $knob timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value'],1))"
done
