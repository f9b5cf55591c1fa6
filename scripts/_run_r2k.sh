mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "conv or dgrad" 2>&1 | tail -3
for gen in 1 2; do
  echo "== OBMAN_CONV64_GEN=$gen"
  OBMAN_CONV64_GEN=$gen AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64" | grep -E "plain|bias\+relu\+add|rev\+mask\+add"
done
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_handnet.py tests/test_gpu_dense.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --quick > gpurun_out/bench_c3_r2k.json 2> gpurun_out/bench_c3_r2k.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c3_r2k.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'), 'launches', d.get('gpu_launches'))
PY
