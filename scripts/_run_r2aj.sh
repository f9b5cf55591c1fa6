mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests_final.log
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/tc_launches_final.txt > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'frac', r.get('frac'), 'gemm_ms', r.get('gemm_ms_per_step'), 'traffic', r.get('traffic'), d.get('clocks'), 'launches', d.get('gpu_launches'))
for k, v in (d.get('secondary') or {}).items():
    print('  secondary', k[:12], {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != 'workload'})
ch = d.get('chamfer', {})
for k in ('workload', 'sweep_2048x10000'):
    print('  chamfer', k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in ch.get(k, {}).items() if a != 'bwd'}, 'bwd', ch.get(k, {}).get('bwd'))
PY
