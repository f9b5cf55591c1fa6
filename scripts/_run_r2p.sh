mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --quick --dump-launches gpurun_out/tc_launches_c3_r2p.txt > gpurun_out/bench_c3_r2p.json 2> gpurun_out/bench_c3_r2p.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c3_r2p.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'), 'launches', d.get('gpu_launches'))
PY
