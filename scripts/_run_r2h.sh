mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --dump-launches gpurun_out/tc_launches_c3_r2h.txt > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2h.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2h.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'), 'launches', d.get('gpu_launches'))
print('secondary', json.dumps(d.get('secondary'))[:900])
print('eager', d.get('gpu_eager_baseline'))
print('cpu', d.get('cpu_baseline'))
print('chamfer', json.dumps(d.get('chamfer'))[:1200])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2h.json 2> gpurun_out/bench_ref_r2h.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_r2h.json
