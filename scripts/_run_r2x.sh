mkdir -p gpurun_out
timeout 300 python scripts/debug_capture.py 64 4 2>&1 | grep -v Warning | tail -8
timeout 300 python -m pytest tests/test_gpu_handnet.py -q -x -k "training_driver" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_dense.py -q -x -k "tail or decoder or linear" 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick --dump-launches gpurun_out/tc_launches_r2x.txt > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2x.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'))
PY
grep "gemm M655872" gpurun_out/tc_launches_r2x.txt | head
