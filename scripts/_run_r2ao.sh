mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -q -x -k "gemm or decoder" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-gpu-eager --quick --dump-launches gpurun_out/tc_launches_r2ao.txt > gpurun_out/bench_r2ao.json 2> gpurun_out/bench_r2ao.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2ao.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), 'traffic', r.get('traffic'), r.get('traffic_source'))
PY
grep "gemm M655872" gpurun_out/tc_launches_r2ao.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --config 3 --steps 1 --warmup 1 --no-graph --quick > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
