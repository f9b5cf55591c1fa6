# Evidence run of the round: smoke, GPU tests, headline bench (+ per-launch table), reference arm, CUDA-core kernel timings,
# ncu launch list with DRAM traffic of one configs[2] step, --set full capture of representative tensor-core kernels.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests_final.log
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/tc_launches_final.txt > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'frac', r.get('frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'), 'launches', d.get('gpu_launches'))
for k, v in (d.get('secondary') or {}).items():
    print('  secondary', k[:12], {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != 'workload'})
print('  eager', d.get('gpu_eager_baseline'))
print('  cpu', d.get('cpu_baseline'))
ch = d.get('chamfer', {})
for k in ('workload', 'sweep_2048x10000'):
    print('  chamfer', k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in ch.get(k, {}).items() if a != 'bwd'}, 'bwd', ch.get(k, {}).get('bwd'))
r = json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[-1])
print('  reference arm', r['value'], r['steps'], r['warmup'], r['cpu_baseline'])
PY
timeout 100 python scripts/time_nn.py 2>&1 | grep "^nn" > gpurun_out/cuda_core_kernels_final.txt
timeout 100 python scripts/time_raycast.py 2>&1 | grep "^raycast" >> gpurun_out/cuda_core_kernels_final.txt
cat gpurun_out/cuda_core_kernels_final.txt
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --config 3 --steps 1 --warmup 1 --no-graph --quick > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
PROF_B=256 PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on \
   -k regex:'gemm_tc_kernel|wgrad_bf16_kernel|conv64_' -c 9 -o gpurun_out/prof_r2 -f python scripts/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
