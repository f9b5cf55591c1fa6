# Final evidence run of the round: smoke, GPU tests, headline bench (+ per-launch table), reference arm, ncu launch list
# with DRAM traffic of the same bench command, --set full capture of representative tensor-core kernels.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 --dump-launches gpurun_out/tc_launches.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_c3", "bench_ref"):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, d.get('dtype'), round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), d.get('roofline', {}).get('frac'), d.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on \
   -k regex:'gemm_tc_kernel|wgrad_bf16_kernel' -c 10 -o gpurun_out/prof_r1j -f python scripts/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
