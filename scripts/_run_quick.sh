# quick A/B: GPU tests + bench with the env given on the command line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/tc_launches.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.err
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_c3"):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, d['dtype'], round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1),
              'tc TF/s', round(d['roofline']['achieved'], 1), 'gemm ms', round(d['roofline']['gemm_ms_per_step'], 3), d['clocks'])
    except Exception as e:
        print(f, 'failed', e)
PY
