# Round 2, first call: stacked-N bring-up (numerics then speed), config 3 / config 4 baselines with per-launch tables,
# ncu launch list of one config 3 step.
mkdir -p gpurun_out
echo "== STACK64 numerics"
OBMAN_GEMM_STACK64=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -5
OBMAN_WGRAD_STACK64=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -5
echo "== A/B"
for knob in "OBMAN_GEMM_STACK64=0" "OBMAN_GEMM_STACK64=1"; do
  echo "-- $knob"
  env $knob AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64"
done
for knob in "OBMAN_WGRAD_STACK64=0" "OBMAN_WGRAD_STACK64=1"; do
  echo "-- $knob"
  env $knob PROF_B=256 PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=10 timeout 120 python scripts/prof_kernels.py 2>&1 | grep -E "wgrad"
done
echo "== config 3"
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/tc_launches_c3.txt > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"
OBMAN_GEMM_STACK64=1 OBMAN_WGRAD_STACK64=1 timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/tc_launches_c3_stack.txt > gpurun_out/bench_c3_stack.json 2> gpurun_out/bench_c3_stack.err; echo "rc=$?"
echo "== config 4 (1 GPU shard)"
timeout 400 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_c3", "bench_c3_stack", "bench_c4"):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
        try:
            print(open('gpurun_out/%s.err' % f).read()[-1500:])
        except Exception:
            pass
PY
echo "== ncu launch list, config 3"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_c3.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches_c3.csv
