# Decoder tail columns (TAIL variant), raycast block sizing, maxpool backward ILP, fused layer-1 backward: parity + timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r2w.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests_r2w.log
timeout 500 python bench.py --steps 20 --warmup 5 --no-secondary --no-gpu-eager --dump-launches gpurun_out/tc_launches_r2w.txt > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2w.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'frac', r.get('frac'), 'gemm_ms', r.get('gemm_ms_per_step'), d.get('clocks'), 'launches', d.get('gpu_launches'))
PY
grep "gemm M655872" gpurun_out/tc_launches_r2w.txt | head
OBMAN_GEMM_TAIL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TAIL=0', round(d['ms_per_step'],3))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv \
   --log-file gpurun_out/launches_r2w.csv python bench.py --config 3 --steps 1 --warmup 1 --no-graph --quick > gpurun_out/ncu_bench_r2w.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r2w.csv')) if len(r) > 10 and r[0].isdigit()]
# last step = after the last but one adam_kernel
names = [r[4] for r in rows]
adam = [i for i, n in enumerate(names) if 'adam_kernel' in n]
seg = rows[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for r in seg:
    n = r[4].split('(')[0][-60:]
    agg[n][0] += 1; agg[n][1] += float(r[-1]) / 1000.0
tot = sum(v[1] for v in agg.values())
print(len(seg), 'launches', round(tot, 1), 'us')
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%-62s %4d %9.1f' % (n, v[0], v[1]))
PY
