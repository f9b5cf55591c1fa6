mkdir -p gpurun_out
OBMAN_NN_PACKED=1 timeout 200 python scripts/time_nn.py 2>&1 | grep "^nn"
OBMAN_NN_PACKED=0 timeout 200 python scripts/time_nn.py 2>&1 | grep "^nn"
timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_losshead.py -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick > gpurun_out/bench_r2ab.json 2> gpurun_out/bench_r2ab.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2ab.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
   --log-file gpurun_out/launches_r2ab.csv python bench.py --config 3 --steps 1 --warmup 1 --no-graph --quick > gpurun_out/ncu_bench_r2ab.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r2ab.csv')) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]
adam = [i for i, n in enumerate(names) if 'adam_kernel' in n]
print('adam launches seen', len(adam), 'rows', len(rows))
seg = rows[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for r in seg:
    n = r[4].split('(')[0][-60:]
    agg[n][0] += 1; agg[n][1] += float(r[-1]) / 1000.0
tot = sum(v[1] for v in agg.values())
print(len(seg), 'launches', round(tot, 1), 'us')
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[6:20]:
    print('%-62s %4d %9.1f' % (n, v[0], v[1]))
PY
