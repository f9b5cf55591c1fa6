set -x
mkdir -p gpurun_out
PROBE_ONLY=wgrad timeout 900 python scripts/probe_dense.py > gpurun_out/probe_wg.log 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/probe_dense.json'))
for k,v in d.items():
    if '_b3' in k: print(k, v.get('rel'), v.get('error','')[:300])
PY
PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=20 timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_bf.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/prof_bf.txt
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_handnet.py tests/test_gpu_dense.py -q -m gpu > gpurun_out/bf_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/bf_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --dump-launches gpurun_out/tc_launches_bf.txt > gpurun_out/bench_bf.json 2> gpurun_out/bench_bf.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/bench_bf.json | head -c 900
