mkdir -p gpurun_out
PROBE_ONLY=wgrad timeout 900 python scripts/probe_dense.py > gpurun_out/probe_b3.log 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/probe_dense.json'))
for k,v in d.items():
    if '_b3' in k: print(k, v.get('rel'), v.get('error','')[:300])
PY
PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=20 timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_bf.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/prof_bf.txt
TRACE_ONLY=wgrad timeout 600 python scripts/trace_kernels.py > gpurun_out/cta_phases_wg.txt 2>&1; echo "trace rc=$?"
grep -v "per-SM" gpurun_out/cta_phases_wg.txt | head -60
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -q -m gpu -x > gpurun_out/bf_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/bf_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_bf.json 2> gpurun_out/bench_bf.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_bf.json').read().strip().splitlines()[-1]); print(d['ms_per_step'],d['value'],d['e2e']['value'], d['roofline']['gemm_ms_per_step'])"
