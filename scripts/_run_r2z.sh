mkdir -p gpurun_out
OBMAN_GEMM_PERSIST=2 timeout 300 python scripts/trace_gemm_persist.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_persist_trace.txt
for CL in 2 4; do
OBMAN_GEMM_PERSIST=0 OBMAN_GEMM_CLUSTER=$CL timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick --dump-launches gpurun_out/tc_launches_r2z_cl$CL.txt > gpurun_out/bench_r2z_cl$CL.json 2> gpurun_out/bench_r2z_cl$CL.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_r2z_cl$CL.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print('CLUSTER=$CL', round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'))
PY
grep "gemm M655872\|conv_nhwc n256 32x32 c128->128 taps9\|conv_nhwc n256 16x16 c256->256\|conv_nhwc n256 8x8 c512->512" gpurun_out/tc_launches_r2z_cl$CL.txt | head -12
done
