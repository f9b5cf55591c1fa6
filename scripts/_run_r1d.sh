# Round-1 re-entry validation call: GPU tests, headline bench, ncu launch list (+ DRAM traffic) of the same
# bench command, one --set full capture of representative tensor-core kernels and of the Chamfer kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 --dump-launches gpurun_out/tc_launches.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['dtype'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'], d['clocks'], d.get('cpu_baseline'))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json | cut -c 1-300
# launch list of the same bench command (eager launches so that every kernel is its own ncu result)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
# one --set full capture: representative tensor-core kernels + the Chamfer nearest-neighbour kernel
PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=1 timeout 500 ncu --set full --clock-control none --import-source on \
   -k regex:'gemm_tc_kernel|wgrad_bf16_kernel' -c 10 -o gpurun_out/prof_r1d -f python scripts/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'nn_kernel' -c 8 -o gpurun_out/prof_r1d_nn -f \
   python -c "
import torch, sys
sys.path.insert(0, '.')
from obman_train_b200 import functional as Fb
g = torch.Generator(device='cuda').manual_seed(0)
for b, n in ((64, 642), (512, 2500)):
    x = torch.randn(b, n, 3, device='cuda', generator=g) * 60
    y = torch.randn(b, n if n > 1000 else 600, 3, device='cuda', generator=g) * 60
    Fb.nearest_neighbours(x, y); Fb.nearest_neighbours(x, y)
torch.cuda.synchronize()
" > gpurun_out/ncu_nn.log 2>&1; echo "ncu nn rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
tail -c 400 gpurun_out/bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
print(d['dtype'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'])
PY
