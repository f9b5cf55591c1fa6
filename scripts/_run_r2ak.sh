mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_r2ak.json 2> gpurun_out/bench_4gpu_r2ak.err; echo "bench8 rc=$?"
tail -3 gpurun_out/bench_4gpu_r2ak.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_4gpu_r2ak.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(d['config']['workload'][:60], round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'sync', d.get('replicas_in_sync'), d.get('clocks'))
print('comm', d.get('comm'))
PY
