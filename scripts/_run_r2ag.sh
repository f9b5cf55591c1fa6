mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m pytest tests/test_gpu_nccl.py -q -s 2>&1 | grep -E "gradient:|update rel|passed|failed|FAILED|Error|^E " | head -20
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r2ag.json 2> gpurun_out/bench_2gpu_r2ag.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_2gpu_r2ag.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(d['config']['workload'][:60], round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'sync', d.get('replicas_in_sync'))
print('comm', d.get('comm'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --strong > gpurun_out/bench_2gpu_strong_r2ag.json 2> gpurun_out/bench_2gpu_strong_r2ag.err; echo "strong rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_2gpu_strong_r2ag.json').read().strip().splitlines()[-1])
print('strong', d['config'].get('global_batch'), d['config'].get('per_gpu_batch'), d['scaling'], round(d['ms_per_step'], 3), round(d['value'], 1), 'sync', d.get('replicas_in_sync'))
PY
