mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "gradient:|update rel|full size|passed|failed|FAILED|Error|^E " | head -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_r2u.json 2> gpurun_out/bench_2gpu_r2u.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_2gpu_r2u.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(d['config']['workload'][:60], round(d['ms_per_step'], 3), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'sync', d.get('replicas_in_sync'))
print('comm', d.get('comm'))
PY
AB_B=256 AB_REPS=5 timeout 100 python scripts/bench_chamfer.py 2>&1 | tail -3
