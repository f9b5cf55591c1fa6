mkdir -p gpurun_out
OBMAN_RAYCAST_PT=2 timeout 200 python scripts/time_raycast.py 2>&1 | grep "^raycast"
OBMAN_RAYCAST_PT=4 timeout 200 python scripts/time_raycast.py 2>&1 | grep "^raycast"
OBMAN_RAYCAST_PACKED=0 timeout 200 python scripts/time_raycast.py 2>&1 | grep "^raycast"
timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_handnet.py -q -x 2>&1 | tail -2
OBMAN_RAYCAST_PT=4 timeout 600 python -m pytest tests/test_gpu_geometry.py -q -x 2>&1 | tail -2
