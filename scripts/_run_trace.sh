mkdir -p gpurun_out
TRACE_ONLY=wgrad timeout 600 python scripts/trace_kernels.py > gpurun_out/cta_phases_wg.txt 2>&1; echo "rc=$?"
grep -v "per-SM" gpurun_out/cta_phases_wg.txt | head -80
