mkdir -p gpurun_out
timeout 600 python scripts/trace_kernels.py > gpurun_out/cta_phases.txt 2>&1; echo "rc=$?"
cat gpurun_out/cta_phases.txt
