set -x
mkdir -p gpurun_out
PROF_PASSES=2 PROF_WG_PASSES=2 PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|wgrad_bf16_kernel' -o gpurun_out/prof_bf16 -f python scripts/prof_kernels.py > gpurun_out/ncu_bf.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_bf.log
ls -la gpurun_out/*.ncu-rep
