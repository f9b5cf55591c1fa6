mkdir -p gpurun_out
PROF_B=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'nn_packed_kernel|raycast_stream_kernel|raycast_prep|pointmlp_l1_fwd|pointmlp_l1_bwd_kernel' -c 5 -o gpurun_out/prof_cuda_core -f python scripts/prof_cuda_core.py > gpurun_out/ncu_cuda_core.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_cuda_core.log; ls -la gpurun_out/prof_cuda_core.ncu-rep
