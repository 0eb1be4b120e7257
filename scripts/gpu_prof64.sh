set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_kernel -s 4 -c 1 -f -o gpurun_out/r02_fused_fp64 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_fused.log 2>&1
bash scripts/ncu_digest.sh gpurun_out/r02_fused_fp64.ncu-rep gpurun_out/r02_fused_fp64 $(python scripts/phase_ranges.py nmpc_ipm.cuh nmpc_model.cuh)
