set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_mixed -s 1 -c 1 -f -o gpurun_out/mixed_r02f python scripts/mixed_probe.py --config 2 --reps 1 > gpurun_out/ncu_mixed.log 2>&1; tail -2 gpurun_out/ncu_mixed.log
