set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/mixed_probe.py --config 2 > gpurun_out/r02a_mixed_c2.json 2> gpurun_out/r02a_mixed_c2.err; tail -c 400 gpurun_out/r02a_mixed_c2.err; cat gpurun_out/r02a_mixed_c2.json
timeout 400 python scripts/mixed_probe.py --config 3 > gpurun_out/r02a_mixed_c3.json 2> gpurun_out/r02a_mixed_c3.err; tail -c 400 gpurun_out/r02a_mixed_c3.err; cat gpurun_out/r02a_mixed_c3.json
timeout 400 python scripts/mixed_probe.py --config 2 --batch 8192 --horizon 40 > gpurun_out/r02a_mixed_n40.json 2> gpurun_out/r02a_mixed_n40.err; tail -c 400 gpurun_out/r02a_mixed_n40.err; cat gpurun_out/r02a_mixed_n40.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_mixed -s 1 -c 1 -f -o gpurun_out/mixed_r02a python scripts/mixed_probe.py --config 2 --reps 1 > gpurun_out/ncu_mixed.log 2>&1; tail -3 gpurun_out/ncu_mixed.log
