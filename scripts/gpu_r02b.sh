set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solver.py -x -q -m gpu -k "mixed" 2>&1 | tail -15
timeout 300 python scripts/mixed_probe.py --config 2 > gpurun_out/r02b_mixed_c2.json 2> gpurun_out/r02b_mixed_c2.err; tail -c 400 gpurun_out/r02b_mixed_c2.err; cat gpurun_out/r02b_mixed_c2.json
timeout 400 python scripts/mixed_probe.py --config 3 > gpurun_out/r02b_mixed_c3.json 2> gpurun_out/r02b_mixed_c3.err; tail -c 400 gpurun_out/r02b_mixed_c3.err; cat gpurun_out/r02b_mixed_c3.json
