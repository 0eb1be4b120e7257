set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 1500 gpurun_out/r02c_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/r02c_bench.err
head -c 6000 gpurun_out/r02c_bench.json; echo; cat gpurun_out/r02c_bench_ref.json
