set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 600 gpurun_out/bench_r01b.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01b_ref.json 2>> gpurun_out/bench_r01b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 3 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm -s 2 -c 1 -f -o gpurun_out/fused_r01b python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kkt_backsolve -s 6 -c 1 -f -o gpurun_out/bs_r01b python tests/tools/bs_check.py > gpurun_out/ncu_bs.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/san_mem.log 2>&1; tail -3 gpurun_out/san_mem.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/san_race.log 2>&1; tail -3 gpurun_out/san_race.log
timeout 300 python scripts/stream_bench.py > gpurun_out/stream_r01b.json 2>&1; timeout 300 python scripts/stream_bench.py --ellipsoids > gpurun_out/stream_r01b_ell.json 2>&1
cat gpurun_out/bench_r01b.json | head -c 3000
