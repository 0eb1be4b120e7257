set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_san_race.log 2>&1; tail -3 gpurun_out/r02_san_race.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_san_mem.log 2>&1; tail -2 gpurun_out/r02_san_mem.log
timeout 300 python scripts/mixed_probe.py --config 2 2>&1 | tail -1 | cut -c1-230
