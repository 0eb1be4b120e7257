# Round-2 validation on ONE B200: GPU tests, bench (both arms), ncu launch list + full captures of the three solver kernels,
# compute-sanitizer, stream benches.  Everything lands in gpurun_out/ (copied to profiles/r02_* afterwards).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 800 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_value_e2e.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu_launch2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_kernel -s 4 -c 1 -f -o gpurun_out/r02_fused_fp64 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_fused.log 2>&1
bash scripts/ncu_digest.sh gpurun_out/r02_fused_fp64.ncu-rep gpurun_out/r02_fused_fp64 $(python scripts/phase_ranges.py nmpc_ipm.cuh nmpc_model.cuh)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_mixed -s 1 -c 1 -f -o gpurun_out/r02_fused_mixed python scripts/mixed_probe.py --config 2 --reps 1 > gpurun_out/ncu_mixed.log 2>&1
bash scripts/ncu_digest.sh gpurun_out/r02_fused_mixed.ncu-rep gpurun_out/r02_fused_mixed $(python scripts/phase_ranges.py nmpc_ipm.cuh nmpc_ipm_mixed.cuh nmpc_model.cuh)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_mixed -s 1 -c 1 -f -o gpurun_out/r02_fused_mixed_c3 python scripts/mixed_probe.py --config 3 --batch 16384 --reps 1 > gpurun_out/ncu_mixed3.log 2>&1
bash scripts/ncu_digest.sh gpurun_out/r02_fused_mixed_c3.ncu-rep gpurun_out/r02_fused_mixed_c3 $(python scripts/phase_ranges.py nmpc_ipm.cuh nmpc_ipm_mixed.cuh nmpc_model.cuh)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kkt_backsolve -s 6 -c 1 -f -o gpurun_out/r02_backsolve python tests/tools/bs_check.py > gpurun_out/ncu_bs.log 2>&1
bash scripts/ncu_digest.sh gpurun_out/r02_backsolve.ncu-rep gpurun_out/r02_backsolve
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_san_mem.log 2>&1; tail -3 gpurun_out/r02_san_mem.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_san_race.log 2>&1; tail -3 gpurun_out/r02_san_race.log
timeout 300 python scripts/stream_bench.py > gpurun_out/r02_stream_fp64.json 2>&1; timeout 300 python scripts/stream_bench.py --mixed > gpurun_out/r02_stream_mixed.json 2>&1
timeout 300 python scripts/stream_bench.py --agents 128 > gpurun_out/r02_stream128_fp64.json 2>&1; timeout 300 python scripts/stream_bench.py --agents 128 --mixed > gpurun_out/r02_stream128_group.json 2>&1
timeout 300 python scripts/group_probe.py > gpurun_out/r02_group_probe.json 2>&1
timeout 300 python scripts/aux_bench.py > gpurun_out/r02_aux_kernels.json 2>&1
tail -c 400 gpurun_out/r02_stream128_group.json
