#!/bin/bash
# Weak-scaling bench at 1/2/4/8 GPUs plus the full-size configs 3 and 4, on one 8xB200 box:
#   gpurun --gpus 8 -- 'bash scripts/scale_run.sh'
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-extras > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  tail -c 300 gpurun_out/scale_$n.json
done
timeout 900 python tests/tools/run_configs.py --config 3 > gpurun_out/config3_full.json 2> gpurun_out/config3.err; tail -c 400 gpurun_out/config3_full.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
  tests/tools/run_configs.py --config 4 > gpurun_out/config4_full_8gpu.json 2> gpurun_out/config4.err; tail -c 400 gpurun_out/config4_full_8gpu.json
