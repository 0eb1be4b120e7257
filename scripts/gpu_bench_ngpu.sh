# usage: bash scripts/gpu_bench_ngpu.sh N   (under `gpurun --gpus N`): the driver's multi-GPU launch of bench.py
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -c 1200 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1])
for k in ("value","e2e","collate","config4","config5"): print(k, json.dumps(d.get(k))[:1500])
PY
