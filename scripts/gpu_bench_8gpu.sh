set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 1500 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_8gpu.json') if l.startswith('{')][-1])
for k in ("value","e2e","collate","config4","config5"): print(k, json.dumps(d.get(k))[:1600])
PY
