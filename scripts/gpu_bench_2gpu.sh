set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 2500 gpurun_out/r02_bench_2gpu.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu.json') if l.startswith('{')][-1])
for k in ("value","e2e","collate","config4","config5"): print(k, json.dumps(d.get(k))[:1600])
PY
