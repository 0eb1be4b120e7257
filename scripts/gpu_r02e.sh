set -x
for ag in 128 256; do
  timeout 300 python scripts/stream_bench.py --agents $ag --replans 300 2>&1 | tail -1
  timeout 300 python scripts/stream_bench.py --agents $ag --replans 300 --mixed 2>&1 | tail -1
done
