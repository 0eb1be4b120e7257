set -x
timeout 1200 python -m pytest tests -x -q -m gpu -k "warp_group or mixed_precision_stream or failed_agent or receding or pipeline" 2>&1 | tail -15
timeout 600 python scripts/group_probe.py 2>&1 | grep -E "B[0-9]+_|\"ms\"|us_per|same_it|max_dz" | paste - - - - - | cut -c1-260
for ag in 128 1024; do
  timeout 300 python scripts/stream_bench.py --agents $ag --replans 300 2>&1 | tail -1 | cut -c1-330
  timeout 300 python scripts/stream_bench.py --agents $ag --replans 300 --mixed 2>&1 | tail -1| cut -c1-330
done
