set -x
timeout 1200 python -m pytest tests -x -q -m gpu -k "rank or stream or mixed or matches_oracle" 2>&1 | tail -4
timeout 400 python scripts/mixed_probe.py --config 3 2>&1 | tail -1 | cut -c1-900
timeout 300 python scripts/stream_bench.py --replans 300 --mixed 2>&1 | tail -1 | cut -c1-330
timeout 300 python scripts/stream_bench.py --replans 300 2>&1 | tail -1 | cut -c1-330
