set -x
timeout 1200 python -m pytest tests -x -q -m gpu -k "warp_group or mixed_precision_stream or failed_agent" 2>&1 | tail -15
