set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "matches_oracle or mixed or kkt or factor or backsolve or predictor" 2>&1 | tail -8
timeout 300 python scripts/mixed_probe.py --config 2 2>&1 | tail -1
timeout 400 python scripts/mixed_probe.py --config 3 2>&1 | tail -1
