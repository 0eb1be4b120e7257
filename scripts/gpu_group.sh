set -x
timeout 600 python scripts/group_probe.py 2>&1 | tail -120
