set -x
mkdir -p gpurun_out
cat > /tmp/gp.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import _lib, solver as S, workloads as W
db = S.DeviceBatch(W.config2(128), np.float64, "cuda:0")
for _ in range(3):
    S.solve_device(db, _lib.default_opts(), lowlatency=True); torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_ipm_group -s 1 -c 1 -f -o gpurun_out/group_r02 python /tmp/gp.py > gpurun_out/ncu_group.log 2>&1; tail -2 gpurun_out/ncu_group.log
bash scripts/ncu_digest.sh gpurun_out/group_r02.ncu-rep gpurun_out/r02_group $(python scripts/phase_ranges.py nmpc_ipm_group.cuh nmpc_ipm.cuh nmpc_model.cuh)
