"""forces_resilient_planner_b200 -- B200-native batched NMPC solver.

Drop-in for the receding-horizon hot path of ZJU-FAST-Lab/forces_resilient_planner: the
ForcesPro-generated `FORCESNLPsolver_{normal,final}_solve` interior-point solve, rebuilt as a
hand-written sm_100a CUDA kernel behind the reference's own C ABI plus a batched entry point.
See DESIGN.md and INTEGRATION.md.
"""
from . import workloads  # noqa: F401
from ._lib import NmpcOpts, default_opts, load  # noqa: F401
from .solver import DeviceBatch, Result, solve, solve_device, solve_host  # noqa: F401
