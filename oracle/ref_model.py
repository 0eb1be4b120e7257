"""ctypes binding to the REFERENCE's own CasADi model callbacks (oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY. The shared objects are compiled by `make -C oracle ref` directly from
/root/reference/src/resilient_planner/plan_manage/solver/{normal,final}/FORCESNLPsolver_*_casadi.c
and *_casadi2forces.c (never copied into this repo).  They expose the callback the ForcesPro
solver core invokes once per stage per iteration:

    FORCESNLPsolver_normal_casadi2forces(x, y, l, p, f, nabla_f, c, nabla_c, h, nabla_h, hess,
                                         stage, iteration, threadID)
    (reference: solver/normal/FORCESNLPsolver_normal_casadi2forces.c:42-245)

Conventions of the reference callback (verified numerically in tests/test_model_parity.py):
  * `*f += stage cost` (accumulates; caller zeroes),
  * outputs are scattered sparse->dense, so the caller must pre-zero them,
  * nabla_c is the 13x17 dynamics Jacobian, nabla_h the 30x17 corridor Jacobian, both COLUMN-major,
  * stage 19 (terminal) skips the dynamics outputs.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NZ, NP, NEQ, NH = 17, 130, 13, 30

_dp = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp)


class RefModel:
    """One of the two reference model flavours: 'normal' or 'final'."""

    def __init__(self, variant: str = "normal"):
        path = os.path.join(_HERE, "_ref", f"libref_model_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` in the build container")
        self.lib = ctypes.CDLL(path)
        self.fn = getattr(self.lib, f"FORCESNLPsolver_{variant}_casadi2forces")
        self.fn.restype = None
        self.fn.argtypes = [_dp] * 11 + [ctypes.c_int] * 3
        self.variant = variant

    def eval(self, z, p, stage: int):
        """Returns dict(f, grad[17], c[13], jc[13,17], h[30], jh[30,17]) at one stage."""
        z = np.ascontiguousarray(z, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        assert z.shape == (NZ,) and p.shape == (NP,)
        f = np.zeros(1)
        g = np.zeros(NZ)
        c = np.zeros(NEQ)
        jc = np.zeros(NEQ * NZ)
        h = np.zeros(NH)
        jh = np.zeros(NH * NZ)
        y = np.zeros(NEQ)
        lam = np.zeros(NH)
        self.fn(_ptr(z), _ptr(y), _ptr(lam), _ptr(p), _ptr(f), _ptr(g), _ptr(c), _ptr(jc),
                _ptr(h), _ptr(jh), None, int(stage), 0, 0)
        return dict(f=float(f[0]), grad=g, c=c,
                    jc=jc.reshape(NZ, NEQ).T.copy(),      # column-major 13x17
                    h=h, jh=jh.reshape(NZ, NH).T.copy())  # column-major 30x17


def available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"libref_model_{v}.so"))
               for v in ("normal", "final"))
