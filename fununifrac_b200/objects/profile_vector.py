"""Profile vector type and the four per-pair primitives.

Mirrors ``src/objects/profile_vector.py`` (:7-23).  In the reference these are
the innermost numpy arithmetic of the all-pairs loop; here the loop lives in
the CUDA kernels (``csrc/pairwise.cu``, ``csrc/diffab.cu``) and these functions
are the N=2 case of the same device path (FP64 validation kernels, so a single
pair matches the reference's float64 arithmetic).
"""
from typing import NewType

import numpy as np

FuncProfileVector = NewType('FuncProfileVector', np.ndarray)


def _pair(P, Q, is_l2: bool, diffab: bool):
    from fununifrac_b200 import engine

    return engine.pair_primitive(np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64), is_l2, diffab)


def get_L1(P: FuncProfileVector, Q: FuncProfileVector):
    return _pair(P, Q, False, False)


def get_L1_diffab(P: FuncProfileVector, Q: FuncProfileVector):
    return _pair(P, Q, False, True)


def get_L2(P: FuncProfileVector, Q: FuncProfileVector):
    return _pair(P, Q, True, False)


def get_L2_diffab(P: FuncProfileVector, Q: FuncProfileVector):
    return _pair(P, Q, True, True)
