"""Mirror of `cutfemx.level_set` for the hot path: `normal` and `level_set_value`
(python/cutfemx/level_set.py:146-208, 553-559 -> cpp/cutfemx/level_set/normal.h, value.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import HOST, check, lib
from .cut import CutData, RuntimeQuadratureRules


def _ls_index(cut_data: CutData, level_set) -> int:
    for i, f in enumerate(cut_data.level_sets):
        if f is level_set:
            return i
    raise ValueError("level set is not part of this CutData")


def normal(cut_data: CutData, level_set, rules: RuntimeQuadratureRules, sign: float = 1.0) -> np.ndarray:
    """sign * grad(phi)/|grad(phi)| at the rule points, shape (npts, gdim) float64
    (level_set/normal.h:39-188).  The device copy stays attached to `rules` for the
    Nitsche kernels (the role of the QuadratureFunction the reference hands to runintgen)."""
    cut_data._activate()
    h = cut_data._ctx.handle
    out = np.zeros((rules.total_points, cut_data.gdim))
    check(h, lib().cfx_evaluate_normals(h, _ls_index(cut_data, level_set), rules._h, C.c_double(sign),
                                        C.c_void_p(out.ctypes.data), HOST))
    rules.normal_sign = sign
    return out


def attach_normal(cut_data: CutData, level_set, rules: RuntimeQuadratureRules, sign: float = 1.0) -> None:
    """Evaluate the normals on the device only (no export)."""
    cut_data._activate()
    h = cut_data._ctx.handle
    check(h, lib().cfx_evaluate_normals(h, _ls_index(cut_data, level_set), rules._h, C.c_double(sign), None, HOST))
    rules.normal_sign = sign


def level_set_value(cut_data: CutData, level_set, rules: RuntimeQuadratureRules) -> np.ndarray:
    """phi at the rule points (level_set/value.h:34-119)."""
    cut_data._activate()
    h = cut_data._ctx.handle
    out = np.zeros(rules.total_points)
    check(h, lib().cfx_evaluate_values(h, _ls_index(cut_data, level_set), rules._h, C.c_void_p(out.ctypes.data), HOST))
    return out
