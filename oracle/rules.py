"""Reference simplex quadrature tables for the ORACLE (test infrastructure only).

TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.  The product (cutfemx_b200/) has its own,
independently written tables in csrc/simplex_rules.h; tests/test_rules.py checks the two
against each other and against exact monomial integrals.

What it stands in for: the sub-simplex rule CutCells maps onto each sub-cell inside
`cutcells::output::quadrature_rules` (called at /root/reference cpp/cutfemx/cut/cut.cpp:1325)
-- CutCells >=0.4,<0.5 is a third-party dependency that is not vendored, so the rule family
is "parity unpinned" (assumption A2): any rule exact to `order`
gives the same integrals for polynomial integrands of that degree.

Convention: points are Cartesian reference coordinates of the unit simplex (vertices
0, e1, e2, e3); weights sum to the reference measure 1/dim!.
"""
from __future__ import annotations

import numpy as np

_TRI6_A1 = 0.4459484909159648863183293
_TRI6_W1 = 0.1116907948390057328475035
_TRI6_A2 = 0.09157621350977074345957146
_TRI6_W2 = 0.05497587182766093381916316

_TET14_A1 = 0.3108859192633006097973457
_TET14_W1 = 0.01878132095300264179986428
_TET14_A2 = 0.09273525031089122640232391
_TET14_W2 = 0.01224884051939365825728503
_TET14_B = 0.04550370412564964949188053
_TET14_W3 = 0.007091003462846911073011571


def _gauss_jacobi_01(n: int, alpha: int):
    """n-point Gauss-Jacobi rule for weight (1-r)^alpha on [0,1]."""
    from scipy.special import roots_jacobi

    t, w = roots_jacobi(n, float(alpha), 0.0)
    return 0.5 * (t + 1.0), w / 2.0 ** (alpha + 1)


def _s21(a):
    c = 1.0 - 2.0 * a
    return [(a, a), (c, a), (a, c)]


def _s31(a):
    c = 1.0 - 3.0 * a
    return [(a, a, a), (c, a, a), (a, c, a), (a, a, c)]


def _s22(b):
    d = 0.5 - b
    return [(b, d, d), (d, b, d), (d, d, b), (d, b, b), (b, d, b), (b, b, d)]


def num_points(dim: int, order: int) -> int:
    return simplex_rule(dim, order)[1].size


def simplex_rule(dim: int, order: int):
    """Return (points[npts, dim], weights[npts]) exact to polynomial degree `order`."""
    if order < 0:
        raise ValueError("order must be >= 0")  # reference: cut.cpp:164-168
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    m = order // 2 + 1
    if dim == 1:
        r, w = _gauss_jacobi_01(m, 0)
        return r.reshape(-1, 1).copy(), w.copy()
    if dim == 2:
        if order <= 1:
            return np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5])
        if order == 2:
            return np.array(_s21(1.0 / 6.0)), np.full(3, 1.0 / 6.0)
        if order <= 4:
            pts = np.array(_s21(_TRI6_A1) + _s21(_TRI6_A2))
            wts = np.array([_TRI6_W1] * 3 + [_TRI6_W2] * 3)
            return pts, wts
        if order == 5:
            s15 = np.sqrt(15.0)
            a, b = (6.0 - s15) / 21.0, (6.0 + s15) / 21.0
            wa, wb = (155.0 - s15) / 2400.0, (155.0 + s15) / 2400.0
            pts = np.array([(1.0 / 3.0, 1.0 / 3.0)] + _s21(a) + _s21(b))
            wts = np.array([0.1125] + [wa] * 3 + [wb] * 3)
            return pts, wts
        r, wr = _gauss_jacobi_01(m, 1)
        s, ws = _gauss_jacobi_01(m, 0)
        pts, wts = [], []
        for i in range(m):
            for j in range(m):
                pts.append((r[i], s[j] * (1.0 - r[i])))
                wts.append(wr[i] * ws[j])
        return np.array(pts), np.array(wts)
    if dim == 3:
        if order <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1.0 / 6.0])
        if order == 2:
            a = (5.0 - np.sqrt(5.0)) / 20.0
            return np.array(_s31(a)), np.full(4, 1.0 / 24.0)
        if order <= 5:
            pts = np.array(_s31(_TET14_A1) + _s31(_TET14_A2) + _s22(_TET14_B))
            wts = np.array([_TET14_W1] * 4 + [_TET14_W2] * 4 + [_TET14_W3] * 6)
            return pts, wts
        r, wr = _gauss_jacobi_01(m, 2)
        s, ws = _gauss_jacobi_01(m, 1)
        t, wt = _gauss_jacobi_01(m, 0)
        pts, wts = [], []
        for i in range(m):
            for j in range(m):
                for k in range(m):
                    pts.append((r[i], s[j] * (1.0 - r[i]), t[k] * (1.0 - r[i]) * (1.0 - s[j])))
                    wts.append(wr[i] * ws[j] * wt[k])
        return np.array(pts), np.array(wts)
    raise ValueError("dim must be 0..3")
