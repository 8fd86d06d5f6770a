"""Solve the moment equations of the fully symmetric simplex rules we embed.

Run once (mpmath, 40 digits) to obtain the orbit parameters printed below; the
values are hard-coded in cutfemx_b200/csrc/simplex_rules.cuh and oracle/rules.py
and re-verified against monomial integrals in tests/test_rules.py.
"""
import itertools
import mpmath as mp

mp.mp.dps = 40


def tet_orbits(params):
    a1, w1, a2, w2, b, w3 = params
    pts = []
    for a, w in ((a1, w1), (a2, w2)):
        c = 1 - 3 * a
        for p in set(itertools.permutations((a, a, a, c))):
            pts.append((p, w))
    for p in set(itertools.permutations((b, b, mp.mpf(1) / 2 - b, mp.mpf(1) / 2 - b))):
        pts.append((p, w3))
    return pts


def mono_tet(a, b, c):
    return mp.factorial(a) * mp.factorial(b) * mp.factorial(c) / mp.factorial(a + b + c + 3)


def tet_residual(*params):
    pts = tet_orbits(params)
    res = []
    # symmetric rule: enough to test one representative per partition
    for (a, b, c) in ((0, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (2, 2, 0), (5, 0, 0)):
        s = sum(w * p[0] ** a * p[1] ** b * p[2] ** c for p, w in pts)
        res.append(s - mono_tet(a, b, c))
    return res


sol = mp.findroot(
    tet_residual,
    (0.31088591926330060980, 0.11268792571801585080 / 6, 0.092735250310891226402,
     0.073493043116361949544 / 6, 0.045503704125649649492, 0.042546020777081466438 / 6),
)
print("tet14:", [mp.nstr(s, 25) for s in sol])
# check all monomials up to degree 5
pts = tet_orbits(list(sol))
err = max(
    abs(sum(w * p[0] ** a * p[1] ** b * p[2] ** c for p, w in pts) - mono_tet(a, b, c))
    for a in range(6) for b in range(6 - a) for c in range(6 - a - b)
)
print("tet14 max moment err deg<=5:", mp.nstr(err, 5), "npts", len(pts))


def tri_orbits(params):
    a1, w1, a2, w2 = params
    pts = []
    for a, w in ((a1, w1), (a2, w2)):
        for p in set(itertools.permutations((a, a, 1 - 2 * a))):
            pts.append((p, w))
    return pts


def mono_tri(a, b):
    return mp.factorial(a) * mp.factorial(b) / mp.factorial(a + b + 2)


def tri_residual(*params):
    pts = tri_orbits(params)
    res = []
    for (a, b) in ((0, 0), (2, 0), (3, 0), (4, 0)):
        s = sum(w * p[0] ** a * p[1] ** b for p, w in pts)
        res.append(s - mono_tri(a, b))
    return res


sol = mp.findroot(tri_residual, (0.445948490915965, 0.223381589678011 / 2, 0.091576213509771, 0.109951743655322 / 2))
print("tri6:", [mp.nstr(s, 25) for s in sol])
pts = tri_orbits(list(sol))
err = max(abs(sum(w * p[0] ** a * p[1] ** b for p, w in pts) - mono_tri(a, b)) for a in range(5) for b in range(5 - a))
print("tri6 max moment err deg<=4:", mp.nstr(err, 5))
