"""-m "not gpu": pin the CPU oracle (oracle/) to what the reference's OWN tests assert for this path.

The reference keeps no per-point golden vectors for the cut path ("pointwise
parity unpinned"); its tests assert sets and integrals.  Each test below restates one of those
assertions (P1..P12 of SURVEY.md section 8c, reference file:line in the docstring) on the same
geometry and runs it through the oracle.  Independent analytic checks (monomial exactness of the
sub-simplex rules, clipped-simplex volumes, sphere/circle convergence) pin the third-party
arithmetic (CutCells / Basix) that is absent from /root/reference.
"""
import math

import numpy as np
import pytest

import oracle as O
from cutfemx_b200 import mesh as M
from oracle import rules as R


def _line_problem(n=3, c=0.51):
    """python/tests/test_cut_api.py:19-33 `_line_level_set`: 3x3 unit square, phi = x - 0.51."""
    mesh = M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0))
    V = M.functionspace(mesh, 1)
    phi = M.interpolate(V, lambda x, y, z: x - c)
    return mesh, V, phi


def _circle_problem(n=21, radius=0.5):
    """test_cut_api.py:1268-1279: [-1,1]^2, 21x21 triangles, phi = |x| - R."""
    mesh = M.create_rectangle(n, n, (-1.0, -1.0), (1.0, 1.0))
    V = M.functionspace(mesh, 1)
    phi = M.interpolate(V, M.sphere_level_set((0.0, 0.0, 0.0), radius))
    return mesh, V, phi


def _scalar(V, rules=None, cells=None):
    m = np.zeros(1)
    O.assemble_cells(V, "one", m, cells, rules, (1.0,))
    return m[0]


# ------------------------------------------------------------------------------------------ P1..P5
def test_p1_line_cut_cell_count():
    """test_cut_api.py:95-103 / test_locate_entities.py:13-35: the x = 0.51 line cuts 6 of the 18
    triangles (ids [2,4,7,10,13,15] in DOLFINx numbering; the numbering is 3P, the count is not)."""
    mesh, V, phi = _line_problem()
    dom = O.classify(V.dofmap, phi)
    cut = O.locate(dom, "phi=0")
    assert cut.size == 6
    # they are exactly the triangles of the middle column of quads
    xc = mesh.x[mesh.x_dofmap[cut]][:, :, 0]
    assert np.all(xc.min(axis=1) >= 1 / 3 - 1e-14) and np.all(xc.max(axis=1) <= 2 / 3 + 1e-14)
    assert np.array_equal(cut, np.sort(cut)) and cut.dtype == np.int32


def test_p3_zero_dofs_are_interface():
    """test_cut_api.py:191-208: phi = x - 0.5 on a 2x1 mesh marks ALL cells intersected."""
    mesh = M.create_rectangle(2, 1, (0.0, 0.0), (1.0, 1.0))
    V = M.functionspace(mesh, 1)
    phi = M.interpolate(V, lambda x, y, z: x - 0.5)
    dom = O.classify(V.dofmap, phi)
    assert np.array_equal(O.locate(dom, "phi=0"), np.arange(mesh.num_cells, dtype=np.int32))
    assert O.locate(dom, "phi<0").size == 0 and O.locate(dom, "phi>0").size == 0


def test_p4_le_is_union_and_rules_equal():
    """test_cut_api.py:142-157 (phi<=0 == phi<0 U phi=0) and :702-710 (<= rules == < rules)."""
    mesh, V, phi = _circle_problem(16)
    dom = O.classify(V.dofmap, phi)
    lt, eq, le = O.locate(dom, "phi<0"), O.locate(dom, "phi=0"), O.locate(dom, "phi<=0")
    assert np.array_equal(le, np.union1d(lt, eq))
    ge, gt = O.locate(dom, "phi>=0"), O.locate(dom, "phi>0")
    assert np.array_equal(ge, np.union1d(gt, eq))
    assert np.array_equal(np.sort(np.concatenate([lt, eq, gt])), np.arange(mesh.num_cells))
    a = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<", 3)
    b = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<=", 3)
    for name in ("points", "weights", "offsets", "parent_map"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_selector_grammar():
    """cut.cpp:47-57 (spaces ignored), docs/user-guide/element-classification.md:149-152 (and/or)."""
    mesh, V, phi = _circle_problem(12)
    phi1 = M.interpolate(V, lambda x, y, z: x - 0.1)
    dom = np.stack([O.classify(V.dofmap, phi), O.classify(V.dofmap, phi1)])
    names = ("phi", "phi1")
    a = O.locate(dom, "phi<0 and phi1<0", names)
    b = O.locate(dom, " phi < 0and  phi1<0 ", names)
    assert np.array_equal(a, b)
    assert np.array_equal(a, np.intersect1d(O.locate(dom, "phi<0", names), O.locate(dom, "phi1<0", names)))
    c = O.locate(dom, "phi<0 or phi1<0", names)
    assert np.array_equal(c, np.union1d(O.locate(dom, "phi<0", names), O.locate(dom, "phi1<0", names)))
    with pytest.raises(ValueError):
        O.locate(dom, "psi<0", names)


@pytest.mark.parametrize("kind", ["<", ">", "="])
def test_p5_rule_container_invariants(kind):
    """test_cut_api.py:405-421."""
    mesh, V, phi = _circle_problem(14)
    dom = O.classify(V.dofmap, phi)
    r = O.runtime_quadrature(mesh, V.dofmap, phi, dom, kind, 4)
    assert r.offsets.dtype == np.int32 and r.parent_map.dtype == np.int32
    assert r.offsets[0] == 0 and r.offsets[-1] == r.weights.size
    assert r.parent_map.size == r.offsets.size - 1
    assert r.points.shape == (r.weights.size, mesh.tdim)
    assert np.all(np.diff(r.offsets) > 0)
    assert np.all(np.isin(r.parent_map, O.locate(dom, "phi=0")))
    assert np.all(np.diff(r.parent_map) > 0)  # assumption A1: ascending parent cells, one rule per cell
    # points are reference coordinates of the parent cell
    assert r.points.min() >= -1e-14 and r.points.sum(axis=1).max() <= 1 + 1e-14


# ------------------------------------------------------------------------------------------ P6, P7, P12
def test_p6_circle_area_and_perimeter():
    """test_cut_api.py:1268-1300: |area - pi R^2| < 1e-2, |perimeter - 2 pi R| < 1e-2."""
    mesh, V, phi = _circle_problem()
    dom = O.classify(V.dofmap, phi)
    inside = O.locate(dom, "phi<0")
    rv = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<", 4)
    ri = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "=", 4)
    area = _scalar(V, rv, inside)
    perimeter = _scalar(V, ri)
    assert abs(area - math.pi * 0.25) < 1.0e-2
    assert abs(perimeter - math.pi) < 1.0e-2
    # sharper: both equal the exact measures of the piecewise-linear interface polygon
    assert abs(rv.weights.sum() + _scalar(V, None, inside) - area) < 1e-13


def test_p7_vector_sums_to_area():
    """test_cut_api.py:858-869: sum(b) == area to 1e-12 for L = 1*v (partition of unity)."""
    for deg in (1, 2):
        mesh, Vphi, phi = _circle_problem(18)
        V = M.functionspace(mesh, deg, permute_seed=3)
        dom = O.classify(Vphi.dofmap, phi)
        inside = O.locate(dom, "phi<0")
        rv = O.runtime_quadrature(mesh, Vphi.dofmap, phi, dom, "<", 4)
        b = np.zeros(V.num_dofs)
        O.assemble_cells(V, "source", b, inside, rv, (1.0,))
        area = _scalar(V, rv, inside)
        np.testing.assert_allclose(b.sum(), area, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("tdim", [2, 3])
def test_p12_opposite_fractions_sum_to_one(tdim):
    """test_extensions_cell_aggregation.py:74-98: volume fractions of phi<0 and phi>0 sum to 1 on
    every cut cell (1e-12)."""
    if tdim == 2:
        mesh, V, phi = _circle_problem(17)
    else:
        mesh = M.create_box(7, 6, 5)
        V = M.functionspace(mesh, 1)
        phi = M.interpolate(V, M.sphere_level_set((0.5, 0.5, 0.5), 0.35))
    dom = O.classify(V.dofmap, phi)
    neg = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<", 2)
    pos = O.runtime_quadrature(mesh, V.dofmap, phi, dom, ">", 2)
    X = mesh.x[mesh.x_dofmap][:, :, :tdim]
    J = (X[:, 1:] - X[:, :1]).transpose(0, 2, 1)
    vol = np.abs(np.linalg.det(J)) / math.factorial(tdim)
    frac = np.zeros(mesh.num_cells)
    np.add.at(frac, neg.parent_map, np.add.reduceat(neg.weights, neg.offsets[:-1]))
    np.add.at(frac, pos.parent_map, np.add.reduceat(pos.weights, pos.offsets[:-1]))
    cut = O.locate(dom, "phi=0")
    np.testing.assert_allclose(frac[cut] / vol[cut], 1.0, atol=1e-12)


# ------------------------------------------------------------------------------------------ P8
@pytest.mark.parametrize("tdim,deg", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_p8_full_cell_runtime_rule_equals_standard_assembly(tdim, deg):
    """test_assembly_poisson.py (P1 Poisson, 1e-12) / test_assembly_elasticity.py (1e-9): a run-time
    rule that covers whole cells (reference points = standard rule, weights = w*|detJ|,
    quadrature_utils.py:53-61) assembles the same matrix/vector as the standard integral."""
    mesh = M.create_rectangle(5, 4) if tdim == 2 else M.create_box(3, 3, 2)
    V = M.functionspace(mesh, deg, permute_seed=5)
    cells = np.arange(mesh.num_cells, dtype=np.int32)
    X = mesh.x[mesh.x_dofmap][:, :, :tdim]
    detJ = np.abs(np.linalg.det((X[:, 1:] - X[:, :1]).transpose(0, 2, 1)))
    row_ptr, cols = O.sparsity(V, cells)
    for kernel, order in (("laplace", 2 * (deg - 1)), ("mass", 2 * deg)):
        p, w = R.simplex_rule(tdim, order)
        nq = w.size
        rules = O.Rules(tdim, np.tile(p, (cells.size, 1)), (detJ[:, None] * w[None, :]).reshape(-1),
                        (np.arange(cells.size + 1) * nq).astype(np.int32), cells.copy())
        a_std = O.assemble_cells(V, kernel, np.zeros(cols.size), cells, None, (1.3,), row_ptr, cols)
        a_rt = O.assemble_cells(V, kernel, np.zeros(cols.size), None, rules, (1.3,), row_ptr, cols)
        assert np.linalg.norm(a_std - a_rt) <= 1e-12 * np.linalg.norm(a_std)
    p, w = R.simplex_rule(tdim, deg)
    rules = O.Rules(tdim, np.tile(p, (cells.size, 1)), (detJ[:, None] * w[None, :]).reshape(-1),
                    (np.arange(cells.size + 1) * w.size).astype(np.int32), cells.copy())
    b_std = O.assemble_cells(V, "source", np.zeros(V.num_dofs), cells, None, (0.7,))
    b_rt = O.assemble_cells(V, "source", np.zeros(V.num_dofs), None, rules, (0.7,))
    assert np.linalg.norm(b_std - b_rt) <= 1e-12 * np.linalg.norm(b_std)
    # and the Laplace matrix has zero row sums, the mass matrix sums to the domain measure
    lap = O.assemble_cells(V, "laplace", np.zeros(cols.size), cells, None, (1.0,), row_ptr, cols)
    mass = O.assemble_cells(V, "mass", np.zeros(cols.size), cells, None, (1.0,), row_ptr, cols)
    rows = np.repeat(np.arange(V.num_dofs), np.diff(row_ptr))
    assert np.abs(np.bincount(rows, lap, V.num_dofs)).max() < 1e-11
    size = np.prod(np.asarray(mesh.p1) - np.asarray(mesh.p0))
    np.testing.assert_allclose(mass.sum(), size, rtol=1e-12)


# ------------------------------------------------------------------------------------------ P9
def test_p9_line_interface_normal_integrals():
    """test_cut_api.py:989-1009: int |n|^2 = int n_x = measure on the x = 0.51 line (1e-12)."""
    mesh, V, phi = _line_problem()
    dom = O.classify(V.dofmap, phi)
    ri = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "=", 2)
    n = O.normals(mesh, V.dofmap, 1, phi, ri)
    measure = ri.weights.sum()
    np.testing.assert_allclose(measure, 1.0, rtol=1e-12)  # the line x = 0.51 crosses the unit square once
    np.testing.assert_allclose((ri.weights * (n * n).sum(axis=1)).sum(), measure, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose((ri.weights * n[:, 0]).sum(), measure, rtol=1e-12, atol=1e-12)
    assert n.dtype == np.float64 and n.shape == (ri.weights.size, 2)


def test_p9_p2_level_set_normal_is_radial():
    """test_cut_api.py:1012-1026: for a P2 level set phi = |x-c|^2 - r^2 (exactly representable) the
    evaluated normal equals the radial normal at every run-time point (error integral < 1e-24)."""
    center, radius = (0.47, 0.43), 0.31
    mesh = M.create_rectangle(6, 6, (0.0, 0.0), (1.0, 1.0))
    V1 = M.functionspace(mesh, 1)
    V2 = M.functionspace(mesh, 2, permute_seed=9)
    f = lambda x, y, z: (x - center[0]) ** 2 + (y - center[1]) ** 2 - radius ** 2  # noqa: E731
    phi1, phi2 = M.interpolate(V1, f), M.interpolate(V2, f)
    dom = O.classify(V1.dofmap, phi1)
    ri = O.runtime_quadrature(mesh, V1.dofmap, phi1, dom, "=", 5)
    n = O.normals(mesh, V2.dofmap, 2, phi2, ri)
    xq = O.physical_points(mesh, ri)
    d = xq[:2].T - np.asarray(center)
    n_exact = d / np.linalg.norm(d, axis=1)[:, None]
    err = (ri.weights * ((n - n_exact) ** 2).sum(axis=1)).sum()
    assert err < 1e-24


# ------------------------------------------------------------------------------------------ P10, P11
def test_p10_active_cells():
    """test_cut_api.py:841-845: active cells = unique(inside U parent_map)."""
    mesh, V, phi = _circle_problem(15)
    dom = O.classify(V.dofmap, phi)
    inside = O.locate(dom, "phi<0")
    rv = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<", 2)
    active = np.unique(np.concatenate([inside, rv.parent_map]))
    row_ptr, cols = O.sparsity(V, active, None, insert_diagonal=False)
    touched = np.unique(V.dofmap[active])
    assert np.array_equal(np.nonzero(np.diff(row_ptr))[0], touched)


def test_p11_ghost_facets_owned_unique_two_celled():
    """test_cut_api.py:1158-1173 and :1176-1196."""
    mesh, V, phi = _line_problem()
    dom = O.classify(V.dofmap, phi)
    cut, inside = O.locate(dom, "phi=0"), O.locate(dom, "phi<0")
    facets = O.ghost_penalty_facets(mesh, cut, inside)
    assert facets.size > 0
    assert np.array_equal(facets, np.unique(facets))
    assert np.all(facets < mesh.num_owned_facets)
    ncell = np.diff(mesh.f2c_offsets)
    assert np.all(ncell[facets] == 2)
    # every facet touches a cut cell and both of its cells are in cut U inside (cut.py:364-379)
    band = np.union1d(cut, inside)
    for f in facets:
        cs = mesh.f2c[mesh.f2c_offsets[f]:mesh.f2c_offsets[f + 1]]
        assert np.all(np.isin(cs, band)) and np.any(np.isin(cs, cut))
    # interior_facets_for_cells: stays inside the cell set; all cells -> all interior facets
    msh = M.create_rectangle(4, 4, (0.0, 0.0), (1.0, 1.0))
    sel = np.arange(0, msh.num_cells, 2, dtype=np.int32)
    fs = O.interior_facets_for_cells(msh, sel)
    for f in fs:
        cs = msh.f2c[msh.f2c_offsets[f]:msh.f2c_offsets[f + 1]]
        assert cs.size == 2 and np.all(np.isin(cs, sel))
    allf = O.interior_facets_for_cells(msh, np.arange(msh.num_cells, dtype=np.int32))
    assert np.array_equal(allf, np.nonzero(np.diff(msh.f2c_offsets) == 2)[0])
    rows = O.facet_rows(msh, allf)
    for (c0, l0, c1, l1), f in zip(rows, allf):
        assert msh.c2f[c0, l0] == f and msh.c2f[c1, l1] == f and c0 < c1


# ------------------------------------------------------------------------------------------ analytic pins
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5, 6, 8])
def test_simplex_rules_integrate_monomials_exactly(dim, order):
    """int x^a y^b z^c over the unit simplex = a! b! c! / (a+b+c+dim)!  for a+b+c <= order."""
    p, w = R.simplex_rule(dim, order)
    p = np.asarray(p).reshape(w.size, dim)
    assert np.all(w > 0)
    for a in range(order + 1):
        for b in range(order + 1 - a if dim > 1 else 1):
            for c in range(order + 1 - a - b if dim > 2 else 1):
                e = (a, b, c)[:dim]
                exact = math.prod(math.factorial(k) for k in e) / math.factorial(sum(e) + dim)
                val = (w * np.prod(p ** np.asarray(e), axis=1)).sum()
                assert abs(val - exact) < 2e-15, (e, val, exact)


def test_minimal_point_counts():
    """SURVEY.md section 8 sizing assumes 6-pt triangle / 14-pt-or-fewer tet rules at degree 4."""
    assert R.simplex_rule(2, 4)[1].size == 6 and R.simplex_rule(2, 2)[1].size == 3
    assert R.simplex_rule(3, 2)[1].size == 4 and R.simplex_rule(3, 4)[1].size <= 14


@pytest.mark.parametrize("tdim", [2, 3])
def test_clipped_reference_simplex_measures(tdim):
    """One cell cut by the plane x_0 = t: the inside part {x_0 < t} of the unit simplex has measure
    (1 - (1-t)^d)/d!, the interface {x_0 = t} has measure (1-t)^(d-1)/(d-1)!  -- exact for the
    marching-simplex case tables whatever the sub-simplex ordering is."""
    nv = tdim + 1
    x = np.zeros((nv, 3))
    for k in range(tdim):
        x[k + 1, k] = 1.0
    xd = np.arange(nv, dtype=np.int32)[None, :]
    mesh = M.Mesh(nv, tdim, tdim, x, xd, None, None, None, 1, 0, 0)
    for t in (0.25, 0.5, 0.9):
        phi = x[:, 0] - t
        dom = O.classify(xd, phi)
        assert dom[0] == O.INTERSECTED
        lt = O.runtime_quadrature(mesh, xd, phi, dom, "<", 3)
        gt = O.runtime_quadrature(mesh, xd, phi, dom, ">", 3)
        eq = O.runtime_quadrature(mesh, xd, phi, dom, "=", 3)
        d = tdim
        np.testing.assert_allclose(gt.weights.sum(), (1 - t) ** d / math.factorial(d), rtol=1e-13)
        np.testing.assert_allclose(lt.weights.sum(), (1 - (1 - t) ** d) / math.factorial(d), rtol=1e-13)
        np.testing.assert_allclose(eq.weights.sum(), (1 - t) ** (d - 1) / math.factorial(d - 1), rtol=1e-13)
        # first moments too (order-3 rule): int x_0 over {x_0 > t}
        m1 = (gt.weights * gt.points[:, 0]).sum()
        s = 1 - t
        exact = (s ** d / d - s ** (d + 1) / (d + 1)) / math.factorial(d - 1)
        np.testing.assert_allclose(m1, exact, rtol=1e-13)


def test_sphere_volume_and_area_converge_second_order():
    errs = []
    for n in (8, 16):
        mesh = M.create_box(n, n, n)
        V = M.functionspace(mesh, 1)
        phi = M.interpolate(V, M.sphere_level_set((0.5, 0.5, 0.5), 0.35))
        dom = O.classify(V.dofmap, phi)
        inside = O.locate(dom, "phi<0")
        rv = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "<", 2)
        ri = O.runtime_quadrature(mesh, V.dofmap, phi, dom, "=", 2)
        vol = _scalar(V, rv, inside)
        area = ri.weights.sum()
        errs.append((abs(vol - 4 / 3 * math.pi * 0.35 ** 3), abs(area - 4 * math.pi * 0.35 ** 2)))
    assert errs[1][0] < errs[0][0] / 3 and errs[1][1] < errs[0][1] / 3
    assert errs[1][0] < 4e-3 and errs[1][1] < 4e-2


def test_cutcell_sizing_matches_survey():
    """SURVEY.md section 8 sizing table, config C1: 64x64 circle -> 1480 inside, 226 cut cells."""
    mesh, V, phi = _circle_problem(64)
    dom = O.classify(V.dofmap, phi)
    assert O.locate(dom, "phi<0").size == 1480 and O.locate(dom, "phi=0").size == 226


# ------------------------------------------------------------------------------------------ vector spaces
@pytest.mark.parametrize("tdim,deg", [(2, 1), (3, 1), (2, 2), (3, 2)])
def test_elasticity_oracle_pins(tdim, deg):
    """Blocked (vector) spaces + inner(sigma(u), eps(v)) (demo_elasticity.py:213-224).  Reference pin:
    test_assembly_elasticity.py:18-70 -- a run-time rule covering whole cells assembles the standard matrix
    (1e-9 there, 1e-12 relative here).  Analytic pins: symmetric; translations and infinitesimal rotations
    are in the kernel; the load vector of a constant force sums to force * volume per component."""
    import scipy.sparse as sp

    mesh = M.create_rectangle(4, 4, (0.0, 0.0), (1.0, 1.0)) if tdim == 2 else M.create_box(3, 2, 2)
    bs = tdim
    V = M.functionspace(mesh, deg, bs=bs, permute_seed=3)
    cells = np.arange(mesh.num_cells, dtype=np.int32)
    rp, cols = O.sparsity(V, cells)
    mu, lam = 3.0, 5.0
    a = O.assemble_cells(V, "elasticity", np.zeros(cols.size * bs * bs), cells, None, (mu, lam), rp, cols)
    A = sp.bsr_matrix((a.reshape(-1, bs, bs), cols, rp), shape=(V.num_dofs * bs,) * 2).tocsr()
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    X = V.dof_coords
    modes = []
    for c in range(bs):
        t = np.zeros((V.num_dofs, bs))
        t[:, c] = 1.0
        modes.append(t.ravel())
    for i, j in ([(0, 1)] if bs == 2 else [(0, 1), (0, 2), (1, 2)]):
        r = np.zeros((V.num_dofs, bs))
        r[:, i], r[:, j] = -X[:, j], X[:, i]
        modes.append(r.ravel())
    for m in modes:
        assert np.abs(A @ m).max() < 1e-10 * abs(A).max()
    Xc = mesh.x[mesh.x_dofmap][:, :, :tdim]
    detJ = np.abs(np.linalg.det((Xc[:, 1:] - Xc[:, :1]).transpose(0, 2, 1)))
    p, w = R.simplex_rule(tdim, 2 * (deg - 1))
    p = np.asarray(p).reshape(w.size, tdim)
    rules = O.Rules(tdim, np.tile(p, (cells.size, 1)), (detJ[:, None] * w[None, :]).reshape(-1),
                    (np.arange(cells.size + 1) * w.size).astype(np.int32), cells.copy())
    a2 = O.assemble_cells(V, "elasticity", np.zeros(cols.size * bs * bs), None, rules, (mu, lam), rp, cols)
    assert np.linalg.norm(a - a2) < 1e-12 * np.linalg.norm(a)
    force = (1.0, 2.0, 3.0)
    b = O.assemble_cells(V, "source_vec", np.zeros(V.num_dofs * bs), cells, None, force)
    vol = np.prod(np.asarray(mesh.p1) - np.asarray(mesh.p0))
    np.testing.assert_allclose(b.reshape(-1, bs).sum(axis=0), np.asarray(force[:bs]) * vol, rtol=1e-12)


# ---------------------------------------------------------------- Dirichlet conditions
def test_dirichlet_restatement_is_consistent_with_the_unconstrained_matrix():
    """assemble_matrix(bcs) zeroes rows/columns of every element tensor (assemble_matrix_impl.h:146-185) and
    lift_bc subtracts alpha * Ae[:, bc] (x_bc - x0) (assemble_vector_impl.h:405-432).  Both are linear in the
    element tensors, so against the unconstrained oracle matrix A: masked == A with rows/cols zeroed, and the
    lifting == alpha * A[:, bc] (g - x0)[bc].  Blocked and scalar spaces."""
    import scipy.sparse as sp

    from cutfemx_b200 import mesh as M
    from util import make_problem

    for kind, n, deg, bs in (("circle", 8, 1, 1), ("circle", 6, 2, 2), ("sphere", 4, 1, 3)):
        mesh, Vphi, phi, _ = make_problem(kind, n, 1)
        V = M.functionspace(mesh, deg, bs=bs)
        dom = O.classify(Vphi.dofmap, phi.x.array)
        inside, cut = O.locate(dom, "phi<0"), O.locate(dom, "phi=0")
        rv = O.runtime_quadrature(mesh, Vphi.dofmap, phi.x.array, dom, "<", 2)
        rows4 = O.facet_rows(mesh, O.ghost_penalty_facets(mesh, cut, inside))
        rp, cols = O.sparsity(V, np.concatenate([inside, rv.parent_map]), rows4)
        kern, cc = ("laplace", (1.0,)) if bs == 1 else ("elasticity", (1.0, 2.0))

        def assemble(out):
            O.assemble_cells(V, kern, out, inside, rv, cc, rp, cols)
            O.assemble_interior_facets(V, "ghost_grad_jump", out, rows4, (0.1,), rp, cols)
            return out

        nb = V.num_dofs * bs
        rng = np.random.default_rng(1)
        markers = (rng.random(nb) < 0.2).astype(np.int8)
        g, x0, alpha = rng.standard_normal(nb), rng.standard_normal(nb), 0.6
        full = assemble(np.zeros(cols.size * bs * bs))
        to_csr = lambda v: (sp.bsr_matrix((v.reshape(-1, bs, bs), cols, rp), shape=(nb, nb)).tocsr() if bs > 1
                            else sp.csr_matrix((v, cols, rp), shape=(nb, nb)))
        Af = to_csr(full)
        with O.dirichlet("matrix", markers, markers):
            masked = assemble(np.zeros_like(full))
        keep = sp.diags((1 - markers).astype(float))
        assert abs(to_csr(masked) - keep @ Af @ keep).max() <= 1e-13 * abs(Af).max()
        b = np.zeros(nb)
        with O.dirichlet("lifting", None, markers, g, x0, alpha, b):
            assemble(np.zeros_like(full))
        expect = -alpha * (Af @ (markers * (g - x0)))
        assert np.linalg.norm(b - expect) <= 1e-12 * np.linalg.norm(expect)
        # set_diagonal / set_bc
        d = np.nonzero(markers)[0]
        O.set_diagonal(rp, cols, masked, d, 2.5, bs)
        assert np.all(to_csr(masked).diagonal()[d] == 2.5)
        O.set_bc(b, d, g, x0, alpha)
        np.testing.assert_allclose(b[d], alpha * (g - x0)[d])


# ---------------------------------------------------------------- Function coefficients
def test_square_functional_of_a_packed_coefficient():
    """demo_poisson.py:213 error functional: int (w_h)^2 over the cut domain with w_h the P1 / P2 interpolant of a
    linear function is exact with the degree-2p rules (standard and run-time), so it equals int (a + b.x)^2 over
    the polygon/polyhedron the cut produces -- computed here from the oracle's own order-2 moments."""
    from cutfemx_b200 import mesh as M
    from util import make_problem

    for kind, n, deg in (("circle", 9, 1), ("circle", 6, 2), ("sphere", 4, 1)):
        mesh, Vphi, phi, _ = make_problem(kind, n, 1)
        V = M.functionspace(mesh, deg)
        dom = O.classify(Vphi.dofmap, phi.x.array)
        inside = O.locate(dom, "phi<0")
        rv = O.runtime_quadrature(mesh, Vphi.dofmap, phi.x.array, dom, "<", 2 * deg)
        a0, b = 0.3, np.array([1.1, -0.7, 0.4])[: mesh.tdim]
        w = a0 + V.dof_coords[:, : mesh.tdim] @ b
        val = np.zeros(1)
        with O.coefficient(w):
            O.assemble_cells(V, "square_fn", val, inside, rv, (2.0,))
        # the same integral from physical quadrature points of a degree-2 rule on every entity
        ref = 0.0
        X = mesh.x[mesh.x_dofmap][:, :, : mesh.tdim]
        from oracle import rules as R

        p2, w2 = R.simplex_rule(mesh.tdim, 2)
        p2 = np.asarray(p2).reshape(w2.size, mesh.tdim)
        for c in inside:
            J = (X[c, 1:] - X[c, :1]).T
            xq = X[c, 0] + p2 @ J.T
            ref += 2.0 * np.sum(w2 * abs(np.linalg.det(J)) * (a0 + xq @ b) ** 2)
        xp = O.physical_points(mesh, rv).T
        ref += 2.0 * np.sum(rv.weights * (a0 + xp @ b) ** 2)
        assert abs(val[0] - ref) <= 1e-12 * abs(ref)


# ---------------------------------------------------------------- facets as hosts
def test_facet_hosted_rules_measure_the_wet_boundary_exactly():
    """Planar level sets cut the boundary of the unit square / cube in straight pieces, so the wet part of the boundary
    (full inside facets + the run-time rules of the cut facets) has a closed form: 2.02 for phi = x - 0.51 on the
    square, 1 + 0.52 + 0.22 + 2 * 0.37 = 2.48 for phi = x + 0.3 y - 0.52 on the cube.  Also the invariants the
    reference's tests assert (test_cut_api.py:171-188, :424-462)."""
    from cutfemx_b200 import mesh as M

    for tdim, n, fn, exact in ((2, 9, lambda x, y, z: x - 0.51, 2.02), (3, 5, lambda x, y, z: x + 0.3 * y - 0.52, 2.48)):
        mesh = M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0)) if tdim == 2 else M.create_box(n, n, n)
        V = M.functionspace(mesh, 1, permute_seed=9)
        phi = M.Function(V, "phi").interpolate(fn)
        facets = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0].astype(np.int32)
        code, verts, _ = O.classify_facets(mesh, V.dofmap, phi.x.array, facets)
        assert set(np.unique(code)) <= {O.INSIDE, O.INTERSECTED, O.OUTSIDE}
        for order in (1, 2, 4):
            r = O.facet_runtime_quadrature(mesh, V.dofmap, phi.x.array, facets, "<", order)
            assert r.offsets[0] == 0 and r.offsets[-1] == r.weights.size and r.parent_map.size == r.offsets.size - 1
            assert set(r.parent_map.tolist()) <= set(facets[code == O.INTERSECTED].tolist())
            X = mesh.x[verts[code == O.INSIDE]]
            full = (np.linalg.norm(X[:, 1] - X[:, 0], axis=1).sum() if tdim == 2 else
                    0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1).sum())
            np.testing.assert_allclose(full + r.weights.sum(), exact, rtol=1e-12)
            rp = O.facet_runtime_quadrature(mesh, V.dofmap, phi.x.array, facets, ">", order)
            Xo = mesh.x[verts[code == O.OUTSIDE]]
            dry = (np.linalg.norm(Xo[:, 1] - Xo[:, 0], axis=1).sum() if tdim == 2 else
                   0.5 * np.linalg.norm(np.cross(Xo[:, 1] - Xo[:, 0], Xo[:, 2] - Xo[:, 0]), axis=1).sum())
            np.testing.assert_allclose(full + r.weights.sum() + dry + rp.weights.sum(), 4.0 if tdim == 2 else 6.0,
                                       rtol=1e-12)
            # the interface inside the boundary facets: two waterline points on the square; on the cube the
            # waterline x + 0.3 y = 0.52 has length 1 + 1 + 2 sqrt(1.09)
            ri = O.facet_runtime_quadrature(mesh, V.dofmap, phi.x.array, facets, "=", order)
            np.testing.assert_allclose(ri.weights.sum(), 2.0 if tdim == 2 else 2.0 + 2.0 * np.sqrt(1.09), rtol=1e-12)


# ---------------------------------------------------------------- vector Nitsche terms (configs[3])
def test_vector_nitsche_kernel_against_an_independent_evaluation():
    """-(sigma(u) n).v - (sigma(v) n).u + gamma (2 mu + lambda)/h u.v on the interface rules: the assembled matrix is
    symmetric, and for P1 vector fields u, v that interpolate LINEAR fields (constant stress) the bilinear form equals
    the same expression evaluated directly at the rules' physical points with the exact stresses."""
    import scipy.sparse as sp

    from cutfemx_b200 import mesh as M
    from util import make_problem

    mu, lam, gam = 1.3, 2.1, 7.0
    for kind, n in (("circle", 8), ("sphere", 4)):
        mesh, Vphi, phi, _ = make_problem(kind, n, 1)
        td = mesh.tdim
        V = M.functionspace(mesh, 1, bs=td)
        dom = O.classify(Vphi.dofmap, phi.x.array)
        ri = O.runtime_quadrature(mesh, Vphi.dofmap, phi.x.array, dom, "=", 2)
        ri.normals = O.normals(mesh, Vphi.dofmap, 1, phi.x.array, ri)
        rp, cols = O.sparsity(V, ri.parent_map)
        A = O.assemble_cells(V, "nitsche_vec", np.zeros(cols.size * td * td), None, ri, (mu, lam, gam), rp, cols)
        nb = V.num_dofs * td
        Ms = sp.bsr_matrix((A.reshape(-1, td, td), cols, rp), shape=(nb, nb)).tocsr()
        assert abs(Ms - Ms.T).max() <= 1e-13 * abs(Ms).max()
        rng = np.random.default_rng(2)
        Bu, cu, Bv, cv = rng.standard_normal((td, td)), rng.standard_normal(td), rng.standard_normal((td, td)), rng.standard_normal(td)
        X = V.dof_coords[:, :td]
        u, v = (X @ Bu.T + cu).reshape(-1), (X @ Bv.T + cv).reshape(-1)
        sig = lambda B: mu * (B + B.T) + lam * np.trace(B) * np.eye(td)
        xp = O.physical_points(mesh, ri).T
        rule_of_pt = np.repeat(np.arange(ri.parent_map.size), np.diff(ri.offsets))
        Xc = mesh.x[mesh.x_dofmap[ri.parent_map]][:, :, :td]
        h = np.array([max(np.linalg.norm(Xc[k, a] - Xc[k, b]) for a in range(td + 1) for b in range(a + 1, td + 1))
                      for k in range(ri.parent_map.size)])
        up, vp, nq = xp @ Bu.T + cu, xp @ Bv.T + cv, ri.normals
        direct = np.sum(ri.weights * (-np.einsum("ab,qb,qa->q", sig(Bu), nq, vp) - np.einsum("ab,qb,qa->q", sig(Bv), nq, up)
                                      + gam * (2 * mu + lam) / h[rule_of_pt] * np.einsum("qa,qa->q", up, vp)))
        assert abs(v @ (Ms @ u) - direct) <= 1e-11 * abs(direct)


# ---------------------------------------------------------------- P2 level sets
def _p2_problem(tdim, n, fn):
    mesh = M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0)) if tdim == 2 else M.create_box(n, n, n)
    V2 = M.functionspace(mesh, 2, permute_seed=5)
    phi = M.interpolate(V2, fn)
    return mesh, V2, phi, O.classify(V2.dofmap, phi)


def test_p2_level_set_reference_assertion_and_measures():
    """test_cut_api.py:1012-1026 on the oracle: the normal of the quadratic circle level set (6 x 6 unit square, order 5)
    is the radial normal, error <= 1e-24; the measures of the cut converge at second order."""
    c, R = (0.47, 0.43), 0.31
    fn = lambda x, y, z: (x - c[0]) ** 2 + (y - c[1]) ** 2 - R * R
    mesh, V2, phi, dom = _p2_problem(2, 6, fn)
    ri = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "=", 5)
    nq = O.normals(mesh, V2.dofmap, 2, phi, ri)
    xp = O.physical_points(mesh, ri)
    d = np.stack([xp[0] - c[0], xp[1] - c[1]], axis=1)
    err = float(np.sum(ri.weights * np.sum((nq - d / np.linalg.norm(d, axis=1, keepdims=True)) ** 2, axis=1)))
    assert ri.weights.size > 0 and err <= 1e-24
    # container invariants (test_cut_api.py:405-421) and parent_map within the cut cells
    assert ri.offsets[0] == 0 and ri.offsets[-1] == ri.weights.size and ri.parent_map.size == ri.offsets.size - 1
    assert set(ri.parent_map.tolist()) <= set(O.locate(dom, "phi=0").tolist())
    errs = []
    for n in (8, 16, 32):
        mesh, V2, phi, dom = _p2_problem(2, n, fn)
        area = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "<", 2).weights.sum() + O.locate(dom, "phi<0").size * 0.5 / n / n
        per = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "=", 2).weights.sum()
        errs.append((abs(area - np.pi * R * R), abs(per - 2 * np.pi * R)))
    for k in (0, 1):
        assert errs[1][k] < errs[0][k] / 3.0 and errs[2][k] < errs[1][k] / 3.0, errs


def test_p2_level_set_plane_is_exact_and_parts_fill_the_cut_cells():
    fn = lambda x, y, z: x + 0.3 * y - 0.2 * z - 0.41
    mesh, V2, phi, dom = _p2_problem(3, 4, fn)
    V1 = M.functionspace(mesh, 1)
    phi1 = M.interpolate(V1, fn)
    dom1 = O.classify(V1.dofmap, phi1)
    cm = 1.0 / (4 ** 3 * 6)
    v2 = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "<", 3).weights.sum() + O.locate(dom, "phi<0").size * cm
    v1 = O.runtime_quadrature(mesh, V1.dofmap, phi1, dom1, "<", 3).weights.sum() + O.locate(dom1, "phi<0").size * cm
    np.testing.assert_allclose(v2, v1, rtol=1e-13)
    a2 = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "=", 2).weights.sum()
    a1 = O.runtime_quadrature(mesh, V1.dofmap, phi1, dom1, "=", 2).weights.sum()
    np.testing.assert_allclose(a2, a1, rtol=1e-13)
    rin = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, "<", 2).weights.sum()
    rout = O.runtime_quadrature(mesh, V2.dofmap, phi, dom, ">", 2).weights.sum()
    np.testing.assert_allclose(rin + rout, O.locate(dom, "phi=0").size * cm, rtol=1e-12)
