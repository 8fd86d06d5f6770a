"""-m gpu: higher-order (P2) level sets: cutfemx.cut(level_set) with a degree-2
level set, default cut options (wrappers/cut.cpp:117-140), run-time quadrature and normals on it.

Against the oracle's restatement (oracle/cutfem_oracle.cpp cut_cell_rule_p2: red refinement through the P2 nodes,
case tables per sub-simplex, closed-form roots of the quadratic edge restriction): containers bit-exact, points 1e-14,
weights 1e-12; plus the reference's own P2 assertion (test_cut_api.py:1012-1026: the normal of a quadratic circle
level set is the radial normal, error 1e-24) and analytic pins (measures of the circle / sphere, exactness for planes)."""
import numpy as np
import pytest

import oracle as O
from cutfemx_b200 import mesh as M

pytestmark = pytest.mark.gpu


def quad_circle(c, R):
    return lambda x, y, z: (x - c[0]) ** 2 + (y - c[1]) ** 2 - R * R


def quad_sphere(c, R):
    return lambda x, y, z: (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2 - R * R


CASES = [(2, 6, quad_circle((0.47, 0.43), 0.31)),       # _quadratic_circle_level_set, test_cut_api.py:36-47
         (2, 13, quad_circle((0.5, 0.5), 0.25)),        # vertices exactly on the circle (zero dofs)
         (3, 4, quad_sphere((0.47, 0.43, 0.41), 0.31)),  # _quadratic_sphere_level_set geometry on tetrahedra
         (3, 6, lambda x, y, z: x + 0.3 * y - 0.2 * z - 0.41)]  # a plane: the P2 interpolant is exact


def _setup(tdim, n, fn):
    import cutfemx_b200 as cfx

    mesh = M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0)) if tdim == 2 else M.create_box(n, n, n)
    V2 = M.functionspace(mesh, 2, permute_seed=5)
    phi = M.Function(V2, "phi").interpolate(fn)
    return cfx, mesh, V2, phi


@pytest.mark.parametrize("tdim,n,fn", CASES, ids=["circle6", "circle13-zeros", "sphere4", "plane6"])
def test_p2_level_set_rules_against_the_oracle(built_lib, tdim, n, fn):
    cfx, mesh, V2, phi = _setup(tdim, n, fn)
    cd = cfx.cut(phi)
    dom = O.classify(V2.dofmap, phi.x.array)
    assert np.array_equal(cd.domain_codes()[: dom.size], dom)
    for sel in ("phi<0", "phi=0", "phi>0"):
        assert np.array_equal(cfx.locate_entities(cd, sel), O.locate(dom, sel))
    for rel, sel, order in (("<", "phi<0", 4), (">", "phi>0", 2), ("=", "phi=0", 5), ("<=", "phi<=0", 4)):
        rg = cfx.runtime_quadrature(cd, sel, order)
        ro = O.runtime_quadrature(mesh, V2.dofmap, phi.x.array, dom, rel, order)
        # container invariants, test_cut_api.py:405-421
        assert rg.offsets[0] == 0 and rg.offsets[-1] == rg.weights.size and rg.parent_map.size == rg.offsets.size - 1
        assert rg.offsets.dtype == np.int32 and rg.parent_map.dtype == np.int32
        assert set(rg.parent_map.tolist()) <= set(O.locate(dom, "phi=0").tolist())
        assert np.array_equal(rg.offsets, ro.offsets) and np.array_equal(rg.parent_map, ro.parent_map)
        np.testing.assert_allclose(rg.points, ro.points, rtol=0, atol=1e-14)
        np.testing.assert_allclose(rg.weights, ro.weights, rtol=1e-12, atol=1e-18)
        assert np.all(rg.weights > 0.0)
        np.testing.assert_allclose(rg.physical_points, O.physical_points(mesh, ro), rtol=0, atol=1e-14)
    # test_cut_api.py:702-710: "<=" rules are the "<" rules
    a, b = cfx.runtime_quadrature(cd, "phi<0", 4), cfx.runtime_quadrature(cd, "phi<=0", 4)
    assert np.array_equal(a.points, b.points) and np.array_equal(a.weights, b.weights)
    # test_extensions volume fractions: the "<" and ">" parts of the cut cells fill them (1e-12)
    r_in, r_out = cfx.runtime_quadrature(cd, "phi<0", 2), cfx.runtime_quadrature(cd, "phi>0", 2)
    cut = O.locate(dom, "phi=0")
    cell_measure = 1.0 / (n ** tdim * (2 if tdim == 2 else 6))
    np.testing.assert_allclose(r_in.weights.sum() + r_out.weights.sum(), cut.size * cell_measure, rtol=1e-12)
    # normals of the P2 level set at the interface points
    ri = cfx.runtime_quadrature(cd, "phi=0", 5)
    ng = cfx.normal(cd, phi, ri)
    oi = O.runtime_quadrature(mesh, V2.dofmap, phi.x.array, dom, "=", 5)
    np.testing.assert_allclose(ng, O.normals(mesh, V2.dofmap, 2, phi.x.array, oi), rtol=0, atol=1e-12)


def test_quadratic_circle_normal_is_the_radial_normal(built_lib):
    """test_cut_api.py:1012-1026: int |n_h - n_exact|^2 over the interface rules of a quadratic circle level set on
    the 6 x 6 unit square, order 5: 0 to 1e-24 (the P2 function represents the quadratic exactly)."""
    c, R = (0.47, 0.43), 0.31
    cfx, mesh, V2, phi = _setup(2, 6, quad_circle(c, R))
    cd = cfx.cut(phi)
    rules = cfx.runtime_quadrature(cd, "phi=0", 5)
    nq = cfx.normal(cd, phi, rules)
    xp = rules.physical_points
    d = np.stack([xp[0] - c[0], xp[1] - c[1]], axis=1)
    n_exact = d / np.linalg.norm(d, axis=1, keepdims=True)
    err = float(np.sum(rules.weights * np.sum((nq - n_exact) ** 2, axis=1)))
    assert rules.weights.size > 0 and err <= 1e-24


@pytest.mark.parametrize("tdim", [2, 3])
def test_measures_converge_at_second_order(built_lib, tdim):
    """Analytic pins: area / perimeter of the circle, volume / area of the sphere; the error falls by ~4 per halving
    and the P2 cut on n cells is as accurate as the P1 cut on 2 n cells (it cuts on the nodes of the refined cell
    with true edge roots)."""
    errs = []
    for n in ((8, 16, 32) if tdim == 2 else (4, 8, 16)):
        fn = quad_circle((0.47, 0.43), 0.31) if tdim == 2 else quad_sphere((0.47, 0.43, 0.41), 0.31)
        cfx, mesh, V2, phi = _setup(tdim, n, fn)
        cd = cfx.cut(phi)
        inside = cfx.locate_entities(cd, "phi<0")
        cell_measure = 1.0 / (n ** tdim * (2 if tdim == 2 else 6))
        vol = cfx.runtime_quadrature(cd, "phi<0", 2).weights.sum() + inside.size * cell_measure
        surf = cfx.runtime_quadrature(cd, "phi=0", 2).weights.sum()
        R = 0.31
        exact_v, exact_s = (np.pi * R * R, 2 * np.pi * R) if tdim == 2 else (4 / 3 * np.pi * R ** 3, 4 * np.pi * R * R)
        errs.append((abs(vol - exact_v) / exact_v, abs(surf - exact_s) / exact_s))
    for k in (0, 1):
        assert errs[1][k] < errs[0][k] / 3.0 and errs[2][k] < errs[1][k] / 3.0, errs
    assert errs[2][0] < (2e-4 if tdim == 2 else 4e-3) and errs[2][1] < (2e-4 if tdim == 2 else 4e-3), errs


def test_plane_is_cut_exactly_and_assembly_runs_on_p2_rules(built_lib):
    """A plane as a P2 level set: the wet volume is exact (1e-13) and equals the P1 cut's; the rules (no moment
    shortcut) feed the assembly kernels: sum b = f x wet volume (test_cut_api.py:858-869)."""
    fn = lambda x, y, z: x + 0.3 * y - 0.2 * z - 0.41
    cfx, mesh, V2, phi2 = _setup(3, 5, fn)
    V1 = M.functionspace(mesh, 1, permute_seed=2)
    phi1 = M.Function(V1, "phi").interpolate(fn)
    cd2 = cfx.cut(phi2)
    rv2 = cfx.runtime_quadrature(cd2, "phi<0", 3)
    in2 = cfx.locate_entities(cd2, "phi<0")
    cd1 = cfx.cut(phi1)
    rv1 = cfx.runtime_quadrature(cd1, "phi<0", 3)
    in1 = cfx.locate_entities(cd1, "phi<0")
    cm = 1.0 / (5 ** 3 * 6)
    v2, v1 = rv2.weights.sum() + in2.size * cm, rv1.weights.sum() + in1.size * cm
    np.testing.assert_allclose(v2, v1, rtol=1e-13)
    L = cfx.fem.CutForm(V1, 1).add_cell_integral("source", in2, rv2, (2.0,))
    b = cfx.fem.assemble_vector(L)
    np.testing.assert_allclose(b.sum(), 2.0 * v2, rtol=1e-12)
    a = cfx.fem.CutForm(V1, 2).add_cell_integral("laplace", in2, rv2, (1.0,))
    A = cfx.fem.assemble_matrix(a)
    dom = O.classify(V2.dofmap, phi2.x.array)
    ro = O.runtime_quadrature(mesh, V2.dofmap, phi2.x.array, dom, "<", 3)
    rp, cols = O.sparsity(V1, np.concatenate([in2, ro.parent_map]), np.zeros((0, 4), dtype=np.int32))
    ref = np.zeros(cols.size)
    O.assemble_cells(V1, "laplace", ref, in2, ro, (1.0,), rp, cols)
    assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, cols)
    assert np.linalg.norm(A.data - ref) <= 1e-11 * np.linalg.norm(ref)
