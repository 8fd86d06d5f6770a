"""-m gpu: facets as hosts of the cut: cutfemx.cut(level_set, facets, entity_dim =
tdim - 1), locate_entities, runtime_quadrature on them -- against the oracle's restatement and the assertions of the
reference's own tests (python/tests/test_cut_api.py:171-188 partition, :349-367 and :424-501 rule invariants, weight
sums independent of the host list).  Bars: lists bit-exact, points 1e-14, weights 1e-12."""
import numpy as np
import pytest

import oracle as O
from cutfemx_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _problem(tdim, n, fn):
    import cutfemx_b200 as cfx

    mesh = M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0)) if tdim == 2 else M.create_box(n, n, n)
    V = M.functionspace(mesh, 1, permute_seed=9)
    phi = M.Function(V, "phi").interpolate(fn)
    return cfx, mesh, V, phi


CASES = [(2, 9, lambda x, y, z: x - 0.51), (2, 11, M.sphere_level_set((0.45, 0.55, 0.0), 0.38)),
         (3, 5, lambda x, y, z: x + 0.3 * y - 0.52), (3, 6, M.sphere_level_set((0.4, 0.5, 0.55), 0.42))]


@pytest.mark.parametrize("tdim,n,fn", CASES, ids=["line2d", "circle2d", "plane3d", "sphere3d"])
def test_facet_hosts_against_the_oracle(built_lib, tdim, n, fn):
    cfx, mesh, V, phi = _problem(tdim, n, fn)
    all_facets = np.nonzero(np.diff(mesh.f2c_offsets) >= 1)[0].astype(np.int32)  # the synthetic numbering has gaps
    exterior = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0].astype(np.int32)
    interior = np.nonzero(np.diff(mesh.f2c_offsets) == 2)[0].astype(np.int32)
    for facets in (all_facets, exterior, interior[::3]):
        cd = cfx.cut(phi, facets, tdim - 1)
        assert cd.tdim == tdim - 1 and cd.entity_dim == tdim - 1
        code, _, _ = O.classify_facets(mesh, V.dofmap, phi.x.array, facets)
        neg, cut, pos = (cfx.locate_entities(cd, s) for s in ("phi<0", "phi=0", "phi>0"))
        assert np.array_equal(neg, facets[code == O.INSIDE])
        assert np.array_equal(cut, facets[code == O.INTERSECTED])
        assert np.array_equal(pos, facets[code == O.OUTSIDE])
        # test_cut_api.py:171-188: the three parts are disjoint and cover the host list
        assert np.array_equal(np.sort(np.concatenate([neg, cut, pos])), np.sort(facets))
        assert np.array_equal(cfx.locate_entities(cd, "phi<=0"), facets[code != O.OUTSIDE])
        for rel, sel, order in (("<", "phi<0", 2), (">", "phi>0", 3), ("<=", "phi<=0", 1), ("=", "phi=0", 2)):
            r = cfx.runtime_quadrature(cd, sel, order)
            ro = O.facet_runtime_quadrature(mesh, V.dofmap, phi.x.array, facets, rel, order)
            # test_cut_api.py:405-421 / :439-455
            assert r.kind == "per_entity" and r.tdim == tdim - 1
            assert r.offsets[0] == 0 and r.offsets[-1] == r.weights.size and r.parent_map.size == r.offsets.size - 1
            assert r.offsets.dtype == np.int32 and r.parent_map.dtype == np.int32
            assert set(r.parent_map.tolist()) <= set(cut.tolist())
            assert np.array_equal(r.offsets, ro.offsets) and np.array_equal(r.parent_map, ro.parent_map)
            np.testing.assert_allclose(r.points, ro.points, rtol=0, atol=1e-14)
            np.testing.assert_allclose(r.weights, ro.weights, rtol=1e-12, atol=1e-18)
            xp = r.with_physical_points().physical_points
            assert xp.shape == (mesh.gdim, r.weights.size) and np.all(np.isfinite(xp))
            # the points lie in the selected part: the P1 level set there has the right sign (or is ~0)
            if r.weights.size:
                verts, _, _ = O.facet_vertices(mesh, r.parent_map)
                rule_of_pt = np.repeat(np.arange(r.parent_map.size), np.diff(r.offsets))
                lam = np.concatenate([1.0 - r.points.sum(axis=1, keepdims=True), r.points], axis=1)
                xref = np.einsum("pk,pkd->pd", lam, mesh.x[verts[rule_of_pt]][:, :, : mesh.gdim])
                np.testing.assert_allclose(xp.T, xref, rtol=0, atol=1e-14)
                _, _, pv = O.classify_facets(mesh, V.dofmap, phi.x.array, r.parent_map)
                phi_h = np.einsum("pk,pk->p", lam, pv[rule_of_pt])
                if rel == "=":
                    assert np.all(np.abs(phi_h) < 1e-12)
                else:
                    assert np.all(phi_h < 1e-12) if rel in ("<", "<=") else np.all(phi_h > -1e-12)


def test_weight_sum_does_not_depend_on_the_host_list(built_lib):
    """test_cut_api.py:424-462: rules from all exterior facets and from the cut exterior facets only sum alike; and
    for phi = x - 0.51 on the unit square the wet part of the boundary has length 0.51 + 0.51 + 1."""
    cfx, mesh, V, phi = _problem(2, 9, lambda x, y, z: x - 0.51)
    exterior = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0].astype(np.int32)
    cd = cfx.cut(phi, exterior, 1)
    cut = cfx.locate_entities(cd, "phi=0")
    inside = cfx.locate_entities(cd, "phi<0")
    r_all = cfx.runtime_quadrature(cd, "phi<0", 2)
    r_cut = cfx.runtime_quadrature(cfx.cut(phi, cut, 1), "phi<0", 2)
    assert cut.size == 2 and np.array_equal(r_all.parent_map, r_cut.parent_map)
    np.testing.assert_allclose(r_all.weights.sum(), r_cut.weights.sum(), rtol=1e-14)
    verts, _, _ = O.facet_vertices(mesh, inside)
    full = np.linalg.norm(mesh.x[verts[:, 1]] - mesh.x[verts[:, 0]], axis=1).sum()
    np.testing.assert_allclose(full + r_all.weights.sum(), 0.51 + 0.51 + 1.0, rtol=1e-12)
    # test_cut_api.py:504-527: one * ds(subdomain_data=rules) assembled as a scalar
    form = cfx.fem.CutForm(V, 0).add_exterior_facet_integral("one", r_all, (2.0,))
    value = cfx.fem.assemble_scalar(form)
    assert np.isfinite(value) and value > 0.0
    np.testing.assert_allclose(value, 2.0 * r_all.weights.sum(), rtol=1e-13)


@pytest.mark.parametrize("tdim,n,fn,deg", [(2, 9, lambda x, y, z: x - 0.51, 1), (2, 7, M.sphere_level_set((0.45, 0.55, 0.0), 0.62), 2),
                                           (3, 5, lambda x, y, z: x + 0.3 * y - 0.52, 1),
                                           (3, 4, M.sphere_level_set((0.4, 0.5, 0.55), 0.7), 2)],
                         ids=["line2d-P1", "circle2d-P2", "plane3d-P1", "sphere3d-P2"])
def test_exterior_facet_integrals_on_cut_boundary_facets(built_lib, tdim, n, fn, deg):
    """SURVEY section 8(f) rank 3, second half: `alpha u v ds(rules)` (Robin / penalty term), `g v ds(rules)` and
    `1 ds(rules)` on the wet part of the boundary -- the run-time exterior-facet integrals of
    _facet_payload_with_rows (_runintgen_adapter.py:605-680).  The facet rules are mapped into the reference
    coordinates of the facets' cells and the cell kernel families evaluate them; corner cells own two or three
    boundary facets.  Against the oracle's restatement (sparsity bit-exact, values 1e-12) and the analytic pins:
    the load vector sums to g x (wet boundary measure), the mass matrix sums to alpha x the same."""
    cfx, mesh, Vphi, phi = _problem(tdim, n, fn)
    V = M.functionspace(mesh, deg, permute_seed=4)
    exterior = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0].astype(np.int32)
    cd = cfx.cut(phi, exterior, tdim - 1)
    rules = cfx.runtime_quadrature(cd, "phi<0", 4)
    ro = O.facet_runtime_quadrature(mesh, Vphi.dofmap, phi.x.array, exterior, "<", 4)
    rc = O.facet_rules_on_cells(mesh, ro)
    assert rules.num_rules == ro.parent_map.size > 0
    assert np.unique(rc.parent_map).size < rc.parent_map.size or tdim == 2 and n == 9  # corner cells appear twice
    alpha, g = 3.5, -1.25
    cells = np.unique(rc.parent_map)
    rp, cols = O.sparsity(V, cells, np.zeros((0, 4), dtype=np.int32))
    ref = np.zeros(cols.size)
    O.assemble_cells(V, "mass", ref, None, rc, (alpha,), rp, cols)
    bref = np.zeros(V.num_dofs)
    O.assemble_cells(V, "source", bref, None, rc, (g,))
    a = cfx.fem.CutForm(V, 2).add_exterior_facet_integral("mass", rules, (alpha,))
    A = cfx.fem.assemble_matrix(a)
    assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, cols)
    assert np.linalg.norm(A.data - ref) <= 1e-12 * np.linalg.norm(ref)
    L = cfx.fem.CutForm(V, 1).add_exterior_facet_integral("source", rules, (g,))
    b = cfx.fem.assemble_vector(L)
    assert np.linalg.norm(b - bref) <= 1e-12 * np.linalg.norm(bref)
    wet = rules.weights.sum()
    np.testing.assert_allclose(b.sum(), g * wet, rtol=1e-12)          # partition of unity on the boundary
    np.testing.assert_allclose(A.data.sum(), alpha * wet, rtol=1e-12)
    m0 = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_exterior_facet_integral("one", rules, (1.0,)))
    np.testing.assert_allclose(m0, wet, rtol=1e-13)


def test_facet_host_errors(built_lib):
    cfx, mesh, V, phi = _problem(2, 5, lambda x, y, z: x - 0.51)
    with pytest.raises(cfx.CfxError):  # validate_local_entities
        cfx.cut(phi, np.array([mesh.num_facets], dtype=np.int32), 1)
    cd = cfx.cut(phi, np.nonzero(np.diff(mesh.f2c_offsets) >= 1)[0].astype(np.int32), 1)
    with pytest.raises(ValueError, match="cell-hosted"):  # python/cutfemx/cut.py:350-351
        cfx.ghost_penalty_facets(cd, "phi<0")
    ri = cfx.runtime_quadrature(cd, "phi=0", 2)  # the interface inside segment facets: the cut points, weight 1
    assert ri.tdim == 1 and np.all(ri.weights == 1.0) and ri.weights.size == ri.parent_map.size
    np.testing.assert_allclose(ri.with_physical_points().physical_points[0], 0.51, rtol=0, atol=1e-14)
    rules = cfx.runtime_quadrature(cd, "phi<0", 2)
    with pytest.raises(cfx.CfxError):  # facet rules do not fit cell integrals
        cfx.fem.CutForm(V, 0).add_cell_integral("one", None, rules, (1.0,))
    with pytest.raises(ValueError):
        cfx.cut(phi, np.arange(3, dtype=np.int32), 0)
