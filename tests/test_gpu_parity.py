"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): classification, cut/inside/outside lists, facet lists and CSR
sparsity bit-exact; quadrature weights summing to the oracle's cut volume/area to 1e-12 relative;
matrix and vector entries to 1e-11 relative in Frobenius norm.
"""
import numpy as np
import pytest

from util import GpuRun, OracleRun, make_problem

pytestmark = pytest.mark.gpu

CASES = [
    ("line", 3, 1),      # test_locate_entities.py:13-35 geometry (3x3, phi = x - 0.51)
    ("circle", 21, 1),   # quickstart / test_cut_api.py:1268-1300
    ("circle", 64, 1),   # BASELINE configs[0] (has phi == 0 vertices -> degenerate cuts)
    ("circle", 32, 2),   # P2 on triangles (configs[1] element)
    ("sphere", 12, 1),   # configs[2] element
    ("sphere", 6, 2),
    ("torus", 14, 1),
]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}{c[1]}-P{c[2]}")
def runs(request, built_lib):
    kind, n, deg = request.param
    mesh, Vphi, phi, V = make_problem(kind, n, deg)
    return OracleRun(mesh, Vphi, phi, V), GpuRun(mesh, Vphi, phi, V)


def test_classification_bit_exact(runs):
    o, g = runs
    assert np.array_equal(o.domain, g.domain[: o.domain.size])
    counts = g.cut_data.counts()
    assert counts == (o.inside.size, o.cut.size, o.outside.size)


def test_lists_bit_exact(runs):
    o, g = runs
    for name in ("inside", "cut", "outside"):
        a, b = getattr(o, name), getattr(g, name)
        assert a.dtype == b.dtype == np.int32
        assert np.array_equal(a, b), name


def test_rules(runs):
    o, g = runs
    for ro, rg in ((o.rv, g.rv), (o.ro, g.ro), (o.ri, g.ri)):
        assert np.array_equal(ro.offsets, rg.offsets)
        assert np.array_equal(ro.parent_map, rg.parent_map)
        assert rg.offsets.dtype == np.int32 and rg.parent_map.dtype == np.int32
        assert rg.points.shape == ro.points.shape
        np.testing.assert_allclose(rg.points, ro.points, rtol=0, atol=1e-14)
        np.testing.assert_allclose(rg.weights, ro.weights, rtol=1e-12, atol=1e-18)
        assert abs(rg.weights.sum() - ro.weights.sum()) <= 1e-12 * abs(ro.weights.sum())


def test_physical_points_and_normals(runs):
    import oracle as O

    o, g = runs
    mesh = o.mesh
    np.testing.assert_allclose(g.ri.physical_points, O.physical_points(mesh, o.ri), rtol=0, atol=1e-14)
    np.testing.assert_allclose(g.normals, o.ri.normals, rtol=0, atol=1e-12)


def test_surface_provenance_bit_exact(runs):
    """RuntimeSurfaceProvenance (cut.cpp:1273-1308): aligned with the interface rules, empty for volume selectors."""
    import oracle as O

    o, g = runs
    ref = O.surface_provenance(o.ri, "=", cut_cells=o.cut)
    pv = g.ri.surface_provenance
    assert pv.selector == g.ri.selector and pv.level_set_index == ref["level_set_index"] == 0
    assert pv.size() == g.ri.num_rules  # "aligned with runtime quadrature rules" (cut.cpp:1289-1295)
    for k in ("cut_cell_ids", "parent_cell_ids", "local_zero_entity_ids", "dimensions"):
        a = getattr(pv, k)
        assert a.dtype == np.int32 and np.array_equal(a, ref[k]), k
    assert np.array_equal(o.cut[pv.cut_cell_ids], pv.parent_cell_ids) and np.all(pv.dimensions == o.mesh.tdim - 1)
    vol = g.rv.surface_provenance
    assert vol.empty() and vol.level_set_index == -1 == O.surface_provenance(o.rv, "<")["level_set_index"]


def test_ghost_facets_bit_exact(runs):
    o, g = runs
    assert np.array_equal(o.ghost, g.ghost)
    assert np.array_equal(o.rows4, g.rows4)


def test_sparsity_bit_exact(runs):
    o, g = runs
    assert np.array_equal(o.row_ptr, g.row_ptr)
    assert np.array_equal(o.cols, g.cols)


def test_matrix_vector_scalar(runs):
    o, g = runs
    assert rel(g.vals, o.vals) < 1e-11
    assert rel(g.b, o.b) < 1e-11
    assert abs(g.volume - o.volume) <= 1e-12 * abs(o.volume)
    assert abs(g.area - o.area) <= 1e-12 * abs(o.area)


def test_deterministic(runs):
    """Two assemblies give bit-identical values (no floating-point atomics anywhere)."""
    import cutfemx_b200 as cfx

    o, g = runs
    A2 = cfx.fem.assemble_matrix(g.a)
    assert np.array_equal(A2.data, g.vals)
    b2 = cfx.fem.assemble_vector(g.L)
    assert np.array_equal(b2, g.b)


def test_active_domain_and_deactivation(runs):
    """cutfemx.fem.active_domain / deactivate_outside (fem/deactivate.h:387-418); reference test:
    test_cut_api.py:841-845 (active cells = unique(inside U parent_map), here plus the facet cells)."""
    import oracle as O

    import cutfemx_b200 as cfx

    o, g = runs
    cells_o, inactive_o = O.active_domain(o.V, [o.inside, o.rv.parent_map, o.ri.parent_map], o.rows4)
    dom = cfx.fem.active_domain(g.a)
    assert np.array_equal(dom.active_cells, cells_o) and dom.active_cells.dtype == np.int32
    assert np.array_equal(dom.inactive_dofs, inactive_o)
    assert np.array_equal(np.unique(np.concatenate([o.inside, o.rv.parent_map])),
                          np.setdiff1d(cells_o, np.setdiff1d(np.concatenate([o.rows4[:, 0], o.rows4[:, 2]]),
                                                             np.concatenate([o.inside, o.rv.parent_map]))))
    vals_o, b_o = o.vals.copy(), o.b.copy()
    O.deactivate_outside(o.row_ptr, o.cols, vals_o, inactive_o, 2.0, b_o, -3.0)
    A = cfx.fem.assemble_matrix(g.a)
    b = g.b.copy()
    cfx.fem.deactivate_outside(A, b, dom, diagonal=2.0, rhs_value=-3.0)
    assert rel(A.data, vals_o) < 1e-11 and rel(b, b_o) < 1e-11
    assert np.all(A.data[np.isin(np.repeat(np.arange(o.V.num_dofs), np.diff(o.row_ptr)), inactive_o)] == 2.0)
    with pytest.raises(ValueError):
        cfx.fem.active_domain(g.L)


def test_assemble_system_is_bit_identical_to_separate_calls(runs):
    """cfx_assemble_system fuses the right-hand side into the matrix gather; same bits as the two calls."""
    import torch

    import cutfemx_b200 as cfx

    o, g = runs
    A = cfx.fem.create_matrix(g.a)
    b = torch.full((o.V.num_dofs,), 7.0, dtype=torch.float64, device="cuda:0")
    cfx.fem.assemble_system(g.a, A, g.L, b)
    assert np.array_equal(A.data, g.vals)
    assert np.array_equal(b.cpu().numpy(), g.b)
    # adding into an existing vector
    b2 = torch.ones(o.V.num_dofs, dtype=torch.float64, device="cuda:0")
    A2 = cfx.fem.create_matrix(g.a)
    cfx.fem.assemble_system(g.a, A2, g.L, b2, zero_b=False)
    ref = np.ones(o.V.num_dofs)
    cfx.fem.assemble_vector(g.L, ref)
    assert np.array_equal(b2.cpu().numpy(), ref)
