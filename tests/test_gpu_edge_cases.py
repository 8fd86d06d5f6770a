"""-m gpu: edge cases of the cut path through the C ABI, against the oracle / the reference's
documented behaviour: empty and degenerate cuts, several level sets with and/or selectors, order 0
and high orders, facet helpers, error behaviour (the exceptions the reference raises at the same
places), re-cutting a moved level set with the same handles."""
import numpy as np
import pytest

import oracle as O
from cutfemx_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _setup(tdim, n, fn, name="phi"):
    import cutfemx_b200 as cfx

    mesh = M.create_rectangle(n, n, (-1.0, -1.0), (1.0, 1.0)) if tdim == 2 else M.create_box(n, n, n)
    V = M.functionspace(mesh, 1, permute_seed=5)
    phi = M.Function(V, name).interpolate(fn)
    return cfx, mesh, V, phi


@pytest.mark.parametrize("sign", [+1.0, -1.0])
def test_no_cut_cells_at_all(sign, built_lib):
    """Level set of one sign everywhere: empty cut list, empty rules (offsets == [0]), no ghost facets,
    and assembly over the whole (or empty) mesh still works."""
    cfx, mesh, V, phi = _setup(2, 6, lambda x, y, z: sign * (3.0 + x))
    cd = cfx.cut(phi)
    assert cfx.locate_entities(cd, "phi=0").size == 0
    inside = cfx.locate_entities(cd, "phi<0")
    assert inside.size == (mesh.num_cells if sign < 0 else 0)
    for sel in ("phi<0", "phi>0", "phi=0"):
        r = cfx.runtime_quadrature(cd, sel, 3)
        assert r.num_rules == 0 and r.total_points == 0
        assert r.offsets.tolist() == [0] and r.parent_map.size == 0 and r.points.shape == (0, 2)
    assert cfx.ghost_penalty_facets(cd, "phi<0").size == 0
    a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", inside, cfx.runtime_quadrature(cd, "phi<0", 2), (1.0,))
    if sign < 0:
        A = cfx.fem.assemble_matrix(a)
        rp, cols = O.sparsity(V, inside)
        ref = O.assemble_cells(V, "laplace", np.zeros(cols.size), inside, None, (1.0,), rp, cols)
        assert np.array_equal(A.indptr, rp) and np.array_equal(A.indices, cols)
        assert np.linalg.norm(A.data - ref) <= 1e-11 * np.linalg.norm(ref)
    else:
        A = cfx.fem.assemble_matrix(a)  # only the deactivation diagonal remains
        assert A.nnz == V.num_dofs and np.array_equal(A.indices, np.arange(V.num_dofs)) and not A.data.any()
        with pytest.raises(cfx.CfxError):
            cfx.fem.active_domain(a)  # deactivate.h:160-164 "found no active background cells"


def test_degenerate_zero_vertices(built_lib):
    """phi == 0 exactly on a mesh line: every touching cell is 'intersected' (level-sets.md:74-92,
    test_cut_api.py:191-208), cut parts may be empty -> such cells get no rule; sums stay exact."""
    cfx, mesh, V, phi = _setup(2, 8, lambda x, y, z: x - 0.25)  # x = 0.25 is a mesh line of the 8x8 mesh
    cd = cfx.cut(phi)
    dom = O.classify(V.dofmap, phi.x.array)
    assert np.array_equal(cd.domain_codes(), dom)
    for rel, sel in (("<", "phi<0"), (">", "phi>0"), ("=", "phi=0")):
        rg = cfx.runtime_quadrature(cd, sel, 4)
        ro = O.runtime_quadrature(mesh, V.dofmap, phi.x.array, dom, rel, 4)
        assert np.array_equal(rg.offsets, ro.offsets) and np.array_equal(rg.parent_map, ro.parent_map)
        np.testing.assert_allclose(rg.weights, ro.weights, rtol=1e-12, atol=1e-18)
    rl = cfx.runtime_quadrature(cd, "phi<0", 2)
    inside = cfx.locate_entities(cd, "phi<0")
    vol = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", inside, rl, (1.0,)))
    np.testing.assert_allclose(vol, 1.25 * 2.0, rtol=1e-13)  # {x < 0.25} in [-1,1]^2


def test_two_level_sets_and_selectors(built_lib):
    """cut([phi, phi1]) + DNF selectors (cut.cpp:877-924; names frozen as phi, phi1: cut.cpp:59-62)."""
    import cutfemx_b200 as cfx

    mesh = M.create_box(7, 7, 7)
    V = M.functionspace(mesh, 1, permute_seed=2)
    f0 = M.Function(V, "f").interpolate(M.sphere_level_set((0.5, 0.5, 0.5), 0.35))
    f1 = M.Function(V, "u").interpolate(lambda x, y, z: z - 0.45)
    cd = cfx.cut([f0, f1])
    assert cd.level_set_names == ("phi", "phi1")
    dom = np.stack([O.classify(V.dofmap, f0.x.array), O.classify(V.dofmap, f1.x.array)])
    for sel in ("phi<0 and phi1<0", "phi<0 or phi1>=0", "phi=0 and phi1<=0 or phi>0 and phi1=0", " phi1 = 0 "):
        assert np.array_equal(cfx.locate_entities(cd, sel), O.locate(dom, sel, ("phi", "phi1"))), sel
    r1 = cfx.runtime_quadrature(cd, "phi1<0", 2)
    o1 = O.runtime_quadrature(mesh, V.dofmap, f1.x.array, dom[1], "<", 2)
    assert np.array_equal(r1.parent_map, o1.parent_map)
    np.testing.assert_allclose(r1.weights, o1.weights, rtol=1e-12, atol=1e-18)
    with pytest.raises(ValueError):
        cfx.locate_entities(cd, "psi<0")
    with pytest.raises(NotImplementedError):
        cfx.runtime_quadrature(cd, "phi<0 and phi1<0", 2)


@pytest.mark.parametrize("order", [0, 1, 3, 6, 9])
def test_orders(order, built_lib):
    cfx, mesh, V, phi = _setup(3, 5, M.sphere_level_set((0.5, 0.5, 0.5), 0.35))
    cd = cfx.cut(phi)
    dom = O.classify(V.dofmap, phi.x.array)
    for rel, sel in (("<", "phi<0"), ("=", "phi=0")):
        rg = cfx.runtime_quadrature(cd, sel, order)
        ro = O.runtime_quadrature(mesh, V.dofmap, phi.x.array, dom, rel, order)
        assert np.array_equal(rg.offsets, ro.offsets)
        np.testing.assert_allclose(rg.points, ro.points, rtol=0, atol=1e-14)
        np.testing.assert_allclose(rg.weights, ro.weights, rtol=1e-12, atol=1e-18)
    with pytest.raises(ValueError):
        cfx.runtime_quadrature(cd, "phi<0", -1)  # cut.cpp:164-168
    with pytest.raises(ValueError):
        cfx.runtime_quadrature(cd, "phi<0", 2, backend="algoim")  # cut.cpp:207-237: no algoim on simplices


def test_facet_helpers(built_lib):
    """interior_facets_for_cells (cut.cpp:926-994) and facet_integration_rows (wrappers/cut.cpp:54-115)."""
    import cutfemx_b200 as cfx

    mesh = M.create_rectangle(4, 4, (0.0, 0.0), (1.0, 1.0))
    sel = np.arange(0, mesh.num_cells, 2, dtype=np.int32)
    assert np.array_equal(cfx.interior_facets_for_cells(mesh, sel), O.interior_facets_for_cells(mesh, sel))
    allc = np.arange(mesh.num_cells, dtype=np.int32)
    allf = cfx.interior_facets_for_cells(mesh, allc)
    assert np.array_equal(allf, np.nonzero(np.diff(mesh.f2c_offsets) == 2)[0])  # test_cut_api.py:1194-1196
    assert np.array_equal(cfx.facet_integration_rows(mesh, allf), O.facet_rows(mesh, allf))
    assert cfx.interior_facets_for_cells(mesh, np.zeros(0, np.int32)).size == 0
    with pytest.raises(cfx.CfxError):  # cut.cpp:963-964 "Cell index is out of range."
        cfx.interior_facets_for_cells(mesh, np.array([mesh.num_cells], dtype=np.int32))
    bnd = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0][:1].astype(np.int32)
    with pytest.raises(cfx.CfxError):  # wrappers/cut.cpp:101-105: facet without two adjacent cells
        cfx.facet_integration_rows(mesh, bnd)


def test_update_recuts_a_moved_level_set(built_lib):
    """cutfemx.update (cut.cpp:845-868): same CutData, new dof values (demo_moving_poisson.py:69-73)."""
    cfx, mesh, V, phi = _setup(3, 6, M.sphere_level_set((0.4, 0.5, 0.5), 0.3))
    cd = cfx.cut(phi)
    first = cfx.locate_entities(cd, "phi=0")
    phi.x.array[:] = M.interpolate(V, M.sphere_level_set((0.6, 0.5, 0.5), 0.3))
    cfx.update(cd)
    dom = O.classify(V.dofmap, phi.x.array)
    second = cfx.locate_entities(cd, "phi=0")
    assert np.array_equal(second, O.locate(dom, "phi=0")) and not np.array_equal(first, second)
    rv = cfx.runtime_quadrature(cd, "phi<0", 2)
    ro = O.runtime_quadrature(mesh, V.dofmap, phi.x.array, dom, "<", 2)
    np.testing.assert_allclose(rv.weights.sum(), ro.weights.sum(), rtol=1e-12)


def test_form_errors(built_lib):
    import cutfemx_b200 as cfx

    cfx_, mesh, V, phi = _setup(2, 6, M.sphere_level_set((0.0, 0.0, 0.0), 0.5))
    cd = cfx.cut(phi)
    ri = cfx.runtime_quadrature(cd, "phi=0", 2)
    with pytest.raises(cfx.CfxError):  # interface kernels need normals
        cfx.fem.CutForm(V, 2).add_cell_integral("nitsche", None, ri, (40.0,))
    with pytest.raises(ValueError):    # rank mismatch
        cfx.fem.CutForm(V, 2).add_cell_integral("source", None, ri, (1.0,))
    with pytest.raises(RuntimeError):  # assembler.h:444-448
        cfx.fem.create_matrix(cfx.fem.CutForm(V, 1))
    with pytest.raises(cfx.CfxError):  # entity index out of range
        a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", np.array([mesh.num_cells + 3], dtype=np.int32))
        cfx.fem.assemble_matrix(a)


@pytest.mark.parametrize("kind,n,deg", [("circle", 12, 1), ("circle", 7, 2), ("sphere", 6, 1), ("sphere", 4, 2)])
def test_square_functional_of_a_function_coefficient(built_lib, kind, n, deg):
    """pack_coefficients (pack_form.h:68-158) + the error functional of demo_poisson.py:213 as c0 * w_h**2."""
    import cutfemx_b200 as cfx
    from util import make_problem

    mesh, Vphi, phi, _ = make_problem(kind, n, 1)
    V = M.functionspace(mesh, deg, permute_seed=3)
    dom = O.classify(Vphi.dofmap, phi.x.array)
    inside = O.locate(dom, "phi<0")
    rv = O.runtime_quadrature(mesh, Vphi.dofmap, phi.x.array, dom, "<", 4)
    w = np.sin(3.0 * V.dof_coords[:, 0]) + V.dof_coords[:, 1] ** 2
    ref = np.zeros(1)
    with O.coefficient(w):
        O.assemble_cells(V, "square_fn", ref, inside, rv, (1.5,))
    cd = cfx.cut(phi)
    form = cfx.fem.CutForm(V, 0).add_cell_integral("square_fn", cfx.locate_entities(cd, "phi<0"),
                                                    cfx.runtime_quadrature(cd, "phi<0", 4), (1.5,))
    with pytest.raises(cfx.CfxError):  # the kernel needs its coefficient
        cfx.fem.assemble_scalar(form)
    form.set_coefficient(w)
    got = cfx.fem.assemble_scalar(form)
    assert abs(got - ref[0]) <= 1e-12 * abs(ref[0])
    with pytest.raises(cfx.CfxError):  # one value per dof of the space
        form.set_coefficient(w[:-1])


def test_cell_subset_as_host(built_lib):
    """cutfemx.cut(level_set, entities, entity_dim = tdim) (test_cut_api.py:160-168, :211-222, :728-745): only the
    listed cells host the cut -- the three domains partition the subset, rules exist only for its cut cells."""
    cfx, mesh, V, phi = _setup(2, 9, M.sphere_level_set((0.1, -0.05, 0.0), 0.55))
    dom_all = O.classify(V.dofmap, phi.x.array)
    rng = np.random.default_rng(5)
    subset = np.sort(rng.choice(mesh.num_cells, size=mesh.num_cells // 3, replace=False)).astype(np.int32)
    cd = cfx.cut(phi, subset, mesh.tdim)
    neg, cut, pos = (cfx.locate_entities(cd, s) for s in ("phi<0", "phi=0", "phi>0"))
    assert np.array_equal(cut, np.intersect1d(O.locate(dom_all, "phi=0"), subset)) and cut.size > 0
    assert np.array_equal(neg, np.intersect1d(O.locate(dom_all, "phi<0"), subset))
    assert np.array_equal(np.sort(np.concatenate([neg, cut, pos])), subset)
    rules = cfx.runtime_quadrature(cd, "phi<0", 2)
    ro = O.runtime_quadrature(mesh, V.dofmap, phi.x.array, dom_all, "<", 2)
    keep = np.isin(ro.parent_map, subset)
    assert np.array_equal(rules.parent_map, ro.parent_map[keep])
    w_ref = sum(ro.weights[ro.offsets[i]:ro.offsets[i + 1]].sum() for i in np.nonzero(keep)[0])
    np.testing.assert_allclose(rules.weights.sum(), w_ref, rtol=1e-12)
    # an unsorted list with a repeated entry: locate_entities answers in the order of the list (cut.cpp:574-576)
    shuffled = np.concatenate([subset[::-1], subset[:3]]).astype(np.int32)
    cd_s = cfx.cut(phi, shuffled, mesh.tdim)
    cut_s = cfx.locate_entities(cd_s, "phi=0")
    assert np.array_equal(cut_s, shuffled[np.isin(shuffled, cut)])
    # the next plain cut sees every cell again
    cd2 = cfx.cut(phi)
    assert np.array_equal(cfx.locate_entities(cd2, "phi=0"), O.locate(dom_all, "phi=0"))
    with pytest.raises(ValueError, match="entity_dim must be supplied"):
        cfx.cut(phi, entities=np.arange(11, dtype=np.int32))
    with pytest.raises(ValueError, match="entity_dim is only valid"):
        cfx.cut(phi, entity_dim=0)
    with pytest.raises(ValueError):  # cut.cpp:545-550: positive-dimensional entities only
        cfx.cut(phi, np.arange(4, dtype=np.int32), 0)
    with pytest.raises(cfx.CfxError):
        cfx.cut(phi, np.array([mesh.num_cells], dtype=np.int32), mesh.tdim)


def test_two_cut_data_on_one_mesh_do_not_alias(built_lib):
    """Reference CutData objects are independent (cut.cpp builds a mesh view and a level-set set per call): an
    earlier CutData keeps answering for ITS level set and host cells after another cut() on the same mesh, and a
    dof index outside the level-set space is refused at bind time (cut.cpp:303-306)."""
    import cutfemx_b200 as cfx
    from cutfemx_b200 import mesh as M

    mesh = M.create_rectangle(16, 16, (-1.0, -1.0), (1.0, 1.0))
    V = M.functionspace(mesh, 1, permute_seed=3)
    phi_a = M.Function(V, "phi").interpolate(M.sphere_level_set((0.0, 0.0, 0.0), 0.5))
    phi_b = M.Function(V, "phi").interpolate(M.sphere_level_set((0.3, 0.1, 0.0), 0.3))
    dom_a, dom_b = O.classify(V.dofmap, phi_a.x.array), O.classify(V.dofmap, phi_b.x.array)
    cd_a = cfx.cut(phi_a)
    in_a = cfx.locate_entities(cd_a, "phi<0")
    cd_b = cfx.cut(phi_b)
    sub = np.arange(0, mesh.num_cells, 3, dtype=np.int32)
    cd_sub = cfx.cut(phi_a, sub, mesh.tdim)
    assert np.array_equal(cfx.locate_entities(cd_b, "phi<0"), O.locate(dom_b, "phi<0"))
    # the first CutData still answers for phi_a on all cells
    assert np.array_equal(cfx.locate_entities(cd_a, "phi<0"), in_a) and np.array_equal(in_a, O.locate(dom_a, "phi<0"))
    assert np.array_equal(cfx.locate_entities(cd_sub, "phi=0"), np.intersect1d(O.locate(dom_a, "phi=0"), sub))
    ra, rb = cfx.runtime_quadrature(cd_a, "phi<0", 2), cfx.runtime_quadrature(cd_b, "phi<0", 2)
    oa = O.runtime_quadrature(mesh, V.dofmap, phi_a.x.array, dom_a, "<", 2)
    ob = O.runtime_quadrature(mesh, V.dofmap, phi_b.x.array, dom_b, "<", 2)
    assert np.array_equal(ra.parent_map, oa.parent_map) and np.array_equal(rb.parent_map, ob.parent_map)
    assert abs(ra.weights.sum() - oa.weights.sum()) < 1e-13 and abs(rb.weights.sum() - ob.weights.sum()) < 1e-13
    assert cd_a.counts() == (O.locate(dom_a, "phi<0").size, O.locate(dom_a, "phi=0").size, O.locate(dom_a, "phi>0").size)
    # update() after the values changed refreshes THIS CutData only
    phi_b.x.array[:] = phi_a.x.array
    cd_b.update()
    assert np.array_equal(cfx.locate_entities(cd_b, "phi<0"), in_a)
    del cd_b, cd_sub
    assert np.array_equal(cfx.locate_entities(cd_a, "phi<0"), in_a)
    # out-of-range level-set dof index
    bad = M.FunctionSpace(mesh, 1, V.dofmap.copy(), V.num_dofs - 5, V.num_dofs - 5, 1, V.dof_coords[:-5])
    fbad = M.Function(bad, "phi", phi_a.x.array[:-5].copy())
    with pytest.raises(cfx.CfxError, match="Level-set dof index is out of range"):
        cfx.cut(fbad)
    assert np.array_equal(cfx.locate_entities(cd_a, "phi<0"), in_a)


def test_assembly_overwrites_every_entry_of_the_active_rows(built_lib):
    """cfx_create_sparsity zeroes the inactive rows' entries only and leaves the active rows' values to the first
    assembly of the pattern's own form (cfx_pattern::values_lazy).  With CFX_POISON_VALUES=1 the library fills the
    values with NaN bit patterns before that, so an entry no kernel writes would surface in the comparison with the
    oracle that smoke() makes."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CFX_POISON_VALUES="1")
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout
