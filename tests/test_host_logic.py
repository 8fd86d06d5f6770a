"""-m "not gpu": host-side logic of the Python mirror (no CUDA calls): selector grammar, level-set
naming, argument validation, synthetic DOLFINx-layout meshes and dofmaps."""
import numpy as np
import pytest

import oracle as O
import importlib

from cutfemx_b200 import _lib, mesh as M

cutmod = importlib.import_module("cutfemx_b200.cut")  # the package also exports a function named `cut`


def test_selector_parser_matches_reference_grammar():
    """cut.cpp:47-57 (whitespace ignored), cut.cpp:881-882 (DNF: `and` within terms, `or` between)."""
    to, cl, cr = _lib.parse_selector(" phi < 0 ", ("phi",))
    assert to.tolist() == [0, 1] and cl.tolist() == [0] and cr.tolist() == [_lib.REL["<"]]
    to, cl, cr = _lib.parse_selector("phi<=0 and phi1>0 or phi1=0", ("phi", "phi1"))
    assert to.tolist() == [0, 2, 3] and cl.tolist() == [0, 1, 1]
    assert cr.tolist() == [_lib.REL["<="], _lib.REL[">"], _lib.REL["="]]
    to, cl, cr = _lib.parse_selector("phi>=0.0", ("phi",))
    assert cr.tolist() == [_lib.REL[">="]]
    # names that contain the keywords
    to, cl, cr = _lib.parse_selector("band<0 and floor>0", ("band", "floor"))
    assert cl.tolist() == [0, 1]
    for bad in ("", "phi", "phi<1", "psi<0", "phi<0 xor phi>0", "phi<0 and", "<0"):
        with pytest.raises(ValueError):
            _lib.parse_selector(bad, ("phi",))
    # the oracle's independent parser agrees
    for expr in ("phi<0", "phi<=0 and phi1>0 or phi1=0", "phi=0 or phi1>=0"):
        a = _lib.parse_selector(expr, ("phi", "phi1"))
        b = O.parse_selector(expr, ("phi", "phi1"))
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_level_set_names_are_frozen_like_the_reference():
    """cut.cpp:59-62,81-137 / test_cut_api.py:750-763: unnamed ('', 'u', 'f') -> phi, phi1, phi2."""
    mesh = M.create_rectangle(2, 2)
    V = M.functionspace(mesh, 1)
    fs = [M.Function(V, n) for n in ("f", "u", "", "levelset")]
    assert cutmod._freeze_names(fs) == ("phi", "phi1", "phi2", "levelset")
    with pytest.raises(ValueError):
        cutmod._freeze_names([M.Function(V, "a"), M.Function(V, "a")])
    with pytest.raises(TypeError):
        cutmod._normalise_level_sets("phi")
    with pytest.raises(ValueError):
        cutmod._normalise_level_sets([])
    with pytest.raises(TypeError):
        cutmod._normalise_level_sets([V])


def test_cut_argument_validation_before_any_device_work():
    mesh = M.create_rectangle(2, 2)
    other = M.create_rectangle(2, 2)
    V, W = M.functionspace(mesh, 1), M.functionspace(other, 1)
    with pytest.raises(ValueError):  # cut.cpp:462-498: same mesh for every level set
        cutmod.cut([M.Function(V, "a"), M.Function(W, "b")])
    with pytest.raises(ValueError, match="entity_dim must be supplied"):  # python/cutfemx/cut.py:157-158
        cutmod.cut(M.Function(V, "a"), entities=np.arange(3))
    with pytest.raises(ValueError, match="entity_dim is only valid"):  # python/cutfemx/cut.py:153-154
        cutmod.cut(M.Function(V, "a"), entity_dim=1)
    with pytest.raises(ValueError, match="positive-dimensional"):  # cut.cpp:545-550
        cutmod.cut(M.Function(V, "a"), entities=np.arange(3), entity_dim=0)
    Vv = M.functionspace(mesh, 1, bs=2)
    with pytest.raises(ValueError):  # cut.cpp:444-460: scalar Lagrange
        cutmod.cut(M.Function(Vv, "a"))


@pytest.mark.parametrize("tdim", [2, 3])
def test_mesh_generator_is_a_valid_dolfinx_layout(tdim):
    mesh = M.create_rectangle(5, 4) if tdim == 2 else M.create_box(3, 4, 2)
    nv = tdim + 1
    assert mesh.x.shape[1] == 3 and mesh.x.dtype == np.float64  # cut.cpp:529: stride 3 even in 2D
    assert mesh.x_dofmap.dtype == np.int32 and mesh.x_dofmap.shape[1] == nv
    X = mesh.x[mesh.x_dofmap][:, :, :tdim]
    vol = np.abs(np.linalg.det((X[:, 1:] - X[:, :1]).transpose(0, 2, 1)))
    assert vol.min() > 0
    size = np.prod(np.asarray(mesh.p1) - np.asarray(mesh.p0))
    np.testing.assert_allclose(vol.sum() / (2 if tdim == 2 else 6), size, rtol=1e-13)
    # facet i is opposite local vertex i; every facet has 1 or 2 cells; shared facets share vertices
    ncell = np.diff(mesh.f2c_offsets)
    used = np.unique(mesh.c2f)
    assert set(ncell[used]) <= {1, 2} and np.all(ncell[np.setdiff1d(np.arange(mesh.num_facets), used)] == 0)
    for f in used[ncell[used] == 2][:200]:
        c0, c1 = mesh.f2c[mesh.f2c_offsets[f]:mesh.f2c_offsets[f] + 2]
        v0 = set(np.delete(mesh.x_dofmap[c0], np.nonzero(mesh.c2f[c0] == f)[0][0]))
        v1 = set(np.delete(mesh.x_dofmap[c1], np.nonzero(mesh.c2f[c1] == f)[0][0]))
        assert v0 == v1 and c0 < c1
    n_int = (ncell == 2).sum()
    n_bnd = (ncell == 1).sum()
    assert nv * mesh.num_cells == 2 * n_int + n_bnd


@pytest.mark.parametrize("tdim", [2, 3])
def test_p2_dofmap(tdim):
    mesh = M.create_rectangle(3, 3) if tdim == 2 else M.create_box(2, 2, 2)
    V = M.functionspace(mesh, 2, permute_seed=4)
    assert V.nd == (6 if tdim == 2 else 10)
    edges = M.TRI_EDGES if tdim == 2 else M.TET_EDGES
    # edge dofs sit at the midpoint of the edge's two vertices (Basix sub-entity numbering)
    for c in range(0, mesh.num_cells, 3):
        for e, (a, b) in enumerate(edges):
            mid = 0.5 * (mesh.x[mesh.x_dofmap[c, a]] + mesh.x[mesh.x_dofmap[c, b]])
            np.testing.assert_allclose(V.dof_coords[V.dofmap[c, tdim + 1 + e]], mid, atol=1e-15)
        np.testing.assert_allclose(V.dof_coords[V.dofmap[c, : tdim + 1]], mesh.x[mesh.x_dofmap[c]], atol=1e-15)
    assert np.unique(V.dofmap).size == V.num_dofs


def test_bench_host_binding_without_gpu_or_numa_information():
    """bench.py binds a rank to its GPU's NUMA node where the machine exposes one; without a GPU (here) or without
    sysfs NUMA entries (the VMs the benchmark ran on) it reports that and leaves the affinity alone."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_numa_node(0)
    assert isinstance(info, dict) and info.get("numa_node") is None
    assert os.sched_getaffinity(0) == before
    # every workload names the configuration of BASELINE.json it stands for
    assert set(bench.WORKLOADS) >= {"C1", "C2", "C3", "C4", "C5"}
    assert bench.WORKLOADS["C3"]["n"] == 256 and bench.WORKLOADS["C4"]["n"] == 192 and bench.WORKLOADS["C2"]["n"] == 4096
    assert bench.reference_parts(256, 1) >= 1
