"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py -- see its header for
provenance: frozen oracle outputs, the reference itself cannot run here).

-m "not gpu": the oracle still reproduces them (the checker has not moved);
-m gpu:       the CUDA path, through the C ABI, reproduces them from the STORED inputs.
Bars: integer/index data bit-exact; weights 1e-12 relative; matrix/vector 1e-11 relative Frobenius.
"""
import glob
import os

import numpy as np
import pytest

from util import GpuRun, OracleRun, make_problem

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
GOLDEN = [p for p in _ALL if not os.path.basename(p).startswith("ext_")]
GOLDEN_EXT = [p for p in _ALL if os.path.basename(p).startswith("ext_")]
INT_KEYS = ("domain", "inside", "cut", "outside", "ghost", "rows4", "row_ptr", "cols", "rv_offsets", "rv_parent_map",
            "ro_offsets", "ro_parent_map", "ri_offsets", "ri_parent_map")


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _check(run, g, points_atol):
    got = dict(domain=run.domain[: g["domain"].size], inside=run.inside, cut=run.cut, outside=run.outside,
               ghost=run.ghost, rows4=run.rows4, row_ptr=run.row_ptr, cols=run.cols)
    for tag, r in (("rv", run.rv), ("ro", run.ro), ("ri", run.ri)):
        got[f"{tag}_offsets"], got[f"{tag}_parent_map"] = r.offsets, r.parent_map
        np.testing.assert_allclose(r.points, g[f"{tag}_points"], rtol=0, atol=points_atol)
        np.testing.assert_allclose(r.weights, g[f"{tag}_weights"], rtol=1e-12, atol=1e-18)
    for k in INT_KEYS:
        assert np.array_equal(np.asarray(got[k]).reshape(-1), g[k].reshape(-1)), k
    assert rel(run.vals, g["vals"]) < 1e-11
    assert rel(run.b, g["b"]) < 1e-11
    assert abs(run.volume - g["volume"]) <= 1e-12 * abs(g["volume"])
    assert abs(run.area - g["area"]) <= 1e-12 * abs(g["area"])


def test_fixtures_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    mesh, Vphi, phi, V = make_problem(str(g["kind"]), int(g["n"]), int(g["degree"]))
    # the stored inputs are what make_problem regenerates (seeded)
    assert np.array_equal(Vphi.dofmap, g["phi_dofmap"]) and np.array_equal(V.dofmap, g["dofmap"])
    assert np.array_equal(phi.x.array, g["phi"])
    _check(OracleRun(mesh, Vphi, phi, V), g, 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_reproduces_golden(path, built_lib):
    from cutfemx_b200 import mesh as M

    g = np.load(path)
    mesh, Vphi, phi, V = make_problem(str(g["kind"]), int(g["n"]), int(g["degree"]))
    Vphi.dofmap[:] = g["phi_dofmap"]
    V.dofmap[:] = g["dofmap"]
    phi = M.Function(Vphi, "phi", np.array(g["phi"]))
    run = GpuRun(mesh, Vphi, phi, V)
    _check(run, g, 1e-14)
    np.testing.assert_allclose(run.normals, g["normals"], rtol=0, atol=1e-12)


# ---------------------------------------------------------------- extended fixtures (rows a13, a14, (f) rank 3)
def _ext_problem(g):
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    inp = mg.ext_inputs(str(g["kind"]), int(g["n"]), int(g["degree"]))
    mesh, Vphi, phi, V, markers, gv, w, exterior = inp
    assert np.array_equal(Vphi.dofmap, g["phi_dofmap"]) and np.array_equal(V.dofmap, g["dofmap"])
    assert np.array_equal(phi.x.array, g["phi"]) and np.array_equal(markers, g["markers"])
    assert np.array_equal(gv, g["g"]) and np.array_equal(w, g["w"]) and np.array_equal(exterior, g["exterior"])
    return mg, inp


def test_ext_fixtures_exist():
    assert len(GOLDEN_EXT) >= 2


@pytest.mark.parametrize("path", GOLDEN_EXT, ids=[os.path.basename(p)[:-4] for p in GOLDEN_EXT])
def test_oracle_reproduces_ext_golden(path):
    g = np.load(path)
    mg, inp = _ext_problem(g)
    o = mg.ext_oracle(*inp)
    for k in ("row_ptr", "cols", "facet_codes", "fr_offsets", "fr_parent_map"):
        assert np.array_equal(o[k], g[k]), k
    for k in ("A_bc", "b_lift", "fr_points", "fr_weights"):
        assert np.array_equal(o[k], g[k]), k
    assert o["square"] == float(g["square"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN_EXT, ids=[os.path.basename(p)[:-4] for p in GOLDEN_EXT])
def test_cuda_reproduces_ext_golden(path, built_lib):
    import cutfemx_b200 as cfx

    g = np.load(path)
    _, (mesh, Vphi, phi, V, markers, gv, w, exterior) = _ext_problem(g)
    cd = cfx.cut(phi)
    inside = cfx.locate_entities(cd, "phi<0")
    rv = cfx.runtime_quadrature(cd, "phi<0", 4)
    a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", inside, rv, (1.3,))
    a.add_interior_facet_integral("ghost_grad_jump", facets=cfx.ghost_penalty_facets(cd, "phi<0"), constants=(0.1,))
    bc = cfx.fem.dirichletbc(gv * markers, np.nonzero(markers)[0], V)
    A = cfx.fem.assemble_matrix(a, bcs=[bc])
    assert np.array_equal(A.indptr, g["row_ptr"]) and np.array_equal(A.indices, g["cols"])
    assert rel(A.data, g["A_bc"]) < 1e-11
    b = np.zeros(V.num_dofs)
    cfx.fem.apply_lifting(b, [a], [[bc]], A=[A])
    assert rel(b, g["b_lift"]) < 1e-11
    M0 = cfx.fem.CutForm(V, 0).add_cell_integral("square_fn", inside, rv, (1.5,)).set_coefficient(w)
    assert abs(cfx.fem.assemble_scalar(M0) - float(g["square"])) <= 1e-12 * abs(float(g["square"]))
    cdf = cfx.cut(phi, exterior, mesh.tdim - 1)
    for sel, code in (("phi<0", 1), ("phi=0", 2), ("phi>0", 3)):
        assert np.array_equal(cfx.locate_entities(cdf, sel), exterior[g["facet_codes"] == code])
    fr = cfx.runtime_quadrature(cdf, "phi<0", 2)
    assert np.array_equal(fr.offsets, g["fr_offsets"]) and np.array_equal(fr.parent_map, g["fr_parent_map"])
    np.testing.assert_allclose(fr.points, g["fr_points"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(fr.weights, g["fr_weights"], rtol=1e-12, atol=1e-18)
