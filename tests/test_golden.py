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

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
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
