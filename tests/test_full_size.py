"""-m gpu: BASELINE.json's full-size configuration (3D sphere, 256^3 Kuhn tetrahedra, P1) through
size-independent properties -- the oracle would need minutes there, the properties need none:

* partition of unity: sum(b) for L = f*v equals f * (volume of the cut domain) to 1e-12 (the reference
  pins the same identity at small size, python/tests/test_cut_api.py:858-869);
* the volume itself: inside cells + run-time weights = assemble_scalar(1*dx), and within the O(h^2)
  geometric error of the exact ball; the interface measure likewise against 4 pi R^2;
* constants are in the kernel of the Laplace form: row sums of the Laplace-only matrix vanish;
* the full Nitsche + ghost-penalty matrix is symmetric (checked at 128^3 where the host can hold A^T);
* classification counts add up and equal SURVEY.md section 8's sizing (691 620 cut cells);
* two assemblies are bit-identical.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(n, device=0):
    import torch

    import cutfemx_b200 as cfx
    from cutfemx_b200 import demo_poisson as dp
    from cutfemx_b200.mesh import Function, FunctionSpace

    mesh = dp.device_mesh(device, [n] * 3, [0.0] * 3, [1.0] * 3)
    vals = dp.device_level_set(mesh, "sphere", (0.5, 0.5, 0.5, 0.35, 0.0))
    V = FunctionSpace(mesh, 1, mesh.x_dofmap, int(mesh.x.shape[0]), int(mesh.x.shape[0]), 1, None)
    phi = Function(V, "phi", vals)
    return cfx, torch, mesh, V, phi


@pytest.mark.parametrize("n", [256])
def test_full_size_properties(n, built_lib):
    cfx, torch, mesh, V, phi = build(n)
    from cutfemx_b200 import cut as _c  # noqa: F401
    import importlib

    cutm = importlib.import_module("cutfemx_b200.cut")
    cd = cfx.cut(phi)
    n_in, n_cut, n_out = cd.counts()
    assert n_in + n_cut + n_out == 6 * n ** 3
    if n == 256:
        assert (n_in, n_cut) == (17738688, 691620)  # SURVEY.md section 8 sizing table, config C3
    inside = cutm.locate_entities_device(cd, "phi<0")
    assert inside.size == n_in
    rv = cfx.runtime_quadrature(cd, "phi<0", 4)
    ri = cfx.runtime_quadrature(cd, "phi=0", 4)
    ro = cfx.runtime_quadrature(cd, "phi>0", 4)
    cfx.level_set.attach_normal(cd, phi, ri)
    # volumes: inside + cut part; complementary parts fill the cube (P12 at full size)
    vol = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", inside, rv, (1.0,)))
    outside = cutm.locate_entities_device(cd, "phi>0")
    vol_out = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", outside, ro, (1.0,)))
    assert abs(vol + vol_out - 1.0) < 1e-12
    exact = 4.0 / 3.0 * math.pi * 0.35 ** 3
    h = 1.0 / n
    assert 0 < exact - vol < 6 * h * h          # the piecewise-linear interface is inscribed: O(h^2) deficit
    area = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", None, ri, (1.0,)))
    assert abs(area - 4 * math.pi * 0.35 ** 2) < 20 * h * h
    # partition of unity
    f = 1.75
    L = cfx.fem.CutForm(V, 1).add_cell_integral("source", inside, rv, (f,))
    b = torch.zeros(V.num_dofs, dtype=torch.float64, device=mesh.x.device)
    cfx.fem.assemble_vector(L, b)
    sb = float(b.sum(dtype=torch.float64))
    assert abs(sb - f * vol) <= 1e-11 * abs(f * vol)
    # Laplace only: row sums vanish (relative to the row's absolute sum)
    a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", inside, rv, (1.0,))
    A = cfx.fem.assemble_matrix(a)
    vals, rp = A.values_device(), A.indptr_device()
    rows = torch.repeat_interleave(torch.arange(V.num_dofs, device=vals.device), rp[1:] - rp[:-1],
                                   output_size=int(A.nnz))
    rs = torch.zeros(V.num_dofs, dtype=torch.float64, device=vals.device).index_add_(0, rows, vals)
    ra = torch.zeros(V.num_dofs, dtype=torch.float64, device=vals.device).index_add_(0, rows, vals.abs())
    assert float((rs.abs() / ra.clamp(min=1e-300)).max()) < 1e-11
    # determinism at full size
    A2 = cfx.fem.assemble_matrix(a)
    assert torch.equal(A2.values_device(), vals) and torch.equal(A2.indices_device(), A.indices_device())


def test_symmetry_128(built_lib):
    import scipy.sparse as sp

    cfx, torch, mesh, V, phi = build(128)
    from cutfemx_b200 import demo_poisson as dp

    prob = dp.CutPoisson(mesh, phi, V, order=4)
    prob.step()
    A = prob.A
    M = sp.csr_matrix((A.data, A.indices, A.indptr), shape=A.shape)
    D = (M - M.T).tocsr()
    assert abs(D).max() <= 1e-12 * abs(M).max()
    assert (M.indptr != M.T.tocsr().indptr).sum() == 0  # structurally symmetric pattern


def test_c2_p2_4096_properties(built_lib):
    """configs[1]: 2D circle R=0.5 on 4096x4096 triangles, P2 u / P1 level set (SURVEY.md sizing: 6 581 196
    inside, 13 994 cut cells).  Properties: partition of unity for P2 (sum b = f * area), area and perimeter
    within the O(h^2) geometric error, Laplace row sums vanish."""
    import torch

    import cutfemx_b200 as cfx
    from cutfemx_b200 import parallel as P
    import importlib

    cutm = importlib.import_module("cutfemx_b200.cut")
    n = 4096
    pipe = P.RankPipeline([n, n], [-1.0, -1.0], [1.0, 1.0], 1, 0, 0, "sphere", (0.0, 0.0, 0.0, 0.5, 0.0), order=4,
                          degree=2)
    V, phi, mesh = pipe.V, pipe.phi, pipe.mesh
    assert V.nd == 6
    cd = pipe.prob.cut_data
    n_in, n_cut, n_out = cd.counts()
    assert (n_in, n_cut) == (6581196, 13994) and n_in + n_cut + n_out == 2 * n * n
    inside = cutm.locate_entities_device(cd, "phi<0")
    rv = cfx.runtime_quadrature(cd, "phi<0", 4)
    ri = cfx.runtime_quadrature(cd, "phi=0", 4)
    area = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", inside, rv, (1.0,)))
    per = cfx.fem.assemble_scalar(cfx.fem.CutForm(V, 0).add_cell_integral("one", None, ri, (1.0,)))
    h = 2.0 / n
    assert 0 < math.pi * 0.25 - area < 2 * h * h and 0 < math.pi - per < 4 * h * h
    f = 0.6
    L = cfx.fem.CutForm(V, 1).add_cell_integral("source", inside, rv, (f,))
    b = torch.zeros(V.num_dofs, dtype=torch.float64, device=mesh.x.device)
    cfx.fem.assemble_vector(L, b)
    assert abs(float(b.sum(dtype=torch.float64)) - f * area) <= 1e-11 * f * area
    a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", inside, rv, (1.0,))
    A = cfx.fem.assemble_matrix(a)
    vals, rp = A.values_device(), A.indptr_device()
    rows = torch.repeat_interleave(torch.arange(V.num_dofs, device=vals.device), rp[1:] - rp[:-1],
                                   output_size=int(A.nnz))
    rs = torch.zeros(V.num_dofs, dtype=torch.float64, device=vals.device).index_add_(0, rows, vals)
    ra = torch.zeros(V.num_dofs, dtype=torch.float64, device=vals.device).index_add_(0, rows, vals.abs())
    assert float((rs.abs() / ra.clamp(min=1e-300)).max()) < 1e-10
