"""CPU: the overlapping-slab oracle driver of tests/test_parity_at_size.py reproduces the single-process oracle
(rows, lists, weights in global numbering), so the at-size GPU parity tests compare against the same thing the small
tests do."""
import numpy as np
import pytest

from cutfemx_b200 import mesh as M
from oracle import pipeline
from test_parity_at_size import compare_with_slabs, oracle_slabs


@pytest.mark.parametrize("shape,nparts,order", [((12, 12, 24), 3, 4), ((10, 9, 16), 4, 2), ((40, 48), 3, 4)])
def test_slabs_reproduce_the_serial_oracle(shape, nparts, order):
    tdim = len(shape)
    p0, p1 = ((0.0,) * 3, (1.0,) * 3) if tdim == 3 else ((-1.0, -1.0), (1.0, 1.0))
    mesh = M.create_box(*shape, p0, p1) if tdim == 3 else M.create_rectangle(*shape, p0, p1)
    V = M.FunctionSpace(mesh, 1, mesh.x_dofmap, mesh.num_nodes, mesh.num_nodes, 1, mesh.x)
    ls = M.sphere_level_set((0.5, 0.5, 0.5), 0.35) if tdim == 3 else M.sphere_level_set((0.0, 0.0, 0.0), 0.5)
    phi = M.interpolate(V, ls)
    ref = pipeline.run_pipeline(mesh, V.dofmap, phi, V, order=order, g_value=2.5)
    slabs = oracle_slabs(shape, p0, p1, mesh.x, phi, order, nparts=nparts)
    rv, ri = ref["rv"], ref["ri"]
    ea, eb = compare_with_slabs(slabs, ref["row_ptr"], ref["cols"], ref["vals"], ref["b"], ref["inside"], ref["cut"],
                                ref["ghost"], float(np.sum(rv.weights)), float(np.sum(ri.weights)), rv.weights.size,
                                ri.weights.size)
    assert ea < 1e-13 and eb < 1e-13
