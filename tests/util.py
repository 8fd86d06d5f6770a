"""Shared problem set-ups: the same seeded inputs go to the oracle and to the CUDA path."""
from __future__ import annotations

import numpy as np

import oracle as O
from cutfemx_b200 import mesh as M


def make_problem(kind: str, n: int, degree: int = 1, permute_seed: int | None = 7, shift=0.0):
    """kind: 'circle' (C1-like, [-1,1]^2, R=0.5), 'sphere' (C3-like, [0,1]^3, R=0.35), 'torus' (C4-like),
    'line' (phi = x - 0.51, test_cut_api.py:19-33)."""
    if kind in ("circle", "line"):
        mesh = M.create_rectangle(n, n, (-1.0, -1.0), (1.0, 1.0)) if kind == "circle" else \
            M.create_rectangle(n, n, (0.0, 0.0), (1.0, 1.0))
        ls = M.sphere_level_set((shift, 0.0, 0.0), 0.5) if kind == "circle" else (lambda x, y, z: x - 0.51)
    elif kind == "sphere":
        mesh = M.create_box(n, n, n)
        ls = M.sphere_level_set((0.5 + shift, 0.5, 0.5), 0.35)
    elif kind == "torus":
        mesh = M.create_box(n, n, n)
        ls = M.torus_level_set((0.5, 0.5, 0.5), 0.3, 0.12)
    else:
        raise ValueError(kind)
    Vphi = M.functionspace(mesh, 1, permute_seed=permute_seed)
    phi = M.Function(Vphi, "phi").interpolate(ls)
    V = M.functionspace(mesh, degree, permute_seed=None if permute_seed is None else permute_seed + 1)
    return mesh, Vphi, phi, V


class OracleRun:
    """The reference pipeline of demo_poisson.py:156-201 through the CPU oracle."""

    def __init__(self, mesh, Vphi, phi, V, order=4, gamma=40.0, gamma_g=0.1, g_value=2.5, f_value=1.0):
        self.mesh, self.V = mesh, V
        vals = phi.x.array
        self.domain = O.classify(Vphi.dofmap, vals)
        self.inside = O.locate(self.domain, "phi<0")
        self.cut = O.locate(self.domain, "phi=0")
        self.outside = O.locate(self.domain, "phi>0")
        self.rv = O.runtime_quadrature(mesh, Vphi.dofmap, vals, self.domain, "<", order)
        self.ro = O.runtime_quadrature(mesh, Vphi.dofmap, vals, self.domain, ">", order)
        self.ri = O.runtime_quadrature(mesh, Vphi.dofmap, vals, self.domain, "=", order)
        self.ri.normals = O.normals(mesh, Vphi.dofmap, 1, vals, self.ri)
        self.ghost = O.ghost_penalty_facets(mesh, self.cut, self.inside)
        self.rows4 = O.facet_rows(mesh, self.ghost)
        self.active = np.concatenate([self.inside, self.rv.parent_map])
        self.row_ptr, self.cols = O.sparsity(V, self.active, self.rows4)
        self.vals = np.zeros(self.cols.size)
        O.assemble_cells(V, "laplace", self.vals, self.inside, self.rv, (1.0,), self.row_ptr, self.cols)
        O.assemble_cells(V, "nitsche", self.vals, None, self.ri, (gamma,), self.row_ptr, self.cols)
        O.assemble_interior_facets(V, "ghost_grad_jump", self.vals, self.rows4, (gamma_g,), self.row_ptr, self.cols)
        self.b = np.zeros(V.num_dofs)
        O.assemble_cells(V, "source", self.b, self.inside, self.rv, (f_value,))
        O.assemble_cells(V, "nitsche_rhs", self.b, None, self.ri, (gamma, g_value))
        m = np.zeros(1)
        O.assemble_cells(V, "one", m, self.inside, self.rv, (1.0,))
        self.volume = m[0]
        m = np.zeros(1)
        O.assemble_cells(V, "one", m, None, self.ri, (1.0,))
        self.area = m[0]


class GpuRun:
    """The same pipeline through cutfemx_b200 (C ABI -> sm_100a kernels)."""

    def __init__(self, mesh, Vphi, phi, V, order=4, gamma=40.0, gamma_g=0.1, g_value=2.5, f_value=1.0):
        import cutfemx_b200 as cfx

        self.cut_data = cd = cfx.cut(phi)
        self.domain = cd.domain_codes()
        self.inside = cfx.locate_entities(cd, "phi<0")
        self.cut = cfx.locate_entities(cd, "phi=0")
        self.outside = cfx.locate_entities(cd, "phi>0")
        self.rv = cfx.runtime_quadrature(cd, "phi<0", order)
        self.ro = cfx.runtime_quadrature(cd, "phi>0", order)
        self.ri = cfx.runtime_quadrature(cd, "phi=0", order)
        self.normals = cfx.normal(cd, phi, self.ri)
        self.ghost = cfx.ghost_penalty_facets(cd, "phi<0")
        self.rows4 = cfx.facet_integration_rows(mesh, self.ghost)
        a = cfx.fem.CutForm(V, 2)
        a.add_cell_integral("laplace", self.inside, self.rv, (1.0,))
        a.add_cell_integral("nitsche", None, self.ri, (gamma,))
        a.add_interior_facet_integral("ghost_grad_jump", facets=self.ghost, constants=(gamma_g,))
        self.a = a
        self.A = cfx.fem.assemble_matrix(a)
        self.row_ptr, self.cols, self.vals = self.A.indptr, self.A.indices, self.A.data
        L = cfx.fem.CutForm(V, 1)
        L.add_cell_integral("source", self.inside, self.rv, (f_value,))
        L.add_cell_integral("nitsche_rhs", None, self.ri, (gamma, g_value))
        self.L = L
        self.b = cfx.fem.assemble_vector(L)
        Mv = cfx.fem.CutForm(V, 0).add_cell_integral("one", self.inside, self.rv, (1.0,))
        self.volume = cfx.fem.assemble_scalar(Mv)
        Ma = cfx.fem.CutForm(V, 0).add_cell_integral("one", None, self.ri, (1.0,))
        self.area = cfx.fem.assemble_scalar(Ma)
