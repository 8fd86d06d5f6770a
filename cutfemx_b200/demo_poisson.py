"""The cut Poisson pipeline of the reference's demo (python/demo/demo_poisson.py:156-201) written
against this package's public API, device-resident where the reference holds numpy arrays.

One `step()` is one pass of the whole hot path:
  update (classify) -> locate inside cells -> volume / interface run-time quadrature -> normals
  -> ghost-penalty facets + integration rows -> sparsity -> matrix + vector assembly.
bench.py times it, __graft_entry__.smoke() runs it once, tests compare it with the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import importlib

_cut = importlib.import_module(__package__ + ".cut")  # the package exports a function named `cut` too
from . import fem as _fem
from . import level_set as _ls
from ._lib import DEVICE, HOST, check, lib
from .cut import Context
from .mesh import TETRAHEDRON, TRIANGLE, Function, FunctionSpace, Mesh


def device_mesh(ctx_device: int, shape, p0, p1):
    """Generate a Kuhn box / right-diagonal rectangle mesh on the GPU (csrc/meshgen.cu)."""
    import torch

    dev = f"cuda:{ctx_device}"
    tdim = len(shape)
    nn = int(np.prod([s + 1 for s in shape]))
    nc = int(np.prod(shape)) * (2 if tdim == 2 else 6)
    nv = tdim + 1
    x = torch.empty((nn, 3), dtype=torch.float64, device=dev)
    x_dofmap = torch.empty((nc, nv), dtype=torch.int32, device=dev)
    c2f = torch.empty((nc, nv), dtype=torch.int32, device=dev)
    nfac = (3 if tdim == 2 else 12) * nn
    mesh = Mesh(TRIANGLE if tdim == 2 else TETRAHEDRON, tdim, tdim, x, x_dofmap, c2f, None, None, nc, nfac, nfac,
                shape=tuple(shape), p0=tuple(p0), p1=tuple(p1))
    mesh.extra["device"] = ctx_device
    # generate into the tensors with a scratch context, then bind them
    gen = Context(ctx_device)
    P0 = (C.c_double * 3)(*(list(p0) + [0.0] * 3)[:3])
    P1 = (C.c_double * 3)(*(list(p1) + [0.0] * 3)[:3])
    if tdim == 3:
        check(gen.handle, lib().cfx_meshgen_box(gen.handle, shape[0], shape[1], shape[2], P0, P1,
                                                C.c_void_p(x.data_ptr()), C.c_void_p(x_dofmap.data_ptr()),
                                                C.c_void_p(c2f.data_ptr())))
    else:
        check(gen.handle, lib().cfx_meshgen_rectangle(gen.handle, shape[0], shape[1], P0, P1,
                                                      C.c_void_p(x.data_ptr()), C.c_void_p(x_dofmap.data_ptr()),
                                                      C.c_void_p(c2f.data_ptr())))
    gen.sync()
    gen.close()
    return mesh


def device_p2_tet_space(mesh: Mesh, bs: int = 1) -> FunctionSpace:
    """P2 Lagrange space (block size `bs`) on a device_mesh() box: vertex dofs + one dof per edge, generated on the
    GPU (csrc/meshgen.cu p2_tet_dofmap_kernel; sparse edge numbering, unused ids at the box boundary)."""
    import torch

    assert mesh.tdim == 3 and mesh.shape is not None
    ctx = _cut._mesh_context(mesh)
    nc = int(mesh.x_dofmap.shape[0])
    dm = torch.empty((nc, 10), dtype=torch.int32, device=mesh.x.device)
    nd = C.c_int64(0)
    check(ctx.handle, lib().cfx_meshgen_p2_tet_dofmap(ctx.handle, mesh.shape[0], mesh.shape[1], mesh.shape[2],
                                                      C.c_void_p(mesh.x_dofmap.data_ptr()), C.c_int64(nc),
                                                      C.c_void_p(dm.data_ptr()), C.byref(nd)))
    return FunctionSpace(mesh, 2, dm, int(nd.value), int(nd.value), bs, None)


def device_level_set(mesh: Mesh, kind: str, params, out=None):
    """Nodal interpolation of a sphere/circle (c, R) or torus (c, R, r) level set on the GPU
    (into `out` if given)."""
    import torch

    ctx = _cut._mesh_context(mesh)
    vals = torch.empty(mesh.x.shape[0], dtype=torch.float64, device=mesh.x.device) if out is None else out
    p = (C.c_double * 5)(*(list(params) + [0.0] * 5)[:5])
    check(ctx.handle, lib().cfx_meshgen_level_set(ctx.handle, C.c_void_p(mesh.x.data_ptr()),
                                                  C.c_int64(mesh.x.shape[0]), 0 if kind == "sphere" else 1, p,
                                                  C.c_void_p(vals.data_ptr())))
    return vals


class CutPoisson:
    """demo_poisson.py:156-201 with constant source f and constant Dirichlet value g."""

    def __init__(self, mesh: Mesh, phi: Function, V: FunctionSpace, order: int = 4, gamma: float = 40.0,
                 gamma_g: float = 0.1, f_value: float = 1.0, g_value: float = 0.0):
        self.mesh, self.phi, self.V = mesh, phi, V
        self.order, self.gamma, self.gamma_g, self.f_value, self.g_value = order, gamma, gamma_g, f_value, g_value
        self.cut_data = _cut.cut(phi)
        self.ctx = self.cut_data._ctx
        self.A = None
        self.b = None
        self.stats = {}
        self.last = None
        self.fused = True  # matrix and right-hand side in one pass (cfx_assemble_system)
        # persistent = True: the result objects of a step (lists, rules, matrix) are refilled in place by the next
        # step instead of being freed and recreated -- what a time loop does (demo_moving_poisson.py:69-107), and
        # what deferred-size mode and graph capture need (cfx_set_deferred / cfx_graph_begin)
        self.persistent = False
        self.keep = {}
        self.graph = None
        # lanes of the three independent branches of build_forms (Context.set_lanes(False) / CFX_NO_LANES=1: serial)
        self.lanes = (1, 2, 3)

    def build_forms(self, assemble_rhs: bool = True):
        """update -> locate -> rules -> normals -> ghost facets -> the forms a and L."""
        cd = self.cut_data
        k = self.keep if self.persistent else {}
        ln = self.lanes
        _cut.update(cd)                                                     # cutfemx.update
        # the four results below depend on the classification only: each on a lane (stream) of its own
        with self.ctx.lane(ln[0]):
            rv = _cut.runtime_quadrature(cd, "phi<0", self.order, out=k.get("rv"))          # volume rules
        with self.ctx.lane(ln[1]):
            ri = _cut.runtime_quadrature(cd, "phi=0", self.order, out=k.get("ri"))          # interface rules
            _ls.attach_normal(cd, self.phi, ri)                             # n = normal(phi)
        with self.ctx.lane(ln[2]):
            ghost = _cut.ghost_penalty_facets_device(cd, "phi<0", out=k.get("ghost"))       # ghost_penalty_facets
            rows = _cut.facet_integration_rows_device(self.mesh, ghost, out=k.get("rows"))  # facet_integration_rows
        inside = _cut.locate_entities_device(cd, "phi<0", out=k.get("inside"))          # locate_entities
        self.ctx.join()
        a = _fem.CutForm(self.V, 2)
        a.add_cell_integral("laplace", inside, rv, (1.0,))
        a.add_cell_integral("nitsche", None, ri, (self.gamma,))
        # (a persistent loop never asks for the size: in deferred-size mode it is not on the host)
        if self.persistent or ghost.size > 0:
            a.add_interior_facet_integral("ghost_grad_jump", rows=rows, constants=(self.gamma_g,))
        L = None
        if assemble_rhs:
            L = _fem.CutForm(self.V, 1)
            L.add_cell_integral("source", inside, rv, (self.f_value,))
            L.add_cell_integral("nitsche_rhs", None, ri, (self.gamma, self.g_value))
        self.last = dict(inside=inside, rv=rv, ri=ri, ghost=ghost, rows=rows, a=a, L=L)
        if self.persistent:
            self.keep = dict(inside=inside, rv=rv, ri=ri, ghost=ghost, rows=rows)
        return a, L

    def assemble(self):
        """create_sparsity_pattern + assemble_matrix (+ assemble_vector) of the current forms."""
        a, L = self.last["a"], self.last["L"]
        self.A = _fem.create_matrix(a, self.A)                              # create_sparsity_pattern
        if L is not None and self.fused:
            import torch

            if self.b is None or not hasattr(self.b, "data_ptr"):
                self.b = torch.empty(self.V.num_dofs * self.V.bs, dtype=torch.float64,
                                     device=f"cuda:{self.ctx.device}")
            _fem.assemble_system(a, self.A, L, self.b)                      # assemble_matrix + assemble_vector
        else:
            _fem.assemble_matrix(a, self.A)                                 # assemble_matrix
            if L is not None:
                self.b = self._assemble_vector_device(L)
        if self.persistent:
            return None  # sizes stay where they are; fetch_stats() asks for them
        return self.fetch_stats()

    def fetch_stats(self):
        """Sizes of the current step's results (synchronises in deferred-size mode)."""
        cd, t = self.cut_data, self.last if self.last else self.keep
        counts = cd.counts()
        self.stats = dict(inside=counts[0], cut=counts[1], outside=counts[2], nnz=self.A.nnz,
                          n_rows=self.A.shape[0], ghost_facets=t["ghost"].size, volume_points=t["rv"].total_points,
                          interface_points=t["ri"].total_points, volume_rules=t["rv"].num_rules)
        return self.stats

    def release_step(self):
        t, self.last = self.last, None
        if not t:
            return
        t["a"].free()
        if t["L"] is not None:
            t["L"].free()
        if self.persistent:
            return  # the lists, rules and the matrix are refilled in place by the next step
        for k in ("inside", "rv", "ri", "ghost", "rows"):
            t[k].free()

    def step(self, assemble_rhs: bool = True, keep: bool = False):
        self.build_forms(assemble_rhs)
        stats = self.assemble()
        if not keep:
            self.release_step()
        return stats

    # ---- time loops: persistent objects, deferred sizes, one graph launch per step
    def capture(self, margin: float = 0.25):
        """Record one step as a CUDA graph.  Runs two eager steps first (objects exist, buffers get `margin` spare
        capacity), one deferred-size step (nothing needs a size on the host any more), then captures."""
        self.persistent = True
        ctx = self.ctx
        ctx.set_deferred(False, margin)
        for _ in range(2):
            self.step()
        ctx.set_deferred(True)
        self.step()
        ctx.check()
        ctx.graph_begin()
        try:
            self.step()
        finally:
            self.graph = ctx.graph_end()
        return self.graph

    def replay(self):
        """One time step = one graph launch (the level-set values are re-read from the bound array)."""
        self.graph.launch()

    def _assemble_vector_device(self, L):
        import torch

        n = self.V.num_dofs * self.V.bs
        if self.b is None or not hasattr(self.b, "data_ptr"):
            self.b = torch.empty(n, dtype=torch.float64, device=f"cuda:{self.ctx.device}")
        check(self.ctx.handle, lib().cfx_assemble_vector(self.ctx.handle, L._h, C.c_void_p(self.b.data_ptr()), 1, DEVICE))
        return self.b
