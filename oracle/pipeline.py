"""demo_poisson.py:156-201 through the CPU oracle, with per-stage wall times.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product.
Label for every number it produces: "CPU restatement of reference loops -- reference binary
unavailable" (BASELINE.md section 3).
"""
from __future__ import annotations

import time

import numpy as np

from . import oracle as O


def run_pipeline(mesh, Vphi_dofmap, phi_vals, V, order=4, gamma=40.0, gamma_g=0.1, f_value=1.0, g_value=0.0,
                 with_rhs=True):
    t = {}
    t0 = time.perf_counter()
    domain = O.classify(Vphi_dofmap, phi_vals, mesh.num_cells_local)
    t["classify"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    inside = O.locate(domain, "phi<0")
    cut = O.locate(domain, "phi=0")
    t["locate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    rv = O.runtime_quadrature(mesh, Vphi_dofmap, phi_vals, domain, "<", order)
    ri = O.runtime_quadrature(mesh, Vphi_dofmap, phi_vals, domain, "=", order)
    t["quadrature"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ri.normals = O.normals(mesh, Vphi_dofmap, 1, phi_vals, ri)
    t["normals"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ghost = O.ghost_penalty_facets(mesh, cut, inside)
    rows4 = O.facet_rows(mesh, ghost)
    t["ghost_facets"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    active = np.concatenate([inside, rv.parent_map])
    row_ptr, cols = O.sparsity(V, active, rows4)
    t["sparsity"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    vals = np.zeros(cols.size)
    O.assemble_cells(V, "laplace", vals, inside, rv, (1.0,), row_ptr, cols)
    O.assemble_cells(V, "nitsche", vals, None, ri, (gamma,), row_ptr, cols)
    O.assemble_interior_facets(V, "ghost_grad_jump", vals, rows4, (gamma_g,), row_ptr, cols)
    t["assemble_matrix"] = time.perf_counter() - t0
    b = None
    if with_rhs:
        t0 = time.perf_counter()
        b = np.zeros(V.num_dofs)
        O.assemble_cells(V, "source", b, inside, rv, (f_value,))
        O.assemble_cells(V, "nitsche_rhs", b, None, ri, (gamma, g_value))
        t["assemble_vector"] = time.perf_counter() - t0
    out = dict(domain=domain, inside=inside, cut=cut, rv=rv, ri=ri, ghost=ghost, rows4=rows4, row_ptr=row_ptr,
               cols=cols, vals=vals, b=b, times=t, total_s=sum(t.values()))
    return out


def run_elasticity_pipeline(mesh, Vphi_dofmap, phi_vals, V, order=4, E=1.0e3, nu=0.3, gamma=40.0, gamma_g=0.05,
                            force=(0.0, 0.0, -1.0)):
    """python/demo/demo_elasticity.py:213-238 through the oracle: vector space V (block size gdim), symmetric Nitsche
    on the interface rules, componentwise ghost penalty, constant body force."""
    mu = E / (2.0 * (1.0 + nu))
    lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    bs = V.bs
    t = {}
    t0 = time.perf_counter()
    domain = O.classify(Vphi_dofmap, phi_vals, mesh.num_cells_local)
    t["classify"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    inside = O.locate(domain, "phi<0")
    cut = O.locate(domain, "phi=0")
    t["locate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    rv = O.runtime_quadrature(mesh, Vphi_dofmap, phi_vals, domain, "<", order)
    ri = O.runtime_quadrature(mesh, Vphi_dofmap, phi_vals, domain, "=", order)
    t["quadrature"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ri.normals = O.normals(mesh, Vphi_dofmap, 1, phi_vals, ri)
    t["normals"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ghost = O.ghost_penalty_facets(mesh, cut, inside)
    rows4 = O.facet_rows(mesh, ghost)
    t["ghost_facets"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    active = np.concatenate([inside, rv.parent_map])
    row_ptr, cols = O.sparsity(V, active, rows4)
    t["sparsity"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    vals = np.zeros(cols.size * bs * bs)
    O.assemble_cells(V, "elasticity", vals, inside, rv, (mu, lam), row_ptr, cols)
    O.assemble_cells(V, "nitsche_vec", vals, None, ri, (mu, lam, gamma), row_ptr, cols)
    O.assemble_interior_facets(V, "ghost_grad_jump", vals, rows4, (gamma_g * (2.0 * mu + lam),), row_ptr, cols)
    t["assemble_matrix"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = np.zeros(V.num_dofs * bs)
    O.assemble_cells(V, "source_vec", b, inside, rv, tuple(force)[:bs])
    t["assemble_vector"] = time.perf_counter() - t0
    return dict(domain=domain, inside=inside, cut=cut, rv=rv, ri=ri, ghost=ghost, rows4=rows4, row_ptr=row_ptr,
                cols=cols, vals=vals, b=b, times=t, total_s=sum(t.values()))
