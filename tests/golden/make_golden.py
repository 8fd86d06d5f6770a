#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz.

    python tests/golden/make_golden.py

PROVENANCE.  The reference (sclaus2/CutFEMx) cannot be imported or compiled in this image (it needs
DOLFINx 0.11, CutCells, runintgen, Basix, MPI -- DESIGN.md "Oracle"), and its own tests hold no
per-point vectors for this path, so these fixtures are NOT outputs of the reference: they are
outputs of the CPU oracle (oracle/cutfem_oracle.cpp) after it passed tests/test_oracle_pins.py
(the reference's own test assertions P1-P12 + analytic checks).  They freeze the oracle so that
(a) a later edit of the oracle cannot silently move the target and (b) the -m gpu tests have a
fixed target that travels to the GPU box.  Inputs are stored too, so the fixtures are
self-contained.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from util import OracleRun, make_problem  # noqa: E402

CASES = [("line", 3, 1), ("circle", 16, 1), ("circle", 8, 2), ("sphere", 6, 1), ("sphere", 4, 2), ("torus", 8, 1)]


def pack(kind, n, deg):
    mesh, Vphi, phi, V = make_problem(kind, n, deg)
    o = OracleRun(mesh, Vphi, phi, V)
    d = dict(kind=kind, n=n, degree=deg, phi_dofmap=Vphi.dofmap, phi=phi.x.array, dofmap=V.dofmap,
             num_dofs=V.num_dofs, domain=o.domain, inside=o.inside, cut=o.cut, outside=o.outside, ghost=o.ghost,
             rows4=o.rows4, row_ptr=o.row_ptr, cols=o.cols, vals=o.vals, b=o.b, volume=o.volume, area=o.area,
             normals=o.ri.normals)
    for tag, r in (("rv", o.rv), ("ro", o.ro), ("ri", o.ri)):
        d.update({f"{tag}_points": r.points, f"{tag}_weights": r.weights, f"{tag}_offsets": r.offsets,
                  f"{tag}_parent_map": r.parent_map})
    return d


# ---- second family: Dirichlet conditions, Function coefficients, facet-hosted rules (rows a13, a14, (f) rank 3)
EXT_CASES = [("circle", 10, 1), ("sphere", 5, 1)]


def ext_inputs(kind, n, deg):
    """Deterministic inputs of the extended fixtures (also used by tests/test_golden.py to re-create them)."""
    mesh, Vphi, phi, V = make_problem(kind, n, deg)
    nb = V.num_dofs
    xs = V.dof_coords[:, 0]
    markers = (xs < np.median(xs)).astype(np.int8)
    rng = np.random.default_rng(20261018)
    g = rng.standard_normal(nb)
    w = np.sin(3.0 * V.dof_coords[:, 0]) + V.dof_coords[:, 1] ** 2
    exterior = np.nonzero(np.diff(mesh.f2c_offsets) == 1)[0].astype(np.int32)
    return mesh, Vphi, phi, V, markers, g, w, exterior


def ext_oracle(mesh, Vphi, phi, V, markers, g, w, exterior):
    import oracle as O

    vals = phi.x.array
    dom = O.classify(Vphi.dofmap, vals)
    inside, cut = O.locate(dom, "phi<0"), O.locate(dom, "phi=0")
    rv = O.runtime_quadrature(mesh, Vphi.dofmap, vals, dom, "<", 4)
    rows4 = O.facet_rows(mesh, O.ghost_penalty_facets(mesh, cut, inside))
    rp, cols = O.sparsity(V, np.concatenate([inside, rv.parent_map]), rows4)

    def assemble(out):
        O.assemble_cells(V, "laplace", out, inside, rv, (1.3,), rp, cols)
        O.assemble_interior_facets(V, "ghost_grad_jump", out, rows4, (0.1,), rp, cols)
        return out

    with O.dirichlet("matrix", markers, markers):
        A_bc = assemble(np.zeros(cols.size))
    O.set_diagonal(rp, cols, A_bc, np.nonzero(markers)[0], 1.0)
    b_lift = np.zeros(V.num_dofs)
    with O.dirichlet("lifting", None, markers, g * markers, None, 1.0, b_lift):
        assemble(np.zeros(cols.size))
    sq = np.zeros(1)
    with O.coefficient(w):
        O.assemble_cells(V, "square_fn", sq, inside, rv, (1.5,))
    code, _, _ = O.classify_facets(mesh, Vphi.dofmap, vals, exterior)
    fr = O.facet_runtime_quadrature(mesh, Vphi.dofmap, vals, exterior, "<", 2)
    return dict(row_ptr=rp, cols=cols, A_bc=A_bc, b_lift=b_lift, square=sq[0], facet_codes=code,
                fr_points=fr.points, fr_weights=fr.weights, fr_offsets=fr.offsets, fr_parent_map=fr.parent_map)


def pack_ext(kind, n, deg):
    inp = ext_inputs(kind, n, deg)
    mesh, Vphi, phi, V, markers, g, w, exterior = inp
    d = dict(kind=kind, n=n, degree=deg, phi_dofmap=Vphi.dofmap, phi=phi.x.array, dofmap=V.dofmap, markers=markers,
             g=g, w=w, exterior=exterior)
    d.update(ext_oracle(*inp))
    return d


if __name__ == "__main__":
    for kind, n, deg in EXT_CASES:
        path = os.path.join(HERE, f"ext_{kind}{n}_P{deg}.npz")
        np.savez_compressed(path, **pack_ext(kind, n, deg))
        print(path, os.path.getsize(path) // 1024, "KiB")
    for kind, n, deg in CASES:
        path = os.path.join(HERE, f"{kind}{n}_P{deg}.npz")
        np.savez_compressed(path, **pack(kind, n, deg))
        print(path, os.path.getsize(path) // 1024, "KiB")
