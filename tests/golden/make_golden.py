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


if __name__ == "__main__":
    for kind, n, deg in CASES:
        path = os.path.join(HERE, f"{kind}{n}_P{deg}.npz")
        np.savez_compressed(path, **pack(kind, n, deg))
        print(path, os.path.getsize(path) // 1024, "KiB")
