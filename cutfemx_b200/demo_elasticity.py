"""The cut linear-elasticity pipeline of the reference's demo (python/demo/demo_elasticity.py:213-238) on this
package's public API: vector Lagrange space (block size = gdim), sigma(u) = 2 mu eps(u) + lambda tr(eps(u)) I,

    a = inner(sigma(u), eps(v)) dx(inside + cut cells, run-time rules)
        - (sigma(u) n).v - (sigma(v) n).u + gamma (2 mu + lambda)/h u.v   on the interface rules   (symmetric Nitsche)
        + gamma_g (2 mu + lambda) avg(h) inner(jump(grad u, n), jump(grad v, n)) dS(ghost-penalty facets)
    L = inner(f, v) dx

BASELINE.json configs[3]: torus level set, P2 vector space on tetrahedra.  Same step structure, persistent objects,
deferred sizes and graph capture as demo_poisson.CutPoisson.
"""
from __future__ import annotations

import importlib

from . import fem as _fem
from . import level_set as _ls
from .demo_poisson import CutPoisson

_cut = importlib.import_module(__package__ + ".cut")


class CutElasticity(CutPoisson):
    def __init__(self, mesh, phi, V, order: int = 4, E: float = 1.0e3, nu: float = 0.3, gamma: float = 40.0,
                 gamma_g: float = 0.05, force=(0.0, 0.0, -1.0)):
        super().__init__(mesh, phi, V, order=order, gamma=gamma, gamma_g=gamma_g)
        self.mu = E / (2.0 * (1.0 + nu))
        self.lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
        self.force = tuple(force)[: mesh.gdim]

    def build_forms(self, assemble_rhs: bool = True):
        cd = self.cut_data
        k = self.keep if self.persistent else {}
        mu, lam = self.mu, self.lam
        ln = self.lanes
        _cut.update(cd)
        with self.ctx.lane(ln[0]):
            rv = _cut.runtime_quadrature(cd, "phi<0", self.order, out=k.get("rv"))
        with self.ctx.lane(ln[1]):
            ri = _cut.runtime_quadrature(cd, "phi=0", self.order, out=k.get("ri"))
            _ls.attach_normal(cd, self.phi, ri)
        with self.ctx.lane(ln[2]):
            ghost = _cut.ghost_penalty_facets_device(cd, "phi<0", out=k.get("ghost"))
            rows = _cut.facet_integration_rows_device(self.mesh, ghost, out=k.get("rows"))
        inside = _cut.locate_entities_device(cd, "phi<0", out=k.get("inside"))
        self.ctx.join()
        a = _fem.CutForm(self.V, 2)
        a.add_cell_integral("elasticity", inside, rv, (mu, lam))
        a.add_cell_integral("nitsche_vec", None, ri, (mu, lam, self.gamma))
        if self.persistent or ghost.size > 0:
            a.add_interior_facet_integral("ghost_grad_jump", rows=rows, constants=(self.gamma_g * (2.0 * mu + lam),))
        L = None
        if assemble_rhs:
            L = _fem.CutForm(self.V, 1)
            L.add_cell_integral("source_vec", inside, rv, self.force)
        self.last = dict(inside=inside, rv=rv, ri=ri, ghost=ghost, rows=rows, a=a, L=L)
        if self.persistent:
            self.keep = dict(inside=inside, rv=rv, ri=ri, ghost=ghost, rows=rows)
        return a, L
