"""cutfemx_b200 -- the cut-cell hot path of CutFEMx on B200 (sm_100a), behind CutFEMx's own API.

Public surface mirrors python/cutfemx/__init__.py:11-68 for the accelerated path only; everything
computes in libcutfemx_b200.so (hand-written CUDA, no CPU fallback).
"""
__version__ = "0.1.0"

from . import fem as fem
from . import level_set as level_set
from . import mesh as mesh
from ._lib import CfxError as CfxError
from .cut import (
    CutData as CutData,
    RuntimeQuadratureRules as RuntimeQuadratureRules,
    cut as cut,
    facet_integration_rows as facet_integration_rows,
    ghost_penalty_facets as ghost_penalty_facets,
    interior_facets_for_cells as interior_facets_for_cells,
    locate_entities as locate_entities,
    runtime_quadrature as runtime_quadrature,
    runtime_quadratures as runtime_quadratures,
    update as update,
)
from .level_set import level_set_value as level_set_value
from .level_set import normal as normal
