"""CPU oracle for the CutFEMx cut-cell hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package (see cutfem_oracle.cpp for the parity status: "pointwise parity unpinned").
"""
from .oracle import *  # noqa: F401,F403
