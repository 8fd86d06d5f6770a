"""ctypes binding of libcutfemx_b200.so (the C ABI in include/cutfemx_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module
raises, and every entry point raises `CfxError` when the C side reports a failure (e.g. no
CUDA device).  The library is built in-tree by `cutfemx_b200._build.build_library()`
(`__graft_entry__.build()` calls it).
"""
from __future__ import annotations

import ctypes as C
import functools
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcutfemx_b200.so")

HOST, DEVICE = 0, 1
TRIANGLE, TETRAHEDRON = 3, 4
DOMAIN_INSIDE, DOMAIN_INTERSECTED, DOMAIN_OUTSIDE = 1, 2, 3
REL = {"<": 0, "<=": 1, ">": 2, ">=": 3, "=": 4}
KERNEL = {"laplace": 1, "mass": 2, "nitsche": 3, "ghost_grad_jump": 4, "source": 5, "nitsche_rhs": 6, "one": 7,
          "elasticity": 8, "source_vec": 9, "square_fn": 10, "nitsche_vec": 11}
KERNEL_RANK = {1: 2, 2: 2, 3: 2, 4: 2, 5: 1, 6: 1, 7: 0, 8: 2, 9: 1, 10: 0, 11: 2}

# every symbol include/cutfemx_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "cfx_ctx_create", "cfx_ctx_destroy", "cfx_last_error", "cfx_sync", "cfx_version", "cfx_launch_count",
    "cfx_mesh_bind", "cfx_topology_bind", "cfx_levelset_bind", "cfx_levelset_unbind", "cfx_set_host_cells", "cfx_cut_facets", "cfx_ecut_locate", "cfx_ecut_runtime_quadrature", "cfx_rules_facets_to_cells", "cfx_ecut_free",
    "cfx_update", "cfx_counts", "cfx_domain_fetch",
    "cfx_locate_entities", "cfx_list_size", "cfx_list_device_ptr", "cfx_list_fetch", "cfx_list_free",
    "cfx_runtime_quadrature", "cfx_rules_sizes", "cfx_rules_fetch", "cfx_rules_physical_points", "cfx_rules_surface_provenance", "cfx_rules_free",
    "cfx_simplex_rule", "cfx_set_simplex_rule", "cfx_evaluate_normals", "cfx_evaluate_values",
    "cfx_ghost_penalty_facets", "cfx_interior_facets_for_cells", "cfx_facet_integration_rows", "cfx_space_bind",
    "cfx_form_create", "cfx_form_set_coefficient", "cfx_form_add_exterior_facet_integral", "cfx_form_add_cell_integral", "cfx_form_add_interior_facet_integral", "cfx_form_free",
    "cfx_create_sparsity", "cfx_pattern_import", "cfx_pattern_sizes", "cfx_pattern_block_size", "cfx_pattern_fetch",
    "cfx_pattern_values_device_ptr", "cfx_pattern_row_ptr_device_ptr", "cfx_pattern_cols_device_ptr",
    "cfx_pattern_values_fetch", "cfx_pattern_free", "cfx_assemble_matrix", "cfx_assemble_system", "cfx_assemble_vector",
    "cfx_assemble_scalar", "cfx_stage_count", "cfx_stage_name", "cfx_stage_timing_enable", "cfx_stage_ms",
    "cfx_stage_reset", "cfx_form_insert_pattern_entries", "cfx_create_sparsity_rows", "cfx_pattern_positions",
    "cfx_gather_f64", "cfx_scatter_add_f64", "cfx_active_domain", "cfx_active_indicator_device_ptr",
    "cfx_inactive_dofs", "cfx_deactivate_outside", "cfx_assemble_matrix_bc", "cfx_assemble_system_bc", "cfx_set_diagonal", "cfx_apply_lifting",
    "cfx_set_bc", "cfx_meshgen_box", "cfx_meshgen_rectangle", "cfx_meshgen_level_set", "cfx_meshgen_p2_tet_dofmap",
    "cfx_device_bytes", "cfx_set_deferred", "cfx_check", "cfx_graph_begin", "cfx_graph_end", "cfx_graph_launch",
    "cfx_lane_begin", "cfx_lane_end", "cfx_lane_join", "cfx_set_lanes",
    "cfx_graph_kernel_nodes", "cfx_graph_free", "cfx_facet_integration_rows_list", "cfx_form_add_cell_integral_list",
    "cfx_form_add_interior_facet_integral_list", "cfx_comm_unique_id", "cfx_comm_init", "cfx_comm_destroy",
    "cfx_xplan_create", "cfx_xplan_free", "cfx_xplan_pack_pattern", "cfx_xplan_exchange", "cfx_xplan_insert_pattern",
    "cfx_xplan_pack_values", "cfx_xplan_unpack_add", "cfx_xplan_buffer", "cfx_space_counters", "cfx_space_forget",
]


class CfxError(RuntimeError):
    """Raised where the reference raises std::runtime_error / invalid_argument / out_of_range."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libcutfemx_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.cfx_last_error.restype = C.c_char_p
        L.cfx_last_error.argtypes = [C.c_void_p]
        L.cfx_stage_name.restype = C.c_char_p
        L.cfx_stage_name.argtypes = [C.c_void_p, C.c_int]
        L.cfx_list_size.restype = C.c_int64
        L.cfx_list_size.argtypes = [C.c_void_p]
        L.cfx_launch_count.restype = C.c_int64
        L.cfx_launch_count.argtypes = [C.c_void_p]
        L.cfx_device_bytes.restype = C.c_int64
        L.cfx_device_bytes.argtypes = [C.c_void_p]
        L.cfx_graph_kernel_nodes.restype = C.c_int64
        L.cfx_graph_kernel_nodes.argtypes = [C.c_void_p]
        L.cfx_set_deferred.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cfx_list_device_ptr.restype = C.c_void_p
        L.cfx_list_device_ptr.argtypes = [C.c_void_p]
        for name in ("cfx_pattern_values_device_ptr", "cfx_pattern_row_ptr_device_ptr", "cfx_pattern_cols_device_ptr"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        L.cfx_active_indicator_device_ptr.restype = C.c_void_p
        L.cfx_active_indicator_device_ptr.argtypes = [C.c_void_p, C.c_void_p]
        L.cfx_ctx_destroy.restype = None
        L.cfx_comm_destroy.restype = None
        L.cfx_comm_destroy.argtypes = [C.c_void_p]
        L.cfx_comm_unique_id.argtypes = [C.c_void_p, C.c_char_p]
        L.cfx_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        for name in ("cfx_list_free", "cfx_rules_free", "cfx_pattern_free", "cfx_form_free", "cfx_ecut_free",
                     "cfx_graph_free", "cfx_xplan_free"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(ctx, status: int):
    if status != 0:
        msg = lib().cfx_last_error(ctx)
        raise CfxError((msg.decode() if msg else "unknown error") + f" [status {status}]")


class _CudaView:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


def device_view(ptr: int, n: int, dtype, device: int, owner=None):
    """torch tensor aliasing `n` elements of `dtype` at device pointer `ptr` (valid while `owner` lives
    and the library does not reallocate the buffer)."""
    import torch

    typestr = {np.float64: "<f8", np.int64: "<i8", np.int32: "<i4"}[dtype]
    if n == 0 or not ptr:
        tdt = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
        return torch.empty(0, dtype=tdt, device=f"cuda:{device}")
    return torch.as_tensor(_CudaView(ptr, n, typestr, owner), device=f"cuda:{device}")


def is_device_array(a) -> bool:
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def as_arg(a, dtype):
    """(pointer, memspace, keepalive) for a numpy array (host) or a torch CUDA tensor (device)."""
    if a is None:
        return None, HOST, None
    if is_device_array(a):
        import torch

        want = {np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64, np.int8: torch.int8,
                np.uint8: torch.uint8}[dtype]
        if a.dtype != want or not a.is_contiguous():
            a = a.to(want).contiguous()
        return C.c_void_p(a.data_ptr()), DEVICE, a
    arr = np.ascontiguousarray(a, dtype=dtype)
    return C.c_void_p(arr.ctypes.data), HOST, arr


def parse_selector(expr: str, names):
    """Compiled selector (cached: the same few expressions are parsed every step of a moving-domain loop)."""
    return _parse_selector_cached(str(expr), tuple(names))


@functools.lru_cache(maxsize=256)
def _parse_selector_cached(expr: str, names: tuple):
    to, cl, cr = _parse_selector(expr, names)
    for a in (to, cl, cr):
        a.setflags(write=False)
    return to, cl, cr


def _parse_selector(expr: str, names):
    """The selector grammar of cutcells::parse_selection_expr as the reference uses it
    (cut.cpp:881-882): whitespace ignored (cut.cpp:47-57), clauses `name rel 0` joined by
    `and` inside a term, terms joined by `or`."""
    import re

    s = "".join(str(expr).split())
    if not s:
        raise ValueError("empty level-set selector")
    clause_re = re.compile(r"([A-Za-z_]\w*?)(<=|>=|<|>|=)(0(?:\.0*)?)(?=and|or|$)")
    term_offsets, cls, crel = [0], [], []
    pos = 0
    while True:
        m = clause_re.match(s, pos)
        if not m:
            raise ValueError(f"cannot parse selector '{expr}' at '{s[pos:]}' (expected `name rel 0`)")
        name, op = m.group(1), m.group(2)
        if name not in names:
            raise ValueError(f"selector references unknown level set '{name}' (known: {list(names)})")
        cls.append(list(names).index(name))
        crel.append(REL[op])
        pos = m.end()
        if pos == len(s):
            break
        if s.startswith("and", pos):
            pos += 3
        elif s.startswith("or", pos):
            pos += 2
            term_offsets.append(len(cls))
        else:
            raise ValueError(f"cannot parse selector '{expr}' at '{s[pos:]}' (expected `and` / `or`)")
    term_offsets.append(len(cls))
    return (np.asarray(term_offsets, dtype=np.int32), np.asarray(cls, dtype=np.int32),
            np.asarray(crel, dtype=np.int32))
