"""Mirror of `cutfemx.fem` for the hot path (python/cutfemx/fem.py of the reference):
`form`, `create_matrix`, `assemble_matrix`, `assemble_vector`, `assemble_scalar`.

The reference compiles arbitrary UFL with runintgen/FFCx; this build ships hand-written kernel
families instead (include/cutfemx_b200.h, CFX_K_*), so a form is described by the same tuples
the reference feeds to `create_form_*` (fem.py:346-351): per integral
`(kernel, entities, constants, custom_data)` with `custom_data` = the run-time rules.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DEVICE, HOST, KERNEL, KERNEL_RANK, CfxError, as_arg, check, is_device_array, lib
from .cut import RuntimeQuadratureRules, _List, _bind_topology, _mesh_context, facet_integration_rows_device
from .mesh import FunctionSpace


class CutForm:
    """Integral table of one form (dolfinx_custom_data::fem::Form, Form.h:119-677)."""

    def __init__(self, space: FunctionSpace, rank: int):
        self.function_space = space
        self.rank = rank
        self.ctx = _mesh_context(space.mesh)
        self._h = C.c_void_p()
        self._owners = []  # fem.py:57,327-328: keeps custom_data alive
        sidx = self.ctx.space_index(space)
        check(self.ctx.handle, lib().cfx_form_create(self.ctx.handle, sidx, rank, C.byref(self._h)))

    def add_cell_integral(self, kernel: str, cells=None, rules: RuntimeQuadratureRules | None = None,
                          constants=(1.0,)):
        """Cell integral over the mixed measure subdomain_data=[cells, rules]
        (demo_poisson.py:165-167): `cells` use the compile-time rule, `rules` the run-time one."""
        kid = KERNEL[kernel]
        if KERNEL_RANK[kid] != self.rank:
            raise ValueError(f"kernel '{kernel}' has rank {KERNEL_RANK[kid]}, form has rank {self.rank}")
        if isinstance(cells, _List):  # library list: borrowed as it is, its length may still be on the device
            cst = np.ascontiguousarray(list(constants), dtype=np.float64)
            h = self.ctx.handle
            check(h, lib().cfx_form_add_cell_integral_list(h, self._h, kid, cells._h,
                                                           rules._h if rules is not None else None,
                                                           C.c_void_p(cst.ctypes.data), int(cst.size)))
            self._owners += [cells, rules]
            return self
        if cells is None:
            p, ms, n, keep = None, HOST, 0, None
        else:
            p, ms, keep = as_arg(cells, np.int32)
            n = int(keep.numel() if is_device_array(keep) else keep.size)
        cst = np.ascontiguousarray(list(constants), dtype=np.float64)
        h = self.ctx.handle
        check(h, lib().cfx_form_add_cell_integral(h, self._h, kid, p, C.c_int64(n), ms,
                                                  rules._h if rules is not None else None,
                                                  C.c_void_p(cst.ctypes.data), int(cst.size)))
        self._owners += [keep, rules]
        return self

    def add_exterior_facet_integral(self, kernel: str, rules: RuntimeQuadratureRules, constants=(1.0,)):
        """`integrand * ds(subdomain_data=rules)` on facet-hosted rules.  Functionals of the measure ("one", rank 0,
        test_cut_api.py:504-527) are summed on the facet rules directly; every other kernel family runs on the rules
        re-expressed in the reference coordinates of the facets' cells (_facet_payload_with_rows,
        _runintgen_adapter.py:605-680): one cell integral per local facet index, so that no cell appears twice in a
        rule set."""
        cst = np.ascontiguousarray(list(constants), dtype=np.float64)
        h = self.ctx.handle
        if self.rank == 0 and kernel == "one":
            check(h, lib().cfx_form_add_exterior_facet_integral(h, self._h, KERNEL[kernel], rules._h,
                                                                C.c_void_p(cst.ctypes.data), int(cst.size)))
            self._owners.append(rules)
            return self
        from .cut import RuntimeQuadratureRules as _Rules
        from .cut import _bind_topology

        msh = self.function_space.mesh
        _bind_topology(msh, self.ctx)
        for lf in range(msh.tdim + 1):
            cr = _Rules(self.ctx, rules.selector, rules.ls, rules.order)
            cr.ctx_gdim = msh.gdim
            check(h, lib().cfx_rules_facets_to_cells(h, rules._h, lf, C.byref(cr._h)))
            if cr.num_rules:
                self.add_cell_integral(kernel, None, cr, constants)
        self._owners.append(rules)
        return self

    def set_coefficient(self, values):
        """The form's ordinary Function coefficient (its x.array over owned+ghost dofs).  The reference packs it per
        entity before every assembly (pack_coefficients, pack_form.h:68-158, called from assembler.h:213-219); here
        the kernels gather the cell's values through the dofmap while they run."""
        p, ms, keep = as_arg(values, np.float64)
        n = int(keep.numel() if is_device_array(keep) else keep.size)
        check(self.ctx.handle, lib().cfx_form_set_coefficient(self.ctx.handle, self._h, p, C.c_int64(n), ms))
        self._owners.append(keep)
        return self

    def add_interior_facet_integral(self, kernel: str, facets=None, rows=None, constants=(1.0,)):
        """Interior-facet integral over raw facet ids (converted with facet_integration_rows,
        _runintgen_adapter.py:416-435) or over ready (cell0, lf0, cell1, lf1) rows."""
        kid = KERNEL[kernel]
        msh = self.function_space.mesh
        _bind_topology(msh, self.ctx)
        if rows is None:
            rows = facet_integration_rows_device(msh, facets)
        if isinstance(rows, _List):
            cst = np.ascontiguousarray(list(constants), dtype=np.float64)
            h = self.ctx.handle
            check(h, lib().cfx_form_add_interior_facet_integral_list(h, self._h, kid, rows._h,
                                                                     C.c_void_p(cst.ctypes.data), int(cst.size)))
            self._owners += [rows]
            return self
        else:
            p, ms, keep = as_arg(rows, np.int32)
            n = int(keep.numel() if is_device_array(keep) else keep.size) // 4
        cst = np.ascontiguousarray(list(constants), dtype=np.float64)
        h = self.ctx.handle
        check(h, lib().cfx_form_add_interior_facet_integral(h, self._h, kid, p, C.c_int64(n), ms,
                                                            C.c_void_p(cst.ctypes.data), int(cst.size)))
        self._owners += [keep]
        return self

    def free(self):
        if self._h:
            lib().cfx_form_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


def form(space: FunctionSpace, rank: int, integrals=()) -> CutForm:
    """Build a form from integral tuples `(itype, kernel, entities, constants, custom_data)`
    with itype in {"cell", "interior_facet"}."""
    f = CutForm(space, rank)
    for itype, kernel, entities, constants, custom_data in integrals:
        if itype == "cell":
            f.add_cell_integral(kernel, entities, custom_data, constants)
        elif itype == "interior_facet":
            f.add_interior_facet_integral(kernel, facets=entities, constants=constants)
        else:
            raise ValueError(f"unsupported integral type '{itype}'")
    return f


class MatrixCSR:
    """dolfinx.la.MatrixCSR stand-in: device-resident pattern + values, exported on demand."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()
        self._cache = {}

    def _sizes(self):
        nr, nnz = C.c_int64(), C.c_int64()
        check(self.ctx.handle, lib().cfx_pattern_sizes(self._h, C.byref(nr), C.byref(nnz)))
        return nr.value, nnz.value

    @property
    def shape(self):
        n = self._sizes()[0]
        return (n, n)

    @property
    def nnz(self) -> int:
        return self._sizes()[1]

    def _fetch_pattern(self):
        if "indptr" not in self._cache:
            nr, nnz = self._sizes()
            rp = np.empty(nr + 1, dtype=np.int64)
            cols = np.empty(nnz, dtype=np.int32)
            h = self.ctx.handle
            check(h, lib().cfx_pattern_fetch(h, self._h, C.c_void_p(rp.ctypes.data), C.c_void_p(cols.ctypes.data), HOST))
            self._cache.update(indptr=rp, indices=cols)

    @property
    def indptr(self):
        self._fetch_pattern()
        return self._cache["indptr"]

    @property
    def indices(self):
        self._fetch_pattern()
        return self._cache["indices"]

    @property
    def block_size(self) -> int:
        return int(lib().cfx_pattern_block_size(self._h))

    @property
    def data(self):
        """Copy of the values on the host: nnz entries, or nnz row-major bs x bs blocks for vector spaces."""
        out = np.empty(self.nnz * self.block_size ** 2)
        if out.size:
            h = self.ctx.handle
            check(h, lib().cfx_pattern_values_fetch(h, self._h, C.c_void_p(out.ctypes.data), HOST))
        return out

    # zero-copy device views (valid until the next create_matrix on this object)
    def indptr_device(self):
        from ._lib import device_view

        nr, _ = self._sizes()
        return device_view(lib().cfx_pattern_row_ptr_device_ptr(self._h), nr + 1, np.int64, self.ctx.device, self)

    def indices_device(self):
        from ._lib import device_view

        return device_view(lib().cfx_pattern_cols_device_ptr(self._h), self.nnz, np.int32, self.ctx.device, self)

    def values_device(self):
        from ._lib import device_view

        return device_view(lib().cfx_pattern_values_device_ptr(self._h), self.nnz * self.block_size ** 2, np.float64,
                           self.ctx.device, self)

    def copy_to_host_async(self, h_indptr, h_indices, h_values, stream=None):
        """Enqueue the device->host copy of the CSR arrays into (pinned) torch host tensors on `stream`
        (a torch.cuda.Stream; default: the current one).  Returns (n_rows + 1, nnz).  The caller owns the
        ordering: record an event after assembly and make `stream` wait for it."""
        import torch

        nr, nnz = self._sizes()
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
            h_indptr[: nr + 1].copy_(self.indptr_device(), non_blocking=True)
            h_indices[:nnz].copy_(self.indices_device(), non_blocking=True)
            bb = self.block_size ** 2
            h_values[: nnz * bb].copy_(self.values_device(), non_blocking=True)
        return nr + 1, nnz

    def to_scipy(self):
        import scipy.sparse as sp

        bs = self.block_size
        if bs > 1:
            n = self.shape[0] * bs
            return sp.bsr_matrix((self.data.reshape(-1, bs, bs), self.indices, self.indptr), shape=(n, n)).tocsr()
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)

    def scatter_reverse(self):  # single rank: no-op (multi-rank: parallel.py)
        return None

    def free(self):
        if self._h:
            lib().cfx_pattern_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


def create_matrix(a: CutForm, A: MatrixCSR | None = None) -> MatrixCSR:
    """create_sparsity_pattern + finalize + MatrixCSR (assembler.h:567-592, wrappers/fem.cpp:266-276)."""
    if a.rank != 2:
        raise RuntimeError("Cannot create sparsity pattern. Form is not a bilinear.")
    if A is None:
        A = MatrixCSR(a.ctx)
    A._cache.clear()
    check(a.ctx.handle, lib().cfx_create_sparsity(a.ctx.handle, a._h, C.byref(A._h)))
    return A


def create_ghost_row_pattern(a: CutForm, row_begin: int, A: MatrixCSR | None = None) -> MatrixCSR:
    """The entries SparsityPattern::finalize() ships to other ranks: pattern of the rows >= row_begin
    (the ghost rows) only, without the deactivation diagonal (cfx_create_sparsity_rows)."""
    if A is None:
        A = MatrixCSR(a.ctx)
    A._cache.clear()
    check(a.ctx.handle, lib().cfx_create_sparsity_rows(a.ctx.handle, a._h, C.c_int64(row_begin), C.byref(A._h)))
    return A


def insert_pattern_entries(a: CutForm, rows, cols) -> None:
    """la::SparsityPattern::insert for entries received from other ranks (sorted by row)."""
    if rows is None or int(rows.numel() if is_device_array(rows) else np.size(rows)) == 0:
        check(a.ctx.handle, lib().cfx_form_insert_pattern_entries(a.ctx.handle, a._h, None, None, C.c_int64(0), HOST))
        return
    pr, msr, kr = as_arg(rows, np.int32)
    pc, msc, kc = as_arg(cols, np.int32)
    if msr != msc:
        raise CfxError("inserted pattern rows and columns must live in the same memory space")
    n = int(kr.numel() if is_device_array(kr) else kr.size)
    check(a.ctx.handle, lib().cfx_form_insert_pattern_entries(a.ctx.handle, a._h, pr, pc, C.c_int64(n), msr))


class DirichletBC:
    """Stand-in for dolfinx.fem.DirichletBC on the flat-array spaces of this mirror: `dofs` are blocked dof
    indices bs*dof + k (what DirichletBC::dof_indices returns unrolled), `value` a scalar or an array over all
    blocked dofs (the `g` function's x.array)."""

    def __init__(self, value, dofs, V: FunctionSpace):
        self.function_space = V
        self.dofs = np.unique(np.asarray(dofs, dtype=np.int32))
        n = V.num_dofs * V.bs
        if self.dofs.size and (self.dofs[0] < 0 or self.dofs[-1] >= n):
            raise ValueError("Dirichlet dof index out of range")
        self.g = np.full(n, float(value)) if np.isscalar(value) else np.ascontiguousarray(value, dtype=np.float64)
        if self.g.size != n:
            raise ValueError("Dirichlet values must cover all (blocked) dofs of the space")

    def owned_dofs(self) -> np.ndarray:
        V = self.function_space
        return self.dofs[self.dofs < V.num_dofs_owned * V.bs]


def dirichletbc(value, dofs, V: FunctionSpace) -> DirichletBC:
    return DirichletBC(value, dofs, V)


def _bc_arrays(V: FunctionSpace, bcs):
    """marker / value arrays over the blocked dofs (assembler.h:657-672: DirichletBC::mark_dofs / set)."""
    n = V.num_dofs * V.bs
    markers, values = np.zeros(n, dtype=np.int8), np.zeros(n)
    for bc in bcs:
        if bc.function_space is not V:
            raise ValueError("Dirichlet condition defined on a different function space")
        markers[bc.dofs] = 1
        values[bc.dofs] = bc.g[bc.dofs]
    return markers, values


def assemble_matrix(a: CutForm, A: MatrixCSR | None = None, *, bcs=None, diag: float = 1.0,
                    diag_inactive: float = 0.0) -> MatrixCSR:
    """fem.py:886-942: create the matrix if none is given, then add the form into it; with `bcs` the
    Dirichlet rows and columns of every element tensor are dropped (assembler.h:643-683) and `diag` is set on
    the owned Dirichlet rows afterwards (insert_diagonal, fem.py:935-941 -> assembler.h:745-787).
    `diag_inactive` writes that value on the diagonal of rows outside the active domain
    (deactivate_outside, fem/deactivate.h:402-418)."""
    zero = 0
    if A is None:
        A = create_matrix(a)
        zero = 1
    h = a.ctx.handle
    if not bcs:
        check(h, lib().cfx_assemble_matrix(h, a._h, A._h, zero, C.c_double(diag_inactive), None, HOST))
        return A
    markers, _ = _bc_arrays(a.function_space, bcs)
    pm = C.c_void_p(markers.ctypes.data)
    check(h, lib().cfx_assemble_matrix_bc(h, a._h, A._h, zero, C.c_double(diag_inactive), pm, pm, HOST, None, HOST))
    for bc in bcs:
        rows = np.ascontiguousarray(bc.owned_dofs(), dtype=np.int32)
        check(h, lib().cfx_set_diagonal(h, A._h, C.c_void_p(rows.ctypes.data), C.c_int64(rows.size), C.c_double(diag),
                                        HOST))
    A._cache.pop("data", None)
    return A


def assemble_system_bc(a: CutForm, A: MatrixCSR, L: CutForm, b, bcs, x0=None, alpha: float = 1.0, diag: float = 1.0):
    """assemble_matrix(a, bcs) + assemble_vector(L) + apply_lifting + set_bc (demo_elasticity.py:78-84) with ONE
    assembly of the unconstrained system (cfx_assemble_system_bc).  `b`: torch CUDA vector (owned+ghost, blocked),
    overwritten like A."""
    import torch

    if not is_device_array(b):
        raise TypeError("assemble_system_bc needs a device vector (torch CUDA tensor)")
    V = a.function_space
    markers, values = _bc_arrays(V, bcs)
    dev = b.device
    tm = torch.from_numpy(markers).to(dev)
    tv = torch.from_numpy(values).to(dev)
    tx = None if x0 is None else torch.as_tensor(np.ascontiguousarray(x0, dtype=np.float64), device=dev)
    rows = np.unique(np.concatenate([bc.owned_dofs() for bc in bcs])).astype(np.int32) if bcs else np.zeros(0, np.int32)
    tr = torch.from_numpy(rows).to(dev)
    h = a.ctx.handle
    check(h, lib().cfx_assemble_system_bc(h, a._h, A._h, L._h, C.c_void_p(b.data_ptr()), C.c_void_p(tm.data_ptr()),
                                          C.c_void_p(tv.data_ptr()), None if tx is None else C.c_void_p(tx.data_ptr()),
                                          C.c_double(alpha), C.c_void_p(tr.data_ptr()) if rows.size else None,
                                          C.c_int64(rows.size), C.c_double(diag)))
    A._cache.pop("data", None)
    return A, b


def apply_lifting(b: np.ndarray, a, bcs, x0=None, alpha: float = 1.0, *, A=None) -> None:
    """cutfemx.fem.apply_lifting (fem.py:604-635 -> assemble_vector_impl.h:383-564):
    b -= alpha * A_j (g_j - x0_j) for every bilinear form a[j] with conditions bcs[j].  `A`: matrices (or one
    matrix) carrying the sparsity pattern of each form; created from the form when omitted."""
    if isinstance(a, CutForm):
        a, bcs = [a], [bcs]
        x0 = None if x0 is None else [x0]
        A = None if A is None else [A]
    for j, form in enumerate(a):
        if form is None or not bcs[j]:
            continue
        Aj = A[j] if A is not None and A[j] is not None else create_matrix(form)
        markers, values = _bc_arrays(form.function_space, bcs[j])
        x0j = None if x0 is None or x0[j] is None else np.ascontiguousarray(x0[j], dtype=np.float64)
        bb = np.ascontiguousarray(b, dtype=np.float64)
        h = form.ctx.handle
        check(h, lib().cfx_apply_lifting(h, form._h, Aj._h, C.c_void_p(bb.ctypes.data), C.c_void_p(values.ctypes.data),
                                         C.c_void_p(markers.ctypes.data),
                                         None if x0j is None else C.c_void_p(x0j.ctypes.data), C.c_double(alpha), HOST))
        if bb is not b:
            b[:] = bb


def set_bc(b: np.ndarray, bcs, x0=None, alpha: float = 1.0) -> None:
    """dolfinx.fem.set_bc as demo_elasticity.py:84 calls it: b[dofs] = alpha * (g - x0)."""
    if not bcs:
        return
    ctx = _mesh_context(bcs[0].function_space.mesh)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    x0a = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    for bc in bcs:
        d = np.ascontiguousarray(bc.dofs, dtype=np.int32)
        check(ctx.handle, lib().cfx_set_bc(ctx.handle, C.c_void_p(bb.ctypes.data), C.c_int64(bb.size),
                                           C.c_void_p(d.ctypes.data), C.c_int64(d.size), C.c_void_p(bc.g.ctypes.data),
                                           None if x0a is None else C.c_void_p(x0a.ctypes.data), C.c_double(alpha),
                                           HOST))
    if bb is not b:
        b[:] = bb


def assemble_system(a: CutForm, A: MatrixCSR, L: CutForm, b, *, zero_b: bool = True, diag_inactive: float = 0.0):
    """assemble_matrix(a, A) and assemble_vector(L, b) in one pass over the mesh where the two forms share
    their integration domains (cfx_assemble_system).  `b`: torch CUDA tensor of owned+ghost entries.
    Adds into A (like assemble_matrix); bit-identical to the two separate calls."""
    if a.rank != 2 or L.rank != 1:
        raise RuntimeError("assemble_system expects a bilinear and a linear form")
    if not is_device_array(b):
        raise TypeError("assemble_system needs a device vector (torch CUDA tensor)")
    check(a.ctx.handle, lib().cfx_assemble_system(a.ctx.handle, a._h, A._h, 0, C.c_double(diag_inactive), L._h,
                                                  C.c_void_p(b.data_ptr()), int(zero_b)))
    return A, b


def assemble_vector(L: CutForm, b: np.ndarray | None = None) -> np.ndarray:
    """fem.py:851-883: adds into `b` (owned+ghost entries) or creates a zero vector first."""
    if L.rank != 1:
        raise RuntimeError("assemble_vector expects a linear form")
    n = L.function_space.num_dofs * L.function_space.bs
    zero = 0
    if b is None:
        b = np.zeros(n)
        zero = 1
    p, ms, keep = as_arg(b, np.float64)
    check(L.ctx.handle, lib().cfx_assemble_vector(L.ctx.handle, L._h, p, zero, ms))
    if ms == HOST and keep is not b:
        b[:] = keep
    return b


class ActiveDomain:
    """Mirror of cutfemx.fem.ActiveDomain (python/cutfemx/fem.py:88-123): `active_cells`, `inactive_dofs`."""

    def __init__(self, form: CutForm, cells: _List, dofs: _List):
        self.form, self._cells, self._dofs = form, cells, dofs
        self.function_space = form.function_space

    @property
    def active_cells(self) -> np.ndarray:
        return self._cells.numpy()

    @property
    def inactive_dofs(self) -> np.ndarray:
        return self._dofs.numpy()


def active_domain(a: CutForm) -> ActiveDomain:
    """cutfemx.fem.active_domain (fem.py:683-695 -> fem/deactivate.h:387-400): the cells the form
    integrates over and the owned dofs none of them touches."""
    if a.rank != 2:
        raise ValueError("cutfemx.fem.active_domain requires a bilinear form")
    cells, dofs = _List(a.ctx), _List(a.ctx)
    check(a.ctx.handle, lib().cfx_active_domain(a.ctx.handle, a._h, C.byref(cells._h), C.byref(dofs._h)))
    return ActiveDomain(a, cells, dofs)


def deactivate_outside(A: MatrixCSR, b_or_active_domain, active_domain_or_none=None, *, diagonal: float = 1.0,
                       rhs_value: float = 0.0) -> ActiveDomain:
    """cutfemx.fem.deactivate_outside (fem.py:698-736 -> deactivate.h:402-418): set the diagonal of the
    inactive rows (and, with a right-hand side, its inactive entries).  `b`: torch CUDA tensor (in place) or
    numpy array (updated in place through a device copy)."""
    if isinstance(b_or_active_domain, ActiveDomain):
        if active_domain_or_none is not None:
            raise TypeError("deactivate_outside(A, active_domain) takes no RHS vector")
        domain, b = b_or_active_domain, None
    else:
        if active_domain_or_none is None:
            raise TypeError("deactivate_outside(A, b, active_domain) requires active_domain")
        domain, b = active_domain_or_none, b_or_active_domain
    dofs = domain._dofs
    h = A.ctx.handle
    if b is None or is_device_array(b):
        pb = C.c_void_p(b.data_ptr()) if b is not None else None
        check(h, lib().cfx_deactivate_outside(h, A._h, C.c_void_p(dofs.device_ptr), C.c_int64(dofs.size), DEVICE,
                                              C.c_double(diagonal), pb, C.c_double(rhs_value)))
    else:
        import torch

        tb = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(f"cuda:{A.ctx.device}")
        check(h, lib().cfx_deactivate_outside(h, A._h, C.c_void_p(dofs.device_ptr), C.c_int64(dofs.size), DEVICE,
                                              C.c_double(diagonal), C.c_void_p(tb.data_ptr()), C.c_double(rhs_value)))
        b[:] = tb.cpu().numpy()
    A._cache.pop("data", None)
    return domain


def assemble_scalar(M: CutForm) -> float:
    """fem.py assemble_scalar -> assemble_scalar_impl.h:26-275 (local value; allreduce is the caller's)."""
    if M.rank != 0:
        raise RuntimeError("assemble_scalar expects a functional")
    out = C.c_double()
    check(M.ctx.handle, lib().cfx_assemble_scalar(M.ctx.handle, M._h, C.byref(out)))
    return float(out.value)
