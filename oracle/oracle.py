"""ctypes front end of oracle/liboracle.so (the CPU restatement in cutfem_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import rules as _rules

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REL = {"<": 0, "<=": 1, ">": 2, ">=": 3, "=": 4}
INSIDE, INTERSECTED, OUTSIDE = 1, 2, 3

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cutfem_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_last_error.restype = C.c_char_p
        for name in ("orc_locate", "orc_runtime_quadrature", "orc_runtime_quadrature_p2", "orc_ghost_penalty_facets",
                     "orc_interior_facets_for_cells", "orc_sparsity"):
            getattr(_LIB, name).restype = C.c_int64
        _registered.clear()
    return _LIB


_registered: set = set()


def _need_rule(dim: int, order: int):
    if (dim, order) in _registered:
        return
    p, w = _rules.simplex_rule(dim, order)
    p = np.ascontiguousarray(p, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    lib().orc_set_rule(dim, order, int(w.size), _p(p, _f64p), _p(w, _f64p))
    _registered.add((dim, order))


def _p(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def _chk(rc):
    if rc < 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return rc


def _ci32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _cf64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def classify(ls_dofmap, values, ncells=None):
    ls_dofmap = _ci32(ls_dofmap)
    values = _cf64(values)
    n = ls_dofmap.shape[0] if ncells is None else int(ncells)
    dom = np.zeros(n, dtype=np.int8)
    lib().orc_classify(_p(ls_dofmap, _i32p), int(ls_dofmap.shape[1]), _p(values, _f64p), C.c_int64(n), _p(dom, _i8p))
    return dom


def parse_selector(expr: str, names):
    """cutcells::parse_selection_expr + compile_selection_expr as used at cut.cpp:881-882:
    spaces ignored (cut.cpp:47-57), 'and' inside terms, 'or' between terms."""
    s = expr.replace(" ", "")
    term_offsets, cls, crel = [0], [], []
    for term in s.split("or"):
        for clause in term.split("and"):
            for op in ("<=", ">=", "<", ">", "="):
                if op in clause:
                    name, rhs = clause.split(op)
                    if rhs != "0":
                        raise ValueError(f"selector clause '{clause}' must compare with 0")
                    if name not in names:
                        raise ValueError(f"unknown level set '{name}'")
                    cls.append(list(names).index(name))
                    crel.append(REL[op])
                    break
            else:
                raise ValueError(f"cannot parse selector clause '{clause}'")
        term_offsets.append(len(cls))
    return _ci32(term_offsets), _ci32(cls), _ci32(crel)


def locate(domain, expr: str, names=("phi",)):
    domain = np.ascontiguousarray(np.atleast_2d(domain), dtype=np.int8)
    to, cl, cr = parse_selector(expr, names)
    ncells = domain.shape[1]
    out = np.zeros(ncells, dtype=np.int32)
    n = lib().orc_locate(_p(domain, _i8p), C.c_int64(domain.shape[1]), C.c_int64(ncells), int(to.size - 1),
                         _p(to, _i32p), _p(cl, _i32p), _p(cr, _i32p), _p(out, _i32p))
    return out[:n].copy()


class Rules:
    def __init__(self, tdim, points, weights, offsets, parent_map):
        self.tdim, self.points, self.weights, self.offsets, self.parent_map = tdim, points, weights, offsets, parent_map
        self.normals = None


def runtime_quadrature(mesh, ls_dofmap, values, domain, relation: str, order: int) -> Rules:
    """relation in '<', '<=', '>', '>=', '='; rules in ascending parent-cell order.  A level-set dofmap of width
    6 / 10 (P2) takes the higher-order cut (orc_runtime_quadrature_p2)."""
    tdim = mesh.tdim
    _need_rule(tdim - 1 if relation == "=" else tdim, order)
    x, xd, ld, v = _cf64(mesh.x), _ci32(mesh.x_dofmap), _ci32(ls_dofmap), _cf64(values)
    fn = lib().orc_runtime_quadrature if ld.shape[1] == tdim + 1 else lib().orc_runtime_quadrature_p2
    dom = np.ascontiguousarray(domain, dtype=np.int8)
    nc = C.c_int64(mesh.num_cells_local)
    nr = C.c_int64(0)
    args = (mesh.cell_type, _p(x, _f64p), _p(xd, _i32p), _p(ld, _i32p), _p(v, _f64p), _p(dom, _i8p), nc, REL[relation],
            order)
    npts = _chk(fn(*args, None, None, None, None, C.byref(nr)))
    pts = np.zeros((npts, tdim))
    wts = np.zeros(npts)
    off = np.zeros(nr.value + 1, dtype=np.int32)
    pm = np.zeros(nr.value, dtype=np.int32)
    _chk(fn(*args, _p(pts, _f64p), _p(wts, _f64p), _p(off, _i32p), _p(pm, _i32p), C.byref(nr)))
    return Rules(tdim, pts, wts, off, pm)


def surface_provenance(r: Rules, relation: str, level_set_index: int = 0, cut_cells=None):
    """make_surface_provenance, cut.cpp:1273-1308 (+ single_equality_level_set_index, :1258-1271): only a single
    "ls = 0" selector whose part has dimension tdim - 1 carries provenance; one SelectedZeroEntityInfo per rule --
    (cut_cell_id, parent_cell_id, local_zero_entity_id, dimension).  cut_cell_id is the position of the parent cell in
    the ascending cut-cell list (`cut_cells`; default: one rule per cut cell, assumption A1 -- then it is the rule
    index); a P1 level set has one zero entity per cut cell."""
    if relation != "=":
        e = np.zeros(0, dtype=np.int32)
        return dict(level_set_index=-1, cut_cell_ids=e, parent_cell_ids=e, local_zero_entity_ids=e, dimensions=e)
    n = r.parent_map.size
    ids = np.arange(n, dtype=np.int32) if cut_cells is None else np.searchsorted(cut_cells, r.parent_map).astype(np.int32)
    return dict(level_set_index=level_set_index, cut_cell_ids=ids,
                parent_cell_ids=r.parent_map.astype(np.int32), local_zero_entity_ids=np.zeros(n, dtype=np.int32),
                dimensions=np.full(n, r.tdim - 1, dtype=np.int32))


def physical_points(mesh, r: Rules):
    out = np.zeros((mesh.gdim, r.weights.size))
    x, xd = _cf64(mesh.x), _ci32(mesh.x_dofmap)
    lib().orc_physical_points(mesh.cell_type, mesh.gdim, _p(x, _f64p), _p(xd, _i32p), _p(r.points, _f64p),
                              _p(r.offsets, _i32p), _p(r.parent_map, _i32p), C.c_int64(r.parent_map.size),
                              C.c_int64(r.weights.size), _p(out, _f64p))
    return out


def normals(mesh, ls_dofmap, ls_degree, values, r: Rules, sign: float = 1.0):
    out = np.zeros((r.weights.size, mesh.gdim))
    x, xd, ld, v = _cf64(mesh.x), _ci32(mesh.x_dofmap), _ci32(ls_dofmap), _cf64(values)
    lib().orc_normals(mesh.cell_type, mesh.gdim, _p(x, _f64p), _p(xd, _i32p), _p(ld, _i32p), int(ld.shape[1]),
                      ls_degree, _p(v, _f64p), _p(r.points, _f64p), _p(r.offsets, _i32p), _p(r.parent_map, _i32p),
                      C.c_int64(r.parent_map.size), C.c_double(sign), _p(out, _f64p))
    return out


def values(mesh, ls_dofmap, ls_degree, vals, r: Rules):
    out = np.zeros(r.weights.size)
    ld, v = _ci32(ls_dofmap), _cf64(vals)
    lib().orc_values(mesh.cell_type, _p(ld, _i32p), int(ld.shape[1]), ls_degree, _p(v, _f64p), _p(r.points, _f64p),
                     _p(r.offsets, _i32p), _p(r.parent_map, _i32p), C.c_int64(r.parent_map.size), _p(out, _f64p))
    return out


def ghost_penalty_facets(mesh, cut_cells, selected_cells, include_ghosts=False):
    cc, sc = _ci32(cut_cells), _ci32(selected_cells)
    c2f, off, f2c = _ci32(mesh.c2f), _ci32(mesh.f2c_offsets), _ci32(mesh.f2c)
    out = np.zeros(cc.size * c2f.shape[1] + 1, dtype=np.int32)
    n = lib().orc_ghost_penalty_facets(_p(cc, _i32p), C.c_int64(cc.size), _p(sc, _i32p), C.c_int64(sc.size),
                                       C.c_int64(mesh.num_cells), _p(c2f, _i32p), int(c2f.shape[1]), _p(off, _i32p),
                                       _p(f2c, _i32p), C.c_int64(mesh.num_owned_facets), int(include_ghosts),
                                       _p(out, _i32p))
    return out[:n].copy()


def interior_facets_for_cells(mesh, cells, include_ghosts=False):
    cc = _ci32(cells)
    c2f, off, f2c = _ci32(mesh.c2f), _ci32(mesh.f2c_offsets), _ci32(mesh.f2c)
    out = np.zeros(cc.size * c2f.shape[1] + 1, dtype=np.int32)
    n = lib().orc_interior_facets_for_cells(_p(cc, _i32p), C.c_int64(cc.size), C.c_int64(mesh.num_cells),
                                            _p(c2f, _i32p), int(c2f.shape[1]), _p(off, _i32p), _p(f2c, _i32p),
                                            C.c_int64(mesh.num_owned_facets), int(include_ghosts), _p(out, _i32p))
    return out[:n].copy()


def facet_rows(mesh, facets):
    f = _ci32(facets)
    c2f, off, f2c = _ci32(mesh.c2f), _ci32(mesh.f2c_offsets), _ci32(mesh.f2c)
    rows = np.zeros((f.size, 4), dtype=np.int32)
    _chk(lib().orc_facet_rows(_p(f, _i32p), C.c_int64(f.size), _p(c2f, _i32p), int(c2f.shape[1]), _p(off, _i32p),
                              _p(f2c, _i32p), _p(rows, _i32p)))
    return rows


def sparsity(space, cells, rows4=None, insert_diagonal=True):
    dm = _ci32(space.dofmap)
    cells = _ci32(cells)
    rows4 = _ci32(rows4 if rows4 is not None else np.zeros((0, 4)))
    n_rows = space.num_dofs
    rp = np.zeros(n_rows + 1, dtype=np.int64)
    args = (_p(dm, _i32p), int(dm.shape[1]), C.c_int64(n_rows), _p(cells, _i32p), C.c_int64(cells.size),
            _p(rows4, _i32p), C.c_int64(rows4.shape[0]), int(insert_diagonal), _p(rp, _i64p))
    nnz = lib().orc_sparsity(*args, None)
    cols = np.zeros(nnz, dtype=np.int32)
    lib().orc_sparsity(*args, _p(cols, _i32p))
    return rp, cols


K = {"laplace": 1, "mass": 2, "nitsche": 3, "ghost_grad_jump": 4, "source": 5, "nitsche_rhs": 6, "one": 7,
     "elasticity": 8, "source_vec": 9, "square_fn": 10, "nitsche_vec": 11}
_RANK = {1: 2, 2: 2, 3: 2, 4: 2, 5: 1, 6: 1, 7: 0, 8: 2, 9: 1, 10: 0, 11: 2}


def _register_std_rules(space, kernel_id):
    td, p = space.mesh.tdim, space.degree
    for o in {2 * (p - 1), 2 * p, p, 0}:  # covers every family incl. elasticity (2(p-1)) and source_vec (p)
        _need_rule(td, o)
    _need_rule(td - 1, 2 * (p - 1))


def assemble_cells(space, kernel: str, out, std_cells=None, rules: Rules | None = None, constants=(1.0,),
                   row_ptr=None, cols=None):
    """Adds the integral over entity list [std_cells ++ rules.parent_map] into `out`
    (CSR values for rank 2, vector for rank 1, out[0] for rank 0)."""
    mesh = space.mesh
    kid = K[kernel]
    _register_std_rules(space, kid)
    x, xd, dm = _cf64(mesh.x), _ci32(mesh.x_dofmap), _ci32(space.dofmap)
    sc = _ci32(std_cells if std_cells is not None else [])
    cst = _cf64(list(constants) + [0.0] * 8)
    if rules is not None:
        pts, wts, off, pm, nr = rules.points, rules.weights, rules.offsets, rules.parent_map, rules.parent_map.size
        nrm = rules.normals
    else:
        pts = wts = off = pm = nrm = None
        nr = 0
    if space.bs > 1:
        _chk(lib().orc_assemble_cells_blocked(kid, _RANK[kid], mesh.cell_type, space.degree, int(space.bs),
                                              _p(x, _f64p), _p(xd, _i32p), _p(dm, _i32p), _p(sc, _i32p),
                                              C.c_int64(sc.size), _p(pts, _f64p), _p(wts, _f64p), _p(off, _i32p),
                                              _p(pm, _i32p), C.c_int64(nr), _p(nrm, _f64p), _p(cst, _f64p),
                                              _p(row_ptr, _i64p), _p(cols, _i32p), _p(out, _f64p)))
        return out
    _chk(lib().orc_assemble_cells(kid, _RANK[kid], mesh.cell_type, space.degree, _p(x, _f64p), _p(xd, _i32p),
                                  _p(dm, _i32p), _p(sc, _i32p), C.c_int64(sc.size), _p(pts, _f64p), _p(wts, _f64p),
                                  _p(off, _i32p), _p(pm, _i32p), C.c_int64(nr), _p(nrm, _f64p), _p(cst, _f64p),
                                  _p(row_ptr, _i64p), _p(cols, _i32p), _p(out, _f64p)))
    return out


def assemble_interior_facets(space, kernel: str, vals, rows4, constants, row_ptr, cols):
    mesh = space.mesh
    kid = K[kernel]
    _register_std_rules(space, kid)
    x, xd, dm = _cf64(mesh.x), _ci32(mesh.x_dofmap), _ci32(space.dofmap)
    rows4 = _ci32(rows4)
    cst = _cf64(list(constants) + [0.0] * 8)
    if space.bs > 1:
        _chk(lib().orc_assemble_interior_facets_blocked(kid, mesh.cell_type, space.degree, int(space.bs),
                                                        _p(x, _f64p), _p(xd, _i32p), _p(dm, _i32p), _p(rows4, _i32p),
                                                        C.c_int64(rows4.shape[0]), _p(cst, _f64p), _p(row_ptr, _i64p),
                                                        _p(cols, _i32p), _p(vals, _f64p)))
        return vals
    _chk(lib().orc_assemble_interior_facets(kid, mesh.cell_type, space.degree, _p(x, _f64p), _p(xd, _i32p),
                                            _p(dm, _i32p), _p(rows4, _i32p), C.c_int64(rows4.shape[0]),
                                            _p(cst, _f64p), _p(row_ptr, _i64p), _p(cols, _i32p), _p(vals, _f64p)))
    return vals


# ---------------------------------------------------------------- active domain / deactivation
def active_domain(space, cell_lists, rows4=None):
    """cpp/cutfemx/fem/deactivate.h:103-185,387-400: active cells = sorted unique owned cells of all
    integral domains (both cells of interior-facet rows); indicator = dofs of those cells; inactive
    dofs = owned dofs with indicator 0."""
    nco = space.mesh.num_cells_local
    parts = [np.asarray(c, dtype=np.int64).ravel() for c in cell_lists if c is not None]
    if rows4 is not None and len(rows4):
        r = np.asarray(rows4, dtype=np.int64).reshape(-1, 4)
        parts += [r[:, 0], r[:, 2]]
    cells = np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.int64)
    cells = cells[(cells >= 0) & (cells < nco)].astype(np.int32)
    bs = int(getattr(space, "bs", 1))
    indicator = np.zeros(space.num_dofs * bs)
    for c in cells:  # mark_cell_dofs (deactivate.h:37-46): values[bs * dof + k] = 1
        for k in range(bs):
            indicator[space.dofmap[c] * bs + k] = 1.0
    inactive = np.nonzero(np.abs(indicator[: space.num_dofs_owned * bs]) < 1.0e-8)[0].astype(np.int32)
    return cells, inactive


def deactivate_outside(row_ptr, cols, vals, inactive_dofs, diagonal=1.0, b=None, rhs_value=0.0, bs=1):
    """deactivate.h:402-418: dolfinx::fem::set_diagonal (sets A[r][r]) and b[r] = rhs_value."""
    if bs > 1:
        set_diagonal(row_ptr, cols, vals, inactive_dofs, diagonal, bs)
        if b is not None:
            b[np.asarray(inactive_dofs, dtype=np.int64)] = rhs_value
        return vals
    for r in inactive_dofs:
        seg = cols[row_ptr[r]:row_ptr[r + 1]]
        k = int(np.searchsorted(seg, r))
        if k >= seg.size or seg[k] != r:
            raise RuntimeError("inactive row has no diagonal entry")
        vals[row_ptr[r] + k] = diagonal
        if b is not None:
            b[r] = rhs_value
    return vals


# ---------------------------------------------------------------- Dirichlet conditions
class dirichlet:
    """Context for the orc_assemble_* calls inside it (rank 2 only).

    mode "matrix": assemble_matrix(A, a, bcs) -- rows bc0 / columns bc1 of every element tensor are zeroed
    before mat_set (assemble_matrix_impl.h:146-185).  mode "lifting": apply_lifting -- the same loops in
    LiftingMode, b -= alpha * Ae[:, bc] (x_bc - x0) (assemble_vector_impl.h:383-439); nothing reaches the matrix.
    Markers are int8 per (blocked) dof index bs*dof + k."""

    def __init__(self, mode, bc0=None, bc1=None, values1=None, x0=None, alpha=1.0, b=None):
        self.mode = {"matrix": 1, "lifting": 2}[mode]
        self.bc0 = None if bc0 is None else np.ascontiguousarray(bc0, dtype=np.int8)
        self.bc1 = None if bc1 is None else np.ascontiguousarray(bc1, dtype=np.int8)
        self.values1 = None if values1 is None else _cf64(values1)
        self.x0 = None if x0 is None else _cf64(x0)
        self.alpha = float(alpha)
        self.b = b
        if self.mode == 2:
            assert self.bc1 is not None and self.values1 is not None and b is not None
            assert b.dtype == np.float64 and b.flags.c_contiguous

    def __enter__(self):
        q = lambda a, t: None if a is None else _p(a, t)
        lib().orc_set_bcs(self.mode, q(self.bc0, _i8p), q(self.bc1, _i8p), q(self.values1, _f64p), q(self.x0, _f64p),
                          C.c_double(self.alpha), q(self.b, _f64p))
        return self

    def __exit__(self, *exc):
        lib().orc_set_bcs(0, None, None, None, None, C.c_double(1.0), None)
        return False


def set_diagonal(row_ptr, cols, vals, rows, diagonal=1.0, bs=1):
    """assembler.h:745-753 with MatrixCSR::mat_set_values: A[r, r] = diagonal for every listed (blocked) row.
    rows are blocked dof indices bs*dof + k (DirichletBC::dof_indices, unrolled)."""
    for r in np.asarray(rows, dtype=np.int64):
        blk, k = divmod(int(r), bs)
        seg = cols[row_ptr[blk]:row_ptr[blk + 1]]
        p = int(np.searchsorted(seg, blk))
        if p >= seg.size or seg[p] != blk:
            raise RuntimeError("Dirichlet row has no diagonal entry")
        vals[(row_ptr[blk] + p) * bs * bs + k * bs + k] = diagonal
    return vals


def set_bc(b, dofs, values, x0=None, alpha=1.0):
    """DirichletBC::set (dolfinx 0.11, fem.set_bc): b[d] = alpha * (g[d] - x0[d]) on the constrained dofs."""
    d = np.asarray(dofs, dtype=np.int64)
    g = np.asarray(values, dtype=np.float64)
    b[d] = alpha * (g[d] - (0.0 if x0 is None else np.asarray(x0)[d]))
    return b


class coefficient:
    """Context: the ordinary Function coefficient of the orc_assemble_cells calls inside it (dof values over the
    space; packed per entity by the loop like pack_coefficients, pack_form.h:68-158)."""

    def __init__(self, values):
        self.values = _cf64(values)

    def __enter__(self):
        lib().orc_set_coefficient(_p(self.values, _f64p))
        return self

    def __exit__(self, *exc):
        lib().orc_set_coefficient(None)
        return False


# ---------------------------------------------------------------- facets as hosts of the cut
def facet_vertices(mesh, facets):
    """Vertices of the listed facets, ascending vertex number per facet (DOLFINx keeps entity vertices sorted;
    build_entity_mesh_view, cut.cpp:540-591), and for each the local vertex index in the facet's first cell."""
    nv = mesh.x_dofmap.shape[1]
    verts = np.zeros((len(facets), nv - 1), dtype=np.int32)
    local = np.zeros((len(facets), nv - 1), dtype=np.int32)
    cells = np.zeros(len(facets), dtype=np.int64)
    for i, f in enumerate(np.asarray(facets, dtype=np.int64)):
        c = int(mesh.f2c[mesh.f2c_offsets[f]])
        lf = int(np.nonzero(mesh.c2f[c] == f)[0][0])
        lj = np.array([j for j in range(nv) if j != lf])  # facet lf is opposite local vertex lf
        v = mesh.x_dofmap[c][lj]
        o = np.argsort(v, kind="stable")
        verts[i], local[i], cells[i] = v[o], lj[o], c
    return verts, local, cells


def facet_rules_on_cells(mesh, r: "Rules") -> "Rules":
    """Facet-hosted rules re-expressed on the facets' first cells -- what _facet_payload_with_rows and runintgen's
    facet_runtime_quadrature_payload (python/cutfemx/_runintgen_adapter.py:605-680) do before an exterior-facet kernel
    runs: facet reference point xi -> barycentric (1 - sum xi, xi) over the facet's vertices (ascending vertex number)
    -> reference coordinates of the cell, where the vertex with local index t >= 1 is the unit vector e_{t-1}.
    Rules keep the order of the facet rules; parent_map = cells (a cell may appear more than once)."""
    verts, local, cells = facet_vertices(mesh, r.parent_map)
    tdim = mesh.tdim
    rule_of_pt = np.repeat(np.arange(r.parent_map.size), np.diff(r.offsets))
    lam = np.concatenate([1.0 - r.points.sum(axis=1, keepdims=True), r.points], axis=1)
    X = np.zeros((r.weights.size, tdim))
    for j in range(tdim):            # facet vertex j
        loc = local[rule_of_pt, j]
        for t in range(tdim):
            X[:, t] += np.where(loc == t + 1, lam[:, j], 0.0)
    return Rules(tdim, X, r.weights.copy(), r.offsets.copy(), cells.astype(np.int32))


def classify_facets(mesh, ls_dofmap, values, facets):
    """Domain codes of facet hosts from the level-set values at their vertices (P1: the entity dofmap of
    fem/entity_dofmap.cpp:11-88 lists the vertices' dofs) -- the rule of cut.cpp:292-321 one dimension down."""
    verts, local, cells = facet_vertices(mesh, facets)
    phi = np.asarray(values)[np.asarray(ls_dofmap)[cells[:, None], local]]
    code = np.full(len(facets), INTERSECTED, dtype=np.int8)
    code[np.all(phi < 0.0, axis=1)] = INSIDE
    code[np.all(phi > 0.0, axis=1)] = OUTSIDE
    return code, verts, phi


def facet_runtime_quadrature(mesh, ls_dofmap, values, facets, relation: str, order: int) -> "Rules":
    """Rules on the cut facets in the facets' own reference coordinates with physical weights (cutfemx.cut(...,
    entity_dim = tdim - 1) + runtime_quadrature, test_cut_api.py:424-501).  Segments (2D meshes) in closed form;
    triangles (3D meshes) through the cell generator applied to an ISOMETRIC planar copy of each facet -- reference
    points do not depend on the embedding and the weights only on lengths and areas."""
    code, verts, phi = classify_facets(mesh, ls_dofmap, values, facets)
    edim = verts.shape[1] - 1
    positive = relation in (">", ">=")
    X = mesh.x[verts][:, :, : edim + 1]
    cutf = np.nonzero(code == INTERSECTED)[0]
    if edim == 1:
        p, w = _rules.simplex_rule(1, order)
        p, w = np.asarray(p).reshape(-1), np.asarray(w)
        pts, wts, off, pm = [], [], [0], []
        for i in cutf:
            inside = (phi[i] > 0.0) if positive else (phi[i] < 0.0)
            if inside.sum() != 1:
                continue
            a = int(np.nonzero(inside)[0][0])
            b = 1 - a
            t = phi[i, a] / (phi[i, a] - phi[i, b])
            xa, xc = float(a), float(a) + t * (float(b) - float(a))  # reference coordinate of vertex q is q
            length = np.linalg.norm(X[i, 1] - X[i, 0])
            if relation == "=":  # the interface inside a segment is the cut point, counting measure
                pts.append(np.array([xc]))
                wts.append(np.array([1.0]))
                off.append(off[-1] + 1)
                pm.append(int(np.asarray(facets)[i]))
                continue
            pts.append(xa + p * (xc - xa))
            wts.append(w * abs(xc - xa) * length)
            off.append(off[-1] + w.size)
            pm.append(int(np.asarray(facets)[i]))
        P = np.concatenate(pts).reshape(-1, 1) if pts else np.zeros((0, 1))
        W = np.concatenate(wts) if wts else np.zeros(0)
        return Rules(1, P, W, np.asarray(off, dtype=np.int32), np.asarray(pm, dtype=np.int32))
    # triangles: planar isometric copies as a pseudo mesh of independent cells
    from types import SimpleNamespace

    e1, e2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
    l1 = np.linalg.norm(e1, axis=1)
    u = e1 / l1[:, None]
    a2 = np.einsum("ij,ij->i", e2, u)
    h2 = np.linalg.norm(e2 - a2[:, None] * u, axis=1)
    n = len(facets)
    xp = np.zeros((3 * n, 3))
    xp[1::3, 0] = l1
    xp[2::3, 0], xp[2::3, 1] = a2, h2
    xd = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    pseudo = SimpleNamespace(x=xp, x_dofmap=xd, cell_type=3, tdim=2, gdim=2, num_cells=n, num_cells_local=n)
    r = runtime_quadrature(pseudo, xd, phi.reshape(-1), code, relation, order)
    return Rules(2, r.points, r.weights, r.offsets, np.asarray(facets, dtype=np.int32)[r.parent_map])
