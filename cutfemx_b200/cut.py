"""Host-side mirror of `cutfemx.cut` (python/cutfemx/cut.py of the reference) on top of the C ABI.

Same names, argument meaning and error behaviour as the reference for the hot path:
`cut`, `update`, `locate_entities`, `runtime_quadrature(s)`, `ghost_penalty_facets`,
`interior_facets_for_cells`; the DOLFINx objects are replaced by the flat-array stand-ins of
`cutfemx_b200.mesh` (Mesh / FunctionSpace / Function).  All arithmetic runs in the sm_100a
library; this module only marshals arrays (numpy = host, torch CUDA tensors = device).
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections.abc import Sequence

import numpy as np

from . import _lib
from ._lib import CfxError, DEVICE, HOST, as_arg, check, lib, parse_selector
from .mesh import Function, Mesh

_DEFAULT_NAMES = ("", "u", "f")  # cut.cpp:59-62: unnamed level sets become phi, phi1, ...


class Context:
    """One cfx_ctx per (mesh, GPU)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        check(None, lib().cfx_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(self._h)))
        self.device = int(device)
        self._keep = []
        self._spaces = {}

    def close(self):
        if self._h:
            lib().cfx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(self._h, lib().cfx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().cfx_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(lib().cfx_device_bytes(self._h))

    # -- deferred sizes / CUDA graphs (include/cutfemx_b200.h "deferred sizes and CUDA graphs")
    def set_deferred(self, on: bool, margin: float = -1.0):
        check(self._h, lib().cfx_set_deferred(self._h, int(on), float(margin)))

    def check(self):
        """Raise if a deferred-size call exceeded a capacity or met an invalid index (synchronises)."""
        check(self._h, lib().cfx_check(self._h))

    def graph_begin(self):
        check(self._h, lib().cfx_graph_begin(self._h))

    def graph_end(self) -> "Graph":
        g = Graph(self)
        check(self._h, lib().cfx_graph_end(self._h, C.byref(g._h)))
        return g

    # -- lanes: independent call sequences of a step on streams of their own (cfx_lane_begin / _end / _join)
    def lane(self, k: int):
        """`with ctx.lane(k): ...` -- the calls inside run on lane k (1..3), concurrently with the other lanes and
        with later main-stream calls; `ctx.join()` before anything uses their results."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            if k <= 0:
                yield
                return
            check(self._h, lib().cfx_lane_begin(self._h, int(k)))
            try:
                yield
            finally:
                check(self._h, lib().cfx_lane_end(self._h))
        return scope()

    def join(self):
        check(self._h, lib().cfx_lane_join(self._h))

    def set_lanes(self, on: bool):
        """off: lanes (the library's internal ones included) collapse onto the main stream -- one kernel at a time."""
        check(self._h, lib().cfx_set_lanes(self._h, int(on)))

    # -- profiling hooks
    def stage_timing(self, on: bool):
        check(self._h, lib().cfx_stage_timing_enable(self._h, int(on)))

    def stage_reset(self):
        check(self._h, lib().cfx_stage_reset(self._h))

    def stages(self):
        out = []
        for i in range(lib().cfx_stage_count(self._h)):
            ms, by = C.c_double(), C.c_double()
            check(self._h, lib().cfx_stage_ms(self._h, i, C.byref(ms), C.byref(by)))
            out.append((lib().cfx_stage_name(self._h, i).decode(), ms.value, by.value))
        return out

    def space_index(self, space) -> int:
        """Bind a FunctionSpace to a context slot (once)."""
        key = id(space)
        if key in self._spaces:
            return self._spaces[key][0]
        idx = len(self._spaces)
        if idx >= 4:
            raise CfxError("too many function spaces bound to one context")
        p, ms, keep = as_arg(space.dofmap, np.int32)
        check(self._h, lib().cfx_space_bind(self._h, idx, p, int(space.nd), int(space.bs), int(space.degree),
                                            C.c_int64(space.num_dofs_owned), C.c_int64(space.num_dofs), ms))
        self._spaces[key] = (idx, space, keep)
        return idx


class Graph:
    """A captured step (cfx_graph): `launch()` replays every call recorded between graph_begin and graph_end."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._h = C.c_void_p()

    def launch(self):
        check(self.ctx.handle, lib().cfx_graph_launch(self.ctx.handle, self._h))

    @property
    def kernel_nodes(self) -> int:
        return int(lib().cfx_graph_kernel_nodes(self._h))

    def free(self):
        if self._h:
            lib().cfx_graph_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


def _mesh_context(mesh: Mesh, device: int | None = None) -> Context:
    ctx = mesh.extra.get("_cfx_ctx")
    if ctx is None:
        if device is None:
            device = int(mesh.extra.get("device", 0))
        ctx = Context(device)
        h = ctx.handle
        px, msx, kx = as_arg(mesh.x, np.float64)
        pd, msd, kd = as_arg(mesh.x_dofmap, np.int32)
        if msx != msd:
            raise CfxError("mesh geometry arrays must live in the same memory space")
        nn, ncell = int(mesh.x.shape[0]), int(mesh.x_dofmap.shape[0])
        check(h, lib().cfx_mesh_bind(h, px, C.c_int64(nn), pd, C.c_int64(mesh.num_cells_local), C.c_int64(ncell),
                                     int(mesh.cell_type), int(mesh.gdim), msx))
        ctx._keep += [kx, kd]
        mesh.extra["_cfx_ctx"] = ctx
    return ctx


def _bind_topology(mesh: Mesh, ctx: Context):
    if mesh.extra.get("_cfx_topo"):
        return
    if mesh.c2f is None:
        raise RuntimeError("Facet-cell connectivity is unavailable.")  # cut.py:361-362
    pc, ms, kc = as_arg(mesh.c2f, np.int32)
    po, mso, ko = as_arg(mesh.f2c_offsets, np.int32)
    pf, msf, kf = as_arg(mesh.f2c, np.int32)
    if mesh.f2c_offsets is None or mesh.f2c is None:
        po = pf = None
    h = ctx.handle
    check(h, lib().cfx_topology_bind(h, pc, po, pf, C.c_int64(mesh.num_facets), C.c_int64(mesh.num_owned_facets), ms))
    ctx._keep += [kc, ko, kf]
    mesh.extra["_cfx_topo"] = True


class _List:
    """Device-resident int32 list returned by the library."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._h = C.c_void_p()

    @property
    def size(self) -> int:
        n = int(lib().cfx_list_size(self._h))
        if n < 0:  # a deferred size could not be fetched (capacity exceeded in the step that produced the list)
            check(self.ctx.handle, -4)
        return n

    @property
    def device_ptr(self) -> int:
        return int(lib().cfx_list_device_ptr(self._h) or 0)

    def numpy(self) -> np.ndarray:
        out = np.empty(self.size, dtype=np.int32)
        if out.size:
            check(self.ctx.handle, lib().cfx_list_fetch(self.ctx.handle, self._h, C.c_void_p(out.ctypes.data), HOST))
        return out

    def torch(self):
        import torch

        out = torch.empty(self.size, dtype=torch.int32, device=f"cuda:{self.ctx.device}")
        if out.numel():
            check(self.ctx.handle, lib().cfx_list_fetch(self.ctx.handle, self._h, C.c_void_p(out.data_ptr()), DEVICE))
        return out

    def free(self):
        if self._h:
            lib().cfx_list_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class SurfaceProvenance:
    """cutfemx::RuntimeSurfaceProvenance (runtime_quadrature.h:30-43)."""

    def __init__(self, selector, level_set_index, cut_cell_ids, parent_cell_ids, local_zero_entity_ids, dimensions):
        self.selector, self.level_set_index = selector, level_set_index
        self.cut_cell_ids, self.parent_cell_ids = cut_cell_ids, parent_cell_ids
        self.local_zero_entity_ids, self.dimensions = local_zero_entity_ids, dimensions

    def empty(self) -> bool:
        return self.cut_cell_ids.size == 0

    def size(self) -> int:
        return int(self.cut_cell_ids.size)


class RuntimeQuadratureRules:
    """Mirror of the runintgen `QuadratureRules` subclass the reference returns
    (python/cutfemx/cut.py:22-57, wrappers/cut.cpp:185-240): `kind`, `tdim`, `points` (npts, tdim),
    `weights`, `offsets` (int32, nrules+1), `parent_map` (int32), lazy `physical_points`
    (gdim, npts).  Arrays are fetched from the device on first access; the device copy stays
    attached (`_h`) and is what the assembly kernels read."""

    kind = "per_entity"

    def __init__(self, ctx: Context, selector: str, ls: int, order: int):
        self.ctx, self.selector, self.ls, self.order = ctx, selector, ls, order
        self._h = C.c_void_p()
        self._cache = {}
        self.normal_sign = None

    def _sizes(self):
        npts, nr, td = C.c_int64(), C.c_int64(), C.c_int()
        check(self.ctx.handle, lib().cfx_rules_sizes(self._h, C.byref(npts), C.byref(nr), C.byref(td)))
        return npts.value, nr.value, td.value

    @property
    def tdim(self) -> int:
        return self._sizes()[2]

    @property
    def total_points(self) -> int:
        return self._sizes()[0]

    @property
    def num_rules(self) -> int:
        return self._sizes()[1]

    def _fetch(self):
        if "weights" in self._cache:
            return
        npts, nr, td = self._sizes()
        pts = np.empty((npts, td))
        wts = np.empty(npts)
        off = np.empty(nr + 1, dtype=np.int32)
        pm = np.empty(nr, dtype=np.int32)
        h = self.ctx.handle
        check(h, lib().cfx_rules_fetch(h, self._h, C.c_void_p(pts.ctypes.data), C.c_void_p(wts.ctypes.data),
                                       C.c_void_p(off.ctypes.data), C.c_void_p(pm.ctypes.data), HOST))
        self._cache.update(points=pts, weights=wts, offsets=off, parent_map=pm)

    @property
    def points(self):
        self._fetch()
        return self._cache["points"]

    @property
    def weights(self):
        self._fetch()
        return self._cache["weights"]

    @property
    def offsets(self):
        self._fetch()
        return self._cache["offsets"]

    @property
    def parent_map(self):
        self._fetch()
        return self._cache["parent_map"]

    @property
    def physical_points(self):
        if "physical_points" not in self._cache:
            npts, _, _ = self._sizes()
            gdim = self.ctx_gdim
            out = np.zeros((gdim, npts))
            if npts:
                h = self.ctx.handle
                check(h, lib().cfx_rules_physical_points(h, self._h, C.c_void_p(out.ctypes.data), HOST))
            self._cache["physical_points"] = out
        return self._cache["physical_points"]

    def with_physical_points(self):
        _ = self.physical_points
        return self

    @property
    def surface_provenance(self):
        """RuntimeSurfaceProvenance (runtime_quadrature.h:30-43, cut.cpp:1273-1308): selector, level_set_index and
        -- for a single "phi = 0" selector -- per rule cut_cell_ids, parent_cell_ids, local_zero_entity_ids,
        dimensions (int32).  Empty (level_set_index = -1, zero-length arrays) for every other selector."""
        if "provenance" not in self._cache:
            _, nr, _ = self._sizes()
            ls = C.c_int32(-1)
            h = self.ctx.handle
            check(h, lib().cfx_rules_surface_provenance(h, self._h, C.byref(ls), None, None, None, None, HOST))
            n = nr if ls.value >= 0 else 0
            arrs = [np.empty(n, dtype=np.int32) for _ in range(4)]
            if n:
                check(h, lib().cfx_rules_surface_provenance(h, self._h, C.byref(ls),
                                                            *[C.c_void_p(a.ctypes.data) for a in arrs], HOST))
            self._cache["provenance"] = SurfaceProvenance(self.selector, int(ls.value), *arrs)
        return self._cache["provenance"]

    def free(self):
        if self._h:
            lib().cfx_rules_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class CutData:
    """Python handle for cut data (python/cutfemx/cut.py:91-142)."""

    def __init__(self, ctx: Context, level_sets: Sequence[Function], names: tuple):
        self._ctx = ctx
        self._level_sets = tuple(level_sets)
        self._names = tuple(names)
        self._keep = []
        self._entities, self._entity_dim = None, None  # python/cutfemx/cut.py:102-108,141-146
        self._ecut = C.c_void_p()  # facet-hosted cut (cfx_cut_facets)
        self._host_cells = None    # cell subset acting as host (int32 array), or None for every owned cell

    # The context (device mirrors of the mesh, classification codes, cut lists) is shared by every CutData of a
    # mesh, but a reference CutData owns its level-set view, its host entities and its cut cells (cut.cpp:742-786):
    # two of them must not alias.  The context remembers which CutData its level-set slots, host mask and
    # classification belong to; a query through another CutData re-binds and re-classifies first.
    def _bind(self) -> None:
        ctx, h = self._ctx, self._ctx.handle
        ctx._active_cut = None  # a bind that fails half-way leaves nobody active: the next query re-binds
        ent = self._host_cells
        if ent is None:
            check(h, lib().cfx_set_host_cells(h, None, C.c_int64(0), HOST))
        else:
            check(h, lib().cfx_set_host_cells(h, C.c_void_p(ent.ctypes.data), C.c_int64(ent.size), HOST))
        self._keep = []
        for i, f in enumerate(self._level_sets):
            V = f.function_space
            pd, msd, kd = as_arg(V.dofmap, np.int32)
            pv, msv, kv = as_arg(f.x.array, np.float64)
            if msd != msv:
                # dofmap is copied/adopted separately from the values; mixed spaces need two calls' worth of
                # bookkeeping the C ABI does not expose -- move the dofmap to where the values are.
                if msv == DEVICE:
                    import torch

                    kd = torch.as_tensor(np.ascontiguousarray(V.dofmap), device=f"cuda:{ctx.device}")
                    pd, msd = C.c_void_p(kd.data_ptr()), DEVICE
                else:
                    kd = V.dofmap.cpu().numpy()
                    pd, msd = C.c_void_p(kd.ctypes.data), HOST
            check(h, lib().cfx_levelset_bind(h, i, pd, int(V.nd), int(V.degree), pv, C.c_int64(V.num_dofs), msv, 1))
            self._keep += [kd, kv]
        for i in range(len(self._level_sets), getattr(ctx, "_n_bound_ls", 0)):
            check(h, lib().cfx_levelset_unbind(h, i))
        ctx._n_bound_ls = len(self._level_sets)
        ctx._active_cut = weakref.ref(self)

    def _activate(self) -> None:
        """Make the context answer for THIS CutData (no-op while it is the one bound last)."""
        ref = getattr(self._ctx, "_active_cut", None)
        if ref is not None and ref() is self:
            return
        self._bind()
        update(self)

    @property
    def facet_hosted(self) -> bool:
        return self._entity_dim is not None and self._entity_dim == self.mesh.tdim - 1

    def __del__(self):
        try:
            if self._ecut and self._ctx._h:
                lib().cfx_ecut_free(self._ctx.handle, self._ecut)
                self._ecut = C.c_void_p()
            ref = getattr(self._ctx, "_active_cut", None)
            if self._ctx._h and ref is not None and ref() in (self, None):
                # the context must not keep pointers into this CutData's value arrays (or their page-locking)
                for i in range(getattr(self._ctx, "_n_bound_ls", 0)):
                    lib().cfx_levelset_unbind(self._ctx.handle, i)
                self._ctx._n_bound_ls = 0
                self._ctx._active_cut = None
        except Exception:
            pass

    def update(self) -> None:
        update(self)

    @property
    def tdim(self) -> int:  # dimension of the host entities (test_cut_api.py:361: cutter.tdim == 1 for facets)
        return self._entity_dim if self.facet_hosted else self.mesh.tdim

    @property
    def gdim(self) -> int:
        return self.mesh.gdim

    @property
    def num_local_cells(self) -> int:
        return self.mesh.num_cells_local

    @property
    def level_set_names(self) -> tuple:
        return self._names

    @property
    def level_sets(self) -> tuple:
        return self._level_sets

    @property
    def mesh(self) -> Mesh:
        return self._level_sets[0].function_space.mesh

    @property
    def entity_dim(self):
        return self._entity_dim

    @property
    def entities(self):
        return self._entities

    def counts(self, ls: int = 0):
        """(inside, intersected, outside) owned-cell counts."""
        self._activate()
        out = (C.c_int64 * 3)()
        check(self._ctx.handle, lib().cfx_counts(self._ctx.handle, ls, out))
        return tuple(int(v) for v in out)

    def domain_codes(self, ls: int = 0) -> np.ndarray:
        self._activate()
        out = np.empty(self.mesh.x_dofmap.shape[0], dtype=np.int8)
        check(self._ctx.handle, lib().cfx_domain_fetch(self._ctx.handle, ls, C.c_void_p(out.ctypes.data), HOST))
        return out


def _normalise_level_sets(level_set):
    # python/cutfemx/cut.py:159-183
    if isinstance(level_set, Function):
        return [level_set]
    if isinstance(level_set, (str, bytes)) or not isinstance(level_set, Sequence):
        raise TypeError("cutfemx.cut expects a Function or a non-empty sequence of Functions")
    level_sets = list(level_set)
    if not level_sets:
        raise ValueError("cutfemx.cut requires at least one level-set function")
    for item in level_sets:
        if not isinstance(item, Function):
            raise TypeError("cutfemx.cut sequence entries must be Function objects")
    return level_sets


def _freeze_names(level_sets):
    # cut.cpp:81-137: unnamed -> phi, phi1, phi2...; duplicates are an error
    names = []
    for i, f in enumerate(level_sets):
        n = f.name
        if n in _DEFAULT_NAMES:
            n = "phi" if i == 0 else f"phi{i}"
        if n in names:
            raise ValueError(f"duplicate level-set name '{n}'")
        names.append(n)
    return tuple(names)


def cut(level_set, entities=None, entity_dim=None, *, cut_approximation: str = "auto",
        cut_approximation_order: int = 1, max_refinement_iterations: int = 8, edge_max_depth: int = 20,
        device: int | None = None) -> CutData:
    """Cut one or more scalar level-set functions on the cells of their mesh
    (python/cutfemx/cut.py:186-249 -> cut.cpp:639-659, 742-786, 845-868)."""
    # python/cutfemx/cut.py:149-160
    if entities is None and entity_dim is not None:
        raise ValueError("entity_dim is only valid when entities are supplied")
    if entities is not None and entity_dim is None:
        raise ValueError("entity_dim must be supplied when entities are supplied")
    if cut_approximation not in ("auto", "linear") or cut_approximation_order != 1:
        raise NotImplementedError("only the straight (order 1) cut approximation is implemented")
    level_sets = _normalise_level_sets(level_set)
    mesh = level_sets[0].function_space.mesh
    for f in level_sets:
        if f.function_space.mesh is not mesh:
            raise ValueError("all level sets must live on the same mesh")  # cut.cpp:462-498
        if f.function_space.bs != 1:
            raise ValueError("level set must be a scalar Lagrange function")  # cut.cpp:444-460
    if len(level_sets) > 4:
        raise ValueError("at most 4 level sets per CutData")
    if entities is not None and int(entity_dim) not in (mesh.tdim, mesh.tdim - 1):
        # cut.cpp:545-550 accepts every positive entity dimension; edges of tetrahedra are not implemented here
        if int(entity_dim) <= 0 or int(entity_dim) > mesh.tdim:
            raise ValueError("cutfemx::cut entity_dim must select positive-dimensional mesh entities")
        raise NotImplementedError("entity-hosted cuts are implemented for cells and facets")
    names = _freeze_names(level_sets)
    ctx = _mesh_context(mesh, device)
    cd = CutData(ctx, level_sets, names)
    h = ctx.handle
    if entities is not None and int(entity_dim) == mesh.tdim - 1:
        # facets as hosts (cut.cpp:540-591, 1022-1063): classified after the level sets are bound (update)
        cd._entities, cd._entity_dim = np.ascontiguousarray(np.asarray(entities, dtype=np.int32)), int(entity_dim)
    elif entities is None:
        cd._entities, cd._entity_dim = None, None
    else:  # cell subset as host (cut.cpp:500-538, test_cut_api.py:160-168); lists come back ascending
        ent = np.ascontiguousarray(np.asarray(entities, dtype=np.int32))
        cd._entities, cd._entity_dim, cd._host_cells = ent, int(entity_dim), ent
    cd._bind()
    update(cd)
    return cd


def update(cut_data: CutData) -> None:
    """Refresh cut data from the current level-set values (cut.cpp:845-868)."""
    h = cut_data._ctx.handle
    ref = getattr(cut_data._ctx, "_active_cut", None)
    if ref is None or ref() is not cut_data:
        cut_data._bind()
    check(h, lib().cfx_update(h))
    if cut_data.facet_hosted:
        _bind_topology(cut_data.mesh, cut_data._ctx)
        ent = cut_data._entities
        check(h, lib().cfx_cut_facets(h, C.c_void_p(ent.ctypes.data), C.c_int64(ent.size), HOST,
                                      C.byref(cut_data._ecut)))


def _selector_args(cut_data: CutData, ls_part: str):
    to, cl, cr = parse_selector(ls_part, cut_data.level_set_names)
    return (int(to.size - 1), C.c_void_p(to.ctypes.data), C.c_void_p(cl.ctypes.data), C.c_void_p(cr.ctypes.data),
            (to, cl, cr))


def locate_entities_device(cut_data: CutData, ls_part: str, out: _List | None = None) -> _List:
    """`out`: a list from an earlier call to refill in place (its buffer is reused; in deferred-size mode the new
    length then stays on the device)."""
    cut_data._activate()
    n, pto, pcl, pcr, keep = _selector_args(cut_data, ls_part)
    out = _List(cut_data._ctx) if out is None else out
    h = cut_data._ctx.handle
    if cut_data.facet_hosted:  # facet ids in the order of the host list (cut.cpp:344-359)
        check(h, lib().cfx_ecut_locate(h, cut_data._ecut, n, pto, pcl, pcr, C.byref(out._h)))
    else:
        check(h, lib().cfx_locate_entities(h, n, pto, pcl, pcr, C.byref(out._h)))
    return out


def locate_entities(cut_data: CutData, ls_part: str) -> np.ndarray:
    """Owned cell ids matching the selector (cut.cpp:877-924): ascending for the whole mesh; for a cell subset as
    host in the ORDER OF THE SUBSET LIST, repeated entries included -- the reference walks the entity view, whose
    parent_entities are the caller's list (cut.cpp:574-576, :344-359).  (The device-resident variant,
    locate_entities_device, and the parent_map of run-time rules stay ascending and unique.)"""
    lst = locate_entities_device(cut_data, ls_part)
    out = lst.numpy()
    lst.free()
    ent = cut_data._entities
    if ent is not None and not cut_data.facet_hosted and ent.size and (np.any(np.diff(ent) <= 0)):
        out = ent[np.isin(ent, out)]  # the caller's order and multiplicity
    return out


def runtime_quadrature(cut_data: CutData, ls_part: str, order: int, *, backend: str = "straight", out=None):
    """Run-time quadrature for the selected part (cut.cpp:1311-1335).  `out`: rules from an earlier call to refill
    in place."""
    if backend != "straight":
        # cut.cpp:207-237: algoim backends exist only for interval/quadrilateral/hexahedron cells
        raise ValueError(f"runtime_quadrature backend '{backend}' is not available for simplex cells")
    if order < 0:
        raise ValueError("runtime_quadrature order must be >= 0")  # cut.cpp:164-168
    to, cl, cr = parse_selector(ls_part, cut_data.level_set_names)
    if cl.size != 1:
        raise NotImplementedError("runtime_quadrature supports single-clause selectors (`name rel 0`)")
    if out is not None:
        rules = out
        rules.selector, rules.ls, rules.order = ls_part, int(cl[0]), int(order)
        rules._cache.clear()
    else:
        rules = RuntimeQuadratureRules(cut_data._ctx, ls_part, int(cl[0]), int(order))
    rules.ctx_gdim = cut_data.gdim
    cut_data._activate()
    h = cut_data._ctx.handle
    if cut_data.facet_hosted:
        check(h, lib().cfx_ecut_runtime_quadrature(h, cut_data._ecut, int(cl[0]), int(cr[0]), int(order),
                                                   C.byref(rules._h)))
    else:
        check(h, lib().cfx_runtime_quadrature(h, int(cl[0]), int(cr[0]), int(order), C.byref(rules._h)))
    return rules


def runtime_quadratures(cut_data: CutData, ls_parts: Sequence[str], order: int, *, backend: str = "straight"):
    return {str(p): runtime_quadrature(cut_data, str(p), order, backend=backend) for p in ls_parts}


def ghost_penalty_facets_device(cut_data: CutData, selector: str, *, include_ghosts: bool = False,
                                out: _List | None = None) -> _List:
    mesh = cut_data.mesh
    cut_data._activate()
    _bind_topology(mesh, cut_data._ctx)
    n, pto, pcl, pcr, keep = _selector_args(cut_data, selector)
    if "phi" not in cut_data.level_set_names:
        raise ValueError("ghost_penalty_facets locates the cut cells with 'phi=0' (python/cutfemx/cut.py:364)")
    out = _List(cut_data._ctx) if out is None else out
    h = cut_data._ctx.handle
    check(h, lib().cfx_ghost_penalty_facets(h, cut_data.level_set_names.index("phi"), n, pto, pcl, pcr,
                                            int(include_ghosts), C.byref(out._h)))
    return out


def ghost_penalty_facets(cut_data: CutData, selector: str, *, depth: int = 1,
                         include_ghosts: bool = False) -> np.ndarray:
    """Owned raw interior facet ids of the cut-cell stabilisation band (cut.py:340-380)."""
    if depth != 1:
        raise NotImplementedError("ghost_penalty_facets currently supports depth=1.")
    if cut_data.entity_dim is not None and cut_data.entity_dim != cut_data.mesh.tdim:
        raise ValueError("ghost_penalty_facets expects cell-hosted CutData.")  # cut.py:350-351
    lst = ghost_penalty_facets_device(cut_data, selector, include_ghosts=include_ghosts)
    out = lst.numpy()
    lst.free()
    return out


def interior_facets_for_cells(msh: Mesh, cells, *, include_ghosts: bool = False) -> np.ndarray:
    """Raw local interior facet ids whose adjacent cells are both in `cells` (cut.cpp:926-994)."""
    ctx = _mesh_context(msh)
    _bind_topology(msh, ctx)
    p, ms, keep = as_arg(np.ascontiguousarray(np.asarray(cells, dtype=np.int32).ravel())
                         if not _lib.is_device_array(cells) else cells, np.int32)
    n = int(keep.numel() if _lib.is_device_array(keep) else keep.size)
    out = _List(ctx)
    check(ctx.handle, lib().cfx_interior_facets_for_cells(ctx.handle, p, C.c_int64(n), ms, int(include_ghosts),
                                                          C.byref(out._h)))
    res = out.numpy()
    out.free()
    return res


def facet_integration_rows_device(msh: Mesh, facets, out: _List | None = None) -> _List:
    ctx = _mesh_context(msh)
    _bind_topology(msh, ctx)
    out = _List(ctx) if out is None else out
    if isinstance(facets, _List):  # stays on the device; the length may be deferred
        check(ctx.handle, lib().cfx_facet_integration_rows_list(ctx.handle, facets._h, C.byref(out._h)))
        out._keep = facets
        return out
    p, ms, keep = as_arg(facets, np.int32)
    n = int(keep.numel() if _lib.is_device_array(keep) else keep.size)
    check(ctx.handle, lib().cfx_facet_integration_rows(ctx.handle, p, C.c_int64(n), ms, C.byref(out._h)))
    return out


def facet_integration_rows(msh: Mesh, facets) -> np.ndarray:
    """(cell0, local_facet0, cell1, local_facet1) per interior facet (wrappers/cut.cpp:54-115)."""
    lst = facet_integration_rows_device(msh, facets)
    out = lst.numpy().reshape(-1, 4)
    lst.free()
    return out
