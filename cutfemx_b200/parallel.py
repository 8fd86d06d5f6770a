"""One rank per GPU: DOLFINx-style partitions, index maps and the ghost exchange.

What this module stands in for (all third-party DOLFINx 0.11 behaviour the reference relies on,
SURVEY.md section 8e):

* the MPI cell partition with a `shared_facet` ghost layer and `[owned | ghost]` local numbering
  of cells, facets and dofs (`partition_slab`: z-slabs in 3D / y-strips in 2D of the synthetic
  Kuhn meshes, dofs owned by the lowest rank that touches them);
* `common::IndexMap` (`IndexMap`: owned range, ghost -> (owner, global index));
* `la::SparsityPattern::finalize()` -- entries of ghost rows travel to the owning rank, which
  merges them into its rows, adding new ghost COLUMNS where it does not know the dof
  (`MatrixExchange.build`);
* `la::MatrixCSR::scatter_rev()` / `la::Vector::scatter_rev(add)` as called after assembly in
  python/demo/demo_poisson.py:52,54 (`MatrixExchange.scatter_reverse`, `VectorExchange`).

Index arithmetic runs in torch on whatever device the arrays live on (CUDA for the product, CPU in
the gloo tests); the data path on the GPU is the library's pack / unpack-add kernels
(`cfx_gather_f64`, `cfx_scatter_add_f64`, `cfx_pattern_positions`) around one neighbour exchange
through `torch.distributed` (NCCL over NVLink).  Nothing here computes element tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .mesh import TETRAHEDRON, TRIANGLE, FunctionSpace, Mesh


# ----------------------------------------------------------------------------- index map
@dataclass
class IndexMap:
    """dolfinx.common.IndexMap stand-in for one rank."""

    rank: int
    world: int
    n_owned: int
    offset: int                    # global index of local index 0
    ghost_global: "object"         # (n_ghost,) int64, ascending
    ghost_owner: "object"          # (n_ghost,) int64
    l2g: "object"                  # (n_owned + n_ghost,) int64

    @property
    def n_total(self) -> int:
        return int(self.l2g.shape[0])

    def owner_ranges(self):
        """[(owner, first local ghost index, one past the last)] -- ghosts are grouped by owner here
        (static; computed once)."""
        if getattr(self, "_ranges", None) is None:
            own = self.ghost_owner.cpu().numpy()
            out, i = [], 0
            while i < own.size:
                j = i
                while j < own.size and own[j] == own[i]:
                    j += 1
                out.append((int(own[i]), self.n_owned + i, self.n_owned + j))
                i = j
            if len({q for q, _, _ in out}) != len(out):
                out = None  # not grouped: callers fall back to a per-entry split
            self._ranges = out if out is not None else False
        return self._ranges or None

    def global_to_local(self, g):
        """Local index of global indices `g` (int64 tensor); -1 where the dof is unknown here.
        No host round trip."""
        import torch

        own = (g >= self.offset) & (g < self.offset + self.n_owned)
        ng = int(self.ghost_global.numel())
        if ng == 0:
            return torch.where(own, g - self.offset, torch.full_like(g, -1))
        pos = torch.searchsorted(self.ghost_global, g).clamp(max=ng - 1)
        hit = self.ghost_global[pos] == g
        return torch.where(own, g - self.offset, torch.where(hit, pos + self.n_owned, torch.full_like(g, -1)))


def slab_ranges(n_layers: int, world: int, weights=None):
    """[lo, hi) cell layers of every rank.  Uniform by default; with per-layer `weights` (estimated work,
    e.g. cells + k * active cells) the cuts balance the cumulative weight -- the role vertex weights
    play in the graph partitioner behind DOLFINx's cell partition.  Every rank gets >= 1 layer."""
    if weights is None:
        return [((n_layers * p) // world, (n_layers * (p + 1)) // world) for p in range(world)]
    w = np.asarray(weights, dtype=np.float64)
    assert w.size == n_layers and world <= n_layers
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for p in range(1, world):
        k = int(np.searchsorted(cum, cum[-1] * p / world))
        k = min(max(k, cuts[-1] + 1), n_layers - (world - p))
        cuts.append(k)
    cuts.append(n_layers)
    return [(cuts[p], cuts[p + 1]) for p in range(world)]


def layer_weights(shape, p0, p1, level_set, active_cost: float = 40.0, samples: int = 48):
    """Work estimate per cell layer along the last axis: 1 per cell + `active_cost` per cell with
    phi < 0 at its centre, from a coarse sample of the level-set function (host, numpy)."""
    tdim = len(shape)
    ax = tdim - 1
    n_ax = shape[ax]
    z = p0[ax] + (np.arange(n_ax) + 0.5) * (p1[ax] - p0[ax]) / n_ax
    g = [p0[d] + (np.arange(samples) + 0.5) * (p1[d] - p0[d]) / samples for d in range(ax)]
    out = np.empty(n_ax)
    for k in range(n_ax):
        if tdim == 3:
            X, Y = np.meshgrid(g[0], g[1], indexing="ij")
            frac = np.mean(level_set(X, Y, np.full_like(X, z[k])) < 0)
        else:
            frac = np.mean(level_set(g[0], np.full_like(g[0], z[k]), np.zeros_like(g[0])) < 0)
        out[k] = 1.0 + active_cost * frac
    return out


def partition_slab(shape, p0, p1, world: int, rank: int, device=None, ranges=None):
    """Local mesh + P1 space + index map of rank `rank` for a slab partition of the Kuhn box /
    right-diagonal rectangle with `shape` cells, split along the last axis.

    device=None: host arrays (numpy, via mesh.create_*); device=int: CUDA tensors generated by the
    library's mesh generator.  Returns (mesh, V, index_map).  Global dof numbering = the global
    lexicographic vertex numbering (dofs are owned plane by plane, lowest rank first), so the
    union of the ranks' owned rows is directly comparable with a serial assembly."""
    import torch

    from . import mesh as M

    shape = tuple(int(s) for s in shape)
    tdim = len(shape)
    ax = tdim - 1
    n_ax = shape[ax]
    ranges = slab_ranges(n_ax, world) if ranges is None else list(ranges)
    lo, hi = ranges[rank]
    if hi <= lo:
        raise ValueError(f"rank {rank} would own no cell layer: {n_ax} layers on {world} ranks")
    g0, g1 = int(rank > 0), int(rank < world - 1)
    L0, L1 = lo - g0, hi + g1
    nl = L1 - L0
    h = (p1[ax] - p0[ax]) / n_ax
    q0, q1 = list(p0), list(p1)
    q0[ax], q1[ax] = p0[ax] + L0 * h, p0[ax] + L1 * h
    lshape = list(shape)
    lshape[ax] = nl
    if device is None:
        lex = M.create_box(*lshape, q0, q1) if tdim == 3 else M.create_rectangle(*lshape, q0, q1)
        dev = torch.device("cpu")
        x = torch.from_numpy(lex.x)
        xd = torch.from_numpy(lex.x_dofmap)
        c2f = torch.from_numpy(lex.c2f)
    else:
        from . import demo_poisson as dp

        lex = dp.device_mesh(int(device), lshape, q0, q1)
        dev = lex.x.device
        x, xd, c2f = lex.x, lex.x_dofmap, lex.c2f
        ctx = lex.extra.pop("_cfx_ctx", None)
        if ctx is not None:
            ctx.close()
    cell_type = TETRAHEDRON if tdim == 3 else TRIANGLE
    nv = tdim + 1
    cpl = int(np.prod(shape[:ax])) * (6 if tdim == 3 else 2)        # cells per layer
    vpp = int(np.prod([s + 1 for s in shape[:ax]]))                 # vertices per plane
    nslot = 12 if tdim == 3 else 3
    n_planes = nl + 1
    # ---- planes: owned first (ascending), then ghosts owned below, then ghosts owned above
    zf = lo + 1 if rank > 0 else 0                                  # first owned global plane
    owned_planes = list(range(zf - L0, hi - L0 + 1))
    below = list(range(0, zf - L0))
    above = list(range(hi - L0 + 1, n_planes))
    plane_order = owned_planes + below + above
    new_plane_pos = torch.empty(n_planes, dtype=torch.int64)
    new_plane_pos[torch.tensor(plane_order)] = torch.arange(n_planes)
    new_plane_pos = new_plane_pos.to(dev)
    # ---- layers: owned block first, then the ghost layer below, then the one above
    layer_order = list(range(g0, g0 + hi - lo)) + ([0] if g0 else []) + ([nl - 1] if g1 else [])
    lt = torch.tensor(layer_order, device=dev)
    if world == 1:
        x_new, xd_new, c2f_new = x, xd, c2f
        n_owned_f = nslot * vpp * n_planes
    else:
        x_new = x.view(n_planes, vpp, 3)[torch.tensor(plane_order, device=dev)].reshape(-1, 3).contiguous()
        v = xd.to(torch.int64)
        v = new_plane_pos[v // vpp] * vpp + v % vpp
        xd_new = v.view(nl, cpl, nv)[lt].reshape(-1, nv).to(torch.int32).contiguous()
        # ---- facets: owner layer by (base plane, slot); [owned | ghost] numbering
        kl = torch.arange(n_planes).view(-1, 1).expand(n_planes, nslot)
        slot = torch.arange(nslot).view(1, -1).expand(n_planes, nslot)
        low = (slot >= 10) if tdim == 3 else (slot == 2)            # facets on the low side along `ax`
        owner_layer = torch.where(low, (kl - 1).clamp(min=0), kl)
        owned_tab = (owner_layer >= g0) & (owner_layer < g0 + hi - lo)
        owned_tab &= ~((kl == nl) & ~low)                           # virtual cells above the top plane: only low-side facets
        m_own = owned_tab.sum(dim=1)                                # owned slots per vertex of the plane
        m_oth = nslot - m_own
        base_own = vpp * (torch.cumsum(m_own, 0) - m_own)
        n_owned_f = int(vpp * m_own.sum())
        base_oth = n_owned_f + vpp * (torch.cumsum(m_oth, 0) - m_oth)
        rank_own = torch.cumsum(owned_tab.to(torch.int64), 1) - owned_tab.to(torch.int64)
        rank_oth = torch.cumsum((~owned_tab).to(torch.int64), 1) - (~owned_tab).to(torch.int64)
        T_base = torch.where(owned_tab, base_own.view(-1, 1), base_oth.view(-1, 1)).to(dev)
        T_m = torch.where(owned_tab, m_own.view(-1, 1), m_oth.view(-1, 1)).to(dev)
        T_rank = torch.where(owned_tab, rank_own, rank_oth).to(dev)
        f = c2f.to(torch.int64)
        fv, fs = f // nslot, f % nslot
        fk, fu = fv // vpp, fv % vpp
        fnew = T_base[fk, fs] + fu * T_m[fk, fs] + T_rank[fk, fs]
        c2f_new = fnew.view(nl, cpl, nv)[lt].reshape(-1, nv).to(torch.int32).contiguous()
    n_facets = nslot * vpp * n_planes
    n_owned_cells = (hi - lo) * cpl
    # ---- index map: global dof = global lexicographic vertex id
    z_of_new = torch.tensor([L0 + k for k in plane_order], dtype=torch.int64, device=dev)
    if world > 1:
        # coordinates along the split axis from the GLOBAL formula (start + z * step, last point exact), so
        # that every rank sees bit-identical vertices and level-set values -- a sub-box linspace differs in
        # the last bit, which flips the classification of vertices lying exactly on the interface
        step = (p1[ax] - p0[ax]) / n_ax
        zc = z_of_new.to(torch.float64) * step + p0[ax]
        zc[z_of_new == n_ax] = p1[ax]
        x_new.view(n_planes, vpp, 3)[:, :, ax] = zc.view(-1, 1)
    l2g = (z_of_new.view(-1, 1) * vpp + torch.arange(vpp, device=dev).view(1, -1)).reshape(-1)
    n_owned = len(owned_planes) * vpp
    owner_of_plane = lambda z: 0 if z == 0 else next(q for q, (a, b) in enumerate(ranges) if a < z <= b)  # noqa: E731
    gown = torch.tensor([owner_of_plane(L0 + k) for k in below + above], dtype=torch.int64, device=dev)
    ghost_owner = gown.view(-1, 1).expand(len(below) + len(above), vpp).reshape(-1)
    imap = IndexMap(rank, world, n_owned, zf * vpp, l2g[n_owned:].contiguous(), ghost_owner.contiguous(), l2g)
    if device is None:
        xa, xda, c2fa = x_new.numpy(), xd_new.numpy(), c2f_new.numpy()
        off, f2c = M.invert_c2f(c2fa, n_facets)
    else:
        xa, xda, c2fa, off, f2c = x_new, xd_new, c2f_new, None, None
    mesh = Mesh(cell_type, tdim, tdim, xa, xda, c2fa, off, f2c, n_owned_cells, n_facets, n_owned_f,
                shape=tuple(lshape), p0=tuple(q0), p1=tuple(q1))
    if device is not None:
        mesh.extra["device"] = int(device)
    n_dofs = int(xda.shape[0] and x_new.shape[0])
    V = FunctionSpace(mesh, 1, xda, n_dofs, n_owned, 1, xa if device is None else None)
    return mesh, V, imap


def p2_tet_global_dofs(gv, shape):
    """P2 dofs of Kuhn tetrahedra from the GLOBAL lexicographic vertex ids `gv` (n_cells, 4) int64 of their
    vertices, in the partition's global numbering: every mesh plane z is a block of B = 8 * vpp ids,

        z * B + d * vpp + u,   u = vertex index inside the plane,
        d = 0: the vertex;  d = 1..7: the edge leaving the vertex in direction (d & 1, d >> 1 & 1, d >> 2)

    (the offsets of a Kuhn tetrahedron's vertices are nested bit patterns, so an edge is its lower vertex and one of 7
    directions; ids of edges that would leave the box stay unused, as in the serial device_p2_tet_space).
    d < 4 lies IN plane z, d >= 4 inside cell layer z -- blocks owned by the plane's / the layer's rank, which
    makes every rank's owned ids one contiguous range.  Basix edge order e0 = (2,3) ... e5 = (0,1)."""
    import torch

    from .mesh import TET_EDGES

    sx = int(shape[0]) + 1
    vpp = sx * (int(shape[1]) + 1)
    B = 8 * vpp
    ea = torch.tensor([e[0] for e in TET_EDGES], device=gv.device)
    eb = torch.tensor([e[1] for e in TET_EDGES], device=gv.device)
    va, vb = gv[:, ea], gv[:, eb]
    a = torch.minimum(va, vb)
    diff = torch.maximum(va, vb) - a
    dk = diff // vpp
    rem = diff - dk * vpp
    dj = rem // sx
    di = rem - dj * sx
    d = di + 2 * dj + 4 * dk
    vert = (gv // vpp) * B + gv % vpp
    edge = (a // vpp) * B + d * vpp + a % vpp
    return torch.cat([vert, edge], dim=1)


def p2_tet_slab_space(mesh: Mesh, imap1: IndexMap, shape, ranges, rank: int, world: int, bs: int = 1):
    """P2 (block size `bs`) space and index map of one rank of a slab partition of the Kuhn box: `mesh`, `imap1` are
    what partition_slab returned (imap1.l2g = global lexicographic vertex ids).  Owned ids of rank r:
    [lo * B + 4 vpp, hi * B + 4 vpp) (rank 0 from 0): the edges inside its cell layers and the planes above them;
    ghosts: the dofs of the ghost cell layer below (planes lo - 1, lo and the layer between) and above."""
    import torch

    shape = tuple(int(v) for v in shape)
    assert len(shape) == 3
    dev = imap1.l2g.device
    vpp = (shape[0] + 1) * (shape[1] + 1)
    B = 8 * vpp
    lo, hi = ranges[rank]
    gv = imap1.l2g[mesh.x_dofmap.to(torch.int64) if hasattr(mesh.x_dofmap, "to") else
                   torch.from_numpy(np.asarray(mesh.x_dofmap, dtype=np.int64))]
    g = p2_tet_global_dofs(gv, shape)
    offset = 0 if rank == 0 else lo * B + 4 * vpp
    end = hi * B + 4 * vpp
    n_owned = end - offset
    owner_of_plane = lambda z: 0 if z == 0 else next(q for q, (a, b) in enumerate(ranges) if a < z <= b)  # noqa: E731
    parts, owners = [], []
    if rank > 0:       # plane lo - 1 | layer lo - 1 | plane lo
        parts.append(torch.arange((lo - 1) * B, lo * B + 4 * vpp, device=dev))
        owners.append(torch.cat([torch.full((4 * vpp,), owner_of_plane(lo - 1), dtype=torch.int64, device=dev),
                                 torch.full((8 * vpp,), rank - 1, dtype=torch.int64, device=dev)]))
    n_below = int(parts[0].numel()) if parts else 0
    if rank < world - 1:  # layer hi | plane hi + 1
        parts.append(torch.arange(end, end + B, device=dev))
        owners.append(torch.full((B,), rank + 1, dtype=torch.int64, device=dev))
    z64 = torch.zeros(0, dtype=torch.int64, device=dev)
    ghost_global = torch.cat(parts) if parts else z64
    ghost_owner = torch.cat(owners) if owners else z64
    below0 = (lo - 1) * B
    local = torch.where((g >= offset) & (g < end), g - offset,
                        torch.where(g < offset, n_owned + g - below0, n_owned + n_below + g - end))
    n_total = n_owned + int(ghost_global.numel())
    if bool(((local < 0) | (local >= n_total)).any()):
        raise RuntimeError("p2_tet_slab_space: a cell touches a dof outside the rank's owned and ghost blocks")
    l2g = torch.cat([torch.arange(offset, end, device=dev), ghost_global])
    imap = IndexMap(rank, world, n_owned, offset, ghost_global.contiguous(), ghost_owner.contiguous(), l2g)
    dm = local.to(torch.int32).contiguous()
    if not hasattr(mesh.x_dofmap, "to"):
        dm = dm.numpy()
    return FunctionSpace(mesh, 2, dm, n_total, n_owned, bs, None), imap


def p2_tet_global_to_serial(g, shape):
    """Partition-global P2 id (p2_tet_global_dofs) -> the id of the same dof in the serial device_p2_tet_space /
    meshgen numbering (vertex v -> v, edge (a, d) -> n_nodes + 7 a + d - 1).  numpy or torch int64."""
    vpp = (int(shape[0]) + 1) * (int(shape[1]) + 1)
    nn = vpp * (int(shape[2]) + 1)
    B = 8 * vpp
    z, r = g // B, g % B
    d, u = r // vpp, r % vpp
    v = z * vpp + u
    return (d == 0) * v + (d != 0) * (nn + 7 * v + d - 1)


# ----------------------------------------------------------------------------- transports
class TorchDistTransport:
    """One rank per process (`torch.distributed`, backend nccl on GPUs / gloo on the CPU)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.local_ranks = [self.rank]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
                else torch.device("cpu")
        self.device = torch.device(device)

    def exchange(self, sends, counts=None, dtype=None):
        """sends: [ {dest rank: 1-D tensor} ] for the one local rank.  `counts`: optional
        [ {src rank: n} ] when the receive sizes are already known (no size exchange).  `dtype`: element
        type of this exchange -- every rank must name the same one, also a rank that sends nothing."""
        import torch

        dist = self.dist
        send = sends[0]
        ref = next(iter(send.values())) if send else None
        device = self.device
        dtype = dtype if dtype is not None else (ref.dtype if ref is not None else torch.float64)
        in_split = [int(send[q].numel()) if q in send else 0 for q in range(self.world)]
        if counts is None:
            s = torch.tensor(in_split, dtype=torch.int64, device=device)
            r = torch.empty_like(s)
            dist.all_to_all_single(r, s, group=self.group)
            out_split = [int(v) for v in r.tolist()]
        else:
            out_split = [int(counts[0].get(q, 0)) for q in range(self.world)]
        inp = torch.cat([send[q].reshape(-1) for q in range(self.world) if q in send]) if send else \
            torch.empty(0, dtype=dtype, device=device)
        out = torch.empty(sum(out_split), dtype=inp.dtype, device=device)
        dist.all_to_all_single(out, inp, out_split, in_split, group=self.group)
        res, o = {}, 0
        for q, n in enumerate(out_split):
            if n:
                res[q] = out[o:o + n]
            o += n
        return [res]


class LocalTransport:
    """All ranks in one process (tests, and N emulated ranks on one GPU): a mailbox."""

    def __init__(self, world: int):
        self.world = world
        self.local_ranks = list(range(world))

    def exchange(self, sends, counts=None, dtype=None):
        res = [dict() for _ in range(self.world)]
        for p, send in enumerate(sends):
            for q, t in send.items():
                if t.numel():
                    res[q][p] = t.clone()
        return res


# ----------------------------------------------------------------------------- device helpers
class _DeviceOps:
    """The library kernels around the exchange (CUDA tensors only: there is no CPU fallback)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def _chk(self, rc):
        from ._lib import check

        check(self.ctx.handle, rc)

    def gather(self, src, index):
        import torch

        from ._lib import lib

        out = torch.empty(index.numel(), dtype=torch.float64, device=src.device)
        self._chk(lib().cfx_gather_f64(self.ctx.handle, C.c_void_p(src.data_ptr()), C.c_void_p(index.data_ptr()),
                                       C.c_int64(index.numel()), C.c_void_p(out.data_ptr())))
        return out

    def scatter_add(self, dst, index, src):
        from ._lib import lib

        self._chk(lib().cfx_scatter_add_f64(self.ctx.handle, C.c_void_p(src.data_ptr()),
                                            C.c_void_p(index.data_ptr()), C.c_int64(index.numel()),
                                            C.c_void_p(dst.data_ptr())))

    def positions(self, A, rows, cols):
        import torch

        from ._lib import lib

        pos = torch.empty(rows.numel(), dtype=torch.int64, device=rows.device)
        r32, c32 = rows.to(torch.int32).contiguous(), cols.to(torch.int32).contiguous()
        self._chk(lib().cfx_pattern_positions(self.ctx.handle, A._h, C.c_void_p(r32.data_ptr()),
                                              C.c_void_p(c32.data_ptr()), C.c_int64(rows.numel()),
                                              C.c_void_p(pos.data_ptr())))
        return pos


# ----------------------------------------------------------------------------- plans (host logic)
def split_by_owner(imap: IndexMap, rows_local):
    """{owner rank: index tensor into rows_local} for ghost rows (ascending local row => grouped)."""
    owners = imap.ghost_owner[rows_local - imap.n_owned]
    out = {}
    for q in [int(v) for v in owners.unique().tolist()]:
        out[q] = (owners == q).nonzero().reshape(-1)
    return out


def ghost_entries_to_send(imap: IndexMap, rows_local, cols_local):
    """Phase A of SparsityPattern::finalize: ghost-row entries as global (row, col) pairs per owner.
    Returns ({owner: int64 tensor (2*n,)}, {owner: index tensor}) -- the index tensors remember which
    local entries went where (same order as sent)."""
    import torch

    sel = split_by_owner(imap, rows_local) if rows_local.numel() else {}
    sends = {}
    for q, idx in sel.items():
        rg, cg = imap.l2g[rows_local[idx]], imap.l2g[cols_local[idx]]
        sends[q] = torch.stack([rg, cg]).reshape(-1).contiguous()
    return sends, sel


def received_entries_to_local(imap: IndexMap, recv):
    """Phase B of finalize on the owner: received global pairs -> local (row, col); dofs this rank
    does not know become NEW ghost columns n_total, n_total+1, ... (ascending global index).
    Returns ({src: (rows, cols)}, new_ghost_globals).  Two host round trips in total."""
    import torch

    if not recv:
        return {}, torch.empty(0, dtype=torch.int64, device=imap.l2g.device)
    srcs = sorted(recv)
    sizes = [int(recv[q].numel()) // 2 for q in srcs]
    rg = torch.cat([recv[q].view(2, -1)[0] for q in srcs])
    cg = torch.cat([recv[q].view(2, -1)[1] for q in srcs])
    rows = rg - imap.offset
    cols = imap.global_to_local(cg)
    bad = ((rows < 0) | (rows >= imap.n_owned)).any()
    new_globals = torch.unique(cg[cols < 0])                       # round trip 1 (size of the selection)
    if bool(bad):                                                  # round trip 2
        raise RuntimeError("received a matrix row this rank does not own")
    if new_globals.numel():
        cols = torch.where(cols < 0, imap.n_total + torch.searchsorted(new_globals, cg), cols)
    out, o = {}, 0
    for q, n in zip(srcs, sizes):
        out[q] = (rows[o:o + n], cols[o:o + n])
        o += n
    return out, new_globals


def merge_inserted_entries(per_src, disjoint_ascending=False):
    """Union of the received (row, col) pairs, rows ascending -- the argument of
    cfx_form_insert_pattern_entries.  `disjoint_ascending`: the sources (in ascending rank order) are
    known to target disjoint, ascending row ranges (slab neighbours), so their concatenation is already
    a valid insertion list (each source's pairs are distinct and row-sorted): no sort needed."""
    import torch

    if not per_src:
        return None, None
    rows = torch.cat([per_src[q][0] for q in sorted(per_src)])
    cols = torch.cat([per_src[q][1] for q in sorted(per_src)])
    if disjoint_ascending:
        return rows.to(torch.int32).contiguous(), cols.to(torch.int32).contiguous()
    key = torch.unique(rows * (1 << 32) + cols)
    return (key >> 32).to(torch.int32).contiguous(), (key & 0xFFFFFFFF).to(torch.int32).contiguous()


class VectorExchange:
    """Static plan for la::Vector::scatter_rev(add) / scatter_fwd from an IndexMap."""

    def __init__(self, imap: IndexMap):
        self.imap = imap
        self.send_sel = None   # {owner: local ghost indices}
        self.recv_pos = None   # {src: owned local positions}

    def begin(self):
        """-> sends for Transport.exchange (ghost global indices to their owners)."""
        import torch

        im = self.imap
        gl = torch.arange(im.n_owned, im.n_total, device=im.l2g.device)
        self.send_sel = {}
        for q in [int(v) for v in im.ghost_owner.unique().tolist()]:
            self.send_sel[q] = gl[im.ghost_owner == q]
        return {q: im.l2g[idx].contiguous() for q, idx in self.send_sel.items()}

    def finish(self, recv):
        self.recv_pos = {q: (t - self.imap.offset).contiguous() for q, t in recv.items()}

    def pack_reverse(self, b):
        return {q: b[idx].contiguous() for q, idx in self.send_sel.items()}

    def recv_counts(self):
        return {q: int(p.numel()) for q, p in self.recv_pos.items()}

    def unpack_reverse_add(self, b, recv, scatter_add):
        """`scatter_add(dst, index, src)`: the product passes the library kernel
        (_DeviceOps.scatter_add); indices are distinct per neighbour."""
        for q in sorted(recv):  # fixed neighbour order: bit-reproducible
            scatter_add(b, self.recv_pos[q], recv[q])
        return b


class MatrixExchange:
    """Per-pattern plan for SparsityPattern::finalize + MatrixCSR::scatter_rev (host logic only)."""

    def __init__(self, imap: IndexMap):
        self.imap = imap
        self.sel = {}          # {owner: indices into the ghost-row COO}
        self.ghost_rows = None
        self.ghost_cols = None
        self.recv_local = {}   # {src: (rows, cols)} in local numbering
        self.new_ghost_globals = None
        self.send_pos = {}     # {owner: positions in the final values array}
        self.recv_pos = {}     # {src: positions in the final values array}

    def begin(self, ghost_rows, ghost_cols, entry_ranges=None):
        """ghost-row entries (local COO, ascending rows) -> sends of global (row, col) pairs.
        `entry_ranges` {owner: (first, last) entry} when the caller already knows the split."""
        import torch

        self.ghost_rows, self.ghost_cols = ghost_rows, ghost_cols
        if entry_ranges is None:
            sends, self.sel = ghost_entries_to_send(self.imap, ghost_rows, ghost_cols)
            return sends
        sends, self.sel = {}, {}
        for q, (a, b) in entry_ranges.items():
            if b > a:
                self.sel[q] = slice(a, b)
                sends[q] = torch.stack([self.imap.l2g[ghost_rows[a:b]], self.imap.l2g[ghost_cols[a:b]]]).reshape(-1)
        return sends

    def inserted_entries(self, recv, disjoint_ascending=False):
        """received pairs -> (rows, cols) to insert into the owner's pattern (int32, rows ascending)."""
        self.recv_local, self.new_ghost_globals = received_entries_to_local(self.imap, recv)
        return merge_inserted_entries(self.recv_local, disjoint_ascending)

    def finish(self, positions):
        """`positions(rows, cols)` -> positions in the final pattern (cfx_pattern_positions); all queries
        of the step go through ONE call."""
        import torch

        keys = [("s", q) for q in sorted(self.sel)] + [("r", q) for q in sorted(self.recv_local)]
        if not keys:
            self.send_pos, self.recv_pos = {}, {}
            return
        rows = [self.ghost_rows[self.sel[q]] if k == "s" else self.recv_local[q][0] for k, q in keys]
        cols = [self.ghost_cols[self.sel[q]] if k == "s" else self.recv_local[q][1] for k, q in keys]
        pos = positions(torch.cat(rows), torch.cat(cols))
        self.send_pos, self.recv_pos, o = {}, {}, 0
        for (k, q), r in zip(keys, rows):
            n = int(r.numel())
            (self.send_pos if k == "s" else self.recv_pos)[q] = pos[o:o + n]
            o += n

    def recv_counts(self):
        return {q: int(p.numel()) for q, p in self.recv_pos.items()}


# ----------------------------------------------------------------------------- static exchange plan
def candidate_entries(indptr, indices, rows):
    """Entries of the CSR rows `rows` (int64 tensor), row by row: (row of each entry, column of each entry,
    offsets (len(rows) + 1) of each row's entries).  torch, any device; setup only."""
    import torch

    counts = indptr[rows + 1] - indptr[rows]
    ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    ptr[1:] = torch.cumsum(counts, 0)
    total = int(ptr[-1])
    flat = torch.repeat_interleave(indptr[rows] - ptr[:-1], counts, output_size=total) + \
        torch.arange(total, device=rows.device)
    rowrep = torch.repeat_interleave(rows, counts, output_size=total)
    return rowrep, indices[flat].to(torch.int64), ptr


def static_plan_tables(imap: IndexMap, send_rows, cand, recv_pairs, recv_vec_rows):
    """Host logic of the static exchange plan of one rank (include/cutfemx_b200.h, cfx_xplan_create).

    send_rows     {owner q: local ghost rows (ascending)}                       -- VectorExchange.send_sel
    cand          {owner q: (ptr (len(rows)+1), local candidate columns)}       -- this rank's static ghost-row pattern
    recv_pairs    {source q: int64 tensor [global rows ..., global cols ...]}   -- what the sources announced
    recv_vec_rows {source q: local owned rows of q's ghost dofs, q's order}     -- VectorExchange.recv_pos

    Returns the argument arrays of cfx_xplan_create (numpy) and the new ghost columns' global indices."""
    import torch

    neigh = sorted(set(send_rows) | set(recv_pairs) | set(recv_vec_rows))
    dev = imap.l2g.device
    z64 = torch.zeros(0, dtype=torch.int64, device=dev)
    s_row_off, s_rows, s_ptr, s_cols = [0], [], [torch.zeros(1, dtype=torch.int64, device=dev)], []
    for q in neigh:
        rows = send_rows.get(q, z64)
        ptr, cols = cand.get(q, (torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev), z64))
        s_rows.append(rows)
        s_ptr.append(ptr[1:] + s_ptr[-1][-1])
        s_cols.append(cols)
        s_row_off.append(s_row_off[-1] + int(rows.numel()))
    r_ent_off, r_row, r_colg, r_row_off, r_vec = [0], [], [], [0], []
    for q in neigh:
        t = recv_pairs.get(q, z64).view(2, -1)
        r_row.append(t[0] - imap.offset)
        r_colg.append(t[1])
        r_ent_off.append(r_ent_off[-1] + int(t.shape[1]))
        v = recv_vec_rows.get(q, z64)
        r_vec.append(v)
        r_row_off.append(r_row_off[-1] + int(v.numel()))
    rows = torch.cat(r_row) if r_row else z64
    cg = torch.cat(r_colg) if r_colg else z64
    if rows.numel() and bool(((rows < 0) | (rows >= imap.n_owned)).any()):
        raise RuntimeError("received a matrix row this rank does not own")
    cols = imap.global_to_local(cg) if cg.numel() else z64
    new_globals = torch.unique(cg[cols < 0]) if cg.numel() else z64
    if new_globals.numel():  # dofs this rank does not know: new ghost columns (ascending global index)
        cols = torch.where(cols < 0, imap.n_total + torch.searchsorted(new_globals, cg), cols)
    perm = torch.sort(rows, stable=True).indices if rows.numel() else z64
    i32 = lambda t: np.ascontiguousarray(t.cpu().numpy().astype(np.int32))    # noqa: E731
    i64 = lambda t: np.ascontiguousarray(t.cpu().numpy().astype(np.int64))    # noqa: E731
    cat = lambda parts: torch.cat(parts) if parts else z64                    # noqa: E731
    tables = dict(neigh=np.asarray(neigh, dtype=np.int32), s_row_off=np.asarray(s_row_off, dtype=np.int64),
                  s_rows=i32(cat(s_rows)), s_ptr=i64(cat(s_ptr)), s_cols=i32(cat(s_cols)),
                  r_ent_off=np.asarray(r_ent_off, dtype=np.int64), r_row=i32(rows), r_col=i32(cols), r_perm=i32(perm),
                  r_row_off=np.asarray(r_row_off, dtype=np.int64), r_vec_row=i32(cat(r_vec)))
    return tables, new_globals


class StaticExchange:
    """cfx_xplan of one rank: SparsityPattern::finalize + scatter_rev with fixed-size messages, all on the device."""

    def __init__(self, ctx, space_index: int, tables, new_ghost_globals):
        from ._lib import check, lib

        self.ctx, self.tables, self.new_ghost_globals = ctx, tables, new_ghost_globals
        self.neigh = [int(q) for q in tables["neigh"]]
        self._h = C.c_void_p()
        p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        t = tables
        check(ctx.handle, lib().cfx_xplan_create(ctx.handle, space_index, len(self.neigh), p(t["neigh"]),
                                                 p(t["s_row_off"]), p(t["s_rows"]), p(t["s_ptr"]), p(t["s_cols"]),
                                                 p(t["r_ent_off"]), p(t["r_row"]), p(t["r_col"]), p(t["r_perm"]),
                                                 p(t["r_row_off"]), p(t["r_vec_row"]), C.byref(self._h)))

    def _call(self, fn, *args):
        from ._lib import check, lib

        check(self.ctx.handle, getattr(lib(), fn)(self.ctx.handle, self._h, *args))

    def pack_pattern(self, a):
        self._call("cfx_xplan_pack_pattern", a._h)

    def insert_pattern(self, a):
        self._call("cfx_xplan_insert_pattern", a._h)

    def pack_values(self, A, b):
        self._call("cfx_xplan_pack_values", A._h, C.c_void_p(b.data_ptr()) if b is not None else None)

    def unpack_add(self, A, b):
        self._call("cfx_xplan_unpack_add", A._h, C.c_void_p(b.data_ptr()) if b is not None else None)

    def exchange(self, which: int):
        """NCCL send/recv with every neighbour on the context's stream (cfx_comm_init first)."""
        self._call("cfx_xplan_exchange", int(which))

    def buffer(self, which: int, k: int):
        """torch uint8 view of a message buffer (emulated transports)."""
        from ._lib import check, device_view, lib

        ptr, nb = C.c_void_p(), C.c_int64()
        check(self.ctx.handle, lib().cfx_xplan_buffer(self.ctx.handle, self._h, which, k, C.byref(ptr), C.byref(nb)))
        return device_view(ptr.value or 0, nb.value // 8, np.int64, self.ctx.device, self) if nb.value % 8 == 0 \
            else device_view(ptr.value or 0, nb.value // 4, np.int32, self.ctx.device, self)

    def free(self):
        from ._lib import lib

        if self._h:
            lib().cfx_xplan_free(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


def init_nccl(ctx, rank: int, world: int, group=None):
    """Give the context its own NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import torch
    import torch.distributed as dist

    from ._lib import check, lib

    uid = (C.c_char * 128)()
    if rank == 0:
        check(None, lib().cfx_comm_unique_id(uid, None))
    backend = dist.get_backend(group)
    dev = torch.device("cuda", ctx.device) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, 0, group=group)
    raw = bytes(t.cpu().numpy().tobytes())
    check(ctx.handle, lib().cfx_comm_init(ctx.handle, raw, rank, world, None))


def coo_of_rows(indptr, indices, row_begin: int, n_rows: int, cuts=()):
    """(rows, cols) int64 of CSR rows [row_begin, n_rows) -- torch, any device.  `cuts`: extra row
    indices whose entry offsets (relative to row_begin) are returned too, fetched in the same
    single device->host copy as the two ends."""
    import torch

    pts = torch.tensor([row_begin, n_rows, *cuts], device=indptr.device)
    vals = [int(v) for v in indptr[pts].tolist()]
    start, end = vals[0], vals[1]
    rp = indptr[row_begin:n_rows + 1]
    counts = rp[1:] - rp[:-1]
    rows = torch.repeat_interleave(torch.arange(row_begin, n_rows, device=indptr.device), counts,
                                   output_size=end - start)
    cols = indices[start:end].to(torch.int64)
    if cuts:
        return rows, cols, [v - start for v in vals[2:]]
    return rows, cols


# ----------------------------------------------------------------------------- the pipeline, one rank
class RankPipeline:
    """demo_poisson.py:156-201 on one rank of a slab partition (device path).

    phase_a: local cut path up to the forms + the ghost-row pattern   -> entries for the owners
    phase_b: insert received entries, sparsity, assemble A and b, pack -> ghost values for the owners
    phase_c: unpack-add (A.scatter_reverse(), b.scatter_reverse(add))
    """

    def __init__(self, shape, p0, p1, world, rank, device, ls_kind, ls_params, order=4, ranges=None, degree=1, **kw):
        from . import demo_poisson as dp
        from .mesh import Function

        self.world, self.rank = world, rank
        self.ranges = slab_ranges(int(shape[-1]), world) if ranges is None else list(ranges)
        self.mesh, self.V, self.imap = partition_slab(shape, p0, p1, world, rank, device=device, ranges=self.ranges)
        if callable(ls_kind):
            # nodal interpolation on the host (numpy), as the tests do for the serial problem
            import torch

            xh = self.mesh.x.cpu().numpy()
            vals = torch.from_numpy(np.ascontiguousarray(ls_kind(xh[:, 0], xh[:, 1], xh[:, 2]))).to(self.mesh.x.device)
        else:
            vals = dp.device_level_set(self.mesh, ls_kind, ls_params)  # phi on owned + ghost dofs ("scattered")
        self.phi = Function(self.V, "phi", vals)  # the level set is P1 on the mesh vertices
        self.ls_kind, self.ls_params = ls_kind, ls_params
        problem = kw.pop("problem", "poisson")
        bs = int(kw.pop("bs", 1))
        # "blocks": the partition's plane-block numbering on one rank too (bench.py: the same rows for every N)
        p2_numbering = kw.pop("p2_numbering", "serial")
        if degree == 2 and self.mesh.tdim == 3:
            # P2 on tetrahedra (BASELINE configs[3]): vertex dofs + edge dofs generated on the device
            import torch

            if world == 1 and p2_numbering != "blocks":
                self.V = dp.device_p2_tet_space(self.mesh, bs)
                nd = self.V.num_dofs
                self.imap = IndexMap(rank, world, nd, 0, self.imap.ghost_global, self.imap.ghost_owner,
                                     torch.arange(nd, device=self.mesh.x.device))
            else:  # dofs numbered plane block by plane block, owned ranges contiguous (p2_tet_slab_space)
                self.V, self.imap = p2_tet_slab_space(self.mesh, self.imap, shape, self.ranges, rank, world, bs)
        elif degree == 2:
            # P2 on triangles: vertex dofs + one dof per edge; on these meshes edge == facet and local edge e is
            # opposite local vertex e (Basix), exactly the c2f convention -> edge dof = n_vertices + facet id
            if world != 1 or self.mesh.tdim != 2:
                raise NotImplementedError("P2 spaces: single-rank triangle meshes only")
            import torch

            nn = int(self.mesh.x.shape[0])
            dm = torch.cat([self.mesh.x_dofmap, self.mesh.c2f + nn], dim=1).contiguous()
            nd = nn + int(self.mesh.num_facets)
            self.V = FunctionSpace(self.mesh, 2, dm, nd, nd, bs, None)
            self.imap = IndexMap(rank, world, nd, 0, self.imap.ghost_global, self.imap.ghost_owner,
                                 torch.arange(nd, device=dm.device))
        elif degree != 1:
            raise ValueError("degree must be 1 or 2")
        elif bs != 1:  # blocked P1: the index map counts blocks, as DOLFINx's does
            self.V = FunctionSpace(self.mesh, 1, self.V.dofmap, self.V.num_dofs, self.V.num_dofs_owned, bs, None)
        if problem == "elasticity":
            from .demo_elasticity import CutElasticity

            self.prob = CutElasticity(self.mesh, self.phi, self.V, order=order, **kw)
        else:
            self.prob = dp.CutPoisson(self.mesh, self.phi, self.V, order=order, **kw)
        self.ctx = self.prob.ctx
        self.ops = _DeviceOps(self.ctx)
        self.vx = VectorExchange(self.imap)
        self.mx = MatrixExchange(self.imap)
        self.Ag = None
        self.xplan = None
        # slab neighbours send rows of different mesh planes: below-neighbour rows < above-neighbour rows
        lo, hi = self.ranges[rank]
        self.disjoint_sources = hi - lo >= 3

    # static vector plan (once)
    def plan_begin(self):
        return self.vx.begin()

    def plan_finish(self, recv):
        self.vx.finish(recv)

    # ---- static exchange plan (once per partition): cfx_xplan
    def static_begin(self):
        """The static superset of this rank's ghost-row entries (every owned cell active, every owned interior
        facet in the band) -> global (row, col) pairs for the owners."""
        import torch

        from . import fem as _fem
        from ._lib import DEVICE, check, lib
        from .cut import _List, _bind_topology, facet_integration_rows_device

        self.xplan = None
        if self.world == 1:
            return {}
        mesh, V, ctx, im = self.mesh, self.V, self.ctx, self.imap
        dev = torch.device("cuda", ctx.device)
        _bind_topology(mesh, ctx)
        nct = int(mesh.x_dofmap.shape[0])
        cells = torch.arange(nct, dtype=torch.int32, device=dev)
        facets = _List(ctx)
        check(ctx.handle, lib().cfx_interior_facets_for_cells(ctx.handle, C.c_void_p(cells.data_ptr()), C.c_int64(nct),
                                                              DEVICE, 0, C.byref(facets._h)))
        rows4 = facet_integration_rows_device(mesh, facets)
        a = _fem.CutForm(V, 2)
        if V.bs > 1:  # any bilinear family of the space does: only the pattern of the form is used
            a.add_cell_integral("elasticity", cells[: mesh.num_cells_local].contiguous(), None, (1.0, 1.0))
        else:
            a.add_cell_integral("laplace", cells[: mesh.num_cells_local].contiguous(), None, (1.0,))
        a.add_interior_facet_integral("ghost_grad_jump", rows=rows4, constants=(1.0,))
        Ag = _fem.create_ghost_row_pattern(a, im.n_owned)
        rp, ci = Ag.indptr_device(), Ag.indices_device()
        self._cand, sends = {}, {}
        for q, sel in self.vx.send_sel.items():          # ghost rows owned by q, ascending
            rowrep, cols, ptr = candidate_entries(rp, ci, sel)
            self._cand[q] = (ptr, cols)
            sends[q] = torch.stack([im.l2g[rowrep], im.l2g[cols]]).reshape(-1).contiguous()
        torch.cuda.synchronize(dev)
        a.free()
        Ag.free()
        rows4.free()
        facets.free()
        # that pattern had every row in the band: do not let it size the per-step lists
        check(ctx.handle, lib().cfx_space_forget(ctx.handle, ctx.space_index(V)))
        return sends

    def static_finish(self, recv):
        if self.world == 1:
            return
        tables, new_globals = static_plan_tables(self.imap, self.vx.send_sel, self._cand, recv, self.vx.recv_pos)
        self._cand = None
        sidx = self.ctx.space_index(self.V)
        self.xplan = StaticExchange(self.ctx, sidx, tables, new_globals)

    # per step with the static plan: every call below is asynchronous (and capturable in deferred-size mode)
    def sphase_a(self):
        a, L = self.prob.build_forms()
        if self.xplan is not None:
            self.xplan.pack_pattern(a)

    def sphase_b(self):
        a = self.prob.last["a"]
        if self.xplan is not None:
            self.xplan.insert_pattern(a)
        stats = self.prob.assemble()
        if self.xplan is not None:
            self.xplan.pack_values(self.prob.A, self.prob.b)
        return stats

    def sphase_c(self):
        if self.xplan is not None:
            self.xplan.unpack_add(self.prob.A, self.prob.b)

    def step_static(self):
        """One step of this rank over NCCL (one process per GPU): no host round trip, no Python-side exchange."""
        self.sphase_a()
        if self.xplan is not None:
            self.xplan.exchange(0)
        stats = self.sphase_b()
        if self.xplan is not None:
            self.xplan.exchange(1)
        self.sphase_c()
        self.prob.release_step()
        return stats

    def capture_static(self, margin: float = 0.25):
        """Record step_static as one CUDA graph (NCCL calls included): two eager steps, one deferred, capture."""
        prob, ctx = self.prob, self.ctx
        prob.persistent = True
        ctx.set_deferred(False, margin)
        for _ in range(2):
            self.step_static()
        ctx.set_deferred(True)
        self.step_static()
        ctx.check()
        ctx.graph_begin()
        try:
            self.step_static()
        finally:
            prob.graph = ctx.graph_end()
        return prob.graph

    def phase_a(self):
        from . import fem as _fem

        a, L = self.prob.build_forms()
        if self.world == 1:
            return {}
        self.Ag = _fem.create_ghost_row_pattern(a, self.imap.n_owned, self.Ag)
        rng = self.imap.owner_ranges()
        if rng is None:
            rows, cols = coo_of_rows(self.Ag.indptr_device(), self.Ag.indices_device(), self.imap.n_owned,
                                     self.imap.n_total)
            return self.mx.begin(rows, cols)
        cuts = [b for _, _, b in rng[:-1]]
        rows, cols, offs = coo_of_rows(self.Ag.indptr_device(), self.Ag.indices_device(), self.imap.n_owned,
                                       self.imap.n_total, cuts=cuts) if cuts else \
            (*coo_of_rows(self.Ag.indptr_device(), self.Ag.indices_device(), self.imap.n_owned, self.imap.n_total), [])
        bounds = [0] + offs + [int(rows.numel())]
        return self.mx.begin(rows, cols, {q: (bounds[k], bounds[k + 1]) for k, (q, _, _) in enumerate(rng)})

    def phase_b(self, recv):
        import torch

        from . import fem as _fem

        a = self.prob.last["a"]
        if self.world > 1:
            xr, xc = self.mx.inserted_entries(recv, self.disjoint_sources)
            _fem.insert_pattern_entries(a, xr, xc)
        stats = self.prob.assemble()
        if self.world == 1:
            return {}, stats
        A, b = self.prob.A, self.prob.b
        self.mx.finish(lambda r, c: self.ops.positions(A, r, c))
        vals = A.values_device()
        sends = {}
        for q in sorted(set(self.mx.send_pos) | set(self.vx.send_sel)):
            parts = []
            if q in self.mx.send_pos:
                parts.append(self.ops.gather(vals, self.mx.send_pos[q]))
            if q in self.vx.send_sel:
                parts.append(b[self.vx.send_sel[q]])
            sends[q] = torch.cat(parts)
        return sends, stats

    def recv_counts(self):
        m, v = self.mx.recv_counts(), self.vx.recv_counts()
        return {q: m.get(q, 0) + v.get(q, 0) for q in set(m) | set(v)}

    def phase_c(self, recv):
        if self.world == 1:
            return
        A, b = self.prob.A, self.prob.b
        vals = A.values_device()
        for q in sorted(recv):  # fixed neighbour order
            t = recv[q]
            nm = int(self.mx.recv_pos[q].numel()) if q in self.mx.recv_pos else 0
            if nm:
                self.ops.scatter_add(vals, self.mx.recv_pos[q], t[:nm].contiguous())
            if q in self.vx.recv_pos:
                self.ops.scatter_add(b, self.vx.recv_pos[q], t[nm:].contiguous())

    def finish_step(self):
        self.prob.release_step()

    def move_level_set(self, params):
        """Re-interpolate the (device-resident) level set with new parameters, in place: the next
        cutfemx.update() re-cuts it (demo_moving_poisson.py:69-73)."""
        from . import demo_poisson as dp

        dp.device_level_set(self.mesh, self.ls_kind, params, out=self.phi.x.array)
        self.ls_params = params

    # results in global numbering (owned rows only) -- for the parity tests
    def owned_matrix_global(self):
        import scipy.sparse as sp

        A = self.prob.A
        rp, cols, vals = A.indptr, A.indices.astype(np.int64), A.data
        n_owned, n_total = self.imap.n_owned, self.imap.n_total
        l2g = self.imap.l2g.cpu().numpy()
        ng = self.xplan.new_ghost_globals if getattr(self, "xplan", None) is not None else self.mx.new_ghost_globals
        extra = ng.cpu().numpy() if ng is not None else np.zeros(0, np.int64)
        colmap = np.concatenate([l2g, extra])
        e = int(rp[n_owned])
        rows = np.repeat(np.arange(n_owned), np.diff(rp[: n_owned + 1]))
        bs = int(getattr(self.V, "bs", 1) or 1)
        b = self.prob.b.cpu().numpy()[: n_owned * bs]
        if bs > 1:  # blocked space: rows / columns are block indices, values (entries, bs, bs), b (n_owned, bs)
            return (rows + self.imap.offset, colmap[cols[:e]], vals[: e * bs * bs].reshape(e, bs, bs),
                    b.reshape(n_owned, bs), self.imap.offset)
        return rows + self.imap.offset, colmap[cols[:e]], vals[:e], b, self.imap.offset


def run_step(pipes, transport):
    """One step of the pipeline for the ranks this process hosts (1 with torch.distributed)."""
    import torch

    recv = transport.exchange([p.phase_a() for p in pipes], dtype=torch.int64)
    out = [p.phase_b(r) for p, r in zip(pipes, recv)]
    recv = transport.exchange([o[0] for o in out], counts=[p.recv_counts() for p in pipes], dtype=torch.float64)
    for p, r in zip(pipes, recv):
        p.phase_c(r)
    return [o[1] for o in out]


def run_step_static(pipes):
    """One step of ranks EMULATED on one GPU with the static plans: the kernels of the exchange are the product's
    (cfx_xplan_*), only the transport is a device-to-device copy between the emulated ranks' message buffers."""
    by_rank = {p.rank: p for p in pipes}

    def transfer(which_send, which_recv):
        for p in pipes:
            if p.xplan is None:
                continue
            for k, q in enumerate(p.xplan.neigh):
                dst = by_rank[q].xplan
                src_buf = p.xplan.buffer(which_send, k)
                if src_buf.numel():
                    dst.buffer(which_recv, dst.neigh.index(p.rank)).copy_(src_buf)

    for p in pipes:
        p.sphase_a()
    transfer(0, 1)
    out = [p.sphase_b() for p in pipes]
    transfer(2, 3)
    for p in pipes:
        p.sphase_c()
    return out


def plan(pipes, transport, static: bool = False):
    import torch

    recv = transport.exchange([p.plan_begin() for p in pipes], dtype=torch.int64)
    for p, r in zip(pipes, recv):
        p.plan_finish(r)
    if static:
        recv = transport.exchange([p.static_begin() for p in pipes], dtype=torch.int64)
        for p, r in zip(pipes, recv):
            p.static_finish(r)
