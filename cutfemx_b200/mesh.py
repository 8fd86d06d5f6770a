"""Synthetic background meshes in DOLFINx array layout (host side, numpy).

DOLFINx 0.11 is not installable in the build image, so this
module produces the flat arrays CutFEMx borrows from a `dolfinx.mesh.Mesh`:

* ``x``          geometry().x(), shape (num_nodes, 3) float64, z padded with 0 in 2D
                 (reference: cpp/cutfemx/cut/cut.cpp:525-531, always stride 3);
* ``x_dofmap``   geometry().dofmaps().front(), shape (num_cells, nv) int32;
* ``c2f``        topology().connectivity(tdim, tdim-1), shape (num_cells, tdim+1) int32,
                 facet ``i`` is the one opposite local vertex ``i`` (Basix convention);
* ``f2c_offsets``/``f2c`` topology().connectivity(tdim-1, tdim) as an AdjacencyList.

Numbering (the same formulas are used by the device generator in csrc/meshgen.cu so that
large meshes never exist on the host): vertices lexicographic; cells grouped per
quad/hexahedron, 2 triangles with the "right" diagonal / 6 Kuhn tetrahedra sharing the
main diagonal; facets numbered per *virtual* cell ``vc`` (one per vertex, the cell whose
lowest corner is that vertex), 3 (2D) or 12 (3D) slots each -- slots that do not exist on
the upper boundary stay empty facets with zero cells.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

TRIANGLE = 3      # cell_type codes == number of vertices (cfx_cell_type in include/cutfemx_b200.h)
TETRAHEDRON = 4

# Kuhn tetrahedra of the unit cube, corner index bits: 1 = +x, 2 = +y, 4 = +z.
KUHN_TETS = np.array(
    [(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)],
    dtype=np.int32,
)
# interior faces (0, m, 7): slot t <-> m
_TET_INTERIOR_M = (1, 2, 4, 3, 5, 6)
# boundary-square triangles (corner sets on the *low* side square of a virtual cube): slot 6..11
_TET_SQUARE_TRIS = (
    ((0, 2, 6), 6), ((0, 4, 6), 7),    # x-low square
    ((0, 1, 5), 8), ((0, 4, 5), 9),    # y-low square
    ((0, 1, 3), 10), ((0, 2, 3), 11),  # z-low square
)


def kuhn_facet_table():
    """For each Kuhn tet and local facet: (di, dj, dk, slot) of the facet's virtual cube."""
    table = np.zeros((6, 4, 4), dtype=np.int32)
    for t, tet in enumerate(KUHN_TETS):
        for lf in range(4):
            face = sorted(int(v) for k, v in enumerate(tet) if k != lf)
            if face[0] == 0 and face[2] == 7:
                table[t, lf] = (0, 0, 0, _TET_INTERIOR_M.index(face[1]))
                continue
            if face[0] == 0:
                shift = 0  # lies on a low-side square of this cube
                rel = tuple(face)
            else:
                # lies on the high side of direction a = lowest corner: shift to the neighbour
                shift = face[0]
                rel = tuple(v - shift for v in face)
            slot = dict(_TET_SQUARE_TRIS)[rel]
            table[t, lf] = (shift & 1, (shift >> 1) & 1, (shift >> 2) & 1, slot)
    return table


@dataclass
class Mesh:
    """Flat-array stand-in for the parts of dolfinx.mesh.Mesh the cut path reads."""

    cell_type: int
    tdim: int
    gdim: int
    x: np.ndarray
    x_dofmap: np.ndarray
    c2f: np.ndarray
    f2c_offsets: np.ndarray
    f2c: np.ndarray
    num_cells_local: int          # cell index_map size_local()  (owned cells)
    num_facets: int
    num_owned_facets: int
    shape: tuple = ()
    p0: tuple = ()
    p1: tuple = ()
    extra: dict = field(default_factory=dict)

    @property
    def num_cells(self) -> int:
        return int(self.x_dofmap.shape[0])

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def nv(self) -> int:
        return int(self.x_dofmap.shape[1])


def invert_c2f(c2f: np.ndarray, num_facets: int):
    """Facet->cell AdjacencyList (cells ascending per facet) from cell->facet."""
    nc, nf = c2f.shape
    flat = c2f.reshape(-1).astype(np.int64)
    order = np.argsort(flat, kind="stable")
    cells = (order // nf).astype(np.int32)
    counts = np.bincount(flat, minlength=num_facets)
    offsets = np.zeros(num_facets + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    return offsets, cells


def create_rectangle(nx: int, ny: int, p0=(-1.0, -1.0), p1=(1.0, 1.0)) -> Mesh:
    """nx x ny quads, each split into 2 triangles along the "right" diagonal."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # [j, i]
    x = np.zeros(((nx + 1) * (ny + 1), 3))
    x[:, 0] = X.reshape(-1)
    x[:, 1] = Y.reshape(-1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    i = i.reshape(-1)
    j = j.reshape(-1)
    sx = nx + 1
    v00 = j * sx + i
    v10, v01, v11 = v00 + 1, v00 + sx, v00 + sx + 1
    cells = np.empty((nx * ny, 2, 3), dtype=np.int32)
    cells[:, 0] = np.stack([v00, v10, v11], axis=1)
    cells[:, 1] = np.stack([v00, v01, v11], axis=1)
    c2f = np.empty((nx * ny, 2, 3), dtype=np.int32)
    # triangle A = (v00, v10, v11): opposite v00 -> x-low edge of quad (i+1, j); v10 -> diagonal; v11 -> y-low
    c2f[:, 0] = np.stack([3 * v10 + 1, 3 * v00 + 0, 3 * v00 + 2], axis=1)
    # triangle B = (v00, v01, v11): opposite v00 -> y-low edge of quad (i, j+1); v01 -> diagonal; v11 -> x-low
    c2f[:, 1] = np.stack([3 * v01 + 2, 3 * v00 + 0, 3 * v00 + 1], axis=1)
    x_dofmap = np.ascontiguousarray(cells.reshape(-1, 3))
    c2f = np.ascontiguousarray(c2f.reshape(-1, 3))
    nfac = 3 * (nx + 1) * (ny + 1)
    off, f2c = invert_c2f(c2f, nfac)
    return Mesh(TRIANGLE, 2, 2, x, x_dofmap, c2f, off, f2c, x_dofmap.shape[0], nfac, nfac,
                shape=(nx, ny), p0=tuple(p0), p1=tuple(p1))


def create_box(nx: int, ny: int, nz: int, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0)) -> Mesh:
    """nx x ny x nz hexahedra, each split into the 6 Kuhn tetrahedra."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # [k, j, i]
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    v0 = (k * (ny + 1) + j) * sx + i
    corner = lambda b: v0 + (b & 1) + ((b >> 1) & 1) * sx + ((b >> 2) & 1) * sy  # noqa: E731
    ncube = nx * ny * nz
    cells = np.empty((ncube, 6, 4), dtype=np.int32)
    c2f = np.empty((ncube, 6, 4), dtype=np.int32)
    ftab = kuhn_facet_table()
    for t in range(6):
        for lv in range(4):
            cells[:, t, lv] = corner(int(KUHN_TETS[t, lv]))
            di, dj, dk, slot = (int(a) for a in ftab[t, lv])
            c2f[:, t, lv] = 12 * (v0 + di + dj * sx + dk * sy) + slot
    x_dofmap = np.ascontiguousarray(cells.reshape(-1, 4))
    c2f = np.ascontiguousarray(c2f.reshape(-1, 4))
    nfac = 12 * (nx + 1) * (ny + 1) * (nz + 1)
    off, f2c = invert_c2f(c2f, nfac)
    return Mesh(TETRAHEDRON, 3, 3, x, x_dofmap, c2f, off, f2c, x_dofmap.shape[0], nfac, nfac,
                shape=(nx, ny, nz), p0=tuple(p0), p1=tuple(p1))


# ----------------------------------------------------------------------------- function spaces
@dataclass
class FunctionSpace:
    """Scalar/blocked Lagrange space: the arrays of dolfinx.fem.DofMap the path reads."""

    mesh: Mesh
    degree: int
    dofmap: np.ndarray            # (num_cells, nd) int32, DOLFINx/Basix local ordering
    num_dofs: int                 # size_local + num_ghosts
    num_dofs_owned: int
    bs: int = 1
    dof_coords: np.ndarray | None = None  # tabulate_dof_coordinates(), (num_dofs, 3)

    @property
    def nd(self) -> int:
        return int(self.dofmap.shape[1])


# Basix sub-entity numbering: edge e of a triangle joins the two vertices other than e;
# tetrahedron edges: e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1).
TRI_EDGES = ((1, 2), (0, 2), (0, 1))
TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def functionspace(mesh: Mesh, degree: int, bs: int = 1, permute_seed: int | None = None) -> FunctionSpace:
    """Lagrange P1/P2 dofmap. `permute_seed` renumbers dofs randomly (DOLFINx dof numbers
    are unrelated to vertex numbers; tests use this to make sure nothing assumes they agree)."""
    if degree == 1:
        dofmap = mesh.x_dofmap.copy()
        ndofs = mesh.num_nodes
        coords = mesh.x.copy()
    elif degree == 2:
        edges = TRI_EDGES if mesh.cell_type == TRIANGLE else TET_EDGES
        nn = mesh.num_nodes
        a = np.stack([mesh.x_dofmap[:, e[0]] for e in edges], axis=1).astype(np.int64)
        b = np.stack([mesh.x_dofmap[:, e[1]] for e in edges], axis=1).astype(np.int64)
        key = np.minimum(a, b) * nn + np.maximum(a, b)
        uniq, inv = np.unique(key.reshape(-1), return_inverse=True)
        edofs = (nn + inv).reshape(key.shape).astype(np.int32)
        dofmap = np.concatenate([mesh.x_dofmap, edofs], axis=1).astype(np.int32)
        ndofs = nn + uniq.size
        coords = np.concatenate([mesh.x, 0.5 * (mesh.x[uniq // nn] + mesh.x[uniq % nn])], axis=0)
    else:
        raise ValueError("only P1 and P2 Lagrange spaces are generated")
    if permute_seed is not None:
        perm = np.random.default_rng(permute_seed).permutation(ndofs).astype(np.int32)
        dofmap = perm[dofmap]
        inv = np.empty_like(perm)
        inv[perm] = np.arange(ndofs, dtype=np.int32)
        coords = coords[inv]
    return FunctionSpace(mesh, degree, np.ascontiguousarray(dofmap, dtype=np.int32), ndofs, ndofs, bs, coords)


# ----------------------------------------------------------------------------- level sets
def interpolate(space: FunctionSpace, fn) -> np.ndarray:
    """Nodal interpolation: values at the dof coordinates (dolfinx Function.interpolate)."""
    c = space.dof_coords
    return np.ascontiguousarray(fn(c[:, 0], c[:, 1], c[:, 2]), dtype=np.float64)


def sphere_level_set(center, radius):
    cx, cy, cz = (tuple(center) + (0.0, 0.0, 0.0))[:3]
    return lambda x, y, z: np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) - radius


def torus_level_set(center, R, r):
    cx, cy, cz = center
    return lambda x, y, z: np.sqrt((np.sqrt((x - cx) ** 2 + (y - cy) ** 2) - R) ** 2 + (z - cz) ** 2) - r


class _Vector:
    """dolfinx.la.Vector stand-in: `.array` is the owned+ghost value array."""

    def __init__(self, array):
        self.array = array

    def scatter_forward(self):  # single-rank meshes: nothing to do (multi-rank: parallel.py)
        return None


class Function:
    """dolfinx.fem.Function stand-in: a function space, a name and `.x.array`."""

    def __init__(self, space: FunctionSpace, name: str = "f", array=None):
        self.function_space = space
        self.name = name
        self.x = _Vector(np.zeros(space.num_dofs) if array is None else array)

    def interpolate(self, fn):
        self.x.array[:] = interpolate(self.function_space, fn)
        return self
