// meshgen.cu -- harness: device-side generator of the synthetic background meshes of
// cutfemx_b200/mesh.py (identical numbering), and level-set nodal interpolation.  DOLFINx is not
// available in the build image, and a 256^3 Kuhn mesh (100.7 M tets) should never exist on the
// host; this stands in for dolfinx.mesh.create_box/create_rectangle + Function.interpolate.
#include "common.cuh"

namespace cfx
{
namespace
{
constexpr int MB = 256;

__constant__ int8_t c_kuhn[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
// (di, dj, dk, slot) per Kuhn tet and local facet: mesh.py kuhn_facet_table()
__constant__ int8_t c_kuhn_facets[6][4][4] = {
    {{1, 0, 0, 6}, {0, 0, 0, 3}, {0, 0, 0, 0}, {0, 0, 0, 10}}, {{1, 0, 0, 7}, {0, 0, 0, 4}, {0, 0, 0, 0}, {0, 0, 0, 8}},
    {{0, 1, 0, 8}, {0, 0, 0, 3}, {0, 0, 0, 1}, {0, 0, 0, 11}}, {{0, 1, 0, 9}, {0, 0, 0, 5}, {0, 0, 0, 1}, {0, 0, 0, 6}},
    {{0, 0, 1, 10}, {0, 0, 0, 4}, {0, 0, 0, 2}, {0, 0, 0, 9}}, {{0, 0, 1, 11}, {0, 0, 0, 5}, {0, 0, 0, 2}, {0, 0, 0, 7}}};

__global__ void box_nodes_kernel(int nx, int ny, int nz, double x0, double y0, double z0, double x1, double y1,
                                 double z1, double* __restrict__ x)
{
  const int64_t n = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  const int64_t v = static_cast<int64_t>(blockIdx.x) * MB + threadIdx.x;
  if (v >= n)
    return;
  const int i = static_cast<int>(v % (nx + 1));
  const int j = static_cast<int>((v / (nx + 1)) % (ny + 1));
  const int k = static_cast<int>(v / (static_cast<int64_t>(nx + 1) * (ny + 1)));
  // numpy.linspace: start + i * step with separately rounded product and sum (no FMA contraction,
  // so that host- and device-generated meshes agree bit for bit), last point exact
  x[3 * v + 0] = i == nx ? x1 : __dadd_rn(x0, __dmul_rn(static_cast<double>(i), (x1 - x0) / nx));
  x[3 * v + 1] = j == ny ? y1 : __dadd_rn(y0, __dmul_rn(static_cast<double>(j), (y1 - y0) / ny));
  x[3 * v + 2] = (nz > 0) ? (k == nz ? z1 : __dadd_rn(z0, __dmul_rn(static_cast<double>(k), (z1 - z0) / nz))) : 0.0;
}

__global__ void box_cells_kernel(int nx, int ny, int nz, int32_t* __restrict__ x_dofmap, int32_t* __restrict__ c2f)
{
  const int64_t ncell = static_cast<int64_t>(nx) * ny * nz * 6;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * MB + threadIdx.x;
  if (c >= ncell)
    return;
  const int t = static_cast<int>(c % 6);
  const int64_t cube = c / 6;
  const int i = static_cast<int>(cube % nx);
  const int j = static_cast<int>((cube / nx) % ny);
  const int k = static_cast<int>(cube / (static_cast<int64_t>(nx) * ny));
  const int64_t sx = nx + 1, sy = static_cast<int64_t>(nx + 1) * (ny + 1);
  const int64_t v0 = (static_cast<int64_t>(k) * (ny + 1) + j) * sx + i;
  int4 dm, cf;
  int32_t* dmp = reinterpret_cast<int32_t*>(&dm);
  int32_t* cfp = reinterpret_cast<int32_t*>(&cf);
#pragma unroll
  for (int lv = 0; lv < 4; ++lv)
  {
    const int b = c_kuhn[t][lv];
    dmp[lv] = static_cast<int32_t>(v0 + (b & 1) + ((b >> 1) & 1) * sx + ((b >> 2) & 1) * sy);
    const int8_t* ft = c_kuhn_facets[t][lv];
    cfp[lv] = static_cast<int32_t>(12 * (v0 + ft[0] + ft[1] * sx + ft[2] * sy) + ft[3]);
  }
  reinterpret_cast<int4*>(x_dofmap)[c] = dm;
  reinterpret_cast<int4*>(c2f)[c] = cf;
}

__global__ void rect_cells_kernel(int nx, int ny, int32_t* __restrict__ x_dofmap, int32_t* __restrict__ c2f)
{
  const int64_t ncell = static_cast<int64_t>(nx) * ny * 2;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * MB + threadIdx.x;
  if (c >= ncell)
    return;
  const int t = static_cast<int>(c & 1);
  const int64_t q = c >> 1;
  const int i = static_cast<int>(q % nx), j = static_cast<int>(q / nx);
  const int32_t sx = nx + 1;
  const int32_t v00 = j * sx + i, v10 = v00 + 1, v01 = v00 + sx, v11 = v00 + sx + 1;
  if (t == 0)
  {
    x_dofmap[3 * c + 0] = v00;
    x_dofmap[3 * c + 1] = v10;
    x_dofmap[3 * c + 2] = v11;
    c2f[3 * c + 0] = 3 * v10 + 1;
    c2f[3 * c + 1] = 3 * v00 + 0;
    c2f[3 * c + 2] = 3 * v00 + 2;
  }
  else
  {
    x_dofmap[3 * c + 0] = v00;
    x_dofmap[3 * c + 1] = v01;
    x_dofmap[3 * c + 2] = v11;
    c2f[3 * c + 0] = 3 * v01 + 2;
    c2f[3 * c + 1] = 3 * v00 + 0;
    c2f[3 * c + 2] = 3 * v00 + 1;
  }
}

__global__ void level_set_kernel(const double* __restrict__ x, int64_t n, int kind, double cx, double cy, double cz,
                                 double R, double r, double* __restrict__ out)
{
  const int64_t v = static_cast<int64_t>(blockIdx.x) * MB + threadIdx.x;
  if (v >= n)
    return;
  const double dx = x[3 * v] - cx, dy = x[3 * v + 1] - cy, dz = x[3 * v + 2] - cz;
  if (kind == 0)
    out[v] = sqrt(dx * dx + dy * dy + dz * dz) - R;
  else
  {
    const double a = sqrt(dx * dx + dy * dy) - R;
    out[v] = sqrt(a * a + dz * dz) - r;
  }
}
// P2 Lagrange dofmap of a Kuhn box mesh: the 4 vertex dofs, then one dof per edge in Basix order
// e0 = (2,3), e1 = (1,3), e2 = (1,2), e3 = (0,3), e4 = (0,2), e5 = (0,1).  The vertex offsets of a Kuhn tetrahedron
// are nested bit patterns, so an edge is (lower vertex, direction) with one of 7 directions (3 axes, 3 face
// diagonals, the cube diagonal): edge dof = n_nodes + 7 * lower vertex + direction -- a sparse numbering like the
// facet numbering of box_cells_kernel (ids of edges that would leave the box stay unused).
__global__ void p2_tet_dofmap_kernel(const int32_t* __restrict__ x_dofmap, int64_t n_cells, int64_t n_nodes,
                                     int64_t sx, int64_t sy, int32_t* __restrict__ dofmap)
{
  const int64_t c = static_cast<int64_t>(blockIdx.x) * MB + threadIdx.x;
  if (c >= n_cells)
    return;
  const int4 v4 = reinterpret_cast<const int4*>(x_dofmap)[c];
  const int32_t v[4] = {v4.x, v4.y, v4.z, v4.w};
  int32_t* o = dofmap + c * 10;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = v[j];
  constexpr int ea[6] = {2, 1, 1, 0, 0, 0}, eb[6] = {3, 3, 2, 3, 2, 1};
#pragma unroll
  for (int e = 0; e < 6; ++e)
  {
    const int64_t a = min(v[ea[e]], v[eb[e]]), b = max(v[ea[e]], v[eb[e]]);
    const int64_t diff = b - a;
    const int dk = static_cast<int>(diff / sy);
    const int64_t rem = diff - dk * sy;
    const int dj = static_cast<int>(rem / sx);
    const int di = static_cast<int>(rem - dj * sx);
    const int dir = (di | (dj << 1) | (dk << 2)) - 1;
    o[4 + e] = static_cast<int32_t>(n_nodes + 7 * a + dir);
  }
}
} // namespace
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_meshgen_box(cfx_ctx* ctx, int nx, int ny, int nz, const double p0[3], const double p1[3], double* x,
                           int32_t* x_dofmap, int32_t* c2f)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && x && x_dofmap && c2f && nx > 0 && ny > 0 && nz > 0, CFX_ERR_INVALID, "cfx_meshgen_box: bad args");
  const int64_t nn = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  CFX_REQUIRE(12 * nn < (int64_t(1) << 31), CFX_ERR_RANGE, "cfx_meshgen_box: facet ids exceed int32");
  CFX_LAUNCH(ctx, box_nodes_kernel, grid_for(nn, MB), MB, 0, nx, ny, nz, p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], x);
  CFX_LAUNCH(ctx, box_cells_kernel, grid_for(static_cast<int64_t>(nx) * ny * nz * 6, MB), MB, 0, nx, ny, nz, x_dofmap,
             c2f);
  CFX_API_END(ctx)
}

cfx_status cfx_meshgen_rectangle(cfx_ctx* ctx, int nx, int ny, const double p0[2], const double p1[2], double* x,
                                 int32_t* x_dofmap, int32_t* c2f)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && x && x_dofmap && c2f && nx > 0 && ny > 0, CFX_ERR_INVALID, "cfx_meshgen_rectangle: bad args");
  const int64_t nn = static_cast<int64_t>(nx + 1) * (ny + 1);
  CFX_LAUNCH(ctx, box_nodes_kernel, grid_for(nn, MB), MB, 0, nx, ny, 0, p0[0], p0[1], 0.0, p1[0], p1[1], 0.0, x);
  CFX_LAUNCH(ctx, rect_cells_kernel, grid_for(static_cast<int64_t>(nx) * ny * 2, MB), MB, 0, nx, ny, x_dofmap, c2f);
  CFX_API_END(ctx)
}

cfx_status cfx_meshgen_level_set(cfx_ctx* ctx, const double* x, int64_t n_nodes, int kind, const double params[5],
                                 double* values)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && x && values && params, CFX_ERR_INVALID, "cfx_meshgen_level_set: bad args");
  CFX_LAUNCH(ctx, level_set_kernel, grid_for(n_nodes, MB), MB, 0, x, n_nodes, kind, params[0], params[1], params[2],
             params[3], params[4], values);
  CFX_API_END(ctx)
}
cfx_status cfx_meshgen_p2_tet_dofmap(cfx_ctx* ctx, int nx, int ny, int nz, const int32_t* x_dofmap, int64_t n_cells,
                                     int32_t* dofmap, int64_t* n_dofs)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && x_dofmap && dofmap && nx > 0 && ny > 0 && nz > 0, CFX_ERR_INVALID,
              "cfx_meshgen_p2_tet_dofmap: bad args");
  const int64_t nn = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  CFX_REQUIRE(8 * nn < (int64_t(1) << 31), CFX_ERR_RANGE, "cfx_meshgen_p2_tet_dofmap: dof ids exceed int32");
  CFX_LAUNCH(ctx, p2_tet_dofmap_kernel, grid_for(n_cells, MB), MB, 0, x_dofmap, n_cells, nn,
             static_cast<int64_t>(nx + 1), static_cast<int64_t>(nx + 1) * (ny + 1), dofmap);
  if (n_dofs)
    *n_dofs = 8 * nn;
  CFX_API_END(ctx)
}
} // extern "C"
