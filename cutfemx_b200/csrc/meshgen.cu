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
} // extern "C"
