// element.cuh -- affine simplex geometry and Lagrange P1/P2 tabulation on the device
// (Basix/DOLFINx reference-cell and dof ordering).  The role Basix tabulation and
// CoordinateElement::compute_jacobian(_inverse) play inside the reference's generated kernels
// and at cpp/cutfemx/level_set/normal.h:84-99,151-152.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace cfx
{
template <int TDIM>
struct Geo
{
  double J[TDIM * TDIM]; // J[r*TDIM + t] = dx_r / dX_t
  double K[TDIM * TDIM]; // K[t*TDIM + r] = dX_t / dx_r
  double x0[TDIM];
  double detJ;
};

// X: vertex coordinates, X[v][r]
template <int TDIM>
__device__ __forceinline__ void make_geo(const double (&X)[TDIM + 1][TDIM], Geo<TDIM>& g)
{
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
  {
    g.x0[r] = X[0][r];
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      g.J[r * TDIM + t] = X[t + 1][r] - X[0][r];
  }
  if constexpr (TDIM == 2)
  {
    const double a = g.J[0], b = g.J[1], c = g.J[2], d = g.J[3];
    g.detJ = a * d - b * c;
    const double id = 1.0 / g.detJ;
    g.K[0] = d * id;
    g.K[1] = -b * id;
    g.K[2] = -c * id;
    g.K[3] = a * id;
  }
  else
  {
    const double* J = g.J;
    const double c00 = J[4] * J[8] - J[5] * J[7];
    const double c01 = J[5] * J[6] - J[3] * J[8];
    const double c02 = J[3] * J[7] - J[4] * J[6];
    g.detJ = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double id = 1.0 / g.detJ;
    g.K[0] = c00 * id;
    g.K[1] = (J[2] * J[7] - J[1] * J[8]) * id;
    g.K[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    g.K[3] = c01 * id;
    g.K[4] = (J[0] * J[8] - J[2] * J[6]) * id;
    g.K[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    g.K[6] = c02 * id;
    g.K[7] = (J[1] * J[6] - J[0] * J[7]) * id;
    g.K[8] = (J[0] * J[4] - J[1] * J[3]) * id;
  }
}

// gather the TDIM+1 vertex coordinates of a cell (x has stride 3, cut.cpp:529)
template <int TDIM>
__device__ __forceinline__ void load_cell_coords(const double* __restrict__ x, const int32_t* __restrict__ x_dofmap,
                                                 int64_t cell, double (&X)[TDIM + 1][TDIM])
{
  constexpr int NV = TDIM + 1;
  int32_t node[NV];
  if constexpr (NV == 4)
  {
    const int4 v = __ldg(reinterpret_cast<const int4*>(x_dofmap) + cell);
    node[0] = v.x;
    node[1] = v.y;
    node[2] = v.z;
    node[3] = v.w;
  }
  else
  {
#pragma unroll
    for (int v = 0; v < NV; ++v)
      node[v] = __ldg(x_dofmap + cell * NV + v);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
      X[v][r] = __ldg(x + 3 * static_cast<int64_t>(node[v]) + r);
}

// static geometry cache record (common.cuh cfx_ctx::geo): K row-major, detJ, h
template <int TDIM>
struct GeoRec
{
  static constexpr int STRIDE = TDIM == 3 ? 12 : 8;
};

// K and detJ of a cell from the cache (J and x0 are not filled).  Records are 32-byte aligned:
// 256-bit loads (one L1 wavefront per 32 B sector instead of two)
template <int TDIM>
__device__ __forceinline__ void load_geo_cached(const double* __restrict__ geo, int64_t cell, Geo<TDIM>& g)
{
  const double* p = geo + cell * GeoRec<TDIM>::STRIDE;
  double a0, a1, a2, a3, b0, b1, b2, b3;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(p + 4));
  if constexpr (TDIM == 3)
  {
    double c0, c1, c2, c3;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(p + 8));
    g.K[0] = a0; g.K[1] = a1; g.K[2] = a2; g.K[3] = a3; g.K[4] = b0;
    g.K[5] = b1; g.K[6] = b2; g.K[7] = b3; g.K[8] = c0; g.detJ = c1;
  }
  else
  {
    g.K[0] = a0; g.K[1] = a1; g.K[2] = a2; g.K[3] = a3; g.detJ = b0;
  }
}

// UFL CellDiameter: largest vertex-to-vertex distance
template <int TDIM>
__device__ __forceinline__ double cell_diameter(const double (&X)[TDIM + 1][TDIM])
{
  double h2 = 0.0;
#pragma unroll
  for (int a = 0; a < TDIM + 1; ++a)
#pragma unroll
    for (int b = a + 1; b < TDIM + 1; ++b)
    {
      double d2 = 0.0;
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
      {
        const double d = X[a][r] - X[b][r];
        d2 += d * d;
      }
      h2 = fmax(h2, d2);
    }
  return sqrt(h2);
}

template <int TDIM, int DEG>
struct Elem
{
  static constexpr int NV = TDIM + 1;
  static constexpr int ND = DEG == 1 ? TDIM + 1 : (TDIM == 2 ? 6 : 10);
};

// Basix sub-entity numbering of edges (vertex pairs)
__device__ __forceinline__ void edge_vertices(int tdim, int e, int& a, int& b)
{
  if (tdim == 2)
  { // (1,2) (0,2) (0,1)
    a = e == 0 ? 1 : 0;
    b = e == 2 ? 1 : 2;
  }
  else
  { // (2,3) (1,3) (1,2) (0,3) (0,2) (0,1)
    constexpr int EA[6] = {2, 1, 1, 0, 0, 0};
    constexpr int EB[6] = {3, 3, 2, 3, 2, 1};
    a = EA[e];
    b = EB[e];
  }
}

// phi[i], dphi[i][t] (reference gradients) at reference point Xr
template <int TDIM, int DEG>
__device__ __forceinline__ void tabulate(const double (&Xr)[TDIM], double (&phi)[Elem<TDIM, DEG>::ND],
                                         double (&dphi)[Elem<TDIM, DEG>::ND][TDIM])
{
  constexpr int NV = TDIM + 1;
  double lam[NV];
  lam[0] = 1.0;
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
  {
    lam[0] -= Xr[t];
    lam[t + 1] = Xr[t];
  }
  auto dlam = [](int v, int t) -> double { return v == 0 ? -1.0 : (v - 1 == t ? 1.0 : 0.0); };
  if constexpr (DEG == 1)
  {
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      phi[v] = lam[v];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        dphi[v][t] = dlam(v, t);
    }
  }
  else
  {
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      phi[v] = lam[v] * (2.0 * lam[v] - 1.0);
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        dphi[v][t] = (4.0 * lam[v] - 1.0) * dlam(v, t);
    }
    constexpr int NE = TDIM == 2 ? 3 : 6;
#pragma unroll
    for (int e = 0; e < NE; ++e)
    {
      int a, b;
      edge_vertices(TDIM, e, a, b);
      phi[NV + e] = 4.0 * lam[a] * lam[b];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        dphi[NV + e][t] = 4.0 * (lam[a] * dlam(b, t) + lam[b] * dlam(a, t));
    }
  }
}

// physical gradients grad[i][r] = sum_t K[t][r] dphi[i][t]
template <int TDIM, int ND>
__device__ __forceinline__ void push_gradients(const Geo<TDIM>& g, const double (&dphi)[ND][TDIM],
                                               double (&grad)[ND][TDIM])
{
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        s += g.K[t * TDIM + r] * dphi[i][t];
      grad[i][r] = s;
    }
}

// binary search: largest r with offsets[r] <= q  (offsets ascending, offsets[0] = 0, q < offsets[n])
__device__ __forceinline__ int64_t find_rule(const int32_t* __restrict__ offsets, int64_t nrules, int64_t q)
{
  int64_t lo = 0, hi = nrules; // invariant: offsets[lo] <= q < offsets[hi]
  while (hi - lo > 1)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= q)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}
} // namespace cfx
