// assemble.cu -- K4 element kernels (standard + run-time quadrature cells, interior facets) and
// K5 deterministic CSR / vector gather.
//
// Replaces, for the hand-written kernel families of include/cutfemx_b200.h, the loops
// assemble_cells_matrix / assemble_interior_facets (cpp/dolfinx_custom_data/fem/
// assemble_matrix_impl.h:68-189, 409-607), assemble_cells (assemble_vector_impl.h:62-122),
// assemble_cells scalar (assemble_scalar_impl.h:26-59), the generated tabulate_tensor calls inside
// them (:141-143, :534-535) and mat_set = MatrixCSR::mat_add_values
// (python/cutfemx/wrappers/fem.cpp:340-385).
//
// Two stages:
//  (1) element stage: one thread per entity (cell or (facet, macro row)) computes the element
//      tensor and stores it in a per-form buffer indexed by the active-cell slot (all cell
//      integrals of a form accumulate into the same slot, in integral order);
//  (2) gather stage ("owner gathers", see sparsity.cu): one thread per matrix row walks the row's
//      cells in ascending order and adds the matching element-tensor row into its CSR row by
//      binary search of the sorted columns -- the same search MatrixCSR::mat_add_values does, but
//      with a fixed summation order and no atomics, so the result is bit-reproducible.
//
// Run-time rule convention: SURVEY.md facts 4-5 (reference points of the parent cell, physical
// weights, rule looked up by loop index).  Standard cells use the compile-time rule of the degree
// FFCx would pick (sum of argument degrees) scaled by |detJ|.
//
// Roofline: HBM for P1/P2 scalar forms (SURVEY.md section 8d "K4", "K5").
#include "common.cuh"
#include "element.cuh"

namespace cfx
{
namespace
{
constexpr int EB = 128;

struct RuleView
{
  const double* pts;  // SoA (tdim, npts)
  const double* wts;
  const double* nrm;  // SoA (gdim, npts) or null
  const int32_t* offsets;
  const int32_t* parent_map;
  int64_t npts;
};

struct StdRule
{
  const double* pts; // AoS (npts, tdim)
  const double* wts;
  int npts;
};

struct Consts
{
  double c[CFX_MAX_CONSTANTS];
};

template <int KID>
struct KernelTraits;
template <>
struct KernelTraits<CFX_K_LAPLACE>
{
  static constexpr int RANK = 2;
  static constexpr bool H = false, N = false;
};
template <>
struct KernelTraits<CFX_K_MASS>
{
  static constexpr int RANK = 2;
  static constexpr bool H = false, N = false;
};
template <>
struct KernelTraits<CFX_K_NITSCHE>
{
  static constexpr int RANK = 2;
  static constexpr bool H = true, N = true;
};
template <>
struct KernelTraits<CFX_K_SOURCE>
{
  static constexpr int RANK = 1;
  static constexpr bool H = false, N = false;
};
template <>
struct KernelTraits<CFX_K_NITSCHE_RHS>
{
  static constexpr int RANK = 1;
  static constexpr bool H = true, N = true;
};
template <>
struct KernelTraits<CFX_K_ONE>
{
  static constexpr int RANK = 0;
  static constexpr bool H = false, N = false;
};

template <int ND, int RANK>
struct ESize
{
  static constexpr int value = RANK == 2 ? ND * ND : (RANK == 1 ? ND : 1);
};

// contribution of one quadrature point (the body of the generated tabulate_tensor)
template <int TDIM, int DEG, int KID>
__device__ __forceinline__ void point_contribution(const Geo<TDIM>& g, const double (&xi)[TDIM], double w,
                                                   const double (&n)[TDIM], double h, const Consts& cs,
                                                   double (&acc)[ESize<Elem<TDIM, DEG>::ND, KernelTraits<KID>::RANK>::value])
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  if constexpr (KID == CFX_K_ONE)
  {
    acc[0] += cs.c[0] * w;
  }
  else
  {
  double phi[ND], dphi[ND][TDIM];
  tabulate<TDIM, DEG>(xi, phi, dphi);
  if constexpr (KID == CFX_K_MASS)
  {
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j)
        acc[i * ND + j] += cs.c[0] * w * phi[i] * phi[j];
  }
  else if constexpr (KID == CFX_K_SOURCE)
  {
#pragma unroll
    for (int i = 0; i < ND; ++i)
      acc[i] += cs.c[0] * w * phi[i];
  }
  else
  {
    double grad[ND][TDIM];
    push_gradients<TDIM, ND>(g, dphi, grad);
    if constexpr (KID == CFX_K_LAPLACE)
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          double s = 0.0;
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
            s += grad[i][r] * grad[j][r];
          acc[i * ND + j] += cs.c[0] * w * s;
        }
    }
    else
    {
      double gn[ND];
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          s += grad[i][r] * n[r];
        gn[i] = s;
      }
      if constexpr (KID == CFX_K_NITSCHE)
      { // demo_poisson.py:186-190
#pragma unroll
        for (int i = 0; i < ND; ++i)
#pragma unroll
          for (int j = 0; j < ND; ++j)
            acc[i * ND + j] += w * (-gn[j] * phi[i] - gn[i] * phi[j] + cs.c[0] / h * phi[i] * phi[j]);
      }
      else
      { // CFX_K_NITSCHE_RHS, demo_poisson.py:201 with constant boundary value c1
#pragma unroll
        for (int i = 0; i < ND; ++i)
          acc[i] += w * (-gn[i] * cs.c[1] + cs.c[0] / h * cs.c[1] * phi[i]);
      }
    }
  }
  }
}

// where a cell kernel puts its element tensor
struct OutCtx
{
  const int32_t* dofmap;     // argument-space dofmap (n_cells, ND)
  const int32_t* row_slot;   // dof -> index of its matrix row among the active rows
  const uint8_t* cell_inc_l; // (n_cells, ND): position of the cell in the incidence list of its i-th dof
  const int32_t* cell_slot;  // cell -> index among the active cells
  double* out;
  uint8_t* written;          // per active cell: an earlier integral already stored this cell's tensor
  int stride;                // incidence slots per matrix row
};

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// One thread per entity of a cell integral.  RUNTIME = false: entity e is standard cell
// cells[e] with the compile-time rule; true: entity e is rule e (parent cell parent_map[e]).
//
// Output layout ("owner-major"): row i of the element tensor goes straight to the storage of the
// matrix row it will be gathered into: slot (row_slot[dof_i] * stride + l) * ND, l = position of
// this cell among the cells around dof_i, columns in ascending global-dof order.  The gather then
// streams contiguous memory (the scattered 32-byte accesses happen here, as stores that the L2
// merges, instead of as dependent random loads in the gather).  For P1 tetrahedra one row is one
// 32-byte sector, written with a single 256-bit store.
template <int TDIM, int DEG, int KID, bool RUNTIME>
__global__ void __launch_bounds__(EB)
    cell_kernel(const int32_t* __restrict__ cells, int64_t n, RuleView rv, StdRule sr, Consts cs,
                const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, OutCtx oc)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int RANK = KernelTraits<KID>::RANK;
  constexpr int ES = ESize<ND, RANK>::value;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (e >= n)
    return;
  const int64_t cell = RUNTIME ? rv.parent_map[e] : cells[e];
  // issue the output-side gathers early: they are independent of the arithmetic
  const int64_t slot = oc.cell_slot[cell];
  int32_t d[ND];
  int64_t dest[ND];
  if constexpr (RANK >= 1)
  {
#pragma unroll
    for (int j = 0; j < ND; ++j)
      d[j] = __ldg(oc.dofmap + cell * ND + j);
#pragma unroll
    for (int j = 0; j < ND; ++j)
      dest[j] = static_cast<int64_t>(__ldg(oc.row_slot + d[j])) * oc.stride + __ldg(oc.cell_inc_l + cell * ND + j);
  }
  const bool add = oc.written[slot] != 0;
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  double h = 1.0;
  if constexpr (KernelTraits<KID>::H)
    h = cell_diameter<TDIM>(X);
  double acc[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i)
    acc[i] = 0.0;
  double nq[TDIM];
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
    nq[r] = 0.0;
  if constexpr (RUNTIME)
  {
    const int32_t q0 = rv.offsets[e], q1 = rv.offsets[e + 1];
    for (int32_t q = q0; q < q1; ++q)
    {
      double xi[TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        xi[t] = rv.pts[static_cast<int64_t>(t) * rv.npts + q];
      if constexpr (KernelTraits<KID>::N)
      {
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          nq[r] = rv.nrm[static_cast<int64_t>(r) * rv.npts + q];
      }
      point_contribution<TDIM, DEG, KID>(g, xi, rv.wts[q], nq, h, cs, acc);
    }
  }
  else
  {
    const double s = fabs(g.detJ);
    for (int q = 0; q < sr.npts; ++q)
    {
      double xi[TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        xi[t] = __ldg(sr.pts + q * TDIM + t);
      point_contribution<TDIM, DEG, KID>(g, xi, __ldg(sr.wts + q) * s, nq, h, cs, acc);
    }
  }
  if (!add)
    oc.written[slot] = 1;
  if constexpr (RANK == 2)
  {
    int rank[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      int rk = 0;
#pragma unroll
      for (int jj = 0; jj < ND; ++jj)
        rk += (d[jj] < d[j]) ? 1 : 0;
      rank[j] = rk;
    }
    if constexpr (ND == 4)
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        double rowv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
          double v = acc[i * 4 + 0];
#pragma unroll
          for (int j = 1; j < 4; ++j)
            v = (rank[j] == r) ? acc[i * 4 + j] : v;
          rowv[r] = v;
        }
        double* p = oc.out + dest[i] * 4;
        if (add)
        {
          double o0, o1, o2, o3;
          ld256(p, o0, o1, o2, o3);
          st256(p, o0 + rowv[0], o1 + rowv[1], o2 + rowv[2], o3 + rowv[3]);
        }
        else
          st256(p, rowv[0], rowv[1], rowv[2], rowv[3]);
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          double* p = oc.out + dest[i] * ND + rank[j];
          *p = add ? *p + acc[i * ND + j] : acc[i * ND + j];
        }
    }
  }
  else if constexpr (RANK == 1)
  {
#pragma unroll
    for (int i = 0; i < ND; ++i)
    {
      double* p = oc.out + dest[i];
      *p = add ? *p + acc[i] : acc[i];
    }
  }
  else
    oc.out[slot] = add ? oc.out[slot] + acc[0] : acc[0];
}

// Interior-facet ghost penalty, one thread per (facet, macro row).
// demo_poisson.py:191-199: c0 * avg(h) * inner(jump(grad(u), n), jump(grad(v), n)) * dS with
// jump(grad u, n) = grad u('+').n('+') + grad u('-').n('-'), n('-') = -n('+'); macro layout
// [[++,+-],[-+,--]] with '+' = first cell of the facet row (assemble_matrix_impl.h:537-542).
// Facet quadrature points are generated on the '+' side and pulled back to each cell's
// reference coordinates through its own affine map, so no quadrature_permutation is needed.
template <int TDIM, int DEG>
__global__ void __launch_bounds__(EB)
    facet_kernel(const int32_t* __restrict__ rows4, int64_t n_facets, StdRule fr, Consts cs,
                 const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, double* __restrict__ Fe,
                 bool accumulate)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int NV = TDIM + 1;
  constexpr int SD = TDIM - 1;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (t >= n_facets * 2 * ND)
    return;
  const int64_t f = t / (2 * ND);
  const int mrow = static_cast<int>(t - f * 2 * ND);
  const int32_t c0 = rows4[4 * f], lf0 = rows4[4 * f + 1], c1 = rows4[4 * f + 2];
  double X0[NV][TDIM], X1[NV][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, c0, X0);
  load_cell_coords<TDIM>(x, x_dofmap, c1, X1);
  Geo<TDIM> g0, g1;
  make_geo<TDIM>(X0, g0);
  make_geo<TDIM>(X1, g1);
  const double havg = 0.5 * (cell_diameter<TDIM>(X0) + cell_diameter<TDIM>(X1));
  // outward normal of cell 0 on local facet lf0: n = -K^T dlam_lf0 / |.|
  double nrm[TDIM], nn = 0.0;
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
  {
    double s = 0.0;
#pragma unroll
    for (int tt = 0; tt < TDIM; ++tt)
    {
      const double dl = lf0 == 0 ? -1.0 : (lf0 - 1 == tt ? 1.0 : 0.0);
      s += g0.K[tt * TDIM + r] * dl;
    }
    nrm[r] = -s;
    nn += s * s;
  }
  nn = sqrt(nn);
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
    nrm[r] /= nn;
  // facet vertices: the vertices of cell 0 other than lf0, ascending
  double Xf[TDIM][TDIM];
  {
    int k = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v != lf0)
      {
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          Xf[k < TDIM ? k : 0][r] = X0[v][r];
        ++k;
      }
  }
  double measure;
  if constexpr (TDIM == 2)
  {
    const double dx = Xf[1][0] - Xf[0][0], dy = Xf[1][1] - Xf[0][1];
    measure = sqrt(dx * dx + dy * dy); // rule weights sum to 1
  }
  else
  {
    const double u0 = Xf[1][0] - Xf[0][0], u1 = Xf[1][1] - Xf[0][1], u2 = Xf[1][2] - Xf[0][2];
    const double w0 = Xf[2][0] - Xf[0][0], w1 = Xf[2][1] - Xf[0][1], w2 = Xf[2][2] - Xf[0][2];
    const double cx = u1 * w2 - u2 * w1, cy = u2 * w0 - u0 * w2, cz = u0 * w1 - u1 * w0;
    measure = sqrt(cx * cx + cy * cy + cz * cz); // 2*area; rule weights sum to 1/2
  }
  double acc[2 * ND];
#pragma unroll
  for (int j = 0; j < 2 * ND; ++j)
    acc[j] = 0.0;
  for (int q = 0; q < fr.npts; ++q)
  {
    double lam[TDIM], l0 = 1.0;
#pragma unroll
    for (int k = 0; k < SD; ++k)
    {
      lam[k] = __ldg(fr.pts + q * SD + k);
      l0 -= lam[k];
    }
    double xq[TDIM];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double v = l0 * Xf[0][r];
#pragma unroll
      for (int k = 0; k < SD; ++k)
        v += lam[k] * Xf[k + 1][r];
      xq[r] = v;
    }
    double jn[2 * ND];
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
      const Geo<TDIM>& g = s ? g1 : g0;
      double Xr[TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
      {
        double v = 0.0;
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          v += g.K[tt * TDIM + r] * (xq[r] - g.x0[r]);
        Xr[tt] = v;
      }
      double phi[ND], dphi[ND][TDIM], grad[ND][TDIM];
      tabulate<TDIM, DEG>(Xr, phi, dphi);
      push_gradients<TDIM, ND>(g, dphi, grad);
      const double sign = s ? -1.0 : 1.0;
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double v = 0.0;
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          v += grad[i][r] * nrm[r];
        jn[s * ND + i] = sign * v;
      }
    }
    double ji = 0.0;
#pragma unroll
    for (int j = 0; j < 2 * ND; ++j)
      ji = (j == mrow) ? jn[j] : ji;
    const double w = __ldg(fr.wts + q) * measure * cs.c[0] * havg;
#pragma unroll
    for (int j = 0; j < 2 * ND; ++j)
      acc[j] += w * ji * jn[j];
  }
  double* o = Fe + t * 2 * ND;
#pragma unroll
  for (int j = 0; j < 2 * ND; ++j)
    o[j] = accumulate ? o[j] + acc[j] : acc[j];
}

// ------------------------------------------------------------------ K5 gather
struct GatherCtx
{
  const int64_t* inc_ptr;
  const int32_t* inc_cell;
  const int32_t* dofmap;
  const uint8_t* cell_flags;
  const uint8_t* row_flag;
  const int32_t* cell_slot;
  const double* Ae;
  const int32_t* c2f;
  const int32_t* facet_slot;
  const int32_t* rows4;
  const double* Fe;
  int nf;
  int stride; // incidence slots per matrix row in the owner-major element storage
};

constexpr int GW = 4; // rows (warps) per block

// rows no active entity touches: optional identity diagonal (deactivate_outside, deactivate.h:402-418)
__global__ void inactive_diag_kernel(const uint8_t* __restrict__ row_flag, int64_t n_rows,
                                     const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                     double* __restrict__ vals, double diag)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (r >= n_rows || row_flag[r])
    return;
  for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
    if (cols[p] == r)
      vals[p] = diag;
}

// One WARP per active row ("owner gathers").
//  phase 1: lanes take the row's incident cells: gather dofmap row, slot and the matching
//           element-tensor row (one 32 B sector for P1 tets) into shared memory -- up to 32
//           independent gather chains in flight per warp;
//  phase 2: lanes own the row's CSR entries (coalesced cols/vals access); every lane walks the
//           staged cells in ascending order (broadcast shared-memory reads) and adds the entry
//           whose dof equals its column.  Fixed order, no atomics -> bit-reproducible.
// Interior-facet macro tensors are added afterwards, cell by cell, facet by facet.
template <int ND>
__global__ void __launch_bounds__(GW * 32)
    gather_matrix_kernel(GatherCtx gc, const int32_t* __restrict__ act_rows, int64_t n_act,
                         const uint8_t* __restrict__ skip_fast, const int64_t* __restrict__ row_ptr,
                         const int32_t* __restrict__ cols, double* __restrict__ vals, int zero_first,
                         int32_t* __restrict__ err)
{
  __shared__ int32_t s_dofs[GW][32][ND];
  __shared__ double s_a[GW][32][ND];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * GW + w;
  if (idx >= n_act)
    return;
  if (skip_fast && (skip_fast[idx] & 1))
    return; // handled by gather_matrix_fast_kernel
  const unsigned full = 0xffffffffu;
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  int matched = 0, expected = 0; // entries found in the pattern vs entries contributed
  for (int kc = 0; kc < rn; kc += 32)
  {
    const bool have_col = kc + lane < rn;
    const int32_t mycol = have_col ? cols[rb + kc + lane] : -2;
    double acc = (have_col && !zero_first) ? vals[rb + kc + lane] : 0.0;
    for (int k0 = 0; k0 < n_inc; k0 += 32)
    {
      // ---- phase 1
      const int k = k0 + lane;
      int64_t c = -1;
      uint8_t fl = 0;
      int li = 0;
      if (k < n_inc)
      {
        c = gc.inc_cell[ib + k];
        fl = gc.cell_flags[c];
      }
      if (fl)
      {
        int32_t d[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          d[j] = gc.dofmap[c * ND + j];
          li = (d[j] == r) ? j : li;
        }
        if (fl & 1)
        {
          const double* a = gc.Ae + (idx * gc.stride + k) * ND; // owner-major: slot k of this row
          // element-tensor rows are stored in ascending-dof column order (cell_kernel)
#pragma unroll
          for (int pass = 0; pass < ND; ++pass)
#pragma unroll
            for (int i = pass & 1; i + 1 < ND; i += 2)
            {
              const int32_t lo = min(d[i], d[i + 1]), hi = max(d[i], d[i + 1]);
              d[i] = lo;
              d[i + 1] = hi;
            }
#pragma unroll
          for (int j = 0; j < ND; ++j)
          {
            s_dofs[w][lane][j] = d[j];
            s_a[w][lane][j] = a[j];
          }
        }
      }
      if (!(fl & 1))
      {
#pragma unroll
        for (int j = 0; j < ND; ++j)
          s_dofs[w][lane][j] = -1;
      }
      __syncwarp();
      // ---- phase 2
      const int nl = (n_inc - k0 < 32) ? n_inc - k0 : 32;
      for (int l = 0; l < nl; ++l)
      {
#pragma unroll
        for (int j = 0; j < ND; ++j)
          if (s_dofs[w][l][j] == mycol)
          {
            acc += s_a[w][l][j];
            ++matched;
          }
      }
      if (kc == 0)
        expected += (fl & 1) ? ND : 0;
      // ---- interior-facet macro rows of the band cells of this chunk
      unsigned band = __ballot_sync(full, (fl & 2) != 0);
      while (band)
      {
        const int l = __ffs(band) - 1;
        band &= band - 1;
        const int64_t cl = __shfl_sync(full, c, l);
        const int lil = __shfl_sync(full, li, l);
        for (int lf = 0; lf < gc.nf; ++lf)
        {
          const int64_t fs = gc.facet_slot[gc.c2f[cl * gc.nf + lf]];
          if (fs < 0)
            continue;
          const int64_t c0 = gc.rows4[4 * fs], c1 = gc.rows4[4 * fs + 2];
          const int mrow = (cl == c0 ? 0 : ND) + lil;
          const double* F = gc.Fe + (fs * 2 * ND + mrow) * 2 * ND;
#pragma unroll
          for (int s = 0; s < 2; ++s)
          {
            const int64_t cc = s ? c1 : c0;
#pragma unroll
            for (int j = 0; j < ND; ++j)
              if (gc.dofmap[cc * ND + j] == mycol)
              {
                acc += F[s * ND + j];
                ++matched;
              }
          }
          if (kc == 0 && lane == 0)
            expected += 2 * ND;
        }
      }
      __syncwarp();
    }
    if (have_col)
      vals[rb + kc + lane] = acc;
  }
  // MatrixCSR::mat_add_values throws when an entry is not in the pattern: every contributed
  // (cell, j) pair must have matched exactly one column of this row.
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    matched += __shfl_down_sync(full, matched, o);
    expected += __shfl_down_sync(full, expected, o);
  }
  if (lane == 0 && matched != expected)
  {
    err[0] = 31;
    err[1] = static_cast<int32_t>(r);
  }
}

// Fast rows (<= 32 columns, <= 32 incident cells): the pattern pass left, per incident cell l,
// gmask[idx*stride + l] = bit mask of the CSR positions of the cell's dofs, and the cell kernels
// stored the cell's tensor row at Ae[(idx*stride + l)*ND ..] in column order.  Lane k owns CSR entry
// k of the row and walks the incident cells in ascending order (masks broadcast by shuffle): cell l
// contributes iff bit k of its mask is set, and the value is entry popc(mask & lanes_below_k) of its
// row.  Everything a row reads is contiguous: stride masks, stride*ND doubles, its CSR segment.
// No shared memory, no dofmap read, no column search, fixed summation order.
template <int ND>
__global__ void __launch_bounds__(GW * 32)
    gather_matrix_fast_kernel(GatherCtx gc, const int32_t* __restrict__ act_rows, int64_t n_act,
                              const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ gmask,
                              const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                              double* __restrict__ vals, int zero_first)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * GW + w;
  if (idx >= n_act)
    return;
  const uint8_t rf = row_fast[idx];
  if (!(rf & 1))
    return; // handled by gather_matrix_kernel
  const unsigned full = 0xffffffffu;
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  const uint32_t gm = lane < n_inc ? gmask[idx * gc.stride + lane] : 0u;
  const bool have_col = lane < rn;
  double acc = (have_col && !zero_first) ? vals[rb + lane] : 0.0;
  const uint32_t below = (1u << lane) - 1u;
  const double* __restrict__ rowbase = gc.Ae + idx * gc.stride * ND;
  // batches of 8 cells: all (predicated) loads of a batch are issued before the first add; the adds
  // keep the ascending-cell order (x + 0.0 == x)
  for (int l0 = 0; l0 < n_inc; l0 += 8)
  {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int l = (l0 + u) & 31;
      const uint32_t M = __shfl_sync(full, gm, l);
      v[u] = ((M >> lane) & 1u) ? __ldg(rowbase + l * ND + __popc(M & below)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      acc += v[u];
  }
  if (rf & 2)
  { // interior-facet macro rows of the band cells (rare): column matched by value
    const int32_t mycol = have_col ? cols[rb + lane] : -2;
    int64_t c = -1;
    uint8_t fl = 0;
    int li = 0;
    if (lane < n_inc)
    {
      c = gc.inc_cell[ib + lane];
      fl = gc.cell_flags[c];
    }
    if (fl & 2)
    {
#pragma unroll
      for (int j = 0; j < ND; ++j)
        li = (gc.dofmap[c * ND + j] == r) ? j : li;
    }
    unsigned band = __ballot_sync(full, (fl & 2) != 0);
    while (band)
    {
      const int l = __ffs(band) - 1;
      band &= band - 1;
      const int64_t cl = __shfl_sync(full, c, l);
      const int lil = __shfl_sync(full, li, l);
      for (int lf = 0; lf < gc.nf; ++lf)
      {
        const int64_t fs = gc.facet_slot[gc.c2f[cl * gc.nf + lf]];
        if (fs < 0)
          continue;
        const int64_t c0 = gc.rows4[4 * fs], c1 = gc.rows4[4 * fs + 2];
        const int mrow = (cl == c0 ? 0 : ND) + lil;
        const double* F = gc.Fe + (fs * 2 * ND + mrow) * 2 * ND;
#pragma unroll
        for (int s = 0; s < 2; ++s)
        {
          const int64_t cc = s ? c1 : c0;
#pragma unroll
          for (int j = 0; j < ND; ++j)
            if (gc.dofmap[cc * ND + j] == mycol)
              acc += F[s * ND + j];
        }
      }
    }
  }
  if (have_col)
    vals[rb + lane] = acc;
}

// Four rows per warp (8 lanes each).  The owner-major vector storage holds one slot per incident
// cell (zero where the cell is inactive), so a row is one contiguous read; fixed-order partial sums
// + fixed shuffle tree -> bit-reproducible.
template <int ND>
__global__ void __launch_bounds__(GW * 32)
    gather_vector_kernel(GatherCtx gc, const int32_t* __restrict__ act_rows, int64_t n_act, double* __restrict__ b,
                         int zero_first)
{
  const int lane = threadIdx.x & 31;
  const int gl = lane & 7;
  const int64_t idx = (static_cast<int64_t>(blockIdx.x) * (GW * 32) + threadIdx.x) >> 3;
  const bool valid = idx < n_act;
  double s = 0.0;
  int64_t r = 0;
  if (valid)
  {
    r = act_rows[idx];
    const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - gc.inc_ptr[r]);
    const double* __restrict__ base = gc.Ae + idx * gc.stride;
#pragma unroll 4
    for (int k = gl; k < n_inc; k += 8)
      s += base[k];
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1)
    s += __shfl_down_sync(0xffffffffu, s, o, 8);
  if (valid && gl == 0)
    b[r] = zero_first ? s : b[r] + s;
}

// fixed-shape two-level sum: bit-reproducible
__global__ void __launch_bounds__(256) sum_partial_kernel(const double* __restrict__ v, int64_t n,
                                                          double* __restrict__ partial)
{
  __shared__ double s[256];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b = chunk * blockIdx.x;
  const int64_t e = b + chunk < n ? b + chunk : n;
  double acc = 0.0;
  for (int64_t i = b + threadIdx.x; i < e; i += 256)
    acc += v[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1)
  {
    if (threadIdx.x < o)
      s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    partial[blockIdx.x] = s[0];
}

// ------------------------------------------------------------------ host dispatch
template <int TDIM, int DEG, int KID>
void launch_cell(cfx_ctx* c, const cfx_integral& I, cfx_form* f)
{
  RuleView rv{};
  StdRule sr{};
  Consts cs;
  for (int k = 0; k < CFX_MAX_CONSTANTS; ++k)
    cs.c[k] = I.constants[k];
  const Space& S = c->spaces[f->space];
  OutCtx oc{S.dofmap, f->prep->row_slot.p, S.cell_inc_l.p, f->prep->cell_slot.p, f->Ae.p, f->written.p, S.stride};
  if (I.n > 0)
  {
    int order = 0;
    switch (KID)
    {
    case CFX_K_LAPLACE: order = 2 * (DEG - 1); break;
    case CFX_K_MASS: order = 2 * DEG; break;
    case CFX_K_SOURCE: order = DEG; break;
    default: order = 0;
    }
    RuleTable& rt = get_rule(c, TDIM, order);
    sr = StdRule{rt.d_pts, rt.d_wts, rt.npts};
    auto k = cell_kernel<TDIM, DEG, KID, false>;
    CFX_LAUNCH(c, k, grid_for(I.n, EB), EB, 0, I.entities, I.n, rv, sr, cs, c->x, c->x_dofmap, oc);
  }
  if (I.rules && I.rules->nrules > 0)
  {
    const cfx_rules* R = I.rules;
    CFX_REQUIRE(R->tdim == TDIM, CFX_ERR_INVALID, "run-time rules have the wrong reference dimension");
    rv = RuleView{R->points.p, R->weights.p, R->has_normals ? R->normals.p : nullptr, R->offsets.p, R->parent_map.p,
                  R->npts};
    auto k = cell_kernel<TDIM, DEG, KID, true>;
    CFX_LAUNCH(c, k, grid_for(R->nrules, EB), EB, 0, nullptr, R->nrules, rv, sr, cs, c->x, c->x_dofmap, oc);
  }
}

template <int TDIM, int DEG>
void dispatch_cell(cfx_ctx* c, const cfx_integral& I, cfx_form* f)
{
  switch (I.kernel)
  {
  case CFX_K_LAPLACE: launch_cell<TDIM, DEG, CFX_K_LAPLACE>(c, I, f); break;
  case CFX_K_MASS: launch_cell<TDIM, DEG, CFX_K_MASS>(c, I, f); break;
  case CFX_K_NITSCHE: launch_cell<TDIM, DEG, CFX_K_NITSCHE>(c, I, f); break;
  case CFX_K_SOURCE: launch_cell<TDIM, DEG, CFX_K_SOURCE>(c, I, f); break;
  case CFX_K_NITSCHE_RHS: launch_cell<TDIM, DEG, CFX_K_NITSCHE_RHS>(c, I, f); break;
  case CFX_K_ONE: launch_cell<TDIM, DEG, CFX_K_ONE>(c, I, f); break;
  default: throw Error(CFX_ERR_UNSUPPORTED, "unknown cell kernel family");
  }
}

void run_cell_integrals(cfx_ctx* c, cfx_form* f, int esize)
{
  const Space& S = c->spaces[f->space];
  // owner-major storage: (active rows, stride, ND) for matrices, (active rows, stride) for vectors
  const size_t n_out = f->rank == 0 ? static_cast<size_t>(f->prep->n_active)
                                    : static_cast<size_t>(f->prep->n_act_rows) * S.stride * (f->rank == 2 ? S.nd : 1);
  (void)esize;
  f->Ae.reserve(c->pool, n_out + 4);
  if (f->rank == 1) // the vector gather sums every slot of a row: slots of inactive cells must read 0
    CFX_CUDA(cudaMemsetAsync(f->Ae.p, 0, (n_out + 4) * sizeof(double), c->stream));
  f->written.reserve(c->pool, static_cast<size_t>(f->prep->n_active) + 1);
  CFX_CUDA(cudaMemsetAsync(f->written.p, 0, static_cast<size_t>(f->prep->n_active) + 1, c->stream));
  for (auto& I : f->integrals)
  {
    if (I.facet)
      continue;
    if (c->tdim == 2 && S.degree == 1)
      dispatch_cell<2, 1>(c, I, f);
    else if (c->tdim == 2)
      dispatch_cell<2, 2>(c, I, f);
    else if (S.degree == 1)
      dispatch_cell<3, 1>(c, I, f);
    else
      dispatch_cell<3, 2>(c, I, f);
  }
}

template <int TDIM, int DEG>
void launch_facet(cfx_ctx* c, const cfx_integral& I, cfx_form* f, bool accumulate)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  RuleTable& rt = get_rule(c, TDIM - 1, 2 * (DEG - 1));
  StdRule fr{rt.d_pts, rt.d_wts, rt.npts};
  Consts cs;
  for (int k = 0; k < CFX_MAX_CONSTANTS; ++k)
    cs.c[k] = I.constants[k];
  auto k = facet_kernel<TDIM, DEG>;
  CFX_LAUNCH(c, k, grid_for(I.n * 2 * ND, EB), EB, 0, I.entities, I.n, fr, cs, c->x, c->x_dofmap, f->Fe.p, accumulate);
}

GatherCtx make_gather_ctx(cfx_ctx* c, cfx_form* f, const cfx_integral* FI)
{
  const Space& S = c->spaces[f->space];
  return GatherCtx{S.inc_ptr.p, S.inc_cell.p,    S.dofmap,           f->prep->cell_flags.p,        f->prep->row_flag.p,
                   f->prep->cell_slot.p, f->Ae.p,       c->c2f,             c->facet_slot.p,        FI ? FI->entities : nullptr,
                   f->Fe.p,        c->tdim + 1, S.stride};
}
} // namespace
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_assemble_matrix(cfx_ctx* ctx, const cfx_form* a_const, cfx_pattern* A, int zero_first,
                               double diag_inactive, double* values_out, int memspace)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && a && A, CFX_ERR_INVALID, "cfx_assemble_matrix: NULL argument");
  CFX_REQUIRE(a->rank == 2, CFX_ERR_INVALID, "cfx_assemble_matrix: form is not bilinear");
  CFX_REQUIRE(A->space == a->space, CFX_ERR_INVALID, "cfx_assemble_matrix: matrix and form use different spaces");
  const Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  const cfx_integral* FI = facet_integral_domain(a);
  const int nd = S.nd;
  {
    StageScope st(ctx, "element_cells");
    run_cell_integrals(ctx, a, nd * nd);
    st.set_bytes(static_cast<double>(a->prep->n_active) * (4.0 * ctx->nv + 8.0 * nd * nd));
  }
  if (FI)
  {
    StageScope st(ctx, "element_facets", static_cast<double>(FI->n) * (16.0 + 8.0 * 4.0 * nd * nd));
    a->Fe.reserve(ctx->pool, static_cast<size_t>(FI->n) * 4 * nd * nd + 1);
    bool acc = false;
    for (auto& I : a->integrals)
    {
      if (!I.facet || I.n == 0)
        continue;
      if (ctx->tdim == 2 && S.degree == 1)
        launch_facet<2, 1>(ctx, I, a, acc);
      else if (ctx->tdim == 2)
        launch_facet<2, 2>(ctx, I, a, acc);
      else if (S.degree == 1)
        launch_facet<3, 1>(ctx, I, a, acc);
      else
        launch_facet<3, 2>(ctx, I, a, acc);
      acc = true;
    }
  }
  {
    StageScope st(ctx, "gather_matrix",
                  12.0 * static_cast<double>(A->nnz) + 8.0 * nd * nd * static_cast<double>(a->prep->n_active));
    set_facet_slots(ctx, FI, false);
    GatherCtx gc = make_gather_ctx(ctx, a, FI);
    if (zero_first)
      CFX_CUDA(cudaMemsetAsync(A->values.p, 0, static_cast<size_t>(A->nnz) * sizeof(double), ctx->stream));
    if (diag_inactive != 0.0)
      CFX_LAUNCH(ctx, inactive_diag_kernel, grid_for(A->n_rows, 256), 256, 0, a->prep->row_flag.p, A->n_rows, A->row_ptr.p,
                 A->cols.p, A->values.p, diag_inactive);
    auto k = nd == 3 ? gather_matrix_kernel<3>
             : nd == 4 ? gather_matrix_kernel<4>
             : nd == 6 ? gather_matrix_kernel<6>
                       : gather_matrix_kernel<10>;
    auto kf = nd == 3 ? gather_matrix_fast_kernel<3>
              : nd == 4 ? gather_matrix_fast_kernel<4>
              : nd == 6 ? gather_matrix_fast_kernel<6>
                        : gather_matrix_fast_kernel<10>;
    // the gather table is valid only for the pattern that was built from this very form
    const bool fast = a->gtab_serial == A->serial && a->gtab_serial > 0;
    if (a->prep->n_act_rows > 0)
    {
      const unsigned g = grid_for(a->prep->n_act_rows, GW);
      if (fast)
        CFX_LAUNCH(ctx, kf, g, GW * 32, 0, gc, a->prep->act_rows.p, a->prep->n_act_rows, a->row_fast.p, a->gmask.p, A->row_ptr.p,
                   A->cols.p, A->values.p, zero_first);
      if (!fast || a->n_slow_rows > 0)
        CFX_LAUNCH(ctx, k, g, GW * 32, 0, gc, a->prep->act_rows.p, a->prep->n_act_rows, fast ? a->row_fast.p : nullptr,
                   A->row_ptr.p, A->cols.p, A->values.p, zero_first, ctx->err_flag.p);
    }
    set_facet_slots(ctx, FI, true);
  }
  if (values_out)
    export_to(ctx, values_out, A->values.p, static_cast<size_t>(A->nnz), memspace);
  check_device_error(ctx, "cfx_assemble_matrix (entry not in sparsity pattern)");
  CFX_API_END(ctx)
}

cfx_status cfx_assemble_vector(cfx_ctx* ctx, const cfx_form* L_const, double* b, int zero_first, int memspace)
{
  CFX_API_BEGIN
  cfx_form* L = const_cast<cfx_form*>(L_const);
  CFX_REQUIRE(ctx && L && b, CFX_ERR_INVALID, "cfx_assemble_vector: NULL argument");
  CFX_REQUIRE(L->rank == 1, CFX_ERR_INVALID, "cfx_assemble_vector: form is not linear");
  const Space& S = ctx->spaces[L->space];
  prepare_form(ctx, L);
  {
    StageScope st(ctx, "element_cells_vector");
    run_cell_integrals(ctx, L, S.nd);
    st.set_bytes(static_cast<double>(L->prep->n_active) * (4.0 * ctx->nv + 8.0 * S.nd));
  }
  DevBuf<double> tmp;
  double* d_b = b;
  if (memspace == CFX_HOST)
  {
    tmp.reserve(ctx->pool, static_cast<size_t>(S.n_total));
    d_b = tmp.p;
    if (!zero_first)
      CFX_CUDA(cudaMemcpyAsync(d_b, b, static_cast<size_t>(S.n_total) * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  }
  {
    StageScope st(ctx, "gather_vector", 8.0 * static_cast<double>(S.n_total) + 8.0 * S.nd * L->prep->n_active);
    GatherCtx gc = make_gather_ctx(ctx, L, nullptr);
    auto k = S.nd == 3 ? gather_vector_kernel<3>
             : S.nd == 4 ? gather_vector_kernel<4>
             : S.nd == 6 ? gather_vector_kernel<6>
                         : gather_vector_kernel<10>;
    if (zero_first)
      CFX_CUDA(cudaMemsetAsync(d_b, 0, static_cast<size_t>(S.n_total) * sizeof(double), ctx->stream));
    if (L->prep->n_act_rows > 0)
      CFX_LAUNCH(ctx, k, grid_for(L->prep->n_act_rows * 8, GW * 32), GW * 32, 0, gc, L->prep->act_rows.p, L->prep->n_act_rows, d_b,
                 zero_first);
  }
  if (memspace == CFX_HOST)
  {
    export_to(ctx, b, tmp.p, static_cast<size_t>(S.n_total), CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}

cfx_status cfx_assemble_scalar(cfx_ctx* ctx, const cfx_form* M_const, double* out)
{
  CFX_API_BEGIN
  cfx_form* M = const_cast<cfx_form*>(M_const);
  CFX_REQUIRE(ctx && M && out, CFX_ERR_INVALID, "cfx_assemble_scalar: NULL argument");
  CFX_REQUIRE(M->rank == 0, CFX_ERR_INVALID, "cfx_assemble_scalar: form is not a functional");
  prepare_form(ctx, M);
  run_cell_integrals(ctx, M, 1);
  constexpr int NB = 256;
  DevBuf<double> partial;
  partial.reserve(ctx->pool, NB + 1);
  CFX_LAUNCH(ctx, sum_partial_kernel, NB, 256, 0, M->Ae.p, M->prep->n_active, partial.p);
  CFX_LAUNCH(ctx, sum_partial_kernel, 1, 256, 0, partial.p, static_cast<int64_t>(NB), partial.p + NB);
  CFX_CUDA(cudaMemcpyAsync(out, partial.p + NB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  partial.release();
  CFX_API_END(ctx)
}
} // extern "C"
