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
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <type_traits>
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
  const double* mom;  // per rule: measure, first (and, interface rules, second) moments, or null (cfx_rules::moments)
  int mom_stride;     // tdim + 1 (volume rules) or 1 + tdim + tdim (tdim + 1) / 2 (interface rules)
  const int64_t* d_sizes; // deferred-size mode: device [nrules, npts] (npts above is then only an upper bound)
  // stride of the SoA point / normal arrays = the exact number of points
  __device__ __forceinline__ int64_t stride() const { return d_sizes ? d_sizes[1] : npts; }
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

// entry i of a small register array without dynamic indexing
template <int N>
__device__ __forceinline__ double pick(const double (&a)[N], int i)
{
  double v = a[0];
#pragma unroll
  for (int j = 1; j < N; ++j)
    v = (j == i) ? a[j] : v;
  return v;
}

template <int N>
__device__ __forceinline__ int32_t pick_i(const int32_t (&a)[N], int i)
{
  int32_t v = a[0];
#pragma unroll
  for (int j = 1; j < N; ++j)
    v = (j == i) ? a[j] : v;
  return v;
}


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

template <>
struct KernelTraits<CFX_K_SQUARE_FN>
{
  static constexpr int RANK = 0;
  static constexpr bool H = false, N = false;
};
// kernels that read an ordinary Function coefficient (its nd cell-local dof values)
template <int KID>
struct KernelCoef
{
  static constexpr bool value = KID == CFX_K_SQUARE_FN;
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
  if constexpr (KID == CFX_K_ONE || KID == CFX_K_SQUARE_FN)
  {
    acc[0] += cs.c[0] * w; // CFX_K_SQUARE_FN: the caller has folded w_h(xi)^2 into the weight
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
  int32_t* mat_slot; // cell -> slot of its materialised tensor (-1: none yet); ranks 1 and 2
  double* out;
  int64_t base;      // first slot (rank 1, 2) / first entity index (rank 0) of this integral;
                     // < 0: FOLLOW mode -- use the slots another form of the same domains has claimed and add
                     // into a zero-initialised buffer (cfx_assemble_system: the linear form follows the bilinear one)
};

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// One thread per entity of a cell integral whose tensor is MATERIALISED: every run-time-rule entity
// (RUNTIME = true: entity e is rule e, parent cell parent_map[e], rule looked up by loop index --
// SURVEY.md fact 5), and the entities of rank-0 functionals.  Standard-quadrature cells of rank-1/2
// forms are NOT materialised: the owner of each matrix row computes their tensor rows on the fly
// (gather kernels below).
//
// Output: cell-major, natural dof order, ES doubles per slot.  The first integral that reaches a
// cell claims slot base + e and records it in mat_slot[cell]; later integrals of the same form
// (launched afterwards on the same stream) add into that slot.  Rules hold at most one entity per
// cell, so there is no race inside a launch.
template <int TDIM, int DEG, int KID, bool RUNTIME>
__global__ void __launch_bounds__(EB)
    cell_kernel(const int32_t* __restrict__ cells, DN n_, RuleView rv, StdRule sr, Consts cs,
                const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, OutCtx oc,
                const double* __restrict__ coeff, const int32_t* __restrict__ dofmap)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int RANK = KernelTraits<KID>::RANK;
  constexpr int ES = ESize<ND, RANK>::value;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (e >= n_.get())
    return;
  const int64_t cell = RUNTIME ? rv.parent_map[e] : cells[e];
  int64_t slot = oc.base + e;
  bool add = false;
  const bool follow = oc.base < 0;
  if constexpr (RANK >= 1)
  {
    const int32_t s0 = oc.mat_slot[cell];
    add = s0 >= 0;
    slot = add ? s0 : slot;
    if (follow && !add)
      return; // cannot happen for forms over the same prepared domains
  }
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
  // pack_coefficients (pack_form.h:98-131) fused into the kernel: the cell's nd dof values of the Function
  double wl[ND];
  if constexpr (KernelCoef<KID>::value)
  {
#pragma unroll
    for (int j = 0; j < ND; ++j)
      wl[j] = coeff[dofmap[cell * ND + j]];
  }
  // value of the coefficient at a reference point, squared (CFX_K_SQUARE_FN), folded into the weight
  auto coef_weight = [&](const double (&xi)[TDIM], double w) -> double
  {
    if constexpr (KernelCoef<KID>::value)
    {
      double phi[ND], dphi[ND][TDIM];
      tabulate<TDIM, DEG>(xi, phi, dphi);
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < ND; ++j)
        v += phi[j] * wl[j];
      return w * (v * v);
    }
    else
      return w;
  };
  // P1 integrands that are constant (Laplace, measure) or linear (source) in xi: the rule's measure W and its
  // first moments give the sum over the points exactly -- one evaluation at the rule's centroid with weight W
  // instead of a pass over the points (the generator computed W and the moments, quadrature.cu)
  constexpr bool LINEAR = DEG == 1 && (KID == CFX_K_LAPLACE || KID == CFX_K_SOURCE || KID == CFX_K_ONE);
  bool done = false;
  if constexpr (RUNTIME && LINEAR)
  {
    if (rv.mom != nullptr)
    {
      const double* mo = rv.mom + e * rv.mom_stride;
      const double W = mo[0];
      double xi[TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        xi[t] = W != 0.0 ? mo[1 + t] / W : 0.0;
      point_contribution<TDIM, DEG, KID>(g, xi, W, nq, h, cs, acc);
      done = true;
    }
  }
  // P1 Nitsche terms on an interface rule: the normal is constant on the cell (P1 level set), the integrand is
  // quadratic in xi -> measure, first and second moments give the point sums exactly:
  //   phi_i = a_i + b_i . xi (a_0 = 1, b_0 = -1, a_t = 0, b_t = e_t)
  //   S_i = sum w phi_i = a_i W + b_i . M1,   Q_ij = sum w phi_i phi_j = a_i a_j W + a_i b_j.M1 + a_j b_i.M1 + b_i^T M2 b_j
  constexpr bool QUAD = DEG == 1 && (KID == CFX_K_NITSCHE || KID == CFX_K_NITSCHE_RHS);
  if constexpr (RUNTIME && QUAD)
  {
    constexpr int NM2 = TDIM * (TDIM + 1) / 2;
    if (rv.mom != nullptr && rv.mom_stride == 1 + TDIM + NM2 && rv.offsets[e + 1] > rv.offsets[e])
    {
      const double* mo = rv.mom + e * rv.mom_stride;
      const double W = mo[0];
      double M1[TDIM], M2[TDIM][TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        M1[t] = mo[1 + t];
      {
        int k = 0;
#pragma unroll
        for (int a = 0; a < TDIM; ++a)
#pragma unroll
          for (int b = a; b < TDIM; ++b)
          {
            M2[a][b] = mo[1 + TDIM + k];
            M2[b][a] = M2[a][b];
            ++k;
          }
      }
      const int32_t qf = rv.offsets[e];
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
        nq[r] = rv.nrm[static_cast<int64_t>(r) * rv.stride() + qf];
      // constant gradients and their normal components
      double xi0[TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        xi0[t] = 0.0;
      double phi0[ND], dphi[ND][TDIM], grad[ND][TDIM], gn[ND];
      tabulate<TDIM, DEG>(xi0, phi0, dphi);
      push_gradients<TDIM, ND>(g, dphi, grad);
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double sgn = 0.0;
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          sgn += grad[i][r] * nq[r];
        gn[i] = sgn;
      }
      double Sv[ND];
      double m1sum = 0.0;
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        m1sum += M1[t];
      Sv[0] = W - m1sum;
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        Sv[t + 1] = M1[t];
      if constexpr (KID == CFX_K_NITSCHE)
      {
        // Q: rows/cols >= 1 are M2; row 0 follows from phi_0 = 1 - sum_t phi_t
        double Q[ND][ND];
#pragma unroll
        for (int a = 0; a < TDIM; ++a)
#pragma unroll
          for (int b = 0; b < TDIM; ++b)
            Q[a + 1][b + 1] = M2[a][b];
#pragma unroll
        for (int b = 0; b < TDIM; ++b)
        {
          double rs = 0.0;
#pragma unroll
          for (int a = 0; a < TDIM; ++a)
            rs += M2[a][b];
          Q[0][b + 1] = M1[b] - rs; // sum w (1 - sum_a xi_a) xi_b
          Q[b + 1][0] = Q[0][b + 1];
        }
        {
          double q00 = Sv[0];
#pragma unroll
          for (int b = 0; b < TDIM; ++b)
            q00 -= Q[0][b + 1]; // sum w phi_0 (1 - sum_b xi_b)
          Q[0][0] = q00;
        }
        const double pen = cs.c[0] / h;
#pragma unroll
        for (int i = 0; i < ND; ++i)
#pragma unroll
          for (int j = 0; j < ND; ++j)
            acc[i * ND + j] += -gn[j] * Sv[i] - gn[i] * Sv[j] + pen * Q[i][j];
      }
      else
      { // CFX_K_NITSCHE_RHS
#pragma unroll
        for (int i = 0; i < ND; ++i)
          acc[i] += -gn[i] * cs.c[1] * W + cs.c[0] / h * cs.c[1] * Sv[i];
      }
      done = true;
    }
  }
  if constexpr (RUNTIME)
  {
    const int32_t q0 = rv.offsets[e], q1 = done ? rv.offsets[e] : rv.offsets[e + 1];
    for (int32_t q = q0; q < q1; ++q)
    {
      double xi[TDIM];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        xi[t] = rv.pts[static_cast<int64_t>(t) * rv.stride() + q];
      if constexpr (KernelTraits<KID>::N)
      {
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          nq[r] = rv.nrm[static_cast<int64_t>(r) * rv.stride() + q];
      }
      point_contribution<TDIM, DEG, KID>(g, xi, coef_weight(xi, rv.wts[q]), nq, h, cs, acc);
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
      point_contribution<TDIM, DEG, KID>(g, xi, coef_weight(xi, __ldg(sr.wts + q) * s), nq, h, cs, acc);
    }
  }
  if constexpr (RANK >= 1)
  {
    if (!add && !follow)
      oc.mat_slot[cell] = static_cast<int32_t>(slot);
  }
  double* p = oc.out + slot * ES;
  if constexpr (ES % 4 == 0)
  {
#pragma unroll
    for (int i = 0; i < ES; i += 4)
    {
      if (add)
      {
        double o0, o1, o2, o3;
        ld256(p + i, o0, o1, o2, o3);
        st256(p + i, o0 + acc[i], o1 + acc[i + 1], o2 + acc[i + 2], o3 + acc[i + 3]);
      }
      else
        st256(p + i, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
  }
  else
  {
#pragma unroll
    for (int i = 0; i < ES; ++i)
      p[i] = add ? p[i] + acc[i] : acc[i];
  }
}

__global__ void reset_slots_kernel(const int32_t* __restrict__ cells, DN n_, int32_t* __restrict__ mat_slot)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n_.get())
    mat_slot[cells[i]] = -1;
}

// Interior-facet ghost penalty, one thread per (facet, macro row).
// demo_poisson.py:191-199: c0 * avg(h) * inner(jump(grad(u), n), jump(grad(v), n)) * dS with
// jump(grad u, n) = grad u('+').n('+') + grad u('-').n('-'), n('-') = -n('+'); macro layout
// [[++,+-],[-+,--]] with '+' = first cell of the facet row (assemble_matrix_impl.h:537-542).
// Facet quadrature points are generated on the '+' side and pulled back to each cell's
// reference coordinates through its own affine map, so no quadrature_permutation is needed.
template <int TDIM, int DEG>
__global__ void __launch_bounds__(EB)
    facet_kernel(const int32_t* __restrict__ rows4, DN n_facets_, StdRule fr, Consts cs,
                 const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, const double* __restrict__ geo,
                 double* __restrict__ Fe, bool accumulate)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int NV = TDIM + 1;
  constexpr int SD = TDIM - 1;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (t >= n_facets_.get() * 2 * ND)
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

// P1 ghost penalty, one thread per facet, FACTORED and COMBINED output.  The normal-gradient jump is constant
// on the facet, so the macro tensor is the rank-one matrix w * jn (x) jn with jn the 2 nd jump coefficients and
// w = c0 * avg(h) * |F|.  The two cells share nd - 1 dofs: adding the two coefficients of a shared dof gives one
// coefficient J per DISTINCT dof (nd + 1 of them), and the assembled contribution of the facet to entry (r, s)
// is w * J_r * J_s.  The record a row owner reads is therefore 80 B -- the distinct dofs, J and w -- instead of
// two dofmap rows, 2 nd coefficients and the weight, and a row meets every facet once instead of once per cell:
//   int32 d[5]  : the distinct dofs, ASCENDING (triangles use d[0..3], d[4] = INT_MAX)
//   int32 c0    : first cell of the facet row (the lane of c0 owns the facet for rows both cells touch)
//   int32 i0 | i1 << 4 : where the dof opposite the facet in cell 0 / in cell 1 sits in d (P1: local facet lf is
//                 opposite local vertex lf);  int32 : the dof opposite the facet in cell 1
//   double J[5] : combined jump coefficients in the order of d;  double w
// Ascending order is what lets a band row add a facet through a bit mask of CSR positions (add_facet_rows_p1_mask).
// Everything follows from the two cached geometry records: grad lam_j = rows of K, n = -grad lam_lf0 / |.|,
// |F| (tdim-1)! = |detJ| |grad lam_lf0|.
constexpr int FREC = 10; // record stride in doubles (80 B)

template <int TDIM>
__global__ void __launch_bounds__(EB, 6)
    facet_p1_kernel(const int32_t* __restrict__ rows4, DN n_facets_, Consts cs, const double* __restrict__ geo,
                    const int32_t* __restrict__ dofmap, double* __restrict__ Frec, bool accumulate)
{
  constexpr int ND = TDIM + 1;
  const int64_t f = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (f >= n_facets_.get())
    return;
  const int32_t c0 = rows4[4 * f], lf0 = rows4[4 * f + 1], c1 = rows4[4 * f + 2], lf1 = rows4[4 * f + 3];
  Geo<TDIM> g[2];
  load_geo_cached<TDIM>(geo, c0, g[0]);
  load_geo_cached<TDIM>(geo, c1, g[1]);
  const double havg = 0.5 * (__ldg(geo + static_cast<int64_t>(c0) * GeoRec<TDIM>::STRIDE + TDIM * TDIM + 1)
                             + __ldg(geo + static_cast<int64_t>(c1) * GeoRec<TDIM>::STRIDE + TDIM * TDIM + 1));
  double G[2][ND][TDIM];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double s0 = 0.0;
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
      {
        G[s][tt + 1][r] = g[s].K[tt * TDIM + r];
        s0 -= g[s].K[tt * TDIM + r];
      }
      G[s][0][r] = s0;
    }
  double nrm[TDIM], nn = 0.0;
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
  {
    double v = G[0][0][r];
#pragma unroll
    for (int j = 1; j < ND; ++j)
      v = (j == lf0) ? G[0][j][r] : v;
    nrm[r] = -v;
    nn += v * v;
  }
  nn = sqrt(nn);
  const double measure = fabs(g[0].detJ) * nn;
  double jn[2][ND];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int i = 0; i < ND; ++i)
    {
      double v = 0.0;
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
        v += G[s][i][r] * (nrm[r] / nn);
      jn[s][i] = s ? -v : v;
    }
  int32_t d0[ND], d1[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j)
  {
    d0[j] = dofmap[static_cast<int64_t>(c0) * ND + j];
    d1[j] = dofmap[static_cast<int64_t>(c1) * ND + j];
  }
  int32_t d[5] = {-3, -3, -3, -3, -3};
  double J[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int k = 0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
  {
    if (j == lf0)
      continue;
    double other = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i)
      other = (d1[i] == d0[j]) ? jn[1][i] : other;
    // k runs 0 .. nd-2 over the shared dofs (compile-time unrolled: the index stays in registers)
#pragma unroll
    for (int q = 0; q < ND - 1; ++q)
      if (q == k)
      {
        d[q] = d0[j];
        J[q] = jn[0][j] + other;
      }
    ++k;
  }
  d[ND - 1] = pick_i<ND>(d0, lf0);
  J[ND - 1] = pick<ND>(jn[0], lf0);
  d[ND] = pick_i<ND>(d1, lf1);
  J[ND] = pick<ND>(jn[1], lf1);
  const double w = (TDIM == 3 ? 0.5 : 1.0) * measure * cs.c[0] * havg;
  // ascending-dof order (rank = number of smaller entries; the triangle's unused fifth entry sorts last)
  if (ND == 3)
    d[4] = 0x7fffffff;
  int32_t ds[5] = {0, 0, 0, 0, 0};
  double Js[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int i0 = 0, i1 = 0;
#pragma unroll
  for (int q = 0; q < 5; ++q)
  {
    int rank = 0;
#pragma unroll
    for (int p = 0; p < 5; ++p)
      rank += (d[p] < d[q]) ? 1 : 0;
#pragma unroll
    for (int t = 0; t < 5; ++t)
      if (t == rank)
      {
        ds[t] = d[q];
        Js[t] = J[q];
      }
    i0 = (q == ND - 1) ? rank : i0;
    i1 = (q == ND) ? rank : i1;
  }
  double* o = Frec + f * FREC;
  int32_t* oi = reinterpret_cast<int32_t*>(o);
  const double w_old = accumulate ? o[9] : 0.0;
  *reinterpret_cast<int4*>(oi) = make_int4(ds[0], ds[1], ds[2], ds[3]);
  *reinterpret_cast<int4*>(oi + 4) = make_int4(ds[4], c0, i0 | (i1 << 4), d[ND]);
#pragma unroll
  for (int q = 0; q < 5; ++q)
    o[4 + q] = Js[q];
  o[9] = w_old + w;
}

// ------------------------------------------------------------------ K5 gather ("owner gathers")
// standard-quadrature cell integrals of a form, evaluated on the fly by the row owners
struct StdTab
{
  int n;
  int kernel[CFX_MAX_STD_LISTS];
  unsigned bit[CFX_MAX_STD_LISTS]; // cell_flags bit of the integral's cell list
  double c[CFX_MAX_STD_LISTS][2];
  const double* pts[CFX_MAX_STD_LISTS]; // AoS (npts, tdim) reference rule, weights sum to 1/tdim!
  const double* wts[CFX_MAX_STD_LISTS];
  int npts[CFX_MAX_STD_LISTS];
  // P1 closed forms: the summed coefficient of the form's standard integrals a cell belongs to, indexed by
  // the cell's list bits (cell_flags >> 2).  t0 = Laplace (rank 2) / source (rank 1), t1 = mass.
  double t0[1 << CFX_MAX_STD_LISTS];
  double t1[1 << CFX_MAX_STD_LISTS];
  int has_mass;
  // P2 closed forms: reference-element tables (RefTab layout below), device memory
  const double* ref;
};

// Reference-element integrals of a Lagrange space on the affine simplex, computed once per (tdim, degree) on the
// device with the library's own tabulate and an exact rule:
//   R[a][b][i][j] = int d_a phi_i d_b phi_j,  M[i][j] = int phi_i phi_j,  S[i] = int phi_i   (reference cell)
// so that on a cell with K = J^-1 and C = K K^T (grad phi = K^T grad_ref phi):
//   Laplace_ij = |detJ| sum_ab C_ab R[a][b][i][j],  mass_ij = |detJ| M[i][j],  source_i = |detJ| S[i].
template <int TDIM, int ND>
struct RefTab
{
  static constexpr int R = 0, M = TDIM * TDIM * ND * ND, S = M + ND * ND, SIZE = S + ND;
};

template <int TDIM, int DEG>
__global__ void ref_tables_kernel(const double* __restrict__ pts, const double* __restrict__ wts, int npts,
                                  double* __restrict__ out)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  using T = RefTab<TDIM, ND>;
  const int t = threadIdx.x;
  if (t >= ND * ND)
    return;
  const int i = t / ND, j = t % ND;
  double r[TDIM][TDIM], m = 0.0, sv = 0.0;
#pragma unroll
  for (int a = 0; a < TDIM; ++a)
#pragma unroll
    for (int b = 0; b < TDIM; ++b)
      r[a][b] = 0.0;
  for (int q = 0; q < npts; ++q)
  {
    double xi[TDIM];
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
      xi[a] = pts[q * TDIM + a];
    double phi[ND], dphi[ND][TDIM];
    tabulate<TDIM, DEG>(xi, phi, dphi);
    const double w = wts[q];
    double pi = 0.0, pj = 0.0, di[TDIM], dj[TDIM];
#pragma unroll
    for (int k = 0; k < ND; ++k)
    {
      pi = (k == i) ? phi[k] : pi;
      pj = (k == j) ? phi[k] : pj;
#pragma unroll
      for (int a = 0; a < TDIM; ++a)
      {
        di[a] = (k == i) ? dphi[k][a] : (k == 0 ? 0.0 : di[a]);
        dj[a] = (k == j) ? dphi[k][a] : (k == 0 ? 0.0 : dj[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
#pragma unroll
      for (int b = 0; b < TDIM; ++b)
        r[a][b] += w * di[a] * dj[b];
    m += w * pi * pj;
    sv += w * pi;
  }
#pragma unroll
  for (int a = 0; a < TDIM; ++a)
#pragma unroll
    for (int b = 0; b < TDIM; ++b)
      out[T::R + ((a * TDIM + b) * ND + i) * ND + j] = r[a][b];
  out[T::M + i * ND + j] = m;
  if (j == 0)
    out[T::S + i] = sv;
}

struct GatherCtx
{
  const int64_t* inc_ptr;
  const int32_t* inc_cell;
  const uint32_t* fperm;
  const uint32_t* fmask;
  const int32_t* dofmap;
  const uint8_t* cell_flags;
  const int32_t* mat_slot;
  const double* Ae;
  const double* geo; // static per-cell geometry records (element.cuh GeoRec)
  const double* lrow; // scalar P1: static Laplace tensor rows, 4 doubles per incidence (Space::lrow); else null
  const uint32_t* fpos; // scalar P1 with a static structure: packed static row positions per incidence (Space::fpos)
  const uint64_t* fpos64; // scalar P2 on triangles with a static structure: the same, 64-bit (Space::fpos64)
  const int64_t* frow_ptr;
  const uint64_t* fclist;
  const int32_t* c2f;
  const int32_t* facet_slot;
  const int32_t* rows4;
  const double* Fe; // full macro tensors (n_facets, 2nd, 2nd), or the P1 records of facet_p1_kernel (n_facets, FREC)
  const double* Fw; // non-null: Fe holds P1 records; null: full tensors
  // fused system assembly (cfx_assemble_system): the linear form's materialised entries (same slots) and
  // the vector the contribution-list kernel also fills; null otherwise
  const double* AeL;
  double* bvec;
  int zero_first_b;
  int nf;
  int stride;
};

// Row `li` of the element tensor of standard cell `c` (sum over the form's standard integrals the
// cell belongs to) plus row `li` of its materialised run-time tensor, natural dof order.
template <int TDIM, int DEG>
__device__ __forceinline__ void std_row_values(const StdTab& st, const Geo<TDIM>& g, unsigned fl, int li,
                                               double (&v)[Elem<TDIM, DEG>::ND])
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  if constexpr (DEG == 1)
  { // P1: no quadrature loop and no loop over the integrals -- one table lookup gives the summed coefficient.
    // Laplace: constant gradients grad lam_0 = -(sum of the rows of K), grad lam_j = row j-1 of K, |T| = |detJ|/tdim!
    const double s = fabs(g.detJ);
    const unsigned m = fl >> 2;
    double G[ND][TDIM];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double s0 = 0.0;
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
      {
        G[t + 1][r] = g.K[t * TDIM + r];
        s0 -= g.K[t * TDIM + r];
      }
      G[0][r] = s0;
    }
    const double w = st.t0[m] * s * (TDIM == 3 ? 1.0 / 6.0 : 0.5);
    double gi[TDIM];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double t = G[0][r];
#pragma unroll
      for (int j = 1; j < ND; ++j)
        t = (j == li) ? G[j][r] : t;
      gi[r] = t * w;
    }
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      double d = 0.0;
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
        d += gi[r] * G[j][r];
      v[j] += d;
    }
    if (st.has_mass)
    { // int lam_i lam_j = |T| (1 + delta_ij) / ((tdim+1)(tdim+2))
      const double wm = st.t1[m] * s * (TDIM == 3 ? 1.0 / 120.0 : 1.0 / 24.0);
#pragma unroll
      for (int j = 0; j < ND; ++j)
        v[j] += (j == li) ? 2.0 * wm : wm;
    }
    return;
  }
  if (st.ref != nullptr)
  { // P2 (any tabulated degree): reference-element tables, no quadrature loop, no loop over the integrals
    using T = RefTab<TDIM, ND>;
    const double s = fabs(g.detJ);
    const unsigned m = fl >> 2;
    const double wl = st.t0[m] * s;
    const double* Rr = st.ref + T::R + li * ND;
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
#pragma unroll
      for (int b = 0; b < TDIM; ++b)
      {
        double cab = 0.0; // C = K K^T
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          cab += g.K[a * TDIM + t] * g.K[b * TDIM + t];
        cab *= wl;
#pragma unroll
        for (int j = 0; j < ND; ++j)
          v[j] += cab * Rr[(a * TDIM + b) * ND * ND + j];
      }
    if (st.has_mass)
    {
      const double wm = st.t1[m] * s;
#pragma unroll
      for (int j = 0; j < ND; ++j)
        v[j] += wm * st.ref[T::M + li * ND + j];
    }
    return;
  }
  {
    const double s = fabs(g.detJ);
    for (int k = 0; k < st.n; ++k)
    {
      if (!(fl & st.bit[k]))
        continue;
      const double c0 = st.c[k][0];
      const bool mass = st.kernel[k] == CFX_K_MASS;
      if (DEG == 1 && !mass)
      { // P1 Laplace: constant gradients grad lam_0 = -(sum of the rows of K), grad lam_j = row j-1 of K;
        // the degree-0 rule's weights sum to 1/tdim!
        double G[ND][TDIM];
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
        {
          double s0 = 0.0;
#pragma unroll
          for (int t = 0; t < TDIM; ++t)
          {
            G[t + 1][r] = g.K[t * TDIM + r];
            s0 -= g.K[t * TDIM + r];
          }
          G[0][r] = s0;
        }
        const double w = c0 * s * (TDIM == 3 ? 1.0 / 6.0 : 0.5);
        double gi[TDIM];
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
        {
          double t = G[0][r];
#pragma unroll
          for (int j = 1; j < ND; ++j)
            t = (j == li) ? G[j][r] : t;
          gi[r] = t * w;
        }
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          double d = 0.0;
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
            d += gi[r] * G[j][r];
          v[j] += d;
        }
        continue;
      }
      for (int q = 0; q < st.npts[k]; ++q)
      {
        double xi[TDIM];
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          xi[t] = __ldg(st.pts[k] + q * TDIM + t);
        const double w = __ldg(st.wts[k] + q) * s * c0;
        double phi[ND], dphi[ND][TDIM];
        tabulate<TDIM, DEG>(xi, phi, dphi);
        if (mass)
        {
          const double pi = pick<ND>(phi, li) * w;
#pragma unroll
          for (int j = 0; j < ND; ++j)
            v[j] += pi * phi[j];
        }
        else
        { // CFX_K_LAPLACE
          double grad[ND][TDIM];
          push_gradients<TDIM, ND>(g, dphi, grad);
          double gi[TDIM];
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
          {
            double t = grad[0][r];
#pragma unroll
            for (int j = 1; j < ND; ++j)
              t = (j == li) ? grad[j][r] : t;
            gi[r] = t * w;
          }
#pragma unroll
          for (int j = 0; j < ND; ++j)
          {
            double d = 0.0;
#pragma unroll
            for (int r = 0; r < TDIM; ++r)
              d += gi[r] * grad[j][r];
            v[j] += d;
          }
        }
      }
    }
  }
}

// P1: the 32-byte record of (cell, li) -- off-diagonal Laplace entries (ascending j != li) and |detJ|
struct P1Rec
{
  double a[4];
};
__device__ __forceinline__ P1Rec load_p1rec(const double* __restrict__ lrow, int64_t pinc)
{
  P1Rec r;
  const double* p = lrow + pinc * 4;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a[0]), "=d"(r.a[1]), "=d"(r.a[2]), "=d"(r.a[3]) : "l"(p));
  return r;
}
// row li of the summed standard-cell tensor from the record (same table lookup of the coefficients as
// std_row_values): Laplace off-diagonals scaled, diagonal = -(their sum), mass from |detJ|
template <int TDIM>
__device__ __forceinline__ void p1_row_from_rec(const StdTab& st, const P1Rec& rec, unsigned fl, int li,
                                                double (&v)[TDIM + 1])
{
  constexpr int ND = TDIM + 1;
  const unsigned m = fl >> 2;
  const double w = st.t0[m];
  double dsum = 0.0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
  {
    // slot of column j among the off-diagonals: j for j < li, j - 1 for j > li
    double o = 0.0;
#pragma unroll
    for (int q = 0; q < ND - 1; ++q)
      o = ((q < li ? q : q + 1) == j) ? rec.a[q] : o;
    const double t = w * o;
    dsum += (j == li) ? 0.0 : t;
    v[j] += (j == li) ? 0.0 : t;
  }
#pragma unroll
  for (int j = 0; j < ND; ++j)
    v[j] -= (j == li) ? dsum : 0.0;
  if (st.has_mass)
  { // int lam_i lam_j = |T| (1 + delta_ij) / ((tdim+1)(tdim+2))
    const double wm = st.t1[m] * rec.a[3] * (TDIM == 3 ? 1.0 / 120.0 : 1.0 / 24.0);
#pragma unroll
    for (int j = 0; j < ND; ++j)
      v[j] += (j == li) ? 2.0 * wm : wm;
  }
}

// pinc: position of (row, cell) in the space's incidence arrays (addresses the P1 tensor-row record)
template <int TDIM, int DEG>
__device__ __forceinline__ void cell_row_values(const GatherCtx& gc, const StdTab& st, int64_t c, unsigned fl, int li,
                                                int64_t pinc, double (&v)[Elem<TDIM, DEG>::ND])
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
#pragma unroll
  for (int j = 0; j < ND; ++j)
    v[j] = 0.0;
  if (fl >> 2)
  {
    if constexpr (DEG == 1)
    {
      if (gc.lrow != nullptr)
      {
        const P1Rec rec = load_p1rec(gc.lrow, pinc);
        p1_row_from_rec<TDIM>(st, rec, fl, li, v);
      }
      else
      {
        Geo<TDIM> g;
        load_geo_cached<TDIM>(gc.geo, c, g);
        std_row_values<TDIM, DEG>(st, g, fl, li, v);
      }
    }
    else
    {
      Geo<TDIM> g;
      load_geo_cached<TDIM>(gc.geo, c, g);
      std_row_values<TDIM, DEG>(st, g, fl, li, v);
    }
  }
  if (fl & 1)
  {
    const double* a = gc.Ae + (static_cast<int64_t>(__ldg(gc.mat_slot + c)) * ND + li) * ND;
#pragma unroll
    for (int j = 0; j < ND; ++j)
      v[j] += a[j];
  }
}

// rank 1: entry `li` of the element vector of cell c
template <int TDIM, int DEG>
__device__ __forceinline__ double cell_entry_value(const GatherCtx& gc, const StdTab& st, int64_t c, unsigned fl,
                                                   int li)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  double e = 0.0;
  if (fl >> 2)
  {
    const double s = fabs(__ldg(gc.geo + c * GeoRec<TDIM>::STRIDE + TDIM * TDIM));
    if constexpr (DEG == 1) // P1 source: int phi_i = |detJ| / (tdim+1)!, summed coefficient from the table
      e += st.t0[fl >> 2] * s * (TDIM == 3 ? 1.0 / 24.0 : 1.0 / 6.0);
    else if (st.ref != nullptr) // int phi_i = |detJ| S[i]
      e += st.t0[fl >> 2] * s * st.ref[RefTab<TDIM, ND>::S + li];
    else
    for (int k = 0; k < st.n; ++k)
    {
      if (!(fl & st.bit[k]))
        continue;
      for (int q = 0; q < st.npts[k]; ++q)
      { // CFX_K_SOURCE
        double xi[TDIM];
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          xi[t] = __ldg(st.pts[k] + q * TDIM + t);
        double phi[ND], dphi[ND][TDIM];
        tabulate<TDIM, DEG>(xi, phi, dphi);
        e += st.c[k][0] * (__ldg(st.wts[k] + q) * s) * pick<ND>(phi, li);
      }
    }
  }
  if (fl & 1)
    e += gc.Ae[static_cast<int64_t>(__ldg(gc.mat_slot + c)) * ND + li];
  return e;
}

constexpr int GW = 4;  // rows (warps) per block
constexpr int GWC = 4; // rows per block of the contribution-list kernel
constexpr int GWM = 4; // rows per block of the mask kernel

// rows no active entity touches: optional identity diagonal (deactivate_outside, deactivate.h:402-418)
__global__ void inactive_diag_kernel(const uint8_t* __restrict__ row_flag, int64_t n_rows,
                                     const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                     double* __restrict__ vals, double diag, int bs, int64_t cap)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (r >= n_rows || row_flag[r] || row_ptr[r + 1] > cap)
    return;
  for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
    if (cols[p] == r)
      for (int k = 0; k < bs; ++k) // blocked matrices: the diagonal of the diagonal block
        vals[p * bs * bs + k * bs + k] = diag;
}

// Right-hand-side entries of a row are summed in ONE order by every kernel that produces them, so that a fused
// system assembly (matrix gather kernels fill b) and a separate vector assembly agree bit for bit: the incident
// cells in groups of four, pairwise inside a group ((e0 + e1) + (e2 + e3)), the groups added in ascending order.
// That order costs a one-thread-per-row kernel nothing (its loop walks four incidences at a time) and a
// warp-per-row kernel two butterfly steps and seven broadcasts.  Returns the sum in every lane.
__device__ __forceinline__ double rhs_sum_groups_of_four(double e)
{
  const unsigned full = 0xffffffffu;
  e += __shfl_xor_sync(full, e, 1);
  e += __shfl_xor_sync(full, e, 2);
  double s = __shfl_sync(full, e, 0);
#pragma unroll
  for (int k = 1; k < 8; ++k)
    s += __shfl_sync(full, e, 4 * k);
  return s;
}

// Interior-facet macro rows of the band cells among this warp's (<= 32) incident cells.  Lane l owns
// incident cell l and probes its own local facets (up to 32 independent gather chains in flight);
// hits are staged in shared memory (the two cells' dofs + the macro-tensor row) and the column lanes
// add them facet round by facet round, cell by cell -- a fixed order.  Returns the number of matches.
// staged macro rows are padded to a multiple of 4 entries so that the column lanes can read them with
// 128-bit broadcast loads (6 shared-memory instructions per P1-tetrahedron hit instead of 16)
template <int ND>
struct FacetStage
{
  static constexpr int W = (2 * ND + 3) / 4 * 4;
};

// P1 (factored, combined records of facet_p1_kernel): the row's contribution from facet F is w J_r J_s over the
// nd + 1 distinct dofs s.  Lane l probes the facets of its band cell; a facet whose two cells both hold the row's
// dof is taken by the lane of its first cell only.  `r` is the row's dof.
template <int ND>
__device__ __forceinline__ int add_facet_rows_p1(const GatherCtx& gc, int32_t (*s_fd)[FacetStage<ND>::W],
                                                 double (*s_fv)[FacetStage<ND>::W], bool band_cell, int64_t c,
                                                 int32_t r, int32_t mycol, double& acc, int& expected, bool count)
{
  static_assert(FacetStage<ND>::W >= 6 || ND == 3, "staging rows hold the nd + 1 combined entries");
  constexpr int NE = ND + 1;                    // distinct dofs of the macro element
  constexpr int NQ = (NE + 1) / 2;              // 128-bit value reads per staged row (entries padded to even)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int matched = 0;
  if (__ballot_sync(full, band_cell) == 0)
    return 0;
  int32_t fct[4] = {0, 0, 0, 0};
  if (band_cell)
  {
    if (gc.nf == 4)
    {
      const int4 q = __ldg(reinterpret_cast<const int4*>(gc.c2f) + c);
      fct[0] = q.x;
      fct[1] = q.y;
      fct[2] = q.z;
      fct[3] = q.w;
    }
    else
      for (int lf = 0; lf < gc.nf; ++lf)
        fct[lf] = gc.c2f[c * gc.nf + lf];
  }
  // all slot probes first: independent gathers in flight together
  int32_t fsl[4] = {-1, -1, -1, -1};
  if (band_cell)
    for (int lf = 0; lf < gc.nf; ++lf)
      fsl[lf] = gc.facet_slot[fct[lf]];
  for (int lf = 0; lf < gc.nf; ++lf)
  {
    const int64_t fs = fsl[lf];
    bool valid = fs >= 0;
    if (valid)
    {
      const double* rec = gc.Fe + fs * FREC;
      const int4 q0 = __ldg(reinterpret_cast<const int4*>(rec));
      const int4 q1 = __ldg(reinterpret_cast<const int4*>(rec) + 1);
      const int32_t dd[5] = {q0.x, q0.y, q0.z, q0.w, q1.x};
      // the facet belongs to this lane if its cell is the facet's first cell, or if the first cell does not
      // hold the row's dof at all (then the row's dof is the one opposite the facet in the second cell)
      valid = (c == q1.y) || (q1.w == r);
      if (valid)
      {
        const double2 j01 = __ldg(reinterpret_cast<const double2*>(rec) + 2);
        const double2 j23 = __ldg(reinterpret_cast<const double2*>(rec) + 3);
        const double2 j4w = __ldg(reinterpret_cast<const double2*>(rec) + 4);
        const double J[5] = {j01.x, j01.y, j23.x, j23.y, j4w.x};
        double jr = 0.0;
#pragma unroll
        for (int k = 0; k < NE; ++k)
          jr = (dd[k] == r) ? J[k] : jr;
        const double jm = jr * j4w.y;
#pragma unroll
        for (int k = 0; k < 2 * NQ; ++k)
        {
          s_fd[lane][k] = (k < NE) ? dd[k < 5 ? k : 4] : -3; // padding matches no column
          s_fv[lane][k] = (k < NE) ? jm * J[k < 5 ? k : 4] : 0.0;
        }
        if (count)
          expected += NE;
      }
    }
    __syncwarp();
    unsigned m = __ballot_sync(full, valid);
    while (m)
    {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const int2* pd = reinterpret_cast<const int2*>(s_fd[l]);
      const double2* pv = reinterpret_cast<const double2*>(s_fv[l]);
#pragma unroll
      for (int q = 0; q < NQ; ++q)
      {
        const int2 d = pd[q];
        const double2 v = pv[q];
        if (d.x == mycol)
        {
          acc += v.x;
          ++matched;
        }
        if (d.y == mycol)
        {
          acc += v.y;
          ++matched;
        }
      }
    }
    __syncwarp();
  }
  return matched;
}

// P1 band rows with a stored position mask (gather_matrix_fast_kernel).  The macro element of a facet is the lane's
// cell plus ONE more dof, and the record lists the dofs ascending, so the facet's CSR positions in the row are the
// bit mask PM = (positions of the cell's dofs, known from the pattern pass) | (position of the extra dof, one
// lower-bound search in the row's staged columns), and column lane k takes entry popc(PM & lanes_below_k) of the
// staged values -- no column compare per entry.  Same facets, same order (local facet by local facet, lanes
// ascending) and the same products as add_facet_rows_p1.
template <int ND>
__device__ __forceinline__ void add_facet_rows_p1_mask(const GatherCtx& gc, int32_t* s_cols, uint32_t* s_pm,
                                                       double (*s_fv)[6], bool band_cell, int64_t c, int32_t r,
                                                       uint32_t Mc, int32_t mycol, double& acc)
{
  constexpr int NE = ND + 1;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  if (__ballot_sync(full, band_cell) == 0)
    return;
  s_cols[lane] = mycol; // ascending, lanes past the row's end hold INT_MAX
  int32_t fct[4] = {0, 0, 0, 0};
  if (band_cell)
  {
    if (gc.nf == 4)
    {
      const int4 q = __ldg(reinterpret_cast<const int4*>(gc.c2f) + c);
      fct[0] = q.x;
      fct[1] = q.y;
      fct[2] = q.z;
      fct[3] = q.w;
    }
    else
      for (int lf = 0; lf < gc.nf; ++lf)
        fct[lf] = gc.c2f[c * gc.nf + lf];
  }
  int32_t fsl[4] = {-1, -1, -1, -1};
  if (band_cell)
    for (int lf = 0; lf < gc.nf; ++lf)
      fsl[lf] = gc.facet_slot[fct[lf]];
  const uint32_t below = (1u << lane) - 1u;
  __syncwarp();
  for (int lf = 0; lf < gc.nf; ++lf)
  {
    const int64_t fs = fsl[lf];
    bool valid = fs >= 0;
    if (valid)
    {
      // the whole record at once: five independent 16-byte loads (the partner lane of a shared facet reads the
      // same record in the same instruction)
      const double* rec = gc.Fe + fs * FREC;
      const int4 q0 = __ldg(reinterpret_cast<const int4*>(rec));
      const int4 q1 = __ldg(reinterpret_cast<const int4*>(rec) + 1);
      const double2 j01 = __ldg(reinterpret_cast<const double2*>(rec) + 2);
      const double2 j23 = __ldg(reinterpret_cast<const double2*>(rec) + 3);
      const double2 j4w = __ldg(reinterpret_cast<const double2*>(rec) + 4);
      // the facet belongs to this lane if its cell is the facet's first cell, or if the first cell does not
      // hold the row's dof at all (then the row's dof is the one opposite the facet in the second cell)
      const bool first = c == q1.y;
      valid = first || (q1.w == r);
      if (valid)
      {
        const int32_t dd[5] = {q0.x, q0.y, q0.z, q0.w, q1.x};
        const double J[5] = {j01.x, j01.y, j23.x, j23.y, j4w.x};
        const int ix = first ? (q1.z >> 4) : (q1.z & 15); // the macro dof this lane's cell does not have
        double jr = 0.0;
        int32_t dx = 0;
#pragma unroll
        for (int k = 0; k < NE; ++k)
        {
          jr = (dd[k] == r) ? J[k] : jr;
          dx = (k == ix) ? dd[k] : dx;
        }
        int pos = 0; // lower bound of dx among the row's columns
#pragma unroll
        for (int st = 16; st > 0; st >>= 1)
          pos += (s_cols[pos + st - 1] < dx) ? st : 0;
        s_pm[lane] = Mc | (1u << pos);
        const double jm = jr * j4w.y;
#pragma unroll
        for (int k = 0; k < NE; ++k)
          s_fv[lane][k] = jm * J[k];
      }
    }
    __syncwarp();
    unsigned m = __ballot_sync(full, valid);
    while (m)
    {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t PM = s_pm[l];
      if ((PM >> lane) & 1u)
        acc += s_fv[l][__popc(PM & below)];
    }
    __syncwarp();
  }
}

template <int ND>
__device__ __forceinline__ int add_facet_rows(const GatherCtx& gc, int32_t (*s_fd)[FacetStage<ND>::W],
                                              double (*s_fv)[FacetStage<ND>::W], bool band_cell, int64_t c, int li,
                                              int32_t r, int32_t mycol, double& acc, int& expected, bool count)
{
  constexpr int W = FacetStage<ND>::W;
  if constexpr (ND <= 4)
  {
    if (gc.Fw) // P1: combined records
      return add_facet_rows_p1<ND>(gc, s_fd, s_fv, band_cell, c, r, mycol, acc, expected, count);
  }
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int matched = 0;
  if (__ballot_sync(full, band_cell) == 0)
    return 0;
  // the cell's facet ids: one 16-byte load for tetrahedra (lane-distinct gathers cost one L1 wavefront per
  // lane and instruction, so rows are fetched with the widest load that fits them)
  int32_t fct[4] = {0, 0, 0, 0};
  if (band_cell)
  {
    if (gc.nf == 4)
    {
      const int4 q = __ldg(reinterpret_cast<const int4*>(gc.c2f) + c);
      fct[0] = q.x;
      fct[1] = q.y;
      fct[2] = q.z;
      fct[3] = q.w;
    }
    else
      for (int lf = 0; lf < gc.nf; ++lf)
        fct[lf] = gc.c2f[c * gc.nf + lf];
  }
  for (int lf = 0; lf < gc.nf; ++lf)
  {
    int64_t fs = -1;
    if (band_cell)
      fs = gc.facet_slot[fct[lf]];
    const bool valid = fs >= 0;
    if (valid)
    {
      const int4 row = __ldg(reinterpret_cast<const int4*>(gc.rows4) + fs); // (cell0, lf0, cell1, lf1)
      const int64_t c0 = row.x, c1 = row.z;
      const int mrow = (c == c0 ? 0 : ND) + li;
#pragma unroll
      for (int j = 0; j < ND; ++j)
      {
        s_fd[lane][j] = gc.dofmap[c0 * ND + j];
        s_fd[lane][ND + j] = gc.dofmap[c1 * ND + j];
      }
#pragma unroll
      for (int j = 2 * ND; j < W; ++j)
      {
        s_fd[lane][j] = -3; // padding: matches no column
        s_fv[lane][j] = 0.0;
      }
      const double* F = gc.Fe + (fs * 2 * ND + mrow) * 2 * ND;
#pragma unroll
      for (int j = 0; j < 2 * ND; ++j)
        s_fv[lane][j] = F[j];
      if (count)
        expected += 2 * ND;
    }
    __syncwarp();
    unsigned m = __ballot_sync(full, valid);
    while (m)
    {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const int4* pd = reinterpret_cast<const int4*>(s_fd[l]);
      const double2* pv = reinterpret_cast<const double2*>(s_fv[l]);
#pragma unroll
      for (int q = 0; q < W / 4; ++q)
      {
        const int4 d = pd[q];
        const double2 v0 = pv[2 * q], v1 = pv[2 * q + 1];
        if (d.x == mycol)
        {
          acc += v0.x;
          ++matched;
        }
        if (d.y == mycol)
        {
          acc += v0.y;
          ++matched;
        }
        if (d.z == mycol)
        {
          acc += v1.x;
          ++matched;
        }
        if (d.w == mycol)
        {
          acc += v1.y;
          ++matched;
        }
      }
    }
    __syncwarp();
  }
  return matched;
}

// Generic rows (any number of columns / incident cells, or a pattern this form did not build):
// one WARP per active row.
//  phase 1: lanes take the row's incident cells and compute (standard cells) or load
//           (materialised cells) the element-tensor row of this matrix row, staged with the
//           cell's dofs in shared memory;
//  phase 2: lanes own the row's CSR entries (coalesced cols/vals access); every lane walks the
//           staged cells in ascending order (broadcast shared-memory reads) and adds the entry
//           whose dof equals its column -- the search MatrixCSR::mat_add_values does, but with a
//           fixed summation order and no atomics -> bit-reproducible.
template <int TDIM, int DEG>
__global__ void __launch_bounds__(GW * 32)
    gather_matrix_kernel(GatherCtx gc, StdTab st, const int32_t* __restrict__ act_rows, DN n_act_,
                         const uint8_t* __restrict__ skip_fast, const int64_t* __restrict__ row_ptr,
                         const int32_t* __restrict__ cols, double* __restrict__ vals, int zero_first,
                         int32_t* __restrict__ err)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  __shared__ int32_t s_dofs[GW][32][ND];
  __shared__ double s_a[GW][32][ND];
  __shared__ __align__(16) int32_t s_fd[GW][32][FacetStage<ND>::W];
  __shared__ __align__(16) double s_fv[GW][32][FacetStage<ND>::W];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * GW + w;
  if (idx >= n_act_.get())
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
      unsigned fl = 0;
      int li = 0;
      if (k < n_inc)
      {
        c = gc.inc_cell[ib + k];
        fl = gc.cell_flags[c];
      }
      const bool contributes = (fl & 0xFDu) != 0;
      if (fl)
      {
        int32_t d[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          d[j] = gc.dofmap[c * ND + j];
          li = (d[j] == r) ? j : li;
        }
        if (contributes)
        {
          double v[ND];
          cell_row_values<TDIM, DEG>(gc, st, c, fl, li, ib + k, v);
#pragma unroll
          for (int j = 0; j < ND; ++j)
          {
            s_dofs[w][lane][j] = d[j];
            s_a[w][lane][j] = v[j];
          }
        }
      }
      if (!contributes)
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
        expected += contributes ? ND : 0;
      matched += add_facet_rows<ND>(gc, s_fd[w], s_fv[w], (fl & 2u) != 0, c, li, static_cast<int32_t>(r), mycol, acc, expected, kc == 0);
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

// Fast rows (<= 32 columns, <= 32 incident cells, pattern built from this form): one WARP per row.
//  phase 1: lane l takes incident cell l: one coalesced read each of inc_cell / fperm / fmask,
//           the cell's flag byte, then the tensor row of this matrix row -- computed on the fly
//           from the cell geometry for standard cells, loaded (one 32 B sector for P1 tets) for
//           materialised cut cells -- staged in shared memory in ascending-dof order.  The CSR
//           positions of the cell's dofs are a bit mask: static rows derive it from the static
//           full-mesh mask and the row's kept-column mask R (prefix popcounts), band rows read the
//           mask the pattern pass stored.
//  phase 2: lane k owns CSR entry k and walks the cells in ascending order (masks broadcast by
//           shuffle): cell l contributes iff bit k of its mask is set; the value is entry
//           popc(mask & lanes_below_k) of its staged row.
// No dofmap read, no column search, no element-tensor round trip through HBM for standard cells,
// fixed summation order -> bit-reproducible.
// (10 blocks of 4 warps per SM: the kernel is one row per warp with a deep dependent chain, 53 % long-scoreboard
//  stalls at 8 blocks; 48 registers without spills buys 40 resident warps: 0.86 -> 0.73 ms; 12 blocks spill)
template <int TDIM, int DEG>
__global__ void __launch_bounds__(GWM * 32, 10)
    gather_matrix_fast_kernel(GatherCtx gc, StdTab st, StdTab stL, const int32_t* __restrict__ act_rows,
                              const int32_t* __restrict__ slots, DN n_act_,
                              const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ gmask,
                              const uint32_t* __restrict__ Rrow, const int64_t* __restrict__ row_ptr,
                              const int32_t* __restrict__ cols, double* __restrict__ vals, int zero_first)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  __shared__ double s_v[GWM][32][ND];
  // facet staging: P1 -- the row's columns, one position mask and nd + 1 values per hit; otherwise dofs + values
  constexpr bool P1F = ND <= 4;
  __shared__ __align__(16) int32_t s_fd[GWM][32][P1F ? 2 : FacetStage<ND>::W];
  __shared__ __align__(16) double s_fv[GWM][32][P1F ? 6 : FacetStage<ND>::W];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = static_cast<int64_t>(blockIdx.x) * GWM + w;
  if (it >= n_act_.get())
    return;
  // two launches: over the band slot list (slots != null), and -- only if some static row has no
  // contribution list -- over all rows, skipping the listed ones (row_fast bit 16)
  const int64_t idx = slots ? slots[it] : it;
  const unsigned rf = row_fast[idx];
  if (!(rf & 1) || (rf & 12u) == 12u || (!slots && (rf & 16u)))
    return; // handled by gather_matrix_kernel / gather_matrix_clist_kernel / the listed launch
  const unsigned full = 0xffffffffu;
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  // ---- phase 1
  int64_t c = -1;
  unsigned fl = 0, fp = 0, Mn = 0;
  if (lane < n_inc)
  {
    c = gc.inc_cell[ib + lane];
    fp = gc.fperm[ib + lane];
    fl = gc.cell_flags[c];
  }
  const bool contributes = (fl & 0xFDu) != 0;
  const int li = static_cast<int>(fp & 15u);
  uint32_t R = 0;
  if (rf & 4)
  {
    R = Rrow[idx];
    if (contributes)
      for (uint32_t t = gc.fmask[ib + lane]; t; t &= t - 1)
        Mn |= 1u << __popc(R & ((1u << (__ffs(t) - 1)) - 1u));
  }
  else if (contributes || (fl & 2u))
    Mn = gmask[idx * gc.stride + lane];
  const uint32_t Mcell = Mn; // CSR positions of the cell's dofs (band cells have them whether or not they contribute)
  Mn = contributes ? Mn : 0u;
  if (contributes)
  {
    double v[ND];
    cell_row_values<TDIM, DEG>(gc, st, c, fl, li, ib + lane, v);
#pragma unroll
    for (int j = 0; j < ND; ++j)
      s_v[w][lane][(fp >> (4 + 4 * j)) & 15u] = v[j]; // ascending-dof order
  }
  if (gc.bvec)
  { // fused right-hand side (cfx_assemble_system): the same entry, tree and update as gather_vector_kernel
    double e = 0.0;
    if (contributes)
    {
      GatherCtx gl = gc;
      gl.Ae = gc.AeL;
      e = cell_entry_value<TDIM, DEG>(gl, stL, c, fl, li);
    }
    e = rhs_sum_groups_of_four(e);
    if (lane == 0)
    {
      const double s0 = 0.0 + e; // gather_vector_kernel adds the chunk sum to a zero accumulator
      gc.bvec[r] = gc.zero_first_b ? s0 : gc.bvec[r] + s0;
    }
  }
  __syncwarp();
  // ---- phase 2
  const uint32_t below = (1u << lane) - 1u;
  const bool have_col = lane < rn;
  double acc = (have_col && !zero_first) ? vals[rb + lane] : 0.0;
  for (int l0 = 0; l0 < n_inc; l0 += 8)
  {
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int l = (l0 + u) & 31;
      const uint32_t M = __shfl_sync(full, Mn, l);
      t[u] = ((M >> lane) & 1u) ? s_v[w][l][__popc(M & below)] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      acc += t[u]; // ascending-cell order (x + 0.0 == x)
  }
  if (rf & 2)
  {
    if constexpr (P1F)
    {
      if (gc.Fw)
      {
        const int32_t mycol = have_col ? cols[rb + lane] : 0x7fffffff;
        int32_t* si = &s_fd[w][0][0];
        add_facet_rows_p1_mask<ND>(gc, si, reinterpret_cast<uint32_t*>(si) + 32, s_fv[w], (fl & 2u) != 0, c,
                                   static_cast<int32_t>(r), Mcell, mycol, acc);
      }
    }
    else
    {
      const int32_t mycol = have_col ? cols[rb + lane] : -2;
      int expected = 0;
      add_facet_rows<ND>(gc, s_fd[w], s_fv[w], (fl & 2u) != 0, c, li, static_cast<int32_t>(r), mycol, acc, expected, false);
    }
  }
  if (have_col)
    vals[rb + lane] = acc;
}

// Static rows whose contribution lists fit (Space::fclist) -- the bulk of the matrix.  One WARP per row.  A row
// needs four DEPENDENT load levels before any arithmetic:
//   (A) by row slot:  row flags, row id, kept-column mask R
//   (B) by row id:    incidence range, CSR row start, full-pattern row start
//   (C) by position:  lane l: incident cell l and the local index of the row's dof in it;
//                     lane k: the contribution list of full-pattern column k and the CSR value it updates
//   (D) by cell:      flag byte + geometry record (+ materialised tensor row for cut cells)
// Per row: the tensor row of every incident cell is computed on the fly and staged (transposed,
// conflict-free) in shared memory; lane k sums the <= 8 listed (cell, local dof) entries in
// ascending cell order, the diagonal takes one entry per cell through a fixed shuffle tree.  No
// atomics, no dofmap read, no column search, no element tensors through HBM for standard cells.
struct ClistD
{
  unsigned fl;
  int32_t ms;
};

// Rows in chunks of 32 per warp.  ncu on the first, per-row four-level software pipeline (row slot -> row id ->
// incidence -> cell record, one row per stage) showed it bound by per-warp latency: time scales 1/warps up to
// the register-limited 16 warps per SM, and the long-scoreboard stalls sat on values the PREVIOUS iteration had
// loaded -- a pipelined load and its consumer are the same two static instructions every iteration, they share
// a scoreboard, and the consumer also waits for the load just issued.  So the two row-level levels leave the
// per-row chain: a warp takes 32 consecutive rows, lane l loads the scalars of row l (coalesced, two dependent
// levels for 32 rows at once, issued one chunk ahead) and parks them in shared memory; the per-row pipeline
// keeps only the per-cell levels (incidence -> cell record, one row ahead each) and reads its row record with
// three broadcast 128-bit shared loads instead of seven global loads.
struct alignas(16) ClistRow
{
  int32_t r;
  uint32_t R;
  int32_t n_inc; // < 0: not a contribution-list row (or past the end)
  int32_t ufl;   // the flag byte every incident cell carries (a standard-quadrature one), 0 = look at the cells
  int64_t ib, rb, fb, pad2;
};

// resident blocks per SM: the P1 kernel keeps 4-double tensor-row records (not 10-double geometry records) in its
// pipeline registers, which fits five blocks of four warps without spilling
#ifndef CFX_CLIST_BLOCKS_P1
#define CFX_CLIST_BLOCKS_P1 5
#endif
template <int DEG>
constexpr int clist_blocks_per_sm()
{
  return DEG == 1 ? CFX_CLIST_BLOCKS_P1 : 4;
}

template <int TDIM, int DEG, bool FUSED>
__global__ void __launch_bounds__(GWC * 32, clist_blocks_per_sm<DEG>())
    gather_matrix_clist_kernel(GatherCtx gc, StdTab st, StdTab stL, const int32_t* __restrict__ act_rows, DN n_act_,
                                const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ Rrow,
                                const uint8_t* __restrict__ row_ufl, const int64_t* __restrict__ row_ptr,
                                double* __restrict__ vals, int zero_first)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  static_assert(ND * 32 <= 255, "contribution-list bytes index the staging array directly");
  __shared__ double s_v[GWC][256];
  __shared__ ClistRow s_rec[GWC][2][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const uint32_t below = (1u << lane) - 1u;
  double* const sv = s_v[w];
  if (lane == 0)
    sv[255] = 0.0;
  const int n_rows = static_cast<int>(n_act_.get());
  const int n_chunks = (n_rows + 31) >> 5;
  const int cstride = static_cast<int>(gridDim.x) * GWC;
  int chunk = static_cast<int>(blockIdx.x) * GWC + w;
  if (chunk >= n_chunks)
    return;

  // lane-parallel row records of one chunk: level 1 (by slot), level 2 (by row id)
  struct L1
  {
    unsigned rf;
    int32_t r;
    uint32_t R;
    int32_t ufl;
  };
  auto load1 = [&](int ch) -> L1
  {
    L1 a;
    const int idx = ch * 32 + lane;
    const bool in = ch < n_chunks && idx < n_rows;
    const int k = in ? idx : n_rows - 1; // clamped: the loads are always legal
    a.rf = in ? row_fast[k] : 0u;
    a.r = act_rows[k];
    a.R = Rrow[k];
    a.ufl = row_ufl[k];
    return a;
  };
  auto load2 = [&](const L1& a) -> ClistRow
  {
    ClistRow q;
    q.r = a.r;
    q.R = a.R;
    q.ib = gc.inc_ptr[a.r];
    const int n_inc = static_cast<int>(gc.inc_ptr[a.r + 1] - q.ib);
    q.n_inc = ((a.rf & 13u) == 13u) ? n_inc : -1;
    q.ufl = a.ufl;
    q.rb = row_ptr[a.r];
    q.fb = gc.frow_ptr[a.r];
    q.pad2 = 0;
    return q;
  };
  struct RowC
  {
    bool ok;
    int32_t r;
    uint32_t R;
    int n_inc;
    int32_t c;
    int li;
    int ufl;
    int64_t pinc;
    uint64_t word;
    double* pv;
    double old;
  };
  // level C of row k of the chunk in buffer `buf`: the row record (three broadcast shared loads), then by position
  // lane l: incident cell l and the local index of the row's dof in it; lane k: contribution list + old value
  auto stageC = [&](int buf, int k, int nk) -> RowC
  {
    RowC c;
    const int kk = k < nk ? k : nk - 1;
    const int4 q0 = *reinterpret_cast<const int4*>(&s_rec[w][buf][kk]);
    const longlong2 q1 = *reinterpret_cast<const longlong2*>(&s_rec[w][buf][kk].ib);
    const int64_t fb = s_rec[w][buf][kk].fb;
    c.ok = k < nk && q0.z >= 0;
    c.r = q0.x;
    c.R = static_cast<uint32_t>(q0.y);
    c.n_inc = q0.z < 0 ? 0 : q0.z;
    const int64_t pos = q1.x + (lane < c.n_inc ? lane : 0);
    c.ufl = q0.w;
    c.pinc = pos;
    c.c = gc.inc_cell[pos];
    c.li = static_cast<int>(gc.fperm[pos] & 15u);
    c.word = __ldg(gc.fclist + fb + lane); // padded allocation: legal for every lane
    c.pv = vals + q1.y + __popc(c.R & below);
    const bool kept = (c.R >> lane) & 1u;
    c.old = (c.ok && kept && !zero_first) ? *c.pv : 0.0;
    return c;
  };
  // per-cell record of level D: P1 -- the 32-byte tensor-row record of (cell, li) (one sector, read by this row
  // only); otherwise the geometry record
  using CellRec = typename std::conditional<DEG == 1, P1Rec, Geo<TDIM>>::type;
  auto stageD = [&](const RowC& c, CellRec& g) -> ClistD
  {
    ClistD d;
    // a row whose cells all carry one standard-quadrature flag byte needs neither the flags nor the tensor slots
    // of its cells (warp-uniform branch): two scattered gathers per incidence less
    const bool in = c.ok && lane < c.n_inc;
    if (c.ufl)
    {
      d.fl = in ? static_cast<unsigned>(c.ufl) : 0u;
      d.ms = 0;
    }
    else
    {
      d.fl = in ? gc.cell_flags[c.c] : 0u;
      d.ms = __ldg(gc.mat_slot + c.c);
    }
    if constexpr (DEG == 1)
      g = load_p1rec(gc.lrow, c.pinc);
    else
      load_geo_cached<TDIM>(gc.geo, c.c, g);
    return d;
  };

  // first chunk: both levels exposed once per warp
  {
    const ClistRow q = load2(load1(chunk));
    s_rec[w][0][lane] = q;
  }
  __syncwarp();
  int buf = 0;
  for (; chunk < n_chunks; chunk += cstride, buf ^= 1)
  {
    // the next chunk's level 1 goes out now, its level 2 after the first rows, both are parked at the chunk's end
    const L1 n1 = load1(chunk + cstride);
    const int nk = min(32, n_rows - chunk * 32);
    RowC c0 = stageC(buf, 0, nk), c1 = stageC(buf, 1, nk);
    CellRec g0;
    ClistD d0 = stageD(c0, g0);
    ClistRow nq = load2(n1);
#pragma unroll 2
    for (int k = 0; k < nk; ++k)
    {
      CellRec g1;
      const ClistD d1 = stageD(c1, g1);
      const RowC c2 = stageC(buf, k + 2, nk);
      if (c0.ok)
      {
        const unsigned fl = d0.fl;
        const bool contributes = (fl & 0xFDu) != 0;
        double v[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
          v[j] = 0.0;
        double e = 0.0;
        if (contributes)
        {
          if (fl >> 2)
          {
            if constexpr (DEG == 1)
              p1_row_from_rec<TDIM>(st, g0, fl, c0.li, v);
            else
              std_row_values<TDIM, DEG>(st, g0, fl, c0.li, v);
            if constexpr (FUSED)
            {
              double s;
              if constexpr (DEG == 1)
                s = g0.a[3];
              else
                s = fabs(g0.detJ);
              if constexpr (DEG == 1)
                e += stL.t0[fl >> 2] * s * (TDIM == 3 ? 1.0 / 24.0 : 1.0 / 6.0);
              else if (stL.ref != nullptr)
                e += stL.t0[fl >> 2] * s * stL.ref[RefTab<TDIM, ND>::S + c0.li];
              else
                for (int kk = 0; kk < stL.n; ++kk)
                {
                  if (!(fl & stL.bit[kk]))
                    continue;
                  for (int q = 0; q < stL.npts[kk]; ++q)
                  {
                    double xi[TDIM];
#pragma unroll
                    for (int t = 0; t < TDIM; ++t)
                      xi[t] = __ldg(stL.pts[kk] + q * TDIM + t);
                    double phi[ND], dphi[ND][TDIM];
                    tabulate<TDIM, DEG>(xi, phi, dphi);
                    e += stL.c[kk][0] * (__ldg(stL.wts[kk] + q) * s) * pick<ND>(phi, c0.li);
                  }
                }
            }
          }
          if (fl & 1)
          {
            const double* a = gc.Ae + (static_cast<int64_t>(d0.ms) * ND + c0.li) * ND;
#pragma unroll
            for (int j = 0; j < ND; ++j)
              v[j] += a[j];
            if constexpr (FUSED)
              e += gc.AeL[static_cast<int64_t>(d0.ms) * ND + c0.li];
          }
        }
        double dval = pick<ND>(v, c0.li);
#pragma unroll
        for (int j = 0; j < ND; ++j)
          sv[j * 32 + lane] = v[j];
        __syncwarp();
        if constexpr (FUSED)
        { // the right-hand-side entry in the one order every kernel uses (rhs_sum_groups_of_four)
          const double x = rhs_sum_groups_of_four(e);
          if (lane == 0)
            gc.bvec[c0.r] = gc.zero_first_b ? x : gc.bvec[c0.r] + x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          dval += __shfl_down_sync(full, dval, o);
        dval = __shfl_sync(full, dval, 0);
        if ((c0.R >> lane) & 1u)
        {
          double acc = c0.old;
          const uint32_t lo = static_cast<uint32_t>(c0.word), hi = static_cast<uint32_t>(c0.word >> 32);
          if ((lo & 0xFFu) == 0xFEu)
            acc += dval;
          else
          {
            const double t0 = sv[lo & 0xFFu], t1 = sv[(lo >> 8) & 0xFFu], t2 = sv[(lo >> 16) & 0xFFu], t3 = sv[lo >> 24];
            const double t4 = sv[hi & 0xFFu], t5 = sv[(hi >> 8) & 0xFFu], t6 = sv[(hi >> 16) & 0xFFu], t7 = sv[hi >> 24];
            acc += t0;
            acc += t1;
            acc += t2;
            acc += t3;
            acc += t4;
            acc += t5;
            acc += t6;
            acc += t7;
          }
          *c0.pv = acc;
        }
        __syncwarp(); // s_v is reused by the next row
      }
      c0 = c1;
      c1 = c2;
      d0 = d1;
      g0 = g1;
    }
    s_rec[w][buf ^ 1][lane] = nq;
    __syncwarp();
  }
}

// Scalar P1 static rows: ONE THREAD PER ROW.  ncu on the warp-per-row contribution-list kernel above showed it bound
// by issue slots and per-warp dependent latency (about 280 warp instructions and 2300 cycles per row: staging through
// shared memory, shuffle trees, pipeline bookkeeping -- all warp-wide instructions that serve one row), not by bytes:
// taking two of its three per-cell gathers away changed nothing.  A P1 row needs per incident cell one 32-byte
// record (three off-diagonal Laplace entries + |detJ|, Space::lrow, stored in incidence order) and one packed word
// of static row positions (Space::fpos), so a thread can walk its row alone: per incidence one 256-bit load, one
// 32-bit load, a handful of FP64 operations and three read-modify-writes of its private accumulator column in shared
// memory (column tid of acc[32][RTB]: the bank depends on the thread only, never conflicts); the diagonal and the
// right-hand-side entry stay in registers.  About 25 warp instructions per row instead of 280, four independent
// record loads in flight per thread, no shuffles, no barriers.  Sums run over the incident cells in ascending order
// (the order of the contribution lists); rows whose cells do not all carry one standard-quadrature flag byte
// (row_ufl == 0) look the flags up per cell and add the materialised tensor rows of cut cells.
constexpr int RTB = 128;

__device__ __forceinline__ uint32_t ldg_keep(const uint32_t* p)
{ // the position words are re-read from the same 32-byte sector by eight consecutive incidences: keep them in L1
  uint32_t v;
  asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ P1Rec ldg_stream_p1rec(const double* __restrict__ lrow, int64_t pinc)
{ // each record is read exactly once per assembly: do not let the stream push the position words out of L1
  P1Rec r;
  const double* p = lrow + pinc * 4;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.a[0]), "=d"(r.a[1]), "=d"(r.a[2]), "=d"(r.a[3])
               : "l"(p));
  return r;
}

// NC: rows of the private accumulator tile = an upper bound of the static row length (Space::max_fcols); 16 for the
// usual P1 meshes (Kuhn tetrahedra: 15 columns, right-diagonal triangles: 7), which halves the shared memory per block
template <int TDIM, bool FUSED, int NC, int MINB>
__global__ void __launch_bounds__(RTB, MINB)
    gather_matrix_p1_kernel(GatherCtx gc, StdTab st, StdTab stL, const int32_t* __restrict__ act_rows, DN n_act_,
                            const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ Rrow,
                            const uint8_t* __restrict__ row_ufl, const uint32_t* __restrict__ fpos,
                            const int64_t* __restrict__ row_ptr, double* __restrict__ vals, int zero_first)
{
  constexpr int ND = TDIM + 1, NO = TDIM; // NO off-diagonal entries per tensor row
  __shared__ double s_acc[NC][RTB];
  const int tid = threadIdx.x;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * RTB + tid;
  if (idx >= n_act_.get())
    return;
  if ((row_fast[idx] & 13u) != 13u)
    return; // band rows, rows without a static structure: the mask / generic kernels
  const int64_t r = act_rows[idx];
  const uint32_t R = Rrow[idx];
  const unsigned ufl = row_ufl[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  double* const out = vals + row_ptr[r];
  {
    int o = 0;
    for (uint32_t m = R; m; m &= m - 1, ++o)
      s_acc[__ffs(m) - 1][tid] = zero_first ? 0.0 : out[o];
  }
  const int pd = static_cast<int>((ldg_keep(fpos + ib) >> 2) & 31u); // the row's own column
  double dacc = zero_first ? 0.0 : s_acc[pd][tid];
  double e = 0.0; // right-hand-side entry, summed in the order of rhs_sum_groups_of_four
  constexpr double MASSW = TDIM == 3 ? 1.0 / 120.0 : 1.0 / 24.0;
  constexpr double SRCW = TDIM == 3 ? 1.0 / 24.0 : 1.0 / 6.0;
  constexpr int U = 4;
  for (int l0 = 0; l0 < n_inc; l0 += U)
  {
    P1Rec rec[U];
    uint32_t word[U];
    unsigned fl[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int l = l0 + u < n_inc ? l0 + u : n_inc - 1;
      rec[u] = ldg_stream_p1rec(gc.lrow, ib + l);
      word[u] = ldg_keep(fpos + ib + l);
      fl[u] = ufl;
    }
    int32_t cell[U] = {0, 0, 0, 0};
    if (!ufl)
    {
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        const int l = l0 + u < n_inc ? l0 + u : n_inc - 1;
        cell[u] = gc.inc_cell[ib + l];
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        fl[u] = gc.cell_flags[cell[u]];
    }
    double eu[U] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (l0 + u >= n_inc || !(fl[u] & 0xFDu))
        continue;
      const unsigned m = fl[u] >> 2;
      double v[NO], vd = 0.0;
#pragma unroll
      for (int q = 0; q < NO; ++q)
        v[q] = 0.0;
      if (m)
      { // the arithmetic of p1_row_from_rec
        const double w = st.t0[m];
        double dsum = 0.0;
#pragma unroll
        for (int q = 0; q < NO; ++q)
        {
          const double t = w * rec[u].a[q];
          dsum += t;
          v[q] += t;
        }
        vd -= dsum;
        if (st.has_mass)
        {
          const double wm = st.t1[m] * rec[u].a[3] * MASSW;
#pragma unroll
          for (int q = 0; q < NO; ++q)
            v[q] += wm;
          vd += 2.0 * wm;
        }
        if constexpr (FUSED)
          eu[u] += stL.t0[m] * rec[u].a[3] * SRCW;
      }
      if (fl[u] & 1u)
      { // materialised run-time-rule tensor row of a cut cell (natural dof order)
        const int li = static_cast<int>(word[u] & 3u);
        const int64_t ms = __ldg(gc.mat_slot + cell[u]);
        const double* a = gc.Ae + (ms * ND + li) * ND;
        double av[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
          av[j] = a[j];
#pragma unroll
        for (int q = 0; q < NO; ++q)
          v[q] += pick<ND>(av, q < li ? q : q + 1);
        vd += pick<ND>(av, li);
        if constexpr (FUSED)
          eu[u] += gc.AeL[ms * ND + li];
      }
#pragma unroll
      for (int q = 0; q < NO; ++q)
      {
        const int p = static_cast<int>((word[u] >> (7 + 5 * q)) & 31u);
        s_acc[p][tid] += v[q];
      }
      dacc += vd;
    }
    if constexpr (FUSED)
    {
      const double g4 = (eu[0] + eu[1]) + (eu[2] + eu[3]);
      e = l0 == 0 ? g4 : e + g4;
    }
  }
  s_acc[pd][tid] = dacc;
  {
    int o = 0;
    for (uint32_t m = R; m; m &= m - 1, ++o)
      out[o] = s_acc[__ffs(m) - 1][tid];
  }
  if constexpr (FUSED)
    gc.bvec[r] = gc.zero_first_b ? e : gc.bvec[r] + e;
}

// Scalar P2 on triangles, static rows: one thread per row like gather_matrix_p1_kernel.  Per incidence one 32-byte
// record (C = K K^T: C00, C01, C11, and |detJ|; Space::lrow) and one 64-bit word of static positions (Space::fpos64:
// local index, the row's own column, the cell's five other dofs); the tensor row is
//   v_j = w (C00 T0[li][j] + C01 T1[li][j] + C11 T2[li][j]),  T0 = R[0][0], T1 = R[0][1] + R[1][0], T2 = R[1][1]
// with the reference-element tables staged in shared memory once per block (RefTab).  Vertex rows have 6 cells and
// 19 columns, edge rows 2 cells and 9 columns; the warp-per-row contribution-list kernel it replaces ran at 5 % of
// the HBM peak on BASELINE configs[1] (4096^2) for exactly that reason -- a 32-lane warp per 2-cell row.
template <bool FUSED>
__global__ void __launch_bounds__(RTB)
    gather_matrix_p2tri_kernel(GatherCtx gc, StdTab st, StdTab stL, const int32_t* __restrict__ act_rows, DN n_act_,
                               const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ Rrow,
                               const uint8_t* __restrict__ row_ufl, const uint64_t* __restrict__ fpos64,
                               const int64_t* __restrict__ row_ptr, double* __restrict__ vals, int zero_first)
{
  constexpr int ND = 6;
  using T = RefTab<2, ND>;
  __shared__ double s_acc[32][RTB];
  __shared__ double s_T[3][ND][ND], s_M[ND][ND], s_S[ND];
  const int tid = threadIdx.x;
  for (int t = tid; t < ND * ND; t += RTB)
  {
    const int i = t / ND, j = t - i * ND;
    s_T[0][i][j] = st.ref[T::R + (0 * ND + i) * ND + j];
    s_T[1][i][j] = st.ref[T::R + (1 * ND + i) * ND + j] + st.ref[T::R + (2 * ND + i) * ND + j];
    s_T[2][i][j] = st.ref[T::R + (3 * ND + i) * ND + j];
    s_M[i][j] = st.ref[T::M + t];
    if (j == 0)
      s_S[i] = st.ref[T::S + i];
  }
  __syncthreads();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * RTB + tid;
  if (idx >= n_act_.get())
    return;
  if ((row_fast[idx] & 13u) != 13u)
    return;
  const int64_t r = act_rows[idx];
  const uint32_t R = Rrow[idx];
  const unsigned ufl = row_ufl[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  double* const out = vals + row_ptr[r];
  {
    int o = 0;
    for (uint32_t m = R; m; m &= m - 1, ++o)
      s_acc[__ffs(m) - 1][tid] = zero_first ? 0.0 : out[o];
  }
  const int pd = static_cast<int>((fpos64[ib] >> 3) & 31u);
  double dacc = zero_first ? 0.0 : s_acc[pd][tid];
  double e = 0.0;
  constexpr int U = 4;
  for (int l0 = 0; l0 < n_inc; l0 += U)
  {
    double eu[U] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int l = l0 + u;
      if (l >= n_inc)
        continue;
      const P1Rec rec = ldg_stream_p1rec(gc.lrow, ib + l);
      const uint64_t word = fpos64[ib + l];
      int32_t cell = 0;
      unsigned fl = ufl;
      if (!ufl)
      {
        cell = gc.inc_cell[ib + l];
        fl = gc.cell_flags[cell];
      }
      if (!(fl & 0xFDu))
        continue;
      const int li = static_cast<int>(word & 7u);
      const unsigned m = fl >> 2;
      double v[ND];
#pragma unroll
      for (int j = 0; j < ND; ++j)
        v[j] = 0.0;
      if (m)
      {
        const double s = rec.a[3];
        const double wl = st.t0[m] * s;
        const double c0 = rec.a[0] * wl, c1 = rec.a[1] * wl, c2 = rec.a[2] * wl;
#pragma unroll
        for (int j = 0; j < ND; ++j)
          v[j] += c0 * s_T[0][li][j] + c1 * s_T[1][li][j] + c2 * s_T[2][li][j];
        if (st.has_mass)
        {
          const double wm = st.t1[m] * s;
#pragma unroll
          for (int j = 0; j < ND; ++j)
            v[j] += wm * s_M[li][j];
        }
        if constexpr (FUSED)
          eu[u] += stL.t0[m] * s * s_S[li];
      }
      if (fl & 1u)
      {
        const int64_t ms = __ldg(gc.mat_slot + cell);
        const double* a = gc.Ae + (ms * ND + li) * ND;
#pragma unroll
        for (int j = 0; j < ND; ++j)
          v[j] += a[j];
        if constexpr (FUSED)
          eu[u] += gc.AeL[ms * ND + li];
      }
      // local dof j != li sits at the q-th of the five packed positions, q = j - (j > li)
      int q = 0;
#pragma unroll
      for (int j = 0; j < ND; ++j)
      {
        if (j == li)
          dacc += v[j];
        else
        {
          const int p = static_cast<int>((word >> (8 + 5 * q)) & 31u);
          s_acc[p][tid] += v[j];
          ++q;
        }
      }
    }
    if constexpr (FUSED)
    {
      const double g4 = (eu[0] + eu[1]) + (eu[2] + eu[3]);
      e = l0 == 0 ? g4 : e + g4;
    }
  }
  s_acc[pd][tid] = dacc;
  {
    int o = 0;
    for (uint32_t m = R; m; m &= m - 1, ++o)
      out[o] = s_acc[__ffs(m) - 1][tid];
  }
  if constexpr (FUSED)
    gc.bvec[r] = gc.zero_first_b ? e : gc.bvec[r] + e;
}

// Scalar P1 BAND rows (rows with a ghost-penalty facet among their cells), one thread per row like
// gather_matrix_p1_kernel.  A band row has columns the static full-mesh row does not have (the dofs across the
// facets opposite the row's dof), so the thread first stages its CSR columns in a private shared-memory column and
// builds the map static position -> CSR position (one lower-bound search per kept static column); cell tensors then
// go exactly as in gather_matrix_p1_kernel.  Facets: per incident band cell the cell's facet ids (one 16-byte load),
// their slots (independent probes), the 80-byte records of facet_p1_kernel; a facet is taken by its first cell, or
// by the second when the first does not hold the row's dof; each of the nd + 1 macro dofs is located by a
// lower-bound search in the staged columns.  Fixed order: cells ascending, then facets cell by cell, local facet by
// local facet -- no atomics, bit-reproducible.  (The warp-per-row mask kernel this replaces for P1 spent 1100 of its
// 1800 warp instructions per row matching facet entries against columns lane by lane.)
constexpr int RTBB = 64;

// G threads per row (a power of two <= 8, groups inside one warp): thread g takes the incident cells l = g, g + G, ...
// into its own accumulator column; the group's columns are added in the fixed order g = 0 .. G-1 at the end.  A
// band row is a long dependent chain (cell -> flags -> facets -> slots -> records -> searches) and there are few of
// them, so one thread per row leaves the machine waiting; G threads cut the chain G times.
template <int TDIM, bool FUSED, int G>
__global__ void __launch_bounds__(RTBB, 12)
    gather_matrix_band_p1_kernel(GatherCtx gc, StdTab st, StdTab stL, const int32_t* __restrict__ act_rows,
                                 const int32_t* __restrict__ slots, DN n_band_, const uint8_t* __restrict__ row_fast,
                                 const int32_t* __restrict__ fcols, const int64_t* __restrict__ row_ptr,
                                 const int32_t* __restrict__ cols, double* __restrict__ vals, int zero_first)
{
  constexpr int ND = TDIM + 1, NO = TDIM, NE = ND + 1;
  __shared__ double s_acc[32][RTBB];
  // the row's columns and the static-position map: ONE copy per group of G threads (its members read the same
  // word: a broadcast; the 8 groups of a warp sit in 8 different banks whatever rows of the arrays they read) --
  // 18.5 KB per block instead of 26 KB, 12 resident blocks per SM instead of 8
  __shared__ int32_t s_cols[32][RTBB / G];
  __shared__ uint8_t s_map[32][RTBB / G];
  const int grp = threadIdx.x / G;
  const int tid = threadIdx.x;
  const int g = tid & (G - 1);
  const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
  const int64_t it = (static_cast<int64_t>(blockIdx.x) * RTBB + tid) / G;
  if (it >= n_band_.get())
    return;
  const int64_t idx = slots[it];
  const unsigned rf = row_fast[idx];
  if (!(rf & 1u) || (rf & 4u))
    return; // more than 32 columns / incident cells: the generic kernel
  const int32_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  const int64_t fb = gc.frow_ptr[r];
  const int nfull = static_cast<int>(gc.frow_ptr[r + 1] - fb);
  // the row's columns and the static-position map: every thread of the group fills a share of the group's copy
  const int t0 = tid - g;
  for (int k = g; k < 32; k += G)
    s_cols[k][grp] = k < rn ? cols[rb + k] : 0x7fffffff;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    s_acc[k][tid] = 0.0;
  __syncwarp(gmask);
  auto find = [&](int32_t d) -> int
  { // lower bound of d among the row's (ascending) columns
    int pos = 0;
#pragma unroll
    for (int sp = 16; sp > 0; sp >>= 1)
      pos += (s_cols[pos + sp - 1][grp] < d) ? sp : 0;
    return pos;
  };
  for (int p = g; p < nfull; p += G)
    s_map[p][grp] = static_cast<uint8_t>(find(fcols[fb + p])); // columns the row did not keep are never looked up
  __syncwarp(gmask);
  double e = 0.0;
  constexpr double MASSW = TDIM == 3 ? 1.0 / 120.0 : 1.0 / 24.0;
  constexpr double SRCW = TDIM == 3 ? 1.0 / 24.0 : 1.0 / 6.0;
  // ---- this thread's cells, two at a time: tensors, then facets
  static_assert(G == 4, "the right-hand-side order (rhs_sum_groups_of_four) needs one group of four cells per pass");
  constexpr int U = 2;
  for (int l0 = g; l0 - g < n_inc; l0 += U * G) // the same trip count for the whole group (it shuffles inside)
  {
    P1Rec rec[U];
    uint32_t word[U], fm[U];
    int32_t cell[U];
    unsigned fl[U];
    bool in[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      in[u] = l0 + u * G < n_inc;
      const int l = in[u] ? l0 + u * G : 0;
      rec[u] = ldg_stream_p1rec(gc.lrow, ib + l);
      word[u] = ldg_keep(gc.fpos + ib + l);
      fm[u] = gc.fmask[ib + l];
      cell[u] = gc.inc_cell[ib + l];
    }
    // by cell: the flag byte and -- whether or not the cell turns out to be a band cell -- its facet ids (one
    // level of the dependent chain less than asking the flags first)
    int32_t fct[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      fl[u] = in[u] ? gc.cell_flags[cell[u]] : 0u;
      if (ND == 4)
      {
        const int4 q = __ldg(reinterpret_cast<const int4*>(gc.c2f) + cell[u]);
        fct[u][0] = q.x;
        fct[u][1] = q.y;
        fct[u][2] = q.z;
        fct[u][3] = q.w;
      }
      else
      {
#pragma unroll
        for (int lf = 0; lf < ND; ++lf)
          fct[u][lf] = gc.c2f[static_cast<int64_t>(cell[u]) * ND + lf];
        fct[u][3] = 0;
      }
    }
    int32_t fsl[U][ND];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int lf = 0; lf < ND; ++lf)
        fsl[u][lf] = (fl[u] & 2u) ? gc.facet_slot[fct[u][lf]] : -1;
    double eu[U] = {0.0, 0.0};
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (!(fl[u] & 0xFDu))
        continue;
      const unsigned m = fl[u] >> 2;
      double v[NO], vd = 0.0;
#pragma unroll
      for (int q = 0; q < NO; ++q)
        v[q] = 0.0;
      if (m)
      {
        const double w = st.t0[m];
        double dsum = 0.0;
#pragma unroll
        for (int q = 0; q < NO; ++q)
        {
          const double t = w * rec[u].a[q];
          dsum += t;
          v[q] += t;
        }
        vd -= dsum;
        if (st.has_mass)
        {
          const double wm = st.t1[m] * rec[u].a[3] * MASSW;
#pragma unroll
          for (int q = 0; q < NO; ++q)
            v[q] += wm;
          vd += 2.0 * wm;
        }
        if constexpr (FUSED)
          eu[u] += stL.t0[m] * rec[u].a[3] * SRCW;
      }
      if (fl[u] & 1u)
      {
        const int li = static_cast<int>(word[u] & 3u);
        const int64_t ms = __ldg(gc.mat_slot + cell[u]);
        const double* a = gc.Ae + (ms * ND + li) * ND;
        double av[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
          av[j] = a[j];
#pragma unroll
        for (int q = 0; q < NO; ++q)
          v[q] += pick<ND>(av, q < li ? q : q + 1);
        vd += pick<ND>(av, li);
        if constexpr (FUSED)
          eu[u] += gc.AeL[ms * ND + li];
      }
#pragma unroll
      for (int q = 0; q < NO; ++q)
        s_acc[s_map[(word[u] >> (7 + 5 * q)) & 31u][grp]][tid] += v[q];
      s_acc[s_map[(word[u] >> 2) & 31u][grp]][tid] += vd;
    }
    if constexpr (FUSED)
    { // the group's four cells of each pass: pairwise, then the passes in order (rhs_sum_groups_of_four)
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        double x = eu[u];
        x += __shfl_xor_sync(gmask, x, 1);
        x += __shfl_xor_sync(gmask, x, 2);
        if (l0 - g + u * G < n_inc)
          e = (l0 - g + u * G == 0) ? x : e + x;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (!(fl[u] & 2u))
        continue;
      // CSR positions of the cell's dofs in ASCENDING dof order: the set bits of its static position mask,
      // mapped (positions are monotone in the dof number)
      int cp[ND];
      {
        uint32_t mm = fm[u];
#pragma unroll
        for (int t = 0; t < ND; ++t)
        {
          cp[t] = s_map[(__ffs(mm) - 1) & 31][grp];
          mm &= mm - 1;
        }
      }
#pragma unroll
      for (int lf = 0; lf < ND; ++lf)
      {
        if (fsl[u][lf] < 0)
          continue;
        const double* frec = gc.Fe + static_cast<int64_t>(fsl[u][lf]) * FREC;
        const int4 q1 = __ldg(reinterpret_cast<const int4*>(frec) + 1);
        const int4 q0 = __ldg(reinterpret_cast<const int4*>(frec));
        const double2 j01 = __ldg(reinterpret_cast<const double2*>(frec) + 2);
        const double2 j23 = __ldg(reinterpret_cast<const double2*>(frec) + 3);
        const double2 j4w = __ldg(reinterpret_cast<const double2*>(frec) + 4);
        const bool first = cell[u] == q1.y;
        if (!first && q1.w != r)
          continue; // the facet's first cell holds the row's dof too: it takes the facet
        const int32_t dd[5] = {q0.x, q0.y, q0.z, q0.w, q1.x};
        const double J[5] = {j01.x, j01.y, j23.x, j23.y, j4w.x};
        // the record lists the macro dofs ascending: all but one (index ix) are this cell's dofs, in the order of cp
        const int ix = first ? (q1.z >> 4) : (q1.z & 15);
        double jr = 0.0;
        int32_t dx = 0;
#pragma unroll
        for (int k = 0; k < NE; ++k)
        {
          jr = (dd[k] == r) ? J[k] : jr;
          dx = (k == ix) ? dd[k] : dx;
        }
        const int px = find(dx);
        const double jm = jr * j4w.y;
#pragma unroll
        for (int k = 0; k < NE; ++k)
        {
          const int own = k < ix ? cp[k < ND ? k : ND - 1] : cp[k > 0 ? k - 1 : 0];
          s_acc[k == ix ? px : own][tid] += jm * J[k];
        }
      }
    }
  }
  __syncwarp(gmask);
  // the group's columns in the fixed order 0 .. G-1 (on top of the old value unless the matrix is being overwritten)
  for (int k = g; k < rn; k += G)
  {
    double x = zero_first ? 0.0 : vals[rb + k];
#pragma unroll
    for (int m = 0; m < G; ++m)
      x += s_acc[k][t0 + m];
    vals[rb + k] = x;
  }
  if constexpr (FUSED)
  {
    if (g == 0)
      gc.bvec[r] = gc.zero_first_b ? e : gc.bvec[r] + e;
  }
}

// One warp per active row: lanes take the incident cells, compute / load the cell's entry for this
// row, fixed shuffle tree -> bit-reproducible.
template <int TDIM, int DEG, bool PERM>
__global__ void __launch_bounds__(GW * 32)
    gather_vector_kernel(GatherCtx gc, StdTab st, const int32_t* __restrict__ act_rows, DN n_act_,
                         double* __restrict__ b, int zero_first, const int32_t* __restrict__ slots,
                         const uint8_t* __restrict__ row_fast, int skip_mode)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = static_cast<int64_t>(blockIdx.x) * GW + w;
  if (it >= n_act_.get())
    return;
  // skip_mode (fused system assembly): 1 = rows of the slot list that gather_matrix_clist_kernel did not
  // fill; 2 = all rows except those it filled and the listed ones (row_fast bit 16)
  const int64_t idx = slots ? slots[it] : it;
  if (skip_mode)
  {
    const unsigned rf = row_fast[idx];
    if (skip_mode == 3)
    { // fused system assembly: the contribution-list and mask kernels have filled b for every fast row
      if (rf & 1u)
        return;
    }
    else if ((rf & 13u) == 13u || (skip_mode == 2 && (rf & 16u)))
      return;
  }
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  double s = 0.0;
  for (int k0 = 0; k0 < n_inc; k0 += 32)
  {
    const int k = k0 + lane;
    double e = 0.0;
    if (k < n_inc)
    {
      const int64_t c = gc.inc_cell[ib + k];
      const unsigned fl = gc.cell_flags[c];
      if (fl & 0xFDu)
      {
        int li = 0;
        if constexpr (PERM)
          li = static_cast<int>(gc.fperm[ib + k] & 15u);
        else
        {
#pragma unroll
          for (int j = 0; j < ND; ++j)
            li = (gc.dofmap[c * ND + j] == r) ? j : li;
        }
        e = cell_entry_value<TDIM, DEG>(gc, st, c, fl, li);
      }
    }
    e = rhs_sum_groups_of_four(e);
    s += e;
  }
  if (lane == 0)
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
// exact entity counts of an integral (on the device in deferred-size mode)
inline DN dn_entities(const cfx_integral& I) { return DN{I.d_n, I.n, I.facet ? 2 : 0}; }
inline DN dn_rules(const cfx_rules* R) { return DN{R->deferred ? R->d_sizes : nullptr, R->nrules, 0}; }
inline RuleView make_rule_view(const cfx_rules* R, bool normals, bool moments, int mom_stride)
{
  return RuleView{R->points.p, R->weights.p, normals ? R->normals.p : nullptr, R->offsets.p, R->parent_map.p, R->npts,
                  moments ? R->moments.p : nullptr, mom_stride, R->deferred ? R->d_sizes : nullptr};
}

int std_rule_order(int kernel, int deg)
{
  switch (kernel)
  {
  case CFX_K_LAPLACE: return 2 * (deg - 1);
  case CFX_K_MASS: return 2 * deg;
  case CFX_K_SOURCE: return deg;
  case CFX_K_SQUARE_FN: return 2 * deg; // (uh - u_exact)**2 with both in the space: estimated degree 2 deg
  default: return 0;
  }
}

template <int TDIM, int DEG, int KID>
void launch_cell(cfx_ctx* c, const cfx_integral& I, cfx_form* f, int64_t& base)
{
  constexpr int RANK = KernelTraits<KID>::RANK;
  RuleView rv{};
  StdRule sr{};
  Consts cs;
  for (int k = 0; k < CFX_MAX_CONSTANTS; ++k)
    cs.c[k] = I.constants[k];
  OutCtx oc{c->mat_slot.p, f->Ae.p, base};
  if (KernelCoef<KID>::value)
    CFX_REQUIRE(f->coeff != nullptr, CFX_ERR_STATE,
                "the kernel reads a Function coefficient: call cfx_form_set_coefficient first");
  if (RANK == 0 && I.n > 0)
  { // only functionals materialise their standard entities (one value each)
    RuleTable& rt = get_rule(c, TDIM, std_rule_order(KID, DEG));
    sr = StdRule{rt.d_pts, rt.d_wts, rt.npts};
    auto k = cell_kernel<TDIM, DEG, KID, false>;
    CFX_LAUNCH(c, k, grid_for(I.n, EB), EB, 0, I.entities, dn_entities(I), rv, sr, cs, c->x, c->x_dofmap, oc, f->coeff,
               c->spaces[f->space].dofmap);
    base += I.n;
    oc.base = base;
  }
  if (I.rules && I.rules->nrules > 0)
  {
    const cfx_rules* R = I.rules;
    CFX_REQUIRE(R->tdim == TDIM, CFX_ERR_INVALID, "run-time rules have the wrong reference dimension");
    rv = make_rule_view(R, R->has_normals, R->has_moments,
                        1 + TDIM + (R->relation == CFX_REL_EQ ? TDIM * (TDIM + 1) / 2 : 0));
    auto k = cell_kernel<TDIM, DEG, KID, true>;
    CFX_LAUNCH(c, k, grid_for(R->nrules, EB), EB, 0, nullptr, dn_rules(R), rv, sr, cs, c->x, c->x_dofmap, oc, f->coeff,
               c->spaces[f->space].dofmap);
    if (base >= 0)
      base += R->nrules;
  }
}

// ================================================================== blocked (vector) Lagrange spaces
// demo_elasticity.py:213-238, test_assembly_elasticity.py: V = ("Lagrange", p, (gdim,)).  The dofmap and
// the sparsity pattern stay scalar (one row / column per dof), every CSR entry is a row-major bs x bs
// block (la::MatrixCSR::mat_add_values<BS, BS>, wrappers/fem.cpp:340-385), vectors hold bs values per dof
// (assemble_vector_impl.h:109-120), local index of (dof i, component a) = i*bs + a.  Parity-first path:
// every element tensor (standard cells too) is materialised, the generic row gather adds the blocks.
// 30 x 30 P2-vector tensors are where FP64 tensor-core MMA would be tried.

// elasticity tensor, one thread per (entity, tensor row ra = i*bs + a): the row's N = nd*bs entries
//   A[(i,a),(j,b)] = w ( lambda d_a phi_i d_b phi_j + mu d_b phi_i d_a phi_j + mu delta_ab grad phi_i . grad phi_j )
// c0 = mu, c1 = lambda.  Slots as in cell_kernel; the threads of one entity race benignly on mat_slot
// (they all store the same value) and recognise their own claim by its value.
template <int TDIM, int DEG, bool RUNTIME>
__global__ void __launch_bounds__(EB)
    elasticity_kernel(const int32_t* __restrict__ cells, DN n_, RuleView rv, StdRule sr, Consts cs,
                      const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, OutCtx oc)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (t >= n_.get() * N)
    return;
  const int64_t e = t / N;
  const int ra = static_cast<int>(t - e * N);
  const int i = ra / BS, a = ra - i * BS;
  const int64_t cell = RUNTIME ? rv.parent_map[e] : cells[e];
  const bool follow = oc.base < 0;
  const int64_t mine = oc.base + e;
  const int32_t s0 = oc.mat_slot[cell];
  const bool add = follow || (s0 >= 0 && s0 != mine);
  if (follow && s0 < 0)
    return;
  const int64_t slot = add ? s0 : mine;
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  double acc[N];
#pragma unroll
  for (int k = 0; k < N; ++k)
    acc[k] = 0.0;
  auto point = [&](const double (&xi)[TDIM], double w)
  {
    double phi[ND], dphi[ND][TDIM], grad[ND][TDIM];
    tabulate<TDIM, DEG>(xi, phi, dphi);
    push_gradients<TDIM, ND>(g, dphi, grad);
    double gi[TDIM];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double v = grad[0][r];
#pragma unroll
      for (int j = 1; j < ND; ++j)
        v = (j == i) ? grad[j][r] : v;
      gi[r] = v;
    }
    double gia = gi[0];
#pragma unroll
    for (int r = 1; r < TDIM; ++r)
      gia = (r == a) ? gi[r] : gia;
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      double gg = 0.0, gja = grad[j][0];
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
      {
        gg += gi[r] * grad[j][r];
        gja = (r == a) ? grad[j][r] : gja;
      }
#pragma unroll
      for (int b = 0; b < BS; ++b)
        acc[j * BS + b] += w * (cs.c[1] * gia * grad[j][b] + cs.c[0] * gi[b] * gja + ((a == b) ? cs.c[0] * gg : 0.0));
    }
  };
  if constexpr (RUNTIME)
  {
    const int32_t q0 = rv.offsets[e], q1 = rv.offsets[e + 1];
    for (int32_t q = q0; q < q1; ++q)
    {
      double xi[TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
        xi[tt] = rv.pts[static_cast<int64_t>(tt) * rv.stride() + q];
      point(xi, rv.wts[q]);
    }
  }
  else
  {
    const double sdet = fabs(g.detJ);
    for (int q = 0; q < sr.npts; ++q)
    {
      double xi[TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
        xi[tt] = __ldg(sr.pts + q * TDIM + tt);
      point(xi, __ldg(sr.wts + q) * sdet);
    }
  }
  if (!add)
    oc.mat_slot[cell] = static_cast<int32_t>(slot);
  double* p = oc.out + (slot * N + ra) * N;
#pragma unroll
  for (int k = 0; k < N; ++k)
    p[k] = add ? p[k] + acc[k] : acc[k];
}

// Symmetric Nitsche terms of linear elasticity on interface rules (per-point unit normals), one thread per
// (rule, tensor row ra = i*bs + a):
//   A[(i,a),(j,b)] = sum_q w ( -phi_i S_ab(j) - phi_j S_ba(i) + delta_ab pen phi_i phi_j ),
//   S_ab(j) = component a of sigma(phi_j e_b) n = mu (delta_ab grad phi_j . n + d_a phi_j n_b) + lambda d_b phi_j n_a,
//   pen = c2 (2 mu + lambda) / h;  c0 = mu, c1 = lambda, c2 = gamma.  Slots as in elasticity_kernel.
template <int TDIM, int DEG>
__global__ void __launch_bounds__(EB)
    nitsche_vec_kernel(DN n_, RuleView rv, Consts cs, const double* __restrict__ x,
                       const int32_t* __restrict__ x_dofmap, OutCtx oc)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (t >= n_.get() * N)
    return;
  const int64_t e = t / N;
  const int ra = static_cast<int>(t - e * N);
  const int i = ra / BS, a = ra - i * BS;
  const int64_t cell = rv.parent_map[e];
  const bool follow = oc.base < 0;
  const int64_t mine = oc.base + e;
  const int32_t s0 = oc.mat_slot[cell];
  const bool add = follow || (s0 >= 0 && s0 != mine);
  if (follow && s0 < 0)
    return;
  const int64_t slot = add ? s0 : mine;
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  const double pen = cs.c[2] * (2.0 * cs.c[0] + cs.c[1]) / cell_diameter<TDIM>(X);
  double acc[N];
#pragma unroll
  for (int k = 0; k < N; ++k)
    acc[k] = 0.0;
  const int32_t q0 = rv.offsets[e], q1 = rv.offsets[e + 1];
  for (int32_t q = q0; q < q1; ++q)
  {
    double xi[TDIM], nr[TDIM];
#pragma unroll
    for (int tt = 0; tt < TDIM; ++tt)
    {
      xi[tt] = rv.pts[static_cast<int64_t>(tt) * rv.stride() + q];
      nr[tt] = rv.nrm[static_cast<int64_t>(tt) * rv.stride() + q];
    }
    const double w = rv.wts[q];
    double phi[ND], dphi[ND][TDIM], grad[ND][TDIM];
    tabulate<TDIM, DEG>(xi, phi, dphi);
    push_gradients<TDIM, ND>(g, dphi, grad);
    // this row's basis function: value, gradient, normal derivative
    double gi[TDIM];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double v = grad[0][r];
#pragma unroll
      for (int j = 1; j < ND; ++j)
        v = (j == i) ? grad[j][r] : v;
      gi[r] = v;
    }
    const double phii = pick<ND>(phi, i);
    double gni = 0.0, na = nr[0], gia = gi[0];
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      gni += gi[r] * nr[r];
      na = (r == a) ? nr[r] : na;
      gia = (r == a) ? gi[r] : gia;
    }
#pragma unroll
    for (int j = 0; j < ND; ++j)
    {
      double gnj = 0.0, gja = grad[j][0];
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
      {
        gnj += grad[j][r] * nr[r];
        gja = (r == a) ? grad[j][r] : gja;
      }
#pragma unroll
      for (int b = 0; b < BS; ++b)
      {
        const double dab = (a == b) ? 1.0 : 0.0;
        const double S_ab_j = cs.c[0] * (dab * gnj + gja * nr[b]) + cs.c[1] * grad[j][b] * na;
        const double S_ba_i = cs.c[0] * (dab * gni + gi[b] * na) + cs.c[1] * gia * nr[b];
        acc[j * BS + b] += w * (-phii * S_ab_j - phi[j] * S_ba_i + dab * pen * phii * phi[j]);
      }
    }
  }
  if (!add)
    oc.mat_slot[cell] = static_cast<int32_t>(slot);
  double* p = oc.out + (slot * N + ra) * N;
#pragma unroll
  for (int k = 0; k < N; ++k)
    p[k] = add ? p[k] + acc[k] : acc[k];
}

// Run-time-rule tensors of blocked spaces, ONE WARP PER RULE (elasticity on volume rules, the symmetric Nitsche
// terms on interface rules).  The row-per-thread kernels above re-tabulate every basis function at every point in
// each of their nd*bs threads (ncu: 242 registers, FP64 pipe 57 % busy -- bound by redundant arithmetic).  Here:
//   phase 1  lane = quadrature point (chunks of 32): tabulate once, push the gradients once, stage phi, grad phi,
//            the normal and the weight in shared memory;
//   phase 2  lane = pair (i <= j) of basis functions (55 for P2 tetrahedra: two per lane): the bs x bs block
//            A[(i,.),(j,.)] accumulated over the staged points -- 2 bs + 2 shared loads per bs^2-block update;
//   output   the block and its mirror image (the tensors are symmetric: A[(j,b),(i,a)] = A[(i,a),(j,b)]).
// Same sums over the points in the same (ascending) order as the row-per-thread kernels, ~10x fewer flops.  This
// is the G^T W G contraction SURVEY.md section 8(d) reserves for FP64 tensor cores: on B200 the measured DMMA peak
// (37.0 TF/s) is within 9 % of the DFMA peak (34.1 TF/s, profiles/r5_fp64_peak.json) and after this restructuring
// the kernel is bound by writing the 7.2 KB tensor, so mma.sync.f64 has nothing to win here.
constexpr int BTW = 4; // warps (rules) per block

template <int TDIM, int DEG, bool NITSCHE>
__global__ void __launch_bounds__(BTW * 32)
    blocked_rule_tensor_kernel(DN n_, RuleView rv, Consts cs, const double* __restrict__ x,
                               const int32_t* __restrict__ x_dofmap, OutCtx oc)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS, B2 = BS * BS;
  constexpr int NPAIR = ND * (ND + 1) / 2, PPL = (NPAIR + 31) / 32;
  constexpr int PW = ND * (TDIM + 1) + TDIM + 1; // staged doubles per point: phi, grad, normal, weight
  __shared__ double s_pt[BTW][32][PW + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * BTW + w;
  if (e >= n_.get())
    return;
  const int64_t cell = rv.parent_map[e];
  const bool follow = oc.base < 0;
  const int64_t mine = oc.base + e;
  const int32_t s0 = oc.mat_slot[cell];
  const bool add = follow || (s0 >= 0 && s0 != mine);
  if (follow && s0 < 0)
    return;
  const int64_t slot = add ? s0 : mine;
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  const double mu = cs.c[0], lam = cs.c[1];
  double pen = 0.0;
  if constexpr (NITSCHE)
    pen = cs.c[2] * (2.0 * mu + lam) / cell_diameter<TDIM>(X);
  // this lane's pairs
  int pi[PPL], pj[PPL];
#pragma unroll
  for (int t = 0; t < PPL; ++t)
  {
    int p = lane + 32 * t, i = 0;
    if (p >= NPAIR)
      p = 0;
    while (p >= ND - i)
    {
      p -= ND - i;
      ++i;
    }
    pi[t] = i;
    pj[t] = i + p;
  }
  double acc[PPL][B2];
#pragma unroll
  for (int t = 0; t < PPL; ++t)
#pragma unroll
    for (int k = 0; k < B2; ++k)
      acc[t][k] = 0.0;
  const int32_t q0 = rv.offsets[e], q1 = rv.offsets[e + 1];
  for (int32_t qb = q0; qb < q1; qb += 32)
  {
    const int nq = min(32, q1 - qb);
    if (lane < nq)
    {
      const int32_t q = qb + lane;
      double xi[TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
        xi[tt] = rv.pts[static_cast<int64_t>(tt) * rv.stride() + q];
      double phi[ND], dphi[ND][TDIM], grad[ND][TDIM];
      tabulate<TDIM, DEG>(xi, phi, dphi);
      push_gradients<TDIM, ND>(g, dphi, grad);
      double* o = s_pt[w][lane];
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        o[i] = phi[i];
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
          o[ND + i * TDIM + r] = grad[i][r];
      }
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
        o[ND * (TDIM + 1) + r] = NITSCHE ? rv.nrm[static_cast<int64_t>(r) * rv.stride() + q] : 0.0;
      o[ND * (TDIM + 1) + TDIM] = rv.wts[q];
    }
    __syncwarp();
    for (int k = 0; k < nq; ++k)
    {
      const double* o = s_pt[w][k];
      const double wq = o[ND * (TDIM + 1) + TDIM];
      double nr[TDIM];
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
        nr[r] = o[ND * (TDIM + 1) + r];
#pragma unroll
      for (int t = 0; t < PPL; ++t)
      {
        const int i = pi[t], j = pj[t];
        double gi[TDIM], gj[TDIM];
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
        {
          gi[r] = o[ND + i * TDIM + r];
          gj[r] = o[ND + j * TDIM + r];
        }
        if constexpr (!NITSCHE)
        { // w ( lambda d_a phi_i d_b phi_j + mu d_b phi_i d_a phi_j + mu delta_ab grad phi_i . grad phi_j )
          double gg = 0.0;
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
            gg += gi[r] * gj[r];
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b)
              acc[t][a * BS + b] += wq * (lam * gi[a] * gj[b] + mu * gi[b] * gj[a] + ((a == b) ? mu * gg : 0.0));
        }
        else
        {
          const double phii = o[i], phij = o[j];
          double gni = 0.0, gnj = 0.0;
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
          {
            gni += gi[r] * nr[r];
            gnj += gj[r] * nr[r];
          }
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b)
            {
              const double dab = (a == b) ? 1.0 : 0.0;
              const double S_ab_j = mu * (dab * gnj + gj[a] * nr[b]) + lam * gj[b] * nr[a];
              const double S_ba_i = mu * (dab * gni + gi[b] * nr[a]) + lam * gi[a] * nr[b];
              acc[t][a * BS + b] += wq * (-phii * S_ab_j - phij * S_ba_i + dab * pen * phii * phij);
            }
        }
      }
    }
    __syncwarp();
  }
  if (!add && lane == 0)
    oc.mat_slot[cell] = static_cast<int32_t>(slot);
  double* out = oc.out + slot * N * N;
#pragma unroll
  for (int t = 0; t < PPL; ++t)
  {
    if (lane + 32 * t >= NPAIR)
      continue;
    const int i = pi[t], j = pj[t];
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b)
      {
        double* p0 = out + (i * BS + a) * N + j * BS + b;
        const double v = acc[t][a * BS + b];
        *p0 = add ? *p0 + v : v;
        if (i != j)
        {
          double* p1 = out + (j * BS + b) * N + i * BS + a;
          *p1 = add ? *p1 + v : v;
        }
      }
  }
}

// inner(f, v) dx with a constant vector f = (c0, c1, c2): one thread per entity, N values
template <int TDIM, int DEG, bool RUNTIME>
__global__ void __launch_bounds__(EB)
    source_vec_kernel(const int32_t* __restrict__ cells, DN n_, RuleView rv, StdRule sr, Consts cs,
                      const double* __restrict__ x, const int32_t* __restrict__ x_dofmap, OutCtx oc)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * EB + threadIdx.x;
  if (e >= n_.get())
    return;
  const int64_t cell = RUNTIME ? rv.parent_map[e] : cells[e];
  const bool follow = oc.base < 0;
  const int32_t s0 = oc.mat_slot[cell];
  const bool add = follow || s0 >= 0;
  if (follow && s0 < 0)
    return;
  const int64_t slot = add ? s0 : oc.base + e;
  double acc[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k)
    acc[k] = 0.0;
  if constexpr (RUNTIME)
  {
    const int32_t q0 = rv.offsets[e], q1 = rv.offsets[e + 1];
    for (int32_t q = q0; q < q1; ++q)
    {
      double xi[TDIM], phi[ND], dphi[ND][TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
        xi[tt] = rv.pts[static_cast<int64_t>(tt) * rv.stride() + q];
      tabulate<TDIM, DEG>(xi, phi, dphi);
#pragma unroll
      for (int k = 0; k < ND; ++k)
        acc[k] += rv.wts[q] * phi[k];
    }
  }
  else
  {
    double X[TDIM + 1][TDIM];
    load_cell_coords<TDIM>(x, x_dofmap, cell, X);
    Geo<TDIM> g;
    make_geo<TDIM>(X, g);
    const double sdet = fabs(g.detJ);
    for (int q = 0; q < sr.npts; ++q)
    {
      double xi[TDIM], phi[ND], dphi[ND][TDIM];
#pragma unroll
      for (int tt = 0; tt < TDIM; ++tt)
        xi[tt] = __ldg(sr.pts + q * TDIM + tt);
      tabulate<TDIM, DEG>(xi, phi, dphi);
#pragma unroll
      for (int k = 0; k < ND; ++k)
        acc[k] += __ldg(sr.wts + q) * sdet * phi[k];
    }
  }
  if (!add)
    oc.mat_slot[cell] = static_cast<int32_t>(slot);
  double* p = oc.out + slot * N;
#pragma unroll
  for (int k = 0; k < ND; ++k)
#pragma unroll
    for (int b = 0; b < BS; ++b)
    {
      const double v = cs.c[b] * acc[k];
      p[k * BS + b] = add ? p[k * BS + b] + v : v;
    }
}

// Generic row gather for blocked spaces: one warp per active row, one row per block.  Lanes stage the
// block row (bs x nd*bs values) of each incident cell's materialised tensor with the cell's dofs; the
// column lanes add the matching bs x bs blocks in ascending cell order (fixed order, no atomics).  A
// scalar interior-facet tensor (ghost penalty on jump(grad u, n)) acts on every component alike: its
// entry goes on the diagonal of the block.
template <int TDIM, int DEG>
__global__ void __launch_bounds__(32)
    gather_matrix_blocked_kernel(GatherCtx gc, const int32_t* __restrict__ act_rows, const int32_t* __restrict__ slots,
                                 DN n_act_, int facets_only, const int64_t* __restrict__ row_ptr,
                                 const int32_t* __restrict__ cols, double* __restrict__ vals, int zero_first,
                                 int32_t* __restrict__ err)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS;
  __shared__ int32_t s_dofs[32][ND];
  __shared__ double s_a[32][BS * N];
  __shared__ __align__(16) int32_t s_fd[32][FacetStage<ND>::W];
  __shared__ __align__(16) double s_fv[32][FacetStage<ND>::W];
  const int lane = threadIdx.x;
  if (static_cast<int64_t>(blockIdx.x) >= n_act_.get())
    return;
  // facets_only: the cell tensors were added by gather_matrix_blocked2_kernel; this pass walks the band rows
  // (slot list) and adds the interior-facet macro tensors
  const int64_t idx = slots ? slots[blockIdx.x] : blockIdx.x;
  const unsigned full = 0xffffffffu;
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  int matched = 0, expected = 0;
  for (int kc = 0; kc < rn; kc += 32)
  {
    const bool have_col = kc + lane < rn;
    const int32_t mycol = have_col ? cols[rb + kc + lane] : -2;
    double acc[BS * BS];
#pragma unroll
    for (int q = 0; q < BS * BS; ++q)
      acc[q] = (have_col && !zero_first) ? vals[(rb + kc + lane) * BS * BS + q] : 0.0;
    for (int k0 = 0; k0 < n_inc; k0 += 32)
    {
      const int k = k0 + lane;
      int64_t c = -1;
      unsigned fl = 0;
      int li = 0;
      if (k < n_inc)
      {
        c = gc.inc_cell[ib + k];
        fl = gc.cell_flags[c];
      }
      const bool contributes = (fl & 1u) != 0 && !facets_only;
      if (fl)
      {
        int32_t d[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j)
        {
          d[j] = gc.dofmap[c * ND + j];
          li = (d[j] == r) ? j : li;
        }
        if (contributes)
        {
          const double* a = gc.Ae + (static_cast<int64_t>(gc.mat_slot[c]) * N + li * BS) * N;
#pragma unroll
          for (int j = 0; j < ND; ++j)
            s_dofs[lane][j] = d[j];
          for (int q = 0; q < BS * N; ++q)
            s_a[lane][q] = a[q];
        }
      }
      if (!contributes)
      {
#pragma unroll
        for (int j = 0; j < ND; ++j)
          s_dofs[lane][j] = -1;
      }
      __syncwarp();
      const int nl = (n_inc - k0 < 32) ? n_inc - k0 : 32;
      for (int l = 0; l < nl; ++l)
      {
#pragma unroll
        for (int j = 0; j < ND; ++j)
          if (s_dofs[l][j] == mycol)
          {
#pragma unroll
            for (int a = 0; a < BS; ++a)
#pragma unroll
              for (int b = 0; b < BS; ++b)
                acc[a * BS + b] += s_a[l][a * N + j * BS + b];
            ++matched;
          }
      }
      if (kc == 0)
        expected += contributes ? ND : 0;
      double facc = 0.0;
      matched += add_facet_rows<ND>(gc, s_fd, s_fv, (fl & 2u) != 0, c, li, static_cast<int32_t>(r), mycol, facc, expected, kc == 0);
#pragma unroll
      for (int a = 0; a < BS; ++a)
        acc[a * BS + a] += facc;
      __syncwarp();
    }
    if (have_col)
    {
#pragma unroll
      for (int q = 0; q < BS * BS; ++q)
        vals[(rb + kc + lane) * BS * BS + q] = acc[q];
    }
  }
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

// Blocked spaces, cell tensors: one warp per row, one LANE PER (incident cell, local column dof j) pair -- 32 bs x bs
// blocks per round.  A standard-quadrature elasticity cell is NOT materialised (a P2 tetrahedron's 30 x 30 tensor is
// 7.2 KB): the lane computes its block from the cached geometry and the reference-element table
//   S^{rs} = int d_r phi_i d_s phi_j = |detJ| sum_tu K[t][r] K[u][s] R[t][u][i][j]          (81 FMAs)
//   block[a][b] = lambda S^{ab} + mu S^{ba} + delta_ab mu tr S
// which is the arithmetic the materialising kernel spends on the entry, without the 2 x 7.2 KB round trip through
// HBM; a cut cell's block is read from its materialised run-time-rule tensor.  The block goes to the CSR position of
// dof j (lower-bound search in the row's staged columns); lanes of one round that hit the same position add in lane
// order (match.any + rank loop), rounds run in order: ascending cells, fixed order, no atomics.  The row's blocks
// live in shared memory until the row is complete (one coalesced store).  Interior-facet tensors are added
// afterwards by gather_matrix_blocked_kernel in facets-only mode over the band rows.
constexpr int BGW = 2;      // rows (warps) per block
constexpr int BMAXC = 96;   // columns per row held in shared memory (P2 tetrahedra: 65 on a Kuhn mesh); longer rows
                            // accumulate in the CSR arrays themselves

template <int TDIM, int DEG>
__global__ void __launch_bounds__(BGW * 32, 16)
    gather_matrix_blocked2_kernel(GatherCtx gc, StdTab st, const int32_t* __restrict__ act_rows, DN n_act_,
                                  const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                  double* __restrict__ vals, int zero_first, int32_t* __restrict__ err)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS, B2 = BS * BS;
  using T = RefTab<TDIM, ND>;
  __shared__ double s_row[BGW][BMAXC * B2];
  __shared__ int32_t s_cols[BGW][BMAXC];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * BGW + w;
  if (idx >= n_act_.get())
    return;
  const unsigned full = 0xffffffffu;
  const uint32_t below = (1u << lane) - 1u;
  const int32_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  const int64_t rb = row_ptr[r];
  const int rn = static_cast<int>(row_ptr[r + 1] - rb);
  const bool staged = rn <= BMAXC;
  double* const row = staged ? s_row[w] : vals + rb * B2;
  const int32_t* const sc = staged ? s_cols[w] : cols + rb;
  if (staged)
  {
    for (int k = lane; k < rn; k += 32)
      s_cols[w][k] = cols[rb + k];
    for (int t = lane; t < rn * B2; t += 32)
      row[t] = zero_first ? 0.0 : vals[rb * B2 + t];
  }
  else if (zero_first)
    for (int t = lane; t < rn * B2; t += 32)
      row[t] = 0.0;
  __syncwarp();
  const int np = n_inc * ND;
  int bad = 0;
  for (int p0 = 0; p0 < np; p0 += 32)
  {
    const int p = p0 + lane;
    const bool valid = p < np;
    const int l = valid ? p / ND : 0;
    const int j = valid ? p - l * ND : 0;
    const int64_t c = gc.inc_cell[ib + l];
    const unsigned fl = valid ? gc.cell_flags[c] : 0u;
    const bool contributes = (fl & 0xFDu) != 0;
    int32_t dj = 0;
    int li = 0;
#pragma unroll
    for (int q = 0; q < ND; ++q)
    {
      const int32_t d = gc.dofmap[c * ND + q];
      li = (d == r) ? q : li;
      dj = (q == j) ? d : dj;
    }
    double blk[B2];
#pragma unroll
    for (int q = 0; q < B2; ++q)
      blk[q] = 0.0;
    if (fl >> 2)
    {
      Geo<TDIM> g;
      load_geo_cached<TDIM>(gc.geo, c, g);
      double S[TDIM][TDIM];
#pragma unroll
      for (int a = 0; a < TDIM; ++a)
#pragma unroll
        for (int b = 0; b < TDIM; ++b)
          S[a][b] = 0.0;
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
#pragma unroll
        for (int u = 0; u < TDIM; ++u)
        {
          const double rr = __ldg(st.ref + T::R + ((t * TDIM + u) * ND + li) * ND + j);
#pragma unroll
          for (int a = 0; a < TDIM; ++a)
          {
            const double ka = g.K[t * TDIM + a] * rr;
#pragma unroll
            for (int b = 0; b < TDIM; ++b)
              S[a][b] += ka * g.K[u * TDIM + b];
          }
        }
      double tr = 0.0;
#pragma unroll
      for (int a = 0; a < TDIM; ++a)
        tr += S[a][a];
      const double sdet = fabs(g.detJ);
      for (int k = 0; k < st.n; ++k)
      {
        if (!(fl & st.bit[k]))
          continue;
        const double mu = st.c[k][0] * sdet, lam = st.c[k][1] * sdet;
#pragma unroll
        for (int a = 0; a < TDIM; ++a)
#pragma unroll
          for (int b = 0; b < TDIM; ++b)
            blk[a * BS + b] += lam * S[a][b] + mu * S[b][a] + ((a == b) ? mu * tr : 0.0);
      }
    }
    if (fl & 1u)
    {
      const double* a = gc.Ae + (static_cast<int64_t>(__ldg(gc.mat_slot + c)) * N + li * BS) * N + j * BS;
#pragma unroll
      for (int aa = 0; aa < BS; ++aa)
#pragma unroll
        for (int b = 0; b < BS; ++b)
          blk[aa * BS + b] += a[aa * N + b];
    }
    // CSR position of dof j of the cell: lower bound in the row's ascending columns
    int pos = 0;
    if (contributes)
    {
      int lo = 0, hi = rn;
      while (lo < hi)
      {
        const int mid = (lo + hi) >> 1;
        if (sc[mid] < dj)
          lo = mid + 1;
        else
          hi = mid;
      }
      pos = lo;
      if (pos >= rn || sc[pos] != dj)
        bad = 1; // MatrixCSR::mat_add_values throws when an entry is not in the pattern
    }
    const int key = (contributes && !bad) ? pos : -1 - lane;
    const unsigned peers = __match_any_sync(full, key);
    const int rank = __popc(peers & below);
    const int maxrank = __reduce_max_sync(full, (contributes && !bad) ? rank : 0);
    for (int rr = 0; rr <= maxrank; ++rr)
    {
      if (contributes && !bad && rank == rr)
      {
#pragma unroll
        for (int q = 0; q < B2; ++q)
          row[pos * B2 + q] += blk[q];
      }
      __syncwarp();
    }
  }
  // ---- interior-facet macro tensors (full (2 nd)^2 scalar tensors; the P1 record format keeps the separate
  // facets-only pass): one lane per (incident band cell, local facet) probes the facet slot; every hit is then
  // spread over the lanes, lane e < 2 nd taking entry e of the hit's macro-tensor row -- its dof from the two cells'
  // dofmap rows, its position by the same search -- and adding it to the diagonal of the block (a scalar tensor
  // acts on every component alike).  Hits in (cell, local facet) order, duplicates of a round in lane order.
  if constexpr (DEG == 2)
  {
    if (gc.Fe != nullptr && gc.rows4 != nullptr)
    {
      constexpr int NF = TDIM + 1;
      const int nq = n_inc * NF;
      for (int p0 = 0; p0 < nq; p0 += 32)
      {
        const int p = p0 + lane;
        const bool valid = p < nq;
        const int l = valid ? p / NF : 0;
        const int lf = valid ? p - l * NF : 0;
        const int64_t c = gc.inc_cell[ib + l];
        const unsigned fl = valid ? gc.cell_flags[c] : 0u;
        int32_t fs = -1;
        if (fl & 2u)
          fs = gc.facet_slot[gc.c2f[c * NF + lf]];
        unsigned hits = __ballot_sync(full, fs >= 0);
        while (hits)
        {
          const int src = __ffs(hits) - 1;
          hits &= hits - 1;
          const int32_t hfs = __shfl_sync(full, fs, src);
          const int32_t hc = static_cast<int32_t>(__shfl_sync(full, static_cast<int32_t>(c), src));
          const int4 rw = __ldg(reinterpret_cast<const int4*>(gc.rows4) + hfs); // (cell0, lf0, cell1, lf1)
          const bool first = hc == rw.x;
          // local index of the row's dof in the hit's cell
          int hli = 0;
#pragma unroll
          for (int q = 0; q < ND; ++q)
            hli = (gc.dofmap[static_cast<int64_t>(hc) * ND + q] == r) ? q : hli;
          const int mrow = (first ? 0 : ND) + hli;
          const bool on = lane < 2 * ND;
          int32_t dj = 0;
          double v = 0.0;
          if (on)
          {
            const int64_t cc = lane < ND ? rw.x : rw.z;
            dj = gc.dofmap[cc * ND + (lane < ND ? lane : lane - ND)];
            v = gc.Fe[(static_cast<int64_t>(hfs) * 2 * ND + mrow) * 2 * ND + lane];
          }
          int pos = 0;
          bool ok = on;
          if (on)
          {
            int lo = 0, hi = rn;
            while (lo < hi)
            {
              const int mid = (lo + hi) >> 1;
              if (sc[mid] < dj)
                lo = mid + 1;
              else
                hi = mid;
            }
            pos = lo;
            if (pos >= rn || sc[pos] != dj)
            {
              bad = 1;
              ok = false;
            }
          }
          const unsigned peers = __match_any_sync(full, ok ? pos : -1 - lane);
          const int rank = __popc(peers & below);
          const int maxrank = __reduce_max_sync(full, ok ? rank : 0);
          for (int rr = 0; rr <= maxrank; ++rr)
          {
            if (ok && rank == rr)
            {
#pragma unroll
              for (int a = 0; a < BS; ++a)
                row[pos * B2 + a * BS + a] += v;
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (staged)
    for (int t = lane; t < rn * B2; t += 32)
      vals[rb * B2 + t] = row[t];
  if (__any_sync(full, bad) && lane == 0)
  {
    err[0] = 31;
    err[1] = r;
  }
}

// b[bs*dof + a] += element-vector entries of the incident cells, ascending cell order + fixed shuffle tree
template <int TDIM, int DEG>
__global__ void __launch_bounds__(GW * 32)
    gather_vector_blocked_kernel(GatherCtx gc, const int32_t* __restrict__ act_rows, DN n_act_,
                                 double* __restrict__ b, int zero_first)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  constexpr int BS = TDIM, N = ND * BS;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * GW + w;
  if (idx >= n_act_.get())
    return;
  const int64_t r = act_rows[idx];
  const int64_t ib = gc.inc_ptr[r];
  const int n_inc = static_cast<int>(gc.inc_ptr[r + 1] - ib);
  double s[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
    s[a] = 0.0;
  for (int k0 = 0; k0 < n_inc; k0 += 32)
  {
    const int k = k0 + lane;
    double e[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a)
      e[a] = 0.0;
    if (k < n_inc)
    {
      const int64_t c = gc.inc_cell[ib + k];
      if (gc.cell_flags[c] & 1u)
      {
        int li = 0;
#pragma unroll
        for (int j = 0; j < ND; ++j)
          li = (gc.dofmap[c * ND + j] == r) ? j : li;
        const double* p = gc.Ae + static_cast<int64_t>(gc.mat_slot[c]) * N + li * BS;
#pragma unroll
        for (int a = 0; a < BS; ++a)
          e[a] = p[a];
      }
    }
#pragma unroll
    for (int a = 0; a < BS; ++a)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        e[a] += __shfl_down_sync(0xffffffffu, e[a], o);
      s[a] += e[a];
    }
  }
  if (lane == 0)
  {
#pragma unroll
    for (int a = 0; a < BS; ++a)
      b[r * BS + a] = zero_first ? s[a] : b[r * BS + a] + s[a];
  }
}

template <int TDIM, int DEG>
void launch_blocked_cell(cfx_ctx* c, const cfx_integral& I, cfx_form* f, int64_t& base, bool fly)
{
  constexpr int N = Elem<TDIM, DEG>::ND * TDIM;
  RuleView rv{};
  StdRule sr{};
  Consts cs;
  for (int k = 0; k < CFX_MAX_CONSTANTS; ++k)
    cs.c[k] = I.constants[k];
  OutCtx oc{c->mat_slot.p, f->Ae.p, base};
  const bool el = I.kernel == CFX_K_ELASTICITY;
  if (I.kernel == CFX_K_NITSCHE_VEC)
  { // interface rules only (validated when the integral was added)
    const cfx_rules* R = I.rules;
    if (R && R->nrules > 0)
    {
      CFX_REQUIRE(R->tdim == TDIM && R->has_normals, CFX_ERR_INVALID, "interface kernels need rules with normals");
      rv = make_rule_view(R, true, false, 0);
      CFX_LAUNCH(c, (blocked_rule_tensor_kernel<TDIM, DEG, true>), grid_for(R->nrules, BTW), BTW * 32, 0, dn_rules(R), rv,
                 cs, c->x, c->x_dofmap, oc);
      if (base >= 0)
        base += R->nrules;
    }
    return;
  }
  CFX_REQUIRE(el || I.kernel == CFX_K_SOURCE_VEC, CFX_ERR_UNSUPPORTED,
              "kernel family is not defined on blocked (vector) spaces");
  if (I.n > 0 && !fly)
  {
    RuleTable& rt = get_rule(c, TDIM, el ? 2 * (DEG - 1) : DEG);
    sr = StdRule{rt.d_pts, rt.d_wts, rt.npts};
    if (el)
      CFX_LAUNCH(c, (elasticity_kernel<TDIM, DEG, false>), grid_for(I.n * N, EB), EB, 0, I.entities, dn_entities(I), rv,
                 sr, cs, c->x, c->x_dofmap, oc);
    else
      CFX_LAUNCH(c, (source_vec_kernel<TDIM, DEG, false>), grid_for(I.n, EB), EB, 0, I.entities, dn_entities(I), rv, sr,
                 cs, c->x, c->x_dofmap, oc);
    if (base >= 0)
    {
      base += I.n;
      oc.base = base;
    }
  }
  if (I.rules && I.rules->nrules > 0)
  {
    const cfx_rules* R = I.rules;
    CFX_REQUIRE(R->tdim == TDIM, CFX_ERR_INVALID, "run-time rules have the wrong reference dimension");
    rv = make_rule_view(R, false, false, 0);
    if (el)
      CFX_LAUNCH(c, (blocked_rule_tensor_kernel<TDIM, DEG, false>), grid_for(R->nrules, BTW), BTW * 32, 0, dn_rules(R), rv,
                 cs, c->x, c->x_dofmap, oc);
    else
      CFX_LAUNCH(c, (source_vec_kernel<TDIM, DEG, true>), grid_for(R->nrules, EB), EB, 0, nullptr, dn_rules(R), rv, sr, cs,
                 c->x, c->x_dofmap, oc);
    if (base >= 0)
      base += R->nrules;
  }
}

// one value per rule: c0 * (sum of the rule's weights) -- the measure functional on rules of any host dimension
__global__ void rule_measure_kernel(const int32_t* __restrict__ offsets, const double* __restrict__ weights,
                                    int64_t nrules, double c0, double* __restrict__ out)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (r >= nrules)
    return;
  double s = 0.0;
  for (int32_t q = offsets[r]; q < offsets[r + 1]; ++q)
    s += weights[q];
  out[r] = c0 * s;
}

template <int TDIM, int DEG>
void dispatch_cell(cfx_ctx* c, const cfx_integral& I, cfx_form* f, int64_t& base)
{
  switch (I.kernel)
  {
  case CFX_K_LAPLACE: launch_cell<TDIM, DEG, CFX_K_LAPLACE>(c, I, f, base); break;
  case CFX_K_MASS: launch_cell<TDIM, DEG, CFX_K_MASS>(c, I, f, base); break;
  case CFX_K_NITSCHE: launch_cell<TDIM, DEG, CFX_K_NITSCHE>(c, I, f, base); break;
  case CFX_K_SOURCE: launch_cell<TDIM, DEG, CFX_K_SOURCE>(c, I, f, base); break;
  case CFX_K_NITSCHE_RHS: launch_cell<TDIM, DEG, CFX_K_NITSCHE_RHS>(c, I, f, base); break;
  case CFX_K_ONE: launch_cell<TDIM, DEG, CFX_K_ONE>(c, I, f, base); break;
  case CFX_K_SQUARE_FN: launch_cell<TDIM, DEG, CFX_K_SQUARE_FN>(c, I, f, base); break;
  default: throw Error(CFX_ERR_UNSUPPORTED, "unknown cell kernel family");
  }
}

// materialise the run-time-rule tensors (ranks 1, 2) / all entity values (rank 0); returns the
// number of slots used
int64_t run_cell_integrals(cfx_ctx* c, cfx_form* f, int64_t follow_slots = -1)
{
  const Space& S = c->spaces[f->space];
  const bool blocked = S.bs > 1 && f->rank > 0;
  const bool fly = blocked_on_the_fly(S, f); // standard elasticity cells are evaluated by the row gather
  const int n1 = S.nd * (blocked ? S.bs : 1);
  const int es = f->rank == 2 ? n1 * n1 : (f->rank == 1 ? n1 : 1);
  auto dispatch = [&](const cfx_integral& I, int64_t& base)
  {
    if (I.rules && I.rules->entity_hosted)
    { // exterior-facet functional on facet-hosted rules (cfx_form_add_exterior_facet_integral): c0 * sum of weights
      if (I.rules->nrules > 0)
        CFX_LAUNCH(c, rule_measure_kernel, grid_for(I.rules->nrules, 256), 256, 0, I.rules->offsets.p, I.rules->weights.p,
                   I.rules->nrules, I.constants[0], f->Ae.p + base);
      base += I.rules->nrules;
      return;
    }
    if (blocked)
    {
      if (c->tdim == 2 && S.degree == 1)
        launch_blocked_cell<2, 1>(c, I, f, base, fly);
      else if (c->tdim == 2)
        launch_blocked_cell<2, 2>(c, I, f, base, fly);
      else if (S.degree == 1)
        launch_blocked_cell<3, 1>(c, I, f, base, fly);
      else
        launch_blocked_cell<3, 2>(c, I, f, base, fly);
    }
    else if (c->tdim == 2 && S.degree == 1)
      dispatch_cell<2, 1>(c, I, f, base);
    else if (c->tdim == 2)
      dispatch_cell<2, 2>(c, I, f, base);
    else if (S.degree == 1)
      dispatch_cell<3, 1>(c, I, f, base);
    else
      dispatch_cell<3, 2>(c, I, f, base);
  };
  if (follow_slots >= 0)
  { // FOLLOW mode: same slots as the form that ran before, zero-initialised accumulation
    f->Ae.reserve(c->pool, static_cast<size_t>(follow_slots) * es + 4);
    CFX_CUDA(cudaMemsetAsync(f->Ae.p, 0, (static_cast<size_t>(follow_slots) * es + 4) * sizeof(double), c->stream));
    for (auto& I : f->integrals)
    {
      if (I.facet)
        continue;
      int64_t base = -1;
      dispatch(I, base);
    }
    return follow_slots;
  }
  int64_t cap = 0;
  for (auto& I : f->integrals)
  {
    if (I.facet)
      continue;
    if (f->rank == 0 || (blocked && !fly))
      cap += I.n;
    if (I.rules)
      cap += I.rules->nrules;
  }
  f->Ae.reserve(c->pool, static_cast<size_t>(cap) * es + 4);
  int64_t base = 0;
  for (auto& I : f->integrals)
  {
    if (I.facet)
      continue;
    dispatch(I, base);
  }
  return base;
}

// reference-element tables (RefTab), built once per (tdim, degree) with an exact rule (mass: degree 2p)
const double* ensure_ref_table(cfx_ctx* c, int degree)
{
  auto key = std::make_pair(c->tdim, degree);
  auto it = c->ref_tabs.find(key);
  if (it == c->ref_tabs.end())
  {
    DevBuf<double> buf;
    RuleTable& rt = get_rule(c, c->tdim, 2 * degree);
    if (c->tdim == 2 && degree == 1)
    {
      buf.reserve(c->pool, RefTab<2, 3>::SIZE);
      CFX_LAUNCH(c, (ref_tables_kernel<2, 1>), 1, 32, 0, rt.d_pts, rt.d_wts, rt.npts, buf.p);
    }
    else if (c->tdim == 2)
    {
      buf.reserve(c->pool, RefTab<2, 6>::SIZE);
      CFX_LAUNCH(c, (ref_tables_kernel<2, 2>), 1, 64, 0, rt.d_pts, rt.d_wts, rt.npts, buf.p);
    }
    else if (degree == 1)
    {
      buf.reserve(c->pool, RefTab<3, 4>::SIZE);
      CFX_LAUNCH(c, (ref_tables_kernel<3, 1>), 1, 32, 0, rt.d_pts, rt.d_wts, rt.npts, buf.p);
    }
    else
    {
      buf.reserve(c->pool, RefTab<3, 10>::SIZE);
      CFX_LAUNCH(c, (ref_tables_kernel<3, 2>), 1, 128, 0, rt.d_pts, rt.d_wts, rt.npts, buf.p);
    }
    it = c->ref_tabs.emplace(key, buf).first;
  }
  return it->second.p;
}

void reset_slots(cfx_ctx* c, cfx_form* f)
{
  const bool blocked = c->spaces[f->space].bs > 1 && f->rank > 0 && !blocked_on_the_fly(c->spaces[f->space], f);
  for (auto& I : f->integrals)
  {
    if (I.facet)
      continue;
    if (I.rules && I.rules->nrules > 0)
      CFX_LAUNCH(c, reset_slots_kernel, grid_for(I.rules->nrules, 256), 256, 0, I.rules->parent_map.p, dn_rules(I.rules),
                 c->mat_slot.p);
    if (blocked && I.n > 0) // blocked spaces materialise the standard cells too
      CFX_LAUNCH(c, reset_slots_kernel, grid_for(I.n, 256), 256, 0, I.entities, dn_entities(I), c->mat_slot.p);
  }
}

StdTab make_std_tab(cfx_ctx* c, cfx_form* f)
{
  const Space& S = c->spaces[f->space];
  StdTab st{};
  if (S.bs > 1)
  {
    if (!blocked_on_the_fly(S, f))
      return st; // standard cells are materialised like cut cells
    // bilinear elasticity forms: (mu, lambda) per standard cell list + the reference-element table
    for (auto& I : f->integrals)
    {
      if (I.facet || I.n == 0)
        continue;
      CFX_REQUIRE(st.n < CFX_MAX_STD_LISTS, CFX_ERR_UNSUPPORTED, "too many standard cell integrals in one form");
      const int k = st.n++;
      st.kernel[k] = I.kernel;
      st.bit[k] = std_list_bit(f->prep, I.entities, I.n);
      st.c[k][0] = I.constants[0];
      st.c[k][1] = I.constants[1];
    }
    st.ref = ensure_ref_table(c, S.degree);
    return st;
  }
  for (auto& I : f->integrals)
  {
    if (I.facet || I.n == 0)
      continue;
    CFX_REQUIRE(st.n < CFX_MAX_STD_LISTS, CFX_ERR_UNSUPPORTED, "too many standard cell integrals in one form");
    const bool ok = f->rank == 2 ? (I.kernel == CFX_K_LAPLACE || I.kernel == CFX_K_MASS) : I.kernel == CFX_K_SOURCE;
    CFX_REQUIRE(ok, CFX_ERR_UNSUPPORTED, "kernel family has no standard-quadrature cell variant");
    RuleTable& rt = get_rule(c, c->tdim, std_rule_order(I.kernel, S.degree));
    const int k = st.n++;
    st.kernel[k] = I.kernel;
    st.bit[k] = std_list_bit(f->prep, I.entities, I.n);
    st.c[k][0] = I.constants[0];
    st.c[k][1] = I.constants[1];
    st.pts[k] = rt.d_pts;
    st.wts[k] = rt.d_wts;
    st.npts[k] = rt.npts;
    if (I.kernel == CFX_K_MASS)
      st.has_mass = 1;
  }
  if (S.degree == 2 && st.n > 0)
    st.ref = ensure_ref_table(c, S.degree);
  // P1 closed forms: coefficient sums per combination of list bits (ascending integral order)
  for (unsigned m = 0; m < (1u << CFX_MAX_STD_LISTS); ++m)
    for (int k = 0; k < st.n; ++k)
      if ((st.bit[k] >> 2) & m)
        (st.kernel[k] == CFX_K_MASS ? st.t1[m] : st.t0[m]) += st.c[k][0];
  return st;
}

template <int TDIM, int DEG>
void launch_facet(cfx_ctx* c, const cfx_integral& I, cfx_form* f, bool accumulate)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  Consts cs;
  for (int k = 0; k < CFX_MAX_CONSTANTS; ++k)
    cs.c[k] = I.constants[k];
  if constexpr (DEG == 1)
  { // factored, combined storage: one 80-byte record per facet (distinct dofs, combined jump coefficients, weight)
    static_assert(FREC <= 4 * ND * ND, "the facet buffer is sized for full macro tensors");
    CFX_LAUNCH(c, facet_p1_kernel<TDIM>, grid_for(I.n, EB), EB, 0, I.entities, dn_entities(I), cs, c->geo.p,
               c->spaces[f->space].dofmap, f->Fe.p, accumulate);
  }
  else
  {
    RuleTable& rt = get_rule(c, TDIM - 1, 2 * (DEG - 1));
    StdRule fr{rt.d_pts, rt.d_wts, rt.npts};
    auto k = facet_kernel<TDIM, DEG>;
    CFX_LAUNCH(c, k, grid_for(I.n * 2 * ND, EB), EB, 0, I.entities, dn_entities(I), fr, cs, c->x, c->x_dofmap, c->geo.p, f->Fe.p,
               accumulate);
  }
}

// Space::lrow: the record of every incidence (row, cell) from the geometry cache; one thread per incidence
template <int TDIM>
__global__ void __launch_bounds__(256)
    lrow_kernel(const double* __restrict__ geo, const int32_t* __restrict__ inc_cell, const uint32_t* __restrict__ fperm,
                int64_t n, double* __restrict__ lrow)
{
  constexpr int ND = TDIM + 1;
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n)
    return;
  const int64_t c = inc_cell[p];
  const int li = static_cast<int>(fperm[p] & 15u);
  Geo<TDIM> g;
  load_geo_cached<TDIM>(geo, c, g);
  const double s = fabs(g.detJ);
  double G[ND][TDIM];
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
  {
    double s0 = 0.0;
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
    {
      G[t + 1][r] = g.K[t * TDIM + r];
      s0 -= g.K[t * TDIM + r];
    }
    G[0][r] = s0;
  }
  const double w = s * (TDIM == 3 ? 1.0 / 6.0 : 0.5);
  double Gi[TDIM];
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
  {
    double t = G[0][r];
#pragma unroll
    for (int j = 1; j < ND; ++j)
      t = (j == li) ? G[j][r] : t;
    Gi[r] = t;
  }
  double rec[4] = {0.0, 0.0, 0.0, s};
#pragma unroll
  for (int q = 0; q < ND - 1; ++q)
  {
    // q-th off-diagonal column: j = q for q < li, q + 1 otherwise
    double d = 0.0;
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
    {
      double gj = G[0][r];
#pragma unroll
      for (int j = 1; j < ND; ++j)
        gj = (j == (q < li ? q : q + 1)) ? G[j][r] : gj;
      d += (Gi[r] * w) * gj;
    }
    rec[q] = d;
  }
  st256(lrow + p * 4, rec[0], rec[1], rec[2], rec[3]);
}

// Space::lrow for scalar P2 on triangles: per incidence (C00, C01, C11, |detJ|) with C = K K^T -- what the Laplace
// tensor row of any local dof is a linear combination of (reference tables T0, T1, T2 of gather_matrix_p2tri_kernel)
__global__ void __launch_bounds__(256)
    lrow_p2tri_kernel(const double* __restrict__ geo, const int32_t* __restrict__ inc_cell, int64_t n,
                      double* __restrict__ lrow)
{
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n)
    return;
  Geo<2> g;
  load_geo_cached<2>(geo, inc_cell[p], g);
  double C[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int t = 0; t < 2; ++t)
  {
    C[0] += g.K[0 * 2 + t] * g.K[0 * 2 + t];
    C[1] += g.K[0 * 2 + t] * g.K[1 * 2 + t];
    C[2] += g.K[1 * 2 + t] * g.K[1 * 2 + t];
  }
  st256(lrow + p * 4, C[0], C[1], C[2], fabs(g.detJ));
}

// Space::fpos64 (nd = 6): bits 0..2 local index li, bits 3..7 position of the row's own column, bits 8 + 5 q ..:
// position of the cell's q-th other dof (local order, li skipped)
__global__ void __launch_bounds__(256)
    fpos64_kernel(const uint32_t* __restrict__ fmask, const uint32_t* __restrict__ fperm, int nd, int64_t n,
                  uint64_t* __restrict__ fpos)
{
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n)
    return;
  const uint32_t fm = fmask[p], fp = fperm[p];
  const int li = static_cast<int>(fp & 15u);
  uint64_t w = static_cast<uint64_t>(li);
  int q = 0;
  for (int j = 0; j < nd; ++j)
  {
    const int rank = static_cast<int>((fp >> (4 + 4 * j)) & 15u);
    const uint64_t pos = __fns(fm, 0, rank + 1) & 31u;
    if (j == li)
      w |= pos << 3;
    else
      w |= pos << (8 + 5 * q++);
  }
  fpos[p] = w;
}

const uint64_t* ensure_fpos64(cfx_ctx* c, Space& S)
{
  if (!S.fpos_built)
  {
    S.fpos64.reserve(c->pool, static_cast<size_t>(S.n_inc) + 16);
    CFX_LAUNCH(c, fpos64_kernel, grid_for(S.n_inc, 256), 256, 0, S.fmask.p, S.fperm.p, S.nd, S.n_inc, S.fpos64.p);
    S.fpos_built = true;
  }
  return S.fpos64.p;
}

const double* ensure_lrow(cfx_ctx* c, Space& S)
{
  if (!S.lrow_built && S.degree == 2)
  {
    S.lrow.reserve(c->pool, static_cast<size_t>(S.n_inc) * 4 + 4);
    CFX_LAUNCH(c, lrow_p2tri_kernel, grid_for(S.n_inc, 256), 256, 0, c->geo.p, S.inc_cell.p, S.n_inc, S.lrow.p);
    S.lrow_built = true;
  }
  if (!S.lrow_built)
  {
    S.lrow.reserve(c->pool, static_cast<size_t>(S.n_inc) * 4 + 4);
    if (c->tdim == 3)
      CFX_LAUNCH(c, lrow_kernel<3>, grid_for(S.n_inc, 256), 256, 0, c->geo.p, S.inc_cell.p, S.fperm.p, S.n_inc, S.lrow.p);
    else
      CFX_LAUNCH(c, lrow_kernel<2>, grid_for(S.n_inc, 256), 256, 0, c->geo.p, S.inc_cell.p, S.fperm.p, S.n_inc, S.lrow.p);
    S.lrow_built = true;
  }
  return S.lrow.p;
}

// Space::fpos: per incidence (row r, cell) one word of positions in r's static full-mesh row --
//   bits 0..1 local index li of r in the cell, bits 2..6 position of r itself (the diagonal),
//   bits 7 + 5 q ..: position of the cell's q-th other dof (local order, li skipped)
// from the position mask (ascending dofs) and the rank of each local dof among the cell's dofs.
__global__ void __launch_bounds__(256)
    fpos_kernel(const uint32_t* __restrict__ fmask, const uint32_t* __restrict__ fperm, int nd, int64_t n,
                uint32_t* __restrict__ fpos)
{
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n)
    return;
  const uint32_t fm = fmask[p], fp = fperm[p];
  const int li = static_cast<int>(fp & 15u);
  uint32_t w = static_cast<uint32_t>(li);
  int q = 0;
  for (int j = 0; j < nd; ++j)
  {
    const int rank = static_cast<int>((fp >> (4 + 4 * j)) & 15u);
    const uint32_t pos = __fns(fm, 0, rank + 1) & 31u;
    if (j == li)
      w |= pos << 2;
    else
      w |= pos << (7 + 5 * q++);
  }
  fpos[p] = w;
}

const uint32_t* ensure_fpos(cfx_ctx* c, Space& S)
{
  if (!S.fpos_built)
  {
    S.fpos.reserve(c->pool, static_cast<size_t>(S.n_inc) + 16);
    CFX_LAUNCH(c, fpos_kernel, grid_for(S.n_inc, 256), 256, 0, S.fmask.p, S.fperm.p, S.nd, S.n_inc, S.fpos.p);
    S.fpos_built = true;
  }
  return S.fpos.p;
}

GatherCtx make_gather_ctx(cfx_ctx* c, cfx_form* f, const cfx_integral* FI)
{
  const Space& S = c->spaces[f->space];
  GatherCtx g{};
  g.inc_ptr = S.inc_ptr.p;
  g.inc_cell = S.inc_cell.p;
  g.fperm = S.fperm.p;
  g.fmask = S.fmask.p;
  g.dofmap = S.dofmap;
  g.cell_flags = f->prep->cell_flags.p;
  g.mat_slot = c->mat_slot.p;
  g.Ae = f->Ae.p;
  g.geo = c->geo.p;
  const bool p1s = S.degree == 1 && S.bs == 1 && S.has_perm;
  const bool p2tri = S.degree == 2 && S.bs == 1 && S.has_perm && S.has_static && c->tdim == 2 && f->rank == 2;
  g.lrow = (p1s || p2tri) ? ensure_lrow(c, c->spaces[f->space]) : nullptr;
  g.fpos = (p1s && S.has_static) ? ensure_fpos(c, c->spaces[f->space]) : nullptr;
  g.fpos64 = p2tri ? ensure_fpos64(c, c->spaces[f->space]) : nullptr;
  g.frow_ptr = S.frow_ptr.p;
  g.fclist = S.fclist.p;
  g.c2f = c->c2f;
  g.facet_slot = c->facet_slot.p;
  g.rows4 = FI ? FI->entities : nullptr;
  g.Fe = f->Fe.p;
  g.Fw = (FI && S.degree == 1) ? f->Fe.p : nullptr; // non-null marks the P1 record format of facet_p1_kernel
  g.nf = c->tdim + 1;
  g.stride = S.stride;
  return g;
}

template <int TDIM, int DEG>
void launch_gather_matrix(cfx_ctx* ctx, cfx_form* a, cfx_pattern* A, const GatherCtx& gc, const StdTab& st,
                          const StdTab& stL, int zero_first)
{
  const Space& S = ctx->spaces[a->space];
  cfx_prepared* PR = a->prep;
  if (PR->n_act_rows == 0)
    return;
  if (S.bs > 1)
  {
    CFX_REQUIRE(S.bs == TDIM, CFX_ERR_UNSUPPORTED, "blocked spaces need block size == gdim");
    auto kb = gather_matrix_blocked_kernel<TDIM, DEG>;
    auto k2 = gather_matrix_blocked2_kernel<TDIM, DEG>;
    const cfx_integral* FI = facet_integral_domain(a);
    {
      // FP64-bound: 2 * (81 + 27) flops per bs x bs block of a standard cell, (nd bs)^2 / bs^2 blocks per cell;
      // bytes: 12 B + 8 bs^2 B per CSR entry + the geometry record and dofmap row per (row, cell) pair
      int64_t n_std = 0;
      for (auto& I : a->integrals)
        if (!I.facet)
          n_std += I.n;
      StageScope sk(ctx, "gather_matrix_blocked2_kernel",
                    (4.0 + 8.0 * S.bs * S.bs) * static_cast<double>(A->nnz)
                        + (96.0 + 4.0 * S.nd) * S.nd * static_cast<double>(n_std));
      CFX_LAUNCH(ctx, k2, grid_for(PR->n_act_rows, BGW), BGW * 32, 0, gc, st, PR->act_rows.p, PR->dn_act(),
                 A->row_ptr.p, A->cols.p, A->values.p, zero_first, ctx->err_flag.p);
    }
    if (FI && PR->n_band > 0 && DEG == 1)
    { // P1 facet records: the separate facets-only pass over the band rows
      StageScope sk(ctx, "gather_matrix_blocked_facets_kernel", 0.0);
      CFX_LAUNCH(ctx, kb, grid_for(PR->n_band, 1), 32, 0, gc, PR->act_rows.p, PR->band_idx.p, PR->dn_band(), 1,
                 A->row_ptr.p, A->cols.p, A->values.p, 0, ctx->err_flag.p);
    }
    return;
  }
  // the gather tables are valid only for the pattern that was built from this very form
  const bool fast = a->gtab_serial == A->serial && a->gtab_serial > 0 && S.has_perm;
  const unsigned g = grid_for(PR->n_act_rows, GW);
  if constexpr (Elem<TDIM, DEG>::ND <= 6)
  { // the mask / contribution-list kernels need the packed incidence tables (nd <= 6)
    if (fast && S.has_static && a->n_clist_rows > 0)
    {
      // static rows and band rows are disjoint: this kernel runs on a lane of its own, concurrently with the band
      // kernel below (joined at the end of this function)
      LaneScope lane(ctx, (a->n_mask_rows != 0) ? 1 : 0);
      // the dominant kernel, timed on its own: fused K4 (every standard cell's dofmap row, coordinates and
      // dofs: 4 nv + 24 nv + 4 nd B) + K5 (12 B per CSR entry of its rows) -- SURVEY.md section 8(d)
      int64_t n_std = 0;
      for (auto& I : a->integrals)
        if (!I.facet)
          n_std += I.n;
      // (the stage carries the name of the kernel that runs: scalar P1 -> gather_matrix_p1_kernel)
      const bool p1_rows = Elem<TDIM, DEG>::ND <= 4 && gc.fpos != nullptr;
      const bool p2_rows = DEG == 2 && TDIM == 2 && gc.fpos64 != nullptr && st.ref != nullptr && !getenv("CFX_OLD_CLIST");
      StageScope sk(ctx, p1_rows ? "gather_matrix_p1_kernel" : (p2_rows ? "gather_matrix_p2tri_kernel" : "gather_matrix_clist_kernel"),
                    12.0 * static_cast<double>(a->n_clist_nnz)
                        + (28.0 * ctx->nv + 4.0 * S.nd) * static_cast<double>(n_std));
      bool done = false;
      if constexpr (DEG == 1)
      {
        if (gc.fpos != nullptr)
        { // scalar P1: one thread per row
          const bool nc16 = S.max_fcols > 0 && S.max_fcols <= 16;
          // measured at C3 (profiles/README, r9i): 32 accumulator rows / 5 blocks per SM 0.674 ms; 16 rows 0.642 ms;
          // 16 rows and 6 blocks per SM (80 registers, 12 bytes spilled) 0.635 ms; 7 / 8 blocks (72 / 64 registers,
          // spills in the record pipeline) 0.707 / 0.744 ms
          auto kp = gc.bvec ? gather_matrix_p1_kernel<TDIM, true, 32, 1> : gather_matrix_p1_kernel<TDIM, false, 32, 1>;
          if (nc16)
            kp = gc.bvec ? gather_matrix_p1_kernel<TDIM, true, 16, 6> : gather_matrix_p1_kernel<TDIM, false, 16, 6>;
          CFX_LAUNCH(ctx, kp, grid_for(PR->n_act_rows, RTB), RTB, 0, gc, st, stL, PR->act_rows.p, PR->dn_act(),
                     a->row_fast.p, a->Rrow.p, a->row_ufl.p, gc.fpos, A->row_ptr.p, A->values.p, zero_first);
          done = true;
        }
      }
      if constexpr (DEG == 2 && TDIM == 2)
      {
        static const bool old_clist = getenv("CFX_OLD_CLIST") != nullptr; // A/B switch: the warp-per-row kernel
        if (gc.fpos64 != nullptr && st.ref != nullptr && !old_clist)
        { // scalar P2 on triangles: one thread per row
          auto kp = gc.bvec ? gather_matrix_p2tri_kernel<true> : gather_matrix_p2tri_kernel<false>;
          CFX_LAUNCH(ctx, kp, grid_for(PR->n_act_rows, RTB), RTB, 0, gc, st, stL, PR->act_rows.p, PR->dn_act(),
                     a->row_fast.p, a->Rrow.p, a->row_ufl.p, gc.fpos64, A->row_ptr.p, A->values.p, zero_first);
          done = true;
        }
      }
      if (!done)
      {
        auto kc = gc.bvec ? gather_matrix_clist_kernel<TDIM, DEG, true> : gather_matrix_clist_kernel<TDIM, DEG, false>;
        // persistent: 4 blocks per SM walk the rows grid-stride through the software pipeline
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
        const unsigned gp = static_cast<unsigned>(std::min<int64_t>(
            static_cast<int64_t>(n_sm) * clist_blocks_per_sm<DEG>(), (PR->n_act_rows + GWC - 1) / GWC));
        CFX_LAUNCH(ctx, kc, gp, GWC * 32, 0, gc, st, stL, PR->act_rows.p, PR->dn_act(), a->row_fast.p, a->Rrow.p,
                   a->row_ufl.p, A->row_ptr.p, A->values.p, zero_first);
      }
    }
    if (fast && (a->n_mask_rows != 0))
    {
      // band rows (facet macro rows) and rows with long contribution lists: cell rows through the position
      // masks of the pattern pass, macro-tensor rows matched by column value
      // K5 for the rows the contribution-list kernel does not own: 12 B per CSR entry of the active rows left
      const double nnz_mask = static_cast<double>(A->nnz) - static_cast<double>(a->n_clist_nnz)
                              - static_cast<double>(A->n_rows - PR->n_act_rows);
      const cfx_integral* FIb = facet_integral_domain(a);
      // band rows, honestly counted: 12 B per CSR entry + per incident cell the 32-byte tensor-row record, the
      // position word, the cell id and flag byte (41 B) + per band cell its facet ids and slots (32 B) + the 80-byte
      // record of every band facet once per row that meets it (~ nd + 1 rows per facet)
      double band_bytes = 12.0 * (nnz_mask > 0.0 ? nnz_mask : 0.0);
      if (FIb)
        band_bytes += 80.0 * (S.nd + 1) * static_cast<double>(FIb->n) + (41.0 + 32.0) * S.stride * static_cast<double>(a->n_band_listed);
      const bool p1_band = Elem<TDIM, DEG>::ND <= 4 && gc.fpos != nullptr && a->n_band_listed > 0
                           && getenv("CFX_OLD_BAND") == nullptr;
      StageScope sk(ctx, p1_band ? "gather_matrix_band_p1_kernel" : "gather_matrix_fast_kernel", band_bytes);
      auto kf = gather_matrix_fast_kernel<TDIM, DEG>;
      bool band_done = false;
      if constexpr (DEG == 1)
      {
        static const bool old_band = getenv("CFX_OLD_BAND") != nullptr; // A/B switch: the warp-per-row mask kernel
        if (a->n_band_listed > 0 && gc.fpos != nullptr && !old_band)
        { // scalar P1 with a static structure: G threads per band row
          constexpr int Gu = 4;
          auto kb = gc.bvec ? gather_matrix_band_p1_kernel<TDIM, true, 4> : gather_matrix_band_p1_kernel<TDIM, false, 4>;
          CFX_LAUNCH(ctx, kb, grid_for(a->n_band_listed * Gu, RTBB), RTBB, 0, gc, st, stL, PR->act_rows.p,
                     PR->band_idx.p, PR->dn_band(), a->row_fast.p, S.fcols.p, A->row_ptr.p, A->cols.p, A->values.p,
                     zero_first);
          band_done = true;
        }
      }
      if (a->n_band_listed > 0 && !band_done)
        CFX_LAUNCH(ctx, kf, grid_for(a->n_band_listed, GWM), GWM * 32, 0, gc, st, stL, PR->act_rows.p, PR->band_idx.p,
                   PR->dn_band(), a->row_fast.p, a->gmask.p, a->Rrow.p, A->row_ptr.p, A->cols.p, A->values.p,
                   zero_first);
      // static rows without a contribution list (an edge with more than 8 cells), or no slot list at all
      // (a deferred-size step takes the decision from what eager steps saw; build_pattern verified it on the device)
      const bool noclist = A->deferred ? !a->expect_noclist_zero
                                       : PR->n_act_rows - a->n_band_listed - a->n_clist_rows > 0;
      if (a->n_band_listed == 0 || noclist)
        CFX_LAUNCH(ctx, kf, grid_for(PR->n_act_rows, GWM), GWM * 32, 0, gc, st, stL, PR->act_rows.p, nullptr,
                   PR->dn_act(), a->row_fast.p, a->gmask.p, a->Rrow.p, A->row_ptr.p, A->cols.p, A->values.p,
                   zero_first);
    }
  }
  if (lanes_enabled(ctx))
    lane_join(ctx);
  if (!fast || a->n_slow_rows > 0)
  {
    auto k = gather_matrix_kernel<TDIM, DEG>;
    CFX_LAUNCH(ctx, k, g, GW * 32, 0, gc, st, PR->act_rows.p, PR->dn_act(), fast ? a->row_fast.p : nullptr,
               A->row_ptr.p, A->cols.p, A->values.p, zero_first, ctx->err_flag.p);
  }
}

template <int TDIM, int DEG>
void launch_gather_vector(cfx_ctx* ctx, cfx_form* L, const GatherCtx& gc, const StdTab& st, double* d_b,
                          int zero_first, const cfx_form* fused_with = nullptr)
{
  const Space& S = ctx->spaces[L->space];
  cfx_prepared* PR = L->prep;
  if (PR->n_act_rows == 0)
    return;
  if (S.bs > 1)
  {
    auto kb = gather_vector_blocked_kernel<TDIM, DEG>;
    CFX_LAUNCH(ctx, kb, grid_for(PR->n_act_rows, GW), GW * 32, 0, gc, PR->act_rows.p, PR->dn_act(), d_b, zero_first);
    return;
  }
  auto run = [&](const int32_t* slots, int64_t n_bound, const uint8_t* row_fast, int skip_mode)
  {
    const unsigned g = grid_for(n_bound, GW);
    const DN n = PR->dn_act();
    if (S.has_perm)
    {
      auto k = gather_vector_kernel<TDIM, DEG, true>;
      CFX_LAUNCH(ctx, k, g, GW * 32, 0, gc, st, PR->act_rows.p, n, d_b, zero_first, slots, row_fast, skip_mode);
    }
    else
    {
      auto k = gather_vector_kernel<TDIM, DEG, false>;
      CFX_LAUNCH(ctx, k, g, GW * 32, 0, gc, st, PR->act_rows.p, n, d_b, zero_first, slots, row_fast, skip_mode);
    }
  };
  if (!fused_with)
  {
    run(nullptr, PR->n_act_rows, nullptr, 0);
    return;
  }
  // the matrix gather kernels (contribution-list and mask kernels) fill b for every fast row; what is left are
  // the rows of the generic kernel
  const cfx_form* a = fused_with;
  if (a->n_slow_rows > 0)
    run(nullptr, PR->n_act_rows, a->row_fast.p, 3);
}

#define CFX_DISPATCH_ELEM(ctx, S, FN, ...)                                                                             \
  do                                                                                                                   \
  {                                                                                                                    \
    if ((ctx)->tdim == 2 && (S).degree == 1)                                                                           \
      FN<2, 1>(__VA_ARGS__);                                                                                           \
    else if ((ctx)->tdim == 2)                                                                                         \
      FN<2, 2>(__VA_ARGS__);                                                                                           \
    else if ((S).degree == 1)                                                                                          \
      FN<3, 1>(__VA_ARGS__);                                                                                           \
    else                                                                                                               \
      FN<3, 2>(__VA_ARGS__);                                                                                           \
  } while (0)
template <int TDIM>
__global__ void __launch_bounds__(256) geo_cache_kernel(const double* __restrict__ x, const int32_t* __restrict__ x_dofmap,
                                                        int64_t n, double* __restrict__ geo)
{
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (c >= n)
    return;
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, c, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  double* o = geo + c * GeoRec<TDIM>::STRIDE;
#pragma unroll
  for (int i = 0; i < TDIM * TDIM; ++i)
    o[i] = g.K[i];
  o[TDIM * TDIM] = g.detJ;
  o[TDIM * TDIM + 1] = cell_diameter<TDIM>(X);
#pragma unroll
  for (int i = TDIM * TDIM + 2; i < GeoRec<TDIM>::STRIDE; ++i)
    o[i] = 0.0;
}
} // namespace

// K = J^-1, detJ and the cell diameter of every local cell: static while the mesh is bound
void build_geometry_cache(cfx_ctx* c)
{
  const int stride = c->tdim == 3 ? GeoRec<3>::STRIDE : GeoRec<2>::STRIDE;
  c->geo.reserve(c->pool, static_cast<size_t>(c->nc_total) * stride + 4);
  if (c->tdim == 3)
    CFX_LAUNCH(c, geo_cache_kernel<3>, grid_for(c->nc_total, 256), 256, 0, c->x, c->x_dofmap, c->nc_total, c->geo.p);
  else
    CFX_LAUNCH(c, geo_cache_kernel<2>, grid_for(c->nc_total, 256), 256, 0, c->x, c->x_dofmap, c->nc_total, c->geo.p);
}
} // namespace cfx

using namespace cfx;


// bytes a cell kernel reads from an integral's run-time rules: every point (coordinates, weight, normal), or --
// P1 kernels served by the rule moments (cell_kernel, LINEAR) -- tdim + 1 doubles per rule
static double rule_read_bytes(const cfx_ctx* ctx, const Space& S, const cfx_integral& I)
{
  const bool ifc = I.rules->relation == CFX_REL_EQ;
  const bool linear = S.degree == 1 && S.bs == 1 && I.rules->has_moments
                      && (I.kernel == CFX_K_LAPLACE || I.kernel == CFX_K_SOURCE || I.kernel == CFX_K_ONE
                          || (ifc && (I.kernel == CFX_K_NITSCHE || I.kernel == CFX_K_NITSCHE_RHS)));
  if (linear)
    return 8.0 * (1 + ctx->tdim + (ifc ? ctx->tdim * (ctx->tdim + 1) / 2 + ctx->tdim : 0))
           * static_cast<double>(I.rules->nrules);
  return 8.0 * (ctx->tdim + 1 + (I.rules->has_normals ? ctx->tdim : 0)) * static_cast<double>(I.rules->npts);
}

// assemble_matrix, optionally fused with the right-hand side of a linear form over the same prepared
// domains (L != null): its run-time-rule entries are materialised into the bilinear form's slots and the
// contribution-list kernel fills b for its rows in the same pass over the incidence.
static void assemble_matrix_impl(cfx_ctx* ctx, cfx_form* a, cfx_pattern* A, int zero_first, double diag_inactive,
                                 cfx_form* L, double* d_b, int zero_first_b)
{
  // adding into a matrix that is known to be zero == overwriting it: no second memset, no read of the old values
  // a fresh pattern whose active rows were not zeroed (cfx_pattern::values_lazy): fine when this form is the one the
  // pattern was built from -- the gathers below then overwrite every entry of every active row; otherwise zero now
  if (A->values_lazy && !(a->gtab_serial == A->serial && a->gtab_serial > 0))
    settle_values(ctx, A);
  A->values_lazy = false;
  const bool fresh = A->values_zero;
  A->values_zero = false;
  if (fresh)
    zero_first = 1;
  const Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  const cfx_integral* FI = facet_integral_domain(a);
  const int nd = S.nd;
  int64_t n_mat = 0, n_std = 0;
  for (auto& I : a->integrals)
    if (!I.facet)
      n_std += I.n;
  if (FI)
  { // the facet tensors do not depend on the cell tensors: on a lane of their own
    a->Fe.reserve(ctx->pool, static_cast<size_t>(FI->n) * 4 * nd * nd + 1);
    LaneScope lane(ctx, 2);
    StageScope st(ctx, "element_facets",
                  static_cast<double>(FI->n) * (16.0 + (S.degree == 1 ? 8.0 * (2.0 * nd + 1.0) : 8.0 * 4.0 * nd * nd)));
    bool acc = false;
    for (auto& I : a->integrals)
    {
      if (!I.facet || I.n == 0)
        continue;
      CFX_DISPATCH_ELEM(ctx, S, launch_facet, ctx, I, a, acc);
      acc = true;
    }
  }
  {
    StageScope st(ctx, "element_cells");
    n_mat = run_cell_integrals(ctx, a);
    double by = 0.0;
    for (auto& I : a->integrals)
      if (!I.facet && I.rules)
        by += rule_read_bytes(ctx, S, I)
              + (8.0 + 28.0 * ctx->nv + 8.0 * nd * nd) * static_cast<double>(I.rules->nrules);
    st.set_bytes(by);
  }
  if (L)
  {
    StageScope st(ctx, "element_cells_vector");
    run_cell_integrals(ctx, L, n_mat);
    double by = 0.0;
    for (auto& I : L->integrals)
      if (!I.facet && I.rules)
        by += rule_read_bytes(ctx, S, I)
              + (8.0 + 28.0 * ctx->nv + 8.0 * nd) * static_cast<double>(I.rules->nrules);
    st.set_bytes(by);
  }
  if (lanes_enabled(ctx))
    lane_join(ctx);
  {
    // fused K4 (standard cells: dofmap row + coordinates + dofs in) + K5 (each CSR value and column once)
    StageScope st(ctx, "gather_matrix",
                  12.0 * static_cast<double>(A->nnz) + (28.0 * ctx->nv + 4.0 * nd) * static_cast<double>(n_std)
                      + 8.0 * nd * nd * static_cast<double>(n_mat));
    set_facet_slots(ctx, FI, false);
    GatherCtx gc = make_gather_ctx(ctx, a, FI);
    const StdTab stt = make_std_tab(ctx, a);
    StdTab stL{};
    const bool fuse_b = L && S.bs == 1 && a->gtab_serial == A->serial && a->gtab_serial > 0 && S.has_perm
                        && S.has_static && a->n_clist_rows > 0 && S.nd <= 6;
    if (fuse_b)
    {
      stL = make_std_tab(ctx, L);
      gc.AeL = L->Ae.p;
      gc.bvec = d_b;
      gc.zero_first_b = zero_first_b;
    }
    if (zero_first && !fresh)
      CFX_CUDA(cudaMemsetAsync(A->values.p, 0, static_cast<size_t>(A->nnz) * S.bs * S.bs * sizeof(double), ctx->stream));
    if (diag_inactive != 0.0)
      CFX_LAUNCH(ctx, inactive_diag_kernel, grid_for(A->n_rows, 256), 256, 0, a->prep->row_flag.p, A->n_rows, A->row_ptr.p,
                 A->cols.p, A->values.p, diag_inactive, S.bs, static_cast<int64_t>(A->cols.cap) - 1);
    CFX_DISPATCH_ELEM(ctx, S, launch_gather_matrix, ctx, a, A, gc, stt, stL, zero_first);
    set_facet_slots(ctx, FI, true);
  }
  if (L)
  {
    // rows the contribution-list kernel did not own (band rows, generic rows); inactive rows keep 0 / b
    const bool fused = S.bs == 1 && a->gtab_serial == A->serial && a->gtab_serial > 0 && S.has_perm && S.has_static
                       && a->n_clist_rows > 0 && S.nd <= 6; // same condition as fuse_b above
    // fused: the matrix gather kernels have written b for every fast row; only generic rows are left
    StageScope st(ctx, "gather_vector", fused ? 8.0 * static_cast<double>(a->n_slow_rows)
                                              : 8.0 * static_cast<double>(S.n_total));
    GatherCtx gl = make_gather_ctx(ctx, L, nullptr);
    const StdTab stl = make_std_tab(ctx, L);
    CFX_DISPATCH_ELEM(ctx, S, launch_gather_vector, ctx, L, gl, stl, d_b, zero_first_b, fused ? a : nullptr);
  }
  reset_slots(ctx, a);
}

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
  assemble_matrix_impl(ctx, a, A, zero_first, diag_inactive, nullptr, nullptr, 0);
  if (values_out)
    export_to(ctx, values_out, A->values.p, static_cast<size_t>(A->nnz) * A->bs * A->bs, memspace);
  check_call(ctx, "cfx_assemble_matrix (entry not in sparsity pattern)");
  CFX_API_END(ctx)
}

cfx_status cfx_assemble_system(cfx_ctx* ctx, const cfx_form* a_const, cfx_pattern* A, int zero_first_A,
                               double diag_inactive, const cfx_form* L_const, double* b, int zero_first_b)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  cfx_form* L = const_cast<cfx_form*>(L_const);
  CFX_REQUIRE(ctx && a && A && L && b, CFX_ERR_INVALID, "cfx_assemble_system: NULL argument");
  CFX_REQUIRE(a->rank == 2 && L->rank == 1, CFX_ERR_INVALID, "cfx_assemble_system: needs a bilinear and a linear form");
  CFX_REQUIRE(A->space == a->space && L->space == a->space, CFX_ERR_INVALID,
              "cfx_assemble_system: forms and matrix must use the same space");
  const Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  prepare_form(ctx, L);
  if (zero_first_b)
    CFX_CUDA(cudaMemsetAsync(b, 0, static_cast<size_t>(S.n_total) * S.bs * sizeof(double), ctx->stream));
  if (a->prep == L->prep)
    assemble_matrix_impl(ctx, a, A, zero_first_A, diag_inactive, L, b, zero_first_b);
  else
  { // different integration domains: nothing to share, two passes
    assemble_matrix_impl(ctx, a, A, zero_first_A, diag_inactive, nullptr, nullptr, 0);
    cfx_status rc = cfx_assemble_vector(ctx, L, b, zero_first_b, CFX_DEVICE);
    if (rc != CFX_OK)
      return rc;
  }
  check_call(ctx, "cfx_assemble_system (entry not in sparsity pattern)");
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
  int64_t n_mat = 0, n_std = 0;
  for (auto& I : L->integrals)
    if (!I.facet)
      n_std += I.n;
  {
    StageScope st(ctx, "element_cells_vector");
    n_mat = run_cell_integrals(ctx, L);
    double by = 0.0;
    for (auto& I : L->integrals)
      if (!I.facet && I.rules)
        by += rule_read_bytes(ctx, S, I)
              + (8.0 + 28.0 * ctx->nv + 8.0 * S.nd) * static_cast<double>(I.rules->nrules);
    st.set_bytes(by);
  }
  DevBuf<double> tmp;
  double* d_b = b;
  if (memspace == CFX_HOST)
  {
    tmp.reserve(ctx->pool, static_cast<size_t>(S.n_total) * S.bs);
    d_b = tmp.p;
    if (!zero_first)
      CFX_CUDA(cudaMemcpyAsync(d_b, b, static_cast<size_t>(S.n_total) * S.bs * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  }
  {
    StageScope st(ctx, "gather_vector",
                  8.0 * static_cast<double>(S.n_total) + (28.0 * ctx->nv + 4.0 * S.nd) * static_cast<double>(n_std)
                      + 8.0 * S.nd * static_cast<double>(n_mat));
    GatherCtx gc = make_gather_ctx(ctx, L, nullptr);
    const StdTab stt = make_std_tab(ctx, L);
    if (zero_first)
      CFX_CUDA(cudaMemsetAsync(d_b, 0, static_cast<size_t>(S.n_total) * S.bs * sizeof(double), ctx->stream));
    CFX_DISPATCH_ELEM(ctx, S, launch_gather_vector, ctx, L, gc, stt, d_b, zero_first);
    reset_slots(ctx, L);
  }
  if (memspace == CFX_HOST)
  {
    export_to(ctx, b, tmp.p, static_cast<size_t>(S.n_total) * S.bs, CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}

cfx_status cfx_form_set_coefficient(cfx_ctx* ctx, cfx_form* f, const double* values, int64_t n, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && f && values, CFX_ERR_INVALID, "cfx_form_set_coefficient: NULL argument");
  const Space& S = ctx->spaces[f->space];
  CFX_REQUIRE(S.bs == 1, CFX_ERR_UNSUPPORTED, "Function coefficients are implemented for scalar spaces");
  CFX_REQUIRE(n == S.n_total, CFX_ERR_INVALID,
              "cfx_form_set_coefficient: the array must hold one value per owned+ghost dof of the form's space");
  if (memspace == CFX_HOST)
  {
    f->coeff_own.reserve(ctx->pool, static_cast<size_t>(n));
    CFX_CUDA(cudaMemcpyAsync(f->coeff_own.p, values, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice,
                             ctx->stream));
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
    f->coeff = f->coeff_own.p;
  }
  else
    f->coeff = values;
  CFX_API_END(ctx)
}

cfx_status cfx_assemble_scalar(cfx_ctx* ctx, const cfx_form* M_const, double* out)
{
  CFX_API_BEGIN
  cfx_form* M = const_cast<cfx_form*>(M_const);
  CFX_REQUIRE(ctx && M && out, CFX_ERR_INVALID, "cfx_assemble_scalar: NULL argument");
  CFX_REQUIRE(M->rank == 0, CFX_ERR_INVALID, "cfx_assemble_scalar: form is not a functional");
  resolve_form(ctx, M); // the sum runs over the exact number of entities
  const int64_t n = run_cell_integrals(ctx, M); // one value per entity, entity order
  constexpr int NB = 256;
  DevBuf<double> partial;
  partial.reserve(ctx->pool, NB + 1);
  CFX_LAUNCH(ctx, sum_partial_kernel, NB, 256, 0, M->Ae.p, n, partial.p);
  CFX_LAUNCH(ctx, sum_partial_kernel, 1, 256, 0, partial.p, static_cast<int64_t>(NB), partial.p + NB);
  CFX_CUDA(cudaMemcpyAsync(out, partial.p + NB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  partial.release();
  CFX_API_END(ctx)
}
} // extern "C"
