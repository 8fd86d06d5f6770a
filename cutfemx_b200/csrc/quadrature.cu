// quadrature.cu -- K2/K3: fixed-topology case-table sub-triangulation of cut cells and batched
// fp64 run-time quadrature generation; plus physical points and level-set normals/values at
// the rule points.
//
// Replaces cutcells::cut (sub-triangulation half), cutcells::select_part and
// cutcells::output::quadrature_rules as called at cpp/cutfemx/cut/cut.cpp:857,1003,1325, and
// RuntimeQuadrature::physical_points (runtime_quadrature.h:102-221), evaluate_normals
// (level_set/normal.h:39-188), evaluate_values (level_set/value.h:34-119).
//
// Conventions: points are parent-cell REFERENCE coordinates,
// weights are PHYSICAL; one rule per intersected cell whose selected part is non-empty, rules in
// ascending parent-cell order; within a rule sub-simplices in case-table order, then rule points.
//
// Local point numbering of a cut cell with inside vertices I (ascending) and outside O:
//   [ V_I0 .. V_I(n-1),  P(I0,O0), P(I0,O1), .., P(I1,O0), .. ],
//   P(a,b) = V_a + t (V_b - V_a),  t = phi_a / (phi_a - phi_b).
// "inside" is phi < 0 for selectors < <= =, and phi > 0 for > >=; a vertex with phi == 0 is
// never inside, so degenerate cuts produce zero-measure sub-simplices instead of special cases.
//
// Roofline: HBM.  Per cut cell in: 36*nv B (geometry dofs, phi, coordinates); out:
// 8*(tdim+1)*nq + 8 B (SURVEY.md section 8d "K2+K3").  Output is SoA and written coalesced:
// phase 1 (thread per cell) stages the local points in shared memory, phase 2 (thread per
// output point) expands them.
#include <cmath>

#include "common.cuh"
#include "element.cuh"

namespace cfx
{
// ---------------------------------------------------------------- built-in simplex rules (host)
namespace
{
void gauss_jacobi_01(int n, int alpha, std::vector<double>& r, std::vector<double>& w)
{
  // Gauss-Jacobi nodes for weight (1-t)^alpha on [-1,1] by Newton iteration with deflation on
  // P_n^(alpha,0); mapped to [0,1].  w_i = 2^(alpha+1) / ((1-x_i^2) P_n'(x_i)^2).
  const double a = alpha, b = 0.0;
  auto eval = [&](double x, double& p, double& dp)
  {
    double p0 = 1.0, p1 = 0.5 * (a - b) + 0.5 * (a + b + 2.0) * x;
    if (n == 0)
    {
      p = 1.0;
      dp = 0.0;
      return;
    }
    for (int k = 1; k < n; ++k)
    {
      const double k2 = 2.0 * k + a + b;
      const double c1 = 2.0 * (k + 1.0) * (k + a + b + 1.0) * k2;
      const double c2 = (k2 + 1.0) * (a * a - b * b);
      const double c3 = k2 * (k2 + 1.0) * (k2 + 2.0);
      const double c4 = 2.0 * (k + a) * (k + b) * (k2 + 2.0);
      const double p2 = ((c2 + c3 * x) * p1 - c4 * p0) / c1;
      p0 = p1;
      p1 = p2;
    }
    p = p1;
    // derivative: (1-x^2)(2n+a+b) P_n' = n[(a-b) - (2n+a+b)x] P_n + 2(n+a)(n+b) P_{n-1}
    const double t = 2.0 * n + a + b;
    dp = (n * (a - b - t * x) * p1 + 2.0 * (n + a) * (n + b) * p0) / (t * (1.0 - x * x));
  };
  std::vector<double> x(n);
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < n; ++k)
  {
    double xk = -std::cos((2.0 * k + 1.0) * pi / (2.0 * n));
    if (k > 0)
      xk = 0.5 * (xk + x[k - 1]);
    for (int it = 0; it < 100; ++it)
    {
      double s = 0.0;
      for (int j = 0; j < k; ++j)
        s += 1.0 / (xk - x[j]);
      double p, dp;
      eval(xk, p, dp);
      const double dx = p / (dp - s * p);
      xk -= dx;
      if (std::fabs(dx) < 1e-16)
        break;
    }
    x[k] = xk;
  }
  r.resize(n);
  w.resize(n);
  for (int k = 0; k < n; ++k)
  {
    double p, dp;
    eval(x[k], p, dp);
    const double wk = std::pow(2.0, a + 1.0) / ((1.0 - x[k] * x[k]) * dp * dp);
    r[k] = 0.5 * (x[k] + 1.0);
    w[k] = wk / std::pow(2.0, a + 1.0);
  }
}

void push_s21(std::vector<double>& p, double a)
{
  const double c = 1.0 - 2.0 * a;
  const double q[3][2] = {{a, a}, {c, a}, {a, c}};
  for (auto& v : q)
    p.insert(p.end(), v, v + 2);
}
void push_s31(std::vector<double>& p, double a)
{
  const double c = 1.0 - 3.0 * a;
  const double q[4][3] = {{a, a, a}, {c, a, a}, {a, c, a}, {a, a, c}};
  for (auto& v : q)
    p.insert(p.end(), v, v + 3);
}
void push_s22(std::vector<double>& p, double b)
{
  const double d = 0.5 - b;
  const double q[6][3] = {{b, d, d}, {d, b, d}, {d, d, b}, {d, b, b}, {b, d, b}, {b, b, d}};
  for (auto& v : q)
    p.insert(p.end(), v, v + 3);
}
} // namespace

// Rule of polynomial degree `order` on the unit simplex of dimension dim; weights sum to 1/dim!.
// Fully symmetric positive rules up to degree 5, collapsed Gauss-Jacobi beyond.
void builtin_simplex_rule(int dim, int order, std::vector<double>& pts, std::vector<double>& wts)
{
  CFX_REQUIRE(order >= 0, CFX_ERR_INVALID, "quadrature order must be >= 0"); // cut.cpp:164-168
  CFX_REQUIRE(order <= 30, CFX_ERR_UNSUPPORTED, "quadrature order above 30 is not supported");
  CFX_REQUIRE(dim >= 0 && dim <= 3, CFX_ERR_INVALID, "simplex dimension must be 0..3");
  pts.clear();
  wts.clear();
  const int m = order / 2 + 1;
  if (dim == 0)
  {
    wts = {1.0};
    return;
  }
  if (dim == 1)
  {
    gauss_jacobi_01(m, 0, pts, wts);
    return;
  }
  if (dim == 2)
  {
    if (order <= 1)
    {
      pts = {1.0 / 3.0, 1.0 / 3.0};
      wts = {0.5};
    }
    else if (order == 2)
    {
      push_s21(pts, 1.0 / 6.0);
      wts.assign(3, 1.0 / 6.0);
    }
    else if (order <= 4)
    {
      push_s21(pts, 0.4459484909159648863183293);
      push_s21(pts, 0.09157621350977074345957146);
      wts = {0.1116907948390057328475035,  0.1116907948390057328475035,  0.1116907948390057328475035,
             0.05497587182766093381916316, 0.05497587182766093381916316, 0.05497587182766093381916316};
    }
    else if (order == 5)
    {
      const double s15 = std::sqrt(15.0);
      pts = {1.0 / 3.0, 1.0 / 3.0};
      push_s21(pts, (6.0 - s15) / 21.0);
      push_s21(pts, (6.0 + s15) / 21.0);
      const double wa = (155.0 - s15) / 2400.0, wb = (155.0 + s15) / 2400.0;
      wts = {0.1125, wa, wa, wa, wb, wb, wb};
    }
    else
    {
      std::vector<double> r, wr, s, ws;
      gauss_jacobi_01(m, 1, r, wr);
      gauss_jacobi_01(m, 0, s, ws);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j)
        {
          pts.push_back(r[i]);
          pts.push_back(s[j] * (1.0 - r[i]));
          wts.push_back(wr[i] * ws[j]);
        }
    }
    return;
  }
  if (order <= 1)
  {
    pts = {0.25, 0.25, 0.25};
    wts = {1.0 / 6.0};
  }
  else if (order == 2)
  {
    push_s31(pts, (5.0 - std::sqrt(5.0)) / 20.0);
    wts.assign(4, 1.0 / 24.0);
  }
  else if (order <= 5)
  {
    push_s31(pts, 0.3108859192633006097973457);
    push_s31(pts, 0.09273525031089122640232391);
    push_s22(pts, 0.04550370412564964949188053);
    wts.assign(4, 0.01878132095300264179986428);
    wts.insert(wts.end(), 4, 0.01224884051939365825728503);
    wts.insert(wts.end(), 6, 0.007091003462846911073011571);
  }
  else
  {
    std::vector<double> r, wr, s, ws, t, wt;
    gauss_jacobi_01(m, 2, r, wr);
    gauss_jacobi_01(m, 1, s, ws);
    gauss_jacobi_01(m, 0, t, wt);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j)
        for (int k = 0; k < m; ++k)
        {
          pts.push_back(r[i]);
          pts.push_back(s[j] * (1.0 - r[i]));
          pts.push_back(t[k] * (1.0 - r[i]) * (1.0 - s[j]));
          wts.push_back(wr[i] * ws[j] * wt[k]);
        }
  }
}

RuleTable& get_rule(cfx_ctx* c, int dim, int order)
{
  auto key = std::make_pair(dim, order);
  auto it = c->rules.find(key);
  if (it == c->rules.end())
  {
    RuleTable t;
    t.dim = dim;
    t.order = order;
    builtin_simplex_rule(dim, order, t.pts, t.wts);
    t.npts = static_cast<int>(t.wts.size());
    it = c->rules.emplace(key, std::move(t)).first;
  }
  RuleTable& t = it->second;
  if (!t.d_wts)
  {
    t.d_wts = static_cast<double*>(c->pool.alloc(t.wts.size() * sizeof(double)));
    t.d_pts = static_cast<double*>(c->pool.alloc((t.pts.size() + 1) * sizeof(double)));
    CFX_CUDA(cudaMemcpyAsync(t.d_wts, t.wts.data(), t.wts.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (!t.pts.empty())
      CFX_CUDA(
          cudaMemcpyAsync(t.d_pts, t.pts.data(), t.pts.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CFX_CUDA(cudaStreamSynchronize(c->stream)); // host vectors may be reallocated later
  }
  return t;
}

// ---------------------------------------------------------------- case tables
namespace
{
// sub-simplices as indices into the local point list; [n_in][sub][vertex]
__constant__ int8_t c_tri_vol[3][2][3] = {{{0, 0, 0}, {0, 0, 0}}, {{0, 1, 2}, {0, 0, 0}}, {{0, 1, 3}, {0, 3, 2}}};
__constant__ int8_t c_tri_ifc[3][1][2] = {{{0, 0}}, {{1, 2}}, {{2, 3}}};
__constant__ int8_t c_tet_vol[4][3][4] = {{{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
                                          {{0, 1, 2, 3}, {0, 0, 0, 0}, {0, 0, 0, 0}},
                                          {{0, 2, 3, 1}, {2, 3, 1, 4}, {3, 1, 4, 5}},
                                          {{0, 1, 2, 3}, {1, 2, 3, 4}, {2, 3, 4, 5}}};
__constant__ int8_t c_tet_ifc[4][2][3] = {{{0, 0, 0}, {0, 0, 0}}, {{1, 2, 3}, {0, 0, 0}}, {{2, 3, 5}, {2, 5, 4}},
                                          {{3, 4, 5}, {0, 0, 0}}};

__host__ __device__ inline int num_sub(int tdim, bool interface, int n_in)
{
  if (n_in <= 0 || n_in > tdim)
    return 0;
  if (tdim == 2)
    return interface ? 1 : n_in;            // vol: 1, 2 ; ifc: 1, 1
  return interface ? (n_in == 2 ? 2 : 1)    // ifc: 1, 2, 1
                   : (n_in == 1 ? 1 : 3);   // vol: 1, 3, 3
}

template <int TDIM>
__device__ __forceinline__ int inside_mask(const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals,
                                           int64_t cell, bool positive, double (&phi)[TDIM + 1])
{
  constexpr int NV = TDIM + 1;
  int mask = 0;
#pragma unroll
  for (int v = 0; v < NV; ++v)
  {
    phi[v] = __ldg(vals + __ldg(ls_dofmap + cell * NV + v));
    const bool in = positive ? (phi[v] > 0.0) : (phi[v] < 0.0);
    mask |= in ? (1 << v) : 0;
  }
  return mask;
}

constexpr int QB = 128; // cut cells per block == threads per block

// pass 1: packed (rule flag << 32 | number of points) per cut cell
template <int TDIM>
__global__ void __launch_bounds__(QB)
    rule_count_kernel(const int32_t* __restrict__ cut_cells, DN n_cut_, const int32_t* __restrict__ ls_dofmap,
                      const double* __restrict__ vals, bool positive, bool interface, int npts_s,
                      int64_t* __restrict__ packed)
{
  const int64_t n_cut = n_cut_.get();
  const int64_t k = static_cast<int64_t>(blockIdx.x) * QB + threadIdx.x;
  if (k >= n_cut)
    return;
  double phi[TDIM + 1];
  const int mask = inside_mask<TDIM>(ls_dofmap, vals, cut_cells[k], positive, phi);
  const int nq = num_sub(TDIM, interface, __popc(mask)) * npts_s;
  packed[k] = nq > 0 ? ((int64_t(1) << 32) | nq) : 0;
}

// pass 2: expand. smem per cell: local points (MAXP x TDIM), per-sub scale, n_in, local offset.
template <int TDIM>
struct CutSmem
{
  static constexpr int MAXP = TDIM == 2 ? 4 : 6;
  static constexpr int MAXS = TDIM == 2 ? 2 : 3;
  static constexpr int PSTRIDE = MAXP * TDIM + 1; // odd stride: fewer bank conflicts
  double P[QB][PSTRIDE];
  double scale[QB][MAXS];
  int off[QB + 1];
  int8_t n_in[QB];
};

// (interface rules carry second moments: 92 registers and 29 % occupancy unbounded; 7 blocks per SM = 72 registers,
// 4 bytes spilled -- profiles/r9_ncu_full_rule_fill_kernel_interface_n256.txt)
template <int TDIM, bool interface>
__global__ void __launch_bounds__(QB, interface ? 7 : 8)
    rule_fill_kernel(const int32_t* __restrict__ cut_cells, DN n_cut_, const int64_t* __restrict__ packed_excl,
                     const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals,
                     const int32_t* __restrict__ x_dofmap, const double* __restrict__ x, bool positive,
                     int npts_s, const double* __restrict__ rule_pts, const double* __restrict__ rule_wts,
                     int64_t cap_pts, int64_t cap_rules, int64_t* __restrict__ d_sizes /* [0] nrules, [1] npts */,
                     int32_t* __restrict__ err, double* __restrict__ points /* SoA (TDIM, npts_total) */,
                     double* __restrict__ weights, int32_t* __restrict__ offsets, int32_t* __restrict__ parent_map,
                     double* __restrict__ moments /* (nrules, TDIM + 1) or null */)
{
  constexpr int NV = TDIM + 1;
  __shared__ CutSmem<TDIM> sm;
  const int tid = threadIdx.x;
  const int64_t n_cut = n_cut_.get();
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * QB;
  if (k0 >= n_cut && blockIdx.x != 0)
    return;
  // totals (the scan's closing entry): the SoA stride of the point arrays is the exact number of points
  const int64_t pk_total = packed_excl[n_cut];
  const int64_t npts_total = pk_total & 0xffffffffLL;
  const bool fits = npts_total <= cap_pts && (pk_total >> 32) <= cap_rules;
  if (blockIdx.x == 0 && tid == 0)
  {
    d_sizes[0] = fits ? (pk_total >> 32) : 0;
    d_sizes[1] = fits ? npts_total : 0;
    if (!fits)
    { // deferred-size mode: this step's rules do not fit the buffers of the object that is being reused
      err[0] = 32;
      err[1] = static_cast<int32_t>(npts_total);
    }
    if (n_cut == 0 && fits)
      offsets[0] = 0;
  }
  if (!fits || k0 >= n_cut)
    return;
  const int64_t k = k0 + tid;
  // packed_excl has n_cut + 1 entries (last = totals)
  const int64_t pk_first = packed_excl[k0];
  const int64_t k_end = (k0 + QB < n_cut) ? k0 + QB : n_cut;
  const int64_t pk_last = packed_excl[k_end];
  const int64_t pt_first = pk_first & 0xffffffffLL;
  const int n_block_pts = static_cast<int>((pk_last & 0xffffffffLL) - pt_first);

  int my_nq = 0;
  if (k < n_cut)
  {
    const int64_t cell = cut_cells[k];
    const int64_t pk = packed_excl[k];
    const int64_t pk_next = packed_excl[k + 1];
    my_nq = static_cast<int>((pk_next & 0xffffffffLL) - (pk & 0xffffffffLL));
    sm.off[tid] = static_cast<int>((pk & 0xffffffffLL) - pt_first);
    double phi[NV];
    const int mask = inside_mask<TDIM>(ls_dofmap, vals, cell, positive, phi);
    const int n_in = __popc(mask);
    sm.n_in[tid] = static_cast<int8_t>(my_nq > 0 ? n_in : 0);
    if (my_nq > 0)
    {
      const int64_t rule = pk >> 32;
      offsets[rule] = static_cast<int32_t>(pk & 0xffffffffLL);
      parent_map[rule] = static_cast<int32_t>(cell);
      // ordered vertex lists: inside ascending, outside ascending
      int I[NV], O[NV];
      int ni = 0, no = 0;
#pragma unroll
      for (int v = 0; v < NV; ++v)
      {
        if (mask & (1 << v))
          I[ni++] = v;
        else
          O[no++] = v;
      }
      double* P = sm.P[tid];
      for (int i = 0; i < ni; ++i)
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          P[i * TDIM + t] = (I[i] == t + 1) ? 1.0 : 0.0;
      int np = ni;
      for (int i = 0; i < ni; ++i)
        for (int o = 0; o < no; ++o)
        {
          const int a = I[i], b = O[o];
          const double tp = phi[a] / (phi[a] - phi[b]);
#pragma unroll
          for (int t = 0; t < TDIM; ++t)
          {
            const double va = (a == t + 1) ? 1.0 : 0.0;
            const double vb = (b == t + 1) ? 1.0 : 0.0;
            P[np * TDIM + t] = va + tp * (vb - va);
          }
          ++np;
        }
      // geometry of the parent cell
      double X[NV][TDIM];
      load_cell_coords<TDIM>(x, x_dofmap, cell, X);
      Geo<TDIM> g;
      make_geo<TDIM>(X, g);
      const int nsub = num_sub(TDIM, interface, n_in);
      double mW = 0.0, mX[TDIM], mXX[TDIM * (TDIM + 1) / 2];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        mX[t] = 0.0;
#pragma unroll
      for (int t = 0; t < TDIM * (TDIM + 1) / 2; ++t)
        mXX[t] = 0.0;
      for (int s = 0; s < nsub; ++s)
      {
        double scale;
        if (!interface)
        {
          const int8_t* sv = TDIM == 2 ? c_tri_vol[n_in][s] : c_tet_vol[n_in][s];
          double M[TDIM][TDIM];
#pragma unroll
          for (int r = 0; r < TDIM; ++r)
#pragma unroll
            for (int cc = 0; cc < TDIM; ++cc)
              M[r][cc] = P[sv[cc + 1] * TDIM + r] - P[sv[0] * TDIM + r];
          double det;
          if constexpr (TDIM == 2)
            det = M[0][0] * M[1][1] - M[0][1] * M[1][0];
          else
            det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
                  + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
          scale = fabs(det) * fabs(g.detJ);
          // measure and first moments of the sub-simplex: W_s = scale / tdim!, centroid = vertex mean
          const double Ws = scale * (TDIM == 2 ? 0.5 : 1.0 / 6.0);
          mW += Ws;
#pragma unroll
          for (int t = 0; t < TDIM; ++t)
          {
            double cs = 0.0;
#pragma unroll
            for (int vtx = 0; vtx < NV; ++vtx)
              cs += P[sv[vtx] * TDIM + t];
            mX[t] += Ws * (cs * (1.0 / NV));
          }
        }
        else
        {
          const int8_t* sv = TDIM == 2 ? c_tri_ifc[n_in][s] : c_tet_ifc[n_in][s];
          double Xp[TDIM][TDIM]; // physical vertices of the interface simplex
#pragma unroll
          for (int kk = 0; kk < TDIM; ++kk)
#pragma unroll
            for (int r = 0; r < TDIM; ++r)
            {
              double v = g.x0[r];
#pragma unroll
              for (int t = 0; t < TDIM; ++t)
                v += g.J[r * TDIM + t] * P[sv[kk] * TDIM + t];
              Xp[kk][r] = v;
            }
          if constexpr (TDIM == 2)
          {
            const double dx = Xp[1][0] - Xp[0][0], dy = Xp[1][1] - Xp[0][1];
            scale = sqrt(dx * dx + dy * dy);
          }
          else
          {
            const double u0 = Xp[1][0] - Xp[0][0], u1 = Xp[1][1] - Xp[0][1], u2 = Xp[1][2] - Xp[0][2];
            const double w0 = Xp[2][0] - Xp[0][0], w1 = Xp[2][1] - Xp[0][1], w2 = Xp[2][2] - Xp[0][2];
            const double cx = u1 * w2 - u2 * w1, cy = u2 * w0 - u0 * w2, cz = u0 * w1 - u1 * w0;
            scale = sqrt(cx * cx + cy * cy + cz * cz); // = 2 * area = area * (tdim-1)!
          }
          // measure, first and second moments (reference coordinates) of the interface simplex with vertices v:
          // int xi_a = |T| mean(v_a),  int xi_a xi_b = |T| (sum v_a sum v_b + sum v_a v_b) / (n (n + 1)), n = tdim
          const double Ws = scale * (TDIM == 3 ? 0.5 : 1.0);
          mW += Ws;
          double sa[TDIM];
#pragma unroll
          for (int t = 0; t < TDIM; ++t)
          {
            double cs = 0.0;
#pragma unroll
            for (int kk = 0; kk < TDIM; ++kk)
              cs += P[sv[kk] * TDIM + t];
            sa[t] = cs;
            mX[t] += Ws * (cs * (1.0 / TDIM));
          }
          int m2 = 0;
#pragma unroll
          for (int a = 0; a < TDIM; ++a)
#pragma unroll
            for (int b = a; b < TDIM; ++b)
            {
              double pp = 0.0;
#pragma unroll
              for (int kk = 0; kk < TDIM; ++kk)
                pp += P[sv[kk] * TDIM + a] * P[sv[kk] * TDIM + b];
              mXX[m2++] += Ws * ((sa[a] * sa[b] + pp) * (1.0 / (TDIM * (TDIM + 1))));
            }
        }
        sm.scale[tid][s] = scale;
      }
      if (moments)
      { // volume rules: (W, first moments); interface rules: (W, first, second moments, upper triangle)
        constexpr int NM2 = TDIM * (TDIM + 1) / 2;
        double* mo = moments + rule * (interface ? 1 + TDIM + NM2 : 1 + TDIM);
        mo[0] = mW;
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          mo[1 + t] = mX[t];
        if (interface)
        {
#pragma unroll
          for (int t = 0; t < NM2; ++t)
            mo[1 + TDIM + t] = mXX[t];
        }
      }
    }
  }
  else
  {
    sm.off[tid] = n_block_pts;
    sm.n_in[tid] = 0;
  }
  if (tid == 0)
    sm.off[QB] = n_block_pts;
  if (k == n_cut - 1)
    offsets[pk_last >> 32] = static_cast<int32_t>(npts_total); // closing offset
  __syncthreads();

  // phase 2: one thread per output point, coalesced SoA stores
  constexpr int SD_VOL = TDIM;
  for (int p = tid; p < n_block_pts; p += QB)
  {
    // largest j with off[j] <= p (cells with no points share their successor's offset)
    int lo = 0, hi = QB;
    while (hi - lo > 1)
    {
      const int mid = (lo + hi) >> 1;
      if (sm.off[mid] <= p)
        lo = mid;
      else
        hi = mid;
    }
    const int j = lo;
    const int local = p - sm.off[j];
    const int s = local / npts_s;
    const int q = local - s * npts_s;
    const int n_in = sm.n_in[j];
    const double* P = sm.P[j];
    double xi[TDIM];
    const int sd = interface ? TDIM - 1 : SD_VOL;
    const int8_t* sv;
    if (!interface)
      sv = TDIM == 2 ? c_tri_vol[n_in][s] : c_tet_vol[n_in][s];
    else
      sv = TDIM == 2 ? c_tri_ifc[n_in][s] : c_tet_ifc[n_in][s];
    double l0 = 1.0;
    double lam[TDIM];
#pragma unroll
    for (int cc = 0; cc < TDIM; ++cc)
    {
      lam[cc] = cc < sd ? __ldg(rule_pts + q * sd + cc) : 0.0;
      l0 -= lam[cc];
    }
#pragma unroll
    for (int d = 0; d < TDIM; ++d)
    {
      double v = l0 * P[sv[0] * TDIM + d];
#pragma unroll
      for (int cc = 0; cc < TDIM; ++cc)
        if (cc < sd)
          v += lam[cc] * P[sv[cc + 1] * TDIM + d];
      xi[d] = v;
    }
    const int64_t gp = pt_first + p;
#pragma unroll
    for (int d = 0; d < TDIM; ++d)
      points[static_cast<int64_t>(d) * npts_total + gp] = xi[d];
    weights[gp] = __ldg(rule_wts + q) * sm.scale[j][s];
  }
}

// ------------------------------------------------------------------ P2 level sets
// cutfemx.cut(level_set) with a degree-2 level set, default options (cut_approximation "auto", order 1,
// wrappers/cut.cpp:117-140): straight pieces between edge roots of the higher-order function.  Algorithm (the
// oracle's cut_cell_rule_p2 states it and the reasons; CutCells itself is absent): one red refinement through the
// P2 nodes, marching-simplex case tables per sub-simplex on the nodal values (a sub-simplex entirely inside
// contributes whole), edge roots = the root in [0, 1] of the quadratic restriction of the level set to the sub-edge,
// points in the PARENT cell's reference coordinates, physical weights, one rule per cut cell.
// One thread per cut cell in both passes: the work per cell is irregular (up to 8 sub-cells x 3 pieces x npts) and
// the feature is a "next" row -- correctness and parity first.
__constant__ int8_t c_red_tri[4][3] = {{0, 5, 4}, {1, 3, 5}, {2, 4, 3}, {3, 4, 5}};
__constant__ int8_t c_red_tet[8][4] = {{0, 9, 8, 7}, {1, 9, 6, 5}, {2, 8, 6, 4}, {3, 7, 5, 4},
                                       {9, 8, 7, 5}, {9, 8, 6, 5}, {8, 7, 5, 4}, {8, 6, 5, 4}};
// reference coordinates of the P2 nodes, times 2
__constant__ int8_t c_node2_tri[6][2] = {{0, 0}, {2, 0}, {0, 2}, {1, 1}, {0, 1}, {1, 0}};
__constant__ int8_t c_node2_tet[10][3] = {{0, 0, 0}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {0, 1, 1},
                                          {1, 0, 1}, {1, 1, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};

__device__ __forceinline__ double quadratic_edge_root(double fa, double fm, double fb)
{
  const double a = 2.0 * fa - 4.0 * fm + 2.0 * fb, b = -3.0 * fa + 4.0 * fm - fb, c = fa;
  const double lin = fa / (fa - fb);
  if (fabs(a) <= 1e-14 * (fabs(b) + fabs(c)))
    return lin;
  const double disc = b * b - 4.0 * a * c;
  if (disc < 0.0)
    return lin;
  const double sq = sqrt(disc);
  const double qq = -0.5 * (b + (b >= 0.0 ? sq : -sq));
  const double r1 = qq / a, r2 = (qq != 0.0) ? c / qq : r1;
  const bool ok1 = r1 >= 0.0 && r1 <= 1.0, ok2 = r2 >= 0.0 && r2 <= 1.0;
  if (ok1 && ok2)
    return fabs(r1 - lin) <= fabs(r2 - lin) ? r1 : r2;
  if (ok1)
    return r1;
  if (ok2)
    return r2;
  return lin;
}

// walks the pieces of one cut cell; emit(P, sv, nsv) gets the local point list (parent reference coordinates,
// stride TDIM) and the piece's vertex indices
template <int TDIM, class Emit>
__device__ __forceinline__ void p2_cell_pieces(const double (&dofs)[TDIM == 2 ? 6 : 10], bool positive, bool interface,
                                               Emit emit)
{
  constexpr int NV = TDIM + 1;
  constexpr int NSC = TDIM == 2 ? 4 : 8;
  for (int sc = 0; sc < NSC; ++sc)
  {
    double V[NV][TDIM], f[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      const int node = TDIM == 2 ? c_red_tri[sc][v] : c_red_tet[sc][v];
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        V[v][t] = 0.5 * (TDIM == 2 ? c_node2_tri[node][t] : c_node2_tet[node][t]);
      double fv = dofs[0];
#pragma unroll
      for (int j = 1; j < (TDIM == 2 ? 6 : 10); ++j)
        fv = (j == node) ? dofs[j] : fv;
      f[v] = fv;
    }
    int I[NV], O[NV], n_in = 0, n_out = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      const bool in = positive ? (f[v] > 0.0) : (f[v] < 0.0);
      if (in)
        I[n_in++] = v;
      else
        O[n_out++] = v;
    }
    if (n_in == 0)
      continue;
    double P[(TDIM == 2 ? 4 : 6) * TDIM];
    if (n_in == NV)
    {
      if (interface)
        continue;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          P[v * TDIM + t] = V[v][t];
      const int8_t whole[4] = {0, 1, 2, 3};
      emit(P, whole, NV);
      continue;
    }
    for (int i = 0; i < n_in; ++i)
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        P[i * TDIM + t] = V[I[i]][t];
    int np = n_in;
    for (int i = 0; i < n_in; ++i)
      for (int o = 0; o < n_out; ++o)
      {
        const int a = I[i], b = O[o];
        double Xm[TDIM];
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          Xm[t] = 0.5 * (V[a][t] + V[b][t]);
        double phi[TDIM == 2 ? 6 : 10], dphi[TDIM == 2 ? 6 : 10][TDIM];
        tabulate<TDIM, 2>(Xm, phi, dphi);
        double fm = 0.0;
#pragma unroll
        for (int j = 0; j < (TDIM == 2 ? 6 : 10); ++j)
          fm += phi[j] * dofs[j];
        const double tt = quadratic_edge_root(f[a], fm, f[b]);
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
          P[np * TDIM + t] = V[a][t] + tt * (V[b][t] - V[a][t]);
        ++np;
      }
    const int nsub = num_sub(TDIM, interface, n_in);
    for (int s = 0; s < nsub; ++s)
    {
      const int8_t* sv = interface ? (TDIM == 2 ? c_tri_ifc[n_in][s] : c_tet_ifc[n_in][s])
                                   : (TDIM == 2 ? c_tri_vol[n_in][s] : c_tet_vol[n_in][s]);
      emit(P, sv, interface ? TDIM : NV);
    }
  }
}

template <int TDIM>
__device__ __forceinline__ void load_p2_dofs(const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals,
                                             int64_t cell, double (&dofs)[TDIM == 2 ? 6 : 10])
{
  constexpr int ND = TDIM == 2 ? 6 : 10;
#pragma unroll
  for (int j = 0; j < ND; ++j)
    dofs[j] = __ldg(vals + __ldg(ls_dofmap + cell * ND + j));
}

template <int TDIM>
__global__ void __launch_bounds__(QB)
    rule_count_p2_kernel(const int32_t* __restrict__ cut_cells, DN n_cut_, const int32_t* __restrict__ ls_dofmap,
                         const double* __restrict__ vals, bool positive, bool interface, int npts_s,
                         int64_t* __restrict__ packed)
{
  const int64_t k = static_cast<int64_t>(blockIdx.x) * QB + threadIdx.x;
  if (k >= n_cut_.get())
    return;
  double dofs[TDIM == 2 ? 6 : 10];
  load_p2_dofs<TDIM>(ls_dofmap, vals, cut_cells[k], dofs);
  int pieces = 0;
  p2_cell_pieces<TDIM>(dofs, positive, interface, [&](const double*, const int8_t*, int) { ++pieces; });
  const int nq = pieces * npts_s;
  packed[k] = nq > 0 ? ((int64_t(1) << 32) | nq) : 0;
}

template <int TDIM>
__global__ void __launch_bounds__(QB)
    rule_fill_p2_kernel(const int32_t* __restrict__ cut_cells, DN n_cut_, const int64_t* __restrict__ packed_excl,
                        const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals,
                        const int32_t* __restrict__ x_dofmap, const double* __restrict__ x, bool positive,
                        bool interface, int npts_s, const double* __restrict__ rule_pts,
                        const double* __restrict__ rule_wts, int64_t cap_pts, int64_t cap_rules,
                        int64_t* __restrict__ d_sizes, int32_t* __restrict__ err, double* __restrict__ points,
                        double* __restrict__ weights, int32_t* __restrict__ offsets, int32_t* __restrict__ parent_map)
{
  constexpr int NV = TDIM + 1;
  const int64_t n_cut = n_cut_.get();
  const int64_t pk_total = packed_excl[n_cut];
  const int64_t npts_total = pk_total & 0xffffffffLL;
  const bool fits = npts_total <= cap_pts && (pk_total >> 32) <= cap_rules;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * QB + threadIdx.x;
  if (k == 0)
  {
    d_sizes[0] = fits ? (pk_total >> 32) : 0;
    d_sizes[1] = fits ? npts_total : 0;
    if (!fits)
    {
      err[0] = 32;
      err[1] = static_cast<int32_t>(npts_total);
    }
    else
      offsets[pk_total >> 32] = static_cast<int32_t>(npts_total); // closing offset
  }
  if (!fits || k >= n_cut)
    return;
  const int64_t pk = packed_excl[k];
  const int my_nq = static_cast<int>((packed_excl[k + 1] & 0xffffffffLL) - (pk & 0xffffffffLL));
  if (my_nq == 0)
    return;
  const int64_t cell = cut_cells[k];
  const int64_t rule = pk >> 32;
  offsets[rule] = static_cast<int32_t>(pk & 0xffffffffLL);
  parent_map[rule] = static_cast<int32_t>(cell);
  double dofs[TDIM == 2 ? 6 : 10];
  load_p2_dofs<TDIM>(ls_dofmap, vals, cell, dofs);
  double X[NV][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  int64_t gp = pk & 0xffffffffLL;
  p2_cell_pieces<TDIM>(dofs, positive, interface,
                       [&](const double* P, const int8_t* sv, int nsv)
                       {
                         double scale;
                         if (!interface)
                         {
                           double M[TDIM][TDIM];
#pragma unroll
                           for (int r = 0; r < TDIM; ++r)
#pragma unroll
                             for (int cc = 0; cc < TDIM; ++cc)
                               M[r][cc] = P[sv[cc + 1] * TDIM + r] - P[sv[0] * TDIM + r];
                           double det;
                           if constexpr (TDIM == 2)
                             det = M[0][0] * M[1][1] - M[0][1] * M[1][0];
                           else
                             det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1])
                                   - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
                                   + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
                           scale = fabs(det) * fabs(g.detJ);
                         }
                         else
                         {
                           double Xp[TDIM][TDIM];
#pragma unroll
                           for (int kk = 0; kk < TDIM; ++kk)
#pragma unroll
                             for (int r = 0; r < TDIM; ++r)
                             {
                               double v = g.x0[r];
#pragma unroll
                               for (int t = 0; t < TDIM; ++t)
                                 v += g.J[r * TDIM + t] * P[sv[kk] * TDIM + t];
                               Xp[kk][r] = v;
                             }
                           if constexpr (TDIM == 2)
                           {
                             const double dx = Xp[1][0] - Xp[0][0], dy = Xp[1][1] - Xp[0][1];
                             scale = sqrt(dx * dx + dy * dy);
                           }
                           else
                           {
                             const double u0 = Xp[1][0] - Xp[0][0], u1 = Xp[1][1] - Xp[0][1], u2 = Xp[1][2] - Xp[0][2];
                             const double w0 = Xp[2][0] - Xp[0][0], w1 = Xp[2][1] - Xp[0][1], w2 = Xp[2][2] - Xp[0][2];
                             const double cx = u1 * w2 - u2 * w1, cy = u2 * w0 - u0 * w2, cz = u0 * w1 - u1 * w0;
                             scale = sqrt(cx * cx + cy * cy + cz * cz);
                           }
                         }
                         const int sd = nsv - 1;
                         for (int q = 0; q < npts_s; ++q)
                         {
                           double l0 = 1.0, lam[TDIM];
#pragma unroll
                           for (int cc = 0; cc < TDIM; ++cc)
                           {
                             lam[cc] = cc < sd ? __ldg(rule_pts + q * sd + cc) : 0.0;
                             l0 -= lam[cc];
                           }
#pragma unroll
                           for (int d = 0; d < TDIM; ++d)
                           {
                             double v = l0 * P[sv[0] * TDIM + d];
#pragma unroll
                             for (int cc = 0; cc < TDIM; ++cc)
                               if (cc < sd)
                                 v += lam[cc] * P[sv[cc + 1] * TDIM + d];
                             points[static_cast<int64_t>(d) * npts_total + gp] = v;
                           }
                           weights[gp] = __ldg(rule_wts + q) * scale;
                           ++gp;
                         }
                       });
}

// SoA (dim, n) -> AoS (n, dim)
__global__ void soa_to_aos_kernel(const double* __restrict__ soa, int64_t n, int dim, double* __restrict__ aos)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * dim)
    return;
  const int64_t p = i / dim;
  const int d = static_cast<int>(i - p * dim);
  aos[i] = soa[static_cast<int64_t>(d) * n + p];
}

template <int TDIM>
__global__ void physical_points_kernel(const double* __restrict__ pts, int64_t npts,
                                       const int32_t* __restrict__ offsets, const int32_t* __restrict__ parent_map,
                                       int64_t nrules, const int32_t* __restrict__ x_dofmap,
                                       const double* __restrict__ x, double* __restrict__ out)
{
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= npts)
    return;
  const int64_t r = find_rule(offsets, nrules, q);
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, parent_map[r], X);
  double xi[TDIM], l0 = 1.0;
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
  {
    xi[t] = pts[static_cast<int64_t>(t) * npts + q];
    l0 -= xi[t];
  }
#pragma unroll
  for (int d = 0; d < TDIM; ++d)
  {
    double v = X[0][d] * l0;
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      v += X[t + 1][d] * xi[t];
    out[static_cast<int64_t>(d) * npts + q] = v;
  }
}

// level_set/normal.h:116-185 (per point: K = J^-1, grad = K^T sum_j dphi_j phi_j, floor 1e-14)
template <int TDIM, int DEG>
__global__ void normals_kernel(const double* __restrict__ pts, DN npts_, const int32_t* __restrict__ offsets,
                               const int32_t* __restrict__ parent_map, DN nrules_,
                               const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                               const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals, double sign,
                               double* __restrict__ out_soa, double* __restrict__ out_aos)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  const int64_t npts = npts_.get(), nrules = nrules_.get();
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= npts)
    return;
  const int64_t r = find_rule(offsets, nrules, q);
  const int64_t cell = parent_map[r];
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  double xi[TDIM];
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
    xi[t] = pts[static_cast<int64_t>(t) * npts + q];
  double phi[ND], dphi[ND][TDIM];
  tabulate<TDIM, DEG>(xi, phi, dphi);
  double gref[TDIM];
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
    gref[t] = 0.0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
  {
    const double v = __ldg(vals + __ldg(ls_dofmap + cell * ND + j));
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      gref[t] += dphi[j][t] * v;
  }
  double gp[TDIM], nrm = 0.0;
#pragma unroll
  for (int i = 0; i < TDIM; ++i)
  {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      s += g.K[t * TDIM + i] * gref[t];
    gp[i] = s;
    nrm += s * s;
  }
  nrm = sqrt(nrm);
  if (nrm < 1.0e-14)
    nrm = 1.0e-14;
#pragma unroll
  for (int i = 0; i < TDIM; ++i)
  {
    const double v = sign * gp[i] / nrm;
    out_soa[static_cast<int64_t>(i) * npts + q] = v;
    if (out_aos)
      out_aos[q * TDIM + i] = v;
  }
}

// P1 level sets: grad phi is constant on the cell, so every point of a rule gets the same normal -- one thread
// per RULE evaluates the reference's per-point expression once (same operations in the same order, hence the
// same bits as normals_kernel<TDIM, 1>; dphi of P1 does not depend on the point) into shared memory; then the
// block's threads walk the block's contiguous point range and store coalesced (a thread storing its own rule's
// points one by one issued 32 scattered sectors per store instruction: 60 lg-throttle stalls per issue, 7 % issue
// utilisation, profiles/r9_ncu_full_normals_p1_kernel_n256.txt).  One geometry evaluation per cut cell.
constexpr int NPB = 128; // rules per block
template <int TDIM>
__global__ void __launch_bounds__(NPB) normals_p1_kernel(const double* __restrict__ pts, DN npts_, const int32_t* __restrict__ offsets,
                                  const int32_t* __restrict__ parent_map, DN nrules_,
                                  const int32_t* __restrict__ x_dofmap, const double* __restrict__ x,
                                  const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals, double sign,
                                  double* __restrict__ out_soa, double* __restrict__ out_aos)
{
  constexpr int ND = TDIM + 1;
  __shared__ int32_t s_off[NPB + 1];
  __shared__ double s_n[TDIM][NPB];
  const int64_t npts = npts_.get(), nrules = nrules_.get();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * NPB;
  if (r0 >= nrules)
    return;
  const int nb = static_cast<int>(nrules - r0 < NPB ? nrules - r0 : NPB); // rules of this block
  const int64_t r = r0 + threadIdx.x;
  if (threadIdx.x <= nb)
    s_off[threadIdx.x] = offsets[r];
  if (threadIdx.x == 0)
    s_off[nb] = offsets[r0 + nb];
  const int32_t q0 = threadIdx.x < nb ? offsets[r] : 0, q1 = threadIdx.x < nb ? offsets[r + 1] : 0;
  if (q1 > q0)
  {
  const int64_t cell = parent_map[r];
  double X[TDIM + 1][TDIM];
  load_cell_coords<TDIM>(x, x_dofmap, cell, X);
  Geo<TDIM> g;
  make_geo<TDIM>(X, g);
  double xi[TDIM];
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
    xi[t] = pts[static_cast<int64_t>(t) * npts + q0];
  double phi[ND], dphi[ND][TDIM];
  tabulate<TDIM, 1>(xi, phi, dphi);
  double gref[TDIM];
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
    gref[t] = 0.0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
  {
    const double v = __ldg(vals + __ldg(ls_dofmap + cell * ND + j));
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      gref[t] += dphi[j][t] * v;
  }
  double gp[TDIM], nrm = 0.0;
#pragma unroll
  for (int i = 0; i < TDIM; ++i)
  {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      s += g.K[t * TDIM + i] * gref[t];
    gp[i] = s;
    nrm += s * s;
  }
  nrm = sqrt(nrm);
  if (nrm < 1.0e-14)
    nrm = 1.0e-14;
#pragma unroll
  for (int i = 0; i < TDIM; ++i)
    s_n[i][threadIdx.x] = sign * gp[i] / nrm;
  }
  __syncthreads();
  // the block's points: rule of point q = the last rule whose offset is <= q (empty rules share their offset with
  // the next one and are skipped by the upper bound)
  for (int32_t q = s_off[0] + threadIdx.x; q < s_off[nb]; q += NPB)
  {
    int lo = 0, hi = nb; // invariant: s_off[lo] <= q < s_off[hi]
    while (hi - lo > 1)
    {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] <= q)
        lo = mid;
      else
        hi = mid;
    }
#pragma unroll
    for (int i = 0; i < TDIM; ++i)
    {
      const double v = s_n[i][lo];
      out_soa[static_cast<int64_t>(i) * npts + q] = v;
      if (out_aos)
        out_aos[static_cast<int64_t>(q) * TDIM + i] = v;
    }
  }
}

template <int TDIM, int DEG>
__global__ void values_kernel(const double* __restrict__ pts, int64_t npts, const int32_t* __restrict__ offsets,
                              const int32_t* __restrict__ parent_map, int64_t nrules,
                              const int32_t* __restrict__ ls_dofmap, const double* __restrict__ vals,
                              double* __restrict__ out)
{
  constexpr int ND = Elem<TDIM, DEG>::ND;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= npts)
    return;
  const int64_t cell = parent_map[find_rule(offsets, nrules, q)];
  double xi[TDIM];
#pragma unroll
  for (int t = 0; t < TDIM; ++t)
    xi[t] = pts[static_cast<int64_t>(t) * npts + q];
  double phi[ND], dphi[ND][TDIM];
  tabulate<TDIM, DEG>(xi, phi, dphi);
  double v = 0.0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
    v += phi[j] * __ldg(vals + __ldg(ls_dofmap + cell * ND + j));
  out[q] = v;
}

template <int TDIM>
void run_quadrature(cfx_ctx* c, LevelSet& L, cfx_rules* R, bool positive, bool interface, RuleTable& rt)
{
  // deferred-size mode: a reused object keeps its buffers and the sizes stay on the device
  const bool defer = c->deferred && R->cap_pts >= 256 && R->cap_rules >= 256 && R->points.p != nullptr;
  if (!defer && L.cut_deferred)
  { // this call needs its sizes on the host, so it needs the cut-cell count too
    check_device_error(c, "cut-cell list (capacity exceeded)");
    L.n_cut = read_back(c, L.d_n_cut, 1)[0];
    L.cut_deferred = false;
  }
  const DN n_cut{L.cut_deferred ? L.d_n_cut : nullptr, L.n_cut};
  if (!R->d_sizes)
  {
    R->d_sizes = alloc_count_slot(c, 2);
    R->ctx = c;
  }
  // moments only where the point sums they replace are exact: built-in rule of degree >= 1, volume part
  // (interface rules also carry second moments: exact from degree 2 on)
  R->has_moments = rt.builtin && rt.order >= (interface ? 2 : 1);
  const size_t mom_w = 1 + TDIM + (interface ? TDIM * (TDIM + 1) / 2 : 0);
  if (n_cut.h == 0)
  {
    R->nrules = R->npts = 0;
    R->deferred = false;
    R->has_moments = false;
    R->points.reserve(c->pool, 1);
    R->weights.reserve(c->pool, 1);
    R->offsets.reserve(c->pool, 1);
    R->parent_map.reserve(c->pool, 1);
    CFX_CUDA(cudaMemsetAsync(R->offsets.p, 0, sizeof(int32_t), c->stream));
    CFX_CUDA(cudaMemsetAsync(R->d_sizes, 0, 2 * sizeof(int64_t), c->stream));
    return;
  }
  DevBuf<int64_t> packed, packed_excl;
  packed.reserve(c->pool, static_cast<size_t>(n_cut.h) + 1);
  packed_excl.reserve(c->pool, static_cast<size_t>(n_cut.h) + 2);
  const bool p2 = L.degree == 2;
  if (p2)
  {
    R->has_moments = false; // the pieces are not the case-table pieces of one simplex: kernels take the point loop
    CFX_LAUNCH(c, rule_count_p2_kernel<TDIM>, grid_for(n_cut.h, QB), QB, 0, L.cut_list.p, n_cut, L.dofmap, L.values,
               positive, interface, rt.npts, packed.p);
  }
  else
    CFX_LAUNCH(c, rule_count_kernel<TDIM>, grid_for(n_cut.h, QB), QB, 0, L.cut_list.p, n_cut, L.dofmap, L.values, positive,
               interface, rt.npts, packed.p);
  exclusive_scan_i64(c, packed.p, n_cut, packed_excl.p);
  if (defer)
  {
    R->nrules = R->cap_rules;
    R->npts = R->cap_pts;
    R->deferred = true;
    note_result(c, R);
  }
  else
  {
    const int64_t tot = read_back(c, c->scratch64.p, 1)[0];
    R->nrules = tot >> 32;
    R->npts = tot & 0xffffffffLL;
    R->deferred = false;
    CFX_REQUIRE(R->npts < (int64_t(1) << 31), CFX_ERR_RANGE, "runtime_quadrature: more than 2^31 points (int32 offsets)");
    if (R->npts > R->cap_pts || R->nrules > R->cap_rules || !R->points.p)
    {
      R->cap_pts = std::max(R->cap_pts, with_margin(c, R->npts));
      R->cap_rules = std::max(R->cap_rules, with_margin(c, R->nrules));
    }
  }
  R->points.reserve(c->pool, static_cast<size_t>(R->cap_pts) * TDIM + 1);
  R->weights.reserve(c->pool, static_cast<size_t>(R->cap_pts) + 1);
  R->offsets.reserve(c->pool, static_cast<size_t>(R->cap_rules) + 1);
  R->parent_map.reserve(c->pool, static_cast<size_t>(R->cap_rules) + 1);
  if (R->has_moments)
    R->moments.reserve(c->pool, static_cast<size_t>(R->cap_rules) * mom_w + 1);
  if (p2)
    CFX_LAUNCH(c, rule_fill_p2_kernel<TDIM>, grid_for(n_cut.h, QB), QB, 0, L.cut_list.p, n_cut, packed_excl.p, L.dofmap,
               L.values, c->x_dofmap, c->x, positive, interface, rt.npts, rt.d_pts, rt.d_wts, R->cap_pts, R->cap_rules,
               R->d_sizes, c->err_flag.p, R->points.p, R->weights.p, R->offsets.p, R->parent_map.p);
  else if (interface)
    CFX_LAUNCH(c, (rule_fill_kernel<TDIM, true>), grid_for(n_cut.h, QB), QB, 0, L.cut_list.p, n_cut, packed_excl.p,
               L.dofmap, L.values, c->x_dofmap, c->x, positive, rt.npts, rt.d_pts, rt.d_wts, R->cap_pts, R->cap_rules,
               R->d_sizes, c->err_flag.p, R->points.p, R->weights.p, R->offsets.p, R->parent_map.p,
               R->has_moments ? R->moments.p : nullptr);
  else
    CFX_LAUNCH(c, (rule_fill_kernel<TDIM, false>), grid_for(n_cut.h, QB), QB, 0, L.cut_list.p, n_cut, packed_excl.p,
               L.dofmap, L.values, c->x_dofmap, c->x, positive, rt.npts, rt.d_pts, rt.d_wts, R->cap_pts, R->cap_rules,
               R->d_sizes, c->err_flag.p, R->points.p, R->weights.p, R->offsets.p, R->parent_map.p,
               R->has_moments ? R->moments.p : nullptr);
  packed.release();
  packed_excl.release();
}
} // namespace
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_simplex_rule(int dim, int order, int* npts, double* points, double* weights, int capacity)
{
  cfx_ctx* ctx = nullptr;
  CFX_API_BEGIN
  std::vector<double> p, w;
  builtin_simplex_rule(dim, order, p, w);
  if (npts)
    *npts = static_cast<int>(w.size());
  if (points && weights)
  {
    CFX_REQUIRE(capacity >= static_cast<int>(w.size()), CFX_ERR_RANGE, "cfx_simplex_rule: capacity too small");
    std::copy(p.begin(), p.end(), points);
    std::copy(w.begin(), w.end(), weights);
  }
  CFX_API_END(ctx)
}

cfx_status cfx_set_simplex_rule(cfx_ctx* ctx, int dim, int order, int npts, const double* points,
                                const double* weights)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && points && weights && npts > 0 && dim >= 1 && dim <= 3 && order >= 0, CFX_ERR_INVALID,
              "cfx_set_simplex_rule: invalid arguments");
  RuleTable t;
  t.dim = dim;
  t.order = order;
  t.npts = npts;
  t.builtin = false;
  t.pts.assign(points, points + static_cast<size_t>(npts) * dim);
  t.wts.assign(weights, weights + npts);
  auto key = std::make_pair(dim, order);
  auto it = ctx->rules.find(key);
  if (it != ctx->rules.end())
  {
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (it->second.d_pts)
      ctx->pool.free(it->second.d_pts);
    if (it->second.d_wts)
      ctx->pool.free(it->second.d_wts);
    ctx->rules.erase(it);
  }
  ctx->rules.emplace(key, std::move(t));
  CFX_API_END(ctx)
}

cfx_status cfx_runtime_quadrature(cfx_ctx* ctx, int ls, int relation, int order, cfx_rules** inout)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->classified, CFX_ERR_STATE, "cfx_runtime_quadrature: call cfx_update first");
  CFX_REQUIRE(inout != nullptr, CFX_ERR_INVALID, "cfx_runtime_quadrature: inout is NULL");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID,
              "cfx_runtime_quadrature: invalid level-set index");
  CFX_REQUIRE(relation >= CFX_REL_LT && relation <= CFX_REL_EQ, CFX_ERR_INVALID,
              "cfx_runtime_quadrature: invalid relation");
  CFX_REQUIRE(order >= 0, CFX_ERR_INVALID, "runtime_quadrature: order must be >= 0"); // cut.cpp:164-168
  LevelSet& L = ctx->ls[ls];
  const bool interface = relation == CFX_REL_EQ;
  const bool positive = relation == CFX_REL_GT || relation == CFX_REL_GE;
  RuleTable& rt = get_rule(ctx, interface ? ctx->tdim - 1 : ctx->tdim, order);
  ensure_cut_list(ctx, ls);
  if (*inout == nullptr)
    *inout = new cfx_rules();
  cfx_rules* R = *inout;
  R->tdim = ctx->tdim;
  R->gdim = ctx->gdim;
  R->relation = relation;
  R->entity_hosted = false;
  R->order = order;
  R->ls = ls;
  R->has_normals = false;
  {
    StageScope st(ctx, interface ? "quadrature_interface" : "quadrature_volume");
    if (ctx->tdim == 2)
      run_quadrature<2>(ctx, L, R, positive, interface, rt);
    else
      run_quadrature<3>(ctx, L, R, positive, interface, rt);
    st.set_bytes(36.0 * ctx->nv * static_cast<double>(L.n_cut)
                 + 8.0 * (ctx->tdim + 1) * static_cast<double>(R->npts) + 8.0 * static_cast<double>(R->nrules));
  }
  CFX_API_END(ctx)
}

namespace cfx
{
namespace
{
// RuntimeSurfaceProvenance of straight codimension-one rules (make_surface_provenance, cut.cpp:1273-1308): rule k
// comes from the k-th cut entity of the level set (the rules are emitted one per cut entity, ascending), its parent
// is parent_map[k], the entity has ONE zero entity per P1 level set (local id 0) of dimension (host dim - 1)
__global__ void provenance_kernel(const int32_t* __restrict__ parent_map, int64_t n, int dim,
                                  const int32_t* __restrict__ cut_list, const int64_t* __restrict__ d_n_cut,
                                  int64_t n_cut_host, int32_t* __restrict__ cut_ids, int32_t* __restrict__ parents,
                                  int32_t* __restrict__ local_ids, int32_t* __restrict__ dims)
{
  const int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (k >= n)
    return;
  const int32_t parent = parent_map[k];
  int64_t id = k;
  if (cut_list)
  { // position of the parent among the cut entities (a cut entity whose selected part is empty -- the zero set only
    // touches a vertex -- has no rule, so this is not the rule index in general)
    int64_t lo = 0, hi = d_n_cut ? *d_n_cut : n_cut_host;
    while (lo < hi)
    {
      const int64_t mid = (lo + hi) >> 1;
      if (cut_list[mid] < parent)
        lo = mid + 1;
      else
        hi = mid;
    }
    id = lo;
  }
  cut_ids[k] = static_cast<int32_t>(id);
  parents[k] = parent;
  local_ids[k] = 0;
  dims[k] = dim;
}
} // namespace
} // namespace cfx

cfx_status cfx_rules_surface_provenance(cfx_ctx* ctx, const cfx_rules* r, int32_t* level_set_index,
                                        int32_t* cut_cell_ids, int32_t* parent_cell_ids,
                                        int32_t* local_zero_entity_ids, int32_t* dimensions, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && r && level_set_index, CFX_ERR_INVALID, "cfx_rules_surface_provenance: NULL argument");
  resolve(ctx, const_cast<cfx_rules*>(r));
  // single_equality_level_set_index (cut.cpp:1258-1271): only "ls = 0" selectors carry provenance
  if (r->relation != CFX_REL_EQ)
  {
    *level_set_index = -1;
    return CFX_OK;
  }
  *level_set_index = r->ls;
  if (r->nrules == 0 || !cut_cell_ids)
    return CFX_OK;
  CFX_REQUIRE(parent_cell_ids && local_zero_entity_ids && dimensions, CFX_ERR_INVALID,
              "cfx_rules_surface_provenance: NULL output array");
  const size_t n = static_cast<size_t>(r->nrules);
  DevBuf<int32_t> tmp;
  int32_t* d[4] = {cut_cell_ids, parent_cell_ids, local_zero_entity_ids, dimensions};
  if (memspace != CFX_DEVICE)
  {
    tmp.reserve(ctx->pool, 4 * n);
    for (int q = 0; q < 4; ++q)
      d[q] = tmp.p + q * n;
  }
  const int32_t* cut_list = nullptr;
  const int64_t* d_n_cut = nullptr;
  int64_t n_cut = 0;
  if (!r->entity_hosted && ctx->classified && r->ls >= 0 && r->ls < CFX_MAX_LEVEL_SETS && ctx->ls[r->ls].bound)
  { // cell hosts: the ascending cut-cell list of the level set (facet hosts keep the rule index: their rules follow
    // the host list, one per cut facet)
    ensure_cut_list(ctx, r->ls);
    LevelSet& L = ctx->ls[r->ls];
    cut_list = L.cut_list.p;
    d_n_cut = L.cut_deferred ? L.d_n_cut : nullptr;
    n_cut = L.n_cut;
  }
  CFX_LAUNCH(ctx, provenance_kernel, grid_for(r->nrules, 256), 256, 0, r->parent_map.p, r->nrules, r->tdim - 1,
             cut_list, d_n_cut, n_cut, d[0], d[1], d[2], d[3]);
  if (memspace != CFX_DEVICE)
  {
    export_to(ctx, cut_cell_ids, d[0], n, CFX_HOST);
    export_to(ctx, parent_cell_ids, d[1], n, CFX_HOST);
    export_to(ctx, local_zero_entity_ids, d[2], n, CFX_HOST);
    export_to(ctx, dimensions, d[3], n, CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}

cfx_status cfx_rules_sizes(const cfx_rules* r, int64_t* npts, int64_t* nrules, int* tdim)
{
  if (!r)
    return CFX_ERR_INVALID;
  if (r->deferred && r->ctx)
  { // sizes are on the device: fetch them now (synchronises)
    try
    {
      resolve(r->ctx, const_cast<cfx_rules*>(r));
    }
    catch (const cfx::Error& e)
    {
      cfx_set_error(r->ctx, e.what());
      return e.code;
    }
  }
  if (npts)
    *npts = r->npts;
  if (nrules)
    *nrules = r->nrules;
  if (tdim)
    *tdim = r->tdim;
  return CFX_OK;
}

cfx_status cfx_rules_fetch(cfx_ctx* ctx, const cfx_rules* r, double* points_aos, double* weights, int32_t* offsets,
                           int32_t* parent_map, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && r, CFX_ERR_INVALID, "cfx_rules_fetch: NULL argument");
  resolve(ctx, const_cast<cfx_rules*>(r));
  if (points_aos && r->npts > 0)
  {
    const int64_t n = r->npts * r->tdim;
    if (memspace == CFX_DEVICE)
      CFX_LAUNCH(ctx, soa_to_aos_kernel, grid_for(n, 256), 256, 0, r->points.p, r->npts, r->tdim, points_aos);
    else
    {
      DevBuf<double> tmp;
      tmp.reserve(ctx->pool, static_cast<size_t>(n));
      CFX_LAUNCH(ctx, soa_to_aos_kernel, grid_for(n, 256), 256, 0, r->points.p, r->npts, r->tdim, tmp.p);
      export_to(ctx, points_aos, tmp.p, static_cast<size_t>(n), CFX_HOST);
      tmp.release();
    }
  }
  export_to(ctx, weights, r->weights.p, static_cast<size_t>(r->npts), memspace);
  export_to(ctx, offsets, r->offsets.p, static_cast<size_t>(r->nrules) + 1, memspace);
  export_to(ctx, parent_map, r->parent_map.p, static_cast<size_t>(r->nrules), memspace);
  CFX_API_END(ctx)
}

cfx_status cfx_rules_physical_points(cfx_ctx* ctx, const cfx_rules* r, double* out_soa, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && r && out_soa, CFX_ERR_INVALID, "cfx_rules_physical_points: NULL argument");
  resolve(ctx, const_cast<cfx_rules*>(r));
  if (r->npts == 0)
    return CFX_OK;
  DevBuf<double> tmp;
  double* dst = out_soa;
  if (memspace == CFX_HOST)
  {
    tmp.reserve(ctx->pool, static_cast<size_t>(r->npts) * r->gdim);
    dst = tmp.p;
  }
  if (r->entity_hosted)
    entity_physical_points(ctx, r, dst);
  else if (r->tdim == 2)
    CFX_LAUNCH(ctx, physical_points_kernel<2>, grid_for(r->npts, 256), 256, 0, r->points.p, r->npts, r->offsets.p,
               r->parent_map.p, r->nrules, ctx->x_dofmap, ctx->x, dst);
  else
    CFX_LAUNCH(ctx, physical_points_kernel<3>, grid_for(r->npts, 256), 256, 0, r->points.p, r->npts, r->offsets.p,
               r->parent_map.p, r->nrules, ctx->x_dofmap, ctx->x, dst);
  if (memspace == CFX_HOST)
  {
    export_to(ctx, out_soa, tmp.p, static_cast<size_t>(r->npts) * r->gdim, CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}

void cfx_rules_free(cfx_ctx* ctx, cfx_rules* r)
{
  (void)ctx;
  if (!r)
    return;
  r->points.release();
  r->weights.release();
  r->offsets.release();
  r->parent_map.release();
  r->normals.release();
  r->moments.release();
  r->rule_verts.release();
  free_count_slot(r->ctx, r->d_sizes, 2);
  delete r;
}

cfx_status cfx_evaluate_normals(cfx_ctx* ctx, int ls, cfx_rules* r, double sign, double* out_aos, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && r, CFX_ERR_INVALID, "Cannot evaluate normals without rules.");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID,
              "Cannot evaluate normals without a level set."); // normal.h:47-48
  CFX_REQUIRE(r->tdim == ctx->tdim, CFX_ERR_INVALID,
              "Normal evaluation points must have cell reference dimension."); // normal.h:62-63
  const LevelSet& L = ctx->ls[ls];
  if (out_aos || r->entity_hosted)
    resolve(ctx, r); // exporting needs the exact number of points on the host
  r->normals.reserve(ctx->pool, static_cast<size_t>(std::max(r->npts, r->cap_pts)) * r->gdim + 1);
  r->has_normals = true;
  if (r->npts == 0)
    return CFX_OK;
  const DN d_npts{r->deferred ? r->d_sizes + 1 : nullptr, r->npts};
  const DN d_nrules{r->deferred ? r->d_sizes : nullptr, r->nrules};
  DevBuf<double> tmp;
  double* aos = nullptr;
  if (out_aos)
  {
    if (memspace == CFX_DEVICE)
      aos = out_aos;
    else
    {
      tmp.reserve(ctx->pool, static_cast<size_t>(r->npts) * r->gdim);
      aos = tmp.p;
    }
  }
  {
    StageScope st(ctx, "normals", static_cast<double>(r->npts) * 8.0 * (2.0 * r->tdim));
    const unsigned g = grid_for(r->npts, 256);
#define NARGS                                                                                                          \
  r->points.p, d_npts, r->offsets.p, r->parent_map.p, d_nrules, ctx->x_dofmap, ctx->x, L.dofmap, L.values, sign,       \
      r->normals.p, aos
    if (L.degree == 1)
    { // constant gradient per cell: one thread per rule
      auto nk = ctx->tdim == 2 ? normals_p1_kernel<2> : normals_p1_kernel<3>;
      CFX_LAUNCH(ctx, nk, grid_for(r->nrules, NPB), NPB, 0, NARGS);
    }
    else
    {
      auto nk = ctx->tdim == 2 ? normals_kernel<2, 2> : normals_kernel<3, 2>;
      CFX_LAUNCH(ctx, nk, g, 256, 0, NARGS);
    }
#undef NARGS
  }
  if (out_aos && memspace == CFX_HOST)
  {
    export_to(ctx, out_aos, tmp.p, static_cast<size_t>(r->npts) * r->gdim, CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}

cfx_status cfx_evaluate_values(cfx_ctx* ctx, int ls, const cfx_rules* r, double* out, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && r && out, CFX_ERR_INVALID, "cfx_evaluate_values: NULL argument");
  resolve(ctx, const_cast<cfx_rules*>(r));
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID,
              "Cannot evaluate values without a level set.");
  const LevelSet& L = ctx->ls[ls];
  if (r->npts == 0)
    return CFX_OK;
  DevBuf<double> tmp;
  double* dst = out;
  if (memspace == CFX_HOST)
  {
    tmp.reserve(ctx->pool, static_cast<size_t>(r->npts));
    dst = tmp.p;
  }
  const unsigned g = grid_for(r->npts, 256);
#define VARGS r->points.p, r->npts, r->offsets.p, r->parent_map.p, r->nrules, L.dofmap, L.values, dst
  auto vk = ctx->tdim == 2 ? (L.degree == 1 ? values_kernel<2, 1> : values_kernel<2, 2>)
                           : (L.degree == 1 ? values_kernel<3, 1> : values_kernel<3, 2>);
  CFX_LAUNCH(ctx, vk, g, 256, 0, VARGS);
#undef VARGS
  if (memspace == CFX_HOST)
  {
    export_to(ctx, out, tmp.p, static_cast<size_t>(r->npts), CFX_HOST);
    tmp.release();
  }
  CFX_API_END(ctx)
}
} // extern "C"
