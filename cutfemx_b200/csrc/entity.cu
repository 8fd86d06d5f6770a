// entity.cu -- facets as hosts of the cut.
//
// Reference: cutfemx.cut(level_set, facets, entity_dim = tdim - 1) builds a mesh view over the listed facets
// (build_entity_mesh_view, cut.cpp:540-591: connectivity = entities_to_geometry, cell type one dimension down) and
// restricts every level set to them through an entity dofmap (build_entity_level_sets, cut.cpp:1022-1063;
// fem/entity_dofmap.cpp:11-88); locate_entities then returns facet ids (host_parent_index, cut.cpp:344-359) and
// runtime_quadrature rules whose points live in the FACET's reference coordinates (tdim - 1 of them) with physical
// weights and parent_map = facet ids (test_cut_api.py:171-188, :349-367, :424-501).
//
// Same algorithm as for cells one dimension down: a facet is classified by the signs of the level set at its
// vertices (P1 level sets: the facet's dofs are its vertices' dofs), an intersected facet is split with the
// marching-simplex case table of its own dimension, the rule of each sub-simplex is mapped to the facet's
// reference coordinates and scaled by the physical measure of the sub-simplex.  Vertex order of a facet: ascending
// vertex number (DOLFINx stores entity vertices sorted); the reference pins only sets, sums and shapes here.
#include "compact.cuh"
#include "element.cuh"

struct cfx_ecut
{
  int edim = 0;
  int64_t n = 0;
  int64_t stride = 0;                 // codes: (CFX_MAX_LEVEL_SETS, stride), stride a multiple of 16
  cfx::DevBuf<int32_t> entities;      // facet ids in the caller's order
  cfx::DevBuf<int32_t> verts;         // (n, edim + 1) mesh nodes, ascending per facet
  cfx::DevBuf<double> phi;            // (CFX_MAX_LEVEL_SETS, n, edim + 1) level-set values at those vertices
  cfx::DevBuf<int8_t> codes;
};

namespace cfx
{
namespace
{
constexpr int EB2 = 128;

struct LsView
{
  const int32_t* dofmap[CFX_MAX_LEVEL_SETS];
  const double* values[CFX_MAX_LEVEL_SETS];
  int nd[CFX_MAX_LEVEL_SETS];
  int bound[CFX_MAX_LEVEL_SETS];
};

// one thread per listed facet: its vertices (sorted), the level-set values there, the domain codes
template <int EDIM>
__global__ void __launch_bounds__(EB2)
    facet_hosts_kernel(const int32_t* __restrict__ facets, int64_t n, int64_t n_facets, const int32_t* __restrict__ f2c2,
                       const int32_t* __restrict__ c2f, const int32_t* __restrict__ x_dofmap, LsView ls, int64_t stride,
                       int32_t* __restrict__ verts, double* __restrict__ phi, int8_t* __restrict__ codes,
                       int32_t* __restrict__ err)
{
  constexpr int NV = EDIM + 2; // vertices of the cell
  constexpr int NE = EDIM + 1; // vertices of the facet
  const int64_t i = static_cast<int64_t>(blockIdx.x) * EB2 + threadIdx.x;
  if (i >= n)
    return;
  const int32_t f = facets[i];
  if (f < 0 || f >= n_facets)
  { // validate_local_entities (cut.cpp:560-561)
    err[0] = 15;
    err[1] = f;
    return;
  }
  const int64_t c = f2c2[2 * static_cast<int64_t>(f)];
  if (c < 0)
  { // an index no cell refers to is not a facet of this mesh
    err[0] = 15;
    err[1] = f;
    return;
  }
  int lf = 0;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    lf = (c2f[c * NV + k] == f) ? k : lf;
  // facet lf is opposite local vertex lf (P1 simplex convention): its vertices are the others
  int32_t v[NE];
  int lj[NE];
  {
    int m = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (j != lf)
      {
#pragma unroll
        for (int q = 0; q < NE; ++q)
          if (q == m)
          {
            v[q] = x_dofmap[c * NV + j];
            lj[q] = j;
          }
        ++m;
      }
  }
  // ascending vertex number (tiny network)
#pragma unroll
  for (int pass = 0; pass < NE; ++pass)
#pragma unroll
    for (int q = pass & 1; q + 1 < NE; q += 2)
      if (v[q] > v[q + 1])
      {
        const int32_t tv = v[q];
        v[q] = v[q + 1];
        v[q + 1] = tv;
        const int tj = lj[q];
        lj[q] = lj[q + 1];
        lj[q + 1] = tj;
      }
#pragma unroll
  for (int q = 0; q < NE; ++q)
    verts[i * NE + q] = v[q];
  for (int l = 0; l < CFX_MAX_LEVEL_SETS; ++l)
  {
    if (!ls.bound[l])
      continue;
    bool all_neg = true, all_pos = true;
#pragma unroll
    for (int q = 0; q < NE; ++q)
    {
      // P1 level set: the dof of local vertex j of the cell
      const double val = ls.values[l][ls.dofmap[l][c * ls.nd[l] + lj[q]]];
      phi[(static_cast<int64_t>(l) * n + i) * NE + q] = val;
      all_neg = all_neg && (val < 0.0);
      all_pos = all_pos && (val > 0.0);
    }
    codes[static_cast<int64_t>(l) * stride + i]
        = static_cast<int8_t>(all_neg ? CFX_DOMAIN_INSIDE : (all_pos ? CFX_DOMAIN_OUTSIDE : CFX_DOMAIN_INTERSECTED));
  }
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ out)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n)
    out[i] = src[idx[i]];
}

// sub-simplices of a cut segment / triangle in its own reference coordinates.  Local points: inside vertices
// (ascending), then the edge cuts in (inside, outside) order -- the order of the cell generator (quadrature.cu).
template <int EDIM>
struct ECut
{
  static constexpr int NE = EDIM + 1;
  int n_in = 0, nsub = 0;
  double P[4][EDIM > 1 ? EDIM : 1];
  int sub[2][NE];
  int ifc[EDIM > 1 ? EDIM : 1]; // the interface simplex (dimension EDIM - 1) inside the facet: local point indices
};

template <int EDIM>
__device__ __forceinline__ void ecut_build(const double* phi, bool positive, ECut<EDIM>& E)
{
  constexpr int NE = EDIM + 1;
  int I[NE], O[NE], ni = 0, no = 0;
#pragma unroll
  for (int q = 0; q < NE; ++q)
  {
    const bool in = positive ? (phi[q] > 0.0) : (phi[q] < 0.0);
    if (in)
      I[ni++] = q;
    else
      O[no++] = q;
  }
  E.n_in = ni;
  E.nsub = 0;
  if (ni == 0 || ni == NE)
    return;
  auto refv = [&](int q, int t) { return (q == t + 1) ? 1.0 : 0.0; };
  int np = 0;
  for (int a = 0; a < ni; ++a, ++np)
#pragma unroll
    for (int t = 0; t < EDIM; ++t)
      E.P[np][t] = refv(I[a], t);
  for (int a = 0; a < ni; ++a)
    for (int b = 0; b < no; ++b, ++np)
    {
      const double tp = phi[I[a]] / (phi[I[a]] - phi[O[b]]);
#pragma unroll
      for (int t = 0; t < EDIM; ++t)
        E.P[np][t] = refv(I[a], t) + tp * (refv(O[b], t) - refv(I[a], t));
    }
  if constexpr (EDIM == 1)
  { // one inside vertex: the segment from it to the cut point; the interface is the cut point
    E.nsub = 1;
    E.sub[0][0] = 0;
    E.sub[0][1] = 1;
    E.ifc[0] = 1;
  }
  else
  { // triangle: 1 inside -> (I0, c00, c01); 2 inside -> (I0, I1, c10), (I0, c10, c00)   [tri case table]
    // interface segment: the two edge cuts (tri interface table {1,2} / {2,3})
    E.ifc[0] = ni == 1 ? 1 : 2;
    E.ifc[1] = ni == 1 ? 2 : 3;
    if (ni == 1)
    {
      E.nsub = 1;
      E.sub[0][0] = 0;
      E.sub[0][1] = 1;
      E.sub[0][2] = 2;
    }
    else
    {
      E.nsub = 2;
      E.sub[0][0] = 0;
      E.sub[0][1] = 1;
      E.sub[0][2] = 3;
      E.sub[1][0] = 0;
      E.sub[1][1] = 3;
      E.sub[1][2] = 2;
    }
  }
}

template <int EDIM>
__global__ void __launch_bounds__(EB2)
    ecut_count_kernel(const int8_t* __restrict__ codes, const double* __restrict__ phi, int64_t n, bool positive,
                      bool interface, int npts_s, int64_t* __restrict__ packed)
{
  constexpr int NE = EDIM + 1;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * EB2 + threadIdx.x;
  if (i >= n)
    return;
  int64_t pk = 0;
  if (codes[i] == CFX_DOMAIN_INTERSECTED)
  {
    ECut<EDIM> E;
    ecut_build<EDIM>(phi + i * NE, positive, E);
    if (E.nsub > 0)
      pk = (int64_t(1) << 32) | static_cast<int64_t>((interface ? 1 : E.nsub) * npts_s);
  }
  packed[i] = pk;
}

template <int EDIM>
__global__ void __launch_bounds__(EB2)
    ecut_fill_kernel(const int32_t* __restrict__ entities, const int32_t* __restrict__ verts,
                     const double* __restrict__ phi, const int8_t* __restrict__ codes, int64_t n,
                     const int64_t* __restrict__ packed_excl, bool positive, bool interface, int npts_s,
                     const double* __restrict__ rule_pts, const double* __restrict__ rule_wts,
                     const double* __restrict__ x, int64_t npts_total, double* __restrict__ points,
                     double* __restrict__ weights, int32_t* __restrict__ offsets, int32_t* __restrict__ parent_map,
                     int32_t* __restrict__ rule_verts)
{
  constexpr int NE = EDIM + 1;
  constexpr int GD = EDIM + 1; // geometric dimension of the mesh the facet lives in
  const int64_t i = static_cast<int64_t>(blockIdx.x) * EB2 + threadIdx.x;
  if (i >= n)
    return;
  const int64_t pk = packed_excl[i], pkn = packed_excl[i + 1];
  if (i == n - 1)
    offsets[pkn >> 32] = static_cast<int32_t>(npts_total);
  if ((pkn >> 32) == (pk >> 32))
    return;
  const int64_t rule = pk >> 32;
  const int64_t p0 = pk & 0xffffffffLL;
  offsets[rule] = static_cast<int32_t>(p0);
  parent_map[rule] = entities[i];
  double X[NE][GD];
#pragma unroll
  for (int q = 0; q < NE; ++q)
  {
    const int32_t node = verts[i * NE + q];
    rule_verts[rule * NE + q] = node;
#pragma unroll
    for (int d = 0; d < GD; ++d)
      X[q][d] = x[3 * static_cast<int64_t>(node) + d];
  }
  // physical measure factor of the facet: |e1| (segment), |e1 x e2| (triangle) = measure * edim!
  double J;
  if constexpr (EDIM == 1)
  {
    const double dx = X[1][0] - X[0][0], dy = X[1][1] - X[0][1];
    J = sqrt(dx * dx + dy * dy);
  }
  else
  {
    const double u0 = X[1][0] - X[0][0], u1 = X[1][1] - X[0][1], u2 = X[1][2] - X[0][2];
    const double w0 = X[2][0] - X[0][0], w1 = X[2][1] - X[0][1], w2 = X[2][2] - X[0][2];
    const double cx = u1 * w2 - u2 * w1, cy = u2 * w0 - u0 * w2, cz = u0 * w1 - u1 * w0;
    J = sqrt(cx * cx + cy * cy + cz * cz);
  }
  ECut<EDIM> E;
  ecut_build<EDIM>(phi + i * NE, positive, E);
  (void)codes;
  int64_t gp = p0;
  if (interface)
  { // phi = 0 inside the facet: a point (segment facets, weight 1) or a segment (triangle facets, physical length)
    if constexpr (EDIM == 1)
    {
      points[gp] = E.P[E.ifc[0]][0];
      weights[gp] = 1.0;
    }
    else
    {
      double Xa[GD], Xb[GD];
#pragma unroll
      for (int d = 0; d < GD; ++d)
      {
        const double a0 = E.P[E.ifc[0]][0], a1 = E.P[E.ifc[0]][1], b0 = E.P[E.ifc[1]][0], b1 = E.P[E.ifc[1]][1];
        Xa[d] = (1.0 - a0 - a1) * X[0][d] + a0 * X[1][d] + a1 * X[2][d];
        Xb[d] = (1.0 - b0 - b1) * X[0][d] + b0 * X[1][d] + b1 * X[2][d];
      }
      double len = 0.0;
#pragma unroll
      for (int d = 0; d < GD; ++d)
        len += (Xb[d] - Xa[d]) * (Xb[d] - Xa[d]);
      len = sqrt(len);
      for (int q = 0; q < npts_s; ++q, ++gp)
      {
        const double lam = rule_pts[q]; // 1-D rule on [0, 1]
#pragma unroll
        for (int d = 0; d < EDIM; ++d)
          points[static_cast<int64_t>(d) * npts_total + gp] = (1.0 - lam) * E.P[E.ifc[0]][d] + lam * E.P[E.ifc[1]][d];
        weights[gp] = rule_wts[q] * len;
      }
    }
    return;
  }
  for (int s = 0; s < E.nsub; ++s)
  {
    double det;
    if constexpr (EDIM == 1)
      det = E.P[E.sub[s][1]][0] - E.P[E.sub[s][0]][0];
    else
    {
      const double a0 = E.P[E.sub[s][1]][0] - E.P[E.sub[s][0]][0], a1 = E.P[E.sub[s][1]][1] - E.P[E.sub[s][0]][1];
      const double b0 = E.P[E.sub[s][2]][0] - E.P[E.sub[s][0]][0], b1 = E.P[E.sub[s][2]][1] - E.P[E.sub[s][0]][1];
      det = a0 * b1 - a1 * b0;
    }
    const double scale = fabs(det) * J;
    for (int q = 0; q < npts_s; ++q, ++gp)
    {
      double lam[EDIM], l0 = 1.0;
#pragma unroll
      for (int t = 0; t < EDIM; ++t)
      {
        lam[t] = rule_pts[q * EDIM + t];
        l0 -= lam[t];
      }
#pragma unroll
      for (int d = 0; d < EDIM; ++d)
      {
        double v = l0 * E.P[E.sub[s][0]][d];
#pragma unroll
        for (int t = 0; t < EDIM; ++t)
          v += lam[t] * E.P[E.sub[s][t + 1]][d];
        points[static_cast<int64_t>(d) * npts_total + gp] = v;
      }
      weights[gp] = rule_wts[q] * scale;
    }
  }
}

// physical points of facet-hosted rules: x = sum_q lambda_q X_q over the facet's vertices
template <int EDIM>
__global__ void entity_physical_points_kernel(const double* __restrict__ pts, int64_t npts,
                                              const int32_t* __restrict__ offsets, int64_t nrules,
                                              const int32_t* __restrict__ rule_verts, const double* __restrict__ x,
                                              double* __restrict__ out)
{
  constexpr int NE = EDIM + 1, GD = EDIM + 1;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= npts)
    return;
  const int64_t r = find_rule(offsets, nrules, q);
  double lam[NE];
  lam[0] = 1.0;
#pragma unroll
  for (int t = 0; t < EDIM; ++t)
  {
    lam[t + 1] = pts[static_cast<int64_t>(t) * npts + q];
    lam[0] -= lam[t + 1];
  }
#pragma unroll
  for (int d = 0; d < GD; ++d)
  {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < NE; ++k)
      v += lam[k] * x[3 * static_cast<int64_t>(rule_verts[r * NE + k]) + d];
    out[static_cast<int64_t>(d) * npts + q] = v;
  }
}
} // namespace

void entity_physical_points(cfx_ctx* ctx, const cfx_rules* r, double* dst)
{
  if (r->tdim == 1)
    CFX_LAUNCH(ctx, entity_physical_points_kernel<1>, grid_for(r->npts, 256), 256, 0, r->points.p, r->npts,
               r->offsets.p, r->nrules, r->rule_verts.p, ctx->x, dst);
  else
    CFX_LAUNCH(ctx, entity_physical_points_kernel<2>, grid_for(r->npts, 256), 256, 0, r->points.p, r->npts,
               r->offsets.p, r->nrules, r->rule_verts.p, ctx->x, dst);
}
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_cut_facets(cfx_ctx* ctx, const int32_t* facets, int64_t n, int memspace, cfx_ecut** inout)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound && inout, CFX_ERR_INVALID, "cfx_cut_facets: NULL argument");
  CFX_REQUIRE(ctx->topo_bound, CFX_ERR_STATE, "Facet-cell connectivity is unavailable.");
  CFX_REQUIRE(n >= 0 && (facets != nullptr || n == 0), CFX_ERR_INVALID, "cfx_cut_facets: NULL facets");
  LsView lv{};
  bool any = false;
  for (int l = 0; l < CFX_MAX_LEVEL_SETS; ++l)
  {
    const LevelSet& L = ctx->ls[l];
    lv.bound[l] = L.bound ? 1 : 0;
    lv.dofmap[l] = L.dofmap;
    lv.values[l] = L.values;
    lv.nd[l] = L.nd;
    if (L.bound)
    {
      any = true;
      CFX_REQUIRE(L.degree == 1, CFX_ERR_UNSUPPORTED, "facet-hosted cuts need P1 level sets");
    }
  }
  CFX_REQUIRE(any, CFX_ERR_STATE, "cfx_cut_facets: no level set bound");
  if (*inout == nullptr)
    *inout = new cfx_ecut();
  cfx_ecut* E = *inout;
  E->edim = ctx->tdim - 1;
  E->n = n;
  E->stride = (n + 15) / 16 * 16 + 16;
  const int ne = E->edim + 1;
  E->entities.reserve(ctx->pool, static_cast<size_t>(n) + 1);
  E->verts.reserve(ctx->pool, static_cast<size_t>(n) * ne + 1);
  E->phi.reserve(ctx->pool, static_cast<size_t>(CFX_MAX_LEVEL_SETS) * n * ne + 1);
  E->codes.reserve(ctx->pool, static_cast<size_t>(CFX_MAX_LEVEL_SETS) * E->stride);
  CFX_CUDA(cudaMemsetAsync(E->codes.p, 0, static_cast<size_t>(CFX_MAX_LEVEL_SETS) * E->stride, ctx->stream));
  if (n > 0)
  {
    CFX_CUDA(cudaMemcpyAsync(E->entities.p, facets, static_cast<size_t>(n) * sizeof(int32_t),
                             memspace == CFX_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (E->edim == 1)
      CFX_LAUNCH(ctx, facet_hosts_kernel<1>, grid_for(n, EB2), EB2, 0, E->entities.p, n, ctx->n_facets, ctx->f2c2.p,
                 ctx->c2f, ctx->x_dofmap, lv, E->stride, E->verts.p, E->phi.p, E->codes.p, ctx->err_flag.p);
    else
      CFX_LAUNCH(ctx, facet_hosts_kernel<2>, grid_for(n, EB2), EB2, 0, E->entities.p, n, ctx->n_facets, ctx->f2c2.p,
                 ctx->c2f, ctx->x_dofmap, lv, E->stride, E->verts.p, E->phi.p, E->codes.p, ctx->err_flag.p);
    check_device_error(ctx, "cfx_cut_facets (facet index out of range)");
  }
  CFX_API_END(ctx)
}

cfx_status cfx_ecut_locate(cfx_ctx* ctx, const cfx_ecut* E, int n_terms, const int32_t* term_offsets,
                           const int32_t* clause_ls, const int32_t* clause_rel, cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && E && out, CFX_ERR_INVALID, "cfx_ecut_locate: NULL argument");
  const Dnf d = make_dnf(ctx, n_terms, term_offsets, clause_ls, clause_rel);
  if (*out == nullptr)
    *out = new cfx_list();
  DevBuf<int32_t> pos;
  DnfPred p{d, E->codes.p, E->stride};
  const int64_t m = compact_indices(ctx, E->n, p, pos);
  (*out)->n = m;
  (*out)->data.reserve(ctx->pool, static_cast<size_t>(m > 0 ? m : 1));
  if (m > 0)
    CFX_LAUNCH(ctx, gather_i32_kernel, grid_for(m, 256), 256, 0, E->entities.p, pos.p, m, (*out)->data.p);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  pos.release();
  CFX_API_END(ctx)
}

cfx_status cfx_ecut_runtime_quadrature(cfx_ctx* ctx, const cfx_ecut* E, int ls, int relation, int order,
                                       cfx_rules** inout)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && E && inout, CFX_ERR_INVALID, "cfx_ecut_runtime_quadrature: NULL argument");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID,
              "cfx_ecut_runtime_quadrature: invalid level-set index");
  CFX_REQUIRE(relation >= CFX_REL_LT && relation <= CFX_REL_EQ, CFX_ERR_INVALID,
              "cfx_ecut_runtime_quadrature: invalid relation");
  CFX_REQUIRE(order >= 0, CFX_ERR_INVALID, "runtime_quadrature: order must be >= 0"); // cut.cpp:164-168
  if (*inout == nullptr)
    *inout = new cfx_rules();
  cfx_rules* R = *inout;
  const int edim = E->edim, ne = edim + 1;
  R->tdim = edim;
  R->gdim = ctx->gdim;
  R->relation = relation;
  R->order = order;
  R->ls = ls;
  R->has_normals = false;
  R->has_moments = false;
  R->entity_hosted = true;
  const bool positive = relation == CFX_REL_GT || relation == CFX_REL_GE;
  const bool interface = relation == CFX_REL_EQ;
  // the interface inside a segment facet is a point: one "rule point" of weight 1, no table needed
  RuleTable* rtp = (interface && edim == 1) ? nullptr : &get_rule(ctx, interface ? edim - 1 : edim, order);
  const int npts_s = rtp ? rtp->npts : 1;
  const double* d_pts = rtp ? rtp->d_pts : nullptr;
  const double* d_wts = rtp ? rtp->d_wts : nullptr;
  const int64_t n = E->n;
  R->nrules = R->npts = 0;
  R->offsets.reserve(ctx->pool, static_cast<size_t>(n) + 2);
  CFX_CUDA(cudaMemsetAsync(R->offsets.p, 0, sizeof(int32_t), ctx->stream));
  if (n == 0)
    return CFX_OK;
  DevBuf<int64_t> packed, packed_excl;
  packed.reserve(ctx->pool, static_cast<size_t>(n) + 1);
  packed_excl.reserve(ctx->pool, static_cast<size_t>(n) + 2);
  const int8_t* codes = E->codes.p + static_cast<size_t>(ls) * E->stride;
  const double* phi = E->phi.p + static_cast<size_t>(ls) * n * ne;
  if (edim == 1)
    CFX_LAUNCH(ctx, ecut_count_kernel<1>, grid_for(n, EB2), EB2, 0, codes, phi, n, positive, interface, npts_s, packed.p);
  else
    CFX_LAUNCH(ctx, ecut_count_kernel<2>, grid_for(n, EB2), EB2, 0, codes, phi, n, positive, interface, npts_s, packed.p);
  exclusive_scan_i64(ctx, packed.p, n, packed_excl.p);
  const int64_t tot = read_back(ctx, ctx->scratch64.p, 1)[0];
  R->nrules = tot >> 32;
  R->npts = tot & 0xffffffffLL;
  R->points.reserve(ctx->pool, static_cast<size_t>(R->npts) * edim + 1);
  R->weights.reserve(ctx->pool, static_cast<size_t>(R->npts) + 1);
  R->parent_map.reserve(ctx->pool, static_cast<size_t>(R->nrules) + 1);
  R->rule_verts.reserve(ctx->pool, static_cast<size_t>(R->nrules) * ne + 1);
  if (edim == 1)
    CFX_LAUNCH(ctx, ecut_fill_kernel<1>, grid_for(n, EB2), EB2, 0, E->entities.p, E->verts.p, phi, codes, n,
               packed_excl.p, positive, interface, npts_s, d_pts, d_wts, ctx->x, R->npts, R->points.p, R->weights.p,
               R->offsets.p, R->parent_map.p, R->rule_verts.p);
  else
    CFX_LAUNCH(ctx, ecut_fill_kernel<2>, grid_for(n, EB2), EB2, 0, E->entities.p, E->verts.p, phi, codes, n,
               packed_excl.p, positive, interface, npts_s, d_pts, d_wts, ctx->x, R->npts, R->points.p, R->weights.p,
               R->offsets.p, R->parent_map.p, R->rule_verts.p);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  packed.release();
  packed_excl.release();
  CFX_API_END(ctx)
}

void cfx_ecut_free(cfx_ctx* ctx, cfx_ecut* E)
{
  (void)ctx;
  if (!E)
    return;
  E->entities.release();
  E->verts.release();
  E->phi.release();
  E->codes.release();
  delete E;
}
} // extern "C"

namespace cfx
{
namespace
{
// Exterior-facet run-time rules for kernels defined on cells (_facet_payload_with_rows +
// facet_runtime_quadrature_payload, _runintgen_adapter.py:605-680): a facet rule is re-expressed in the reference
// coordinates of the facet's (first) cell.  key[k] = local index of the facet in that cell.
__global__ void facet_rule_key_kernel(const int32_t* __restrict__ parent_map, int64_t n, const int32_t* __restrict__ f2c2,
                                      const int32_t* __restrict__ c2f, int nf, int32_t* __restrict__ key,
                                      int32_t* __restrict__ cell_of)
{
  const int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (k >= n)
    return;
  const int32_t f = parent_map[k];
  const int32_t c = f2c2[2 * static_cast<int64_t>(f)];
  int lf = 0;
  for (int j = 0; j < nf; ++j)
    lf = (c2f[static_cast<int64_t>(c) * nf + j] == f) ? j : lf;
  key[k] = lf;
  cell_of[k] = c;
}
struct KeyPred
{
  const int32_t* key;
  int32_t want;
  __device__ unsigned operator()(int64_t base, int64_t n) const
  { // 16-bit mask for base .. base + 15 (compact.cuh)
    unsigned m = 0;
    for (int k = 0; k < 16; ++k)
      if (base + k < n && key[base + k] == want)
        m |= 1u << k;
    return m;
  }
};
__global__ void facet_rule_count_kernel(const int32_t* __restrict__ idx, int64_t m, const int32_t* __restrict__ offsets,
                                        int32_t* __restrict__ cnt)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < m)
    cnt[i] = offsets[idx[i] + 1] - offsets[idx[i]];
}
// one thread per selected rule: points mapped facet reference -> cell reference
//   x = sum_j lam_j v_j (facet vertices v_j ascending; lam_0 = 1 - sum xi, lam_j = xi_{j-1})  and vertex v of the cell
//   with local index t has cell reference coordinates e_t (e_0 = 0): X_{t-1} += lam_j for t = loc(v_j) >= 1
template <int EDIM>
__global__ void facet_rule_map_kernel(const int32_t* __restrict__ idx, int64_t m, const int32_t* __restrict__ offsets,
                                      const double* __restrict__ pts, int64_t npts_in,
                                      const double* __restrict__ wts, const int32_t* __restrict__ rule_verts,
                                      const int32_t* __restrict__ cell_of, const int32_t* __restrict__ x_dofmap,
                                      const int64_t* __restrict__ out_off64, int64_t npts_out,
                                      double* __restrict__ out_pts, double* __restrict__ out_wts,
                                      int32_t* __restrict__ out_offsets, int32_t* __restrict__ out_parent)
{
  constexpr int TDIM = EDIM + 1, NE = EDIM + 1, NV = TDIM + 1;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= m)
    return;
  const int64_t k = idx[i];
  const int32_t c = cell_of[k];
  int loc[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j)
  {
    const int32_t v = rule_verts[k * NE + j];
    int t = 0;
#pragma unroll
    for (int q = 0; q < NV; ++q)
      t = (x_dofmap[static_cast<int64_t>(c) * NV + q] == v) ? q : t;
    loc[j] = t;
  }
  const int64_t o = out_off64[i];
  out_offsets[i] = static_cast<int32_t>(o);
  if (i == m - 1)
    out_offsets[m] = static_cast<int32_t>(npts_out);
  out_parent[i] = c;
  const int32_t q0 = offsets[k], q1 = offsets[k + 1];
  for (int32_t q = q0; q < q1; ++q)
  {
    double lam[NE];
    lam[0] = 1.0;
#pragma unroll
    for (int j = 1; j < NE; ++j)
    {
      lam[j] = pts[static_cast<int64_t>(j - 1) * npts_in + q];
      lam[0] -= lam[j];
    }
    double X[TDIM];
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      X[t] = 0.0;
#pragma unroll
    for (int j = 0; j < NE; ++j)
#pragma unroll
      for (int t = 0; t < TDIM; ++t)
        X[t] += (loc[j] == t + 1) ? lam[j] : 0.0;
    const int64_t oq = o + (q - q0);
#pragma unroll
    for (int t = 0; t < TDIM; ++t)
      out_pts[static_cast<int64_t>(t) * npts_out + oq] = X[t];
    out_wts[oq] = wts[q];
  }
}
} // namespace
} // namespace cfx

extern "C" cfx_status cfx_rules_facets_to_cells(cfx_ctx* ctx, const cfx_rules* fr, int local_facet, cfx_rules** inout)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && fr && inout, CFX_ERR_INVALID, "cfx_rules_facets_to_cells: NULL argument");
  CFX_REQUIRE(fr->entity_hosted && fr->tdim == ctx->tdim - 1, CFX_ERR_INVALID,
              "cfx_rules_facets_to_cells: the rules must come from a facet-hosted cut");
  CFX_REQUIRE(ctx->c2f != nullptr && ctx->f2c2.p != nullptr, CFX_ERR_STATE,
              "cfx_rules_facets_to_cells: bind the topology first (cfx_topology_bind)");
  CFX_REQUIRE(local_facet >= 0 && local_facet <= ctx->tdim, CFX_ERR_INVALID, "local facet index out of range");
  if (*inout == nullptr)
    *inout = new cfx_rules();
  cfx_rules* R = *inout;
  R->tdim = ctx->tdim;
  R->gdim = ctx->gdim;
  R->relation = fr->relation;
  R->order = fr->order;
  R->ls = fr->ls;
  R->has_normals = false;
  R->has_moments = false;
  R->entity_hosted = false;
  R->deferred = false;
  R->ctx = ctx;
  R->nrules = R->npts = 0;
  R->offsets.reserve(ctx->pool, 2);
  CFX_CUDA(cudaMemsetAsync(R->offsets.p, 0, sizeof(int32_t), ctx->stream));
  const int64_t n = fr->nrules;
  if (n == 0)
    return CFX_OK;
  DevBuf<int32_t> key, cell_of, idx, cnt;
  DevBuf<int64_t> off64;
  key.reserve(ctx->pool, static_cast<size_t>(n) + 1);
  cell_of.reserve(ctx->pool, static_cast<size_t>(n) + 1);
  CFX_LAUNCH(ctx, facet_rule_key_kernel, grid_for(n, 256), 256, 0, fr->parent_map.p, n, ctx->f2c2.p, ctx->c2f,
             ctx->tdim + 1, key.p, cell_of.p);
  KeyPred pred{key.p, local_facet};
  const int64_t m = compact_indices(ctx, n, pred, idx);
  if (m > 0)
  {
    cnt.reserve(ctx->pool, static_cast<size_t>(m) + 1);
    off64.reserve(ctx->pool, static_cast<size_t>(m) + 2);
    CFX_LAUNCH(ctx, facet_rule_count_kernel, grid_for(m, 256), 256, 0, idx.p, m, fr->offsets.p, cnt.p);
    exclusive_scan_i32_to_i64(ctx, cnt.p, m, off64.p);
    const int64_t npts = read_back(ctx, ctx->scratch64.p, 1)[0];
    R->nrules = m;
    R->npts = npts;
    R->points.reserve(ctx->pool, static_cast<size_t>(npts) * ctx->tdim + 1);
    R->weights.reserve(ctx->pool, static_cast<size_t>(npts) + 1);
    R->offsets.reserve(ctx->pool, static_cast<size_t>(m) + 2);
    R->parent_map.reserve(ctx->pool, static_cast<size_t>(m) + 1);
    if (ctx->tdim == 2)
      CFX_LAUNCH(ctx, facet_rule_map_kernel<1>, grid_for(m, 256), 256, 0, idx.p, m, fr->offsets.p, fr->points.p, fr->npts,
                 fr->weights.p, fr->rule_verts.p, cell_of.p, ctx->x_dofmap, off64.p, npts, R->points.p, R->weights.p,
                 R->offsets.p, R->parent_map.p);
    else
      CFX_LAUNCH(ctx, facet_rule_map_kernel<2>, grid_for(m, 256), 256, 0, idx.p, m, fr->offsets.p, fr->points.p, fr->npts,
                 fr->weights.p, fr->rule_verts.p, cell_of.p, ctx->x_dofmap, off64.p, npts, R->points.p, R->weights.p,
                 R->offsets.p, R->parent_map.p);
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  key.release();
  cell_of.release();
  idx.release();
  cnt.release();
  off64.release();
  CFX_API_END(ctx)
}
