// facets.cu -- ghost-penalty facet band, interior facets of a cell set, facet integration rows.
//
// Replaces cutfemx.ghost_penalty_facets (python/cutfemx/cut.py:340-380, a pure-Python set loop),
// cutfemx::interior_facets_for_cells (cpp/cutfemx/cut/cut.cpp:926-994) and
// facet_integration_rows / local_facet_index (python/cutfemx/wrappers/cut.cpp:38-115).
//
// All integer work; results are sorted-unique facet ids exactly as the reference's
// sorted(set(...)) / sort+unique produce.  Marking a facet is an idempotent byte store (no
// atomics), the ascending list comes from the order-preserving compaction in compact.cuh.
// Roofline: HBM; bytes ~ Nc (active flags) + 4*(tdim+1)*Ncut (c2f rows) + Nf (facet flags).
#include "compact.cuh"

namespace cfx
{
namespace
{
constexpr int FB = 256;

// facet -> (min cell, max cell) by integer atomics: deterministic, any c2f numbering
__global__ void f2c_minmax_kernel(const int32_t* __restrict__ c2f, int64_t n_entries, int nf, int64_t n_facets,
                                  int32_t* __restrict__ f2c2, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (i >= n_entries)
    return;
  const int32_t f = c2f[i];
  if (f < 0 || f >= n_facets)
  {
    err[0] = 11;
    err[1] = f;
    return;
  }
  const int32_t cell = static_cast<int32_t>(i / nf);
  atomicMin(&f2c2[2 * static_cast<int64_t>(f)], cell);
  atomicMax(&f2c2[2 * static_cast<int64_t>(f) + 1], cell);
}

__global__ void f2c_fix_kernel(int64_t n_facets, int32_t* __restrict__ f2c2)
{
  const int64_t f = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (f >= n_facets)
    return;
  const int32_t lo = f2c2[2 * f], hi = f2c2[2 * f + 1];
  if (lo == 0x7fffffff)
  { // no cell
    f2c2[2 * f] = -1;
    f2c2[2 * f + 1] = -1;
  }
  else if (hi == lo)
    f2c2[2 * f + 1] = -1; // boundary facet
}

__global__ void f2c_init_kernel(int64_t n_facets, int32_t* __restrict__ f2c2)
{
  const int64_t f = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (f >= n_facets)
    return;
  f2c2[2 * f] = 0x7fffffff;
  f2c2[2 * f + 1] = -1;
}

__global__ void f2c_from_adj_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ data,
                                    int64_t n_facets, int32_t* __restrict__ f2c2, int32_t* __restrict__ err)
{
  const int64_t f = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (f >= n_facets)
    return;
  const int32_t b = off[f], e = off[f + 1];
  int32_t c0 = -1, c1 = -1;
  if (e - b >= 1)
    c0 = data[b];
  if (e - b >= 2)
    c1 = data[b + 1];
  if (e - b > 2)
  {
    err[0] = 12;
    err[1] = static_cast<int32_t>(f);
  }
  f2c2[2 * f] = c0;
  f2c2[2 * f + 1] = c1;
}

// active(c) = intersected by cut_ls  OR  selector matches   (cut.py:364-366), for ANY local cell
struct BandPredDnf
{
  Dnf d;
  int cut_ls;
  const int8_t* domain;
  int64_t stride;
  __device__ bool operator()(int64_t c) const
  {
    return domain[static_cast<int64_t>(cut_ls) * stride + c] == CFX_DOMAIN_INTERSECTED
           || dnf_match(d, domain, stride, c);
  }
};
struct BandPredFlags
{
  const uint8_t* flag;
  __device__ bool operator()(int64_t c) const { return flag[c] != 0; }
};

__global__ void set_flag_kernel(const int32_t* __restrict__ cells, int64_t n, int64_t nc, uint8_t* __restrict__ flag,
                                int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (i >= n)
    return;
  const int32_t c = cells[i];
  if (c < 0 || c >= nc)
  { // cut.cpp:963-964 "Cell index is out of range."
    err[0] = 13;
    err[1] = c;
    return;
  }
  flag[c] = 1;
}

// cut.py:369-379 / cut.cpp:967-987: for every source cell, every facet that is owned, has two
// cells, both flagged -> mark
template <class Active>
__global__ void mark_facets_kernel(const int32_t* __restrict__ src_cells, DN n_src_, int nf,
                                   const int32_t* __restrict__ c2f, const int32_t* __restrict__ f2c2, Active active,
                                   int64_t n_owned_facets, int include_ghosts, uint8_t* __restrict__ facet_flag,
                                   int64_t n_cells, int32_t* __restrict__ tile_counts)
{
  const int64_t n_src = n_src_.get();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (i >= n_src * nf)
    return;
  const int64_t cell = src_cells[i / nf];
  if (cell < 0 || cell >= n_cells)
    return; // caller-supplied list with a bad index: set_flag_kernel has raised the error flag
  const int32_t f = c2f[cell * nf + (i % nf)];
  if (!include_ghosts && f >= n_owned_facets)
    return;
  const int32_t c0 = f2c2[2 * static_cast<int64_t>(f)], c1 = f2c2[2 * static_cast<int64_t>(f) + 1];
  if (c1 < 0)
    return; // not exactly two cells
  if (active(c0) && active(c1))
  { // first writer of the flag also counts the facet for its compaction tile (integer atomics: exact totals), so
    // the compaction needs no counting pass over the whole facet flag array
    unsigned int* w = reinterpret_cast<unsigned int*>(facet_flag) + (f >> 2);
    const unsigned int bit = 1u << (8 * (f & 3));
    if (!(atomicOr(w, bit) & bit))
      atomicAdd(&tile_counts[f / CP_TILE], 1);
  }
}

__global__ void clear_flags_kernel(const int32_t* __restrict__ idx, DN n_, uint8_t* __restrict__ flag)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (i < n_.get())
    flag[idx[i]] = 0;
}

// deferred-size mode, capacity of the output list exceeded: the list comes back empty, so the flags it would have
// cleared are cleared here (a pass over the flag array that only runs when the list is empty: after an overflow,
// or -- redundantly -- when there is no band at all)
__global__ void clear_all_flags_if_empty_kernel(const int64_t* __restrict__ d_n, int64_t n_words,
                                                unsigned int* __restrict__ flag_words)
{
  if (*d_n != 0)
    return;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x; i < n_words;
       i += static_cast<int64_t>(gridDim.x) * FB)
    flag_words[i] = 0u;
}

__global__ void facet_rows_kernel(const int32_t* __restrict__ facets, DN n_, int nf, int64_t n_facets,
                                  const int32_t* __restrict__ c2f, const int32_t* __restrict__ f2c2,
                                  int32_t* __restrict__ rows4, int32_t* __restrict__ err,
                                  int64_t* __restrict__ d_n_out /* receives 4 n: entries of the row list */)
{
  const int64_t n = n_.get();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * FB + threadIdx.x;
  if (i == 0 && d_n_out)
    *d_n_out = 4 * n;
  if (i >= n)
    return;
  const int32_t f = facets[i];
  if (f < 0 || f >= n_facets)
  {
    err[0] = 14;
    err[1] = f;
    return;
  }
  const int32_t cs[2] = {f2c2[2 * static_cast<int64_t>(f)], f2c2[2 * static_cast<int64_t>(f) + 1]};
  if (cs[0] < 0 || cs[1] < 0)
  { // wrappers/cut.cpp:101-105 "Interior facet domain contains a facet without two adjacent cells."
    err[0] = 15;
    err[1] = f;
    return;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s)
  {
    int lf = -1;
    for (int k = 0; k < nf; ++k)
      if (c2f[static_cast<int64_t>(cs[s]) * nf + k] == f)
      {
        lf = k;
        break;
      }
    if (lf < 0)
    { // wrappers/cut.cpp:49-50 "Could not resolve local facet index."
      err[0] = 16;
      err[1] = f;
    }
    rows4[4 * i + 2 * s] = cs[s];
    rows4[4 * i + 2 * s + 1] = lf;
  }
}

template <class Active>
int64_t band_from_flags(cfx_ctx* c, const int32_t* src_cells, DN n_src, Active active, int include_ghosts,
                        cfx_list* out)
{
  if (!out->d_n)
  {
    out->d_n = alloc_count_slot(c);
    out->ctx = c;
  }
  const int nf = c->tdim + 1;
  const unsigned nt = grid_for(c->n_facets, CP_TILE);
  c->blk_counts.reserve(c->pool, nt);
  CFX_CUDA(cudaMemsetAsync(c->blk_counts.p, 0, static_cast<size_t>(nt) * sizeof(int32_t), c->stream));
  if (n_src.h > 0)
    CFX_LAUNCH(c, mark_facets_kernel<Active>, grid_for(n_src.h * nf, FB), FB, 0, src_cells, n_src, nf, c->c2f,
               c->f2c2.p, active, c->n_owned_facets, include_ghosts, c->facet_flag.p, c->nc_total, c->blk_counts.p);
  FlagPred p{c->facet_flag.p};
  out->n = compact_indices(c, dn_exact(c->n_facets), p, out->data, /*counted*/ true, out->d_n, &out->deferred);
  note_result(c, out);
  if (out->n > 0)
    CFX_LAUNCH(c, clear_flags_kernel, grid_for(out->n, FB), FB, 0, out->data.p, dn_of(out), c->facet_flag.p);
  if (out->deferred)
    CFX_LAUNCH(c, clear_all_flags_if_empty_kernel, 296, FB, 0, out->d_n, (c->n_facets + 4) / 4,
               reinterpret_cast<unsigned int*>(c->facet_flag.p));
  return out->n;
}
} // namespace

void derive_f2c(cfx_ctx* c)
{
  const int nf = c->tdim + 1;
  CFX_LAUNCH(c, f2c_init_kernel, grid_for(c->n_facets, FB), FB, 0, c->n_facets, c->f2c2.p);
  const int64_t n_entries = c->nc_total * nf;
  CFX_LAUNCH(c, f2c_minmax_kernel, grid_for(n_entries, FB), FB, 0, c->c2f, n_entries, nf, c->n_facets, c->f2c2.p,
             c->err_flag.p);
  CFX_LAUNCH(c, f2c_fix_kernel, grid_for(c->n_facets, FB), FB, 0, c->n_facets, c->f2c2.p);
}

void dense_f2c_from_adjacency(cfx_ctx* c, const int32_t* off_dev, const int32_t* data_dev)
{
  CFX_LAUNCH(c, f2c_from_adj_kernel, grid_for(c->n_facets, FB), FB, 0, off_dev, data_dev, c->n_facets, c->f2c2.p,
             c->err_flag.p);
}
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_ghost_penalty_facets(cfx_ctx* ctx, int cut_ls, int n_terms, const int32_t* term_offsets,
                                    const int32_t* clause_ls, const int32_t* clause_rel, int include_ghosts,
                                    cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->classified, CFX_ERR_STATE, "cfx_ghost_penalty_facets: call cfx_update first");
  CFX_REQUIRE(ctx->topo_bound, CFX_ERR_STATE, "Facet-cell connectivity is unavailable."); // cut.py:361-362
  CFX_REQUIRE(out != nullptr, CFX_ERR_INVALID, "cfx_ghost_penalty_facets: out is NULL");
  CFX_REQUIRE(cut_ls >= 0 && cut_ls < CFX_MAX_LEVEL_SETS && ctx->ls[cut_ls].bound, CFX_ERR_INVALID,
              "cfx_ghost_penalty_facets: invalid cut level set");
  const Dnf d = make_dnf(ctx, n_terms, term_offsets, clause_ls, clause_rel);
  ensure_cut_list_all(ctx, cut_ls);
  if (*out == nullptr)
    *out = new cfx_list();
  LevelSet& L = ctx->ls[cut_ls];
  const bool all = ctx->nc_total != ctx->nc_owned;
  const int32_t* src = all ? L.cut_list_all.p : L.cut_list.p;
  const DN n_src = all ? DN{L.cut_all_deferred ? L.d_n_cut_all : nullptr, L.n_cut_all}
                       : DN{L.cut_deferred ? L.d_n_cut : nullptr, L.n_cut};
  StageScope st(ctx, "ghost_penalty_facets",
                4.0 * (ctx->tdim + 1) * static_cast<double>(n_src.h) + 2.0 * static_cast<double>(ctx->n_facets));
  band_from_flags(ctx, src, n_src, BandPredDnf{d, cut_ls, ctx->domain.p, ctx->domain_stride}, include_ghosts, *out);
  CFX_API_END(ctx)
}

cfx_status cfx_interior_facets_for_cells(cfx_ctx* ctx, const int32_t* cells, int64_t n, int memspace,
                                         int include_ghosts, cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->topo_bound, CFX_ERR_STATE, "Facet-cell connectivity is unavailable."); // cut.cpp:945-946
  CFX_REQUIRE(out != nullptr && (cells != nullptr || n == 0), CFX_ERR_INVALID,
              "cfx_interior_facets_for_cells: NULL argument");
  if (*out == nullptr)
    *out = new cfx_list();
  DevBuf<int32_t> own;
  const int32_t* d_cells = n > 0 ? adopt(ctx, own, cells, static_cast<size_t>(n), memspace) : nullptr;
  ctx->scratch8.reserve(ctx->pool, static_cast<size_t>(ctx->nc_total) + 16);
  CFX_CUDA(cudaMemsetAsync(ctx->scratch8.p, 0, static_cast<size_t>(ctx->nc_total), ctx->stream));
  if (n > 0)
    CFX_LAUNCH(ctx, set_flag_kernel, grid_for(n, FB), FB, 0, d_cells, n, ctx->nc_total, ctx->scratch8.p,
               ctx->err_flag.p);
  band_from_flags(ctx, d_cells, dn_exact(n), BandPredFlags{ctx->scratch8.p}, include_ghosts, *out);
  own.release();
  check_call(ctx, "cfx_interior_facets_for_cells (Cell index is out of range.)");
  CFX_API_END(ctx)
}

static void facet_rows_impl(cfx_ctx* ctx, const int32_t* d_f, DN n, cfx_list* o)
{
  if (!o->d_n)
  {
    o->d_n = alloc_count_slot(ctx);
    o->ctx = ctx;
  }
  o->data.reserve(ctx->pool, static_cast<size_t>(4 * n.h) + 1);
  o->n = 4 * n.h;
  o->deferred = n.d != nullptr;
  note_result(ctx, o);
  if (n.h > 0)
    CFX_LAUNCH(ctx, facet_rows_kernel, grid_for(n.h, FB), FB, 0, d_f, n, ctx->tdim + 1, ctx->n_facets, ctx->c2f,
               ctx->f2c2.p, o->data.p, ctx->err_flag.p, o->d_n);
  else
    CFX_CUDA(cudaMemsetAsync(o->d_n, 0, sizeof(int64_t), ctx->stream));
}

cfx_status cfx_facet_integration_rows(cfx_ctx* ctx, const int32_t* facets, int64_t n, int memspace, cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->topo_bound, CFX_ERR_STATE, "Facet-cell connectivity is unavailable.");
  CFX_REQUIRE(out != nullptr && (facets != nullptr || n == 0), CFX_ERR_INVALID,
              "cfx_facet_integration_rows: NULL argument");
  if (*out == nullptr)
    *out = new cfx_list();
  DevBuf<int32_t> own;
  const int32_t* d_f = n > 0 ? adopt(ctx, own, facets, static_cast<size_t>(n), memspace) : nullptr;
  facet_rows_impl(ctx, d_f, dn_exact(n), *out);
  own.release();
  if (n > 0)
    check_call(ctx, "cfx_facet_integration_rows (Interior facet domain contains a facet without two "
                    "adjacent cells / could not resolve local facet index)");
  CFX_API_END(ctx)
}

cfx_status cfx_facet_integration_rows_list(cfx_ctx* ctx, const cfx_list* facets, cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->topo_bound, CFX_ERR_STATE, "Facet-cell connectivity is unavailable.");
  CFX_REQUIRE(out != nullptr && facets != nullptr && facets != *out, CFX_ERR_INVALID,
              "cfx_facet_integration_rows_list: NULL argument");
  if (*out == nullptr)
    *out = new cfx_list();
  facet_rows_impl(ctx, facets->data.p, dn_of(facets), *out);
  check_call(ctx, "cfx_facet_integration_rows (Interior facet domain contains a facet without two adjacent cells / "
                  "could not resolve local facet index)");
  CFX_API_END(ctx)
}
} // extern "C"
