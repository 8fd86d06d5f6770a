// classify.cu -- K1: level-set cell classification, and cutfemx::locate_entities.
//
// Replaces the classification half of cutcells::cut (called at cpp/cutfemx/cut/cut.cpp:857) and
// the selector scan cut.cpp:877-924.  Rule (cut.cpp:292-321): all dofs < 0 -> inside, all > 0 ->
// outside, otherwise intersected (a zero dof makes the cell intersected).
//
// Roofline: HBM.  Algorithmic bytes = Nc*(4*nd + 1) + 8*Nd (dofmap row in, one code byte out,
// every level-set value once) -- SURVEY.md section 8(d) "K1".
#include "compact.cuh"

namespace cfx
{
namespace
{
constexpr int CB = 256;

template <int ND>
__device__ __forceinline__ void load_row(const int32_t* __restrict__ dofmap, int64_t c, int32_t (&d)[ND])
{
  if constexpr (ND == 4)
  {
    const int4 v = __ldg(reinterpret_cast<const int4*>(dofmap) + c);
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  }
  else
  {
#pragma unroll
    for (int k = 0; k < ND; ++k)
      d[k] = __ldg(dofmap + c * ND + k);
  }
}

// CPT cells per thread, strided by the warp: in round i a warp reads 32 consecutive dofmap rows (coalesced)
// and writes 32 consecutive code bytes, and the gathers of all CPT rounds are in flight together (the kernel
// is a dependent pair of loads per cell, so bytes in flight per thread set its bandwidth).  Ghost cells
// (c >= nc_owned) are classified too (their codes feed the ghost-penalty band across partition boundaries)
// but are not counted.
constexpr int CPT = 4;

template <int ND>
__global__ void __launch_bounds__(CB)
    classify_kernel(const int32_t* __restrict__ dofmap, const double* __restrict__ vals, int64_t nc_total,
                    int64_t nc_owned, int8_t* __restrict__ domain, unsigned long long* __restrict__ counts,
                    const uint8_t* __restrict__ host /* null: every cell is a host of the cut */,
                    int32_t* __restrict__ blk_counts /* (blocks, 4): owned inside / intersected / outside */)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x) >> 5;
  const int64_t c0 = warp * (32 * CPT) + lane;
  int32_t d[CPT][ND];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
  {
    const int64_t c = c0 + 32 * i;
    if (c < nc_total)
      load_row<ND>(dofmap, c, d[i]);
    else
    {
#pragma unroll
      for (int k = 0; k < ND; ++k)
        d[i][k] = 0;
    }
  }
  double v[CPT][ND];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
#pragma unroll
    for (int k = 0; k < ND; ++k)
      v[i][k] = __ldg(vals + d[i][k]);
  int n_in = 0, n_cut = 0, n_out = 0;
#pragma unroll
  for (int i = 0; i < CPT; ++i)
  {
    const int64_t c = c0 + 32 * i;
    bool all_neg = true, all_pos = true;
#pragma unroll
    for (int k = 0; k < ND; ++k)
    {
      all_neg = all_neg && (v[i][k] < 0.0);
      all_pos = all_pos && (v[i][k] > 0.0);
    }
    int code = all_neg ? CFX_DOMAIN_INSIDE : (all_pos ? CFX_DOMAIN_OUTSIDE : CFX_DOMAIN_INTERSECTED);
    if (host != nullptr && c < nc_total && !host[c])
      code = 0; // not a host entity of this cut (cut.cpp:507-537: the view holds the selected cells only)
    if (c < nc_total)
      domain[c] = static_cast<int8_t>(code);
    const bool owned = c < nc_owned;
    n_in += __popc(__ballot_sync(0xffffffffu, owned && code == CFX_DOMAIN_INSIDE));
    n_cut += __popc(__ballot_sync(0xffffffffu, owned && code == CFX_DOMAIN_INTERSECTED));
    n_out += __popc(__ballot_sync(0xffffffffu, owned && code == CFX_DOMAIN_OUTSIDE));
  }
  __shared__ int s_cnt[3];
  if (threadIdx.x < 3)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (lane == 0)
  {
    atomicAdd(&s_cnt[0], n_in);
    atomicAdd(&s_cnt[1], n_cut);
    atomicAdd(&s_cnt[2], n_out);
  }
  __syncthreads();
  if (threadIdx.x < 3)
    blk_counts[static_cast<int64_t>(blockIdx.x) * 4 + threadIdx.x] = s_cnt[threadIdx.x];
  if (threadIdx.x < 3 && s_cnt[threadIdx.x] != 0)
    atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// per compaction tile (CP_TILE cells = CP_TILE / (CB * CPT) classification blocks): owned cells whose domain code is
// in `relmask` -- the counts compact_count_kernel<DnfPred> would produce for a single-clause selector
__global__ void tile_counts_from_classes_kernel(const int32_t* __restrict__ blk_counts, int64_t n_blocks, unsigned relmask,
                                                int64_t n_tiles, int32_t* __restrict__ tile_counts)
{
  constexpr int PER = CP_TILE / (CB * CPT);
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= n_tiles)
    return;
  int s = 0;
  for (int q = 0; q < PER; ++q)
  {
    const int64_t b = t * PER + q;
    if (b >= n_blocks)
      break;
    const int4 v = *reinterpret_cast<const int4*>(blk_counts + b * 4);
    s += ((relmask >> CFX_DOMAIN_INSIDE) & 1u) ? v.x : 0;
    s += ((relmask >> CFX_DOMAIN_INTERSECTED) & 1u) ? v.y : 0;
    s += ((relmask >> CFX_DOMAIN_OUTSIDE) & 1u) ? v.z : 0;
  }
  tile_counts[t] = s;
}

template <int ND>
void launch_classify(cfx_ctx* c, const LevelSet& L, int8_t* domain, unsigned long long* counts, int32_t* blk_counts)
{
  CFX_LAUNCH(c, classify_kernel<ND>, grid_for(c->nc_total, CB * CPT), CB, 0, L.dofmap, L.values, c->nc_total, c->nc_owned,
             domain, counts, c->has_host_mask ? c->host_mask.p : nullptr, blk_counts);
}
} // namespace

void classify_all(cfx_ctx* c)
{
  c->scratch64.reserve(c->pool, 64);
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(c->scratch64.p) + 8;
  CFX_CUDA(cudaMemsetAsync(counts, 0, 3 * CFX_MAX_LEVEL_SETS * sizeof(unsigned long long), c->stream));
  static_assert(CP_TILE % (CB * CPT) == 0, "a compaction tile is a whole number of classification blocks");
  c->cls_blocks = grid_for(c->nc_total, CB * CPT);
  c->cls_counts.reserve(c->pool, static_cast<size_t>(CFX_MAX_LEVEL_SETS) * c->cls_blocks * 4 + 4);
  for (int l = 0; l < CFX_MAX_LEVEL_SETS; ++l)
  {
    LevelSet& L = c->ls[l];
    if (!L.bound)
      continue;
    StageScope st(c, "classify",
                  static_cast<double>(c->nc_total) * (4.0 * L.nd + 1.0) + 8.0 * static_cast<double>(L.n_dofs));
    int8_t* dom = c->domain.p + static_cast<size_t>(l) * c->domain_stride;
    unsigned long long* cnt = counts + 3 * l;
    int32_t* blk = c->cls_counts.p + static_cast<size_t>(l) * c->cls_blocks * 4;
    switch (L.nd)
    {
    case 3: launch_classify<3>(c, L, dom, cnt, blk); break;
    case 4: launch_classify<4>(c, L, dom, cnt, blk); break;
    case 6: launch_classify<6>(c, L, dom, cnt, blk); break;
    case 10: launch_classify<10>(c, L, dom, cnt, blk); break;
    default: throw Error(CFX_ERR_UNSUPPORTED, "classify: unsupported level-set dofmap width");
    }
  }
  // the counts stay on the device until somebody asks (cfx_counts): a read-back here would idle the GPU once per
  // update for numbers the path itself never needs
  c->counts_pending = true;
}

void sync_counts(cfx_ctx* c)
{
  if (!c->counts_pending)
    return;
  const int64_t* h = read_back(c, c->scratch64.p + 8, 3 * CFX_MAX_LEVEL_SETS);
  for (int l = 0; l < CFX_MAX_LEVEL_SETS; ++l)
    for (int k = 0; k < 3; ++k)
      c->ls[l].counts[k] = h[3 * l + k];
  c->counts_pending = false;
}

Dnf make_dnf(cfx_ctx* c, int n_terms, const int32_t* term_offsets, const int32_t* clause_ls, const int32_t* clause_rel)
{
  CFX_REQUIRE(n_terms >= 1 && term_offsets && clause_ls && clause_rel, CFX_ERR_INVALID, "selector: empty expression");
  CFX_REQUIRE(n_terms <= CFX_MAX_CLAUSES && term_offsets[0] == 0 && term_offsets[n_terms] <= CFX_MAX_CLAUSES,
              CFX_ERR_RANGE, "selector: too many clauses");
  Dnf d;
  std::memset(&d, 0, sizeof(d));
  d.n_terms = n_terms;
  for (int t = 0; t <= n_terms; ++t)
    d.term_off[t] = term_offsets[t];
  for (int k = 0; k < term_offsets[n_terms]; ++k)
  {
    // cut.cpp:895-902: "Compiled selector contains an invalid level-set index"
    CFX_REQUIRE(clause_ls[k] >= 0 && clause_ls[k] < CFX_MAX_LEVEL_SETS && c->ls[clause_ls[k]].bound, CFX_ERR_INVALID,
                "Compiled selector contains an invalid level-set index");
    CFX_REQUIRE(clause_rel[k] >= CFX_REL_LT && clause_rel[k] <= CFX_REL_EQ, CFX_ERR_INVALID,
                "selector: invalid relation");
    d.ls[k] = static_cast<int8_t>(clause_ls[k]);
    d.relmask[k] = relation_mask(clause_rel[k]);
  }
  return d;
}

// owned cells matching a compiled selector.  A single clause `name rel 0` needs no counting pass: the per-tile
// counts follow from the classification's block counts.
int64_t compact_owned_cells(cfx_ctx* c, const Dnf& d, DevBuf<int32_t>& out, int64_t* d_total = nullptr,
                            bool* deferred = nullptr)
{
  DnfPred p{d, c->domain.p, c->domain_stride};
  const int64_t n = c->nc_owned;
  const bool single = d.n_terms == 1 && d.term_off[1] == 1 && n > 0 && c->cls_blocks > 0;
  if (single)
  {
    const int64_t nt = grid_for(n, CP_TILE);
    c->blk_counts.reserve(c->pool, static_cast<size_t>(nt));
    CFX_LAUNCH(c, tile_counts_from_classes_kernel, grid_for(nt, 256), 256, 0,
               c->cls_counts.p + static_cast<size_t>(d.ls[0]) * c->cls_blocks * 4, c->cls_blocks,
               static_cast<unsigned>(d.relmask[0]), nt, c->blk_counts.p);
  }
  return compact_indices(c, dn_exact(n), p, out, single, d_total, deferred);
}

void ensure_cut_list(cfx_ctx* c, int ls)
{
  LevelSet& L = c->ls[ls];
  if (L.n_cut >= 0)
    return;
  Dnf d;
  std::memset(&d, 0, sizeof(d));
  d.n_terms = 1;
  d.term_off[1] = 1;
  d.ls[0] = static_cast<int8_t>(ls);
  d.relmask[0] = relation_mask(CFX_REL_EQ);
  StageScope st(c, "locate_cut", static_cast<double>(c->nc_owned) * 2.0);
  if (!L.d_n_cut)
    L.d_n_cut = alloc_count_slot(c);
  L.n_cut = compact_owned_cells(c, d, L.cut_list, L.d_n_cut, &L.cut_deferred);
  st.set_bytes(static_cast<double>(c->nc_owned) * 2.0 + 4.0 * static_cast<double>(L.n_cut));
}
// intersected cells among ALL local cells (owned + ghost).  The reference's Python loop
// (cut.py:364-379) starts from locate_entities(...), i.e. owned cells only, and therefore drops band
// facets whose cut cell is a ghost of the facet's owner; classifying ghosts too makes the assembled
// matrix independent of the partition (equal to the serial one).
void ensure_cut_list_all(cfx_ctx* c, int ls)
{
  LevelSet& L = c->ls[ls];
  if (c->nc_total == c->nc_owned)
  {
    ensure_cut_list(c, ls);
    return;
  }
  if (L.n_cut_all >= 0)
    return;
  Dnf d;
  std::memset(&d, 0, sizeof(d));
  d.n_terms = 1;
  d.term_off[1] = 1;
  d.ls[0] = static_cast<int8_t>(ls);
  d.relmask[0] = relation_mask(CFX_REL_EQ);
  DnfPred p{d, c->domain.p, c->domain_stride};
  if (!L.d_n_cut_all)
    L.d_n_cut_all = alloc_count_slot(c);
  L.n_cut_all = compact_indices(c, dn_exact(c->nc_total), p, L.cut_list_all, false, L.d_n_cut_all, &L.cut_all_deferred);
}
} // namespace cfx

using namespace cfx;

namespace cfx
{
namespace
{
__global__ void host_mask_kernel(const int32_t* __restrict__ cells, int64_t n, int64_t nc, uint8_t* __restrict__ mask,
                                 int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * CB + threadIdx.x;
  if (i >= n)
    return;
  const int32_t c = cells[i];
  if (c < 0 || c >= nc)
  { // cut.cpp validate_local_entities: entity index outside the local range
    err[0] = 14;
    err[1] = c;
    return;
  }
  mask[c] = 1;
}
} // namespace
} // namespace cfx

// cutfemx.cut(level_set, entities, entity_dim = tdim): only the listed (owned) cells host the cut
// (cut.cpp:500-538 build_mesh_view over the selected cells; test_cut_api.py:160-168).  cells == NULL: all cells.
// Takes effect at the next cfx_update.
extern "C" cfx_status cfx_set_host_cells(cfx_ctx* ctx, const int32_t* cells, int64_t n, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound, CFX_ERR_STATE, "cfx_set_host_cells: bind the mesh first");
  CFX_REQUIRE(n >= 0 && (cells != nullptr || n == 0), CFX_ERR_INVALID, "cfx_set_host_cells: NULL argument");
  ctx->classified = false;
  if (cells == nullptr)
  {
    ctx->has_host_mask = false;
    return CFX_OK;
  }
  ctx->host_mask.reserve(ctx->pool, static_cast<size_t>(ctx->nc_total) + 16);
  CFX_CUDA(cudaMemsetAsync(ctx->host_mask.p, 0, static_cast<size_t>(ctx->nc_total), ctx->stream));
  if (n > 0)
  {
    DevBuf<int32_t> own;
    const int32_t* d = adopt(ctx, own, cells, static_cast<size_t>(n), memspace);
    CFX_LAUNCH(ctx, host_mask_kernel, grid_for(n, CB), CB, 0, d, n, ctx->nc_owned, ctx->host_mask.p, ctx->err_flag.p);
    check_call(ctx, "cfx_set_host_cells (entity index out of range)");
    own.release();
  }
  ctx->has_host_mask = true;
  CFX_API_END(ctx)
}

extern "C" cfx_status cfx_locate_entities(cfx_ctx* ctx, int n_terms, const int32_t* term_offsets,
                                          const int32_t* clause_ls, const int32_t* clause_rel, cfx_list** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->classified, CFX_ERR_STATE, "cfx_locate_entities: call cfx_update first");
  CFX_REQUIRE(out != nullptr, CFX_ERR_INVALID, "cfx_locate_entities: out is NULL");
  const Dnf d = make_dnf(ctx, n_terms, term_offsets, clause_ls, clause_rel);
  if (*out == nullptr)
    *out = new cfx_list();
  cfx_list* l = *out;
  if (!l->d_n)
  {
    l->d_n = alloc_count_slot(ctx);
    l->ctx = ctx;
  }
  StageScope st(ctx, "locate", static_cast<double>(ctx->nc_owned) * 2.0);
  l->n = compact_owned_cells(ctx, d, l->data, l->d_n, &l->deferred);
  note_result(ctx, l);
  st.set_bytes(static_cast<double>(ctx->nc_owned) * 2.0 + 4.0 * static_cast<double>((*out)->n));
  CFX_API_END(ctx)
}
