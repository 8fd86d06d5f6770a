// sparsity.cu -- dof->cell incidence, form preparation (active cells / rows) and the CSR
// sparsity pattern built on the device.
//
// Replaces dolfinx_custom_data::fem::create_sparsity_pattern + SparsityPattern::finalize +
// MatrixCSR(sp) (cpp/dolfinx_custom_data/fem/assembler.h:442-592, python/cutfemx/wrappers/fem.cpp:
// 266-276): pattern = union of cell cliques over the cell-integral domains, macro cliques over the
// interior-facet domains, and the diagonal of every owned+ghost row
// (insert_deactivation_diagonal, assembler.h:538-560); columns sorted ascending per row.
//
// Design: "owner gathers".  A static incidence dof -> cells (ascending), built once per
// cfx_space_bind, lets one thread own one matrix row: it walks the row's cells, unions their
// dofs into a sorted unique list (pattern) or accumulates their element-tensor rows
// (assemble.cu) in a fixed order.  No sort of (row, col) pairs, no atomics, bit-reproducible.
#include "compact.cuh"

namespace cfx
{
namespace
{
constexpr int SBK = 256;

__global__ void inc_count_kernel(const int32_t* __restrict__ dofmap, int64_t n_entries, int64_t n_dofs,
                                 int32_t* __restrict__ deg, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n_entries)
    return;
  const int32_t d = dofmap[i];
  if (d < 0 || d >= n_dofs)
  {
    err[0] = 21;
    err[1] = d;
    return;
  }
  atomicAdd(&deg[d], 1);
}

__global__ void inc_fill_kernel(const int32_t* __restrict__ dofmap, int64_t n_entries, int nd,
                                const int64_t* __restrict__ inc_ptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ inc_cell)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n_entries)
    return;
  const int32_t d = dofmap[i];
  const int pos = atomicAdd(&cursor[d], 1);
  inc_cell[inc_ptr[d] + pos] = static_cast<int32_t>(i / nd);
}

// the fill order above depends on scheduling; sorting every (short) segment makes the incidence,
// and with it every floating-point summation order downstream, deterministic
__global__ void inc_sort_kernel(int64_t n_dofs, const int64_t* __restrict__ inc_ptr, int32_t* __restrict__ inc_cell)
{
  const int64_t d = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (d >= n_dofs)
    return;
  const int64_t b = inc_ptr[d], e = inc_ptr[d + 1];
  for (int64_t i = b + 1; i < e; ++i)
  {
    const int32_t v = inc_cell[i];
    int64_t j = i;
    while (j > b && inc_cell[j - 1] > v)
    {
      inc_cell[j] = inc_cell[j - 1];
      --j;
    }
    inc_cell[j] = v;
  }
}

__global__ void or_flag_kernel(const int32_t* __restrict__ idx, int64_t n, int stride, int64_t limit, uint8_t bit,
                               uint8_t* __restrict__ flags, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n)
    return;
  const int32_t c = idx[i * stride];
  if (c < 0 || c >= limit)
  {
    err[0] = 22;
    err[1] = c;
    return;
  }
  flags[c] |= bit; // every writer of one launch stores the same bit: idempotent
}

__global__ void scatter_slot_kernel(const int32_t* __restrict__ act, int64_t n, int32_t* __restrict__ slot)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i < n)
    slot[act[i]] = static_cast<int32_t>(i);
}

// row_flag[dof] = 1 for every dof of a flagged cell
__global__ void row_flag_kernel(const uint8_t* __restrict__ cell_flags, int64_t nc, const int32_t* __restrict__ dofmap,
                                int nd, uint8_t* __restrict__ row_flag)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= nc * nd)
    return;
  if (cell_flags[i / nd])
    row_flag[dofmap[i]] = 1;
}

__global__ void facet_slot_set_kernel(const int32_t* __restrict__ rows4, int64_t n, int nf,
                                      const int32_t* __restrict__ c2f, int32_t* __restrict__ facet_slot, bool clear)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n)
    return;
  const int32_t f = c2f[static_cast<int64_t>(rows4[4 * i]) * nf + rows4[4 * i + 1]];
  facet_slot[f] = clear ? -1 : static_cast<int32_t>(i);
}

template <int CAP>
__device__ __forceinline__ bool insert_sorted(int32_t (&a)[CAP], int& n, int32_t v)
{
  int lo = 0, hi = n;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < n && a[lo] == v)
    return true;
  if (n >= CAP)
    return false;
  for (int i = n; i > lo; --i)
    a[i] = a[i - 1];
  a[lo] = v;
  ++n;
  return true;
}

struct RowCtx
{
  const int64_t* inc_ptr;
  const int32_t* inc_cell;
  const int32_t* dofmap;
  const uint8_t* cell_flags;
  const uint8_t* row_flag;
  const int32_t* c2f;
  const int32_t* f2c2;
  const int32_t* facet_slot;
  int nf;
  int insert_diagonal;
};

template <int ND, int CAP>
__device__ __forceinline__ int collect_row(const RowCtx& rc, int64_t r, int32_t (&cols)[CAP], bool& overflow)
{
  int n = 0;
  if (rc.insert_diagonal)
    cols[n++] = static_cast<int32_t>(r);
  if (!rc.row_flag[r])
    return n;
  for (int64_t k = rc.inc_ptr[r]; k < rc.inc_ptr[r + 1]; ++k)
  {
    const int64_t c = rc.inc_cell[k];
    const uint8_t fl = rc.cell_flags[c];
    if (!fl)
      continue;
    bool own_needed = fl & 1;
    if (fl & 2)
    {
      for (int lf = 0; lf < rc.nf; ++lf)
      {
        const int64_t f = rc.c2f[c * rc.nf + lf];
        if (rc.facet_slot[f] < 0)
          continue;
        own_needed = true;
        const int32_t c0 = rc.f2c2[2 * f], c1 = rc.f2c2[2 * f + 1];
        const int64_t other = (c0 == c) ? c1 : c0;
#pragma unroll
        for (int j = 0; j < ND; ++j)
          overflow |= !insert_sorted<CAP>(cols, n, rc.dofmap[other * ND + j]);
      }
    }
    if (own_needed)
    {
#pragma unroll
      for (int j = 0; j < ND; ++j)
        overflow |= !insert_sorted<CAP>(cols, n, rc.dofmap[c * ND + j]);
    }
  }
  return n;
}

template <int ND, int CAP>
__global__ void __launch_bounds__(128)
    pattern_count_kernel(RowCtx rc, int64_t n_rows, int32_t* __restrict__ row_nnz, int32_t* __restrict__ err)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
  if (r >= n_rows)
    return;
  int32_t cols[CAP];
  bool overflow = false;
  row_nnz[r] = collect_row<ND, CAP>(rc, r, cols, overflow);
  if (overflow)
  {
    err[0] = 23;
    err[1] = static_cast<int32_t>(r);
  }
}

template <int ND, int CAP>
__global__ void __launch_bounds__(128)
    pattern_fill_kernel(RowCtx rc, int64_t n_rows, const int64_t* __restrict__ row_ptr, int32_t* __restrict__ out)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
  if (r >= n_rows)
    return;
  int32_t cols[CAP];
  bool overflow = false;
  const int n = collect_row<ND, CAP>(rc, r, cols, overflow);
  const int64_t b = row_ptr[r];
  for (int i = 0; i < n; ++i)
    out[b + i] = cols[i];
}

__global__ void check_sorted_kernel(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                    int64_t n_rows, int32_t* __restrict__ err)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (r >= n_rows)
    return;
  for (int64_t p = row_ptr[r] + 1; p < row_ptr[r + 1]; ++p)
    if (cols[p] <= cols[p - 1])
    {
      err[0] = 24;
      err[1] = static_cast<int32_t>(r);
    }
}
} // namespace

void build_incidence(cfx_ctx* c, Space& S)
{
  const int64_t n_entries = c->nc_total * S.nd;
  DevBuf<int32_t> deg;
  deg.reserve(c->pool, static_cast<size_t>(S.n_total) + 1);
  CFX_CUDA(cudaMemsetAsync(deg.p, 0, (static_cast<size_t>(S.n_total) + 1) * sizeof(int32_t), c->stream));
  CFX_LAUNCH(c, inc_count_kernel, grid_for(n_entries, SBK), SBK, 0, S.dofmap, n_entries, S.n_total, deg.p,
             c->err_flag.p);
  check_device_error(c, "cfx_space_bind (dof index out of range)");
  S.inc_ptr.reserve(c->pool, static_cast<size_t>(S.n_total) + 2);
  exclusive_scan_i32_to_i64(c, deg.p, S.n_total, S.inc_ptr.p);
  S.n_inc = n_entries;
  S.inc_cell.reserve(c->pool, static_cast<size_t>(n_entries) + 1);
  CFX_CUDA(cudaMemsetAsync(deg.p, 0, (static_cast<size_t>(S.n_total) + 1) * sizeof(int32_t), c->stream));
  CFX_LAUNCH(c, inc_fill_kernel, grid_for(n_entries, SBK), SBK, 0, S.dofmap, n_entries, S.nd, S.inc_ptr.p, deg.p,
             S.inc_cell.p);
  CFX_LAUNCH(c, inc_sort_kernel, grid_for(S.n_total, SBK), SBK, 0, S.n_total, S.inc_ptr.p, S.inc_cell.p);
  deg.release();
}

// cell_flags / cell_slot / active list / row flags of a form (Form.h:46-89 domains)
void prepare_form(cfx_ctx* c, cfx_form* f)
{
  if (!f->dirty)
    return;
  Space& S = c->spaces[f->space];
  StageScope st(c, "prepare_form", 6.0 * static_cast<double>(c->nc_total));
  f->cell_flags.reserve(c->pool, static_cast<size_t>(c->nc_total) + 16);
  CFX_CUDA(cudaMemsetAsync(f->cell_flags.p, 0, static_cast<size_t>(c->nc_total), c->stream));
  for (auto& I : f->integrals)
  {
    if (I.facet)
      continue;
    if (I.n > 0)
      CFX_LAUNCH(c, or_flag_kernel, grid_for(I.n, SBK), SBK, 0, I.entities, I.n, 1, c->nc_total, uint8_t(1),
                 f->cell_flags.p, c->err_flag.p);
    if (I.rules && I.rules->nrules > 0)
      CFX_LAUNCH(c, or_flag_kernel, grid_for(I.rules->nrules, SBK), SBK, 0, I.rules->parent_map.p, I.rules->nrules, 1,
                 c->nc_total, uint8_t(1), f->cell_flags.p, c->err_flag.p);
  }
  // active list + slots (before the facet bit is added, so the predicate is just "byte != 0")
  {
    FlagPred p{f->cell_flags.p};
    f->n_active = compact_indices(c, c->nc_total, p, f->active);
  }
  f->cell_slot.reserve(c->pool, static_cast<size_t>(c->nc_total) + 1);
  if (f->n_active > 0)
    CFX_LAUNCH(c, scatter_slot_kernel, grid_for(f->n_active, SBK), SBK, 0, f->active.p, f->n_active, f->cell_slot.p);
  for (auto& I : f->integrals)
  {
    if (!I.facet || I.n == 0)
      continue;
    CFX_LAUNCH(c, or_flag_kernel, grid_for(I.n, SBK), SBK, 0, I.entities, I.n, 4, c->nc_total, uint8_t(2),
               f->cell_flags.p, c->err_flag.p);
    CFX_LAUNCH(c, or_flag_kernel, grid_for(I.n, SBK), SBK, 0, I.entities + 2, I.n, 4, c->nc_total, uint8_t(2),
               f->cell_flags.p, c->err_flag.p);
  }
  f->row_flag.reserve(c->pool, static_cast<size_t>(S.n_total) + 16);
  CFX_CUDA(cudaMemsetAsync(f->row_flag.p, 0, static_cast<size_t>(S.n_total), c->stream));
  CFX_LAUNCH(c, row_flag_kernel, grid_for(c->nc_total * S.nd, SBK), SBK, 0, f->cell_flags.p, c->nc_total, S.dofmap,
             S.nd, f->row_flag.p);
  check_device_error(c, "form domains (entity index out of range)");
  f->dirty = false;
}

const cfx_integral* facet_integral_domain(const cfx_form* f)
{
  const cfx_integral* first = nullptr;
  for (auto& I : f->integrals)
  {
    if (!I.facet || I.n == 0)
      continue;
    if (!first)
      first = &I;
    else
      CFX_REQUIRE(I.entities == first->entities && I.n == first->n, CFX_ERR_UNSUPPORTED,
                  "forms with several different interior-facet domains are not implemented");
  }
  return first;
}

void set_facet_slots(cfx_ctx* c, const cfx_integral* I, bool clear)
{
  if (!I)
    return;
  CFX_REQUIRE(c->topo_bound, CFX_ERR_STATE, "interior-facet integrals need cfx_topology_bind");
  CFX_LAUNCH(c, facet_slot_set_kernel, grid_for(I->n, SBK), SBK, 0, I->entities, I->n, c->tdim + 1, c->c2f,
             c->facet_slot.p, clear);
}
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_form_create(cfx_ctx* ctx, int space, int rank, cfx_form** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && out, CFX_ERR_INVALID, "cfx_form_create: NULL argument");
  CFX_REQUIRE(space >= 0 && space < CFX_MAX_SPACES && ctx->spaces[space].bound, CFX_ERR_INVALID,
              "cfx_form_create: function space not bound");
  CFX_REQUIRE(rank >= 0 && rank <= 2, CFX_ERR_INVALID, "cfx_form_create: rank must be 0, 1 or 2");
  cfx_form* f = new cfx_form();
  f->space = space;
  f->rank = rank;
  *out = f;
  CFX_API_END(ctx)
}

static int kernel_rank(int k)
{
  switch (k)
  {
  case CFX_K_LAPLACE:
  case CFX_K_MASS:
  case CFX_K_NITSCHE:
  case CFX_K_GHOST_GRAD_JUMP: return 2;
  case CFX_K_SOURCE:
  case CFX_K_NITSCHE_RHS: return 1;
  case CFX_K_ONE: return 0;
  }
  return -1;
}

cfx_status cfx_form_add_cell_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* cells, int64_t n_cells,
                                      int memspace, cfx_rules* rules, const double* constants, int n_constants)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && f, CFX_ERR_INVALID, "cfx_form_add_cell_integral: NULL argument");
  CFX_REQUIRE(kernel_rank(kernel) == f->rank && kernel != CFX_K_GHOST_GRAD_JUMP, CFX_ERR_INVALID,
              "cfx_form_add_cell_integral: kernel family does not match the form rank / integral type");
  CFX_REQUIRE(n_constants >= 0 && n_constants <= CFX_MAX_CONSTANTS, CFX_ERR_INVALID, "too many constants");
  CFX_REQUIRE(n_cells == 0 || cells != nullptr, CFX_ERR_INVALID, "cfx_form_add_cell_integral: NULL cells");
  if (kernel == CFX_K_NITSCHE || kernel == CFX_K_NITSCHE_RHS)
  {
    CFX_REQUIRE(n_cells == 0, CFX_ERR_INVALID, "interface kernels take run-time rules only");
    CFX_REQUIRE(rules && rules->has_normals, CFX_ERR_STATE,
                "interface kernels need rules with normals (cfx_evaluate_normals)");
  }
  f->integrals.emplace_back();
  cfx_integral& I = f->integrals.back();
  I.kernel = kernel;
  I.facet = false;
  I.n = n_cells;
  I.entities = n_cells > 0 ? adopt(ctx, I.own, cells, static_cast<size_t>(n_cells), memspace) : nullptr;
  I.rules = rules;
  for (int k = 0; k < n_constants; ++k)
    I.constants[k] = constants[k];
  f->dirty = true;
  if (memspace == CFX_HOST)
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  CFX_API_END(ctx)
}

cfx_status cfx_form_add_interior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* rows4,
                                                int64_t n_facets, int memspace, const double* constants,
                                                int n_constants)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && f, CFX_ERR_INVALID, "cfx_form_add_interior_facet_integral: NULL argument");
  CFX_REQUIRE(kernel == CFX_K_GHOST_GRAD_JUMP && f->rank == 2, CFX_ERR_INVALID,
              "cfx_form_add_interior_facet_integral: unsupported kernel family");
  CFX_REQUIRE(n_constants >= 0 && n_constants <= CFX_MAX_CONSTANTS, CFX_ERR_INVALID, "too many constants");
  CFX_REQUIRE(n_facets == 0 || rows4 != nullptr, CFX_ERR_INVALID, "NULL facet rows");
  f->integrals.emplace_back();
  cfx_integral& I = f->integrals.back();
  I.kernel = kernel;
  I.facet = true;
  I.n = n_facets;
  I.entities = n_facets > 0 ? adopt(ctx, I.own, rows4, static_cast<size_t>(4 * n_facets), memspace) : nullptr;
  for (int k = 0; k < n_constants; ++k)
    I.constants[k] = constants[k];
  f->dirty = true;
  if (memspace == CFX_HOST)
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  CFX_API_END(ctx)
}

void cfx_form_free(cfx_ctx* ctx, cfx_form* f)
{
  (void)ctx;
  if (!f)
    return;
  for (auto& I : f->integrals)
    I.own.release();
  f->cell_flags.release();
  f->cell_slot.release();
  f->active.release();
  f->row_flag.release();
  f->Ae.release();
  f->written.release();
  f->Fe.release();
  delete f;
}

cfx_status cfx_create_sparsity(cfx_ctx* ctx, const cfx_form* a_const, cfx_pattern** inout)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && a && inout, CFX_ERR_INVALID, "cfx_create_sparsity: NULL argument");
  // assembler.h:444-448 "Cannot create sparsity pattern. Form is not a bilinear."
  CFX_REQUIRE(a->rank == 2, CFX_ERR_INVALID, "Cannot create sparsity pattern. Form is not a bilinear.");
  Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  const cfx_integral* FI = facet_integral_domain(a);
  if (*inout == nullptr)
    *inout = new cfx_pattern();
  cfx_pattern* P = *inout;
  P->space = a->space;
  P->n_rows = S.n_total;
  StageScope st(ctx, "create_sparsity");
  set_facet_slots(ctx, FI, false);
  RowCtx rc{S.inc_ptr.p, S.inc_cell.p, S.dofmap, a->cell_flags.p, a->row_flag.p, ctx->c2f,
            ctx->f2c2.p, ctx->facet_slot.p, ctx->tdim + 1, 1};
  DevBuf<int32_t> row_nnz;
  row_nnz.reserve(ctx->pool, static_cast<size_t>(S.n_total) + 1);
  const unsigned g = grid_for(S.n_total, 128);
  auto kcount = S.nd == 3 ? pattern_count_kernel<3, 48>
                : S.nd == 4 ? pattern_count_kernel<4, 96>
                : S.nd == 6 ? pattern_count_kernel<6, 96>
                            : pattern_count_kernel<10, 320>;
  auto kfill = S.nd == 3 ? pattern_fill_kernel<3, 48>
               : S.nd == 4 ? pattern_fill_kernel<4, 96>
               : S.nd == 6 ? pattern_fill_kernel<6, 96>
                           : pattern_fill_kernel<10, 320>;
  CFX_LAUNCH(ctx, kcount, g, 128, 0, rc, S.n_total, row_nnz.p, ctx->err_flag.p);
  P->row_ptr.reserve(ctx->pool, static_cast<size_t>(S.n_total) + 2);
  exclusive_scan_i32_to_i64(ctx, row_nnz.p, S.n_total, P->row_ptr.p);
  P->nnz = read_back(ctx, ctx->scratch64.p, 1)[0];
  P->cols.reserve(ctx->pool, static_cast<size_t>(P->nnz) + 1);
  P->values.reserve(ctx->pool, static_cast<size_t>(P->nnz) + 1);
  CFX_LAUNCH(ctx, kfill, g, 128, 0, rc, S.n_total, P->row_ptr.p, P->cols.p);
  set_facet_slots(ctx, FI, true);
  CFX_CUDA(cudaMemsetAsync(P->values.p, 0, (static_cast<size_t>(P->nnz) + 1) * sizeof(double), ctx->stream));
  row_nnz.release();
  st.set_bytes(12.0 * static_cast<double>(P->nnz) + 8.0 * static_cast<double>(S.n_total));
  check_device_error(ctx, "cfx_create_sparsity (row capacity exceeded)");
  CFX_API_END(ctx)
}

cfx_status cfx_pattern_import(cfx_ctx* ctx, int space, const int64_t* row_ptr, const int32_t* cols, int64_t n_rows,
                              int memspace, cfx_pattern** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && row_ptr && cols && out, CFX_ERR_INVALID, "cfx_pattern_import: NULL argument");
  CFX_REQUIRE(space >= 0 && space < CFX_MAX_SPACES && ctx->spaces[space].bound, CFX_ERR_INVALID,
              "cfx_pattern_import: function space not bound");
  CFX_REQUIRE(n_rows == ctx->spaces[space].n_total, CFX_ERR_INVALID,
              "cfx_pattern_import: row count must equal owned+ghost dofs");
  int64_t nnz = 0;
  if (memspace == CFX_HOST)
    nnz = row_ptr[n_rows];
  else
  {
    CFX_CUDA(cudaMemcpyAsync(&nnz, row_ptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (*out == nullptr)
    *out = new cfx_pattern();
  cfx_pattern* P = *out;
  P->space = space;
  P->n_rows = n_rows;
  P->nnz = nnz;
  P->row_ptr.reserve(ctx->pool, static_cast<size_t>(n_rows) + 1);
  P->cols.reserve(ctx->pool, static_cast<size_t>(nnz) + 1);
  P->values.reserve(ctx->pool, static_cast<size_t>(nnz) + 1);
  const cudaMemcpyKind kind = memspace == CFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  CFX_CUDA(cudaMemcpyAsync(P->row_ptr.p, row_ptr, (static_cast<size_t>(n_rows) + 1) * sizeof(int64_t), kind,
                           ctx->stream));
  CFX_CUDA(cudaMemcpyAsync(P->cols.p, cols, static_cast<size_t>(nnz) * sizeof(int32_t), kind, ctx->stream));
  CFX_CUDA(cudaMemsetAsync(P->values.p, 0, (static_cast<size_t>(nnz) + 1) * sizeof(double), ctx->stream));
  CFX_LAUNCH(ctx, check_sorted_kernel, grid_for(n_rows, SBK), SBK, 0, P->row_ptr.p, P->cols.p, n_rows,
             ctx->err_flag.p);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  check_device_error(ctx, "cfx_pattern_import (columns must be sorted and unique per row)");
  CFX_API_END(ctx)
}

cfx_status cfx_pattern_sizes(const cfx_pattern* p, int64_t* n_rows, int64_t* nnz)
{
  if (!p)
    return CFX_ERR_INVALID;
  if (n_rows)
    *n_rows = p->n_rows;
  if (nnz)
    *nnz = p->nnz;
  return CFX_OK;
}

cfx_status cfx_pattern_fetch(cfx_ctx* ctx, const cfx_pattern* p, int64_t* row_ptr, int32_t* cols, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && p, CFX_ERR_INVALID, "cfx_pattern_fetch: NULL argument");
  export_to(ctx, row_ptr, p->row_ptr.p, static_cast<size_t>(p->n_rows) + 1, memspace);
  export_to(ctx, cols, p->cols.p, static_cast<size_t>(p->nnz), memspace);
  CFX_API_END(ctx)
}

const double* cfx_pattern_values_device_ptr(const cfx_pattern* p) { return p ? p->values.p : nullptr; }

cfx_status cfx_pattern_values_fetch(cfx_ctx* ctx, const cfx_pattern* p, double* values, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && p && values, CFX_ERR_INVALID, "cfx_pattern_values_fetch: NULL argument");
  export_to(ctx, values, p->values.p, static_cast<size_t>(p->nnz), memspace);
  CFX_API_END(ctx)
}

void cfx_pattern_free(cfx_ctx* ctx, cfx_pattern* p)
{
  (void)ctx;
  if (!p)
    return;
  p->row_ptr.release();
  p->cols.release();
  p->values.release();
  delete p;
}
} // extern "C"
