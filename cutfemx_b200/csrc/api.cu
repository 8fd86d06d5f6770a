// api.cu -- context lifetime, array binding and result export of the C ABI
// (include/cutfemx_b200.h).  Host-side validation mirrors the exceptions the reference
// throws at the same seams (cut.cpp:97-106, 444-460, 748-751; runtime_quadrature.h:120-137).
#include "common.cuh"

static thread_local std::string g_thread_err;

void cfx_set_error(cfx_ctx* ctx, const char* msg)
{
  g_thread_err = msg ? msg : "";
  if (ctx)
    ctx->err = g_thread_err;
}

namespace cfx
{
namespace
{
// values first, then (after a system-wide fence) the ticket of this read-back in slot 64
__global__ void to_mapped_kernel(const int64_t* __restrict__ src, int n, volatile int64_t* dst, int64_t ticket)
{
  if (threadIdx.x < n)
    dst[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    dst[64] = ticket;
}
} // namespace

// The path has about a dozen of these per step (list sizes the host needs to size the next buffers), so the
// wake-up latency of a blocking stream synchronisation is paid a dozen times with the GPU idle.  The host spins
// on the ticket in mapped memory instead and only falls back to the stream (to pick up an error, or after a
// long wait) -- the kernel is the last work on the stream, so seeing its ticket means everything before it is done.
const int64_t* read_back(cfx_ctx* c, const int64_t* dev, int n)
{
  CFX_REQUIRE(n >= 0 && n <= 64, CFX_ERR_RANGE, "read_back: at most 64 values");
  const int64_t ticket = ++c->read_ticket;
  to_mapped_kernel<<<1, 64, 0, c->stream>>>(dev, n, c->h_pinned_dev, ticket);
  ++c->launches;
  CFX_CUDA(cudaGetLastError());
  volatile int64_t* flag = c->h_pinned + 64;
  for (int64_t spins = 0; *flag != ticket; ++spins)
  {
    if ((spins & 0x3fff) == 0x3fff)
    { // every 16 K polls: has the stream failed or finished without the ticket becoming visible?
      const cudaError_t q = cudaStreamQuery(c->stream);
      if (q == cudaSuccess)
        break;
      if (q != cudaErrorNotReady)
        CFX_CUDA(q);
    }
  }
  if (*flag != ticket)
    CFX_CUDA(cudaStreamSynchronize(c->stream));
  return c->h_pinned;
}

void check_device_error(cfx_ctx* c, const char* where)
{
  // err_flag is int32[4]; read as two int64
  const int64_t* h = read_back(c, reinterpret_cast<const int64_t*>(c->err_flag.p), 2);
  const int32_t* f = reinterpret_cast<const int32_t*>(h);
  if (f[0] != 0)
  {
    const int32_t code = f[0], detail = f[1];
    CFX_CUDA(cudaMemsetAsync(c->err_flag.p, 0, 4 * sizeof(int32_t), c->stream));
    throw Error(CFX_ERR_RANGE, std::string(where) + ": device-side check failed (code " + std::to_string(code)
                                   + ", detail " + std::to_string(detail) + ")");
  }
}
} // namespace cfx

using namespace cfx;

extern "C"
{
int cfx_version(void) { return 100; }

const char* cfx_last_error(const cfx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_err.c_str(); }

cfx_status cfx_ctx_create(int device, void* stream, cfx_ctx** out)
{
  cfx_ctx* ctx = nullptr;
  CFX_API_BEGIN
  CFX_REQUIRE(out != nullptr, CFX_ERR_INVALID, "cfx_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(CFX_ERR_CUDA, "cfx_ctx_create: no CUDA device available (this library has no CPU fallback)");
  CFX_REQUIRE(device >= 0 && device < ndev, CFX_ERR_INVALID, "cfx_ctx_create: invalid device ordinal");
  CFX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CFX_CUDA(cudaGetDeviceProperties(&prop, device));
  CFX_REQUIRE(prop.major >= 10, CFX_ERR_UNSUPPORTED,
              "cfx_ctx_create: libcutfemx_b200 is built for sm_100a only (Blackwell B200 required)");
  ctx = new cfx_ctx();
  ctx->device = device;
  ctx->stream = static_cast<cudaStream_t>(stream);
  CFX_CUDA(cudaHostAlloc(&ctx->h_pinned, 72 * sizeof(int64_t), cudaHostAllocMapped));
  ctx->h_pinned[64] = 0;
  CFX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->h_pinned_dev), ctx->h_pinned, 0));
  ctx->err_flag.reserve(ctx->pool, 4);
  CFX_CUDA(cudaMemsetAsync(ctx->err_flag.p, 0, 4 * sizeof(int32_t), ctx->stream));
  ctx->scratch64.reserve(ctx->pool, 64);
  *out = ctx;
  ctx = nullptr;
  CFX_API_END(ctx)
}

void cfx_ctx_destroy(cfx_ctx* ctx)
{
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& l : ctx->ls)
    if (l.host_pinned && l.host_values)
      cudaHostUnregister(const_cast<double*>(l.host_values + l.pin_begin));
  for (auto& s : ctx->stages)
  {
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  ctx->pool.release_all();
  if (ctx->h_pinned)
    cudaFreeHost(ctx->h_pinned);
  delete ctx;
}

cfx_status cfx_sync(cfx_ctx* ctx)
{
  CFX_API_BEGIN
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  CFX_API_END(ctx)
}

int64_t cfx_launch_count(const cfx_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- mesh / topology
cfx_status cfx_mesh_bind(cfx_ctx* ctx, const double* x, int64_t n_nodes, const int32_t* x_dofmap,
                         int64_t n_cells_owned, int64_t n_cells_total, int cell_type, int gdim, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && x && x_dofmap, CFX_ERR_INVALID, "cfx_mesh_bind: NULL argument");
  CFX_REQUIRE(cell_type == CFX_TRIANGLE || cell_type == CFX_TETRAHEDRON, CFX_ERR_UNSUPPORTED,
              "cfx_mesh_bind: only affine triangle and tetrahedron meshes are supported");
  const int tdim = cell_type - 1;
  CFX_REQUIRE(gdim == tdim, CFX_ERR_UNSUPPORTED, "cfx_mesh_bind: geometric dimension must equal cell dimension");
  CFX_REQUIRE(n_cells_owned >= 0 && n_cells_owned <= n_cells_total && n_nodes > 0, CFX_ERR_INVALID,
              "cfx_mesh_bind: inconsistent sizes");
  CFX_REQUIRE(n_cells_total < (int64_t(1) << 31), CFX_ERR_RANGE, "cfx_mesh_bind: cell count exceeds int32");
  ctx->x = adopt(ctx, ctx->x_own, x, static_cast<size_t>(n_nodes) * 3, memspace);
  ctx->x_dofmap = adopt(ctx, ctx->x_dofmap_own, x_dofmap, static_cast<size_t>(n_cells_total) * cell_type, memspace);
  ctx->n_nodes = n_nodes;
  ctx->nc_owned = n_cells_owned;
  ctx->nc_total = n_cells_total;
  ctx->cell_type = cell_type;
  ctx->nv = cell_type;
  ctx->tdim = tdim;
  ctx->gdim = gdim;
  ctx->mesh_bound = true;
  ctx->classified = false;
  ctx->domain_stride = (n_cells_total + 15) & ~int64_t(15);
  ctx->domain.reserve(ctx->pool, static_cast<size_t>(ctx->domain_stride) * CFX_MAX_LEVEL_SETS);
  CFX_CUDA(cudaMemsetAsync(ctx->domain.p, 0, static_cast<size_t>(ctx->domain_stride) * CFX_MAX_LEVEL_SETS,
                           ctx->stream));
  ctx->mat_slot.reserve(ctx->pool, static_cast<size_t>(n_cells_total) + 1);
  CFX_CUDA(cudaMemsetAsync(ctx->mat_slot.p, 0xff, (static_cast<size_t>(n_cells_total) + 1) * sizeof(int32_t),
                           ctx->stream));
  build_geometry_cache(ctx);
  if (memspace == CFX_HOST)
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  CFX_API_END(ctx)
}

cfx_status cfx_topology_bind(cfx_ctx* ctx, const int32_t* c2f, const int32_t* f2c_offsets, const int32_t* f2c,
                             int64_t n_facets, int64_t n_owned_facets, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound, CFX_ERR_STATE, "cfx_topology_bind: bind the mesh first");
  CFX_REQUIRE(c2f != nullptr && n_facets > 0 && n_owned_facets <= n_facets, CFX_ERR_INVALID,
              "cfx_topology_bind: invalid arguments");
  CFX_REQUIRE(n_facets < (int64_t(1) << 31), CFX_ERR_RANGE, "cfx_topology_bind: facet count exceeds int32");
  const int nf = ctx->tdim + 1;
  ctx->c2f = adopt(ctx, ctx->c2f_own, c2f, static_cast<size_t>(ctx->nc_total) * nf, memspace);
  ctx->n_facets = n_facets;
  ctx->n_owned_facets = n_owned_facets;
  ctx->f2c2.reserve(ctx->pool, static_cast<size_t>(n_facets) * 2);
  ctx->facet_flag.reserve(ctx->pool, static_cast<size_t>(n_facets) + 4); // marked through 32-bit atomics
  CFX_CUDA(cudaMemsetAsync(ctx->facet_flag.p, 0, static_cast<size_t>(n_facets) + 4, ctx->stream));
  ctx->facet_slot.reserve(ctx->pool, static_cast<size_t>(n_facets));
  CFX_CUDA(cudaMemsetAsync(ctx->facet_slot.p, 0xff, static_cast<size_t>(n_facets) * sizeof(int32_t), ctx->stream));
  ctx->topo_bound = true;
  if (f2c_offsets && f2c)
  {
    DevBuf<int32_t> off_own, dat_own;
    const int32_t* off = adopt(ctx, off_own, f2c_offsets, static_cast<size_t>(n_facets) + 1, memspace);
    int64_t n_data = 0;
    if (memspace == CFX_HOST)
      n_data = f2c_offsets[n_facets];
    else
    {
      int32_t last = 0;
      CFX_CUDA(cudaMemcpyAsync(&last, f2c_offsets + n_facets, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      CFX_CUDA(cudaStreamSynchronize(ctx->stream));
      n_data = last;
    }
    const int32_t* dat = adopt(ctx, dat_own, f2c, static_cast<size_t>(n_data), memspace);
    dense_f2c_from_adjacency(ctx, off, dat);
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
    off_own.release();
    dat_own.release();
  }
  else
  {
    derive_f2c(ctx);
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  check_device_error(ctx, "cfx_topology_bind");
  CFX_API_END(ctx)
}

// ---------------------------------------------------------------- level sets
cfx_status cfx_levelset_bind(cfx_ctx* ctx, int ls, const int32_t* dofmap, int nd, int degree, const double* values,
                             int64_t n_dofs, int memspace, int pin_host)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound, CFX_ERR_STATE, "cfx_levelset_bind: bind the mesh first");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS, CFX_ERR_INVALID, "cfx_levelset_bind: level-set index out of range");
  CFX_REQUIRE(values != nullptr && n_dofs > 0, CFX_ERR_INVALID, "cfx_levelset_bind: NULL values");
  // validate_level_set, cut.cpp:444-460: scalar Lagrange. Degree 1 is the fixed-topology
  // case-table path; higher-order level sets (iterative edge roots) are SURVEY 8(f) rank 4.
  CFX_REQUIRE(degree == 1 || degree == 2, CFX_ERR_UNSUPPORTED, "cfx_levelset_bind: level-set degree must be 1 or 2");
  LevelSet& L = ctx->ls[ls];
  if (L.host_pinned && L.host_values)
  {
    cudaHostUnregister(const_cast<double*>(L.host_values + L.pin_begin));
    L.host_pinned = false;
  }
  if (dofmap == nullptr)
  {
    CFX_REQUIRE(degree == 1 && nd == ctx->nv, CFX_ERR_INVALID,
                "cfx_levelset_bind: dofmap may be NULL only for P1 level sets on the geometry numbering");
    L.dofmap = ctx->x_dofmap;
  }
  else
    L.dofmap = adopt(ctx, L.dofmap_own, dofmap, static_cast<size_t>(ctx->nc_total) * nd, memspace);
  L.nd = nd;
  L.degree = degree;
  L.n_dofs = n_dofs;
  if (memspace == CFX_DEVICE)
  {
    L.values = values;
    L.host_values = nullptr;
  }
  else
  {
    L.values_own.reserve(ctx->pool, static_cast<size_t>(n_dofs));
    L.values = L.values_own.p;
    L.host_values = values;
    L.pin_begin = L.pin_end = 0;
    if (pin_host)
    {
      // Page-lock only the whole pages INSIDE the array: registering partial pages would also pin
      // whatever neighbours share them, and a later copy from such a neighbour that straddles
      // pinned and pageable memory fails with cudaErrorInvalidValue.
      const uintptr_t page = 4096;
      const uintptr_t a = (reinterpret_cast<uintptr_t>(values) + page - 1) & ~(page - 1);
      const uintptr_t b = (reinterpret_cast<uintptr_t>(values + n_dofs)) & ~(page - 1);
      if (b > a && b - a >= (uintptr_t(1) << 16))
      {
        cudaError_t e = cudaHostRegister(reinterpret_cast<void*>(a), b - a, cudaHostRegisterDefault);
        if (e == cudaSuccess)
        {
          L.host_pinned = true;
          L.pin_begin = (a - reinterpret_cast<uintptr_t>(values)) / sizeof(double);
          L.pin_end = (b - reinterpret_cast<uintptr_t>(values)) / sizeof(double);
        }
        else
          cudaGetLastError(); // already pinned by the caller (e.g. torch pinned memory) is fine
      }
    }
  }
  L.bound = true;
  L.n_cut = -1;
  L.n_cut_all = -1;
  ctx->classified = false;
  if (memspace == CFX_HOST)
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  CFX_API_END(ctx)
}

cfx_status cfx_update(cfx_ctx* ctx)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound, CFX_ERR_STATE, "cfx_update: bind the mesh first");
  bool any = false;
  ++ctx->update_serial;
  for (auto& L : ctx->ls)
  {
    if (!L.bound)
      continue;
    any = true;
    if (L.host_values) // cut.cpp:854-855: re-bind dof_values to the (possibly changed) array
    {
      const size_t n = static_cast<size_t>(L.n_dofs);
      const size_t pb = L.host_pinned ? L.pin_begin : n, pe = L.host_pinned ? L.pin_end : n;
      auto copy = [&](size_t b, size_t e)
      {
        if (e > b)
          CFX_CUDA(cudaMemcpyAsync(L.values_own.p + b, L.host_values + b, (e - b) * sizeof(double),
                                   cudaMemcpyHostToDevice, ctx->stream));
      };
      copy(0, pb);  // pageable head
      copy(pb, pe); // page-locked interior (full PCIe rate)
      copy(pe, n);  // pageable tail
    }
    L.n_cut = -1;
    L.n_cut_all = -1;
  }
  CFX_REQUIRE(any, CFX_ERR_STATE, "cfx_update: no level set bound");
  classify_all(ctx);
  ctx->classified = true;
  CFX_API_END(ctx)
}

cfx_status cfx_counts(cfx_ctx* ctx, int ls, int64_t counts[3])
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->classified, CFX_ERR_STATE, "cfx_counts: call cfx_update first");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID, "cfx_counts: bad level set");
  sync_counts(ctx);
  for (int k = 0; k < 3; ++k)
    counts[k] = ctx->ls[ls].counts[k];
  CFX_API_END(ctx)
}

cfx_status cfx_domain_fetch(cfx_ctx* ctx, int ls, int8_t* out, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->classified, CFX_ERR_STATE, "cfx_domain_fetch: call cfx_update first");
  CFX_REQUIRE(ls >= 0 && ls < CFX_MAX_LEVEL_SETS && ctx->ls[ls].bound, CFX_ERR_INVALID, "cfx_domain_fetch: bad ls");
  export_to(ctx, out, ctx->domain.p + static_cast<size_t>(ls) * ctx->domain_stride,
            static_cast<size_t>(ctx->nc_total), memspace);
  CFX_API_END(ctx)
}

// ---------------------------------------------------------------- lists
int64_t cfx_list_size(const cfx_list* l) { return l ? l->n : 0; }
const int32_t* cfx_list_device_ptr(const cfx_list* l) { return l ? l->data.p : nullptr; }

cfx_status cfx_list_fetch(cfx_ctx* ctx, const cfx_list* l, int32_t* out, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && l, CFX_ERR_INVALID, "cfx_list_fetch: NULL argument");
  export_to(ctx, out, l->data.p, static_cast<size_t>(l->n), memspace);
  CFX_API_END(ctx)
}

void cfx_list_free(cfx_ctx* ctx, cfx_list* l)
{
  (void)ctx;
  if (!l)
    return;
  l->data.release();
  delete l;
}

// ---------------------------------------------------------------- spaces
cfx_status cfx_space_bind(cfx_ctx* ctx, int space, const int32_t* dofmap, int nd, int bs, int degree,
                          int64_t n_dofs_owned, int64_t n_dofs_total, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && ctx->mesh_bound, CFX_ERR_STATE, "cfx_space_bind: bind the mesh first");
  CFX_REQUIRE(space >= 0 && space < CFX_MAX_SPACES, CFX_ERR_INVALID, "cfx_space_bind: space index out of range");
  CFX_REQUIRE(dofmap != nullptr, CFX_ERR_INVALID, "cfx_space_bind: NULL dofmap");
  CFX_REQUIRE(degree == 1 || degree == 2, CFX_ERR_UNSUPPORTED, "cfx_space_bind: Lagrange degree must be 1 or 2");
  const int expect = degree == 1 ? ctx->nv : (ctx->tdim == 2 ? 6 : 10);
  CFX_REQUIRE(nd == expect, CFX_ERR_INVALID, "cfx_space_bind: dofmap width does not match the element");
  CFX_REQUIRE(bs == 1 || bs == ctx->gdim, CFX_ERR_UNSUPPORTED,
              "cfx_space_bind: block size must be 1 (scalar) or the geometric dimension (vector space)");
  CFX_REQUIRE(n_dofs_total < (int64_t(1) << 31), CFX_ERR_RANGE, "cfx_space_bind: dof count exceeds int32");
  Space& S = ctx->spaces[space];
  S.dofmap = adopt(ctx, S.dofmap_own, dofmap, static_cast<size_t>(ctx->nc_total) * nd, memspace);
  S.nd = nd;
  S.bs = bs;
  S.degree = degree;
  S.n_owned = n_dofs_owned;
  S.n_total = n_dofs_total;
  S.bound = true;
  build_incidence(ctx, S);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  check_device_error(ctx, "cfx_space_bind");
  CFX_API_END(ctx)
}

// ---------------------------------------------------------------- stage timing
cfx_status cfx_stage_timing_enable(cfx_ctx* ctx, int on)
{
  CFX_API_BEGIN
  ctx->timing = on != 0;
  CFX_API_END(ctx)
}
int cfx_stage_count(const cfx_ctx* ctx) { return ctx ? static_cast<int>(ctx->stages.size()) : 0; }
const char* cfx_stage_name(const cfx_ctx* ctx, int i)
{
  return (ctx && i >= 0 && i < static_cast<int>(ctx->stages.size())) ? ctx->stages[i].name.c_str() : "";
}
cfx_status cfx_stage_ms(cfx_ctx* ctx, int i, double* ms, double* bytes)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && i >= 0 && i < static_cast<int>(ctx->stages.size()), CFX_ERR_INVALID, "cfx_stage_ms: bad index");
  CFX_CUDA(cudaEventSynchronize(ctx->stages[i].e1));
  float t = 0.f;
  CFX_CUDA(cudaEventElapsedTime(&t, ctx->stages[i].e0, ctx->stages[i].e1));
  if (ms)
    *ms = t;
  if (bytes)
    *bytes = ctx->stages[i].bytes;
  CFX_API_END(ctx)
}
cfx_status cfx_stage_reset(cfx_ctx* ctx)
{
  CFX_API_BEGIN
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& s : ctx->stages)
  {
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  ctx->stages.clear();
  CFX_API_END(ctx)
}
} // extern "C"
