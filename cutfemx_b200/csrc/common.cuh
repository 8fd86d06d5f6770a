// common.cuh -- context, device-memory pool, launch/scan helpers shared by all translation units
// of libcutfemx_b200 (sm_100a only; there is no CPU path in this library).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cutfemx_b200.h"

namespace cfx
{
struct Error : std::runtime_error
{
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CFX_CUDA(call)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess)                                                                                             \
      throw cfx::Error(CFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                              \
  } while (0)

#define CFX_REQUIRE(cond, code, msg)                                                                                   \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!(cond))                                                                                                       \
      throw cfx::Error(code, msg);                                                                                     \
  } while (0)

// ---- caching device allocator: handles are created/freed every step of a moving-level-set
// loop; cudaMalloc/cudaFree would serialise the stream each time.
class DevPool
{
public:
  void* alloc(size_t bytes)
  {
    if (bytes == 0)
      bytes = 256;
    bytes = (bytes + 255) & ~size_t(255);
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= 2 * bytes + (size_t(1) << 20))
    {
      void* p = it->second;
      size_t sz = it->first;
      free_.erase(it);
      live_[p] = sz;
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess)
    {
      release_cached();
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess)
      throw Error(CFX_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    live_[p] = bytes;
    total_ += bytes;
    return p;
  }
  void free(void* p)
  {
    if (!p)
      return;
    auto it = live_.find(p);
    if (it == live_.end())
      return;
    free_.emplace(it->second, p);
    live_.erase(it);
  }
  void release_cached()
  {
    for (auto& kv : free_)
    {
      cudaFree(kv.second);
      total_ -= kv.first;
    }
    free_.clear();
  }
  void release_all()
  {
    release_cached();
    for (auto& kv : live_)
      cudaFree(kv.first);
    live_.clear();
    total_ = 0;
  }
  size_t total_bytes() const { return total_; }

private:
  std::multimap<size_t, void*> free_;
  std::unordered_map<void*, size_t> live_;
  size_t total_ = 0;
};

// grow-only typed device buffer bound to a pool
template <class T>
struct DevBuf
{
  T* p = nullptr;
  size_t cap = 0; // elements
  DevPool* pool = nullptr;
  void reserve(DevPool& pl, size_t n)
  {
    if (n <= cap && p)
      return;
    if (p)
      pool->free(p);
    pool = &pl;
    p = static_cast<T*>(pl.alloc(n * sizeof(T)));
    cap = n;
  }
  void release()
  {
    if (p && pool)
      pool->free(p);
    p = nullptr;
    cap = 0;
  }
};

struct LevelSet
{
  bool bound = false;
  const int32_t* dofmap = nullptr; // device
  int nd = 0, degree = 0;
  int64_t n_dofs = 0;
  const double* values = nullptr;      // device (owned copy or borrowed)
  const double* host_values = nullptr; // re-read by cfx_update when bound from the host
  bool host_pinned = false;
  size_t pin_begin = 0, pin_end = 0; // element range [begin, end) of host_values that is page-locked
  DevBuf<double> values_own;
  DevBuf<int32_t> dofmap_own;
  int64_t counts[3] = {0, 0, 0};
  DevBuf<int32_t> cut_list; // intersected owned cells, ascending (cached per update)
  int64_t n_cut = -1;
  DevBuf<int32_t> cut_list_all; // intersected owned + ghost cells (sources of the ghost-penalty band)
  int64_t n_cut_all = -1;
};

struct Space
{
  bool bound = false;
  const int32_t* dofmap = nullptr; // device (n_cells_total, nd)
  DevBuf<int32_t> dofmap_own;
  int nd = 0, bs = 1, degree = 0;
  int64_t n_owned = 0, n_total = 0;
  // static incidence dof -> cells (ascending), built once per bind
  DevBuf<int64_t> inc_ptr;
  DevBuf<int32_t> inc_cell;
  int64_t n_inc = 0;
  int stride = 0;               // max number of cells around a dof
  // per incidence (nd <= 6): bits 0..3 = local index of the dof in the cell, bits 4+4j.. = rank of
  // the cell's j-th dof among the cell's dofs (ascending global number)
  bool has_perm = false;
  DevBuf<uint32_t> fperm;
  // static full-mesh structure (every cell active), only if all its rows have <= 32 columns:
  bool has_static = false;
  DevBuf<int64_t> frow_ptr;     // full pattern row pointers
  DevBuf<int32_t> fcols;        // full pattern columns (sorted)
  DevBuf<uint32_t> fmask;       // per incidence: bit mask of the full-row positions of the cell's dofs
  // per full-pattern entry (row r, column k): the (incident cell l, local dof j) pairs that contribute
  // to it, 8 bits each (l | j << 5), ascending l, 0xFF = none; the diagonal entry is 0x..FE (every
  // incident cell contributes).  frow_ok[r] = 0 if some entry of the row has more than 8 cells.
  DevBuf<uint64_t> fclist;
  DevBuf<uint8_t> frow_ok;
};

struct RuleTable
{
  int dim = 0, order = 0, npts = 0;
  bool builtin = true; // false: supplied through cfx_set_simplex_rule (exactness unknown)
  std::vector<double> pts, wts; // host
  double* d_pts = nullptr;      // device, (npts, dim) AoS
  double* d_wts = nullptr;
};

struct Stage
{
  std::string name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double bytes = 0.0;
};
} // namespace cfx

struct cfx_list
{
  cfx::DevBuf<int32_t> data;
  int64_t n = 0; // number of int32 entries
};

struct cfx_rules
{
  int tdim = 0, gdim = 0, relation = 0, order = 0, ls = 0;
  int64_t npts = 0, nrules = 0;
  cfx::DevBuf<double> points;  // SoA (tdim, npts)
  cfx::DevBuf<double> weights; // (npts)
  cfx::DevBuf<int32_t> offsets;    // (nrules + 1)
  cfx::DevBuf<int32_t> parent_map; // (nrules)
  cfx::DevBuf<double> normals;     // SoA (gdim, npts), valid if has_normals
  bool has_normals = false;
  // volume rules from the built-in tables (exact for degree >= 1): per rule W = sum_q w_q and the first moments
  // sum_q w_q xi_q, computed in the generator from the sub-simplex measures and centroids -- the P1 kernels whose
  // integrand is linear in xi (Laplace, source, measure) need nothing else.  Layout (nrules, tdim + 1).
  // Interface rules (exact for degree >= 2) add the second moments sum_q w_q xi_a xi_b (upper triangle): layout
  // (nrules, 1 + tdim + tdim (tdim+1)/2) -- enough for the P1 Nitsche kernels, whose normal is constant per rule.
  cfx::DevBuf<double> moments;
  bool has_moments = false;
  // facet-hosted rules (entity.cu): tdim = mesh tdim - 1, parent_map = facet ids, and the facet's vertices per rule
  // (nrules, tdim + 1) so that physical points need no topology lookup
  bool entity_hosted = false;
  cfx::DevBuf<int32_t> rule_verts;
};

struct cfx_pattern
{
  int space = 0;
  int bs = 1; // values: nnz * bs * bs
  int64_t n_rows = 0, nnz = 0;
  int64_t serial = 0;
  cfx::DevBuf<int64_t> row_ptr;
  cfx::DevBuf<int32_t> cols;
  cfx::DevBuf<double> values;
};

struct cfx_integral
{
  int kernel = 0;
  bool facet = false;
  const int32_t* entities = nullptr; // device: std cells (cells) or rows4 (facets)
  int64_t n = 0;
  cfx::DevBuf<int32_t> own;
  cfx_rules* rules = nullptr; // borrowed
  double constants[CFX_MAX_CONSTANTS] = {};
};

// active cells / rows of a set of integration domains (Form.h:46-89 domains)
enum { CFX_MAX_STD_LISTS = 6 };
struct cfx_prepared
{
  int refs = 0;
  int space = 0;
  int64_t update_serial = 0;
  std::vector<std::pair<const void*, int64_t>> std_key;  // distinct standard-quadrature cell lists (bit 2+i)
  std::vector<std::pair<const void*, int64_t>> rule_key; // distinct (rules parent_map, nrules)
  std::pair<const void*, int64_t> facet_key{nullptr, 0};
  std::pair<const void*, int64_t> extra_key{nullptr, 0}; // inserted pattern entries (rows made active + generic)
  // bit0: has a materialised (run-time rule) tensor, bit1: touches a facet-integral facet,
  // bit 2+i: member of standard cell list i (its tensor rows are computed on the fly by the row owner)
  cfx::DevBuf<uint8_t> cell_flags;
  cfx::DevBuf<uint8_t> row_flag;   // per dof: bit0 touched by an active cell, bit1 by a facet-band cell
  cfx::DevBuf<int32_t> act_rows;   // ascending dofs with row_flag set
  int64_t n_act_rows = 0;
  cfx::DevBuf<int32_t> band_idx;   // ascending slots (indices into act_rows) of the rows with row_flag bit1:
  int64_t n_band = 0;              // facet-band rows and rows with inserted entries (generic pattern path)
  int64_t n_active_entities = 0;   // sum of the list sizes (for the byte accounting only)
};

struct cfx_form
{
  int space = 0, rank = 0;
  std::vector<cfx_integral> integrals;
  // SparsityPattern::insert entries from other ranks (cfx_form_insert_pattern_entries), sorted by row
  cfx::DevBuf<int32_t> xrows, xcols;
  int64_t n_x = 0;
  // prepared state (recomputed when dirty); shared between forms with the same cell domains
  bool dirty = true;
  cfx_prepared* prep = nullptr;
  // gather table of the pattern built from this form (sparsity.cu)
  cfx::DevBuf<uint32_t> gmask;  // (n_act_rows, stride): CSR positions of the dofs of incident cell l (band rows)
  cfx::DevBuf<uint32_t> Rrow;   // (n_act_rows): static rows: which full-mesh columns the row keeps
  cfx::DevBuf<uint8_t> row_fast; // bit0: mask path, bit1: has band cells, bit2: static row
  int64_t n_slow_rows = 0;
  int64_t n_mask_rows = -1; // rows of the mask gather kernel (band rows, long contribution lists); -1 unknown
  int64_t n_clist_rows = 0, n_clist_nnz = 0; // rows / CSR entries of the contribution-list gather kernel
  int64_t n_band_listed = 0; // generic-pattern rows reached through the prepared band slot list (row_fast bit 16)
  int64_t gtab_serial = -1;
  cfx::DevBuf<double> Ae;      // materialised run-time-rule tensors, cell-major (slot, nd^rank) natural order;
                               // rank 0: one value per entity
  cfx::DevBuf<double> Fe;      // facet macro tensors
  // ordinary Function coefficient (cfx_form_set_coefficient): dof values over the form's space
  const double* coeff = nullptr;
  cfx::DevBuf<double> coeff_own;
};

struct cfx_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  cfx::DevPool pool;
  int64_t launches = 0;
  int64_t pattern_serial = 0;
  int64_t update_serial = 0;
  std::vector<cfx_prepared*> preps; // live prepared domains (owned by the forms that reference them)

  // mesh views
  bool mesh_bound = false;
  const double* x = nullptr;
  const int32_t* x_dofmap = nullptr;
  cfx::DevBuf<double> x_own;
  cfx::DevBuf<int32_t> x_dofmap_own;
  int64_t n_nodes = 0, nc_owned = 0, nc_total = 0;
  int cell_type = 0, nv = 0, tdim = 0, gdim = 0;

  // topology
  bool topo_bound = false;
  const int32_t* c2f = nullptr;
  cfx::DevBuf<int32_t> c2f_own;
  cfx::DevBuf<int32_t> f2c2; // dense (n_facets, 2), -1 padded, ascending
  cfx::DevBuf<uint8_t> facet_flag;  // zero between calls
  cfx::DevBuf<int32_t> facet_slot;  // -1 between calls
  cfx::DevBuf<int32_t> xslot;       // per dof: first inserted pattern entry of the row; -1 between calls
  size_t xslot_init = 0;            // capacity that has been initialised to -1
  int64_t n_facets = 0, n_owned_facets = 0;

  cfx::LevelSet ls[CFX_MAX_LEVEL_SETS];
  // host cells of the cut (cutfemx.cut(level_set, entities, entity_dim = tdim)): cells outside the subset get
  // domain code 0, which no selector matches; null = every local cell is a host
  cfx::DevBuf<uint8_t> host_mask;
  bool has_host_mask = false;
  // per classification block (1024 cells) and level set: owned cells inside / intersected / outside -- lets a
  // single-clause locate skip its counting pass over the domain codes (classify.cu)
  cfx::DevBuf<int32_t> cls_counts; // (CFX_MAX_LEVEL_SETS, cls_blocks, 4)
  int64_t cls_blocks = 0;
  cfx::DevBuf<int8_t> domain; // (CFX_MAX_LEVEL_SETS, domain_stride)
  int64_t domain_stride = 0;
  bool classified = false;
  bool counts_pending = false; // the classification counts are still on the device (fetched on demand: cfx_counts)

  cfx::Space spaces[CFX_MAX_SPACES];
  cfx::DevBuf<int32_t> mat_slot; // (nc_total): slot of a cell's materialised tensor; -1 between assemblies
  // static per-cell affine geometry (the mesh does not move between cfx_update calls): K = J^-1
  // row-major (tdim^2), detJ, cell diameter; record stride 12 doubles (3D, 96 B) / 8 doubles (2D, 64 B)
  cfx::DevBuf<double> geo;

  std::map<std::pair<int, int>, cfx::RuleTable> rules; // (dim, order) -> table (built-in or override)
  std::map<std::pair<int, int>, cfx::DevBuf<double>> ref_tabs; // (tdim, degree) -> reference-element integrals

  // scratch
  cfx::DevBuf<int32_t> blk_counts;
  cfx::DevBuf<int64_t> blk_offsets;
  cfx::DevBuf<int64_t> scratch64;
  cfx::DevBuf<uint8_t> scratch8;
  cfx::DevBuf<int32_t> err_flag; // device int[4]
  int64_t* h_pinned = nullptr;   // 64 x int64 mapped pinned host scratch + the read-back ticket in slot 64
  int64_t read_ticket = 0;
  int64_t* h_pinned_dev = nullptr; // its device address

  bool timing = false;
  std::vector<cfx::Stage> stages;
};

namespace cfx
{
// ---------------------------------------------------------------- launch helper
#define CFX_LAUNCH(ctx, kernel, grid, block, smem, ...)                                                                \
  do                                                                                                                   \
  {                                                                                                                    \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                                   \
    ++(ctx)->launches;                                                                                                 \
    cudaError_t le_ = cudaGetLastError();                                                                              \
    if (le_ != cudaSuccess)                                                                                            \
      throw cfx::Error(CFX_ERR_CUDA, std::string(#kernel) + " launch: " + cudaGetErrorString(le_));                    \
  } while (0)

inline unsigned grid_for(int64_t n, int per_block)
{
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1)
    g = 1;
  if (g > 2147483647LL)
    throw Error(CFX_ERR_RANGE, "grid too large");
  return static_cast<unsigned>(g);
}

struct StageScope
{
  cfx_ctx* c;
  int idx = -1;
  StageScope(cfx_ctx* ctx, const char* name, double bytes = 0.0) : c(ctx)
  {
    if (!c->timing)
      return;
    Stage s;
    s.name = name;
    s.bytes = bytes;
    cudaEventCreate(&s.e0);
    cudaEventCreate(&s.e1);
    cudaEventRecord(s.e0, c->stream);
    c->stages.push_back(s);
    idx = static_cast<int>(c->stages.size()) - 1;
  }
  void set_bytes(double b)
  {
    if (idx >= 0)
      c->stages[idx].bytes = b;
  }
  ~StageScope()
  {
    if (idx >= 0)
      cudaEventRecord(c->stages[idx].e1, c->stream);
  }
};

// copy helpers ------------------------------------------------------------------------
template <class T>
inline const T* adopt(cfx_ctx* c, DevBuf<T>& own, const T* src, size_t n, int memspace)
{
  if (memspace == CFX_DEVICE)
    return src;
  own.reserve(c->pool, n);
  CFX_CUDA(cudaMemcpyAsync(own.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return own.p;
}

template <class T>
inline void export_to(cfx_ctx* c, T* dst, const T* src_dev, size_t n, int memspace)
{
  if (!dst || n == 0)
    return;
  CFX_CUDA(cudaMemcpyAsync(dst, src_dev, n * sizeof(T),
                           memspace == CFX_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  if (memspace == CFX_HOST)
    CFX_CUDA(cudaStreamSynchronize(c->stream));
}

// read `n` (<= 64) int64 from device memory into host scratch (synchronises the stream).  A tiny
// kernel writes them straight into MAPPED pinned host memory: no DMA copy is enqueued, so the
// read-back never waits behind a large device->host transfer the caller has in flight on another
// stream (api.cu).
const int64_t* read_back(cfx_ctx* c, const int64_t* dev, int n);

void check_device_error(cfx_ctx* c, const char* where); // api.cu
void sync_counts(cfx_ctx* c);                            // classify.cu: fetch pending classification counts

// ---------------------------------------------------------------- device primitives
#ifdef __CUDACC__
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o)
      v += t;
  }
  return v;
}

// exclusive scan of one int per thread over a block of NT threads; *total = block sum
template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int* total)
{
  __shared__ int s_warp[NT / 32];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int incl = warp_incl_scan(v);
  if (lane == 31)
    s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0)
  {
    int w = lane < NT / 32 ? s_warp[lane] : 0;
    int wi = warp_incl_scan(w);
    if (lane < NT / 32)
      s_warp[lane] = wi - w;
    if (lane == NT / 32 - 1)
      s_total = wi;
  }
  __syncthreads();
  const int r = s_warp[wid] + incl - v;
  *total = s_total;
  __syncthreads();
  return r;
}

__device__ __forceinline__ long long warp_incl_scan_ll(long long v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    long long t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o)
      v += t;
  }
  return v;
}
#endif

// host-callable device-wide utilities (scan.cu)
// out[i] = sum_{j<i} in[j] (int32 in, int64 out), out[n] = total; total also left in c->scratch64[0]
void exclusive_scan_i32_to_i64(cfx_ctx* c, const int32_t* in, int64_t n, int64_t* out);
void exclusive_scan_i64(cfx_ctx* c, const int64_t* in, int64_t n, int64_t* out);
// block-count scan used by the compaction passes: offsets[b] = sum_{j<b} counts[j]; total -> scratch64[0]
void scan_block_counts(cfx_ctx* c, const int32_t* counts, int64_t nblocks, int64_t* offsets);

RuleTable& get_rule(cfx_ctx* c, int dim, int order); // quadrature.cu: built-in or override, uploaded on demand
void builtin_simplex_rule(int dim, int order, std::vector<double>& pts, std::vector<double>& wts);
void classify_all(cfx_ctx* c);                          // classify.cu
void ensure_cut_list(cfx_ctx* c, int ls);               // classify.cu
void ensure_cut_list_all(cfx_ctx* c, int ls);           // classify.cu
void build_incidence(cfx_ctx* c, Space& s);             // sparsity.cu
void derive_f2c(cfx_ctx* c);                            // facets.cu
void entity_physical_points(cfx_ctx* c, const cfx_rules* r, double* dst_soa); // entity.cu
void build_geometry_cache(cfx_ctx* c);                  // assemble.cu
void dense_f2c_from_adjacency(cfx_ctx* c, const int32_t* off_dev, const int32_t* data_dev); // facets.cu
void prepare_form(cfx_ctx* c, cfx_form* f);             // sparsity.cu
void release_prepared(cfx_ctx* c, cfx_form* f);         // sparsity.cu
uint8_t std_list_bit(const cfx_prepared* P, const void* entities, int64_t n); // sparsity.cu
const cfx_integral* facet_integral_domain(const cfx_form* f);          // sparsity.cu
void set_facet_slots(cfx_ctx* c, const cfx_integral* I, bool clear);   // sparsity.cu
} // namespace cfx

#define CFX_API_BEGIN                                                                                                  \
  try                                                                                                                  \
  {
#define CFX_API_END(ctx)                                                                                               \
  }                                                                                                                    \
  catch (const cfx::Error& e)                                                                                          \
  {                                                                                                                    \
    cfx_set_error(ctx, e.what());                                                                                      \
    return e.code;                                                                                                     \
  }                                                                                                                    \
  catch (const std::exception& e)                                                                                      \
  {                                                                                                                    \
    cfx_set_error(ctx, e.what());                                                                                      \
    return CFX_ERR_INVALID;                                                                                            \
  }                                                                                                                    \
  return CFX_OK;

void cfx_set_error(cfx_ctx* ctx, const char* msg);
