// common.cuh -- context, device-memory pool, launch/scan helpers shared by all translation units
// of libcutfemx_b200 (sm_100a only; there is no CPU path in this library).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
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

// A size that may be known on the device only ("deferred-size mode", cfx_set_deferred): kernels take a DN where
// they used to take an int64 bound and read the exact value themselves; the host sizes grids and buffers with the
// upper bound h (h is the exact value and d is null when the host knows it).
struct DN
{
  const int64_t* d;
  int64_t h;
  int shift; // the device value counts 2^shift units per item (facet-row lists: 4 int32 entries per facet)
#ifdef __CUDACC__
  __device__ __forceinline__ int64_t get() const { return d ? (*d >> shift) : h; }
#endif
};

// ---- caching device allocator: handles are created/freed every step of a moving-level-set
// loop; cudaMalloc/cudaFree would serialise the stream each time.
class DevPool
{
public:
  // CUDA-graph capture (cfx_graph_begin / _end): blocks the captured calls free must stay reserved for the graph's
  // replays, and blocks cached before the capture must not become graph temporaries that a later eager call could
  // be handed as well.  begin_capture() sets the cache aside; end_capture() returns everything freed during the
  // capture (now owned by the graph) and restores the cache.
  // Lanes (cfx_lane_begin): every lane of a context issues its work on a stream of its own, so a block one lane
  // frees must not be handed to another lane whose stream is not ordered after the first -- each lane has its own
  // cache of free blocks (a block returns to the cache of the lane that is current when it is freed: its last
  // users are that lane's stream, or work that the fork / join events order before it).
  enum { LANES = 4 };
  void set_lane(int k) { cur_ = k; }
  void begin_capture()
  {
    for (int k = 0; k < LANES; ++k)
      stash_[k].swap(free_[k]);
  }
  std::vector<std::pair<size_t, void*>> end_capture()
  {
    std::vector<std::pair<size_t, void*>> owned;
    for (int k = 0; k < LANES; ++k)
    {
      owned.insert(owned.end(), free_[k].begin(), free_[k].end());
      free_[k].clear();
      stash_[k].swap(free_[k]);
    }
    return owned;
  }
  void give_back(const std::vector<std::pair<size_t, void*>>& blocks)
  {
    for (auto& b : blocks)
      free_[0].emplace(b.first, b.second);
  }
  void* alloc(size_t bytes)
  {
    if (bytes == 0)
      bytes = 256;
    bytes = (bytes + 255) & ~size_t(255);
    auto& fr = free_[cur_];
    auto it = fr.lower_bound(bytes);
    if (it != fr.end() && it->first <= 2 * bytes + (size_t(1) << 20))
    {
      void* p = it->second;
      size_t sz = it->first;
      fr.erase(it);
      live_[p] = {sz, cur_};
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
    live_[p] = {bytes, cur_};
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
    // freed on the main lane: back to the cache of the lane that allocated it (that lane's next use follows its
    // next fork, which orders it after everything the main stream has been given); freed inside a lane: that
    // lane's cache
    free_[cur_ == 0 ? it->second.second : cur_].emplace(it->second.first, p);
    live_.erase(it);
  }
  void release_cached()
  {
    for (int k = 0; k < LANES; ++k)
      for (auto* m : {&free_[k], &stash_[k]})
      {
        for (auto& kv : *m)
        {
          cudaFree(kv.second);
          total_ -= kv.first;
        }
        m->clear();
      }
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
  std::multimap<size_t, void*> free_[LANES], stash_[LANES];
  int cur_ = 0;
  std::unordered_map<void*, std::pair<size_t, int>> live_; // block -> (bytes, lane that allocated it)
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
  int64_t n_cut = -1;       // exact count, or (cut_deferred) an upper bound; -1: list not built since the update
  DevBuf<int32_t> cut_list_all; // intersected owned + ghost cells (sources of the ghost-penalty band)
  int64_t n_cut_all = -1;
  // deferred-size mode: the exact counts live in these device slots
  int64_t* d_n_cut = nullptr;
  int64_t* d_n_cut_all = nullptr;
  bool cut_deferred = false, cut_all_deferred = false;
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
  // deferred-size mode: upper bounds for the per-step active-row / band-row lists of forms over this space
  // (largest counts seen in eager steps, with the context's capacity margin); 0 = not known yet
  int64_t cap_act_rows = 0, cap_band = 0;
  // what eager assemblies on this space have seen: rows the fast gather paths cannot handle (generic kernel) and
  // static rows without a contribution list; -1 = never observed.  A deferred-size step skips those kernels when
  // the numbers were 0 and verifies on the device that they still are.
  int64_t seen_slow_rows = -1, seen_noclist_rows = -1;
  int stride = 0;               // max number of cells around a dof
  // per incidence (nd <= 6): bits 0..3 = local index of the dof in the cell, bits 4+4j.. = rank of
  // the cell's j-th dof among the cell's dofs (ascending global number)
  bool has_perm = false;
  DevBuf<uint32_t> fperm;
  // static full-mesh structure (every cell active), only if all its rows have <= 32 columns:
  bool has_static = false;
  int max_fcols = 0; // longest row of the static full-mesh pattern
  DevBuf<int64_t> frow_ptr;     // full pattern row pointers
  DevBuf<int32_t> fcols;        // full pattern columns (sorted)
  DevBuf<uint32_t> fmask;       // per incidence: bit mask of the full-row positions of the cell's dofs
  // per full-pattern entry (row r, column k): the (incident cell l, local dof j) pairs that contribute
  // to it, 8 bits each (l | j << 5), ascending l, 0xFF = none; the diagonal entry is 0x..FE (every
  // incident cell contributes).  frow_ok[r] = 0 if some entry of the row has more than 8 cells.
  DevBuf<uint64_t> fclist;
  DevBuf<uint8_t> frow_ok;
  // scalar P1: Laplace tensor rows, one 32-byte record per INCIDENCE (row r, incident cell l), stored in incidence
  // order so that a matrix row reads its records as one contiguous run: the off-diagonal entries
  // |detJ|/tdim! grad(lam_i).grad(lam_j), j != i ascending, and |detJ| in the last slot (the diagonal is minus the
  // sum of the others: constants are in the kernel of the Laplace form).  Static while the mesh and the space are
  // bound; built on first use from the geometry cache.
  DevBuf<double> lrow;
  bool lrow_built = false;
  // scalar P1 with a static structure: per incidence one packed word of positions in the row's full-mesh row
  // (assemble.cu fpos_kernel), what the one-thread-per-row gather reads instead of masks and contribution lists
  DevBuf<uint32_t> fpos;
  DevBuf<uint64_t> fpos64; // scalar P2 on triangles (nd = 6): li, own column, five other dofs
  bool fpos_built = false;
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

struct cfx_ctx;
struct cfx_list
{
  cfx::DevBuf<int32_t> data;
  int64_t n = 0; // number of int32 entries: exact, or -- deferred -- an upper bound (the buffer's capacity)
  int64_t* d_n = nullptr; // device copy of the exact count (slot of the context's count slab)
  bool deferred = false;
  int64_t n_bound = 0;    // the upper bound n had when the size was last left on the device
  cfx_ctx* ctx = nullptr; // owner (set when d_n is allocated): size queries of a deferred list synchronise on it
};

struct cfx_rules
{
  int tdim = 0, gdim = 0, relation = 0, order = 0, ls = 0;
  int64_t npts = 0, nrules = 0; // exact, or -- deferred -- upper bounds (d_sizes holds the exact pair)
  int64_t* d_sizes = nullptr;   // device: [0] nrules, [1] npts
  int64_t cap_pts = 0, cap_rules = 0; // capacities of the point / rule arrays (>= the sizes, with margin)
  bool deferred = false;
  cfx_ctx* ctx = nullptr;
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
  int64_t n_rows = 0, nnz = 0; // nnz: exact, or -- deferred -- an upper bound (row_ptr[n_rows] holds the exact value)
  bool deferred = false;
  cfx_ctx* ctx = nullptr;
  int64_t serial = 0;
  cfx::DevBuf<int64_t> row_ptr;
  cfx::DevBuf<int32_t> cols;
  cfx::DevBuf<double> values;
  // the values are known to be all zero (fresh from cfx_create_sparsity / cfx_pattern_import, nothing has written
  // them since): an assembly that ADDS into them may overwrite instead and skip reading 8 B per entry
  bool values_zero = false;
  // values_zero is claimed but only the INACTIVE rows' entries have really been zeroed (cfx_create_sparsity): the
  // first assembly of the pattern's own form overwrites every entry of every active row, so a zero fill of the
  // whole array (8 B per entry: 0.5 GB at 256^3, 15.6 GB for the blocked 192^3 matrix) would be written twice.
  // Anything else that reads or partly writes the values first settles them (settle_values: the zero fill, late).
  bool values_lazy = false;
};

struct cfx_integral
{
  int kernel = 0;
  bool facet = false;
  const int32_t* entities = nullptr; // device: std cells (cells) or rows4 (facets)
  int64_t n = 0;                     // exact, or an upper bound when d_n is set (list with a deferred size)
  const int64_t* d_n = nullptr;      // device count of `entities` in int32 entries (4 per facet row)
  cfx::DevBuf<int32_t> own;
  cfx_rules* rules = nullptr; // borrowed
  double constants[CFX_MAX_CONSTANTS] = {};
};

// active cells / rows of a set of integration domains (Form.h:46-89 domains)
enum { CFX_MAX_STD_LISTS = 6 };
// Bilinear forms on blocked (vector) spaces whose standard-quadrature cell integrals are all linear elasticity:
// those cells are not materialised, the row gather evaluates their blocks (assemble.cu gather_matrix_blocked2_kernel).
inline bool blocked_on_the_fly(const cfx::Space& S, const cfx_form* f);

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
  // deferred-size mode: n_act_rows / n_band are upper bounds, the exact values live in d_counts[0] / [1]
  bool act_deferred = false, band_deferred = false;
  int64_t* d_counts = nullptr;
  cfx::DN dn_act() const { return cfx::DN{act_deferred ? d_counts : nullptr, n_act_rows, 0}; }
  cfx::DN dn_band() const { return cfx::DN{band_deferred ? d_counts + 1 : nullptr, n_band, 0}; }
  int64_t n_active_entities = 0;   // sum of the list sizes (for the byte accounting only)
  bool lists_built = false;        // act_rows / band_idx (a caller that needs the flags only leaves them for later)
};

struct cfx_form
{
  int space = 0, rank = 0;
  std::vector<cfx_integral> integrals;
  // SparsityPattern::insert entries from other ranks (cfx_form_insert_pattern_entries), sorted by row
  cfx::DevBuf<int32_t> xrows, xcols;
  int64_t n_x = 0;
  const int64_t* d_n_x = nullptr; // deferred: exact number of inserted entries (n_x is then the capacity)
  bool deferred = false;          // some size this form depends on is known on the device only
  // prepared state (recomputed when dirty); shared between forms with the same cell domains
  bool dirty = true;
  // the only change since the form was prepared is a list of inserted pattern entries: prepare_form marks their
  // rows in the existing prepared domain instead of preparing it again (the multi-rank step prepares the form for
  // the ghost-row bits first and receives the entries afterwards)
  bool dirty_x_only = false;
  cfx_prepared* prep = nullptr;
  // gather table of the pattern built from this form (sparsity.cu)
  cfx::DevBuf<uint32_t> gmask;  // (n_act_rows, stride): CSR positions of the dofs of incident cell l (band rows)
  cfx::DevBuf<uint32_t> Rrow;   // (n_act_rows): static rows: which full-mesh columns the row keeps
  // (n_act_rows): static rows all of whose incident cells carry the SAME flag byte, a standard-quadrature one
  // (no materialised tensor, no band facet): that byte, else 0.  Such a row needs no per-cell state in the gather.
  cfx::DevBuf<uint8_t> row_ufl;
  cfx::DevBuf<uint8_t> row_fast; // bit0: mask path, bit1: has band cells, bit2: static row
  int64_t n_slow_rows = 0;
  int64_t n_mask_rows = -1; // rows of the mask gather kernel (band rows, long contribution lists); -1 unknown
  int64_t n_clist_rows = 0, n_clist_nnz = 0; // rows / CSR entries of the contribution-list gather kernel
  int64_t n_band_listed = 0; // generic-pattern rows reached through the prepared band slot list (row_fast bit 16)
  bool expect_noclist_zero = false; // deferred step: no static row without a contribution list (verified on the device)
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
  // deferred-size mode (cfx_set_deferred): calls that reuse objects whose buffers already have a capacity leave
  // their result sizes on the device and return without synchronising; the host fetches them on demand
  bool deferred = false;
  double margin = 0.125;            // extra capacity given to size-dependent buffers (so later steps fit)
  bool capturing = false;           // between cfx_graph_begin and cfx_graph_end: synchronising is an error
  cudaStream_t user_stream = nullptr;
  void* comm = nullptr;             // ncclComm_t of this rank (cfx_comm_init), one rank per GPU
  int comm_rank = 0, comm_size = 1;
  // objects whose sizes were left on the device by captured calls: a replay of the graph makes them deferred again
  std::vector<cfx_list*> cap_lists;
  std::vector<cfx_rules*> cap_rules;
  std::vector<cfx_pattern*> cap_patterns;
  cfx::DevBuf<int64_t> count_slab;  // device int64 slots handed to objects with device-side sizes
  std::vector<int> free_slots;
  int next_slot = 0;
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

  // Lanes: independent call sequences of one step (volume rules | interface rules + normals | ghost-penalty facets |
  // cell lists) issued on streams of their own between cfx_lane_begin / cfx_lane_end and joined into the main
  // stream by cfx_lane_join -- concurrent kernels on the device, parallel branches in a captured graph.  A lane has
  // its own stream, its own cache of free blocks (DevPool) and its own scan / compaction scratch.
  struct Lane
  {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool open = false; // has work that the main stream has not joined yet
    cfx::DevBuf<int32_t> blk_counts;
    cfx::DevBuf<int64_t> blk_offsets;
    cfx::DevBuf<int64_t> scratch64;
    cfx::DevBuf<uint8_t> scratch8;
  };
  Lane lanes[cfx::DevPool::LANES];
  bool lanes_on = true;               // cfx_set_lanes: off = every lane is the main stream (serial, for per-kernel timing)
  int lane = 0;                       // the current lane (0 = the main stream)
  cudaStream_t main_stream = nullptr; // the main stream while a lane is current
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
    // inside a graph capture an ordinary record would only mark a dependency; the external flag makes it a real
    // event-record node, so every replay of the graph re-times the stage
    cudaEventRecordWithFlags(s.e0, c->stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
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
      cudaEventRecordWithFlags(c->stages[idx].e1, c->stream,
                               c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
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
int64_t* alloc_count_slot(cfx_ctx* c, int n = 1); // api.cu: n consecutive device int64 slots (zero-initialised)
void free_count_slot(cfx_ctx* c, int64_t* p, int n = 1);
inline int64_t with_margin(const cfx_ctx* c, int64_t n) { return n + static_cast<int64_t>(c->margin * n) + 256; }
// make a deferred size known to the host (synchronises); api.cu
void resolve(cfx_ctx* c, cfx_list* l);
void resolve(cfx_ctx* c, cfx_rules* r);
void resolve(cfx_ctx* c, cfx_pattern* p);
template <class T>
inline void note_deferred(std::vector<T*>& v, T* o)
{
  for (T* q : v)
    if (q == o)
      return;
  v.push_back(o);
}
// call after a deferred-capable entry point has (re)filled the object
inline void note_result(cfx_ctx* c, cfx_list* l)
{
  if (l->deferred)
  {
    l->n_bound = l->n;
    if (c->capturing)
      note_deferred(c->cap_lists, l);
  }
}
inline void note_result(cfx_ctx* c, cfx_rules* r)
{
  if (r->deferred && c->capturing)
    note_deferred(c->cap_rules, r);
}
inline void note_result(cfx_ctx* c, cfx_pattern* p)
{
  if (p->deferred && c->capturing)
    note_deferred(c->cap_patterns, p);
}
inline DN dn_of(const cfx_list* l) { return DN{l->deferred ? l->d_n : nullptr, l->n, 0}; }
inline DN dn_exact(int64_t n) { return DN{nullptr, n, 0}; }

void check_device_error(cfx_ctx* c, const char* where); // api.cu: reads the device error flag back (synchronises)
// the per-call check of the eager mode; in deferred-size mode the flag is looked at when a size is resolved, a
// result is fetched or cfx_check is called
inline void check_call(cfx_ctx* c, const char* where)
{
  if (!c->deferred)
    check_device_error(c, where);
}
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
// the same over a device-side length (n.h bounds the grids; entries at or past the exact length count as 0 and
// out[exact length] = total)
void exclusive_scan_i64(cfx_ctx* c, const int64_t* in, DN n, int64_t* out);
// block-count scan used by the compaction passes: offsets[b] = sum_{j<b} counts[j]; total -> scratch64[0]
// `total_out` (default: scratch64[0]) receives the total; with cap >= 0 a total above cap raises the device error
// flag and is replaced by 0 (the consumers of a deferred size then do nothing instead of overrunning a buffer)
void scan_block_counts(cfx_ctx* c, const int32_t* counts, int64_t nblocks, int64_t* offsets,
                       int64_t* total_out = nullptr, int64_t cap = -1);

RuleTable& get_rule(cfx_ctx* c, int dim, int order); // quadrature.cu: built-in or override, uploaded on demand
void builtin_simplex_rule(int dim, int order, std::vector<double>& pts, std::vector<double>& wts);
void classify_all(cfx_ctx* c);                          // classify.cu
// lanes (api.cu): see cfx_lane_begin in the header.  Library calls fork INTERNALLY too where they launch kernels over
// disjoint row sets (static | band | inactive rows of a pattern, static | band rows of the matrix gather).
void lane_begin(cfx_ctx* c, int k);
void lane_end(cfx_ctx* c) noexcept;
void lane_join(cfx_ctx* c);
inline bool lanes_enabled(const cfx_ctx* c)
{
  return c->lanes_on && c->lane == 0;
}
// the launches in this scope go to lane k if the context is on its main stream (else they stay where they are)
struct LaneScope
{
  cfx_ctx* c;
  bool on;
  LaneScope(cfx_ctx* ctx, int k) : c(ctx), on(k > 0 && lanes_enabled(ctx))
  {
    if (on)
      lane_begin(c, k);
  }
  ~LaneScope()
  {
    if (on)
      lane_end(c);
  }
};
void ensure_cut_list(cfx_ctx* c, int ls);               // classify.cu
void ensure_cut_list_all(cfx_ctx* c, int ls);           // classify.cu
void build_incidence(cfx_ctx* c, Space& s);             // sparsity.cu
void derive_f2c(cfx_ctx* c);                            // facets.cu
void entity_physical_points(cfx_ctx* c, const cfx_rules* r, double* dst_soa); // entity.cu
void build_geometry_cache(cfx_ctx* c);                  // assemble.cu
void dense_f2c_from_adjacency(cfx_ctx* c, const int32_t* off_dev, const int32_t* data_dev); // facets.cu
void settle_values(cfx_ctx* c, cfx_pattern* P);          // sparsity.cu
void prepare_form(cfx_ctx* c, cfx_form* f, bool lists = true); // sparsity.cu (lists = false: cell / row flags only)
void resolve_form(cfx_ctx* c, cfx_form* f);             // sparsity.cu: deferred sizes of a form -> host
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

inline bool blocked_on_the_fly(const cfx::Space& S, const cfx_form* f)
{
  if (S.bs <= 1 || f->rank != 2)
    return false;
  bool any = false;
  for (auto& I : f->integrals)
  {
    if (I.facet || I.n == 0)
      continue;
    if (I.kernel != CFX_K_ELASTICITY)
      return false;
    any = true;
  }
  return any;
}
