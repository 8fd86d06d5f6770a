// exchange.cu -- pack / unpack-add kernels of the ghost exchange (one rank per GPU).
//
// The role of la::MatrixCSR::scatter_rev and la::Vector::scatter_rev(add) (DOLFINx 0.11; called by
// the user after assembly, python/demo/demo_poisson.py:52,54): ghost-row values and ghost vector
// entries are packed into a send buffer, exchanged with the owning ranks (NCCL send/recv on these
// device buffers, cutfemx_b200/parallel.py) and added into the owner's entries.
// Roofline: HBM, 20 B per exchanged value (index + value in, value out); the exchanged volume is one
// mesh plane per neighbour (~8 MB at 256^3 on 8 ranks), so these kernels are
// launch-latency sized.
#include "compact.cuh"

namespace cfx
{
namespace
{
__global__ void gather_f64_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t n,
                                  double* __restrict__ dst)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n)
    dst[i] = src[idx[i]];
}

// indices are distinct within one call: plain read-modify-write, no atomics
__global__ void scatter_add_f64_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t n,
                                       double* __restrict__ dst)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n)
    dst[idx[i]] += src[i];
}
} // namespace
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_gather_f64(cfx_ctx* ctx, const double* src, const int64_t* index, int64_t n, double* dst)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && (n == 0 || (src && index && dst)), CFX_ERR_INVALID, "cfx_gather_f64: NULL argument");
  if (n > 0)
    CFX_LAUNCH(ctx, gather_f64_kernel, grid_for(n, 256), 256, 0, src, index, n, dst);
  CFX_API_END(ctx)
}

cfx_status cfx_scatter_add_f64(cfx_ctx* ctx, const double* src, const int64_t* index, int64_t n, double* dst)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && (n == 0 || (src && index && dst)), CFX_ERR_INVALID, "cfx_scatter_add_f64: NULL argument");
  if (n > 0)
    CFX_LAUNCH(ctx, scatter_add_f64_kernel, grid_for(n, 256), 256, 0, src, index, n, dst);
  CFX_API_END(ctx)
}
} // extern "C"

// ====================================================================================================
// Static exchange plan: SparsityPattern::finalize() + MatrixCSR::scatter_rev() + Vector::scatter_rev(add)
// without a size exchange, a host round trip or Python in the step.
//
// Once per partition the ranks agree on a STATIC superset of the ghost-row entries: for every ghost row the
// sender lists the columns the row can ever have (all cells active, every interior facet in the band); the
// owner translates each candidate (row, column) to its own numbering (unknown dofs become new ghost columns).
// Per step only fixed-size messages travel:
//   pattern:  one BIT per candidate -- "this entry is in my pattern of the step" (ghost_bits_kernel)
//   values:   one double per candidate (0 where the bit is clear) + one per ghost vector entry
// so every ncclSend/ncclRecv has a size known when the plan is made, and the whole exchange can be captured in
// the step's CUDA graph.  The owner expands the received bits into SparsityPattern::insert entries on the device
// (their number stays on the device: the form's inserted-entry list is a deferred-size list with the static
// capacity), and after assembly adds the received values at positions it finds by binary search in its own rows,
// neighbour by neighbour in rank order (fixed order, no atomics: bit-reproducible).
#include <dlfcn.h>

namespace
{
// minimal NCCL surface, resolved at run time from the libnccl the process already has (torch's) or the system one
struct UniqueId // ncclUniqueId: a 128-byte struct, passed BY VALUE to ncclCommInitRank
{
  char bytes[128];
};
struct NcclApi
{
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int NCCL_UINT32 = 3, NCCL_FLOAT64 = 8;

void load_nccl(const char* path)
{
  if (g_nccl.lib)
    return;
  const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names)
  {
    if (!n || !*n)
      continue;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib)
      break;
  }
  if (!g_nccl.lib)
    throw cfx::Error(CFX_ERR_UNSUPPORTED, std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
  auto sym = [&](const char* s)
  {
    void* p = dlsym(g_nccl.lib, s);
    if (!p)
      throw cfx::Error(CFX_ERR_UNSUPPORTED, std::string("NCCL symbol missing: ") + s);
    return p;
  };
  g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, UniqueId, int)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
  g_nccl.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclSend"));
  g_nccl.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclRecv"));
  g_nccl.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
}

void nccl_check(int rc, const char* what)
{
  if (rc != 0)
    throw cfx::Error(CFX_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}
} // namespace

struct cfx_xplan
{
  int space = 0, n_neigh = 0;
  int bs = 1; // block size of the space: bs*bs doubles per candidate entry, bs per ghost vector entry
  std::vector<int> ranks;
  // per neighbour: [k] .. [k+1] ranges (host)
  std::vector<int64_t> s_word_off, r_word_off; // bit messages, in 32-bit words
  std::vector<int64_t> s_val_off, r_val_off;   // value messages, in doubles
  std::vector<int64_t> r_ent_off, r_row_off;   // received candidate entries / vector rows per neighbour
  int64_t n_s_rows = 0, n_s_ent = 0, n_r_ent = 0, n_r_rows = 0;
  // sender tables
  cfx::DevBuf<int32_t> s_rows;     // (n_s_rows) local ghost rows, grouped by owner
  cfx::DevBuf<int64_t> s_ptr;      // (n_s_rows + 1) candidate ranges
  cfx::DevBuf<int32_t> s_cols;     // (n_s_ent) candidate columns, local numbering, ascending per row
  cfx::DevBuf<int32_t> s_erow;     // (n_s_ent) index of the entry's row in s_rows
  cfx::DevBuf<int64_t> s_bit;      // (n_s_rows) bit index of the row's first candidate in the send bit buffer
  cfx::DevBuf<int64_t> s_vpos;     // (n_s_ent) position of the entry's value in the send value buffer
  cfx::DevBuf<int64_t> s_vec_vpos; // (n_s_rows) position of the row's vector entry in the send value buffer
  // receiver tables (arrival order: neighbour by neighbour, the sender's entry order)
  cfx::DevBuf<int32_t> r_row, r_col; // (n_r_ent) local owned row / local column (>= n_total: new ghost column)
  cfx::DevBuf<int64_t> r_bit, r_vpos; // (n_r_ent) bit index in the recv bit buffer / position in the recv value buffer
  cfx::DevBuf<int32_t> r_perm;       // (n_r_ent) arrival indices sorted by (row, arrival): insertion order
  cfx::DevBuf<int32_t> r_vec_row;    // (n_r_rows) local owned row of each received vector entry
  cfx::DevBuf<int64_t> r_vec_vpos;   // (n_r_rows)
  // messages
  cfx::DevBuf<uint32_t> bits_send, bits_recv;
  cfx::DevBuf<double> vals_send, vals_recv;
  // per-step selection of the received candidates whose bit is set (indices into r_perm order)
  cfx::DevBuf<int32_t> sel;
  int64_t* d_n_sel = nullptr;
  // COMPACT value messages (every space but scalar P1, where the static superset of a ghost row is small): only the
  // blocks whose bit is set travel, in candidate order -- per neighbour [bs per ghost row][bs*bs per set bit].  Both
  // ends know the bits, so both find a block's slot as (set bits before it) from a per-word prefix of popcounts.
  // Message sizes: exact in eager steps (one read-back of the counts per side), and in deferred-size steps the
  // capacity the eager steps needed plus the context's margin (both ends saw the same counts, so they agree); a
  // count above the capacity raises the device error flag.
  bool compact = false;
  std::vector<int64_t> s_row_off, s_ent_off;       // host copies: ghost rows / candidates per neighbour
  cfx::DevBuf<int32_t> s_wcnt, r_wcnt;             // popcount of every bit word
  cfx::DevBuf<int64_t> s_wpre, r_wpre;             // exclusive prefix over the words (n_words + 1)
  cfx::DevBuf<int64_t> d_s_word_off, d_r_word_off; // device copies of the word offsets
  int64_t *d_s_cnt = nullptr, *d_r_cnt = nullptr;  // set bits per neighbour, this step
  std::vector<int64_t> s_seen, r_seen, s_slots, r_slots; // largest counts of eager steps; slots of the current messages
};

namespace cfx
{
namespace
{
__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ bits, int64_t i)
{
  return (bits[i >> 5] >> (i & 31)) & 1u;
}

// One warp per ghost row: which of the row's static candidate columns are in this step's pattern?  Mirrors the
// column rule of pattern_rows_kernel (sparsity.cu): a flagged incident cell brings its dofs; a band cell (flag
// bit 1) also brings the dofs of the cell across each of its facets that is in the facet integral.
template <int ND>
__global__ void __launch_bounds__(128)
    ghost_bits_kernel(int64_t n_rows, const int32_t* __restrict__ s_rows, const int64_t* __restrict__ s_ptr,
                      const int32_t* __restrict__ s_cols, const int64_t* __restrict__ s_bit,
                      const int64_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_cell,
                      const int32_t* __restrict__ dofmap, const uint8_t* __restrict__ cell_flags,
                      const int32_t* __restrict__ c2f, const int32_t* __restrict__ f2c2,
                      const int32_t* __restrict__ facet_slot, uint32_t* __restrict__ bits)
{
  constexpr int NF = (ND == 3 || ND == 6) ? 3 : 4;
  const int lane = threadIdx.x & 31;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x) >> 5;
  if (i >= n_rows)
    return;
  const int64_t r = s_rows[i];
  const int64_t cb = s_ptr[i];
  const int ncand = static_cast<int>(s_ptr[i + 1] - cb);
  const int64_t bb = s_bit[i];
  const int64_t ib = inc_ptr[r];
  const int n_inc = static_cast<int>(inc_ptr[r + 1] - ib);
  for (int k0 = 0; k0 < n_inc; k0 += 32)
  {
    const int k = k0 + lane;
    int32_t d[ND], od[NF][ND];
#pragma unroll
    for (int j = 0; j < ND; ++j)
      d[j] = -1;
#pragma unroll
    for (int lf = 0; lf < NF; ++lf)
#pragma unroll
      for (int j = 0; j < ND; ++j)
        od[lf][j] = -1;
    if (k < n_inc)
    {
      const int64_t c = inc_cell[ib + k];
      const uint8_t fl = cell_flags[c];
      if (fl != 0)
      {
#pragma unroll
        for (int j = 0; j < ND; ++j)
          d[j] = dofmap[c * ND + j];
      }
      if (fl & 2)
      {
        if (facet_slot != nullptr)
        {
#pragma unroll
          for (int lf = 0; lf < NF; ++lf)
          {
            const int64_t f = c2f[c * NF + lf];
            if (facet_slot[f] >= 0)
            {
              const int32_t c0 = f2c2[2 * f], c1 = f2c2[2 * f + 1];
              const int64_t oc = (c0 == c) ? c1 : c0;
              if (oc >= 0)
              {
#pragma unroll
                for (int j = 0; j < ND; ++j)
                  od[lf][j] = dofmap[oc * ND + j];
              }
            }
          }
        }
      }
    }
    for (int j = 0; j < ncand; ++j)
    {
      const int32_t cj = s_cols[cb + j];
      bool hit = false;
#pragma unroll
      for (int q = 0; q < ND; ++q)
        hit = hit || (d[q] == cj);
#pragma unroll
      for (int lf = 0; lf < NF; ++lf)
#pragma unroll
        for (int q = 0; q < ND; ++q)
          hit = hit || (od[lf][q] == cj);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && m)
        atomicOr(&bits[(bb + j) >> 5], 1u << ((bb + j) & 31));
    }
  }
}

// received candidates (insertion order) whose bit is set: compact.cuh predicate over 16 consecutive positions
struct RecvBitPred
{
  const int32_t* perm;
  const int64_t* r_bit;
  const uint32_t* bits;
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
#pragma unroll 4
    for (int k = 0; k < 16; ++k)
      if (base + k < n && bit_at(bits, r_bit[perm[base + k]]))
        m |= 1u << k;
    return m;
  }
};

__global__ void expand_entries_kernel(const int32_t* __restrict__ sel, DN n_, const int32_t* __restrict__ perm,
                                      const int32_t* __restrict__ r_row, const int32_t* __restrict__ r_col,
                                      int32_t* __restrict__ xrows, int32_t* __restrict__ xcols)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_.get())
    return;
  const int32_t e = perm[sel[i]];
  xrows[i] = r_row[e];
  xcols[i] = r_col[e];
}

__device__ __forceinline__ int64_t find_entry(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                              int64_t r, int32_t c)
{
  int64_t lo = row_ptr[r], hi = row_ptr[r + 1];
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < row_ptr[r + 1] && cols[lo] == c) ? lo : -1;
}

// one thread per (candidate entry, component of its bs x bs block); a block's components are consecutive in the
// matrix (values[p * bs * bs + q]) and in the message
__global__ void pack_values_kernel(int64_t n_ent, int bs2, const int32_t* __restrict__ s_rows,
                                   const int64_t* __restrict__ s_ptr, const int32_t* __restrict__ s_cols,
                                   const int32_t* __restrict__ s_erow, const int64_t* __restrict__ s_bit,
                                   const int64_t* __restrict__ s_vpos, const uint32_t* __restrict__ bits,
                                   const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                   const double* __restrict__ vals, double* __restrict__ out, int32_t* __restrict__ err)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t e = t / bs2;
  if (e >= n_ent)
    return;
  const int q = static_cast<int>(t - e * bs2);
  const int32_t i = s_erow[e];
  double v = 0.0;
  if (bit_at(bits, s_bit[i] + (e - s_ptr[i])))
  {
    const int64_t p = find_entry(row_ptr, cols, s_rows[i], s_cols[e]);
    if (p < 0)
    { // a ghost-row entry this rank announced is missing from its own matrix
      err[0] = 41;
      err[1] = s_rows[i];
    }
    else
      v = vals[p * bs2 + q];
  }
  out[s_vpos[e] + q] = v;
}

__global__ void pack_vector_kernel(int64_t n_rows, int bs, const int32_t* __restrict__ s_rows,
                                   const int64_t* __restrict__ s_vec_vpos, const double* __restrict__ b,
                                   double* __restrict__ out)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t i = t / bs;
  if (i >= n_rows)
    return;
  const int k = static_cast<int>(t - i * bs);
  out[s_vec_vpos[i] + k] = b ? b[static_cast<int64_t>(s_rows[i]) * bs + k] : 0.0;
}

// entries [e0, e1) of one neighbour: distinct matrix positions, so plain read-modify-write
__global__ void unpack_values_kernel(int64_t e0, int64_t e1, int bs2, const int32_t* __restrict__ r_row,
                                     const int32_t* __restrict__ r_col, const int64_t* __restrict__ r_bit,
                                     const int64_t* __restrict__ r_vpos, const uint32_t* __restrict__ bits,
                                     const double* __restrict__ in, const int64_t* __restrict__ row_ptr,
                                     const int32_t* __restrict__ cols, double* __restrict__ vals,
                                     int32_t* __restrict__ err)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t e = e0 + t / bs2;
  if (e >= e1 || !bit_at(bits, r_bit[e]))
    return;
  const int q = static_cast<int>(t % bs2);
  const int64_t p = find_entry(row_ptr, cols, r_row[e], r_col[e]);
  if (p < 0)
  {
    err[0] = 42;
    err[1] = r_row[e];
    return;
  }
  vals[p * bs2 + q] += in[r_vpos[e] + q];
}

__global__ void unpack_vector_kernel(int64_t i0, int64_t i1, int bs, const int32_t* __restrict__ r_vec_row,
                                     const int64_t* __restrict__ r_vec_vpos, const double* __restrict__ in,
                                     double* __restrict__ b)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t i = i0 + t / bs;
  if (i >= i1)
    return;
  const int k = static_cast<int>(t % bs);
  b[static_cast<int64_t>(r_vec_row[i]) * bs + k] += in[r_vec_vpos[i] + k];
}

// ---- compact value messages
__global__ void popc_words_kernel(const uint32_t* __restrict__ bits, int64_t n_words, int32_t* __restrict__ out)
{
  const int64_t w = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (w < n_words)
    out[w] = __popc(bits[w]);
}

__global__ void neighbour_counts_kernel(const int64_t* __restrict__ wpre, const int64_t* __restrict__ word_off,
                                        int n_neigh, int64_t* __restrict__ cnt)
{
  const int k = threadIdx.x;
  if (k < n_neigh)
    cnt[k] = wpre[word_off[k + 1]] - wpre[word_off[k]];
}

// slot of the set bit `bi` among the set bits of its neighbour's message (whose first word is wbase)
__device__ __forceinline__ int64_t compact_slot(const uint32_t* __restrict__ bits, const int64_t* __restrict__ wpre,
                                                int64_t wbase, int64_t bi)
{
  const int64_t w = bi >> 5;
  return wpre[w] - wpre[wbase] + __popc(bits[w] & ((1u << (bi & 31)) - 1u));
}

__global__ void pack_values_compact_kernel(int64_t eb, int64_t ee, int bs2, const int32_t* __restrict__ s_rows,
                                           const int64_t* __restrict__ s_ptr, const int32_t* __restrict__ s_cols,
                                           const int32_t* __restrict__ s_erow, const int64_t* __restrict__ s_bit,
                                           const uint32_t* __restrict__ bits, const int64_t* __restrict__ wpre,
                                           int64_t wbase, int64_t n_slots, const int64_t* __restrict__ row_ptr,
                                           const int32_t* __restrict__ cols, const double* __restrict__ vals,
                                           double* __restrict__ out, int32_t* __restrict__ err)
{
  // one thread per candidate: most bits are clear, a set one copies its bs x bs block (consecutive doubles)
  const int64_t e = eb + static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (e >= ee)
    return;
  const int32_t i = s_erow[e];
  const int64_t bi = s_bit[i] + (e - s_ptr[i]);
  if (!bit_at(bits, bi))
    return;
  const int64_t slot = compact_slot(bits, wpre, wbase, bi);
  if (slot >= n_slots)
  { // more ghost-row entries than the message was sized for (deferred-size mode: capacity exceeded)
    err[0] = 43;
    err[1] = s_rows[i];
    return;
  }
  const int64_t p = find_entry(row_ptr, cols, s_rows[i], s_cols[e]);
  if (p < 0)
  {
    err[0] = 41;
    err[1] = s_rows[i];
    return;
  }
  for (int q = 0; q < bs2; ++q)
    out[slot * bs2 + q] = vals[p * bs2 + q];
}

__global__ void pack_vector_compact_kernel(int64_t rb, int64_t re, int bs, const int32_t* __restrict__ s_rows,
                                           const double* __restrict__ b, double* __restrict__ out)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t i = rb + t / bs;
  if (i >= re)
    return;
  const int k = static_cast<int>(t % bs);
  out[(i - rb) * bs + k] = b ? b[static_cast<int64_t>(s_rows[i]) * bs + k] : 0.0;
}

__global__ void unpack_values_compact_kernel(int64_t e0, int64_t e1, int bs2, const int32_t* __restrict__ r_row,
                                             const int32_t* __restrict__ r_col, const int64_t* __restrict__ r_bit,
                                             const uint32_t* __restrict__ bits, const int64_t* __restrict__ wpre,
                                             int64_t wbase, int64_t n_slots, const double* __restrict__ in,
                                             const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                             double* __restrict__ vals, int32_t* __restrict__ err)
{
  const int64_t e = e0 + static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (e >= e1)
    return;
  const int64_t bi = r_bit[e];
  if (!bit_at(bits, bi))
    return;
  const int64_t slot = compact_slot(bits, wpre, wbase, bi);
  if (slot >= n_slots)
  {
    err[0] = 43;
    err[1] = r_row[e];
    return;
  }
  const int64_t p = find_entry(row_ptr, cols, r_row[e], r_col[e]);
  if (p < 0)
  {
    err[0] = 42;
    err[1] = r_row[e];
    return;
  }
  for (int q = 0; q < bs2; ++q)
    vals[p * bs2 + q] += in[slot * bs2 + q];
}

__global__ void unpack_vector_compact_kernel(int64_t i0, int64_t i1, int bs, const int32_t* __restrict__ r_vec_row,
                                             const double* __restrict__ in, double* __restrict__ b)
{
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t i = i0 + t / bs;
  if (i >= i1)
    return;
  const int k = static_cast<int>(t % bs);
  b[static_cast<int64_t>(r_vec_row[i]) * bs + k] += in[(i - i0) * bs + k];
}

template <class T>
void upload(cfx_ctx* c, DevBuf<T>& dst, const T* src, int64_t n)
{
  dst.reserve(c->pool, static_cast<size_t>(n > 0 ? n : 1));
  if (n > 0)
    CFX_CUDA(cudaMemcpyAsync(dst.p, src, static_cast<size_t>(n) * sizeof(T), cudaMemcpyHostToDevice, c->stream));
}

// Per-word prefix of the set bits of one side's bit buffer and this step's message layout of that side:
// val_off[k] .. val_off[k + 1] = [bs per ghost row of neighbour k][bs*bs per slot], slots = the exact number of set
// bits (eager step: one read-back) or the learned capacity (deferred-size step).
void compact_layout(cfx_ctx* c, cfx_xplan* P, bool send)
{
  const int nn = P->n_neigh;
  const std::vector<int64_t>& word_off = send ? P->s_word_off : P->r_word_off;
  const int64_t n_words = word_off[nn];
  DevBuf<int32_t>& wcnt = send ? P->s_wcnt : P->r_wcnt;
  DevBuf<int64_t>& wpre = send ? P->s_wpre : P->r_wpre;
  const uint32_t* bits = send ? P->bits_send.p : P->bits_recv.p;
  int64_t* d_cnt = send ? P->d_s_cnt : P->d_r_cnt;
  std::vector<int64_t>& seen = send ? P->s_seen : P->r_seen;
  std::vector<int64_t>& slots = send ? P->s_slots : P->r_slots;
  std::vector<int64_t>& val_off = send ? P->s_val_off : P->r_val_off;
  wcnt.reserve(c->pool, static_cast<size_t>(n_words) + 1);
  wpre.reserve(c->pool, static_cast<size_t>(n_words) + 2);
  if (n_words > 0)
    CFX_LAUNCH(c, popc_words_kernel, grid_for(n_words, 256), 256, 0, bits, n_words, wcnt.p);
  exclusive_scan_i32_to_i64(c, wcnt.p, n_words, wpre.p);
  if (!c->deferred)
  {
    CFX_LAUNCH(c, neighbour_counts_kernel, 1, 64, 0, wpre.p, (send ? P->d_s_word_off : P->d_r_word_off).p, nn, d_cnt);
    const int64_t* h = read_back(c, d_cnt, nn);
    for (int k = 0; k < nn; ++k)
    {
      slots[k] = h[k];
      seen[k] = std::max(seen[k], h[k]);
    }
  }
  else
    for (int k = 0; k < nn; ++k)
    {
      const int64_t n_cand = send ? P->s_ent_off[k + 1] - P->s_ent_off[k] : P->r_ent_off[k + 1] - P->r_ent_off[k];
      slots[k] = std::min(n_cand, with_margin(c, seen[k]));
    }
  const int64_t bs = P->bs, bs2 = bs * bs;
  for (int k = 0; k < nn; ++k)
  {
    const int64_t rows = send ? P->s_row_off[k + 1] - P->s_row_off[k] : P->r_row_off[k + 1] - P->r_row_off[k];
    val_off[k + 1] = val_off[k] + rows * bs + slots[k] * bs2;
  }
  (send ? P->vals_send : P->vals_recv).reserve(c->pool, static_cast<size_t>(val_off[nn]) + 1);
}
} // namespace
} // namespace cfx

extern "C"
{
cfx_status cfx_comm_unique_id(void* out128, const char* nccl_path)
{
  cfx_ctx* ctx = nullptr;
  CFX_API_BEGIN
  CFX_REQUIRE(out128 != nullptr, CFX_ERR_INVALID, "cfx_comm_unique_id: NULL argument");
  load_nccl(nccl_path);
  nccl_check(g_nccl.GetUniqueId(out128), "ncclGetUniqueId");
  CFX_API_END(ctx)
}

cfx_status cfx_comm_init(cfx_ctx* ctx, const void* unique_id128, int rank, int n_ranks, const char* nccl_path)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && unique_id128 && rank >= 0 && rank < n_ranks, CFX_ERR_INVALID, "cfx_comm_init: invalid arguments");
  CFX_REQUIRE(ctx->comm == nullptr, CFX_ERR_STATE, "cfx_comm_init: the context already has a communicator");
  load_nccl(nccl_path);
  CFX_CUDA(cudaSetDevice(ctx->device));
  UniqueId id;
  std::memcpy(id.bytes, unique_id128, 128);
  nccl_check(g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank), "ncclCommInitRank");
  ctx->comm_rank = rank;
  ctx->comm_size = n_ranks;
  CFX_API_END(ctx)
}

void cfx_comm_destroy(cfx_ctx* ctx)
{
  if (ctx && ctx->comm && g_nccl.CommDestroy)
  {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
}

cfx_status cfx_xplan_create(cfx_ctx* ctx, int space, int n_neigh, const int32_t* neigh_ranks,
                            const int64_t* s_row_off, const int32_t* s_rows, const int64_t* s_ptr, const int32_t* s_cols,
                            const int64_t* r_ent_off, const int32_t* r_row, const int32_t* r_col, const int32_t* r_perm,
                            const int64_t* r_row_off, const int32_t* r_vec_row, cfx_xplan** out)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && out && n_neigh >= 0 && space >= 0 && space < CFX_MAX_SPACES && ctx->spaces[space].bound,
              CFX_ERR_INVALID, "cfx_xplan_create: invalid arguments");
  cfx_xplan* P = new cfx_xplan();
  *out = P;
  P->space = space;
  P->bs = ctx->spaces[space].bs;
  const int64_t bs = P->bs, bs2 = bs * bs;
  P->n_neigh = n_neigh;
  P->ranks.assign(neigh_ranks, neigh_ranks + n_neigh);
  P->n_s_rows = n_neigh ? s_row_off[n_neigh] : 0;
  P->n_s_ent = P->n_s_rows ? s_ptr[P->n_s_rows] : 0;
  P->n_r_ent = n_neigh ? r_ent_off[n_neigh] : 0;
  P->n_r_rows = n_neigh ? r_row_off[n_neigh] : 0;
  P->r_ent_off.assign(r_ent_off, r_ent_off + n_neigh + 1);
  P->r_row_off.assign(r_row_off, r_row_off + n_neigh + 1);
  P->s_row_off.assign(s_row_off, s_row_off + n_neigh + 1);
  P->s_ent_off.assign(n_neigh + 1, 0);
  for (int k = 0; k <= n_neigh; ++k)
    P->s_ent_off[k] = P->n_s_rows ? s_ptr[s_row_off[k]] : 0;
  P->compact = !(bs == 1 && ctx->spaces[space].nd <= 4);
  CFX_REQUIRE(!P->compact || n_neigh <= 64, CFX_ERR_UNSUPPORTED, "cfx_xplan_create: at most 64 neighbours");
  // message layouts: per neighbour [bits: one per candidate, padded to 32-bit words]
  //                                [values: bs*bs per candidate, then bs per ghost row (vector entries)]
  P->s_word_off.assign(n_neigh + 1, 0);
  P->r_word_off.assign(n_neigh + 1, 0);
  P->s_val_off.assign(n_neigh + 1, 0);
  P->r_val_off.assign(n_neigh + 1, 0);
  std::vector<int64_t> h_s_bit(P->n_s_rows), h_s_vpos(P->n_s_ent), h_s_vec(P->n_s_rows), h_r_bit(P->n_r_ent),
      h_r_vpos(P->n_r_ent), h_r_vec(P->n_r_rows);
  std::vector<int32_t> h_s_erow(P->n_s_ent);
  for (int k = 0; k < n_neigh; ++k)
  {
    const int64_t rb = s_row_off[k], re = s_row_off[k + 1];
    const int64_t eb = s_ptr[rb], ee = s_ptr[re];
    for (int64_t i = rb; i < re; ++i)
    {
      CFX_REQUIRE(s_ptr[i + 1] >= s_ptr[i], CFX_ERR_INVALID, "cfx_xplan_create: candidate offsets must ascend");
      h_s_bit[i] = P->s_word_off[k] * 32 + (s_ptr[i] - eb);
      h_s_vec[i] = P->s_val_off[k] + (ee - eb) * bs2 + (i - rb) * bs;
      for (int64_t e = s_ptr[i]; e < s_ptr[i + 1]; ++e)
      {
        h_s_erow[e] = static_cast<int32_t>(i);
        h_s_vpos[e] = P->s_val_off[k] + (e - eb) * bs2;
      }
    }
    P->s_word_off[k + 1] = P->s_word_off[k] + (ee - eb + 31) / 32;
    P->s_val_off[k + 1] = P->s_val_off[k] + (ee - eb) * bs2 + (re - rb) * bs;
    const int64_t qb = r_ent_off[k], qe = r_ent_off[k + 1], vb = r_row_off[k], ve = r_row_off[k + 1];
    for (int64_t e = qb; e < qe; ++e)
    {
      h_r_bit[e] = P->r_word_off[k] * 32 + (e - qb);
      h_r_vpos[e] = P->r_val_off[k] + (e - qb) * bs2;
    }
    for (int64_t i = vb; i < ve; ++i)
      h_r_vec[i] = P->r_val_off[k] + (qe - qb) * bs2 + (i - vb) * bs;
    P->r_word_off[k + 1] = P->r_word_off[k] + (qe - qb + 31) / 32;
    P->r_val_off[k + 1] = P->r_val_off[k] + (qe - qb) * bs2 + (ve - vb) * bs;
  }
  upload(ctx, P->s_rows, s_rows, P->n_s_rows);
  upload(ctx, P->s_ptr, s_ptr, P->n_s_rows + 1);
  upload(ctx, P->s_cols, s_cols, P->n_s_ent);
  upload(ctx, P->s_erow, h_s_erow.data(), P->n_s_ent);
  upload(ctx, P->s_bit, h_s_bit.data(), P->n_s_rows);
  if (!P->compact)
  {
    upload(ctx, P->s_vpos, h_s_vpos.data(), P->n_s_ent);
    upload(ctx, P->s_vec_vpos, h_s_vec.data(), P->n_s_rows);
    upload(ctx, P->r_vpos, h_r_vpos.data(), P->n_r_ent);
    upload(ctx, P->r_vec_vpos, h_r_vec.data(), P->n_r_rows);
  }
  else
  {
    upload(ctx, P->d_s_word_off, P->s_word_off.data(), n_neigh + 1);
    upload(ctx, P->d_r_word_off, P->r_word_off.data(), n_neigh + 1);
    P->d_s_cnt = alloc_count_slot(ctx, 64);
    P->d_r_cnt = alloc_count_slot(ctx, 64);
    for (auto* v : {&P->s_seen, &P->r_seen, &P->s_slots, &P->r_slots})
      v->assign(n_neigh, 0);
    for (int k = 0; k <= n_neigh; ++k) // until the first step lays the messages out: vector entries only
    {
      P->s_val_off[k] = s_row_off[k] * bs;
      P->r_val_off[k] = r_row_off[k] * bs;
    }
  }
  upload(ctx, P->r_row, r_row, P->n_r_ent);
  upload(ctx, P->r_col, r_col, P->n_r_ent);
  upload(ctx, P->r_perm, r_perm, P->n_r_ent);
  upload(ctx, P->r_bit, h_r_bit.data(), P->n_r_ent);
  upload(ctx, P->r_vec_row, r_vec_row, P->n_r_rows);
  P->bits_send.reserve(ctx->pool, static_cast<size_t>(P->s_word_off[n_neigh]) + 1);
  P->bits_recv.reserve(ctx->pool, static_cast<size_t>(P->r_word_off[n_neigh]) + 1);
  P->vals_send.reserve(ctx->pool, static_cast<size_t>(P->s_val_off[n_neigh]) + 1);
  P->vals_recv.reserve(ctx->pool, static_cast<size_t>(P->r_val_off[n_neigh]) + 1);
  CFX_CUDA(cudaMemsetAsync(P->bits_recv.p, 0, (static_cast<size_t>(P->r_word_off[n_neigh]) + 1) * sizeof(uint32_t),
                           ctx->stream));
  // the selection list has the static capacity "every candidate": it never overflows and never needs learning
  P->sel.reserve(ctx->pool, static_cast<size_t>(std::max<int64_t>(P->n_r_ent, 256)));
  P->d_n_sel = alloc_count_slot(ctx);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream)); // the host staging vectors go out of scope
  CFX_API_END(ctx)
}

void cfx_xplan_free(cfx_ctx* ctx, cfx_xplan* P)
{
  if (!P)
    return;
  if (ctx)
    cudaStreamSynchronize(ctx->stream);
  for (auto* b : {&P->s_rows, &P->s_cols, &P->s_erow, &P->r_row, &P->r_col, &P->r_perm, &P->r_vec_row, &P->sel})
    b->release();
  for (auto* b : {&P->s_ptr, &P->s_bit, &P->s_vpos, &P->s_vec_vpos, &P->r_bit, &P->r_vpos, &P->r_vec_vpos})
    b->release();
  P->bits_send.release();
  P->bits_recv.release();
  P->vals_send.release();
  P->vals_recv.release();
  for (auto* b : {&P->s_wpre, &P->r_wpre, &P->d_s_word_off, &P->d_r_word_off})
    b->release();
  P->s_wcnt.release();
  P->r_wcnt.release();
  if (P->d_s_cnt)
    free_count_slot(ctx, P->d_s_cnt, 64);
  if (P->d_r_cnt)
    free_count_slot(ctx, P->d_r_cnt, 64);
  free_count_slot(ctx, P->d_n_sel);
  delete P;
}

// message buffers of neighbour k (for transports other than NCCL, e.g. ranks emulated on one GPU):
// which = 0 pattern bits to send, 1 pattern bits received, 2 values to send, 3 values received
cfx_status cfx_xplan_buffer(cfx_ctx* ctx, const cfx_xplan* P, int which, int k, void** ptr, int64_t* bytes)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && P && ptr && bytes && k >= 0 && k < P->n_neigh && which >= 0 && which <= 3, CFX_ERR_INVALID,
              "cfx_xplan_buffer: invalid arguments");
  switch (which)
  {
  case 0: *ptr = P->bits_send.p + P->s_word_off[k]; *bytes = (P->s_word_off[k + 1] - P->s_word_off[k]) * 4; break;
  case 1: *ptr = P->bits_recv.p + P->r_word_off[k]; *bytes = (P->r_word_off[k + 1] - P->r_word_off[k]) * 4; break;
  case 2: *ptr = P->vals_send.p + P->s_val_off[k]; *bytes = (P->s_val_off[k + 1] - P->s_val_off[k]) * 8; break;
  default: *ptr = P->vals_recv.p + P->r_val_off[k]; *bytes = (P->r_val_off[k + 1] - P->r_val_off[k]) * 8; break;
  }
  CFX_API_END(ctx)
}

// grouped ncclSend / ncclRecv with every neighbour on the context's stream; which = 0 pattern bits, 1 values
cfx_status cfx_xplan_exchange(cfx_ctx* ctx, cfx_xplan* P, int which)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && P && (which == 0 || which == 1), CFX_ERR_INVALID, "cfx_xplan_exchange: invalid arguments");
  if (P->n_neigh == 0)
    return CFX_OK;
  CFX_REQUIRE(ctx->comm != nullptr, CFX_ERR_STATE, "cfx_xplan_exchange: call cfx_comm_init first");
  StageScope st(ctx, which == 0 ? "exchange_pattern" : "exchange_values");
  nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
  for (int k = 0; k < P->n_neigh; ++k)
  {
    if (which == 0)
    {
      const int64_t ns = P->s_word_off[k + 1] - P->s_word_off[k], nr = P->r_word_off[k + 1] - P->r_word_off[k];
      if (ns > 0)
        nccl_check(g_nccl.Send(P->bits_send.p + P->s_word_off[k], static_cast<size_t>(ns), NCCL_UINT32, P->ranks[k],
                               ctx->comm, ctx->stream), "ncclSend");
      if (nr > 0)
        nccl_check(g_nccl.Recv(P->bits_recv.p + P->r_word_off[k], static_cast<size_t>(nr), NCCL_UINT32, P->ranks[k],
                               ctx->comm, ctx->stream), "ncclRecv");
    }
    else
    {
      const int64_t ns = P->s_val_off[k + 1] - P->s_val_off[k], nr = P->r_val_off[k + 1] - P->r_val_off[k];
      if (ns > 0)
        nccl_check(g_nccl.Send(P->vals_send.p + P->s_val_off[k], static_cast<size_t>(ns), NCCL_FLOAT64, P->ranks[k],
                               ctx->comm, ctx->stream), "ncclSend");
      if (nr > 0)
        nccl_check(g_nccl.Recv(P->vals_recv.p + P->r_val_off[k], static_cast<size_t>(nr), NCCL_FLOAT64, P->ranks[k],
                               ctx->comm, ctx->stream), "ncclRecv");
    }
  }
  nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
  ++ctx->launches;
  CFX_API_END(ctx)
}

// SparsityPattern::finalize(), sender half: one bit per static candidate of this rank's ghost rows
cfx_status cfx_xplan_pack_pattern(cfx_ctx* ctx, cfx_xplan* P, const cfx_form* a_const)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && P && a && a->rank == 2 && a->space == P->space, CFX_ERR_INVALID,
              "cfx_xplan_pack_pattern: invalid arguments");
  if (P->n_neigh == 0)
    return CFX_OK;
  const Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a, false); // the flags only: the row lists are built once the neighbours' entries are in
  StageScope st(ctx, "ghost_row_bits", 4.0 * static_cast<double>(P->n_s_ent));
  const cfx_integral* FI = facet_integral_domain(a);
  set_facet_slots(ctx, FI, false);
  CFX_CUDA(cudaMemsetAsync(P->bits_send.p, 0, (static_cast<size_t>(P->s_word_off[P->n_neigh]) + 1) * sizeof(uint32_t),
                           ctx->stream));
  if (P->n_s_rows > 0)
  {
    const unsigned g = grid_for(P->n_s_rows, 4);
    const int32_t* fs = FI ? ctx->facet_slot.p : nullptr;
#define GB_ARGS                                                                                                        \
  P->n_s_rows, P->s_rows.p, P->s_ptr.p, P->s_cols.p, P->s_bit.p, S.inc_ptr.p, S.inc_cell.p, S.dofmap,                  \
      a->prep->cell_flags.p, ctx->c2f, ctx->f2c2.p, fs, P->bits_send.p
    switch (S.nd)
    {
    case 3: CFX_LAUNCH(ctx, ghost_bits_kernel<3>, g, 128, 0, GB_ARGS); break;
    case 4: CFX_LAUNCH(ctx, ghost_bits_kernel<4>, g, 128, 0, GB_ARGS); break;
    case 6: CFX_LAUNCH(ctx, ghost_bits_kernel<6>, g, 128, 0, GB_ARGS); break;
    case 10: CFX_LAUNCH(ctx, ghost_bits_kernel<10>, g, 128, 0, GB_ARGS); break;
    default: throw Error(CFX_ERR_UNSUPPORTED, "cfx_xplan_pack_pattern: unsupported element");
    }
#undef GB_ARGS
  }
  set_facet_slots(ctx, FI, true);
  if (P->compact)
    compact_layout(ctx, P, true);
  CFX_API_END(ctx)
}

// SparsityPattern::finalize(), owner half: the received bits become the form's inserted entries (device side)
cfx_status cfx_xplan_insert_pattern(cfx_ctx* ctx, cfx_xplan* P, cfx_form* a)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && P && a && a->rank == 2 && a->space == P->space, CFX_ERR_INVALID,
              "cfx_xplan_insert_pattern: invalid arguments");
  if (P->n_neigh == 0)
    return CFX_OK;
  StageScope st(ctx, "insert_received_pattern", 12.0 * static_cast<double>(P->n_r_ent));
  if (P->compact)
    compact_layout(ctx, P, false);
  if (P->n_r_ent == 0)
    return CFX_OK;
  RecvBitPred pred{P->r_perm.p, P->r_bit.p, P->bits_recv.p};
  bool deferred = false;
  const int64_t n = compact_indices(ctx, dn_exact(P->n_r_ent), pred, P->sel, false, P->d_n_sel, &deferred);
  a->xrows.reserve(ctx->pool, static_cast<size_t>(P->n_r_ent));
  a->xcols.reserve(ctx->pool, static_cast<size_t>(P->n_r_ent));
  a->n_x = n;
  a->d_n_x = deferred ? P->d_n_sel : nullptr;
  if (deferred)
    a->deferred = true;
  if (n > 0)
    CFX_LAUNCH(ctx, expand_entries_kernel, grid_for(n, 256), 256, 0, P->sel.p, DN{a->d_n_x, n, 0}, P->r_perm.p,
               P->r_row.p, P->r_col.p, a->xrows.p, a->xcols.p);
  a->dirty_x_only = !a->dirty && a->prep != nullptr;
  a->dirty = true;
  CFX_API_END(ctx)
}

// MatrixCSR::scatter_rev + Vector::scatter_rev(add), sender half (b may be NULL)
cfx_status cfx_xplan_pack_values(cfx_ctx* ctx, cfx_xplan* P, const cfx_pattern* A, const double* b)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && P && A && A->space == P->space && A->bs == P->bs, CFX_ERR_INVALID,
              "cfx_xplan_pack_values: invalid arguments");
  const int bs = P->bs, bs2 = bs * bs;
  if (P->n_neigh == 0)
    return CFX_OK;
  settle_values(ctx, const_cast<cfx_pattern*>(A));
  StageScope st(ctx, "pack_ghost_values", (12.0 + 8.0 * bs2) * static_cast<double>(P->n_s_ent));
  if (P->compact)
  {
    for (int k = 0; k < P->n_neigh; ++k)
    {
      const int64_t rb = P->s_row_off[k], re = P->s_row_off[k + 1], eb = P->s_ent_off[k], ee = P->s_ent_off[k + 1];
      double* msg = P->vals_send.p + P->s_val_off[k];
      if (re > rb)
        CFX_LAUNCH(ctx, pack_vector_compact_kernel, grid_for((re - rb) * bs, 256), 256, 0, rb, re, bs, P->s_rows.p, b,
                   msg);
      if (ee > eb)
        CFX_LAUNCH(ctx, pack_values_compact_kernel, grid_for(ee - eb, 256), 256, 0, eb, ee, bs2, P->s_rows.p,
                   P->s_ptr.p, P->s_cols.p, P->s_erow.p, P->s_bit.p, P->bits_send.p, P->s_wpre.p, P->s_word_off[k],
                   P->s_slots[k], A->row_ptr.p, A->cols.p, A->values.p, msg + (re - rb) * bs, ctx->err_flag.p);
    }
    return CFX_OK;
  }
  if (P->n_s_ent > 0)
    CFX_LAUNCH(ctx, pack_values_kernel, grid_for(P->n_s_ent * bs2, 256), 256, 0, P->n_s_ent, bs2, P->s_rows.p,
               P->s_ptr.p, P->s_cols.p, P->s_erow.p, P->s_bit.p, P->s_vpos.p, P->bits_send.p, A->row_ptr.p, A->cols.p,
               A->values.p, P->vals_send.p, ctx->err_flag.p);
  if (P->n_s_rows > 0)
    CFX_LAUNCH(ctx, pack_vector_kernel, grid_for(P->n_s_rows * bs, 256), 256, 0, P->n_s_rows, bs, P->s_rows.p,
               P->s_vec_vpos.p, b, P->vals_send.p);
  CFX_API_END(ctx)
}

// ... owner half: add what the neighbours sent, neighbour by neighbour in the plan's (rank) order
cfx_status cfx_xplan_unpack_add(cfx_ctx* ctx, cfx_xplan* P, cfx_pattern* A, double* b)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && P && A && A->space == P->space && A->bs == P->bs, CFX_ERR_INVALID,
              "cfx_xplan_unpack_add: invalid arguments");
  const int bs = P->bs, bs2 = bs * bs;
  if (P->n_neigh == 0)
    return CFX_OK;
  settle_values(ctx, A);
  A->values_zero = false;
  StageScope st(ctx, "unpack_add_ghost_values", (12.0 + 16.0 * bs2) * static_cast<double>(P->n_r_ent));
  for (int k = 0; k < P->n_neigh; ++k)
  {
    const int64_t e0 = P->r_ent_off[k], e1 = P->r_ent_off[k + 1], i0 = P->r_row_off[k], i1 = P->r_row_off[k + 1];
    if (P->compact)
    {
      const double* msg = P->vals_recv.p + P->r_val_off[k];
      if (e1 > e0)
        CFX_LAUNCH(ctx, unpack_values_compact_kernel, grid_for(e1 - e0, 256), 256, 0, e0, e1, bs2, P->r_row.p,
                   P->r_col.p, P->r_bit.p, P->bits_recv.p, P->r_wpre.p, P->r_word_off[k], P->r_slots[k],
                   msg + (i1 - i0) * bs, A->row_ptr.p, A->cols.p, A->values.p, ctx->err_flag.p);
      if (b && i1 > i0)
        CFX_LAUNCH(ctx, unpack_vector_compact_kernel, grid_for((i1 - i0) * bs, 256), 256, 0, i0, i1, bs,
                   P->r_vec_row.p, msg, b);
      continue;
    }
    if (e1 > e0)
      CFX_LAUNCH(ctx, unpack_values_kernel, grid_for((e1 - e0) * bs2, 256), 256, 0, e0, e1, bs2, P->r_row.p, P->r_col.p,
                 P->r_bit.p, P->r_vpos.p, P->bits_recv.p, P->vals_recv.p, A->row_ptr.p, A->cols.p, A->values.p,
                 ctx->err_flag.p);
    if (b && i1 > i0)
      CFX_LAUNCH(ctx, unpack_vector_kernel, grid_for((i1 - i0) * bs, 256), 256, 0, i0, i1, bs, P->r_vec_row.p,
                 P->r_vec_vpos.p, P->vals_recv.p, b);
  }
  check_call(ctx, "cfx_xplan_unpack_add (received entry not in the sparsity pattern)");
  CFX_API_END(ctx)
}
} // extern "C"
