// compact.cuh -- order-preserving stream compaction "indices i in [0,n) with pred(i)" -> ascending
// int32 list.  Three launches: per-tile count, single-block scan of the tile counts, write.
// No atomics, so the list is identical from run to run (the reference builds the same lists with a
// serial push_back loop, cut.cpp:887-921).
#pragma once
#include "common.cuh"

namespace cfx
{
// Pred: __device__ unsigned operator()(int64_t base, int64_t n) -> 16-bit mask for base..base+15
// (base is a multiple of 16; bits for indices >= n must be 0).  One 16-byte load per thread for
// byte-array predicates, so a 100 M-cell scan streams at HBM rate.
constexpr int CP_BLOCK = 256;
constexpr int CP_ITEMS = 16;
constexpr int CP_TILE = CP_BLOCK * CP_ITEMS;
template <class Pred>
__global__ void __launch_bounds__(CP_BLOCK) compact_count_kernel(Pred pred, DN n_, int32_t* __restrict__ counts)
{
  const int64_t n = n_.get();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * CP_TILE + threadIdx.x * CP_ITEMS;
  int cnt = base < n ? __popc(pred(base, n)) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  __shared__ int s[CP_BLOCK / 32];
  if ((threadIdx.x & 31) == 0)
    s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int t = 0;
#pragma unroll
    for (int w = 0; w < CP_BLOCK / 32; ++w)
      t += s[w];
    counts[blockIdx.x] = t;
  }
}

template <class Pred>
__global__ void __launch_bounds__(CP_BLOCK)
    compact_write_kernel(Pred pred, DN n_, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ d_total,
                         int32_t* __restrict__ out)
{
  const int64_t n = n_.get();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * CP_TILE + threadIdx.x * CP_ITEMS;
  unsigned mask = base < n ? pred(base, n) : 0u;
  int tot;
  const int excl = block_excl_scan<CP_BLOCK>(__popc(mask), &tot);
  // the tile's selected indices are staged in shared memory and leave as one contiguous, coalesced run
  // (a thread storing its up to 16 hits one by one issues 16 strided store instructions per warp)
  __shared__ int32_t s_out[CP_TILE];
  int o = excl;
  while (mask)
  {
    const int k = __ffs(mask) - 1;
    mask &= mask - 1;
    s_out[o++] = static_cast<int32_t>(base + k);
  }
  __syncthreads();
  // a total that exceeded the output capacity was replaced by 0 (scan_block_counts): nothing is written then
  if (tile_off[blockIdx.x] + tot > *d_total)
    return;
  int32_t* dst = out + tile_off[blockIdx.x];
  for (int i = threadIdx.x; i < tot; i += CP_BLOCK)
    dst[i] = s_out[i];
}

// Returns the number of selected indices.  `counted`: c->blk_counts already holds the per-tile counts (the caller
// reserved it for grid_for(n, CP_TILE) tiles and filled it), so the counting pass over the predicate is skipped.
// `d_total`: device slot that receives the exact total (null: context scratch).
// Deferred-size mode: when `out` already has a capacity and `deferred` is given, the total stays on the device --
// the return value is the capacity (an upper bound), *deferred = true, and a total above the capacity raises the
// device error flag and selects nothing (no host round trip anywhere in that case).
template <class Pred>
int64_t compact_indices(cfx_ctx* c, DN n, Pred pred, DevBuf<int32_t>& out, bool counted = false,
                        int64_t* d_total = nullptr, bool* deferred = nullptr)
{
  if (deferred)
    *deferred = false;
  c->scratch64.reserve(c->pool, 64);
  if (!d_total)
    d_total = c->scratch64.p;
  if (n.h <= 0)
  {
    out.reserve(c->pool, 1);
    CFX_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), c->stream));
    return 0;
  }
  const unsigned nb = grid_for(n.h, CP_TILE);
  c->blk_counts.reserve(c->pool, nb);
  c->blk_offsets.reserve(c->pool, nb);
  if (!counted)
    CFX_LAUNCH(c, compact_count_kernel<Pred>, nb, CP_BLOCK, 0, pred, n, c->blk_counts.p);
  const bool defer = c->deferred && deferred != nullptr && out.p != nullptr && out.cap >= 256;
  scan_block_counts(c, c->blk_counts.p, nb, c->blk_offsets.p, d_total, defer ? static_cast<int64_t>(out.cap) : -1);
  int64_t total;
  if (defer)
  {
    total = static_cast<int64_t>(out.cap);
    *deferred = true;
  }
  else
  {
    total = read_back(c, d_total, 1)[0];
    if (static_cast<size_t>(total) > out.cap || !out.p)
      out.reserve(c->pool, static_cast<size_t>(with_margin(c, total)));
  }
  if (total > 0)
    CFX_LAUNCH(c, compact_write_kernel<Pred>, nb, CP_BLOCK, 0, pred, n, c->blk_offsets.p, d_total, out.p);
  return total;
}
template <class Pred>
int64_t compact_indices(cfx_ctx* c, int64_t n, Pred pred, DevBuf<int32_t>& out, bool counted = false)
{
  return compact_indices(c, dn_exact(n), pred, out, counted);
}

// byte-array predicate: flag[i] != 0
__device__ __forceinline__ unsigned nonzero_bytes16(uint4 v)
{
  unsigned m = 0;
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      m |= ((w[q] >> (8 * k)) & 0xffu) ? (1u << (4 * q + k)) : 0u;
  return m;
}

struct FlagPred
{
  const uint8_t* flag; // 16-byte aligned base
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    if (base + 16 <= n)
      return nonzero_bytes16(*reinterpret_cast<const uint4*>(flag + base));
    unsigned m = 0;
    for (int k = 0; base + k < n; ++k)
      m |= flag[base + k] ? (1u << k) : 0u;
    return m;
  }
};

// compiled selector (cutcells::SelectionExpr after compile_selection_expr, cut.cpp:881-882)
struct Dnf
{
  int n_terms;
  int term_off[CFX_MAX_CLAUSES + 1];
  int8_t ls[CFX_MAX_CLAUSES];
  uint8_t relmask[CFX_MAX_CLAUSES]; // bit d set <=> relation matches domain code d (cut.cpp:323-342)
};

__host__ __device__ inline uint8_t relation_mask(int rel)
{
  switch (rel)
  {
  case CFX_REL_LT: return 1u << CFX_DOMAIN_INSIDE;
  case CFX_REL_LE: return (1u << CFX_DOMAIN_INSIDE) | (1u << CFX_DOMAIN_INTERSECTED);
  case CFX_REL_GT: return 1u << CFX_DOMAIN_OUTSIDE;
  case CFX_REL_GE: return (1u << CFX_DOMAIN_OUTSIDE) | (1u << CFX_DOMAIN_INTERSECTED);
  case CFX_REL_EQ: return 1u << CFX_DOMAIN_INTERSECTED;
  }
  return 0;
}

__device__ __forceinline__ bool dnf_match(const Dnf& d, const int8_t* __restrict__ domain, int64_t stride, int64_t cell)
{
  for (int t = 0; t < d.n_terms; ++t)
  {
    bool ok = true;
    for (int k = d.term_off[t]; k < d.term_off[t + 1]; ++k)
    {
      const int dom = domain[static_cast<int64_t>(d.ls[k]) * stride + cell];
      ok = ok && ((d.relmask[k] >> dom) & 1u);
    }
    if (ok)
      return true;
  }
  return false;
}

struct DnfPred
{
  Dnf d;
  const int8_t* domain;
  int64_t stride; // multiple of 16
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
    if (d.n_terms == 1 && d.term_off[1] == 1 && base + 16 <= n)
    { // single clause fast path: one 16-byte load of sixteen domain codes
      const uint4 v = *reinterpret_cast<const uint4*>(domain + static_cast<int64_t>(d.ls[0]) * stride + base);
      const unsigned rm = d.relmask[0];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          m |= ((rm >> ((w[q] >> (8 * k)) & 0x7u)) & 1u) << (4 * q + k);
      return m;
    }
    for (int k = 0; k < 16 && base + k < n; ++k)
      m |= dnf_match(d, domain, stride, base + k) ? (1u << k) : 0u;
    return m;
  }
};

Dnf make_dnf(cfx_ctx* c, int n_terms, const int32_t* term_offsets, const int32_t* clause_ls,
             const int32_t* clause_rel); // classify.cu
} // namespace cfx
