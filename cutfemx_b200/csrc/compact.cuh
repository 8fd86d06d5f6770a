// compact.cuh -- order-preserving stream compaction "indices i in [0,n) with pred(i)" -> ascending
// int32 list.  Three launches: per-tile count, single-block scan of the tile counts, write.
// No atomics, so the list is identical from run to run (the reference builds the same lists with a
// serial push_back loop, cut.cpp:887-921).
#pragma once
#include "common.cuh"

namespace cfx
{
// Pred: __device__ unsigned operator()(int64_t base, int64_t n) -> 4-bit mask for base..base+3
// (base is a multiple of 4; bits for indices >= n must be 0).
template <class Pred>
__global__ void __launch_bounds__(SCAN_BLOCK) compact_count_kernel(Pred pred, int64_t n, int32_t* __restrict__ counts)
{
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int cnt = base < n ? __popc(pred(base, n)) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  __shared__ int s[SCAN_BLOCK / 32];
  if ((threadIdx.x & 31) == 0)
    s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int t = 0;
#pragma unroll
    for (int w = 0; w < SCAN_BLOCK / 32; ++w)
      t += s[w];
    counts[blockIdx.x] = t;
  }
}

template <class Pred>
__global__ void __launch_bounds__(SCAN_BLOCK)
    compact_write_kernel(Pred pred, int64_t n, const int64_t* __restrict__ tile_off, int32_t* __restrict__ out)
{
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  const unsigned mask = base < n ? pred(base, n) : 0u;
  int tot;
  const int excl = block_excl_scan<SCAN_BLOCK>(__popc(mask), &tot);
  int64_t o = tile_off[blockIdx.x] + excl;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (mask & (1u << k))
      out[o++] = static_cast<int32_t>(base + k);
}

// Returns the number of selected indices; `out` is (re)allocated to exactly that size.
template <class Pred>
int64_t compact_indices(cfx_ctx* c, int64_t n, Pred pred, DevBuf<int32_t>& out)
{
  if (n <= 0)
  {
    out.reserve(c->pool, 1);
    return 0;
  }
  const unsigned nb = grid_for(n, SCAN_TILE);
  c->blk_counts.reserve(c->pool, nb);
  c->blk_offsets.reserve(c->pool, nb);
  CFX_LAUNCH(c, compact_count_kernel<Pred>, nb, SCAN_BLOCK, 0, pred, n, c->blk_counts.p);
  scan_block_counts(c, c->blk_counts.p, nb, c->blk_offsets.p);
  const int64_t total = read_back(c, c->scratch64.p, 1)[0];
  out.reserve(c->pool, static_cast<size_t>(total > 0 ? total : 1));
  if (total > 0)
    CFX_LAUNCH(c, compact_write_kernel<Pred>, nb, SCAN_BLOCK, 0, pred, n, c->blk_offsets.p, out.p);
  return total;
}

// byte-array predicate: flag[i] != 0
struct FlagPred
{
  const uint8_t* flag; // 4-byte aligned base
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
    if (base + 4 <= n)
    {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(flag + base);
      m = ((v & 0xffu) ? 1u : 0u) | ((v & 0xff00u) ? 2u : 0u) | ((v & 0xff0000u) ? 4u : 0u)
          | ((v & 0xff000000u) ? 8u : 0u);
    }
    else
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
  int64_t stride;
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
    if (d.n_terms == 1 && d.term_off[1] == 1 && base + 4 <= n)
    { // single clause fast path: one 32-bit load of four domain codes
      const uint32_t v
          = *reinterpret_cast<const uint32_t*>(domain + static_cast<int64_t>(d.ls[0]) * stride + base);
      const unsigned rm = d.relmask[0];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        m |= ((rm >> ((v >> (8 * k)) & 0xffu)) & 1u) << k;
      return m;
    }
    for (int k = 0; k < 4 && base + k < n; ++k)
      m |= dnf_match(d, domain, stride, base + k) ? (1u << k) : 0u;
    return m;
  }
};

Dnf make_dnf(cfx_ctx* c, int n_terms, const int32_t* term_offsets, const int32_t* clause_ls,
             const int32_t* clause_rel); // classify.cu
} // namespace cfx
