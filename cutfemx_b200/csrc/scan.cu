// scan.cu -- device-wide exclusive scans (hand-written; no CUB).  All counts are integers,
// so every result is bit-reproducible.
#include "common.cuh"

namespace cfx
{
namespace
{
constexpr int SB = 1024;

// single block: offsets[i] = sum_{j<i} in[j]; *total = sum.  n up to a few 10^5.
// Each thread scans SI consecutive items per round, so 24.5 K tile counts take 3 rounds.
constexpr int SI = 8;
template <class Tin>
__global__ void __launch_bounds__(SB) scan_small_kernel(const Tin* __restrict__ in, int64_t n,
                                                        int64_t* __restrict__ offsets, int64_t* __restrict__ total,
                                                        int64_t cap, int32_t* __restrict__ err)
{
  __shared__ long long s_warp[SB / 32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0)
    s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += SB * SI)
  {
    const int64_t i0 = base + static_cast<int64_t>(threadIdx.x) * SI;
    long long v[SI];
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < SI; ++k)
    {
      v[k] = (i0 + k < n) ? static_cast<long long>(in[i0 + k]) : 0;
      sum += v[k];
    }
    const long long incl = warp_incl_scan_ll(sum);
    if (lane == 31)
      s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
      long long w = s_warp[lane];
      long long wi = warp_incl_scan_ll(w);
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    long long run = s_carry + s_warp[wid] + incl - sum;
#pragma unroll
    for (int k = 0; k < SI; ++k)
    {
      if (i0 + k < n)
        offsets[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == SB - 1)
      s_carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    long long t = s_carry;
    if (cap >= 0 && t > cap)
    { // deferred-size mode: the result would not fit the buffer the consumers were given
      err[0] = 31;
      err[1] = static_cast<int32_t>(t > 0x7fffffffLL ? 0x7fffffffLL : t);
      t = 0;
    }
    *total = t;
  }
}

constexpr int TB = 256;
constexpr int TILE = TB * 4;

template <class Tin>
__global__ void __launch_bounds__(TB) tile_sum_kernel(const Tin* __restrict__ in, DN n_,
                                                      int64_t* __restrict__ sums)
{
  const int64_t n = n_.get();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * TILE + threadIdx.x * 4;
  long long v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n)
      v += in[base + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  __shared__ long long s[TB / 32];
  if ((threadIdx.x & 31) == 0)
    s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    long long t = 0;
#pragma unroll
    for (int w = 0; w < TB / 32; ++w)
      t += s[w];
    sums[blockIdx.x] = t;
  }
}

// exclusive scan of one long long per thread over a block of TB threads
__device__ __forceinline__ long long block_excl_scan_ll(long long v)
{
  __shared__ long long s_w[TB / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long incl = warp_incl_scan_ll(v);
  if (lane == 31)
    s_w[wid] = incl;
  __syncthreads();
  if (wid == 0)
  {
    long long w = lane < TB / 32 ? s_w[lane] : 0;
    long long wi = warp_incl_scan_ll(w);
    if (lane < TB / 32)
      s_w[lane] = wi - w;
  }
  __syncthreads();
  const long long r = s_w[wid] + incl - v;
  __syncthreads();
  return r;
}

template <class Tin>
__global__ void __launch_bounds__(TB) tile_scan_kernel(const Tin* __restrict__ in, DN n_,
                                                       const int64_t* __restrict__ tile_off,
                                                       const int64_t* __restrict__ total, int64_t* __restrict__ out)
{
  const int64_t n = n_.get();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * TILE + threadIdx.x * 4;
  long long v[4];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    v[k] = (base + k < n) ? static_cast<long long>(in[base + k]) : 0;
    s += v[k];
  }
  const long long excl = block_excl_scan_ll(s);
  long long run = tile_off[blockIdx.x] + excl;
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    if (base + k < n)
      out[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    out[n] = *total;
}

// ---- single-pass scan (decoupled look-back).  Tiles take their index from a ticket counter, so every tile a block
// waits for has already started; a tile publishes its aggregate (flag 1), looks back over its predecessors' status
// words (a warp at a time) until it meets an inclusive prefix (flag 2), then publishes its own.  Integer sums: the
// result does not depend on the order the tiles finish in.
constexpr int LB = 256, LI = 8, LTILE = LB * LI;
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = (1ull << 62) - 1;

template <class Tin>
__global__ void __launch_bounds__(LB) lookback_scan_kernel(const Tin* __restrict__ in, DN n_,
                                                           unsigned long long* __restrict__ status,
                                                           int64_t* __restrict__ total, int64_t* __restrict__ out)
{
  __shared__ long long s_w[LB / 32];
  __shared__ long long s_prefix;
  __shared__ unsigned s_tile;
  const int64_t n = n_.get();
  if (threadIdx.x == 0)
    s_tile = static_cast<unsigned>(atomicAdd(status, 1ull));
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t ntiles = n > 0 ? (n + LTILE - 1) / LTILE : 1;
  if (tile >= ntiles)
    return;
  volatile unsigned long long* st = status + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t base = tile * LTILE + static_cast<int64_t>(threadIdx.x) * LI;
  long long v[LI];
  long long sum = 0;
#pragma unroll
  for (int k = 0; k < LI; ++k)
  {
    v[k] = (base + k < n) ? static_cast<long long>(in[base + k]) : 0;
    sum += v[k];
  }
  const long long incl = warp_incl_scan_ll(sum);
  if (lane == 31)
    s_w[wid] = incl;
  __syncthreads();
  if (wid == 0)
  {
    const long long w = lane < LB / 32 ? s_w[lane] : 0;
    const long long wi = warp_incl_scan_ll(w);
    const long long agg = __shfl_sync(0xffffffffu, wi, LB / 32 - 1); // the tile's aggregate
    if (lane < LB / 32)
      s_w[lane] = wi - w;
    long long prefix = 0;
    if (tile == 0)
    {
      if (lane == 0)
        st[0] = ST_INC | static_cast<unsigned long long>(agg);
    }
    else
    {
      if (lane == 0)
        st[tile] = ST_AGG | static_cast<unsigned long long>(agg);
      int64_t t = tile - 1 - lane; // this lane's predecessor in the current window
      while (true)
      {
        unsigned long long w = t >= 0 ? st[t] : (2ull << 62); // before tile 0: an inclusive prefix of 0
        while (__any_sync(0xffffffffu, (w >> 62) == 0))
          w = t >= 0 ? st[t] : (2ull << 62);
        const unsigned inc = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        const int first = inc ? __ffs(inc) - 1 : 32; // nearest predecessor with an inclusive prefix
        long long x = lane <= first ? static_cast<long long>(w & ST_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          x += __shfl_xor_sync(0xffffffffu, x, o);
        prefix += x;
        if (inc)
          break;
        t -= 32;
      }
      if (lane == 0)
      {
        __threadfence();
        st[tile] = ST_INC | static_cast<unsigned long long>(prefix + agg);
      }
    }
    if (lane == 0)
    {
      s_prefix = prefix;
      if (tile == ntiles - 1)
      {
        out[n] = prefix + agg;
        *total = prefix + agg;
      }
    }
  }
  __syncthreads();
  long long run = s_prefix + s_w[wid] + incl - sum;
#pragma unroll
  for (int k = 0; k < LI; ++k)
  {
    if (base + k < n)
      out[base + k] = run;
    run += v[k];
  }
}
} // namespace

void scan_block_counts(cfx_ctx* c, const int32_t* counts, int64_t nblocks, int64_t* offsets, int64_t* total_out,
                       int64_t cap)
{
  c->scratch64.reserve(c->pool, 64);
  CFX_LAUNCH(c, scan_small_kernel<int32_t>, 1, SB, 0, counts, nblocks, offsets, total_out ? total_out : c->scratch64.p,
             cap, c->err_flag.p);
}

template <class Tin>
static void exclusive_scan_impl(cfx_ctx* c, const Tin* in, DN nn, int64_t* out)
{
  c->scratch64.reserve(c->pool, 64);
  const int64_t n = nn.h;
  if (n == 0)
  {
    CFX_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), c->stream));
    CFX_CUDA(cudaMemsetAsync(c->scratch64.p, 0, sizeof(int64_t), c->stream));
    return;
  }
  // A/B on B200 (256^3, 17 M row counts): the three-pass scan (sum, scan of sums, scan) takes 0.10 ms, the single
  // pass 0.13 ms -- the one-pass kernel is for the many SMALL scans of a step (rule offsets, list offsets), where
  // it replaces three dependent launches by one
  static const bool three_pass = getenv("CFX_SCAN_3PASS") != nullptr; // A/B switch
  if (!three_pass && n <= (int64_t(1) << 22))
  { // one pass over the data: tile aggregates published through a status word, decoupled look-back
    const int64_t ntiles = (n + LTILE - 1) / LTILE;
    DevBuf<unsigned long long> status; // [0] ticket counter, [1 + t] status of tile t
    status.reserve(c->pool, static_cast<size_t>(ntiles) + 2);
    CFX_CUDA(cudaMemsetAsync(status.p, 0, (static_cast<size_t>(ntiles) + 2) * sizeof(unsigned long long), c->stream));
    CFX_LAUNCH(c, lookback_scan_kernel<Tin>, static_cast<unsigned>(ntiles), LB, 0, in, nn, status.p, c->scratch64.p, out);
    status.release();
    return;
  }
  const int64_t ntiles = (n + TILE - 1) / TILE;
  DevBuf<int64_t> sums, offs;
  sums.reserve(c->pool, ntiles);
  offs.reserve(c->pool, ntiles);
  CFX_LAUNCH(c, tile_sum_kernel<Tin>, grid_for(n, TILE), TB, 0, in, nn, sums.p);
  CFX_LAUNCH(c, scan_small_kernel<int64_t>, 1, SB, 0, sums.p, ntiles, offs.p, c->scratch64.p, int64_t(-1),
             c->err_flag.p);
  CFX_LAUNCH(c, tile_scan_kernel<Tin>, grid_for(n, TILE), TB, 0, in, nn, offs.p, c->scratch64.p, out);
  sums.release(); // stream-ordered reuse is safe: the pool hands memory back to this stream only
  offs.release();
}

void exclusive_scan_i32_to_i64(cfx_ctx* c, const int32_t* in, int64_t n, int64_t* out)
{
  exclusive_scan_impl<int32_t>(c, in, dn_exact(n), out);
}
void exclusive_scan_i64(cfx_ctx* c, const int64_t* in, int64_t n, int64_t* out)
{
  exclusive_scan_impl<int64_t>(c, in, dn_exact(n), out);
}
void exclusive_scan_i64(cfx_ctx* c, const int64_t* in, DN n, int64_t* out) { exclusive_scan_impl<int64_t>(c, in, n, out); }
} // namespace cfx
