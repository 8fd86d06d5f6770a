// exchange.cu -- pack / unpack-add kernels of the ghost exchange (one rank per GPU).
//
// The role of la::MatrixCSR::scatter_rev and la::Vector::scatter_rev(add) (DOLFINx 0.11; called by
// the user after assembly, python/demo/demo_poisson.py:52,54): ghost-row values and ghost vector
// entries are packed into a send buffer, exchanged with the owning ranks (NCCL send/recv on these
// device buffers, cutfemx_b200/parallel.py) and added into the owner's entries.
// Roofline: HBM, 20 B per exchanged value (index + value in, value out); the exchanged volume is one
// mesh plane per neighbour (~8 MB at 256^3 on 8 ranks), so these kernels are
// launch-latency sized.
#include "common.cuh"

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
