// Dirichlet conditions on the assembled system.
//
// Reference: assemble_matrix with bcs zeroes the Dirichlet rows / columns of every element tensor before mat_set
// (assemble_matrix_impl.h:146-185 cells, :537-603 interior facets; assembler.h:643-683), set_diagonal then writes the
// diagonal of the owned Dirichlet rows (assembler.h:745-787), apply_lifting re-runs the matrix kernels in LiftingMode
// and subtracts alpha * Ae[:, bc] (x_bc - x0) from b (assemble_vector_impl.h:383-564), and DirichletBC::set writes
// alpha * (x_bc - x0) into the constrained entries.
//
// Here the row owners assemble the UNCONSTRAINED matrix (assemble.cu) and one pass over its CSR rows applies the
// conditions: entry (r, c) is an element-tensor sum, so dropping the contributions of Dirichlet rows / columns of
// every element tensor equals zeroing the assembled entry, and the lifting sum over elements equals the product of
// the assembled Dirichlet columns with alpha (x_bc - x0).  One warp per block row, fixed shuffle tree -> the result
// does not depend on scheduling.  Accumulating assembly (zero_first == 0) keeps the reference semantics (previous
// values stay, masked contributions add nothing) by assembling into the zeroed matrix and adding the saved values back.
#include "common.cuh"

namespace cfx
{
namespace
{
constexpr int DW = 8; // rows (warps) per block

// mask != 0: zero entry (r*bs+a, c*bs+k) when bc0[r*bs+a] or bc1[c*bs+k]   (either marker array may be null)
// lift != 0: b[r*bs+a] -= sum_{c,k : bc1[c*bs+k]} A[r*bs+a, c*bs+k] * alpha * (g[c*bs+k] - x0[c*bs+k])
template <int BS>
__global__ void __launch_bounds__(DW * 32)
    dirichlet_rows_kernel(int64_t n_rows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                          double* __restrict__ vals, const int8_t* __restrict__ bc0, const int8_t* __restrict__ bc1,
                          int mask, int lift, const double* __restrict__ g, const double* __restrict__ x0, double alpha,
                          double* __restrict__ b)
{
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * DW + (threadIdx.x >> 5);
  if (r >= n_rows)
    return;
  const int64_t pb = row_ptr[r], pe = row_ptr[r + 1];
  bool row_bc[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
    row_bc[a] = bc0 && bc0[r * BS + a];
  double acc[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
    acc[a] = 0.0;
  for (int64_t p0 = pb; p0 < pe; p0 += 32)
  { // ascending columns, 32 at a time; per chunk a fixed tree, chunks added in order
    const int64_t p = p0 + lane;
    double part[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a)
      part[a] = 0.0;
    if (p < pe)
    {
      const int64_t c = cols[p];
      double* blk = vals + p * BS * BS;
#pragma unroll
      for (int k = 0; k < BS; ++k)
      {
        const bool col_bc = bc1 && bc1[c * BS + k];
        double gv = 0.0;
        if (lift && col_bc)
          gv = alpha * (g[c * BS + k] - (x0 ? x0[c * BS + k] : 0.0));
#pragma unroll
        for (int a = 0; a < BS; ++a)
        {
          if (lift && col_bc)
            part[a] += blk[a * BS + k] * gv;
          if (mask && (col_bc || row_bc[a]))
            blk[a * BS + k] = 0.0;
        }
      }
    }
    if (lift)
    {
#pragma unroll
      for (int a = 0; a < BS; ++a)
      {
        double v = part[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          v += __shfl_down_sync(0xffffffffu, v, o);
        acc[a] += v;
      }
    }
  }
  if (lift && lane == 0)
  {
#pragma unroll
    for (int a = 0; a < BS; ++a)
      if (acc[a] != 0.0)
        b[r * BS + a] -= acc[a];
  }
}

__global__ void add_values_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n)
    dst[i] += src[i];
}

// rows are blocked dof indices bs*dof + k (DirichletBC::dof_indices unrolled)
__global__ void set_diagonal_blocked_kernel(const int32_t* __restrict__ rows, int64_t n, int64_t n_rows, int bs,
                                            const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                            double* __restrict__ vals, double diagonal, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n)
    return;
  const int32_t rb = rows[i];
  const int32_t r = rb / bs, k = rb % bs;
  if (rb < 0 || r >= n_rows)
  {
    err[0] = 29;
    err[1] = rb;
    return;
  }
  int64_t lo = row_ptr[r], hi = row_ptr[r + 1];
  const int64_t end = hi;
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < r)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo >= end || cols[lo] != r)
  {
    err[0] = 30;
    err[1] = rb;
    return;
  }
  vals[lo * bs * bs + k * bs + k] = diagonal;
}

__global__ void set_bc_kernel(const int32_t* __restrict__ dofs, int64_t n, int64_t n_total,
                              const double* __restrict__ g, const double* __restrict__ x0, double alpha,
                              double* __restrict__ b, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n)
    return;
  const int32_t d = dofs[i];
  if (d < 0 || d >= n_total)
  {
    err[0] = 29;
    err[1] = d;
    return;
  }
  b[d] = alpha * (g[d] - (x0 ? x0[d] : 0.0));
}

__global__ void set_bc_marked_kernel(const int8_t* __restrict__ markers, int64_t n, const double* __restrict__ g,
                                     const double* __restrict__ x0, double alpha, double* __restrict__ b)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n && markers[i])
    b[i] = alpha * (g[i] - (x0 ? x0[i] : 0.0));
}

void launch_rows(cfx_ctx* ctx, cfx_pattern* A, const int8_t* bc0, const int8_t* bc1, int mask, int lift,
                 const double* g, const double* x0, double alpha, double* b)
{
  if (A->n_rows == 0)
    return;
  const unsigned grid = grid_for(A->n_rows, DW);
  switch (A->bs)
  {
  case 1:
    CFX_LAUNCH(ctx, dirichlet_rows_kernel<1>, grid, DW * 32, 0, A->n_rows, A->row_ptr.p, A->cols.p, A->values.p, bc0,
               bc1, mask, lift, g, x0, alpha, b);
    break;
  case 2:
    CFX_LAUNCH(ctx, dirichlet_rows_kernel<2>, grid, DW * 32, 0, A->n_rows, A->row_ptr.p, A->cols.p, A->values.p, bc0,
               bc1, mask, lift, g, x0, alpha, b);
    break;
  case 3:
    CFX_LAUNCH(ctx, dirichlet_rows_kernel<3>, grid, DW * 32, 0, A->n_rows, A->row_ptr.p, A->cols.p, A->values.p, bc0,
               bc1, mask, lift, g, x0, alpha, b);
    break;
  default:
    throw Error(CFX_ERR_UNSUPPORTED, "Dirichlet conditions: block size must be 1, 2 or 3");
  }
}
} // namespace
} // namespace cfx

using namespace cfx;

extern "C"
{
cfx_status cfx_assemble_matrix_bc(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, int zero_first, double diag_inactive,
                                  const int8_t* bc_markers0, const int8_t* bc_markers1, int memspace_bc,
                                  double* values_out, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && a && A, CFX_ERR_INVALID, "cfx_assemble_matrix_bc: NULL argument");
  resolve_form(ctx, const_cast<cfx_form*>(a)); // not part of the deferred-size step: sizes on the host first
  resolve(ctx, A);
  const Space& S = ctx->spaces[A->space];
  const size_t nb = static_cast<size_t>(S.n_total) * S.bs, nv = static_cast<size_t>(A->nnz) * A->bs * A->bs;
  DevBuf<int8_t> own0, own1;
  DevBuf<double> prior;
  const int8_t* d0 = bc_markers0 ? adopt(ctx, own0, bc_markers0, nb, memspace_bc) : nullptr;
  const int8_t* d1 = bc_markers1 ? adopt(ctx, own1, bc_markers1, nb, memspace_bc) : nullptr;
  const bool any = d0 || d1;
  if (any && !zero_first && nv > 0)
  { // keep what the matrix holds: the masked contributions must add nothing to it
    prior.reserve(ctx->pool, nv);
    CFX_CUDA(cudaMemcpyAsync(prior.p, A->values.p, nv * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  cfx_status rc = cfx_assemble_matrix(ctx, a, A, any ? 1 : zero_first, diag_inactive, nullptr, CFX_DEVICE);
  if (rc != CFX_OK)
    return rc;
  if (any)
  {
    StageScope st(ctx, "dirichlet_rows", 12.0 * static_cast<double>(A->nnz));
    launch_rows(ctx, A, d0, d1, 1, 0, nullptr, nullptr, 0.0, nullptr);
    if (prior.p)
      CFX_LAUNCH(ctx, add_values_kernel, grid_for(static_cast<int64_t>(nv), 256), 256, 0, A->values.p, prior.p,
                 static_cast<int64_t>(nv));
  }
  if (values_out)
    export_to(ctx, values_out, A->values.p, nv, memspace);
  else
    CFX_CUDA(cudaStreamSynchronize(ctx->stream)); // the marker copies must not outlive this call
  own0.release();
  own1.release();
  prior.release();
  CFX_API_END(ctx)
}

cfx_status cfx_set_diagonal(cfx_ctx* ctx, cfx_pattern* A, const int32_t* rows, int64_t n, double diagonal, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && A && (n == 0 || rows), CFX_ERR_INVALID, "cfx_set_diagonal: NULL argument");
  resolve(ctx, A);
  settle_values(ctx, A);
  A->values_zero = false;
  if (n > 0)
  {
    DevBuf<int32_t> own;
    const int32_t* d = adopt(ctx, own, rows, static_cast<size_t>(n), memspace);
    CFX_LAUNCH(ctx, set_diagonal_blocked_kernel, grid_for(n, 256), 256, 0, d, n, A->n_rows, A->bs, A->row_ptr.p,
               A->cols.p, A->values.p, diagonal, ctx->err_flag.p);
    check_device_error(ctx, "cfx_set_diagonal (row out of range or without a diagonal entry)");
    own.release();
  }
  CFX_API_END(ctx)
}

cfx_status cfx_apply_lifting(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, double* b, const double* bc_values1,
                             const int8_t* bc_markers1, const double* x0, double alpha, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && a && A && b && bc_values1 && bc_markers1, CFX_ERR_INVALID, "cfx_apply_lifting: NULL argument");
  resolve_form(ctx, const_cast<cfx_form*>(a));
  resolve(ctx, A);
  const Space& S = ctx->spaces[A->space];
  const size_t nb = static_cast<size_t>(S.n_total) * S.bs, nv = static_cast<size_t>(A->nnz) * A->bs * A->bs;
  DevBuf<int8_t> ownm;
  DevBuf<double> owng, ownx, ownb, prior;
  const int8_t* dm = adopt(ctx, ownm, bc_markers1, nb, memspace);
  const double* dg = adopt(ctx, owng, bc_values1, nb, memspace);
  const double* dx = x0 ? adopt(ctx, ownx, x0, nb, memspace) : nullptr;
  double* db = b;
  if (memspace == CFX_HOST)
  {
    ownb.reserve(ctx->pool, nb);
    CFX_CUDA(cudaMemcpyAsync(ownb.p, b, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    db = ownb.p;
  }
  // the Dirichlet columns of the unconstrained matrix: assembled into A's storage, whose values are put back after
  if (nv > 0)
  {
    prior.reserve(ctx->pool, nv);
    CFX_CUDA(cudaMemcpyAsync(prior.p, A->values.p, nv * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  cfx_status rc = cfx_assemble_matrix(ctx, a, A, 1, 0.0, nullptr, CFX_DEVICE);
  if (rc == CFX_OK)
  {
    StageScope st(ctx, "dirichlet_lifting", 12.0 * static_cast<double>(A->nnz));
    launch_rows(ctx, A, nullptr, dm, 0, 1, dg, dx, alpha, db);
  }
  if (nv > 0)
    CFX_CUDA(cudaMemcpyAsync(A->values.p, prior.p, nv * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (rc != CFX_OK)
    return rc;
  if (memspace == CFX_HOST)
    export_to(ctx, b, ownb.p, nb, CFX_HOST);
  else
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  ownm.release();
  owng.release();
  ownx.release();
  ownb.release();
  prior.release();
  CFX_API_END(ctx)
}

// assemble_matrix(a, bcs) + assemble_vector(L) + apply_lifting + set_bc in ONE assembly (the flow of
// demo_elasticity.py:78-84): the unconstrained system is assembled once, one pass over its rows lifts the
// Dirichlet columns into b and zeroes the Dirichlet rows and columns, then the diagonal and the constrained
// entries of b are set.  A and b are overwritten (zero_first semantics); all arrays are DEVICE arrays.
cfx_status cfx_assemble_system_bc(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, const cfx_form* L, double* b,
                                  const int8_t* bc_markers, const double* bc_values, const double* x0, double alpha,
                                  const int32_t* bc_rows_owned, int64_t n_bc_rows, double diagonal)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && a && A && L && b && bc_markers && bc_values && (n_bc_rows == 0 || bc_rows_owned), CFX_ERR_INVALID,
              "cfx_assemble_system_bc: NULL argument");
  const Space& S = ctx->spaces[A->space];
  resolve_form(ctx, const_cast<cfx_form*>(a));
  resolve_form(ctx, const_cast<cfx_form*>(L));
  resolve(ctx, A);
  cfx_status rc = cfx_assemble_system(ctx, a, A, 1, 0.0, L, b, 1);
  if (rc != CFX_OK)
    return rc;
  {
    StageScope st(ctx, "dirichlet_rows", 12.0 * static_cast<double>(A->nnz));
    launch_rows(ctx, A, bc_markers, bc_markers, 1, 1, bc_values, x0, alpha, b);
  }
  if (n_bc_rows > 0)
    CFX_LAUNCH(ctx, set_diagonal_blocked_kernel, grid_for(n_bc_rows, 256), 256, 0, bc_rows_owned, n_bc_rows, A->n_rows,
               A->bs, A->row_ptr.p, A->cols.p, A->values.p, diagonal, ctx->err_flag.p);
  // every constrained entry of b this rank holds (owned AND ghost: a rank may hold only ghost copies of the
  // constrained dofs, so this does not depend on n_bc_rows): the marker array is the list
  CFX_LAUNCH(ctx, set_bc_marked_kernel, grid_for(static_cast<int64_t>(S.n_total) * S.bs, 256), 256, 0, bc_markers,
             static_cast<int64_t>(S.n_total) * S.bs, bc_values, x0, alpha, b);
  if (n_bc_rows > 0)
    check_device_error(ctx, "cfx_assemble_system_bc (Dirichlet row out of range or without a diagonal entry)");
  CFX_API_END(ctx)
}

cfx_status cfx_set_bc(cfx_ctx* ctx, double* b, int64_t n_total, const int32_t* dofs, int64_t n, const double* bc_values,
                      const double* x0, double alpha, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && b && (n == 0 || (dofs && bc_values)), CFX_ERR_INVALID, "cfx_set_bc: NULL argument");
  if (n > 0)
  {
    const size_t nb = static_cast<size_t>(n_total);
    DevBuf<int32_t> ownd;
    DevBuf<double> owng, ownx, ownb;
    const int32_t* dd = adopt(ctx, ownd, dofs, static_cast<size_t>(n), memspace);
    const double* dg = adopt(ctx, owng, bc_values, nb, memspace);
    const double* dx = x0 ? adopt(ctx, ownx, x0, nb, memspace) : nullptr;
    double* db = b;
    if (memspace == CFX_HOST)
    {
      ownb.reserve(ctx->pool, nb);
      CFX_CUDA(cudaMemcpyAsync(ownb.p, b, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      db = ownb.p;
    }
    CFX_LAUNCH(ctx, set_bc_kernel, grid_for(n, 256), 256, 0, dd, n, n_total, dg, dx, alpha, db, ctx->err_flag.p);
    if (memspace == CFX_HOST)
      export_to(ctx, b, ownb.p, nb, CFX_HOST);
    check_device_error(ctx, "cfx_set_bc (dof out of range)");
    ownd.release();
    owng.release();
    ownx.release();
    ownb.release();
  }
  CFX_API_END(ctx)
}
} // extern "C"
