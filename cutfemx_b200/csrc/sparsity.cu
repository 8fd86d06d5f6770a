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
#include <algorithm>

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

// largest incidence-list length (device max, exact)
__global__ void max_degree_kernel(const int64_t* __restrict__ inc_ptr, int64_t n_dofs, int* __restrict__ out)
{
  const int64_t d = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  int v = d < n_dofs ? static_cast<int>(inc_ptr[d + 1] - inc_ptr[d]) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = max(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0)
    atomicMax(out, v);
}

// fperm[k] for incidence k = (dof d, cell c): bits 0..3 local index of d in c, bits 4+4j.. rank of
// the cell's j-th dof among the cell's dofs (ascending).  Static per space, nd <= 6.
__global__ void fperm_kernel(int64_t n_dofs, const int64_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_cell,
                             const int32_t* __restrict__ dofmap, int nd, uint32_t* __restrict__ fperm)
{
  const int64_t d = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (d >= n_dofs)
    return;
  const int64_t b = inc_ptr[d], e = inc_ptr[d + 1];
  for (int64_t k = b; k < e; ++k)
  {
    const int64_t c = inc_cell[k];
    int32_t dd[6];
    for (int i = 0; i < nd; ++i)
      dd[i] = dofmap[c * nd + i];
    uint32_t v = 0;
    for (int i = 0; i < nd; ++i)
    {
      if (dd[i] == d)
        v |= static_cast<uint32_t>(i);
      uint32_t rk = 0;
      for (int j = 0; j < nd; ++j)
        rk += dd[j] < dd[i] ? 1u : 0u;
      v |= rk << (4 + 4 * i);
    }
    fperm[k] = v;
  }
}

// MC listed cells per thread (cells[i * stride]), warp-strided: set `bit` in each cell's flag byte and store
// `rowval` in the row flag of each of its dofs.  All writers of one launch store the same values, so the plain
// byte stores are race-free in effect (idempotent).  The kernel is a chain list -> (flag byte, dofmap row) ->
// stores and nothing else: the loads of the MC cells are batched level by level.
constexpr int MC = 4;
__global__ void mark_cells_kernel(const int32_t* __restrict__ cells, DN n_, int stride, int64_t limit, uint8_t bit,
                                  const int32_t* __restrict__ dofmap, int nd, uint8_t rowval,
                                  uint8_t* __restrict__ cell_flags, uint8_t* __restrict__ row_flag,
                                  int32_t* __restrict__ err)
{
  constexpr int NDMAX = 10;
  const int64_t n = n_.get();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x) >> 5;
  const int64_t i0 = warp * (32 * MC) + lane;
  int32_t c[MC];
#pragma unroll
  for (int q = 0; q < MC; ++q)
  {
    const int64_t i = i0 + 32 * q;
    c[q] = i < n ? cells[i * stride] : -1;
    if (i < n && (c[q] < 0 || c[q] >= limit))
    {
      err[0] = 22;
      err[1] = c[q];
      c[q] = -1;
    }
  }
  uint8_t f[MC];
  int32_t d[MC][NDMAX];
#pragma unroll
  for (int q = 0; q < MC; ++q)
  {
    f[q] = c[q] >= 0 ? cell_flags[c[q]] : uint8_t(0);
#pragma unroll
    for (int j = 0; j < NDMAX; ++j)
      d[q][j] = (c[q] >= 0 && j < nd) ? dofmap[static_cast<int64_t>(c[q]) * nd + j] : -1;
  }
#pragma unroll
  for (int q = 0; q < MC; ++q)
  {
    if (c[q] < 0)
      continue;
    cell_flags[c[q]] = f[q] | bit;
#pragma unroll
    for (int j = 0; j < NDMAX; ++j)
      if (d[q][j] >= 0)
        row_flag[d[q][j]] = rowval;
  }
}

// rows of inserted pattern entries: active (bit0) and generic (bit1)
__global__ void mark_rows_kernel(const int32_t* __restrict__ rows, DN n_, int64_t limit,
                                 uint8_t* __restrict__ row_flag, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n_.get())
    return;
  const int32_t r = rows[i];
  if (r < 0 || r >= limit || (i > 0 && rows[i - 1] > r))
  {
    err[0] = 25;
    err[1] = r;
    return;
  }
  row_flag[r] = 3;
}

__global__ void xslot_set_kernel(const int32_t* __restrict__ rows, DN n_, int32_t* __restrict__ xslot, bool clear)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n_.get())
    return;
  if (i == 0 || rows[i - 1] != rows[i])
    xslot[rows[i]] = clear ? -1 : static_cast<int32_t>(i);
}

// number of entries of the ascending list `a` that are < bound (single thread, binary search)
__global__ void lower_bound_kernel(const int32_t* __restrict__ a, int64_t n, int64_t bound, int64_t* __restrict__ out)
{
  int64_t lo = 0, hi = n;
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < bound)
      lo = mid + 1;
    else
      hi = mid;
  }
  *out = lo;
}

__global__ void facet_slot_set_kernel(const int32_t* __restrict__ rows4, DN n_, int nf,
                                      const int32_t* __restrict__ c2f, int32_t* __restrict__ facet_slot, bool clear)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n_.get())
    return;
  const int32_t f = c2f[static_cast<int64_t>(rows4[4 * i]) * nf + rows4[4 * i + 1]];
  facet_slot[f] = clear ? -1 : static_cast<int32_t>(i);
}

// slots of act_rows whose row has flag bit1 (compact.cuh predicate: 16 consecutive slots per call)
struct BandSlotPred
{
  const int32_t* act_rows;
  const uint8_t* row_flag;
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
#pragma unroll 4
    for (int k = 0; k < 16; ++k)
      if (base + k < n && (row_flag[act_rows[base + k]] & 2))
        m |= 1u << k;
    return m;
  }
};

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
  const int32_t* xslot; // per row: first inserted entry (or -1); null if the form has none
  const int32_t* xrows;
  const int32_t* xcols;
  DN n_x;
};

constexpr int RW = 4;      // rows (warps) per block
// partner-dof candidates per row (facet macro cliques): 128 * ND
constexpr int IMAX = 0x7fffffff;

// inactive rows: only the deactivation diagonal (assembler.h:538-560)
__global__ void pattern_inactive_count_kernel(const uint8_t* __restrict__ row_flag, int64_t n_rows, int diag,
                                              int32_t* __restrict__ row_nnz)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (r < n_rows && !row_flag[r])
    row_nnz[r] = diag;
}

// (vals != null: the row's one bs x bs block is zeroed here -- the active rows' values are left to the assembly)
__global__ void pattern_inactive_fill_kernel(const uint8_t* __restrict__ row_flag, int64_t n_rows, int diag,
                                             const int64_t* __restrict__ row_ptr, int32_t* __restrict__ cols,
                                             int64_t cap /* entries `cols` can hold */, double* __restrict__ vals,
                                             int bs2)
{
  const int64_t r = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (r < n_rows && !row_flag[r] && diag && row_ptr[r] < cap)
  {
    const int64_t p = row_ptr[r];
    cols[p] = static_cast<int32_t>(r);
    if (vals)
      for (int q = 0; q < bs2; ++q)
        vals[p * bs2 + q] = 0.0;
  }
}

// deferred-size mode: the pattern of this step must fit the arrays of the matrix object that is being reused.  If
// it does not, raise the error flag and empty the active-row list, so that no kernel after this one (pattern fill,
// assembly) touches the too-small arrays.
// It also verifies the two launch decisions a deferred step takes from earlier eager steps (counters = the
// pattern pass's scratch: [1] rows left to the generic kernels, [2] rows with a contribution list).
__global__ void pattern_capacity_kernel(const int64_t* __restrict__ d_nnz, int64_t cap,
                                        const int64_t* __restrict__ counters, int expect_no_slow, int expect_no_noclist,
                                        int64_t* __restrict__ d_counts, int32_t* __restrict__ err)
{
  if (*d_nnz > cap)
  {
    err[0] = 33;
    err[1] = static_cast<int32_t>(*d_nnz > 0x7fffffffLL ? 0x7fffffffLL : *d_nnz);
    d_counts[0] = 0;
    d_counts[1] = 0;
    return;
  }
  if (expect_no_slow && counters[1] != 0)
  {
    err[0] = 34;
    err[1] = static_cast<int32_t>(counters[1]);
  }
  if (expect_no_noclist && d_counts[0] - d_counts[1] - counters[2] != 0)
  {
    err[0] = 35;
    err[1] = static_cast<int32_t>(d_counts[0] - d_counts[1] - counters[2]);
  }
}

// ascending sort of a small register array (odd-even transposition network, fully unrolled)
template <int N>
__device__ __forceinline__ void sort_small(int32_t (&v)[N])
{
#pragma unroll
  for (int pass = 0; pass < N; ++pass)
#pragma unroll
    for (int i = pass & 1; i + 1 < N; i += 2)
    {
      const int32_t lo = min(v[i], v[i + 1]), hi = max(v[i], v[i + 1]);
      v[i] = lo;
      v[i + 1] = hi;
    }
}

// One WARP per row.  Lanes take the row's incident cells (coalesced inc_cell read, ND-wide dofmap
// gathers in parallel), sort their cell's dofs, and the warp extracts the sorted unique column set
// by repeated warp-wide minimum over the list heads (one REDUX per column).
//
// FILL = false (first pass): counts the columns and, for rows with at most 32 columns and 32
// incident cells ("fast rows"), also stores
//   tmp[idx*ts + k]        = k-th column                 (copied to the CSR after the scan)
//   mask[idx*stride + l]   = bit mask of the CSR positions of the dofs of incident cell l
// so that the assembly gather (assemble.cu) needs neither the dofmap nor a column search.  ts = 32 for scalar
// spaces (the mask gather handles rows of at most 32 columns); blocked spaces, whose gather searches the columns and
// reads no masks, pass ts = 96 so that the 30 - 65 column rows of P2 tetrahedra are not merged a second time by the
// fill pass (columns are flushed to tmp 32 at a time).
// FILL = true (second pass, slow rows only): writes the columns straight into the CSR.
//
// Modes: act_rows == nullptr walks ALL rows with every cell active (static full-mesh structure,
// masks stored per global incidence); otherwise the active rows of a prepared form, optionally
// only those with a facet-band cell (only_band: the others take pattern_static_kernel).
template <int ND, bool FILL>
__global__ void __launch_bounds__(RW * 32)
    pattern_rows_kernel(RowCtx rc, const int32_t* __restrict__ act_rows, const int32_t* __restrict__ slots,
                        DN n_rows_in, int only_band, int stride, int ts, int32_t* __restrict__ row_nnz,
                        const int64_t* __restrict__ row_ptr,
                        int32_t* __restrict__ cols_out, int32_t* __restrict__ tmp, uint32_t* __restrict__ mask_out,
                        uint8_t* __restrict__ row_fast, unsigned long long* __restrict__ n_slow,
                        int32_t* __restrict__ err)
{
  constexpr int XCAP = 128 * ND; // partner-dof candidates per row (facet macro cliques)
  __shared__ int32_t s_extra[RW][XCAP];
  __shared__ int s_nextra[RW];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = static_cast<int64_t>(blockIdx.x) * RW + w;
  if (it >= n_rows_in.get())
    return;
  // `slots` (optional): the band rows only, so that no warp is launched just to find out it has no work
  const int64_t idx = slots ? slots[it] : it;
  const bool all_mode = act_rows == nullptr;
  const int64_t r = all_mode ? idx : act_rows[idx];
  if (only_band && !(rc.row_flag[r] & 2))
    return; // non-band rows of a mesh with a static structure: pattern_static_kernel
  if constexpr (FILL)
  {
    if (row_fast[idx] & 1)
      return; // already copied from tmp
  }
  const unsigned full = 0xffffffffu;
  const int64_t ib = rc.inc_ptr[r];
  const int n_inc = static_cast<int>(rc.inc_ptr[r + 1] - ib);
  if (lane == 0)
    s_nextra[w] = 0;
  __syncwarp();
  int32_t h[ND]; // this lane's sorted candidate list (first 32 incidences live in registers)
#pragma unroll
  for (int j = 0; j < ND; ++j)
    h[j] = IMAX;
  bool band = false;
  // P1, count pass: a band cell brings at most ONE new column (the dof opposite the facet opposite the row's
  // dof), kept in a register and merged like a fifth list entry instead of going through shared memory
  int32_t xreg = IMAX;
  for (int k0 = 0; k0 < n_inc; k0 += 32)
  {
    const int k = k0 + lane;
    if (k < n_inc)
    {
      const int64_t c = rc.inc_cell[ib + k];
      const uint8_t fl = all_mode ? uint8_t(1) : rc.cell_flags[c];
      bool own_needed = (fl & 0xFD) != 0; // has a cell tensor (materialised or on the fly)
      if (fl & 2)
      {
        // facet macro cliques: the partner cell's dofs that this cell does not have.  All NF
        // facets are probed with independent loads (c2f row -> slots -> partner cells -> their
        // dofmap rows in flight together) instead of one dependent chain per facet.
        band = true;
        constexpr int NF = (ND == 3 || ND == 6) ? 3 : 4;
        int64_t fct[NF];
        int32_t slot[NF], oc[NF], d[ND];
#pragma unroll
        for (int lf = 0; lf < NF; ++lf)
          fct[lf] = rc.c2f[c * NF + lf];
#pragma unroll
        for (int j = 0; j < ND; ++j)
          d[j] = rc.dofmap[c * ND + j];
        // a band cell has a band facet by construction, so its own dofs are needed whatever the probes find
        own_needed = true;
        // P1: local facet lf is opposite local vertex lf, so the only facet that does not touch the row's dof is
        // the one with the dof's own local index -- one probe instead of NF (the others could only bring
        // dofs of cells that are themselves incident to the row, see below)
        int li = -1;
        if constexpr (ND == NF)
        {
#pragma unroll
          for (int j = 0; j < ND; ++j)
            li = (d[j] == r) ? j : li;
        }
#pragma unroll
        for (int lf = 0; lf < NF; ++lf)
          slot[lf] = (ND != NF || lf == li) ? rc.facet_slot[fct[lf]] : -1;
#pragma unroll
        for (int lf = 0; lf < NF; ++lf)
        {
          oc[lf] = -1;
          if (slot[lf] >= 0)
          {
            const int32_t c0 = rc.f2c2[2 * fct[lf]], c1 = rc.f2c2[2 * fct[lf] + 1];
            oc[lf] = (c0 == c) ? c1 : c0;
          }
        }
#pragma unroll
        for (int lf = 0; lf < NF; ++lf)
        {
          if (oc[lf] < 0)
            continue;
          own_needed = true;
          int32_t od[ND];
          bool has_r = false;
#pragma unroll
          for (int j = 0; j < ND; ++j)
          {
            od[j] = rc.dofmap[static_cast<int64_t>(oc[lf]) * ND + j];
            has_r = has_r || (od[j] == r);
          }
          // a partner cell that holds this row's dof is itself an incident cell of the row (and needed, it
          // has the same band facet): its dofs enter through its own lane.  Only facets that do not touch
          // the row's dof bring new columns.
          if (has_r)
            continue;
#pragma unroll
          for (int j = 0; j < ND; ++j)
          {
            bool dup = false;
#pragma unroll
            for (int jj = 0; jj < ND; ++jj)
              dup = dup || (od[j] == d[jj]);
            if (!dup)
            {
              if (ND == NF && !FILL && k0 == 0)
                xreg = od[j];
              else
              {
                const int pos = atomicAdd(&s_nextra[w], 1);
                if (pos < XCAP)
                  s_extra[w][pos] = od[j];
              }
            }
          }
        }
      }
      if (own_needed)
      {
        if (k0 == 0)
        {
#pragma unroll
          for (int j = 0; j < ND; ++j)
            h[j] = rc.dofmap[c * ND + j];
        }
        else
        { // rare: more than 32 incident cells -> spill this cell's dofs to the shared candidate list
          const int pos = atomicAdd(&s_nextra[w], ND);
          if (pos + ND <= XCAP)
          {
#pragma unroll
            for (int j = 0; j < ND; ++j)
              s_extra[w][pos + j] = rc.dofmap[c * ND + j];
          }
        }
      }
    }
  }
  if (rc.xslot != nullptr && !all_mode)
  { // SparsityPattern::insert entries received from other ranks
    const int32_t xs = rc.xslot[r];
    if (xs >= 0)
      for (int64_t i = xs + lane, nx = rc.n_x.get(); i < nx && rc.xrows[i] == r; i += 32)
      {
        const int pos = atomicAdd(&s_nextra[w], 1);
        if (pos < XCAP)
          s_extra[w][pos] = rc.xcols[i];
      }
  }
  __syncwarp();
  if (s_nextra[w] != 0)
  { // the general merge below reads its extra candidates from shared memory only: move the register ones there
    __syncwarp();
    if (xreg != IMAX)
    {
      const int pos = atomicAdd(&s_nextra[w], 1);
      if (pos < XCAP)
        s_extra[w][pos] = xreg;
      xreg = IMAX;
    }
    __syncwarp();
  }
  const int n_extra = s_nextra[w];
  if (n_extra > XCAP)
  {
    if (lane == 0)
    {
      err[0] = 23;
      err[1] = static_cast<int32_t>(r);
    }
    return;
  }
  sort_small<ND>(h);
  // A row with at least one contributing cell finds its own dof among the cells' dofs; rows
  // without any (all_mode cannot have them, active rows neither) are the inactive-row kernels' job.
  const int64_t ob = FILL ? row_ptr[r] : 0;
  int32_t keep = 0;
  uint32_t M = 0;
  int count = 0;
  if (!FILL && n_extra == 0)
  { // common case: no facet partners, columns kept in lane registers (first 32)
    uint32_t bit = 1u;
    while (true)
    {
      const int32_t m = __reduce_min_sync(full, min(h[0], xreg));
      if (m == IMAX)
        break;
      const bool p = h[0] == m;
      M |= p ? bit : 0u;
#pragma unroll
      for (int j = 0; j + 1 < ND; ++j)
        h[j] = p ? h[j + 1] : h[j];
      h[ND - 1] = p ? IMAX : h[ND - 1];
      xreg = (xreg == m) ? IMAX : xreg;
      keep = (lane == (count & 31)) ? m : keep;
      if ((count & 31) == 31 && count < ts)
        tmp[idx * ts + (count & ~31) + lane] = keep; // a full group of 32 columns
      bit <<= 1;
      ++count;
    }
  }
  else
  {
    int32_t last = -1;
    while (true)
    {
      int32_t m = h[0];
      for (int e = lane; e < n_extra; e += 32)
      {
        const int32_t v = s_extra[w][e];
        m = (v > last && v < m) ? v : m;
      }
      m = __reduce_min_sync(full, m);
      if (m == IMAX)
        break;
      if (h[0] == m)
      { // pop: this cell's dof sits at CSR position `count` of the row
        M |= (count < 32) ? (1u << count) : 0u;
#pragma unroll
        for (int j = 0; j + 1 < ND; ++j)
          h[j] = h[j + 1];
        h[ND - 1] = IMAX;
      }
      if constexpr (FILL)
      {
        if (lane == (count & 31))
          keep = m;
        if ((count & 31) == 31)
          cols_out[ob + (count & ~31) + lane] = keep; // coalesced flush of 32 columns
      }
      else
      {
        if (lane == (count & 31))
          keep = m;
        if ((count & 31) == 31 && count < ts)
          tmp[idx * ts + (count & ~31) + lane] = keep;
      }
      last = m;
      ++count;
    }
  }
  if constexpr (FILL)
  {
    if ((count & 31) != 0 && lane < (count & 31))
      cols_out[ob + (count & ~31) + lane] = keep;
  }
  else
  {
    // ts == 32: rows the mask gather can take; ts > 32 (blocked spaces): rows whose columns fit tmp
    const bool fast = ts > 32 ? count <= ts : (count <= 32 && n_inc <= 32);
    const bool any_band = __any_sync(full, band);
    if (fast)
    {
      if (lane < (count & 31))
        tmp[idx * ts + (count & ~31) + lane] = keep; // the last, partial group
      if (ts == 32 && lane < n_inc)
        mask_out[(all_mode ? ib : idx * stride) + lane] = M;
    }
    if (lane == 0)
    {
      row_nnz[r] = count;
      if (row_fast)
        row_fast[idx] = (fast ? 1 : 0) | (any_band ? 2 : 0) | (slots ? 16 : 0);
      if (!fast)
        atomicAdd(n_slow, 1ULL);
    }
  }
}

// Band rows of a scalar P1 space WITH a static structure, a few threads per row (the structure of
// assemble.cu gather_matrix_band_p1_kernel; the warp-per-row pattern_rows_kernel above spends a whole warp and a
// repeated warp-wide minimum on rows that differ from their static row by a handful of columns).  The row's columns
// are its static columns of the needed cells (OR of the static position masks, like pattern_static_kernel) plus
//   * per band cell, the dof across the facet OPPOSITE the row's dof when that facet is a band facet (the partner
//     cell does not hold the row's dof, so that dof is the only new column the ghost-penalty macro element brings;
//     partner cell and its local facet index come from the facet's integration row),
//   * the entries other ranks inserted for this row (SparsityPattern::insert, multi-rank runs).
// Thread g of a group of BPG takes cells g, g + BPG, ...; extras meet in a small shared list, thread 0 sorts them,
// drops duplicates and static columns, and merges them with the kept static columns into tmp (copied to the CSR
// after the scan by pattern_copy_kernel).  Rows that end up with more than 32 columns are left to the generic fill
// pass (row_fast bit 0 clear).  No position masks are stored: the P1 band gather does not read them.
constexpr int BPG = 4;    // threads per band row
constexpr int BPB = 64;   // threads per block
constexpr int BPX = 64;   // extra candidates per row: <= 32 band cells + the entries other ranks inserted

template <int ND>
__global__ void __launch_bounds__(BPB)
    pattern_band_p1_kernel(RowCtx rc, const int32_t* __restrict__ act_rows, const int32_t* __restrict__ slots,
                           DN n_band_, const uint32_t* __restrict__ fmask, const uint32_t* __restrict__ fperm,
                           const int64_t* __restrict__ frow_ptr, const int32_t* __restrict__ fcols,
                           const int32_t* __restrict__ rows4, int32_t* __restrict__ row_nnz,
                           int32_t* __restrict__ tmp, uint8_t* __restrict__ row_fast,
                           unsigned long long* __restrict__ n_slow, int32_t* __restrict__ err)
{
  __shared__ int32_t s_ex[BPB / BPG][BPX];
  __shared__ int s_nex[BPB / BPG];
  const int tid = threadIdx.x, g = tid & (BPG - 1), lrow = tid / BPG;
  const unsigned gmask = ((1u << BPG) - 1u) << ((tid & 31) & ~(BPG - 1));
  const int64_t it = (static_cast<int64_t>(blockIdx.x) * BPB + tid) / BPG;
  if (it >= n_band_.get())
    return;
  const int64_t idx = slots[it];
  const int32_t r = act_rows[idx];
  const int64_t ib = rc.inc_ptr[r];
  const int n_inc = static_cast<int>(rc.inc_ptr[r + 1] - ib);
  if (g == 0)
    s_nex[lrow] = 0;
  __syncwarp(gmask);
  uint32_t R = 0;
  bool band = false;
  constexpr int U = 2;
  for (int l0 = g; l0 < n_inc; l0 += U * BPG)
  {
    int32_t c[U];
    uint32_t fm[U], fp[U];
    unsigned fl[U];
    bool in[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      in[u] = l0 + u * BPG < n_inc;
      const int l = in[u] ? l0 + u * BPG : 0;
      c[u] = rc.inc_cell[ib + l];
      fm[u] = fmask[ib + l];
      fp[u] = fperm[ib + l];
    }
    int32_t fct[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      fl[u] = in[u] ? rc.cell_flags[c[u]] : 0u;
      // the facet opposite the row's dof (P1: local facet li is opposite local vertex li); read whether or not the
      // cell turns out to be a band cell -- one level of the dependent chain less
      fct[u] = rc.c2f[static_cast<int64_t>(c[u]) * ND + (fp[u] & 15u)];
    }
    int32_t fs[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      fs[u] = (fl[u] & 2u) ? rc.facet_slot[fct[u]] : -1;
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (fl[u] & 0xFFu) // a cell with a tensor of its own, or a band cell (its dofs are needed for the facets)
        R |= ((fl[u] & 0xFDu) || (fl[u] & 2u)) ? fm[u] : 0u;
      band = band || (fl[u] & 2u);
      if (fs[u] >= 0)
      {
        const int4 rw = __ldg(reinterpret_cast<const int4*>(rows4) + fs[u]); // (cell0, lf0, cell1, lf1)
        const bool first = rw.x == c[u];
        const int64_t oc = first ? rw.z : rw.x;
        const int olf = first ? rw.w : rw.y;
        const int32_t x = rc.dofmap[oc * ND + olf];
        const int pos = atomicAdd(&s_nex[lrow], 1);
        if (pos < BPX)
          s_ex[lrow][pos] = x;
      }
    }
  }
  if (g == 0 && rc.xslot != nullptr)
  { // SparsityPattern::insert entries received from other ranks
    const int32_t xs = rc.xslot[r];
    if (xs >= 0)
      for (int64_t i = xs, nx = rc.n_x.get(); i < nx && rc.xrows[i] == r; ++i)
      {
        const int pos = atomicAdd(&s_nex[lrow], 1);
        if (pos < BPX)
          s_ex[lrow][pos] = rc.xcols[i];
      }
  }
#pragma unroll
  for (int o = 1; o < BPG; o <<= 1)
  {
    R |= __shfl_xor_sync(gmask, R, o);
    band = __shfl_xor_sync(gmask, band ? 1 : 0, o) || band;
  }
  __syncwarp(gmask);
  if (g != 0)
    return;
  int nex = s_nex[lrow];
  if (nex > BPX)
  {
    err[0] = 38; // more extra column candidates than the list holds
    err[1] = r;
    nex = BPX;
  }
  int32_t* ex = s_ex[lrow];
  const int64_t fb = frow_ptr[r];
  const int nfull = static_cast<int>(frow_ptr[r + 1] - fb);
  int count = __popc(R);
  // extras: ascending, unique, not already a kept static column
  for (int i = 1; i < nex; ++i)
  {
    const int32_t v = ex[i];
    int j = i - 1;
    for (; j >= 0 && ex[j] > v; --j)
      ex[j + 1] = ex[j];
    ex[j + 1] = v;
  }
  int m = 0;
  for (int i = 0; i < nex; ++i)
  {
    const int32_t v = ex[i];
    if (m > 0 && ex[m - 1] == v)
      continue;
    int lo = 0, hi = nfull; // is v a static column of the row, and kept?
    while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (fcols[fb + mid] < v)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo < nfull && fcols[fb + lo] == v)
    {
      if (!((R >> lo) & 1u))
      { // a static column that no needed cell brought: keep it as a static column
        R |= 1u << lo;
        ++count;
      }
      continue;
    }
    ex[m++] = v;
  }
  nex = m;
  count += nex;
  row_nnz[r] = count;
  const bool fast = count <= 32 && n_inc <= 32;
  if (fast)
  { // merge the kept static columns with the extras
    int32_t* dst = tmp + idx * 32;
    int e = 0, k = 0;
    for (uint32_t mm = R; mm; mm &= mm - 1)
    {
      const int32_t col = fcols[fb + __ffs(mm) - 1];
      while (e < nex && ex[e] < col)
        dst[k++] = ex[e++];
      dst[k++] = col;
    }
    while (e < nex)
      dst[k++] = ex[e++];
  }
  else
    atomicAdd(n_slow, 1ULL); // the generic fill pass writes this row's columns
  row_fast[idx] = (fast ? 1 : 0) | (band ? 2 : 0) | 16;
}

// Per-step pattern of a row WITHOUT facet-band cells, from the static full-mesh structure: the
// row keeps the full-mesh columns whose bit is set in R = OR of the full-row masks of its ACTIVE
// incident cells.  Count pass: one warp per row, coalesced inc_cell / fmask reads, one REDUX.OR,
// one POPC; only R (4 B) and the count are stored -- the columns are expanded after the scan by
// pattern_static_fill_kernel, and the assembly gather recomputes each cell's CSR positions from
// (fmask, R) instead of reading a gather table.  Only meshes whose full rows have <= 32 columns.
__global__ void __launch_bounds__(256)
    pattern_static_kernel(RowCtx rc, const int32_t* __restrict__ act_rows, DN n_act_,
                          const uint32_t* __restrict__ fmask, const uint8_t* __restrict__ frow_ok,
                          int32_t* __restrict__ row_nnz, uint32_t* __restrict__ Rrow, uint8_t* __restrict__ row_fast,
                          uint8_t* __restrict__ row_ufl,
                          unsigned long long* __restrict__ n_clist /* [0] rows, [1] nnz of contribution-list rows */)
{
  // 8 lanes per row and CROWS rows per lane group, the dependent load levels (slot -> row -> incidence ->
  // cell flags) of all of them batched: up to 3 * CROWS independent gather chains in flight per lane
  constexpr int CROWS = 4;
  const int64_t n_act = n_act_.get();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t idx0 = (t >> 3) * CROWS;
  const int sl = static_cast<int>(t & 7);
  bool go[CROWS];
  int64_t r[CROWS];
#pragma unroll
  for (int i = 0; i < CROWS; ++i)
  {
    go[i] = idx0 + i < n_act;
    r[i] = go[i] ? act_rows[idx0 + i] : 0;
  }
  int64_t ib[CROWS];
  int n_inc[CROWS];
#pragma unroll
  for (int i = 0; i < CROWS; ++i)
  {
    go[i] = go[i] && !(rc.row_flag[r[i]] & 2); // band rows: pattern_rows_kernel
    ib[i] = rc.inc_ptr[r[i]];
    n_inc[i] = go[i] ? static_cast<int>(rc.inc_ptr[r[i] + 1] - ib[i]) : 0;
  }
  uint32_t M[CROWS];
  unsigned FO[CROWS]; // low byte: OR of the incident cells' flag bytes, second byte: their AND
  bool clist = false;
  int nz = 0;
#pragma unroll
  for (int i = 0; i < CROWS; ++i)
  {
    uint32_t m = 0;
    unsigned fo = 0xFF00u;
#pragma unroll 4
    for (int k = sl; k < n_inc[i]; k += 8)
    {
      const uint32_t fm = fmask[ib[i] + k];
      const unsigned fl = rc.cell_flags[rc.inc_cell[ib[i] + k]];
      m |= (fl & 0xFD) ? fm : 0u;
      fo = (fo | fl) & ((fl << 8) | 0xFFu);
    }
    M[i] = m;
    FO[i] = fo;
  }
#pragma unroll
  for (int i = 0; i < CROWS; ++i)
  {
    uint32_t m = M[i];
    m |= __shfl_xor_sync(0xffffffffu, m, 1);
    m |= __shfl_xor_sync(0xffffffffu, m, 2);
    m |= __shfl_xor_sync(0xffffffffu, m, 4);
    unsigned fo = FO[i];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
    {
      const unsigned x = __shfl_xor_sync(0xffffffffu, fo, o);
      fo = ((fo | x) & 0xFFu) | (fo & x & 0xFF00u);
    }
    if (go[i] && sl == 0)
    {
      const bool cl = frow_ok[r[i]] != 0;
      row_nnz[r[i]] = __popc(m);
      Rrow[idx0 + i] = m;
      // every incident cell has the same flag byte and it is a standard-quadrature one
      const unsigned fl = fo & 0xFFu;
      row_ufl[idx0 + i] = (fl == (fo >> 8) && (fl & 3u) == 0u) ? static_cast<uint8_t>(fl) : uint8_t(0);
      row_fast[idx0 + i] = 1 | 4 | (cl ? 8 : 0);
      if (cl)
      {
        clist = true; // per thread: number of its contribution-list rows and their entries
        nz += __popc(m) | (1 << 16);
      }
    }
  }
  // integer counters (exact, order-independent): one atomic pair per block
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // nz packs (rows << 16 | entries) of this thread's <= CROWS rows; a warp holds 4 lane groups: no overflow
  const unsigned b = __ballot_sync(0xffffffffu, clist);
  int nrows = nz >> 16, nent = nz & 0xFFFF;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    nrows += __shfl_down_sync(0xffffffffu, nrows, o);
    nent += __shfl_down_sync(0xffffffffu, nent, o);
  }
  if ((threadIdx.x & 31) == 0 && b)
  {
    atomicAdd(&s_cnt[0], nrows);
    atomicAdd(&s_cnt[1], nent);
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x])
    atomicAdd(&n_clist[threadIdx.x], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// One THREAD per row variant of pattern_static_kernel (same outputs).  Rows of P2 triangle spaces have 2 (edge
// dofs) or 6 (vertex dofs) incident cells and a P1 tetrahedron row 24: a lane group of 8 is mostly idle in the
// first case and loops in the second, while a thread walking its own row keeps four independent
// (incidence -> cell -> flag) chains in flight and needs no shuffles.
__global__ void __launch_bounds__(256)
    pattern_static_thread_kernel(RowCtx rc, const int32_t* __restrict__ act_rows, DN n_act_,
                                 const uint32_t* __restrict__ fmask, const uint8_t* __restrict__ frow_ok,
                                 int32_t* __restrict__ row_nnz, uint32_t* __restrict__ Rrow,
                                 uint8_t* __restrict__ row_fast, uint8_t* __restrict__ row_ufl,
                                 unsigned long long* __restrict__ n_clist)
{
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  int my_rows = 0, my_nnz = 0;
  if (idx < n_act_.get())
  {
    const int64_t r = act_rows[idx];
    if (!(rc.row_flag[r] & 2)) // band rows: pattern_rows_kernel
    {
      const int64_t ib = rc.inc_ptr[r];
      const int n_inc = static_cast<int>(rc.inc_ptr[r + 1] - ib);
      uint32_t m = 0;
      unsigned fo = 0, fa = 0xFFu;
      for (int k0 = 0; k0 < n_inc; k0 += 4)
      {
        uint32_t fm[4];
        int32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          const int k = k0 + u < n_inc ? k0 + u : n_inc - 1;
          fm[u] = fmask[ib + k];
          c[u] = rc.inc_cell[ib + k];
        }
        unsigned fl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          fl[u] = rc.cell_flags[c[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          if (k0 + u >= n_inc)
            continue;
          m |= (fl[u] & 0xFD) ? fm[u] : 0u;
          fo |= fl[u];
          fa &= fl[u];
        }
      }
      const bool cl = frow_ok[r] != 0;
      const int cnt = __popc(m);
      row_nnz[r] = cnt;
      Rrow[idx] = m;
      row_ufl[idx] = (fo == fa && (fo & 3u) == 0u) ? static_cast<uint8_t>(fo) : uint8_t(0);
      row_fast[idx] = 1 | 4 | (cl ? 8 : 0);
      if (cl)
      {
        my_rows = 1;
        my_nnz = cnt;
      }
    }
  }
  // integer counters (exact, order-independent): one atomic pair per block
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    my_rows += __shfl_down_sync(0xffffffffu, my_rows, o);
    my_nnz += __shfl_down_sync(0xffffffffu, my_nnz, o);
  }
  if ((threadIdx.x & 31) == 0 && my_rows)
  {
    atomicAdd(&s_cnt[0], my_rows);
    atomicAdd(&s_cnt[1], my_nnz);
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x])
    atomicAdd(&n_clist[threadIdx.x], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// one THREAD per row variant of pattern_static_fill_kernel: the row's kept full-mesh columns, in order
__global__ void __launch_bounds__(256)
    pattern_static_fill_thread_kernel(const int32_t* __restrict__ act_rows, DN n_act_,
                                      const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ Rrow,
                                      const int64_t* __restrict__ frow_ptr, const int32_t* __restrict__ fcols,
                                      const int64_t* __restrict__ row_ptr, int32_t* __restrict__ cols)
{
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= n_act_.get() || !(row_fast[idx] & 4))
    return;
  const int64_t r = act_rows[idx];
  const uint32_t R = Rrow[idx];
  const int32_t* __restrict__ src = fcols + frow_ptr[r];
  int32_t* __restrict__ dst = cols + row_ptr[r];
  int o = 0;
  for (uint32_t m = R; m; m &= m - 1, ++o)
    dst[o] = src[__ffs(m) - 1];
}

// static rows after the scan: cols[row_ptr[r] + k] = k-th kept full-mesh column.  16 lanes per row, SROWS rows
// per half-warp with the three dependent load levels (slot -> row -> columns) of all of them batched: a warp
// that lives for one row spends its whole life waiting on three round trips.
constexpr int SROWS = 4;
__global__ void __launch_bounds__(256)
    pattern_static_fill_kernel(const int32_t* __restrict__ act_rows, DN n_act_,
                               const uint8_t* __restrict__ row_fast, const uint32_t* __restrict__ Rrow,
                               const int64_t* __restrict__ frow_ptr, const int32_t* __restrict__ fcols,
                               const int64_t* __restrict__ row_ptr, int32_t* __restrict__ cols)
{
  const int64_t n_act = n_act_.get();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t idx0 = (t >> 4) * SROWS;
  const int sl = static_cast<int>(t & 15);
  bool ok[SROWS];
  int64_t r[SROWS];
  uint32_t R[SROWS];
#pragma unroll
  for (int i = 0; i < SROWS; ++i)
  {
    const int64_t idx = idx0 + i;
    ok[i] = idx < n_act && (row_fast[idx] & 4);
    r[i] = ok[i] ? act_rows[idx] : 0;
    R[i] = ok[i] ? Rrow[idx] : 0u;
  }
  int64_t fb[SROWS], ob[SROWS];
#pragma unroll
  for (int i = 0; i < SROWS; ++i)
  {
    fb[i] = frow_ptr[r[i]];
    ob[i] = row_ptr[r[i]];
  }
  int32_t cv[SROWS][2];
#pragma unroll
  for (int i = 0; i < SROWS; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
      const int bit = sl + 16 * h;
      cv[i][h] = ((R[i] >> bit) & 1u) ? fcols[fb[i] + bit] : 0;
    }
#pragma unroll
  for (int i = 0; i < SROWS; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
      const int bit = sl + 16 * h;
      if ((R[i] >> bit) & 1u)
        cols[ob[i] + __popc(R[i] & ((1u << bit) - 1u))] = cv[i][h];
    }
}

// fast NON-static rows: columns were staged in tmp during the count pass.  One warp moves 32 rows:
// the per-row scalars are loaded once, lane-parallel, then every row is one coalesced copy.
__global__ void __launch_bounds__(256)
    pattern_copy_kernel(const int32_t* __restrict__ act_rows, DN n_act_, const uint8_t* __restrict__ row_fast,
                        const int32_t* __restrict__ tmp, int ts, const int64_t* __restrict__ row_ptr,
                        int32_t* __restrict__ cols)
{
  const int64_t n_act = n_act_.get();
  const int lane = threadIdx.x & 31;
  const int64_t idx0 = ((static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5) * 32;
  if (idx0 >= n_act)
    return;
  const int64_t idx = idx0 + lane;
  int64_t b = 0;
  int n = 0;
  if (idx < n_act && (!row_fast || (row_fast[idx] & 5) == 1))
  {
    const int64_t r = act_rows ? act_rows[idx] : idx;
    b = row_ptr[r];
    n = static_cast<int>(row_ptr[r + 1] - b);
  }
  if (__ballot_sync(0xffffffffu, n > 0) == 0)
    return;
#pragma unroll 4
  for (int i = 0; i < 32; ++i)
  {
    const int64_t bi = __shfl_sync(0xffffffffu, b, i);
    const int ni = __shfl_sync(0xffffffffu, n, i);
    for (int k = lane; k < ni; k += 32)
      cols[bi + k] = tmp[(idx0 + i) * ts + k];
  }
}

// static contribution lists (Space::fclist): one warp per row; lanes own the row's full-pattern
// columns, the incident cells' dofs are staged in shared memory and walked in ascending cell order.
template <int ND>
__global__ void __launch_bounds__(RW * 32)
    clist_kernel(int64_t n_rows, const int64_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_cell,
                 const int32_t* __restrict__ dofmap, const int64_t* __restrict__ frow_ptr,
                 const int32_t* __restrict__ fcols, uint64_t* __restrict__ fclist, uint8_t* __restrict__ frow_ok)
{
  __shared__ int32_t s_d[RW][32][ND];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * RW + w;
  if (r >= n_rows)
    return;
  const int64_t ib = inc_ptr[r];
  const int n_inc = static_cast<int>(inc_ptr[r + 1] - ib);
  const int64_t fb = frow_ptr[r];
  const int nfull = static_cast<int>(frow_ptr[r + 1] - fb);
  if (lane < n_inc)
  {
    const int64_t c = inc_cell[ib + lane];
#pragma unroll
    for (int j = 0; j < ND; ++j)
      s_d[w][lane][j] = dofmap[c * ND + j];
  }
  __syncwarp();
  const int32_t mycol = lane < nfull ? fcols[fb + lane] : -1;
  uint64_t word = ~0ull;
  int cnt = 0;
  bool ok = true;
  if (mycol == r)
    word = ~0ull ^ 1ull; // lowest byte 0xFE: diagonal
  else if (mycol >= 0)
    for (int l = 0; l < n_inc; ++l)
#pragma unroll
      for (int j = 0; j < ND; ++j)
        if (s_d[w][l][j] == mycol)
        {
          if (cnt < 8)
            word = (word & ~(0xFFull << (8 * cnt))) | (static_cast<uint64_t>(l | (j << 5)) << (8 * cnt));
          else
            ok = false;
          ++cnt;
        }
  if (lane < nfull)
    fclist[fb + lane] = word;
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0)
    frow_ok[r] = all_ok ? 1 : 0;
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

__global__ void max_i32_kernel(const int32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  int m = i < n ? v[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0)
    atomicMax(out, static_cast<unsigned long long>(m));
}

// Static full-mesh structure (every cell active): full pattern + per-incidence position masks.
// Built once per cfx_space_bind with the generic row kernel; kept only if every full row has at
// most 32 columns and 32 incident cells (P1 and P2 triangles, P1 tetrahedra on usual meshes).
template <int ND>
static void build_static_structure_nd(cfx_ctx* c, Space& S)
{
  S.has_static = false;
  if (S.stride > 32)
    return;
  RowCtx rc{S.inc_ptr.p, S.inc_cell.p, S.dofmap, nullptr, nullptr, nullptr, nullptr, nullptr, c->tdim + 1, 0};
  DevBuf<int32_t> row_nnz, tmp;
  row_nnz.reserve(c->pool, static_cast<size_t>(S.n_total) + 1);
  tmp.reserve(c->pool, static_cast<size_t>(S.n_total) * 32);
  S.fmask.reserve(c->pool, static_cast<size_t>(S.n_inc) + 1);
  unsigned long long* n_slow = reinterpret_cast<unsigned long long*>(c->scratch64.p) + 1;
  CFX_CUDA(cudaMemsetAsync(n_slow, 0, sizeof(unsigned long long), c->stream));
  CFX_CUDA(cudaMemsetAsync(row_nnz.p, 0, (static_cast<size_t>(S.n_total) + 1) * sizeof(int32_t), c->stream));
  auto k = pattern_rows_kernel<ND, false>;
  CFX_LAUNCH(c, k, grid_for(S.n_total, RW), RW * 32, 0, rc, nullptr, nullptr, dn_exact(S.n_total), 0, S.stride, 32,
             row_nnz.p, nullptr, nullptr, tmp.p, S.fmask.p, nullptr, n_slow, c->err_flag.p);
  S.frow_ptr.reserve(c->pool, static_cast<size_t>(S.n_total) + 2);
  exclusive_scan_i32_to_i64(c, row_nnz.p, S.n_total, S.frow_ptr.p);
  CFX_CUDA(cudaMemsetAsync(n_slow + 1, 0, sizeof(unsigned long long), c->stream));
  CFX_LAUNCH(c, max_i32_kernel, grid_for(S.n_total, 256), 256, 0, row_nnz.p, S.n_total, n_slow + 1);
  const int64_t* h = read_back(c, c->scratch64.p, 3);
  const int64_t fnnz = h[0], slow = h[1];
  S.max_fcols = static_cast<int>(h[2]);
  if (slow == 0)
  {
    S.fcols.reserve(c->pool, static_cast<size_t>(fnnz) + 1);
    CFX_LAUNCH(c, pattern_copy_kernel, grid_for(S.n_total, 256), 256, 0, nullptr, dn_exact(S.n_total), nullptr, tmp.p, 32,
               S.frow_ptr.p, S.fcols.p);
    S.frow_ok.reserve(c->pool, static_cast<size_t>(S.n_total) + 16);
    if (ND <= 4 && S.bs == 1)
    { // scalar P1: the one-thread-per-row gather (assemble.cu gather_matrix_p1_kernel) reads packed positions
      // (Space::fpos), not contribution lists, and takes every static row
      S.fclist.release();
      CFX_CUDA(cudaMemsetAsync(S.frow_ok.p, 1, static_cast<size_t>(S.n_total), c->stream));
    }
    else
    {
      S.fclist.reserve(c->pool, static_cast<size_t>(fnnz) + 64);
      CFX_LAUNCH(c, clist_kernel<ND>, grid_for(S.n_total, RW), RW * 32, 0, S.n_total, S.inc_ptr.p, S.inc_cell.p,
                 S.dofmap, S.frow_ptr.p, S.fcols.p, S.fclist.p, S.frow_ok.p);
    }
    S.has_static = true;
  }
  else
  {
    S.fmask.release();
    S.frow_ptr.release();
  }
  row_nnz.release();
  tmp.release();
}

static void build_static_structure(cfx_ctx* c, Space& S)
{
  switch (S.nd)
  {
  case 3: build_static_structure_nd<3>(c, S); break;
  case 4: build_static_structure_nd<4>(c, S); break;
  case 6: build_static_structure_nd<6>(c, S); break;
  default: S.has_static = false; // P2 tetrahedra: rows have more than 32 columns
  }
}

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
  // stride = longest incidence list; cell -> position-in-list table
  int* d_max = reinterpret_cast<int*>(c->scratch64.p + 2);
  CFX_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int64_t), c->stream));
  CFX_LAUNCH(c, max_degree_kernel, grid_for(S.n_total, SBK), SBK, 0, S.inc_ptr.p, S.n_total, d_max);
  S.stride = static_cast<int>(read_back(c, c->scratch64.p + 2, 1)[0] & 0xffffffffLL);
  CFX_REQUIRE(S.stride >= 1 && S.stride <= 255, CFX_ERR_UNSUPPORTED,
              "cfx_space_bind: a dof with more than 255 (or no) incident cells is not supported");
  S.lrow_built = false;
  S.fpos_built = false;
  S.has_perm = S.nd <= 6;
  if (S.has_perm)
  {
    S.fperm.reserve(c->pool, static_cast<size_t>(n_entries) + 16);
    CFX_LAUNCH(c, fperm_kernel, grid_for(S.n_total, SBK), SBK, 0, S.n_total, S.inc_ptr.p, S.inc_cell.p, S.dofmap, S.nd,
               S.fperm.p);
  }
  deg.release();
  build_static_structure(c, S);
}

void release_prepared(cfx_ctx* c, cfx_form* f)
{
  cfx_prepared* p = f->prep;
  f->prep = nullptr;
  if (!p || --p->refs > 0)
    return;
  for (auto it = c->preps.begin(); it != c->preps.end(); ++it)
    if (*it == p)
    {
      c->preps.erase(it);
      break;
    }
  p->cell_flags.release();
  p->row_flag.release();
  p->act_rows.release();
  p->band_idx.release();
  free_count_slot(c, p->d_counts, 2);
  delete p;
}

// act_rows / band_idx of a prepared domain from its row flags
static void build_row_lists(cfx_ctx* c, Space& S, cfx_prepared* P)
{
  if (!P->d_counts)
    P->d_counts = alloc_count_slot(c, 2);
  {
    // deferred-size mode: the list gets the capacity earlier steps on this space needed (with margin) and its
    // length stays on the device
    if (c->deferred && S.cap_act_rows >= 256)
      P->act_rows.reserve(c->pool, static_cast<size_t>(S.cap_act_rows));
    FlagPred p{P->row_flag.p};
    P->n_act_rows = compact_indices(c, dn_exact(S.n_total), p, P->act_rows, false, P->d_counts, &P->act_deferred);
    if (!P->act_deferred)
      S.cap_act_rows = std::max(S.cap_act_rows, with_margin(c, P->n_act_rows));
  }
  P->n_band = 0;
  if (P->facet_key.first || P->extra_key.first)
  {
    if (c->deferred && S.cap_band >= 256)
      P->band_idx.reserve(c->pool, static_cast<size_t>(S.cap_band));
    BandSlotPred bp{P->act_rows.p, P->row_flag.p};
    P->n_band = compact_indices(c, DN{P->act_deferred ? P->d_counts : nullptr, P->n_act_rows, 0}, bp, P->band_idx,
                                false, P->d_counts + 1, &P->band_deferred);
    if (!P->band_deferred)
      S.cap_band = std::max(S.cap_band, with_margin(c, P->n_band));
  }
  else
    CFX_CUDA(cudaMemsetAsync(P->d_counts + 1, 0, sizeof(int64_t), c->stream));
  P->lists_built = true;
}

// cell flags / row flags / active rows of a form (Form.h:46-89 domains).  Forms over the same
// cell domains (the bilinear and the linear form of one problem) share the result; a form without
// facet integrals may also reuse the prepared domain of one with them (a superset of rows).
void prepare_form(cfx_ctx* c, cfx_form* f, bool lists)
{
  Space& S = c->spaces[f->space];
  if (!f->dirty && f->prep)
  {
    if (lists && !f->prep->lists_built)
    {
      StageScope st(c, "prepare_form");
      build_row_lists(c, S, f->prep);
      check_call(c, "form domains (entity index out of range)");
    }
    return;
  }
  if (f->dirty_x_only && f->prep && f->prep->refs == 1 && f->prep->update_serial == c->update_serial
      && f->prep->extra_key.first == nullptr && f->n_x > 0)
  { // only inserted pattern entries have arrived since the form was prepared: mark their rows, (re)build the lists
    cfx_prepared* P = f->prep;
    StageScope st(c, "prepare_form", 4.0 * static_cast<double>(f->n_x) + 3.0 * static_cast<double>(S.n_total));
    P->extra_key = {f->xrows.p, f->n_x};
    CFX_LAUNCH(c, mark_rows_kernel, grid_for(f->n_x, SBK), SBK, 0, f->xrows.p, DN{f->d_n_x, f->n_x, 0}, S.n_total,
               P->row_flag.p, c->err_flag.p);
    if (lists || P->lists_built)
      build_row_lists(c, S, P);
    check_call(c, "form domains (entity index out of range)");
    f->gtab_serial = -1;
    f->dirty = false;
    f->dirty_x_only = false;
    return;
  }
  f->dirty_x_only = false;
  release_prepared(c, f);
  std::vector<std::pair<const void*, int64_t>> skey, rkey;
  std::pair<const void*, int64_t> fkey{nullptr, 0};
  std::map<std::pair<const void*, int64_t>, DN> key_dn; // exact sizes of the lists (on the device when deferred)
  for (auto& I : f->integrals)
  {
    if (I.facet)
    {
      if (I.n > 0)
      {
        fkey = {I.entities, I.n};
        key_dn[fkey] = DN{I.d_n, I.n, 2};
      }
      continue;
    }
    if (I.n > 0)
    {
      skey.emplace_back(I.entities, I.n);
      key_dn[skey.back()] = DN{I.d_n, I.n, 0};
    }
    if (I.rules && I.rules->nrules > 0)
    {
      rkey.emplace_back(I.rules->parent_map.p, I.rules->nrules);
      key_dn[rkey.back()] = DN{I.rules->deferred ? I.rules->d_sizes : nullptr, I.rules->nrules, 0};
    }
  }
  if (S.bs > 1 && f->rank > 0 && !blocked_on_the_fly(S, f))
  { // blocked spaces: every element tensor is materialised (bit0), standard cells included (elasticity: only
    // when the form mixes kernel families on its standard cells)
    rkey.insert(rkey.end(), skey.begin(), skey.end());
    skey.clear();
  }
  for (auto* k : {&skey, &rkey})
  {
    std::sort(k->begin(), k->end());
    k->erase(std::unique(k->begin(), k->end()), k->end());
  }
  CFX_REQUIRE(static_cast<int>(skey.size()) <= CFX_MAX_STD_LISTS, CFX_ERR_UNSUPPORTED,
              "a form may use at most 6 distinct standard-quadrature cell lists");
  const std::pair<const void*, int64_t> xkey{f->n_x > 0 ? f->xrows.p : nullptr, f->n_x};
  for (cfx_prepared* p : c->preps)
    if (p->space == f->space && p->update_serial == c->update_serial && p->std_key == skey && p->rule_key == rkey
        && (p->facet_key == fkey || fkey.first == nullptr) && (p->extra_key == xkey || xkey.first == nullptr))
    {
      f->prep = p;
      ++p->refs;
      f->gtab_serial = -1;
      f->dirty = false;
      if (lists && !p->lists_built)
      {
        StageScope st(c, "prepare_form");
        build_row_lists(c, S, p);
        check_call(c, "form domains (entity index out of range)");
      }
      return;
    }
  cfx_prepared* P = new cfx_prepared();
  P->refs = 1;
  P->space = f->space;
  P->update_serial = c->update_serial;
  P->std_key = skey;
  P->rule_key = rkey;
  P->facet_key = fkey;
  P->extra_key = xkey;
  c->preps.push_back(P);
  f->prep = P;
  StageScope st(c, "prepare_form");
  P->cell_flags.reserve(c->pool, static_cast<size_t>(c->nc_total) + 16);
  CFX_CUDA(cudaMemsetAsync(P->cell_flags.p, 0, static_cast<size_t>(c->nc_total) + 16, c->stream));
  P->row_flag.reserve(c->pool, static_cast<size_t>(S.n_total) + 16);
  CFX_CUDA(cudaMemsetAsync(P->row_flag.p, 0, static_cast<size_t>(S.n_total) + 16, c->stream));
  P->n_active_entities = 0;
  for (size_t i = 0; i < skey.size(); ++i)
  {
    CFX_LAUNCH(c, mark_cells_kernel, grid_for((skey[i].second + MC - 1) / MC, SBK), SBK, 0, static_cast<const int32_t*>(skey[i].first),
               key_dn[skey[i]], 1, c->nc_total, static_cast<uint8_t>(4u << i), S.dofmap, S.nd, uint8_t(1),
               P->cell_flags.p, P->row_flag.p, c->err_flag.p);
    P->n_active_entities += skey[i].second;
  }
  for (auto& k : rkey)
  {
    CFX_LAUNCH(c, mark_cells_kernel, grid_for((k.second + MC - 1) / MC, SBK), SBK, 0, static_cast<const int32_t*>(k.first), key_dn[k], 1,
               c->nc_total, uint8_t(1), S.dofmap, S.nd, uint8_t(1), P->cell_flags.p, P->row_flag.p, c->err_flag.p);
    P->n_active_entities += k.second;
  }
  if (fkey.first)
  { // both cells of every facet row (cell0, lf0, cell1, lf1); launched after the cell lists: 3 supersedes 1
    const int32_t* rows4 = static_cast<const int32_t*>(fkey.first);
    CFX_LAUNCH(c, mark_cells_kernel, grid_for((fkey.second + MC - 1) / MC, SBK), SBK, 0, rows4, key_dn[fkey], 4, c->nc_total, uint8_t(2),
               S.dofmap, S.nd, uint8_t(3), P->cell_flags.p, P->row_flag.p, c->err_flag.p);
    CFX_LAUNCH(c, mark_cells_kernel, grid_for((fkey.second + MC - 1) / MC, SBK), SBK, 0, rows4 + 2, key_dn[fkey], 4, c->nc_total,
               uint8_t(2), S.dofmap, S.nd, uint8_t(3), P->cell_flags.p, P->row_flag.p, c->err_flag.p);
  }
  if (xkey.first)
    CFX_LAUNCH(c, mark_rows_kernel, grid_for(xkey.second, SBK), SBK, 0, static_cast<const int32_t*>(xkey.first),
               DN{f->d_n_x, f->n_x, 0}, S.n_total, P->row_flag.p, c->err_flag.p);
  if (lists)
    build_row_lists(c, S, P);
  st.set_bytes(static_cast<double>(c->nc_total) + 3.0 * static_cast<double>(S.n_total)
               + (4.0 + 4.0 * S.nd) * static_cast<double>(P->n_active_entities) + 4.0 * static_cast<double>(P->n_act_rows));
  check_call(c, "form domains (entity index out of range)");
  f->gtab_serial = -1;
  f->dirty = false;
}

// bit mask (cell_flags) of the standard cell list `entities` inside a prepared domain
uint8_t std_list_bit(const cfx_prepared* P, const void* entities, int64_t n)
{
  for (size_t i = 0; i < P->std_key.size(); ++i)
    if (P->std_key[i].first == entities && P->std_key[i].second == n)
      return static_cast<uint8_t>(4u << i);
  throw Error(CFX_ERR_STATE, "internal: standard cell list missing from the prepared domain");
}

void settle_values(cfx_ctx* c, cfx_pattern* P)
{
  if (!P || !P->values_lazy)
    return;
  if (P->values.p)
    CFX_CUDA(cudaMemsetAsync(P->values.p, 0, (static_cast<size_t>(P->nnz) * P->bs * P->bs + 1) * sizeof(double),
                             c->stream));
  P->values_lazy = false;
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

// Make every size a form depends on known to the host (entry points that are not part of the deferred-size step
// call this first): integral entity counts, rule sizes.  Synchronises once per deferred size.
void resolve_form(cfx_ctx* c, cfx_form* f)
{
  if (!f || !f->deferred)
    return;
  check_device_error(c, "deferred sizes of a form's integration domains");
  for (auto& I : f->integrals)
  {
    if (I.d_n)
    {
      const int64_t v = read_back(c, I.d_n, 1)[0];
      I.n = I.facet ? v / 4 : v;
      I.d_n = nullptr;
    }
    if (I.rules)
      resolve(c, I.rules);
  }
  CFX_REQUIRE(f->d_n_x == nullptr, CFX_ERR_UNSUPPORTED, "form with device-side inserted pattern entries");
  f->deferred = false;
  f->dirty_x_only = false;
  f->dirty = true;
}

void set_facet_slots(cfx_ctx* c, const cfx_integral* I, bool clear)
{
  if (!I)
    return;
  CFX_REQUIRE(c->topo_bound, CFX_ERR_STATE, "interior-facet integrals need cfx_topology_bind");
  CFX_LAUNCH(c, facet_slot_set_kernel, grid_for(I->n, SBK), SBK, 0, I->entities, DN{I->d_n, I->n, 2}, c->tdim + 1,
             c->c2f, c->facet_slot.p, clear);
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
  case CFX_K_GHOST_GRAD_JUMP:
  case CFX_K_ELASTICITY:
  case CFX_K_NITSCHE_VEC: return 2;
  case CFX_K_SOURCE:
  case CFX_K_SOURCE_VEC:
  case CFX_K_NITSCHE_RHS: return 1;
  case CFX_K_ONE:
  case CFX_K_SQUARE_FN: return 0;
  }
  return -1;
}

// exterior-facet integral over facet-hosted run-time rules (ds measure with subdomain_data = rules,
// test_cut_api.py:504-527).  Shipped family: CFX_K_ONE (rank 0): c0 * measure of the selected part of the facets.
cfx_status cfx_form_add_exterior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, cfx_rules* rules,
                                                const double* constants, int n_constants)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && f && rules, CFX_ERR_INVALID, "cfx_form_add_exterior_facet_integral: NULL argument");
  CFX_REQUIRE(rules->entity_hosted, CFX_ERR_INVALID,
              "cfx_form_add_exterior_facet_integral: the rules must come from a facet-hosted cut");
  CFX_REQUIRE(kernel == CFX_K_ONE && f->rank == 0, CFX_ERR_UNSUPPORTED,
              "run-time exterior-facet integrals: only the measure functional (CFX_K_ONE, rank 0) is shipped");
  CFX_REQUIRE(n_constants >= 0 && n_constants <= CFX_MAX_CONSTANTS, CFX_ERR_INVALID, "too many constants");
  f->integrals.emplace_back();
  cfx_integral& I = f->integrals.back();
  I.kernel = kernel;
  I.facet = false;
  I.entities = nullptr;
  I.n = 0;
  I.rules = rules;
  for (int k = 0; k < n_constants; ++k)
    I.constants[k] = constants[k];
  if (n_constants == 0)
    I.constants[0] = 1.0;
  f->dirty_x_only = false;
  f->dirty = true;
  CFX_API_END(ctx)
}

// shared by the pointer and the list variants: `cells` is a HOST/DEVICE array (list == null) or the list's data
static void add_cell_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* cells, int64_t n_cells, int memspace,
                              const cfx_list* list, cfx_rules* rules, const double* constants, int n_constants)
{
  CFX_REQUIRE(ctx && f, CFX_ERR_INVALID, "cfx_form_add_cell_integral: NULL argument");
  CFX_REQUIRE(kernel_rank(kernel) == f->rank && kernel != CFX_K_GHOST_GRAD_JUMP, CFX_ERR_INVALID,
              "cfx_form_add_cell_integral: kernel family does not match the form rank / integral type");
  CFX_REQUIRE(n_constants >= 0 && n_constants <= CFX_MAX_CONSTANTS, CFX_ERR_INVALID, "too many constants");
  CFX_REQUIRE(n_cells == 0 || cells != nullptr, CFX_ERR_INVALID, "cfx_form_add_cell_integral: NULL cells");
  CFX_REQUIRE(!rules || !rules->entity_hosted, CFX_ERR_UNSUPPORTED,
              "facet-hosted rules belong to exterior / interior facet integrals, which the shipped kernel families do "
              "not cover yet");
  {
    const bool vec_kernel = kernel == CFX_K_ELASTICITY || kernel == CFX_K_SOURCE_VEC || kernel == CFX_K_NITSCHE_VEC;
    const int bs = ctx->spaces[f->space].bs;
    CFX_REQUIRE(f->rank == 0 || (vec_kernel ? bs == ctx->gdim : bs == 1), CFX_ERR_INVALID,
                "cfx_form_add_cell_integral: kernel family does not match the block size of the space");
  }
  if (kernel == CFX_K_NITSCHE || kernel == CFX_K_NITSCHE_RHS || kernel == CFX_K_NITSCHE_VEC)
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
  if (list)
  { // borrowed from the list object (which the caller keeps alive, like a DEVICE pointer); its length may be deferred
    I.entities = n_cells > 0 ? cells : nullptr;
    I.d_n = list->deferred ? list->d_n : nullptr;
  }
  else
    I.entities = n_cells > 0 ? adopt(ctx, I.own, cells, static_cast<size_t>(n_cells), memspace) : nullptr;
  I.rules = rules;
  for (int k = 0; k < n_constants; ++k)
    I.constants[k] = constants[k];
  f->dirty_x_only = false;
  f->dirty = true;
  if (I.d_n || (rules && rules->deferred))
    f->deferred = true;
  if (!list && memspace == CFX_HOST && n_cells > 0) // the caller's host array was read by an asynchronous copy
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
}

cfx_status cfx_form_add_cell_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* cells, int64_t n_cells,
                                      int memspace, cfx_rules* rules, const double* constants, int n_constants)
{
  CFX_API_BEGIN
  add_cell_integral(ctx, f, kernel, cells, n_cells, memspace, nullptr, rules, constants, n_constants);
  CFX_API_END(ctx)
}

cfx_status cfx_form_add_cell_integral_list(cfx_ctx* ctx, cfx_form* f, int kernel, const cfx_list* cells, cfx_rules* rules,
                                           const double* constants, int n_constants)
{
  CFX_API_BEGIN
  add_cell_integral(ctx, f, kernel, cells ? cells->data.p : nullptr, cells ? cells->n : 0, CFX_DEVICE, cells, rules,
                    constants, n_constants);
  CFX_API_END(ctx)
}

static void add_interior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* rows4, int64_t n_facets,
                                        int memspace, const cfx_list* list, const double* constants, int n_constants)
{
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
  if (list)
  {
    I.entities = n_facets > 0 ? rows4 : nullptr;
    I.d_n = list->deferred ? list->d_n : nullptr; // counts int32 entries: 4 per facet row
  }
  else
    I.entities = n_facets > 0 ? adopt(ctx, I.own, rows4, static_cast<size_t>(4 * n_facets), memspace) : nullptr;
  for (int k = 0; k < n_constants; ++k)
    I.constants[k] = constants[k];
  f->dirty_x_only = false;
  f->dirty = true;
  if (I.d_n)
    f->deferred = true;
  if (!list && memspace == CFX_HOST && n_facets > 0)
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
}

cfx_status cfx_form_add_interior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* rows4,
                                                int64_t n_facets, int memspace, const double* constants,
                                                int n_constants)
{
  CFX_API_BEGIN
  add_interior_facet_integral(ctx, f, kernel, rows4, n_facets, memspace, nullptr, constants, n_constants);
  CFX_API_END(ctx)
}

cfx_status cfx_form_add_interior_facet_integral_list(cfx_ctx* ctx, cfx_form* f, int kernel, const cfx_list* rows4,
                                                     const double* constants, int n_constants)
{
  CFX_API_BEGIN
  CFX_REQUIRE(rows4 != nullptr && rows4->n % 4 == 0, CFX_ERR_INVALID,
              "cfx_form_add_interior_facet_integral_list: not a list of (cell0, lf0, cell1, lf1) rows");
  add_interior_facet_integral(ctx, f, kernel, rows4->data.p, rows4->n / 4, CFX_DEVICE, rows4, constants, n_constants);
  CFX_API_END(ctx)
}

void cfx_form_free(cfx_ctx* ctx, cfx_form* f)
{
  (void)ctx;
  if (!f)
    return;
  for (auto& I : f->integrals)
    I.own.release();
  if (ctx)
    release_prepared(ctx, f);
  f->xrows.release();
  f->xcols.release();
  f->gmask.release();
  f->Rrow.release();
  f->row_ufl.release();
  f->row_fast.release();
  f->Ae.release();
  f->coeff_own.release();
  f->Fe.release();
  delete f;
}

// row_begin == 0: the whole pattern (create_sparsity_pattern + finalize).  row_begin > 0: only the
// active rows >= row_begin, generic kernels, no diagonal -- what finalize() ships to other ranks.
static void build_pattern(cfx_ctx* ctx, cfx_form* a, cfx_pattern* P, int64_t row_begin)
{
  Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  const cfx_integral* FI = facet_integral_domain(a);
  cfx_prepared* PR = a->prep;
  const bool part = row_begin > 0;
  P->space = a->space;
  P->bs = S.bs;
  P->n_rows = S.n_total;
  StageScope st(ctx, part ? "ghost_row_pattern" : "create_sparsity");
  set_facet_slots(ctx, FI, false);
  const bool has_x = a->n_x > 0 && !part;
  if (has_x)
  {
    ctx->xslot.reserve(ctx->pool, static_cast<size_t>(S.n_total) + 1);
    if (ctx->xslot.cap != ctx->xslot_init)
    {
      CFX_CUDA(cudaMemsetAsync(ctx->xslot.p, 0xff, ctx->xslot.cap * sizeof(int32_t), ctx->stream));
      ctx->xslot_init = ctx->xslot.cap;
    }
    CFX_LAUNCH(ctx, xslot_set_kernel, grid_for(a->n_x, SBK), SBK, 0, a->xrows.p, DN{a->d_n_x, a->n_x, 0}, ctx->xslot.p,
               false);
  }
  RowCtx rc{S.inc_ptr.p, S.inc_cell.p, S.dofmap, PR->cell_flags.p, PR->row_flag.p, ctx->c2f, ctx->f2c2.p,
            ctx->facet_slot.p, ctx->tdim + 1, 1, has_x ? ctx->xslot.p : nullptr, a->xrows.p, a->xcols.p,
            DN{a->d_n_x, a->n_x, 0}};
  // active rows handled by this call: the tail of the ascending list when row_begin > 0
  int64_t i0 = 0;
  if (part && PR->n_act_rows > 0)
  {
    CFX_REQUIRE(!PR->act_deferred, CFX_ERR_UNSUPPORTED,
                "cfx_create_sparsity_rows: not available on forms prepared in deferred-size mode");
    CFX_LAUNCH(ctx, lower_bound_kernel, 1, 1, 0, PR->act_rows.p, PR->n_act_rows, row_begin, ctx->scratch64.p + 4);
    i0 = read_back(ctx, ctx->scratch64.p + 4, 1)[0];
  }
  const int32_t* act = PR->act_rows.p + i0;
  const int64_t n_act = PR->n_act_rows - i0;
  const DN d_act{PR->act_deferred ? PR->d_counts : nullptr, n_act, 0};
  DevBuf<int32_t> row_nnz;
  row_nnz.reserve(ctx->pool, static_cast<size_t>(S.n_total) + 1);
  auto kcount = S.nd == 3 ? pattern_rows_kernel<3, false>
                : S.nd == 4 ? pattern_rows_kernel<4, false>
                : S.nd == 6 ? pattern_rows_kernel<6, false>
                            : pattern_rows_kernel<10, false>;
  auto kfill = S.nd == 3 ? pattern_rows_kernel<3, true>
               : S.nd == 4 ? pattern_rows_kernel<4, true>
               : S.nd == 6 ? pattern_rows_kernel<6, true>
                           : pattern_rows_kernel<10, true>;
  if (part)
    CFX_CUDA(cudaMemsetAsync(row_nnz.p + row_begin, 0, (static_cast<size_t>(S.n_total - row_begin) + 1) * sizeof(int32_t),
                             ctx->stream));
  else
    CFX_LAUNCH(ctx, pattern_inactive_count_kernel, grid_for(S.n_total, SBK), SBK, 0, PR->row_flag.p, S.n_total, 1,
               row_nnz.p);
  const unsigned ga = grid_for(n_act, RW);
  DevBuf<int32_t> tmp;
  unsigned long long* n_slow = reinterpret_cast<unsigned long long*>(ctx->scratch64.p) + 1;
  CFX_CUDA(cudaMemsetAsync(n_slow, 0, 3 * sizeof(unsigned long long), ctx->stream)); // [1] slow, [2..3] clist rows/nnz
  const bool use_static = S.has_static && !part;
  const int only_band = use_static ? 1 : 0;
  const int ts = S.bs > 1 ? 96 : 32; // columns per row kept from the count pass (pattern_rows_kernel)
  const bool need_generic = !use_static || PR->facet_key.first != nullptr || PR->extra_key.first != nullptr;
  // with a static structure the generic kernels visit the band rows only, through their slot list
  const int32_t* gslots = use_static ? PR->band_idx.p : nullptr;
  const int64_t n_generic = use_static ? PR->n_band : n_act;
  const DN d_generic = use_static ? PR->dn_band() : d_act;
  const unsigned gg = grid_for(n_generic, RW);
  a->n_band_listed = use_static ? PR->n_band : 0;
  if (n_act > 0)
  {
    a->row_fast.reserve(ctx->pool, static_cast<size_t>(n_act) + 16);
    if (use_static)
    {
      a->Rrow.reserve(ctx->pool, static_cast<size_t>(n_act) + 1);
      a->row_ufl.reserve(ctx->pool, static_cast<size_t>(n_act) + 16);
      // measured (profiles/README): rows with few incident cells (triangles: 2-6) are faster one thread per row
      // (C2: 2.56 -> 2.21 ms), rows of 24 cells (P1 tetrahedra) with the 8-lanes-per-row kernel (C3: 1.12 vs 1.16 ms)
      static const bool force_lanes = getenv("CFX_PATTERN_LANES") != nullptr; // A/B switch
      const bool lanes = force_lanes || S.stride > 12;
      // inactive, static and band rows are disjoint row sets: their count kernels (and, after the scan, their fill
      // kernels) run on lanes of their own, concurrently (every buffer is reserved on the main lane first)
      if (need_generic)
        tmp.reserve(ctx->pool, static_cast<size_t>(n_act) * ts);
      LaneScope lane(ctx, 1);
      if (lanes)
        CFX_LAUNCH(ctx, pattern_static_kernel, grid_for((n_act + 3) / 4 * 8, 256), 256, 0, rc, act, d_act, S.fmask.p,
                   S.frow_ok.p, row_nnz.p, a->Rrow.p, a->row_fast.p, a->row_ufl.p, n_slow + 1);
      else
        CFX_LAUNCH(ctx, pattern_static_thread_kernel, grid_for(n_act, 256), 256, 0, rc, act, d_act, S.fmask.p,
                   S.frow_ok.p, row_nnz.p, a->Rrow.p, a->row_fast.p, a->row_ufl.p, n_slow + 1);
    }
    if (need_generic)
    {
      tmp.reserve(ctx->pool, static_cast<size_t>(n_act) * ts);
      // scalar P1 with a static structure (the rows gather_matrix_band_p1_kernel takes): a few threads per band row
      static const bool old_band = getenv("CFX_OLD_BAND") != nullptr; // A/B switch, shared with assemble.cu
      const bool threads = use_static && S.degree == 1 && S.bs == 1 && S.has_perm && !old_band
                           && (PR->facet_key.first == nullptr || FI != nullptr);
      if (threads)
      {
        const int32_t* rows4 = FI ? FI->entities : nullptr;
        auto kb = S.nd == 3 ? pattern_band_p1_kernel<3> : pattern_band_p1_kernel<4>;
        LaneScope lane(ctx, use_static ? 2 : 0);
        CFX_LAUNCH(ctx, kb, grid_for(n_generic * BPG, BPB), BPB, 0, rc, act, gslots, d_generic, S.fmask.p, S.fperm.p,
                   S.frow_ptr.p, S.fcols.p, rows4, row_nnz.p, tmp.p, a->row_fast.p, n_slow, ctx->err_flag.p);
      }
      else
      {
        a->gmask.reserve(ctx->pool, static_cast<size_t>(n_act) * S.stride);
        LaneScope lane(ctx, use_static ? 2 : 0);
        CFX_LAUNCH(ctx, kcount, gg, RW * 32, 0, rc, act, gslots, d_generic, only_band, S.stride, ts, row_nnz.p, nullptr,
                   nullptr, tmp.p, a->gmask.p, a->row_fast.p, n_slow, ctx->err_flag.p);
      }
    }
  }
  P->row_ptr.reserve(ctx->pool, static_cast<size_t>(S.n_total) + 2);
  if (lanes_enabled(ctx))
    lane_join(ctx);
  if (part)
  { // only the tail rows can be non-empty: scan them alone
    CFX_CUDA(cudaMemsetAsync(P->row_ptr.p, 0, static_cast<size_t>(row_begin) * sizeof(int64_t), ctx->stream));
    exclusive_scan_i32_to_i64(ctx, row_nnz.p + row_begin, S.n_total - row_begin, P->row_ptr.p + row_begin);
  }
  else
    exclusive_scan_i32_to_i64(ctx, row_nnz.p, S.n_total, P->row_ptr.p);
  // deferred-size mode: a reused matrix object keeps its arrays, nnz stays on the device (row_ptr[n_rows])
  const int64_t bs2 = static_cast<int64_t>(S.bs) * S.bs;
  const bool defer = ctx->deferred && !part && P->cols.p && P->values.p && P->cols.cap >= 256
                     && P->values.cap >= (P->cols.cap - 1) * static_cast<size_t>(bs2) + 1;
  if (defer)
  {
    P->nnz = static_cast<int64_t>(P->cols.cap) - 1;
    P->deferred = true;
    P->ctx = ctx;
    note_result(ctx, P);
    a->deferred = true;
    // launch decisions that used to read counters back: what eager steps on this space saw, verified on the device
    const bool no_slow = S.seen_slow_rows == 0, no_noclist = S.seen_noclist_rows == 0;
    a->n_slow_rows = no_slow ? 0 : n_generic;
    a->n_clist_rows = use_static ? n_act : 0;
    a->n_clist_nnz = 0;
    a->n_mask_rows = n_act;
    a->expect_noclist_zero = no_noclist;
    if (n_act > 0)
      CFX_LAUNCH(ctx, pattern_capacity_kernel, 1, 1, 0, P->row_ptr.p + S.n_total, P->nnz, ctx->scratch64.p,
                 no_slow ? 1 : 0, (no_noclist && use_static) ? 1 : 0, PR->d_counts, ctx->err_flag.p);
  }
  else
  {
    const int64_t* h = read_back(ctx, ctx->scratch64.p, 4);
    P->nnz = h[0];
    P->deferred = false;
    P->ctx = ctx;
    a->n_slow_rows = h[1];
    a->n_clist_rows = h[2];
    a->n_clist_nnz = h[3];
    int64_t na = n_act, nb = use_static ? PR->n_band : 0;
    if (PR->act_deferred || PR->band_deferred)
    { // the lists were prepared with deferred sizes: fetch the exact lengths for the bookkeeping below
      const int64_t* hc = read_back(ctx, PR->d_counts, 2);
      na = hc[0];
      nb = use_static ? hc[1] : 0;
    }
    a->n_mask_rows = na - a->n_slow_rows - a->n_clist_rows;
    a->deferred = PR->act_deferred || PR->band_deferred;
    a->expect_noclist_zero = false;
    if (!part)
    {
      S.seen_slow_rows = std::max(S.seen_slow_rows, a->n_slow_rows);
      S.seen_noclist_rows = std::max(S.seen_noclist_rows, use_static ? na - nb - a->n_clist_rows : int64_t(0));
    }
    const size_t want = static_cast<size_t>(P->nnz) + 1;
    if (want > P->cols.cap || !P->cols.p || want * bs2 > P->values.cap || !P->values.p)
    {
      const size_t capn = static_cast<size_t>(with_margin(ctx, P->nnz)) + 1;
      P->cols.reserve(ctx->pool, std::max(capn, P->cols.cap));
      P->values.reserve(ctx->pool, (std::max(capn, P->cols.cap) - 1) * bs2 + 1);
    }
  }
  const int64_t cols_cap = static_cast<int64_t>(P->cols.cap) - 1;
  // the values: a full pattern zeroes its inactive rows' diagonal blocks only (cfx_pattern::values_lazy)
  static const bool eager_zero = getenv("CFX_EAGER_ZERO") != nullptr;     // A/B switch: zero everything here
  static const bool poison = getenv("CFX_POISON_VALUES") != nullptr;      // test switch: NaN where nothing was written
  const bool lazy = !part && !eager_zero;
  if (lazy && poison)
    CFX_CUDA(cudaMemsetAsync(P->values.p, 0xff, (static_cast<size_t>(P->nnz) * bs2 + 1) * sizeof(double), ctx->stream));
  if (!part)
    CFX_LAUNCH(ctx, pattern_inactive_fill_kernel, grid_for(S.n_total, SBK), SBK, 0, PR->row_flag.p, S.n_total, 1,
               P->row_ptr.p, P->cols.p, cols_cap, lazy ? P->values.p : nullptr, static_cast<int>(bs2));
  if (n_act > 0)
  {
    if (use_static)
    {
      static const bool force_lanes_fill = getenv("CFX_PATTERN_LANES") != nullptr;
      LaneScope lane(ctx, 1);
      if (force_lanes_fill || S.stride > 12)
        CFX_LAUNCH(ctx, pattern_static_fill_kernel, grid_for((n_act + SROWS - 1) / SROWS * 16, 256), 256, 0, act, d_act,
                   a->row_fast.p, a->Rrow.p, S.frow_ptr.p, S.fcols.p, P->row_ptr.p, P->cols.p);
      else
        CFX_LAUNCH(ctx, pattern_static_fill_thread_kernel, grid_for(n_act, 256), 256, 0, act, d_act, a->row_fast.p,
                   a->Rrow.p, S.frow_ptr.p, S.fcols.p, P->row_ptr.p, P->cols.p);
    }
    {
      LaneScope lane(ctx, use_static ? 2 : 0);
      if (need_generic)
        CFX_LAUNCH(ctx, pattern_copy_kernel, grid_for(n_act, 256), 256, 0, act, d_act, a->row_fast.p, tmp.p, ts,
                   P->row_ptr.p, P->cols.p);
      if (a->n_slow_rows > 0)
        CFX_LAUNCH(ctx, kfill, gg, RW * 32, 0, rc, act, gslots, d_generic, only_band, S.stride, ts, nullptr,
                   P->row_ptr.p, P->cols.p, nullptr, nullptr, a->row_fast.p, nullptr, ctx->err_flag.p);
    }
  }
  if (!lazy) // (zeroed on the main stream while the lanes fill the columns)
    CFX_CUDA(cudaMemsetAsync(P->values.p, 0, (static_cast<size_t>(P->nnz) * S.bs * S.bs + 1) * sizeof(double),
                             ctx->stream));
  P->values_zero = true;
  P->values_lazy = lazy;
  if (lanes_enabled(ctx))
    lane_join(ctx);
  tmp.release();
  P->serial = ++ctx->pattern_serial;
  a->gtab_serial = part ? -1 : P->serial; // the gather tables of a partial build index a sub-list of rows
  set_facet_slots(ctx, FI, true);
  if (has_x)
    CFX_LAUNCH(ctx, xslot_set_kernel, grid_for(a->n_x, SBK), SBK, 0, a->xrows.p, DN{a->d_n_x, a->n_x, 0}, ctx->xslot.p,
               true);
  row_nnz.release();
  st.set_bytes(12.0 * static_cast<double>(P->nnz) + 8.0 * static_cast<double>(S.n_total));
  check_call(ctx, "cfx_create_sparsity (row capacity exceeded / inserted entries not sorted by row)");
}

cfx_status cfx_create_sparsity(cfx_ctx* ctx, const cfx_form* a_const, cfx_pattern** inout)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && a && inout, CFX_ERR_INVALID, "cfx_create_sparsity: NULL argument");
  // assembler.h:444-448 "Cannot create sparsity pattern. Form is not a bilinear."
  CFX_REQUIRE(a->rank == 2, CFX_ERR_INVALID, "Cannot create sparsity pattern. Form is not a bilinear.");
  if (*inout == nullptr)
    *inout = new cfx_pattern();
  build_pattern(ctx, a, *inout, 0);
  CFX_API_END(ctx)
}

cfx_status cfx_create_sparsity_rows(cfx_ctx* ctx, const cfx_form* a_const, int64_t row_begin, cfx_pattern** inout)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && a && inout, CFX_ERR_INVALID, "cfx_create_sparsity_rows: NULL argument");
  CFX_REQUIRE(a->rank == 2, CFX_ERR_INVALID, "Cannot create sparsity pattern. Form is not a bilinear.");
  CFX_REQUIRE(row_begin > 0 && row_begin <= ctx->spaces[a->space].n_total, CFX_ERR_INVALID,
              "cfx_create_sparsity_rows: row_begin must be in (0, owned+ghost dofs]");
  if (*inout == nullptr)
    *inout = new cfx_pattern();
  resolve_form(ctx, a);
  build_pattern(ctx, a, *inout, row_begin);
  CFX_API_END(ctx)
}

cfx_status cfx_form_insert_pattern_entries(cfx_ctx* ctx, cfx_form* f, const int32_t* rows, const int32_t* cols,
                                           int64_t n, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && f, CFX_ERR_INVALID, "cfx_form_insert_pattern_entries: NULL argument");
  CFX_REQUIRE(f->rank == 2, CFX_ERR_INVALID, "cfx_form_insert_pattern_entries: form is not bilinear");
  CFX_REQUIRE(n == 0 || (rows && cols), CFX_ERR_INVALID, "cfx_form_insert_pattern_entries: NULL entries");
  f->n_x = n;
  if (n > 0)
  {
    const cudaMemcpyKind kind = memspace == CFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    f->xrows.reserve(ctx->pool, static_cast<size_t>(n));
    f->xcols.reserve(ctx->pool, static_cast<size_t>(n));
    CFX_CUDA(cudaMemcpyAsync(f->xrows.p, rows, static_cast<size_t>(n) * sizeof(int32_t), kind, ctx->stream));
    CFX_CUDA(cudaMemcpyAsync(f->xcols.p, cols, static_cast<size_t>(n) * sizeof(int32_t), kind, ctx->stream));
    if (memspace == CFX_HOST)
      CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  f->dirty_x_only = !f->dirty && f->prep != nullptr && n > 0;
  f->dirty = true;
  CFX_API_END(ctx)
}

__global__ void positions_kernel(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols_csr,
                                 int64_t n_rows, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                                 int64_t n, int64_t* __restrict__ pos, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n)
    return;
  const int32_t r = rows[i], c = cols[i];
  if (r < 0 || r >= n_rows)
  {
    err[0] = 26;
    err[1] = r;
    return;
  }
  int64_t lo = row_ptr[r], hi = row_ptr[r + 1];
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols_csr[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo >= row_ptr[r + 1] || cols_csr[lo] != c)
  {
    err[0] = 27;
    err[1] = r;
    return;
  }
  pos[i] = lo;
}

cfx_status cfx_pattern_positions(cfx_ctx* ctx, const cfx_pattern* p, const int32_t* rows, const int32_t* cols,
                                 int64_t n, int64_t* positions)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && p && (n == 0 || (rows && cols && positions)), CFX_ERR_INVALID,
              "cfx_pattern_positions: NULL argument");
  CFX_REQUIRE(p->bs == 1, CFX_ERR_UNSUPPORTED, "cfx_pattern_positions: blocked matrices are not supported yet");
  if (n > 0)
  {
    CFX_LAUNCH(ctx, positions_kernel, grid_for(n, 256), 256, 0, p->row_ptr.p, p->cols.p, p->n_rows, rows, cols, n,
               positions, ctx->err_flag.p);
    check_device_error(ctx, "cfx_pattern_positions (entry not in the sparsity pattern)");
  }
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
  P->bs = ctx->spaces[space].bs;
  const size_t bb = static_cast<size_t>(P->bs) * P->bs;
  P->n_rows = n_rows;
  P->nnz = nnz;
  P->serial = ++ctx->pattern_serial;
  P->row_ptr.reserve(ctx->pool, static_cast<size_t>(n_rows) + 1);
  P->cols.reserve(ctx->pool, static_cast<size_t>(nnz) + 1);
  P->values.reserve(ctx->pool, static_cast<size_t>(nnz) * bb + 1);
  const cudaMemcpyKind kind = memspace == CFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  CFX_CUDA(cudaMemcpyAsync(P->row_ptr.p, row_ptr, (static_cast<size_t>(n_rows) + 1) * sizeof(int64_t), kind,
                           ctx->stream));
  CFX_CUDA(cudaMemcpyAsync(P->cols.p, cols, static_cast<size_t>(nnz) * sizeof(int32_t), kind, ctx->stream));
  CFX_CUDA(cudaMemsetAsync(P->values.p, 0, (static_cast<size_t>(nnz) * bb + 1) * sizeof(double), ctx->stream));
  P->values_lazy = false;
  CFX_LAUNCH(ctx, check_sorted_kernel, grid_for(n_rows, SBK), SBK, 0, P->row_ptr.p, P->cols.p, n_rows,
             ctx->err_flag.p);
  CFX_CUDA(cudaStreamSynchronize(ctx->stream));
  check_device_error(ctx, "cfx_pattern_import (columns must be sorted and unique per row)");
  CFX_API_END(ctx)
}

cfx_status cfx_pattern_sizes(const cfx_pattern* p, int64_t* n_rows, int64_t* nnz)
{
  if (p && p->deferred && p->ctx)
  { // nnz is on the device: fetch it now (synchronises)
    try
    {
      resolve(p->ctx, const_cast<cfx_pattern*>(p));
    }
    catch (const cfx::Error& e)
    {
      cfx_set_error(p->ctx, e.what());
      return e.code;
    }
  }
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
  resolve(ctx, const_cast<cfx_pattern*>(p));
  export_to(ctx, row_ptr, p->row_ptr.p, static_cast<size_t>(p->n_rows) + 1, memspace);
  export_to(ctx, cols, p->cols.p, static_cast<size_t>(p->nnz), memspace);
  CFX_API_END(ctx)
}

const double* cfx_pattern_values_device_ptr(const cfx_pattern* p)
{
  if (!p)
    return nullptr;
  if (p->values_lazy && p->ctx) // the caller may read it: the zero fill that cfx_create_sparsity left for later
  {
    try
    {
      settle_values(p->ctx, const_cast<cfx_pattern*>(p));
    }
    catch (...)
    {
      return nullptr;
    }
  }
  const_cast<cfx_pattern*>(p)->values_zero = false; // the caller may write through the pointer
  return p->values.p;
}
const int64_t* cfx_pattern_row_ptr_device_ptr(const cfx_pattern* p) { return p ? p->row_ptr.p : nullptr; }
const int32_t* cfx_pattern_cols_device_ptr(const cfx_pattern* p) { return p ? p->cols.p : nullptr; }

cfx_status cfx_pattern_values_fetch(cfx_ctx* ctx, const cfx_pattern* p, double* values, int memspace)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && p && values, CFX_ERR_INVALID, "cfx_pattern_values_fetch: NULL argument");
  resolve(ctx, const_cast<cfx_pattern*>(p));
  settle_values(ctx, const_cast<cfx_pattern*>(p));
  export_to(ctx, values, p->values.p, static_cast<size_t>(p->nnz) * p->bs * p->bs, memspace);
  CFX_API_END(ctx)
}

int cfx_pattern_block_size(const cfx_pattern* p) { return p ? p->bs : 0; }

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

namespace cfx
{
namespace
{
struct ZeroBytePred
{ // indices whose byte is 0 (compact.cuh predicate)
  const uint8_t* flag;
  __device__ unsigned operator()(int64_t base, int64_t n) const
  {
    unsigned m = 0;
    for (int k = 0; k < 16 && base + k < n; ++k)
      m |= flag[base + k] == 0 ? (1u << k) : 0u;
    return m;
  }
};

__global__ void set_diagonal_kernel(const int32_t* __restrict__ rows, int64_t n, int64_t n_rows,
                                    const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                    double* __restrict__ vals, double diagonal, double* __restrict__ b,
                                    double rhs_value, int32_t* __restrict__ err)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i >= n)
    return;
  const int32_t r = rows[i];
  if (r < 0 || r >= n_rows)
  {
    err[0] = 29;
    err[1] = r;
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
    err[1] = r;
    return;
  }
  vals[lo] = diagonal;
  if (b)
    b[r] = rhs_value;
}
// blocked spaces: dof d -> blocked indices bs*d .. bs*d + bs-1 (deactivate.h:37-64 works on the unrolled array)
__global__ void unroll_blocked_kernel(const int32_t* __restrict__ in, int64_t n, int bs, int32_t* __restrict__ out)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i < n * bs)
    out[i] = in[i / bs] * bs + static_cast<int32_t>(i % bs);
}

__global__ void set_entries_kernel(const int32_t* __restrict__ idx, int64_t n, double* __restrict__ b, double value)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SBK + threadIdx.x;
  if (i < n)
    b[idx[i]] = value;
}
} // namespace
} // namespace cfx

extern "C"
{
const uint8_t* cfx_active_indicator_device_ptr(cfx_ctx* ctx, const cfx_form* a_const)
{
  cfx_form* a = const_cast<cfx_form*>(a_const);
  if (!ctx || !a)
    return nullptr;
  try
  {
    resolve_form(ctx, a);
    prepare_form(ctx, a);
  }
  catch (const std::exception& e)
  {
    cfx_set_error(ctx, e.what());
    return nullptr;
  }
  return a->prep->row_flag.p;
}

cfx_status cfx_inactive_dofs(cfx_ctx* ctx, const uint8_t* indicator, int64_t n_dofs_owned, cfx_list** inactive_dofs)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && indicator && inactive_dofs && n_dofs_owned >= 0, CFX_ERR_INVALID,
              "cfx_inactive_dofs: NULL argument");
  if (*inactive_dofs == nullptr)
    *inactive_dofs = new cfx_list();
  ZeroBytePred p{indicator};
  (*inactive_dofs)->n = compact_indices(ctx, n_dofs_owned, p, (*inactive_dofs)->data);
  CFX_API_END(ctx)
}

cfx_status cfx_active_domain(cfx_ctx* ctx, const cfx_form* a_const, cfx_list** active_cells, cfx_list** inactive_dofs)
{
  CFX_API_BEGIN
  cfx_form* a = const_cast<cfx_form*>(a_const);
  CFX_REQUIRE(ctx && a && active_cells && inactive_dofs, CFX_ERR_INVALID, "cfx_active_domain: NULL argument");
  // deactivate.h:76-101: a bilinear form on one space
  CFX_REQUIRE(a->rank == 2, CFX_ERR_INVALID, "cutfemx.fem.active_domain requires a bilinear form");
  resolve_form(ctx, a);
  Space& S = ctx->spaces[a->space];
  prepare_form(ctx, a);
  if (*active_cells == nullptr)
    *active_cells = new cfx_list();
  {
    FlagPred p{a->prep->cell_flags.p};
    (*active_cells)->n = compact_indices(ctx, ctx->nc_owned, p, (*active_cells)->data);
  }
  // deactivate.h:160-164
  CFX_REQUIRE((*active_cells)->n > 0, CFX_ERR_INVALID, "cutfemx.fem.active_domain found no active background cells");
  if (*inactive_dofs == nullptr)
    *inactive_dofs = new cfx_list();
  ZeroBytePred p{a->prep->row_flag.p};
  (*inactive_dofs)->n = compact_indices(ctx, S.n_owned, p, (*inactive_dofs)->data);
  if (S.bs > 1 && (*inactive_dofs)->n > 0)
  { // the reference lists the rows of the unrolled (blocked) array: bs*dof + k
    const int64_t n = (*inactive_dofs)->n;
    DevBuf<int32_t> un;
    un.reserve(ctx->pool, static_cast<size_t>(n) * S.bs);
    CFX_LAUNCH(ctx, unroll_blocked_kernel, grid_for(n * S.bs, SBK), SBK, 0, (*inactive_dofs)->data.p, n, S.bs, un.p);
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
    (*inactive_dofs)->data.release();
    (*inactive_dofs)->data = un;
    (*inactive_dofs)->n = n * S.bs;
  }
  CFX_API_END(ctx)
}

cfx_status cfx_deactivate_outside(cfx_ctx* ctx, cfx_pattern* A, const int32_t* inactive_dofs, int64_t n, int memspace,
                                  double diagonal, double* b, double rhs_value)
{
  CFX_API_BEGIN
  CFX_REQUIRE(ctx && A && (n == 0 || inactive_dofs), CFX_ERR_INVALID, "cfx_deactivate_outside: NULL argument");
  resolve(ctx, A);
  settle_values(ctx, A);
  A->values_zero = false;
  if (n > 0 && A->bs > 1)
  { // blocked matrix: the rows are blocked indices bs*dof + k (cfx_active_domain on a blocked space)
    DevBuf<int32_t> own;
    const int32_t* d = adopt(ctx, own, inactive_dofs, static_cast<size_t>(n), memspace);
    cfx_status rc = cfx_set_diagonal(ctx, A, d, n, diagonal, CFX_DEVICE);
    if (rc != CFX_OK)
      return rc;
    if (b)
      CFX_LAUNCH(ctx, set_entries_kernel, grid_for(n, SBK), SBK, 0, d, n, b, rhs_value);
    CFX_CUDA(cudaStreamSynchronize(ctx->stream));
    own.release();
    return CFX_OK;
  }
  if (n > 0)
  {
    DevBuf<int32_t> own;
    const int32_t* d = adopt(ctx, own, inactive_dofs, static_cast<size_t>(n), memspace);
    CFX_LAUNCH(ctx, set_diagonal_kernel, grid_for(n, SBK), SBK, 0, d, n, A->n_rows, A->row_ptr.p, A->cols.p,
               A->values.p, diagonal, b, rhs_value, ctx->err_flag.p);
    check_device_error(ctx, "cfx_deactivate_outside (row out of range or without a diagonal entry)");
    own.release();
  }
  CFX_API_END(ctx)
}
} // extern "C"
