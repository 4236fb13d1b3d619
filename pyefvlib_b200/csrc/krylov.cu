// krylov.cu -- north_star kernel 5: Jacobi-preconditioned CG and BiCGStab on a vectorised CSR/BSR SpMV.
//
// The matrix is BSR over the vertex graph (block = ndof x ndof, ndof = 1 for heat => plain CSR).  SpMV assigns a
// group of L lanes to a graph row; the L lanes stream the row's values and column indices with coalesced loads,
// gather x through L1/L2 and combine with warp shuffles.  Every scalar of the Krylov recurrences lives on the
// device: dot products are reduced in two deterministic stages (fixed grid, per-block partial -> every block of
// the NEXT kernel re-sums the partials in the same fixed order), so an iteration is 3 launches with no host
// round trip and no float atomics; the host only polls a `done` flag every few dozen iterations.
// These solvers replace the reference's explicit inverses (apps/heat_transfer.py:78-81,134-135,
// apps/stress_equilibrium.py:170-173, apps/geomechanics.py:140,254).
#include "krylov.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace efv {

KrylovWork *work(efv_system *s) {
  if (!s->kw) {
    KrylovWork *w = new KrylovWork();
    w->n = s->rows;
    w->grid = num_sms() * 8;
    w->partial.alloc((size_t)4 * w->grid);
    w->sc.alloc(64);
    w->flags.alloc(4);
    EFV_CUDA(cudaMallocHost(&w->h_flags, 4 * sizeof(int)));
    EFV_CUDA(cudaMallocHost(&w->h_sc, 64 * sizeof(double)));
    EFV_CUDA(cudaEventCreateWithFlags(&w->ev, cudaEventDisableTiming));
    EFV_CUDA(cudaEventCreate(&w->e0));
    EFV_CUDA(cudaEventCreate(&w->e1));
    s->kw = w;
  }
  return s->kw;
}
void krylov_free(efv_system *s) {
  if (!s->kw) return;
  KrylovWork *w = s->kw;
  if (w->h_flags) cudaFreeHost(w->h_flags);
  if (w->h_sc) cudaFreeHost(w->h_sc);
  if (w->ev) cudaEventDestroy(w->ev);
  if (w->e0) cudaEventDestroy(w->e0);
  if (w->e1) cudaEventDestroy(w->e1);
  delete w;
  s->kw = nullptr;
}

// ---- SpMV epilogue ---------------------------------------------------------------------------------------------------------
// Every SpMV kernel ends a row with epi_apply(row entry, (A v)_entry).  Besides the plain product the epilogue fuses the
// row-wise part of what follows the product, so that a Chebyshev smoothing step or a residual costs ONE pass over the
// matrix and no separate vector kernels (amg.cu):
//   kEpiPlain     y = sigma A v (* rowscale);                                   dot += w . y
//   kEpiResid     y = b - sigma A v
//   kEpiChebFirst r = dinv (b - sigma A v);  d_out = r / theta                  (v = current iterate x, left untouched)
//   kEpiChebMid   r = r_in - dinv sigma A v; d_out = c1 d_in + c2 r;  x += d_in (v = d_in, the previous direction)
//   kEpiChebLast  as Mid, but x += d_in + d_new and nothing else is written;    dot += w . x
// kEpiXZero: x is known to be zero (pre-smoothing), it is not read; kEpiRFromB: r_in = dinv b is formed on the fly.
// The scalars 1/theta, c1, c2 of the step are read from device memory (coef), computed on the device from the
// Gershgorin bound of the level, so that no host round trip sits between the assembly and the solve.
__device__ __forceinline__ double epi_apply(const SpmvEpi &e, int64_t i, double acc) {
  const double av = e.sigma * acc;
  switch (e.mode) {
    case kEpiPlain: {
      double yi = av;
      if (e.rowscale) yi *= e.rowscale[i];
      e.y[i] = yi;
      return e.w ? e.w[i] * yi : 0.0;
    }
    case kEpiResid: {
      const double yi = e.b[i] - av;
      e.y[i] = yi;
      return e.w ? e.w[i] * yi : 0.0;
    }
    case kEpiChebFirst: {
      const double r = e.dinv[i] * (e.b[i] - av);
      e.r_out[i] = r;
      e.d_out[i] = r * e.coef[0];
      return 0.0;
    }
    default: {
      const double di = e.dinv[i];
      const double r = ((e.flags & kEpiRFromB) ? di * e.b[i] : e.r_in[i]) - di * av;
      const double dprev = e.d_in[i];
      const double dn = e.coef[1] * dprev + e.coef[2] * r;
      const double x0 = (e.flags & kEpiXZero) ? 0.0 : e.x[i];
      if (e.mode == kEpiChebMid) {
        e.r_out[i] = r;
        e.d_out[i] = dn;
        e.x[i] = x0 + dprev;
        return 0.0;
      }
      const double xi = x0 + dprev + dn;
      e.x[i] = xi;
      return e.w ? e.w[i] * xi : 0.0;
    }
  }
}
// The same epilogue split in two for the persistent TMA kernel: the row-wise operands of a Chebyshev step are loaded when
// the warp STARTS a slice, so that their latency hides behind the slice's gathers instead of stalling the warp at its end
// (measured: 0.63 -> see DESIGN.md ms per fused smoothing pass at C2, against 0.44 ms for the plain product).
struct EpiOperands { double dinv, rb, dprev, x0; };
__device__ __forceinline__ void epi_prefetch(const SpmvEpi &e, int64_t i, EpiOperands &o) {
  if (e.mode < kEpiChebFirst) return;
  o.dinv = e.dinv[i];
  if (e.mode == kEpiChebFirst) { o.rb = e.b[i]; return; }
  o.rb = (e.flags & kEpiRFromB) ? e.b[i] : e.r_in[i];
  o.dprev = e.d_in[i];
  o.x0 = (e.flags & kEpiXZero) ? 0.0 : e.x[i];
}
__device__ __forceinline__ double epi_finish(const SpmvEpi &e, const EpiOperands &o, int64_t i, double acc) {
  if (e.mode < kEpiChebFirst) return epi_apply(e, i, acc);
  const double av = e.sigma * acc;
  if (e.mode == kEpiChebFirst) {
    const double r = o.dinv * (o.rb - av);
    e.r_out[i] = r;
    e.d_out[i] = r * e.coef[0];
    return 0.0;
  }
  const double r = ((e.flags & kEpiRFromB) ? o.dinv * o.rb : o.rb) - o.dinv * av;
  const double dn = e.coef[1] * o.dprev + e.coef[2] * r;
  if (e.mode == kEpiChebMid) {
    e.r_out[i] = r;
    e.d_out[i] = dn;
    e.x[i] = o.x0 + o.dprev;
    return 0.0;
  }
  const double xi = o.x0 + o.dprev + dn;
  e.x[i] = xi;
  return e.w ? e.w[i] * xi : 0.0;
}
SpmvEpi plain_epi(double *y, double sigma, const double *rowscale, const double *w) {
  SpmvEpi e;
  e.y = y; e.sigma = sigma; e.rowscale = rowscale; e.w = w;
  return e;
}

// ---- SpMV: y = sigma * A x  (optionally y *= rowscale), partial[blockIdx] = sum_i w_i * y_i -------------------------
// A group of L lanes owns a row.  Each lane first issues U independent (value, column) loads -- U*L >= the usual row
// length, so a whole row is in flight at once -- then the U gathers of x, then the FMAs; longer rows continue in a
// plain loop.  Streaming data (values, columns) is read with evict-first hints so that x stays in L1/L2.
template <int NDOF, int L, int U>
__global__ void __launch_bounds__(kBlock) k_spmv(int64_t nV, const int64_t *__restrict__ gptr,
                                                 const int32_t *__restrict__ gcol, const double *__restrict__ vals,
                                                 const double *__restrict__ x, SpmvEpi epi,
                                                 double *__restrict__ partial, const int *__restrict__ done, PeerAR ar) {
  __shared__ double red[32];
  if (done && *done) return;
  const int gl = threadIdx.x % L;
  const int64_t rows_per_block = kBlock / L;
  double dot = 0.0;
  // block-uniform trip count: every lane executes the same shuffles (sub-warp groups share a warp)
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < nV; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t v = base + threadIdx.x / L;
    const bool valid = v < nV;
    const int64_t r0 = valid ? gptr[v] : 0, r1 = valid ? gptr[v + 1] : 0;
    double acc[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) acc[i] = 0.0;
    int64_t t = r0 + gl;
    if (NDOF == 1) {
      double vv[U];
      int cc[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t tk = t + (int64_t)k * L;
        const bool ok = tk < r1;
        cc[k] = ok ? __ldcs(gcol + tk) : -1;
        vv[k] = ok ? __ldcs(vals + tk) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < U; ++k) acc[0] += (cc[k] >= 0) ? vv[k] * __ldg(x + cc[k]) : 0.0;
      for (t += (int64_t)U * L; t < r1; t += L) acc[0] += __ldcs(vals + t) * __ldg(x + __ldcs(gcol + t));
    } else {
      for (; t < r1; t += L) {
        const int64_t c = __ldcs(gcol + t);
        double xv[NDOF];
#pragma unroll
        for (int j = 0; j < NDOF; ++j) xv[j] = __ldg(x + c * NDOF + j);
        const double *blk = vals + t * (NDOF * NDOF);
#pragma unroll
        for (int i = 0; i < NDOF; ++i)
#pragma unroll
          for (int j = 0; j < NDOF; ++j) acc[i] += __ldcs(blk + i * NDOF + j) * xv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < NDOF; ++i) acc[i] = group_sum<L>(acc[i]);
    if (valid && gl == 0) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) dot += epi_apply(epi, v * NDOF + i, acc[i]);
    }
  }
  if (partial) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = dot;
    if (ar.enabled) ar_publish(ar, partial, (int)gridDim.x, 1, red);          // fused all-reduce: producer side
  }
}

// ---- sliced ELL (SELL-32) for the solver ---------------------------------------------------------------------------------
// The CSR SpMV above is bound by L1 wavefronts (ncu: 86 % L1/TEX throughput, 16 of 32 B used per sector) because the
// 27-entry row segments start at arbitrary 8-byte offsets.  For scalar systems the solver therefore keeps a copy of the
// matrix in slices of 32 consecutive rows stored column-major: lane r of a warp owns row r of the slice, every load of
// the warp is one aligned 256-byte (values) or 128-byte (columns) request, neighbouring rows gather neighbouring x, and
// no shuffles or row pointers are needed.  CSR stays the canonical structure (export, parity, assembly); the copy is
// rebuilt when the values change (one pass over the matrix, ~1 % of a solve).
__global__ void k_sell_widths(int64_t nrows, const int64_t *__restrict__ gptr, int32_t *__restrict__ width) {
  int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t nslices = (nrows + 31) >> 5;
  for (; slice < nslices; slice += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t v = slice * 32 + lane;
    int len = (v < nrows) ? (int)(gptr[v + 1] - gptr[v]) : 0;
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) width[slice] = len;
  }
}
// slice has a column beyond the owned rows (a ghost of a partitioned system)
__global__ void k_sell_ghost_flags(int64_t nrows, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol, uint8_t *__restrict__ flag) {
  int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nslices = (nrows + 31) >> 5;
  for (; slice < nslices; slice += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t v = slice * 32 + lane;
    bool g = false;
    if (v < nrows)
      for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) g |= gcol[t] >= nrows;
    const unsigned any = __ballot_sync(0xffffffffu, g);
    if (lane == 0) flag[slice] = any ? 1 : 0;
  }
}
constexpr int32_t kNoDelta = INT32_MIN;
// Offsets of slice s start at a 16-byte aligned position of the offset array so that the TMA kernel can bulk-copy
// them: q(s) = floor4(off[s] + 3 s) leaves room for width(s) entries before q(s+1).  Array length: off[n] + 3 n + 4.
__device__ __host__ __forceinline__ int64_t sell_delta_pos(int64_t o0, int64_t slice) { return (o0 + 3 * slice) & ~(int64_t)3; }
__global__ void k_sell_fill(int64_t nrows, int64_t ncols, int B, int compress, int values_only, const int64_t *__restrict__ gptr,
                            const int32_t *__restrict__ gcol, const double *__restrict__ vals,
                            const int64_t *__restrict__ off, int32_t *__restrict__ scol, int32_t *__restrict__ sdelta,
                            double *__restrict__ sval, double *__restrict__ rowabs) {
  // rowabs (optional): sum_j |a_ij| of every scalar row, for the Gershgorin bound of the multigrid smoother -- this
  // pass reads every value anyway
  const int nd = (B == 1) ? 1 : (B == 4 ? 2 : (B == 9 ? 3 : 4));
  // block rows (vertices) of a slice are the lanes; values of entry k are stored as B consecutive 32-lane columns.
  // Index compression: when every lane of a 32-entry column has the same (column - row) offset -- the rule on meshes
  // numbered like a structured grid -- the column is described by that one offset (sdelta) and the kernel never
  // reads its 32 indices; otherwise sdelta = kNoDelta and the indices are read.  Lossless either way.
  int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t nslices = (nrows + 31) >> 5;
  for (; slice < nslices; slice += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t v = slice * 32 + lane;
    const bool valid = v < nrows;
    const int64_t r0 = valid ? gptr[v] : 0;
    const int len = valid ? (int)(gptr[v + 1] - r0) : 0;
    const int width = (int)(off[slice + 1] - off[slice]);
    const int64_t o0 = off[slice];
    double ra[4] = {0.0, 0.0, 0.0, 0.0};
    if (values_only) {                                             // the index part only depends on the pattern
      for (int k = 0; k < width; ++k) {
        const bool in = k < len;
        for (int q = 0; q < B; ++q) {
          const double a = in ? vals[(r0 + k) * B + q] : 0.0;
          sval[((o0 + k) * B + q) * 32 + lane] = a;
          ra[q / nd] += fabs(a);
        }
      }
      if (rowabs && valid)
        for (int i = 0; i < nd; ++i) rowabs[v * nd + i] = ra[i];
      continue;
    }
    for (int k = 0; k < width; ++k) {
      const bool in = k < len;
      int32_t col = in ? gcol[r0 + k] : 0;
      const unsigned real = __ballot_sync(0xffffffffu, in);        // non-empty: width = longest row of the slice
      const int64_t d = (int64_t)col - v;
      const int64_t dref = __shfl_sync(0xffffffffu, d, __ffs(real) - 1);
      // padded lanes multiply a zero by some valid x entry; they can join the common offset if it stays in range
      const bool fits = in ? (d == dref) : (v + dref >= 0 && v + dref < ncols);
      const bool uniform = compress && __all_sync(0xffffffffu, fits);
      if (!in) col = uniform ? (int32_t)(v + dref) : (int32_t)(valid ? v : 0);
      scol[(o0 + k) * 32 + lane] = col;
      if (lane == 0) sdelta[sell_delta_pos(o0, slice) + k] = uniform ? (int32_t)dref : kNoDelta;
      for (int q = 0; q < B; ++q) {
        const double a = in ? vals[(r0 + k) * B + q] : 0.0;
        sval[((o0 + k) * B + q) * 32 + lane] = a;
        ra[q / nd] += fabs(a);
      }
    }
    if (rowabs && valid)
      for (int i = 0; i < nd; ++i) rowabs[v * nd + i] = ra[i];
  }
}
template <int U, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_spmv_sell(int64_t nrows, const int64_t *__restrict__ off,
                                                      const int32_t *__restrict__ scol, const int32_t *__restrict__ sdelta,
                                                      const double *__restrict__ sval,
                                                      const double *__restrict__ x, SpmvEpi epi,
                                                      double *__restrict__ partial, const int *__restrict__ done, PeerAR ar) {
  static_assert(32 % U == 0, "a batch must not straddle the 32 offsets a warp holds");
  __shared__ double red[32];
  if (done && *done) return;
  const int lane = threadIdx.x & 31;
  const int64_t nslices = (nrows + 31) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; slice < nslices; slice += warps) {
    const int64_t o0 = off[slice];
    const int width = (int)(off[slice + 1] - o0);
    const int32_t *c = scol + o0 * 32 + lane;
    const double *a = sval + o0 * 32 + lane;
    const int vi = (int)(slice * 32 + lane);
    double acc = 0.0;
    for (int k0 = 0; k0 < width; k0 += 32) {
      // one coalesced load brings the offsets of the next 32 columns; lane j keeps column k0+j's
      const int myd = (k0 + lane < width) ? __ldg(sdelta + sell_delta_pos(o0, slice) + k0 + lane) : 0;
      const int kend = min(width, k0 + 32);
      // predicated batches of U columns: every batch issues all its (column, value) loads before the gathers
      for (int k = k0; k < kend; k += U) {
        int cc[U];
        double vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ok = k + u < width;
          const int d = __shfl_sync(0xffffffffu, myd, (k + u - k0) & 31);     // warp-uniform
          cc[u] = !ok ? -1 : (d != kNoDelta ? vi + d : __ldcs(c + (int64_t)(k + u) * 32));
          vv[u] = ok ? __ldcs(a + (int64_t)(k + u) * 32) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += (cc[u] >= 0) ? vv[u] * __ldg(x + cc[u]) : 0.0;
      }
    }
    const int64_t v = slice * 32 + lane;
    if (v < nrows) dot += epi_apply(epi, v, acc);
  }
  if (partial) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = dot;
    if (ar.enabled) ar_publish(ar, partial, (int)gridDim.x, 1, red);          // fused all-reduce: producer side
  }
}

// ---- TMA-staged SELL SpMV (scalar matrices) ----------------------------------------------------------------------------
// The register-staged kernel above keeps at most (resident warps x U) 256-byte loads in flight, which is what bounds it
// once the index stream is compressed away.  Here the value stream does not pass through registers at all: every warp
// owns a ring of kSellStages shared-memory buffers and one lane issues 1-D bulk copies (cp.async.bulk, the TMA engine)
// of the next chunks -- a chunk is kSellChunk consecutive 32-entry columns of a slice, contiguous in the SELL layout --
// completing on an mbarrier.  Registers only hold the x gathers.  One persistent block per SM.
template <int CH, int ST, int W> struct SellTma {
  static constexpr int kChunk = CH, kStages = ST, kWarps = W;      // columns per chunk, ring depth, warps per block
  static constexpr size_t kStageBytes = (size_t)CH * 32 * sizeof(double) + CH * sizeof(int32_t);
  static constexpr size_t kSmem = (size_t)W * ST * kStageBytes + (size_t)W * ST * 8 + 32 * 8;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
template <class C, bool HALO>
__global__ void __launch_bounds__(C::kWarps * 32, 1)
k_spmv_sell_tma(int64_t nrows, const int64_t *__restrict__ off, const int32_t *__restrict__ scol,
                const int32_t *__restrict__ sdelta_al, const double *__restrict__ sval, const double *__restrict__ x,
                SpmvEpi epi, double *__restrict__ partial, int npartial, const int *__restrict__ done, PeerAR ar, HaloIn hin,
                int64_t s_first, int64_t s_last) {
  // [s_first, s_last): the slices this launch covers (s_last < 0: all).  Partitioned systems split a product into an
  // interior launch (slices without ghost columns, started right after the halo push) and boundary launches.
  constexpr int kSellChunk = C::kChunk, kSellStages = C::kStages, kSellWarps = C::kWarps;
  constexpr size_t kSellStageBytes = C::kStageBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  if (done && *done) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *vbuf = reinterpret_cast<double *>(smem) + (size_t)wid * kSellStages * kSellChunk * 32;
  int32_t *dbuf = reinterpret_cast<int32_t *>(smem + (size_t)kSellWarps * kSellStages * kSellChunk * 32 * sizeof(double)) +
                  (size_t)wid * kSellStages * kSellChunk;
  unsigned char *tail = smem + kSellWarps * kSellStages * kSellStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail) + wid * kSellStages;
  double *red = reinterpret_cast<double *>(tail + kSellWarps * kSellStages * 8);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kSellStages; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  const int64_t nslices = (nrows + 31) >> 5;
  // every warp owns a contiguous range of slices; the slice offsets are fetched 32 at a time (lane j holds slice
  // base+j's) so that neither cursor waits on a global load per slice
  if (s_last < 0) { s_first = 0; s_last = nslices; }
  const int64_t nsub = s_last - s_first;
  const int64_t per = (nsub + (int64_t)gridDim.x * kSellWarps - 1) / ((int64_t)gridDim.x * kSellWarps);
  const int64_t s_begin = s_first + min(nsub, ((int64_t)blockIdx.x * kSellWarps + wid) * per);
  const int64_t s_end = min(s_last, s_begin + per);
  struct Cursor { int64_t s, base, h0, h1, o0; int k, width; };
  auto fetch = [&](Cursor &c) {            // (o0, width) of slice c.s
    if (c.s >= s_end) return;
    if (c.s >= c.base + 32) {
      c.base = c.s;
      c.h0 = off[min(c.s + lane, nslices)];
      c.h1 = off[min(c.s + lane + 1, nslices)];
    }
    const int j = (int)(c.s - c.base);
    c.o0 = __shfl_sync(0xffffffffu, c.h0, j);
    c.width = (int)(__shfl_sync(0xffffffffu, c.h1, j) - c.o0);
  };
  Cursor pc{s_begin, s_begin - 32, 0, 0, 0, 0, 0}, cc = pc;
  fetch(pc);
  fetch(cc);
  unsigned issued = 0, consumed = 0;
  auto issue = [&]() {
    const int n = min(kSellChunk, pc.width - pc.k);
    const unsigned st = issued % kSellStages;
    if (lane == 0) {
      const uint32_t bar = smem_u32(bars + st);
      const uint32_t vbytes = (uint32_t)n * 256u, dbytes = (uint32_t)((n + 3) & ~3) * 4u;
      mbar_expect_tx(bar, vbytes + dbytes);
      if (n > 0) {
        bulk_g2s(smem_u32(vbuf + (size_t)st * kSellChunk * 32), sval + (pc.o0 + pc.k) * 32, vbytes, bar, policy);
        bulk_g2s(smem_u32(dbuf + (size_t)st * kSellChunk), sdelta_al + sell_delta_pos(pc.o0, pc.s) + pc.k, dbytes, bar,
                 policy);
      }
    }
    ++issued;
    pc.k += kSellChunk;
    if (pc.k >= pc.width) { ++pc.s; pc.k = 0; fetch(pc); }
  };
#pragma unroll 1
  for (int i = 0; i < kSellStages - 1 && pc.s < s_end; ++i) issue();
  double dot = 0.0, acc = 0.0;
  bool waited = false;
  EpiOperands eo = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  while (cc.s < s_end) {
    if (pc.s < s_end) issue();
    if (cc.k == 0) {                                                         // new slice: start loading its epilogue operands
      const int64_t v0 = cc.s * 32 + lane;
      if (v0 < nrows && epi.prefetch) epi_prefetch(epi, v0, eo);
    }
    const unsigned st = consumed % kSellStages;
    mbar_wait(smem_u32(bars + st), (consumed / kSellStages) & 1u);
    const int n = min(kSellChunk, cc.width - cc.k);
    const double *sv = vbuf + (size_t)st * kSellChunk * 32 + lane;
    const int myd = dbuf[(size_t)st * kSellChunk + (lane & (kSellChunk - 1))];
    static_assert(kSellChunk <= 32 && (kSellChunk & (kSellChunk - 1)) == 0, "chunk offsets live in one warp");
    const int vi = (int)(cc.s * 32 + lane);
    // indices first, gathers second: a column load between two gathers would make every gather wait for the
    // previous one on the shared scoreboard
    int idx[kSellChunk];
    if (__ballot_sync(0xffffffffu, lane < n && myd == kNoDelta) == 0u) {
#pragma unroll
      for (int u = 0; u < kSellChunk; ++u) idx[u] = vi + __shfl_sync(0xffffffffu, myd, u);
    } else {
#pragma unroll
      for (int u = 0; u < kSellChunk; ++u) {
        const int d = __shfl_sync(0xffffffffu, myd, u);                       // warp-uniform
        idx[u] = (u < n && d == kNoDelta) ? __ldcs(scol + (cc.o0 + cc.k + u) * 32 + lane) : vi + d;
      }
    }
    double xv[kSellChunk];
    if (!HALO) {
#pragma unroll
      for (int u = 0; u < kSellChunk; ++u) xv[u] = (u < n) ? __ldg(x + idx[u]) : 0.0;
    } else {
      // partitioned system: ghost entries of x are read straight from the halo staging buffer that the neighbours fill
      // over NVLink.  The warp waits for their flags the first time it meets a ghost column, so the rows that touch no
      // ghost (all but the interface planes) run while the transfer is still in flight.
      const int no = (int)hin.n_owned;
      bool g = false;
#pragma unroll
      for (int u = 0; u < kSellChunk; ++u) g |= (u < n) && (idx[u] >= no);
      if (!waited && __any_sync(0xffffffffu, g)) {
        if (lane < hin.n_nbr) peer_wait_flag(hin.flags + lane, hin.seq, hin.err, 2);
        __syncwarp();
        __threadfence_system();
        waited = true;
      }
      // one load instruction for owned and ghost entries (mixing load flavours between the gathers serialises them on the
      // scoreboard): the ghost block is addressed relative to x.  Plain cached loads are safe for the ghosts: L1 is
      // invalidated at kernel start and no thread reads a ghost line before the flags of this exchange were seen.
      const int64_t gofs = (hin.ghost - x) - (int64_t)no;
#pragma unroll
      for (int u = 0; u < kSellChunk; ++u) {
        const int64_t a = (int64_t)idx[u] + (idx[u] >= no ? gofs : (int64_t)0);
        xv[u] = (u < n) ? x[a] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < kSellChunk; ++u) acc += (u < n) ? sv[u * 32] * xv[u] : 0.0;
    __syncwarp();                                                            // buffer free for the copy issued next turn
    ++consumed;
    cc.k += kSellChunk;
    if (cc.k >= cc.width) {
      const int64_t v = cc.s * 32 + lane;
      if (v < nrows) dot += epi.prefetch ? epi_finish(epi, eo, v, acc) : epi_apply(epi, v, acc);
      acc = 0.0;
      ++cc.s;
      cc.k = 0;
      fetch(cc);
    }
  }
  if (partial) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = dot;
    for (int i = blockIdx.x + (int)gridDim.x * (1 + (int)threadIdx.x); i < npartial; i += gridDim.x * blockDim.x)
      partial[i] = 0.0;                                                      // the consumers sum npartial entries
    if (ar.enabled) ar_publish(ar, partial, npartial, 1, red);               // fused all-reduce: producer side
  }
}

// block version: lane = block row (vertex), NDOF outputs per lane, one entry = one column index + NDOF*NDOF values
template <int NDOF, int U>
__global__ void __launch_bounds__(kBlock) k_spmv_sell_bsr(int64_t nrows, const int64_t *__restrict__ off,
                                                          const int32_t *__restrict__ scol, const double *__restrict__ sval,
                                                          const double *__restrict__ x, SpmvEpi epi,
                                                          double *__restrict__ partial, const int *__restrict__ done, PeerAR ar) {
  constexpr int B = NDOF * NDOF;
  __shared__ double red[32];
  if (done && *done) return;
  const int lane = threadIdx.x & 31;
  const int64_t nslices = (nrows + 31) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int64_t slice = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; slice < nslices; slice += warps) {
    const int64_t o0 = off[slice];
    const int width = (int)(off[slice + 1] - o0);
    double acc[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) acc[i] = 0.0;
    for (int k = 0; k < width; k += U) {
      int cc[U];
      double vv[U][B];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = k + u < width;
        cc[u] = ok ? __ldcs(scol + (o0 + k + u) * 32 + lane) : -1;
#pragma unroll
        for (int q = 0; q < B; ++q) vv[u][q] = ok ? __ldcs(sval + ((o0 + k + u) * B + q) * 32 + lane) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (cc[u] >= 0) {
          double xv[NDOF];
#pragma unroll
          for (int j = 0; j < NDOF; ++j) xv[j] = __ldg(x + (int64_t)cc[u] * NDOF + j);
#pragma unroll
          for (int i = 0; i < NDOF; ++i)
#pragma unroll
            for (int j = 0; j < NDOF; ++j) acc[i] += vv[u][i * NDOF + j] * xv[j];
        }
      }
    }
    const int64_t v = slice * 32 + lane;
    if (v < nrows) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) dot += epi_apply(epi, v * NDOF + i, acc[i]);
    }
  }
  if (partial) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = dot;
    if (ar.enabled) ar_publish(ar, partial, (int)gridDim.x, 1, red);          // fused all-reduce: producer side
  }
}

// TMA-staged block version: a chunk = E consecutive entries of a slice = E * NDOF^2 value columns, contiguous in the
// block-SELL layout.  The offsets of a slice are read once per 32 entries with a coalesced load (their start is not
// 16-byte aligned at entry granularity, so they do not ride on the bulk copy).
template <int NDOF, int E, int ST, int W> struct SellBsrTma {
  static constexpr int kB = NDOF * NDOF;
  static constexpr size_t kStageBytes = (size_t)E * kB * 32 * sizeof(double);
  static constexpr size_t kSmem = (size_t)W * ST * kStageBytes + (size_t)W * ST * 8 + 32 * 8;
};
template <int NDOF, int E, int ST, int W>
__global__ void __launch_bounds__(W * 32, 1)
k_spmv_sell_bsr_tma(int64_t nrows, const int64_t *__restrict__ off, const int32_t *__restrict__ scol,
                    const int32_t *__restrict__ sdelta_al, const double *__restrict__ sval, const double *__restrict__ x,
                    SpmvEpi epi, double *__restrict__ partial, int npartial, const int *__restrict__ done, PeerAR ar) {
  using C = SellBsrTma<NDOF, E, ST, W>;
  constexpr int B = C::kB;
  extern __shared__ __align__(128) unsigned char smem[];
  if (done && *done) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *vbuf = reinterpret_cast<double *>(smem) + (size_t)wid * ST * E * B * 32;
  unsigned char *tail = smem + (size_t)W * ST * C::kStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail) + wid * ST;
  double *red = reinterpret_cast<double *>(tail + W * ST * 8);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < ST; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const int64_t nslices = (nrows + 31) >> 5;
  const int64_t per = (nslices + (int64_t)gridDim.x * W - 1) / ((int64_t)gridDim.x * W);
  const int64_t s_begin = min(nslices, ((int64_t)blockIdx.x * W + wid) * per);
  const int64_t s_end = min(nslices, s_begin + per);
  struct Cursor { int64_t s, base, h0, h1, o0; int k, width; };
  auto fetch = [&](Cursor &c) {
    if (c.s >= s_end) return;
    if (c.s >= c.base + 32) {
      c.base = c.s;
      c.h0 = off[min(c.s + lane, nslices)];
      c.h1 = off[min(c.s + lane + 1, nslices)];
    }
    const int j = (int)(c.s - c.base);
    c.o0 = __shfl_sync(0xffffffffu, c.h0, j);
    c.width = (int)(__shfl_sync(0xffffffffu, c.h1, j) - c.o0);
  };
  Cursor pc{s_begin, s_begin - 32, 0, 0, 0, 0, 0}, cc = pc;
  fetch(pc);
  fetch(cc);
  unsigned issued = 0, consumed = 0;
  auto issue = [&]() {
    const int n = min(E, pc.width - pc.k);
    const unsigned st = issued % ST;
    if (lane == 0) {
      const uint32_t bar = smem_u32(bars + st);
      const uint32_t vbytes = (uint32_t)(n > 0 ? n : 0) * B * 256u;
      mbar_expect_tx(bar, vbytes);
      if (n > 0) bulk_g2s(smem_u32(vbuf + (size_t)st * E * B * 32), sval + (pc.o0 + pc.k) * B * 32, vbytes, bar, policy);
    }
    ++issued;
    pc.k += E;
    if (pc.k >= pc.width) { ++pc.s; pc.k = 0; fetch(pc); }
  };
#pragma unroll 1
  for (int i = 0; i < ST - 1 && pc.s < s_end; ++i) issue();
  double dot = 0.0, acc[NDOF];
#pragma unroll
  for (int i = 0; i < NDOF; ++i) acc[i] = 0.0;
  int myd = 0;
#pragma unroll 1
  while (cc.s < s_end) {
    if (pc.s < s_end) issue();
    if ((cc.k & 31) == 0)                                  // offsets of the next 32 entries of this slice
      myd = (cc.k + lane < cc.width) ? __ldg(sdelta_al + sell_delta_pos(cc.o0, cc.s) + cc.k + lane) : 0;
    const unsigned st = consumed % ST;
    mbar_wait(smem_u32(bars + st), (consumed / ST) & 1u);
    const int n = min(E, cc.width - cc.k);
    const double *sv = vbuf + (size_t)st * E * B * 32 + lane;
    const int vi = (int)(cc.s * 32 + lane);
    int idx[E];
#pragma unroll
    for (int u = 0; u < E; ++u) {                          // indices first, gathers second (see k_spmv_sell_tma)
      const int d = __shfl_sync(0xffffffffu, myd, (cc.k + u) & 31);
      idx[u] = vi + d;
      if (u < n && d == kNoDelta) idx[u] = __ldcs(scol + (cc.o0 + cc.k + u) * 32 + lane);
    }
    double xv[E][NDOF];
#pragma unroll
    for (int u = 0; u < E; ++u)
#pragma unroll
      for (int j = 0; j < NDOF; ++j) xv[u][j] = (u < n) ? __ldg(x + (int64_t)idx[u] * NDOF + j) : 0.0;
#pragma unroll
    for (int u = 0; u < E; ++u)
      if (u < n) {
#pragma unroll
        for (int i = 0; i < NDOF; ++i)
#pragma unroll
          for (int j = 0; j < NDOF; ++j) acc[i] += sv[(u * B + i * NDOF + j) * 32] * xv[u][j];
      }
    __syncwarp();
    ++consumed;
    cc.k += E;
    if (cc.k >= cc.width) {
      const int64_t v = cc.s * 32 + lane;
      if (v < nrows) {
#pragma unroll
        for (int i = 0; i < NDOF; ++i) dot += epi_apply(epi, v * NDOF + i, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < NDOF; ++i) acc[i] = 0.0;
      ++cc.s;
      cc.k = 0;
      fetch(cc);
    }
  }
  if (partial) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = dot;
    for (int i = blockIdx.x + (int)gridDim.x * (1 + (int)threadIdx.x); i < npartial; i += gridDim.x * blockDim.x) partial[i] = 0.0;
    if (ar.enabled) ar_publish(ar, partial, npartial, 1, red);
  }
}

// (re)build the SELL copy if the values changed; returns whether the SELL path should be used
// rowabs (optional, rows entries): filled with sum_j |a_ij| whenever the copy is (re)filled; *rowabs_done tells the caller
bool sell_prepare(efv_system *s, KrylovWork *w, double *rowabs, bool *rowabs_done) {
  if (rowabs_done) *rowabs_done = false;
  const char *fmt = getenv("EFV_SPMV_FORMAT");
  if (fmt && strcmp(fmt, "csr") == 0) { w->sell_use = false; return false; }
  w->sell_use = true;
  const int64_t nrows = s->n_owned;
  if (w->sell_version == s->vals_version && w->sell_rows == nrows) return w->sell_ok;
  w->sell_version = s->vals_version;
  if (w->sell_rows != nrows) w->sell_slices = 0;          // different row range: rebuild the structure
  w->sell_rows = nrows;
  const int64_t nslices = (nrows + 31) >> 5;
  if (w->sell_slices != nslices) {
    // structure pass (once per pattern): slice widths -> offsets
    DBuf<int32_t> width(nslices);
    EFV_LAUNCH(k_sell_widths, grid_for(nslices * 32, 256), 256, 0, nrows, s->gptr.p, width.p);
    w->sell_off.alloc(nslices + 1);
    const int64_t cols32 = exclusive_scan_i32_to_i64(width.p, w->sell_off.p, nslices);
    w->sell_slices = nslices;
    const double padded = 32.0 * (double)cols32;
    int64_t nnz_owned = 0;
    {
      int64_t ends[2];
      EFV_CUDA(cudaMemcpyAsync(&ends[0], s->gptr.p, sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
      EFV_CUDA(cudaMemcpyAsync(&ends[1], s->gptr.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
      EFV_CUDA(cudaStreamSynchronize(stream()));
      nnz_owned = ends[1] - ends[0];
    }
    const bool force = fmt && strcmp(fmt, "sell") == 0;
    // irregular rows: stay on CSR; so do small matrices (coarse multigrid levels), where the persistent one-block-per-SM
    // kernels cost 10-17 us against 5 us of the plain CSR kernel
    w->sell_ok = force || (padded <= 1.2 * (double)nnz_owned && nrows >= 32768);
    if (!w->sell_ok) { w->sell_cols.release(); w->sell_delta.release(); w->sell_vals.release(); return false; }
    w->sell_indexed = false;
    w->sell_cols.alloc((size_t)cols32 * 32);
    w->sell_delta.alloc((size_t)cols32 + 3 * (size_t)nslices + 4);
    w->sell_vals.alloc((size_t)cols32 * 32 * s->ndof * s->ndof);
  }
  if (!w->sell_ok) return false;
  const char *cmp = getenv("EFV_SELL_DELTA");
  const int compress = !(cmp && atoi(cmp) == 0) ? 1 : 0;
  const int values_only = (w->sell_indexed && w->sell_compress == compress) ? 1 : 0;
  EFV_LAUNCH(k_sell_fill, grid_for(nslices * 32, 256), 256, 0, nrows, s->nV, s->ndof * s->ndof, compress, values_only,
             s->gptr.p, s->gcol.p, s->vals.p, w->sell_off.p, w->sell_cols.p, w->sell_delta.p, w->sell_vals.p, rowabs);
  if (rowabs_done) *rowabs_done = rowabs != nullptr;
  w->sell_indexed = true;
  w->sell_compress = compress;
  if (is_dist(s) && w->int_rows != nrows) {
    // longest run of slices without ghost columns (once per pattern): the interior launch of a partitioned product
    DBuf<uint8_t> gh(nslices);
    EFV_LAUNCH(k_sell_ghost_flags, grid_for(nslices * 32, 256), 256, 0, nrows, s->gptr.p, s->gcol.p, gh.p);
    std::vector<uint8_t> h(nslices);
    gh.download(h.data(), nslices);
    int64_t best_lo = 0, best_hi = 0, run = 0;
    for (int64_t i = 0; i <= nslices; ++i) {
      if (i < nslices && !h[i]) { ++run; continue; }
      if (run > best_hi - best_lo) { best_lo = i - run; best_hi = i; }
      run = 0;
    }
    w->int_lo = best_lo; w->int_hi = best_hi; w->int_rows = nrows;
  }
  return true;
}

// persistent SpMV kernels launch one block per SM; EFV_SPMV_BLOCKS overrides that (test hook: with a handful of blocks
// every warp walks many slices of a SMALL matrix, which exercises the slice-offset refresh and the mbarrier phase wrap
// that otherwise only occur beyond ~3.6 M rows)
static int spmv_blocks() {
  const char *e = getenv("EFV_SPMV_BLOCKS");
  const int b = e ? atoi(e) : 0;
  return b > 0 ? std::min(b, num_sms()) : num_sms();
}
// sub-range of slices for the TMA scalar kernel (interior / boundary split of partitioned products); blocks: 0 = one per SM
static thread_local int64_t g_sub_first = 0, g_sub_last = -1;
static thread_local int g_sub_blocks = 0;
static int pick_lanes(const efv_system *s) {
  const char *env = getenv("EFV_SPMV_LANES");
  if (env) { int l = atoi(env); if (l == 2 || l == 4 || l == 8 || l == 16 || l == 32) return l; }
  double avg = (double)s->nnzg / (double)(s->nV > 0 ? s->nV : 1);
  if (avg <= 6) return 2;
  if (avg <= 16) return 4;
  if (avg <= 48) return 8;
  if (avg <= 96) return 16;
  return 32;
}

template <int NDOF>
static void launch_spmv_l(efv_system *s, int grid, const double *x, const SpmvEpi &epi, double *partial, const int *done,
                          const PeerAR &ar) {
#define EFV_SPMV_CASE(LL, UU)                                                                                          \
  EFV_LAUNCH((k_spmv<NDOF, LL, UU>), grid, kBlock, 0, s->n_owned, s->gptr.p, s->gcol.p, s->vals.p, x, epi, partial, done, ar)
  const char *envu = getenv("EFV_SPMV_UNROLL");
  const int lanes = pick_lanes(s);
  int unroll = envu ? atoi(envu) : 0;
  if (!unroll) unroll = lanes <= 4 ? 8 : (lanes == 8 ? 4 : (lanes == 16 ? 2 : 1));
  switch (lanes * 16 + unroll) {
    case 2 * 16 + 8: EFV_SPMV_CASE(2, 8); break;
    case 4 * 16 + 4: EFV_SPMV_CASE(4, 4); break;
    case 4 * 16 + 8: EFV_SPMV_CASE(4, 8); break;
    case 8 * 16 + 2: EFV_SPMV_CASE(8, 2); break;
    case 8 * 16 + 4: EFV_SPMV_CASE(8, 4); break;
    case 8 * 16 + 8: EFV_SPMV_CASE(8, 8); break;
    case 16 * 16 + 1: EFV_SPMV_CASE(16, 1); break;
    case 16 * 16 + 2: EFV_SPMV_CASE(16, 2); break;
    case 16 * 16 + 4: EFV_SPMV_CASE(16, 4); break;
    case 32 * 16 + 1: EFV_SPMV_CASE(32, 1); break;
    case 32 * 16 + 2: EFV_SPMV_CASE(32, 2); break;
    default: throw Error("unsupported EFV_SPMV_LANES / EFV_SPMV_UNROLL combination");
  }
#undef EFV_SPMV_CASE
}
// does the SpMV of this system run on the TMA-staged scalar SELL kernel (the one that can read ghosts from the halo stage)?
static bool uses_tma_scalar(const efv_system *s) {
  const KrylovWork *kw = s->kw;
  const char *ek = getenv("EFV_SELL_KERNEL");
  return kw && kw->sell_use && kw->sell_ok && kw->sell_version == s->vals_version && kw->sell_rows == s->n_owned && s->ndof == 1 &&
         !(ek && strcmp(ek, "reg") == 0);
}
// SpMV whose input vector needs its ghosts refreshed first (partitioned systems; plain launch otherwise).  On the peer
// path the exchange is only PUSHED here; the TMA kernel picks the ghosts up from the staging buffer itself, every other
// kernel gets them copied behind the owned entries first.
void spmv_halo(efv_system *s, int grid, double *x, const SpmvEpi &epi, double *partial, const int *done) {
  if (!is_dist(s)) { launch_spmv(s, grid, x, epi, partial, done); return; }
  // Interior / boundary split (EFV_HALO_SPLIT=1, opt-in): the slices of the longest run of rows that have no ghost column
  // are multiplied right after the halo PUSH -- while the neighbours' values are still in flight -- and the few boundary
  // slices afterwards, by launches of the same kernel that read the ghosts from the staging buffer once the flags are up.
  // Parity-tested (tests/test_gpu_distributed.py), but measured neutral to slightly slower than push + copy kernel at
  // N = 2 (5.10 vs 4.99 ms per multigrid-CG iteration) and N = 8 (5.16 vs 5.07): the halo of one plane is not what the
  // partitioned iteration waits for, and the two extra persistent launches cost what the hidden latency saves.
  KrylovWork *kw = s->kw;
  static const int split_on = [] { const char *e = getenv("EFV_HALO_SPLIT"); return (e && atoi(e) == 1) ? 1 : 0; }();
  if (split_on && uses_tma_scalar(s) && kw->int_hi > kw->int_lo && halo_push(s->halo, x, s->ndof)) {
    const HaloIn hin = halo_in(s->halo, s->ndof);
    const int64_t ns = (s->n_owned + 31) >> 5;
    const int Gb = num_sms();                                   // partial slots of one boundary launch
    const bool lo = kw->int_lo > 0, hi = kw->int_hi < ns;
    double *pA = partial, *pB1 = partial ? partial + (grid - 2 * Gb) : nullptr, *pB2 = partial ? partial + (grid - Gb) : nullptr;
    g_sub_first = kw->int_lo; g_sub_last = kw->int_hi; g_sub_blocks = 0;
    launch_spmv(s, grid - 2 * Gb, x, epi, pA, done);            // interior: no ghosts, no wait
    auto boundary = [&](int64_t a, int64_t b, double *p, bool present) {
      if (!present) {                                           // keep the partial slots defined (the consumers sum all of them)
        if (p) EFV_CUDA(cudaMemsetAsync(p, 0, Gb * sizeof(double), stream()));
        return;
      }
      g_sub_first = a; g_sub_last = b;
      g_sub_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(Gb, (b - a + 23) / 24));
      launch_spmv(s, Gb, x, epi, p, done, PeerAR(), &hin);
    };
    boundary(0, kw->int_lo, pB1, lo);
    boundary(kw->int_hi, ns, pB2, hi);
    g_sub_first = 0; g_sub_last = -1; g_sub_blocks = 0;
    return;
  }
  // EFV_HALO_FUSED=1 (opt-in): the whole product in one launch that tests every gather for a ghost -- measured SLOWER at
  // N = 2 (5.11 vs 4.60 ms per multigrid-CG iteration at C2 weak): the test costs more in the issue-bound main loop than
  // the copy kernel and the exposed latency it saves
  const char *ef = getenv("EFV_HALO_FUSED");
  if (!(ef && atoi(ef) == 1) || !halo_push(s->halo, x, s->ndof)) {
    halo_exchange(s, x);
    launch_spmv(s, grid, x, epi, partial, done);
    return;
  }
  if (uses_tma_scalar(s)) {
    const HaloIn hin = halo_in(s->halo, s->ndof);
    launch_spmv(s, grid, x, epi, partial, done, PeerAR(), &hin);
  } else {
    halo_wait_copy(s->halo, x, s->ndof);
    launch_spmv(s, grid, x, epi, partial, done);
  }
}
void launch_spmv(efv_system *s, int grid, const double *x, const SpmvEpi &epi_in, double *partial, const int *done,
                 const PeerAR &ar, const HaloIn *hin) {
  const int64_t sub_first = g_sub_first, sub_last = g_sub_last;
  const int nblk = g_sub_blocks > 0 ? std::min(g_sub_blocks, spmv_blocks()) : spmv_blocks();
  SpmvEpi epi = epi_in;
  { static const int pf = [] { const char *e = getenv("EFV_EPI_PREFETCH"); return (e && atoi(e) == 0) ? 0 : 1; }(); epi.prefetch = pf; }
  KrylovWork *kw = s->kw;
  if (kw && kw->sell_use && kw->sell_ok && kw->sell_version == s->vals_version && kw->sell_rows == s->n_owned) {
#define EFV_SELL_ARGS s->n_owned, kw->sell_off.p, kw->sell_cols.p, kw->sell_vals.p, x, epi, partial, done, ar
#define EFV_SELL1_ARGS s->n_owned, kw->sell_off.p, kw->sell_cols.p, kw->sell_delta.p, kw->sell_vals.p, x, epi, partial, done, ar
    const char *ek = getenv("EFV_SELL_KERNEL");
    if (s->ndof == 1 && !(ek && strcmp(ek, "reg") == 0)) {
#define EFV_SELL_TMA(CH, ST, W)                                                                                       \
  do {                                                                                                                 \
    using C = SellTma<CH, ST, W>;                                                                                      \
    static bool configured = false;                                                                                    \
    if (!configured) {                                                                                                 \
      EFV_CUDA(cudaFuncSetAttribute((k_spmv_sell_tma<C, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));  \
      EFV_CUDA(cudaFuncSetAttribute((k_spmv_sell_tma<C, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));   \
      configured = true;                                                                                               \
    }                                                                                                                  \
    if (hin)                                                                                                           \
      EFV_LAUNCH((k_spmv_sell_tma<C, true>), nblk, W * 32, C::kSmem, s->n_owned, kw->sell_off.p, kw->sell_cols.p,          \
                 kw->sell_delta.p, kw->sell_vals.p, x, epi, partial, grid, done, ar, *hin, sub_first, sub_last);         \
    else                                                                                                               \
      EFV_LAUNCH((k_spmv_sell_tma<C, false>), nblk, W * 32, C::kSmem, s->n_owned, kw->sell_off.p, kw->sell_cols.p,         \
                 kw->sell_delta.p, kw->sell_vals.p, x, epi, partial, grid, done, ar, HaloIn(), sub_first, sub_last);     \
  } while (0)
      const char *ec = getenv("EFV_SELL_TMA");
      switch (ec ? atoi(ec) : 0) {
        case 1: EFV_SELL_TMA(32, 2, 12); break;
        case 2: EFV_SELL_TMA(32, 3, 8); break;
        case 3: EFV_SELL_TMA(16, 2, 24); break;
        case 4: EFV_SELL_TMA(8, 4, 24); break;
        case 5: EFV_SELL_TMA(8, 3, 32); break;
        case 6: EFV_SELL_TMA(16, 3, 16); break;
        default: EFV_SELL_TMA(16, 2, 24); break;
      }
#undef EFV_SELL_TMA
    } else if (s->ndof == 1) {
      const char *eu = getenv("EFV_SELL_UNROLL");
      const char *eb = getenv("EFV_SELL_MINB");
      switch ((eu ? atoi(eu) : 8) * 16 + (eb ? atoi(eb) : 4)) {
        case 4 * 16 + 4: EFV_LAUNCH((k_spmv_sell<4, 4>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        case 4 * 16 + 6: EFV_LAUNCH((k_spmv_sell<4, 6>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        case 4 * 16 + 8: EFV_LAUNCH((k_spmv_sell<4, 8>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        case 16 * 16 + 4: EFV_LAUNCH((k_spmv_sell<16, 3>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        case 8 * 16 + 5: EFV_LAUNCH((k_spmv_sell<8, 5>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        case 8 * 16 + 6: EFV_LAUNCH((k_spmv_sell<8, 6>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
        default: EFV_LAUNCH((k_spmv_sell<8, 4>), grid, kBlock, 0, EFV_SELL1_ARGS); break;
      }
    } else if (s->ndof == 3 && !(ek && strcmp(ek, "reg") == 0)) {
      // measured: 3x3 blocks (C3) 5.64 -> 5.11 s; 4x4 blocks with one entry per chunk (C4) 7 % slower per iteration
      // than the register-staged kernel, which therefore keeps ndof = 2 and 4
#define EFV_BSR_TMA(ND, EE)                                                                                                 \
  do {                                                                                                                      \
    using C = SellBsrTma<ND, EE, 2, 20>;                                                                                    \
    static bool configured = false;                                                                                         \
    if (!configured) {                                                                                                      \
      EFV_CUDA(cudaFuncSetAttribute((k_spmv_sell_bsr_tma<ND, EE, 2, 20>), cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    (int)C::kSmem));                                                                        \
      configured = true;                                                                                                    \
    }                                                                                                                       \
    EFV_LAUNCH((k_spmv_sell_bsr_tma<ND, EE, 2, 20>), spmv_blocks(), 20 * 32, C::kSmem, s->n_owned, kw->sell_off.p,              \
               kw->sell_cols.p, kw->sell_delta.p, kw->sell_vals.p, x, epi, partial, grid, done, ar);                        \
  } while (0)
      EFV_BSR_TMA(3, 2);
#undef EFV_BSR_TMA
    } else if (s->ndof == 2) {
      EFV_LAUNCH((k_spmv_sell_bsr<2, 4>), grid, kBlock, 0, EFV_SELL_ARGS);
    } else if (s->ndof == 3) {
      EFV_LAUNCH((k_spmv_sell_bsr<3, 2>), grid, kBlock, 0, EFV_SELL_ARGS);
    } else {
      EFV_LAUNCH((k_spmv_sell_bsr<4, 2>), grid, kBlock, 0, EFV_SELL_ARGS);
    }
#undef EFV_SELL_ARGS
#undef EFV_SELL1_ARGS
    return;
  }
  switch (s->ndof) {
    case 1: launch_spmv_l<1>(s, grid, x, epi, partial, done, ar); break;
    case 2: launch_spmv_l<2>(s, grid, x, epi, partial, done, ar); break;
    case 3: launch_spmv_l<3>(s, grid, x, epi, partial, done, ar); break;
    case 4: launch_spmv_l<4>(s, grid, x, epi, partial, done, ar); break;
    default: throw Error("unsupported ndof");
  }
}

void spmv(efv_system *s, const double *x, double *y) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  sell_prepare(s, work(s));
  launch_spmv(s, num_sms() * 8, x, plain_epi(y, 1.0, nullptr, nullptr), nullptr, nullptr);
}

// multi-GPU: collapse `narr` partial arrays of length G into narr scalars (then all-reduced over the ranks); the
// consumer kernels read them as "partial arrays of length 1"
__global__ void k_reduce_partials(const double *__restrict__ partial, int G, int narr, double *__restrict__ out) {
  __shared__ double red[32];
  for (int a = 0; a < narr; ++a) {
    double v = sum_partials(partial + (size_t)a * G, G, red);
    if (threadIdx.x == 0) out[a] = v;
  }
}
// returns the (pointer, length) pair the consumer kernels should use for `narr` consecutive partial arrays
// a partitioned system in a multi-rank job (a plain system created next to it keeps solving alone)
bool is_dist(const efv_system *s) { return comm_active() && s->halo.active; }
const double *consolidate(efv_system *s, const double *partial, int narr, int slot, int *Gout) {
  KrylovWork *w = s->kw;
  if (!is_dist(s)) { *Gout = w->grid; return partial; }
  double *out = w->sc.p + 32 + 4 * slot;
  if (comm_peer_active()) {
    comm_allreduce_partials(partial, w->grid, narr, out, 0);     // one launch: fold partials + all-reduce over peer memory
  } else {
    EFV_LAUNCH(k_reduce_partials, 1, kBlock, 0, partial, w->grid, narr, out);
    comm_allreduce_sum(out, narr);
  }
  *Gout = 1;
  return out;
}

// ---- Jacobi -----------------------------------------------------------------------------------------------------------
__global__ void k_jacobi(int64_t nV, int ndof, const int32_t *__restrict__ diag_slot, const double *__restrict__ vals,
                         double sigma, double *__restrict__ dinv) {
  int64_t n = nV * ndof;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i / ndof;
    int f = (int)(i % ndof);
    double d = sigma * vals[((int64_t)diag_slot[v] * ndof + f) * ndof + f];
    dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
  }
}

// ---- PCG kernels ---------------------------------------------------------------------------------------------------------
// r = sigma*b - y (y = sigma*A*x0); z = dinv r; p = z; partials: [0] r.z  [1] r.r  [2] b.b
__global__ void __launch_bounds__(kBlock) k_pcg_init(int64_t n, const double *__restrict__ b, const double *__restrict__ y,
                                                     const double *__restrict__ dinv, double sigma, double *__restrict__ r,
                                                     double *__restrict__ p, double *__restrict__ partial) {
  __shared__ double red[32];
  double rz = 0, rr = 0, bb = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = b[i];
    double ri = sigma * bi - y[i];
    double zi = dinv[i] * ri;
    r[i] = ri; p[i] = zi;
    rz += ri * zi; rr += ri * ri; bb += bi * bi;
  }
  int G = gridDim.x;
  rz = block_sum(rz, red); if (threadIdx.x == 0) partial[blockIdx.x] = rz;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = rr;
  bb = block_sum(bb, red); if (threadIdx.x == 0) partial[2 * G + blockIdx.x] = bb;
}
// sc: [0],[1] rz ping-pong, [2] b.b, [3] r.r, [4] rtol^2
__global__ void k_pcg_init_scalars(int G, const double *__restrict__ partial, double *__restrict__ sc, int *__restrict__ flags,
                                   double rtol) {
  __shared__ double red[32];
  double rz = sum_partials(partial, G, red);
  double rr = sum_partials(partial + G, G, red);
  double bb = sum_partials(partial + 2 * G, G, red);
  if (threadIdx.x == 0) {
    sc[0] = rz; sc[1] = 0.0; sc[2] = bb; sc[3] = rr; sc[4] = rtol * rtol;
    flags[0] = (rr <= rtol * rtol * bb) ? 1 : 0;
    flags[1] = 0; flags[2] = 0;
  }
}
// alpha = rz / p.Ap ; x += alpha p ; r -= alpha Ap ; partials [0] r.z(new) [1] r.r(new)
__global__ void __launch_bounds__(kBlock) k_pcg_update_xr(int64_t n, int it, const double *__restrict__ sc,
                                                          const double *__restrict__ pAp_partial, int G,
                                                          const double *__restrict__ p, const double *__restrict__ Ap,
                                                          const double *__restrict__ dinv, double *__restrict__ x,
                                                          double *__restrict__ r, double *__restrict__ partial, int Gout,
                                                          const int *__restrict__ flags, PeerAR ar_in, PeerAR ar_out) {
  __shared__ double red[32];
  if (flags[0]) return;
  double pAp;
  if (ar_in.enabled) { ar_wait(ar_in); pAp = ar_value(ar_in, 0); }            // fused all-reduce: consumer side
  else pAp = sum_partials(pAp_partial, G, red);
  if (!(pAp > 0.0) || !(pAp < 1e300)) return;        // not SPD / non-finite: k_pcg_update_p raises the breakdown flag, x stays as it is
  double alpha = sc[it & 1] / pAp;
  double rz = 0, rr = 0;
  // 16-byte accesses (buffers come from the pool: 256-byte aligned); the odd tail entry is done by one thread
  const int64_t n2 = n >> 1;
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *Ap2 = reinterpret_cast<const double2 *>(Ap);
  const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
  double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 pv = p2[i], av = Ap2[i], dv = d2[i];
    double2 xv = x2[i], rv = r2[i];
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * av.x; rv.y -= alpha * av.y;
    x2[i] = xv; r2[i] = rv;
    rz += rv.x * rv.x * dv.x + rv.y * rv.y * dv.y;
    rr += rv.x * rv.x + rv.y * rv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    rz += ri * ri * dinv[i];
    rr += ri * ri;
  }
  rz = block_sum(rz, red); if (threadIdx.x == 0) partial[blockIdx.x] = rz;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[Gout + blockIdx.x] = rr;
  if (ar_out.enabled) ar_publish(ar_out, partial, Gout, 2, red);
}
// beta = rz_new / rz_old ; p = dinv r + beta p ; block 0 publishes the scalars and the convergence flag
__global__ void __launch_bounds__(kBlock) k_pcg_update_p(int64_t n, int it, double *__restrict__ sc,
                                                         const double *__restrict__ partial, int G,
                                                         const double *__restrict__ r, const double *__restrict__ dinv,
                                                         double *__restrict__ p, int *__restrict__ flags, PeerAR ar_in,
                                                         const double *__restrict__ pAp_partial, int GpAp) {
  __shared__ double red[32];
  if (flags[0]) return;
  if (pAp_partial) {
    const double pAp = sum_partials(pAp_partial, GpAp, red);
    if (!(pAp > 0.0) || !(pAp < 1e300)) {            // same test as k_pcg_update_xr, which left x and r untouched
      if (blockIdx.x == 0 && threadIdx.x == 0) { flags[2] = 1; flags[0] = 1; }
      return;
    }
  }
  double rz_new, rr;
  if (ar_in.enabled) { ar_wait(ar_in); rz_new = ar_value(ar_in, 0); rr = ar_value(ar_in, 1); }
  else { rz_new = sum_partials(partial, G, red); rr = sum_partials(partial + G, G, red); }
  double beta = rz_new / sc[it & 1];
  const int64_t n2 = n >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r), *d2 = reinterpret_cast<const double2 *>(dinv);
  double2 *p2 = reinterpret_cast<double2 *>(p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 rv = r2[i], dv = d2[i];
    double2 pv = p2[i];
    pv.x = dv.x * rv.x + beta * pv.x;
    pv.y = dv.y * rv.y + beta * pv.y;
    p2[i] = pv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = dinv[n - 1] * r[n - 1] + beta * p[n - 1];
  // every block must have read flags[0]/sc before block 0 overwrites them: the values written here are only
  // consumed by the NEXT kernel (sc[(it+1)&1] is the slot nobody reads in this one)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[(it + 1) & 1] = rz_new;
    sc[3] = rr;
    flags[1] = it + 1;
    // benign race: a block of THIS kernel that has not started yet may see done=1 and skip its p update,
    // which is irrelevant once converged (x and r are final after k_pcg_update_xr)
    if (!(rr > sc[4] * sc[2])) flags[0] = 1;
  }
}

void solve_pcg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  ensure_vectors(s);
  KrylovWork *w = work(s);
  const int64_t n = s->n_owned * s->ndof;          // owned entries; ghosts of x / p live behind them
  if (!w->r.n) { w->r.alloc(s->rows); w->p.alloc(s->rows); w->Ap.alloc(s->rows); w->dinv.alloc(s->rows); }
  const double sigma = negate ? -1.0 : 1.0;
  const int G = w->grid;
  double *P = w->partial.p;
  int Gc = G;
  EFV_CUDA(cudaEventRecord(w->e0, stream()));
  sell_prepare(s, w);
  EFV_LAUNCH(k_jacobi, grid_for(n, 256), 256, 0, s->n_owned, s->ndof, s->diag_slot.p, s->vals.p, sigma, w->dinv.p);
  EFV_CUDA(cudaMemsetAsync(w->flags.p, 0, 4 * sizeof(int), stream()));
  halo_exchange(s, s->x.p);
  launch_spmv(s, G, s->x.p, plain_epi(w->Ap.p, sigma, nullptr, nullptr), nullptr, nullptr);
  EFV_LAUNCH(k_pcg_init, G, kBlock, 0, n, s->b.p, w->Ap.p, w->dinv.p, sigma, w->r.p, w->p.p, P);
  const double *c0 = consolidate(s, P, 3, 0, &Gc);
  EFV_LAUNCH(k_pcg_init_scalars, 1, kBlock, 0, Gc, c0, w->sc.p, w->flags.p, rtol);
  const int check_every = 32;
  int it = 0;
  bool done = false;
  const char *efuse = getenv("EFV_FUSED_ALLREDUCE");
  // opt-in: measured neutral at N=2 (135.6 vs 135.9 ms/step) and 10 % slower at N=8 (146.8 vs 133.4 ms/step at C2 weak),
  // where every block of the consumer polls the flags the peers are writing over NVLink
  const bool fused = is_dist(s) && comm_peer_active() && efuse && atoi(efuse) == 1;
  unsigned *counter = reinterpret_cast<unsigned *>(w->flags.p + 3);        // zeroed with the flags above
  while (!done && it < maxit) {
    int batch_end = std::min(maxit, it + check_every);
    for (; it < batch_end; ++it) {
      if (fused) {
        halo_exchange(s, w->p.p);
        // the two all-reduces of the iteration ride on the kernels: the last block of the producer publishes to every
        // rank's peer region, the blocks of the consumer wait for all ranks (peer.cuh)
        const PeerAR ar1 = comm_peer_ar(counter), ar2 = comm_peer_ar(counter);
        launch_spmv(s, G, w->p.p, plain_epi(w->Ap.p, sigma, nullptr, w->p.p), P + 3 * G, w->flags.p, ar1);
        EFV_LAUNCH(k_pcg_update_xr, G, kBlock, 0, n, it, w->sc.p, P + 3 * G, G, w->p.p, w->Ap.p, w->dinv.p, s->x.p, w->r.p, P, G,
                   w->flags.p, ar1, ar2);
        EFV_LAUNCH(k_pcg_update_p, G, kBlock, 0, n, it, w->sc.p, P, G, w->r.p, w->dinv.p, w->p.p, w->flags.p, ar2, (const double *)nullptr, 0);
        continue;
      }
      spmv_halo(s, G, w->p.p, plain_epi(w->Ap.p, sigma, nullptr, w->p.p), P + 3 * G, w->flags.p);
      int Gc1 = G;
      const double *c1 = consolidate(s, P + 3 * G, 1, 1, &Gc1);
      Gc = Gc1;
      EFV_LAUNCH(k_pcg_update_xr, G, kBlock, 0, n, it, w->sc.p, c1, Gc, w->p.p, w->Ap.p, w->dinv.p, s->x.p, w->r.p, P, G,
                 w->flags.p, PeerAR(), PeerAR());
      const double *c2 = consolidate(s, P, 2, 2, &Gc);
      EFV_LAUNCH(k_pcg_update_p, G, kBlock, 0, n, it, w->sc.p, c2, Gc, w->r.p, w->dinv.p, w->p.p, w->flags.p, PeerAR(), c1, Gc1);
    }
    EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    done = w->h_flags[0] != 0;
  }
  EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaMemcpyAsync(w->h_sc, w->sc.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaEventRecord(w->e1, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  float ms = 0;
  EFV_CUDA(cudaEventElapsedTime(&ms, w->e0, w->e1));
  s->t_solve_ms = ms;
  if (is_dist(s) && comm_peer_active()) EFV_REQUIRE(comm_peer_error() == 0, "peer-memory collective timed out (a rank did not arrive)");
  if (iters) *iters = w->h_flags[1];
  // breakdown (p.Ap <= 0 or a non-finite scalar: the operator is not SPD): x is the last good iterate, the residual is
  // reported as NaN so that callers never mistake it for convergence
  if (relres) *relres = w->h_flags[2] ? NAN : ((w->h_sc[2] > 0) ? std::sqrt(w->h_sc[3] / w->h_sc[2]) : 0.0);
}

// ---- BiCGStab on the row-scaled system D^-1 A x = D^-1 b ---------------------------------------
// sc: [0] rho, [1] alpha, [2] omega, [3] ||b~||^2, [4] ||r~||^2, [5] rtol^2
__global__ void __launch_bounds__(kBlock) k_bi_init(int64_t n, const double *__restrict__ b, const double *__restrict__ y,
                                                    const double *__restrict__ dinv, double *__restrict__ r,
                                                    double *__restrict__ rhat, double *__restrict__ p, double *__restrict__ v,
                                                    double *__restrict__ partial) {
  __shared__ double red[32];
  double rr = 0, bb = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double bi = dinv[i] * b[i];
    double ri = bi - y[i];
    r[i] = ri; rhat[i] = ri; p[i] = 0.0; v[i] = 0.0;
    rr += ri * ri; bb += bi * bi;
  }
  int G = gridDim.x;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[blockIdx.x] = rr;
  bb = block_sum(bb, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = bb;
}
__global__ void k_bi_init_scalars(int G, const double *__restrict__ partial, double *__restrict__ sc, int *__restrict__ flags,
                                  double rtol) {
  __shared__ double red[32];
  double rr = sum_partials(partial, G, red);
  double bb = sum_partials(partial + G, G, red);
  if (threadIdx.x == 0) {
    sc[0] = 1.0; sc[1] = 1.0; sc[2] = 1.0; sc[3] = bb; sc[4] = rr; sc[5] = rtol * rtol; sc[6] = rr;   // sc[6] = (rhat, r)
    sc[8] = rr;                                                                                       // |rhat|^2
    flags[0] = (rr <= rtol * rtol * bb) ? 1 : 0;
    flags[1] = 0; flags[2] = 0; flags[3] = 0;
  }
}
// beta = (rho_new/rho)(alpha/omega); p = r + beta (p - omega v)
__global__ void __launch_bounds__(kBlock) k_bi_p(int64_t n, const double *__restrict__ sc, const double *__restrict__ r,
                                                 const double *__restrict__ v, double *__restrict__ p,
                                                 double *__restrict__ rhat, const int *__restrict__ flags) {
  if (flags[0]) return;
  if (flags[2]) {     // restart after a breakdown ((rhat, r) or omega vanished): shadow residual <- r, p <- r
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double ri = r[i];
      rhat[i] = ri;
      p[i] = ri;
    }
    return;
  }
  double rho_new = sc[6], rho = sc[0], alpha = sc[1], omega = sc[2];
  double beta = (rho_new / rho) * (alpha / omega);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = r[i] + beta * (p[i] - omega * v[i]);
}
// alpha = rho_new / (rhat, v); s = r - alpha v ; block 0 stores alpha
__global__ void __launch_bounds__(kBlock) k_bi_s(int64_t n, double *__restrict__ sc, const double *__restrict__ rv_partial, int G,
                                                 const double *__restrict__ r, const double *__restrict__ v,
                                                 double *__restrict__ sv, const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  double rv = sum_partials(rv_partial, G, red);
  double alpha = sc[6] / rv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    sv[i] = r[i] - alpha * v[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) sc[7] = alpha;     // staged alpha (sc[1] still read by nobody in this kernel)
}
__global__ void k_bi_clear_restart(int *flags) { flags[2] = 0; }
// t = A s done by spmv with partial (t,s); (t,t) needs its own pass fused here:
__global__ void __launch_bounds__(kBlock) k_bi_tt(int64_t n, const double *__restrict__ t, double *__restrict__ partial,
                                                  const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  double tt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) tt += t[i] * t[i];
  tt = block_sum(tt, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tt;
}
// omega = (t,s)/(t,t); x += alpha p + omega s; r = s - omega t; partials (rhat,r), (r,r)
// px / sx: the vectors that update x (p and s, or M^-1 p and M^-1 s under right preconditioning)
__global__ void __launch_bounds__(kBlock) k_bi_xr(int64_t n, const double *__restrict__ sc, const double *__restrict__ ts_partial,
                                                  const double *__restrict__ tt_partial, int G, const double *__restrict__ px,
                                                  const double *__restrict__ sv, const double *__restrict__ sx, const double *__restrict__ t,
                                                  const double *__restrict__ rhat, double *__restrict__ x,
                                                  double *__restrict__ r, double *__restrict__ partial,
                                                  const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  double ts = sum_partials(ts_partial, G, red);
  double tt = sum_partials(tt_partial, G, red);
  double omega = (tt > 0) ? ts / tt : 0.0;
  double alpha = sc[7];
  double hr = 0, rr = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * px[i] + omega * sx[i];
    double ri = sv[i] - omega * t[i];
    r[i] = ri;
    hr += rhat[i] * ri; rr += ri * ri;
  }
  hr = block_sum(hr, red); if (threadIdx.x == 0) partial[blockIdx.x] = hr;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = rr;
}
__global__ void k_bi_scalars(int it, int G, const double *__restrict__ partial, const double *__restrict__ ts_partial,
                             const double *__restrict__ tt_partial, double *__restrict__ sc, int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  double hr = sum_partials(partial, G, red);
  double rr = sum_partials(partial + G, G, red);
  double ts = sum_partials(ts_partial, G, red);
  double tt = sum_partials(tt_partial, G, red);
  if (threadIdx.x == 0) {
    sc[0] = sc[6];            // rho <- rho_new
    sc[1] = sc[7];            // alpha
    sc[2] = (tt > 0) ? ts / tt : 0.0;
    sc[6] = hr;               // next rho_new = (rhat, r)
    sc[4] = rr;
    flags[1] = it + 1;
    if (!(rr > sc[5] * sc[3])) flags[0] = 1;
    else if (hr * hr <= 1e-28 * sc[8] * rr || sc[2] == 0.0 || !(fabs(sc[6]) < 1e300)) {
      // breakdown: restart from the current residual (rhat <- r in the next k_bi_p)
      sc[0] = 1.0; sc[1] = 1.0; sc[2] = 1.0; sc[6] = rr; sc[8] = rr;
      flags[2] = 1;
      flags[3] += 1;
    }
  }
}

void solve_bicgstab(efv_system *s, double rtol, int maxit, int *iters, double *relres) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  ensure_vectors(s);
  KrylovWork *w = work(s);
  // Partitioned runs: vector kernels cover the owned entries, the ghosts of the SpMV operands are refreshed by a halo
  // exchange, and every set of per-block partials is all-reduced IN PLACE (sum, 0, 0, ...) so that the kernels below,
  // which re-sum G partials, see the global scalars unchanged.
  const int64_t n = s->n_owned * s->ndof;
  const int64_t nalloc = s->rows;
  if (!w->r.n) { w->r.alloc(nalloc); w->p.alloc(nalloc); w->Ap.alloc(nalloc); w->dinv.alloc(nalloc); }
  if (!w->rhat.n) { w->rhat.alloc(nalloc); w->v.alloc(nalloc); w->sv.alloc(nalloc); w->t.alloc(nalloc); }
  // coupled Biot operator with the multigrid preconditioner selected: block-diagonal V-cycles (amg.cu: BlockPrec)
  const bool blockp = s->precond == 1 && s->physics == 2 && s->ndof >= 3;
  if (blockp && !w->phat.n) { w->phat.alloc(nalloc); w->shat.alloc(nalloc); w->phat.zero(); w->shat.zero(); }
  const int G = w->grid;
  double *P = w->partial.p;
  auto global_sums = [&](double *part, int narr) { if (is_dist(s)) comm_allreduce_partials_inplace(part, G, narr); };
  EFV_CUDA(cudaEventRecord(w->e0, stream()));
  sell_prepare(s, w);
  EFV_LAUNCH(k_jacobi, grid_for(n, 256), 256, 0, s->n_owned, s->ndof, s->diag_slot.p, s->vals.p, 1.0, w->dinv.p);
  if (blockp) block_prec_prepare(s);
  EFV_CUDA(cudaMemsetAsync(w->flags.p, 0, 4 * sizeof(int), stream()));
  halo_exchange(s, s->x.p);
  launch_spmv(s, G, s->x.p, plain_epi(w->Ap.p, 1.0, w->dinv.p, nullptr), nullptr, nullptr);
  EFV_LAUNCH(k_bi_init, G, kBlock, 0, n, s->b.p, w->Ap.p, w->dinv.p, w->r.p, w->rhat.p, w->p.p, w->v.p, P);
  global_sums(P, 2);
  EFV_LAUNCH(k_bi_init_scalars, 1, kBlock, 0, G, P, w->sc.p, w->flags.p, rtol);
  const int check_every = 16;
  int it = 0;
  bool done = false;
  while (!done && it < maxit) {
    int batch_end = std::min(maxit, it + check_every);
    for (; it < batch_end; ++it) {
      EFV_LAUNCH(k_bi_p, G, kBlock, 0, n, w->sc.p, w->r.p, w->v.p, w->p.p, w->rhat.p, w->flags.p);
      EFV_LAUNCH(k_bi_clear_restart, 1, 1, 0, w->flags.p);
      // right preconditioning: the products take M^-1 p and M^-1 s, and so does the update of x
      double *pin = w->p.p, *sin = w->sv.p;
      if (blockp) { block_prec_apply(s, w->p.p, w->dinv.p, w->phat.p, w->flags.p); pin = w->phat.p; }
      halo_exchange(s, pin);
      launch_spmv(s, G, pin, plain_epi(w->v.p, 1.0, w->dinv.p, w->rhat.p), P + 2 * G, w->flags.p);            // v = D^-1 A M^-1 p, (rhat,v)
      global_sums(P + 2 * G, 1);
      EFV_LAUNCH(k_bi_s, G, kBlock, 0, n, w->sc.p, P + 2 * G, G, w->r.p, w->v.p, w->sv.p, w->flags.p);
      if (blockp) { block_prec_apply(s, w->sv.p, w->dinv.p, w->shat.p, w->flags.p); sin = w->shat.p; }
      halo_exchange(s, sin);
      launch_spmv(s, G, sin, plain_epi(w->t.p, 1.0, w->dinv.p, w->sv.p), P + 2 * G, w->flags.p);              // t = D^-1 A M^-1 s, (t,s)
      EFV_LAUNCH(k_bi_tt, G, kBlock, 0, n, w->t.p, P + 3 * G, w->flags.p);
      global_sums(P + 2 * G, 2);                                                                    // (t,s) and (t,t)
      EFV_LAUNCH(k_bi_xr, G, kBlock, 0, n, w->sc.p, P + 2 * G, P + 3 * G, G, pin, w->sv.p, sin, w->t.p, w->rhat.p, s->x.p,
                 w->r.p, P, w->flags.p);
      global_sums(P, 2);                                                                            // (rhat,r) and (r,r)
      EFV_LAUNCH(k_bi_scalars, 1, kBlock, 0, it, G, P, P + 2 * G, P + 3 * G, w->sc.p, w->flags.p);
    }
    EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    done = w->h_flags[0] != 0;
  }
  EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaMemcpyAsync(w->h_sc, w->sc.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaEventRecord(w->e1, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  float ms = 0;
  EFV_CUDA(cudaEventElapsedTime(&ms, w->e0, w->e1));
  s->t_solve_ms = ms;
  if (is_dist(s) && comm_peer_active()) EFV_REQUIRE(comm_peer_error() == 0, "peer-memory collective timed out (a rank did not arrive)");
  if (iters) *iters = w->h_flags[1];
  if (relres) *relres = (w->h_sc[3] > 0) ? std::sqrt(w->h_sc[4] / w->h_sc[3]) : 0.0;
}

__global__ void k_max_final(int G, const double *__restrict__ partial, double *__restrict__ out);
// ---- true residual ||b - A x|| / ||b|| with one extra SpMV (never the recurrence residual), checksums -------------------------
__global__ void __launch_bounds__(kBlock) k_resid_norms(int64_t n, const double *__restrict__ b, const double *__restrict__ y,
                                                        double sigma, double *__restrict__ partial) {
  __shared__ double red[32];
  double rr = 0, bb = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = sigma * b[i], ri = bi - y[i];
    rr += ri * ri; bb += bi * bi;
  }
  const int G = gridDim.x;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[blockIdx.x] = rr;
  bb = block_sum(bb, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = bb;
}
__global__ void __launch_bounds__(kBlock) k_sums(int64_t n, const double *__restrict__ a, double *__restrict__ partial) {
  __shared__ double red[32];
  double s1 = 0, s2 = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = a[i];
    s1 += v; s2 += v * v;
  }
  const int G = gridDim.x;
  s1 = block_sum(s1, red); if (threadIdx.x == 0) partial[blockIdx.x] = s1;
  s2 = block_sum(s2, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = s2;
}
// two global sums (over the owned entries of all ranks) from two arrays of G per-block partials
static void two_global_sums(efv_system *s, double *out2) {
  KrylovWork *w = s->kw;
  double *dev = w->sc.p + 48;
  EFV_LAUNCH(k_reduce_partials, 1, kBlock, 0, w->partial.p, w->grid, 2, dev);
  if (is_dist(s)) comm_allreduce_sum(dev, 2);
  EFV_CUDA(cudaMemcpyAsync(out2, dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
}
double true_relres(efv_system *s, int negate) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  ensure_vectors(s);
  KrylovWork *w = work(s);
  if (!w->Ap.n) { w->r.alloc(s->rows); w->p.alloc(s->rows); w->Ap.alloc(s->rows); w->dinv.alloc(s->rows); }
  const double sigma = negate ? -1.0 : 1.0;
  sell_prepare(s, w);
  halo_exchange(s, s->x.p);
  launch_spmv(s, w->grid, s->x.p, plain_epi(w->Ap.p, sigma, nullptr, nullptr), nullptr, nullptr);
  EFV_LAUNCH(k_resid_norms, w->grid, kBlock, 0, s->n_owned * s->ndof, s->b.p, w->Ap.p, sigma, w->partial.p);
  double h[2];
  two_global_sums(s, h);
  return h[1] > 0 ? std::sqrt(h[0] / h[1]) : std::sqrt(h[0]);
}
void vec_sums(efv_system *s, const double *vec, double *sum, double *sumsq) {
  KrylovWork *w = work(s);
  EFV_LAUNCH(k_sums, w->grid, kBlock, 0, s->n_owned * s->ndof, vec, w->partial.p);
  double h[2];
  two_global_sums(s, h);
  if (sum) *sum = h[0];
  if (sumsq) *sumsq = h[1];
}
// max_ij |a_ij - a_ji| / max_ij |a_ij| over the owned scalar entries (column inside the owned block), by binary search
// of the transposed entry: tells the apps at assembly time whether CG applies (general quadrilaterals / hexahedra give
// a non-symmetric EbFVM operator)
__global__ void __launch_bounds__(kBlock) k_symmetry(int64_t n_owned, int ndof, const int64_t *__restrict__ gptr,
                                                     const int32_t *__restrict__ gcol, const double *__restrict__ vals,
                                                     double *__restrict__ partial) {
  __shared__ double red[32];
  double dmax = 0.0, amax = 0.0;
  const int B = ndof * ndof;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_owned; v += (int64_t)gridDim.x * blockDim.x) {
    for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) {
      const int64_t c = gcol[t];
      for (int q = 0; q < B; ++q) amax = fmax(amax, fabs(vals[t * B + q]));
      if (c >= n_owned || c <= v) continue;
      int64_t lo = gptr[c], hi = gptr[c + 1] - 1, pos = -1;
      while (lo <= hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t cm = gcol[mid];
        if (cm == v) { pos = mid; break; }
        if (cm < v) lo = mid + 1; else hi = mid - 1;
      }
      for (int i = 0; i < ndof; ++i)
        for (int j = 0; j < ndof; ++j) {
          const double a = vals[t * B + i * ndof + j];
          const double at = pos >= 0 ? vals[pos * B + j * ndof + i] : 0.0;
          dmax = fmax(dmax, fabs(a - at));
        }
    }
  }
  const int G = gridDim.x;
  dmax = block_max(dmax, red); if (threadIdx.x == 0) partial[blockIdx.x] = dmax;
  amax = block_max(amax, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = amax;
}
double symmetry_defect(efv_system *s) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  KrylovWork *w = work(s);
  EFV_LAUNCH(k_symmetry, w->grid, kBlock, 0, s->n_owned, s->ndof, s->gptr.p, s->gcol.p, s->vals.p, w->partial.p);
  EFV_LAUNCH(k_max_final, 1, kBlock, 0, w->grid, w->partial.p, w->sc.p + 48);
  EFV_LAUNCH(k_max_final, 1, kBlock, 0, w->grid, w->partial.p + w->grid, w->sc.p + 49);
  if (is_dist(s)) comm_allreduce_max(w->sc.p + 48, 2);
  double h[2];
  EFV_CUDA(cudaMemcpyAsync(h, w->sc.p + 48, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  return h[1] > 0 ? h[0] / h[1] : 0.0;
}

// ---- max |x - xold| (heat_transfer.py:142) ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_max_abs_diff(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                                                         double *__restrict__ partial) {
  __shared__ double red[32];
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs(a[i] - b[i]));
  m = block_max(m, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = m;
}
__global__ void k_max_final(int G, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) m = fmax(m, partial[i]);
  m = block_max(m, red);
  if (threadIdx.x == 0) *out = m;
}
double max_abs_diff(efv_system *s) {
  ensure_vectors(s);
  KrylovWork *w = work(s);
  EFV_LAUNCH(k_max_abs_diff, w->grid, kBlock, 0, s->n_owned * s->ndof, s->x.p, s->xold.p, w->partial.p);
  EFV_LAUNCH(k_max_final, 1, kBlock, 0, w->grid, w->partial.p, w->sc.p + 15);
  if (is_dist(s)) comm_allreduce_max(w->sc.p + 15, 1);
  double out = 0;
  EFV_CUDA(cudaMemcpyAsync(&out, w->sc.p + 15, sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  return out;
}

}  // namespace efv
