// krylov.cuh -- pieces of the Krylov layer shared between krylov.cu (SpMV kernels, PCG, BiCGStab) and amg.cu (the
// aggregation multigrid preconditioner, which is built from the same SpMV kernels through their fused epilogues).
#pragma once
#include "views.cuh"
#include "peer.cuh"

namespace efv {

constexpr int kBlock = 256;

struct KrylovWork {
  int64_t n = 0;
  int grid = 0;
  DBuf<double> r, p, Ap, dinv, rhat, v, sv, t, phat, shat;
  DBuf<double> partial;     // 4 * grid
  DBuf<double> sc;          // device scalars
  DBuf<int> flags;          // [0] done, [1] iterations, [2] breakdown
  // sliced-ELL copy of a scalar matrix for the solver (slices of 32 rows, column-major inside a slice)
  DBuf<int64_t> sell_off;   // nslices+1, in units of 32-entry columns
  DBuf<int32_t> sell_cols;
  DBuf<int32_t> sell_delta; // one per 32-entry column: common (column - row) offset, or kNoDelta
  DBuf<double> sell_vals;
  int64_t sell_slices = 0, sell_rows = 0;
  uint64_t sell_version = 0;
  bool sell_ok = false, sell_use = false;
  bool sell_indexed = false;  // columns / offsets of the current pattern are in place (values-only refills)
  int sell_compress = -1;
  int64_t int_lo = 0, int_hi = 0, int_rows = -1;   // partitioned systems: longest run of slices without ghost columns
  int *h_flags = nullptr;   // pinned
  double *h_sc = nullptr;   // pinned
  cudaEvent_t ev = nullptr, e0 = nullptr, e1 = nullptr;
};

// fixed-order sum of `count` partials, identical in every block
__device__ __forceinline__ double sum_partials(const double *__restrict__ part, int count, double *red) {
  double a = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) a += part[i];
  a = block_sum(a, red);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = a;
  __syncthreads();
  double r = bc;
  __syncthreads();
  return r;
}

enum { kEpiPlain = 0, kEpiResid = 1, kEpiChebFirst = 2, kEpiChebMid = 3, kEpiChebLast = 4 };
enum { kEpiXZero = 1, kEpiRFromB = 2 };
// what a SpMV kernel does with (A v)_i at the end of a row: see epi_apply in krylov.cu
struct SpmvEpi {
  int mode = kEpiPlain, flags = 0;
  int prefetch = 1;                    // TMA kernel: load the row-wise operands at the start of a slice (EFV_EPI_PREFETCH=0: at its end)
  double sigma = 1.0;
  double *y = nullptr;
  const double *rowscale = nullptr, *w = nullptr;
  const double *b = nullptr, *dinv = nullptr, *r_in = nullptr, *d_in = nullptr, *coef = nullptr;
  double *r_out = nullptr, *d_out = nullptr, *x = nullptr;
};
SpmvEpi plain_epi(double *y, double sigma, const double *rowscale, const double *w);

KrylovWork *work(efv_system *s);
bool sell_prepare(efv_system *s, KrylovWork *w, double *rowabs = nullptr, bool *rowabs_done = nullptr);
bool is_dist(const efv_system *s);
const double *consolidate(efv_system *s, const double *partial, int narr, int slot, int *Gout);
// y-epilogue(A x) on the owned rows of s (SELL copy when prepared, CSR otherwise); partial: per-block partials of the
// epilogue's dot product (grid entries), done: device flag that turns the launch into a no-op
void launch_spmv(efv_system *s, int grid, const double *x, const SpmvEpi &epi, double *partial, const int *done,
                 const PeerAR &ar = PeerAR(), const HaloIn *hin = nullptr);
void spmv_halo(efv_system *s, int grid, double *x, const SpmvEpi &epi, double *partial, const int *done);

}  // namespace efv
