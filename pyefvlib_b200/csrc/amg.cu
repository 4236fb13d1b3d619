// amg.cu -- aggregation multigrid preconditioner for the CG solver (replaces the Jacobi diagonal where the iteration
// count, not the bandwidth, decides the time to solution: SURVEY.md section 7.2-1; the reference inverts the matrix
// instead, apps/heat_transfer.py:78-81).
//
// Hierarchy.  Level 0 is the assembled system.  Level l+1 is the Galerkin operator of PLAIN aggregation,
// A_c[I][J] = sum of a_ij over i in aggregate I, j in aggregate J (blocks are summed entry-wise for ndof > 1, i.e. the
// coarse space of a vertex aggregate is one value per DOF).  Aggregates hold ~8 vertices: 2x2x2 blocks computed
// arithmetically when the system knows that its vertices are numbered like a lattice (efv_system_set_lattice; the
// synthetic meshes and their slabs), otherwise three rounds of pairwise matching on the graph (host, once per pattern).
// Aggregates never cross ranks, so the transfer operators and the Galerkin sums need no communication.
//
// Everything structural is computed once per pattern (symbolic): members of every aggregate, the sorted coarse CSR
// pattern, and for every coarse entry the ascending list of fine entries that sum into it.  The numeric set-up that
// runs after every assembly is then one gather-sum kernel per level (fixed order: bit-reproducible, no float atomics),
// the inverse diagonal, the Gershgorin bound of D^-1 A (the Chebyshev interval) and a dense inverse of the coarsest
// operator (<= 256 rows).
//
// Cycle.  V(2,2) with Chebyshev smoothing on [lmax/6, lmax] and an over-corrected coarse-grid update (x += s P e_c,
// s = 1.8 for scalar problems: plain aggregation under-estimates the correction, Braess 1995).  Every smoothing step
// and every residual is ONE SpMV launch of krylov.cu with a fused epilogue (SpmvEpi): 4 matrix passes per level and
// cycle + 3 short vector kernels (first direction, restriction, prolongation).  The cycle is a fixed symmetric
// linear operator, so plain CG applies.  Measured iteration counts are in DESIGN.md.
#include "krylov.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>

namespace efv {

// EFV_AMG_TIMING=1: CUDA events at the phase boundaries of the first iterations of a solve, printed per label after the solve
// (diagnostics; the events serialise nothing, they only cost their own recording)
struct PhaseTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> label;
  size_t used = 0;
  void mark(const std::string &name) {
    if (!on || used >= 4000) return;
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); label.push_back(name); }
    label[used] = name;
    cudaEventRecord(ev[used++], stream());
  }
  void report(int rank, long long rows) {
    if (!on || used < 2) { used = 0; return; }
    cudaEventSynchronize(ev[used - 1]);
    std::vector<std::string> order;
    std::vector<double> sum;
    std::vector<int> cnt;
    for (size_t i = 1; i < used; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      size_t k = 0;
      for (; k < order.size(); ++k) if (order[k] == label[i]) break;
      if (k == order.size()) { order.push_back(label[i]); sum.push_back(0); cnt.push_back(0); }
      sum[k] += ms; cnt[k]++;
    }
    fprintf(stderr, "[efv amg timing rank %d rows %lld]", rank, rows);
    for (size_t k = 0; k < order.size(); ++k) fprintf(stderr, " %s=%.3fms/%d", order[k].c_str(), sum[k], cnt[k]);
    fprintf(stderr, "\n");
    used = 0;
  }
};
static PhaseTimer g_pt;
#define EFV_PT(name) g_pt.mark(name)

constexpr int kMaxDeg = 4;
constexpr int kCoarsestRows = 256;     // scalar rows of the dense coarsest solve
constexpr int kMaxLevels = 24;

struct AmgLevel {
  efv_system *sys = nullptr;           // level 0: the caller's system; coarser levels: pattern + values owned here
  bool owns_sys = false;
  int64_t n = 0, ng = 0;               // owned / ghost block rows
  int lat[3] = {0, 0, 0};              // lattice extents of the owned rows (0: unknown)
  DBuf<double> dinv, x, b, r, d0, d1, res;
  DBuf<double> coef;                   // kMaxDeg steps x (1/theta, c1, c2, -) followed by (lmax, -, -, -)
  DBuf<double> ev;                     // running estimate of the dominant eigenvector of D^-1 A (warm start of the power iteration)
  bool ev_ready = false;
  // transfer to the next level
  int64_t nc = 0;
  DBuf<int32_t> agg;                   // aggregate (coarse row) of every owned row
  DBuf<int64_t> mem_ptr;               // members of every aggregate, ascending
  DBuf<int32_t> mem;
  DBuf<int64_t> cl_ptr;                // per coarse entry: the fine entries summed into it, ascending
  DBuf<int32_t> cl;
  // partitioned runs: for every ghost row of THIS level its owner rank and its row number there (levels >= 1)
  std::vector<int> ghost_rank;
  std::vector<int64_t> ghost_owner_local;
  bool gathered = false;               // partitioned level whose operator is replicated on every rank as the NEXT level of the list
};

struct AmgHierarchy {
  std::vector<AmgLevel *> lv;
  uint64_t numeric_version = 0;
  double numeric_sigma = 0.0;
  int deg = 2;
  int power_steps = 2;                 // power-iteration steps per numeric set-up (x6 the first time); 0: Gershgorin bound only
  double ratio = 6.0, scale = 1.8;
  int64_t nd = 0;                      // scalar rows of the coarsest level (all ranks together)
  DBuf<double> dense;                  // nd x 2 nd work matrix; the inverse is its right half
  // partitioned runs: the coarsest operator is gathered on every rank and inverted redundantly
  bool dist = false, dense_dist = false;
  int64_t nd_off = 0, nd_local = 0;    // first scalar row / scalar rows of this rank inside the gathered coarsest system
  GatherStage *stage = nullptr;
  DBuf<int32_t> colmap;                // coarsest level: local column -> global row of the gathered system
  DBuf<double> bfull, dense_local;
  // early gather: below gather_rows rows (all ranks together) the operator of a partitioned level is replicated on every
  // rank and the rest of the hierarchy runs redundantly without communication (tiny levels are pure halo latency otherwise)
  int64_t gather_rows = 262144;
  GatherStage *rep_stage = nullptr;
  int64_t rep_row_off = 0, rep_rows = 0, rep_nnz_off = 0, rep_nnz = 0, rep_total_rows = 0, rep_total_nnz = 0;
  double t_numeric_ms = 0.0;
  bool numeric_timed = false;
  // one CG iteration (product, update, check, V-cycle, direction) captured as a CUDA graph: ~60 launches, most of them
  // a few microseconds long on the small levels, replayed with one launch (single-GPU systems only: the peer collectives
  // carry per-call sequence numbers)
  cudaGraphExec_t iter_graph = nullptr;
  double graph_sigma = 0.0;
  int graph_nodes = 0;
  ~AmgHierarchy() {
    for (AmgLevel *L : lv) {
      if (L->owns_sys && L->sys) { krylov_free(L->sys); delete L->sys; }
      delete L;
    }
    if (stage) gather_stage_destroy(stage);
    if (rep_stage) gather_stage_destroy(rep_stage);
    if (iter_graph) cudaGraphExecDestroy(iter_graph);
  }
};

// ---- symbolic: aggregates -----------------------------------------------------------------------------------------------
__global__ void k_agg_lattice(int64_t n, int nx, int ny, int cx, int cy, int32_t *__restrict__ agg) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(v % nx), j = (int)((v / nx) % ny);
    const int64_t k = v / ((int64_t)nx * ny);
    agg[v] = (int32_t)((i >> 1) + (int64_t)cx * ((j >> 1) + (int64_t)cy * (k >> 1)));
  }
}

// pairwise matching on the host (general meshes; once per pattern): `passes` rounds, every round pairs each unmatched
// vertex with its unmatched neighbour of smallest degree and then contracts the graph
static std::vector<int32_t> host_aggregate(int64_t n, const std::vector<int64_t> &ptr0, const std::vector<int32_t> &col0, int passes,
                                           int64_t *n_coarse) {
  std::vector<int32_t> map(n);
  std::iota(map.begin(), map.end(), 0);
  std::vector<int64_t> ptr = ptr0;
  std::vector<int32_t> col = col0;
  int64_t m = n;
  for (int pass = 0; pass < passes && m > 1; ++pass) {
    std::vector<int32_t> id(m, -1);
    int32_t next = 0;
    for (int64_t v = 0; v < m; ++v) {
      if (id[v] >= 0) continue;
      int64_t best = -1, best_deg = INT64_MAX;
      for (int64_t t = ptr[v]; t < ptr[v + 1]; ++t) {
        const int64_t u = col[t];
        if (u == v || u >= m || id[u] >= 0) continue;
        const int64_t deg = ptr[u + 1] - ptr[u];
        if (deg < best_deg) { best_deg = deg; best = u; }
      }
      id[v] = next;
      if (best >= 0) id[best] = next;
      ++next;
    }
    // contract
    std::vector<std::vector<int32_t>> rows(next);
    for (int64_t v = 0; v < m; ++v)
      for (int64_t t = ptr[v]; t < ptr[v + 1]; ++t)
        if (col[t] < m) rows[id[v]].push_back(id[col[t]]);
    std::vector<int64_t> nptr(next + 1, 0);
    std::vector<int32_t> ncol;
    for (int32_t a = 0; a < next; ++a) {
      auto &r = rows[a];
      std::sort(r.begin(), r.end());
      r.erase(std::unique(r.begin(), r.end()), r.end());
      nptr[a + 1] = nptr[a] + (int64_t)r.size();
      ncol.insert(ncol.end(), r.begin(), r.end());
    }
    for (int64_t v = 0; v < n; ++v) map[v] = id[map[v]];
    ptr.swap(nptr);
    col.swap(ncol);
    m = next;
  }
  *n_coarse = m;
  return map;
}

// ---- symbolic: members, coarse pattern, gather lists ------------------------------------------------------------------------
__global__ void k_count_members(int64_t n, const int32_t *__restrict__ agg, int32_t *__restrict__ cnt) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) atomicAdd(cnt + agg[v], 1);
}
__global__ void k_fill_members(int64_t n, const int32_t *__restrict__ agg, const int64_t *__restrict__ ptr, int32_t *__restrict__ cursor,
                               int32_t *__restrict__ mem) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t a = agg[v];
    mem[ptr[a] + atomicAdd(cursor + a, 1)] = (int32_t)v;
  }
}
// ascending order inside every list (the fill order above depends on the schedule; the sorted lists do not)
__global__ void k_sort_lists(int64_t nlists, const int64_t *__restrict__ ptr, int32_t *__restrict__ items) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nlists; a += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = ptr[a], p1 = ptr[a + 1];
    for (int64_t i = p0 + 1; i < p1; ++i) {
      const int32_t key = items[i];
      int64_t j = i - 1;
      while (j >= p0 && items[j] > key) { items[j + 1] = items[j]; --j; }
      items[j + 1] = key;
    }
  }
}

__global__ void k_i32_to_f64(int64_t n, const int32_t *__restrict__ a, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)a[i];
}

// replication of a partitioned level: row lengths and columns (renumbered to the gathered numbering) as doubles for the
// gather stage, and back
__global__ void k_pattern_to_f64(int64_t n, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                                 const int32_t *__restrict__ colmap, double *__restrict__ len_out, double *__restrict__ col_out) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    len_out[v] = (double)(gptr[v + 1] - gptr[v]);
    for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) col_out[t] = (double)colmap[gcol[t]];
  }
}
__global__ void k_f64_to_i32(int64_t n, const double *__restrict__ a, int32_t *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)a[i];
}
__global__ void k_find_diag_slot(int64_t n, const int64_t *__restrict__ ptr, const int32_t *__restrict__ col, int32_t *__restrict__ diag) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    int32_t found = (int32_t)ptr[v];
    for (int64_t t = ptr[v]; t < ptr[v + 1]; ++t)
      if (col[t] == v) found = (int32_t)t;
    diag[v] = found;
  }
}
__global__ void __launch_bounds__(256) k_copy_slice(int64_t n, const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

constexpr int kCWarps = 4;
constexpr int kCMaxCand = 2048;
// one warp per coarse row: candidates = aggregates of the columns of the member rows; distinct values in ascending order
template <bool FILL>
__global__ void __launch_bounds__(kCWarps * 32) k_coarse_rows(int64_t nc, int64_t n_owned_fine, const int64_t *__restrict__ mem_ptr,
                                                              const int32_t *__restrict__ mem, const int64_t *__restrict__ gptr,
                                                              const int32_t *__restrict__ gcol, const int32_t *__restrict__ agg_all,
                                                              int32_t *__restrict__ rowlen, const int64_t *__restrict__ cptr,
                                                              int32_t *__restrict__ ccol, int32_t *__restrict__ cdiag,
                                                              int *__restrict__ error_flag) {
  __shared__ int32_t cand_s[kCWarps][kCMaxCand];
  __shared__ uint8_t first_s[kCWarps][kCMaxCand];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t *cand = cand_s[warp];
  uint8_t *first = first_s[warp];
  for (int64_t I = (int64_t)blockIdx.x * kCWarps + warp; I < nc; I += (int64_t)gridDim.x * kCWarps) {
    int n = 0;
    bool overflow = false;
    for (int64_t q = mem_ptr[I]; q < mem_ptr[I + 1] && !overflow; ++q) {
      const int64_t v = mem[q];
      const int64_t r0 = gptr[v];
      const int len = (int)(gptr[v + 1] - r0);
      if (n + len > kCMaxCand) { overflow = true; break; }
      for (int k = lane; k < len; k += 32) cand[n + k] = agg_all[gcol[r0 + k]];
      n += len;
    }
    if (overflow) { if (lane == 0) atomicExch(error_flag, 1); continue; }
    __syncwarp();
    int count = 0;
    for (int i = lane; i < n; i += 32) {
      const int32_t c = cand[i];
      bool is_first = true;
      for (int t = 0; t < i; ++t) is_first &= (cand[t] != c);
      first[i] = is_first;
      count += is_first;
    }
    __syncwarp();
    if (!FILL) {
      for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
      if (lane == 0) rowlen[I] = count;
    } else {
      for (int i = lane; i < n; i += 32) {
        if (!first[i]) continue;
        const int32_t c = cand[i];
        int rank = 0;
        for (int t = 0; t < n; ++t) rank += (first[t] && cand[t] < c);
        ccol[cptr[I] + rank] = c;
        if (c == (int32_t)I) cdiag[I] = (int32_t)(cptr[I] + rank);
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ int64_t coarse_slot(const int64_t *__restrict__ cptr, const int32_t *__restrict__ ccol, int32_t I, int32_t J) {
  int64_t lo = cptr[I], hi = cptr[I + 1] - 1;
  while (lo <= hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t c = ccol[mid];
    if (c == J) return mid;
    if (c < J) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
// pass 0: count the fine entries of every coarse entry; pass 1: fill the lists through a cursor
template <int PASS>
__global__ void k_gather_lists(int64_t n_owned, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                               const int32_t *__restrict__ agg_all, const int64_t *__restrict__ cptr, const int32_t *__restrict__ ccol,
                               int32_t *__restrict__ cnt, const int64_t *__restrict__ cl_ptr, int32_t *__restrict__ cl,
                               int *__restrict__ error_flag) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_owned; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t I = agg_all[v];
    for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) {
      const int64_t e = coarse_slot(cptr, ccol, I, agg_all[gcol[t]]);
      if (e < 0) { atomicExch(error_flag, 2); continue; }
      if (PASS == 0) atomicAdd(cnt + e, 1);
      else cl[cl_ptr[e] + atomicAdd(cnt + e, 1)] = (int32_t)t;
    }
  }
}

// ---- numeric ------------------------------------------------------------------------------------------------------------------
// coarse value = sum of its fine values in ascending entry order (one thread per coarse entry and block component)
__global__ void k_galerkin(int64_t cnnz, int B, const int64_t *__restrict__ cl_ptr, const int32_t *__restrict__ cl,
                           const double *__restrict__ fine, double *__restrict__ coarse) {
  const int64_t total = cnnz * B;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / B;
    const int q = (int)(i % B);
    double acc = 0.0;
    for (int64_t k = cl_ptr[e]; k < cl_ptr[e + 1]; ++k) acc += fine[(int64_t)cl[k] * B + q];
    coarse[i] = acc;
  }
}
// dinv = 1 / (sigma a_ii) per scalar row and the Gershgorin bound max_i sum_j |a_ij| / |a_ii| of D^-1 A: a group of 8
// lanes per block row
__global__ void __launch_bounds__(256) k_diag_gershgorin(int64_t n_owned, int ndof, const int64_t *__restrict__ gptr,
                                                         const int32_t *__restrict__ gcol, const double *__restrict__ vals,
                                                         const int32_t *__restrict__ diag_slot, double sigma,
                                                         double *__restrict__ dinv, double *__restrict__ partial) {
  __shared__ double red[32];
  const int gl = threadIdx.x & 7;
  const int B = ndof * ndof;
  double worst = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < n_owned; base += (int64_t)gridDim.x * 32) {
    const int64_t v = base + (threadIdx.x >> 3);
    const bool valid = v < n_owned;
    const int64_t r0 = valid ? gptr[v] : 0, r1 = valid ? gptr[v + 1] : 0;
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t t = r0 + gl; t < r1; t += 8)
      for (int i = 0; i < ndof; ++i)
        for (int j = 0; j < ndof; ++j) sum[i] += fabs(vals[t * B + i * ndof + j]);
    for (int i = 0; i < ndof; ++i) {
      double sv = sum[i];
      for (int o = 4; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (valid && gl == 0) {
        const double d = vals[(int64_t)diag_slot[v] * B + i * ndof + i];
        dinv[v * ndof + i] = (d != 0.0) ? 1.0 / (sigma * d) : 1.0;
        if (d != 0.0) worst = fmax(worst, sv / fabs(d));
      }
    }
  }
  worst = block_max(worst, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = worst;
}
// power iteration on D^-1 A: deterministic start vector, (y.y, v.v) partials, v <- y / |y|
__global__ void k_pow_init(int64_t n, double *__restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ 0x9e3779b9u;
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
    v[i] = (((i & 1) ? -1.0 : 1.0)) * (0.5 + (double)(h & 0xffffu) / 65536.0);
  }
}
__global__ void __launch_bounds__(256) k_pow_dots(int64_t n, const double *__restrict__ y, const double *__restrict__ v,
                                                  double *__restrict__ partial) {
  __shared__ double red[32];
  double yy = 0, vv = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    yy += y[i] * y[i]; vv += v[i] * v[i];
  }
  const int G = gridDim.x;
  yy = block_sum(yy, red); if (threadIdx.x == 0) partial[blockIdx.x] = yy;
  vv = block_sum(vv, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = vv;
}
__global__ void __launch_bounds__(256) k_pow_scale(int64_t n, const double *__restrict__ y, const double *__restrict__ dots, int G,
                                                   double *__restrict__ v) {
  __shared__ double red[32];
  const double yy = sum_partials(dots, G, red);
  const double sc = yy > 0.0 ? rsqrt(yy) : 1.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = y[i] * sc;
}
// the same two results from the row sums that the SELL fill pass left behind (no second pass over the matrix)
__global__ void __launch_bounds__(256) k_diag_from_rowabs(int64_t n_owned, int ndof, const double *__restrict__ vals,
                                                          const int32_t *__restrict__ diag_slot, const double *__restrict__ rowabs,
                                                          double sigma, double *__restrict__ dinv, double *__restrict__ partial) {
  __shared__ double red[32];
  const int B = ndof * ndof;
  const int64_t n = n_owned * ndof;
  double worst = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / ndof;
    const int f = (int)(i % ndof);
    const double d = vals[(int64_t)diag_slot[v] * B + f * ndof + f];
    dinv[i] = (d != 0.0) ? 1.0 / (sigma * d) : 1.0;
    if (d != 0.0) worst = fmax(worst, rowabs[i] / fabs(d));
  }
  worst = block_max(worst, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = worst;
}
// Chebyshev coefficients of the interval [lmax / ratio, lmax] (one thread; partial: per-block maxima of the bound)
// lmax = min(Gershgorin bound, 1.1 x power-iteration estimate sqrt(y.y / v.v)) (pow_dots == nullptr: the bound alone)
__global__ void k_cheb_coef(const double *__restrict__ partial, int G, double ratio, const double *__restrict__ pow_dots, int Gp,
                            double *__restrict__ coef) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) m = fmax(m, partial[i]);
  m = block_max(m, red);
  __shared__ double bound;
  if (threadIdx.x == 0) bound = m;
  __syncthreads();
  double est = 0.0;
  if (pow_dots) {
    const double yy = sum_partials(pow_dots, Gp, red);
    const double vv = sum_partials(pow_dots + Gp, Gp, red);
    est = (vv > 0.0 && yy > 0.0) ? 1.1 * sqrt(yy / vv) : 0.0;
  }
  if (threadIdx.x == 0) {
    m = bound;
    double lmax = (m > 0.0) ? m : 1.0;
    if (est > 0.0 && est < lmax) lmax = est;
    const double lmin = lmax / ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
    double rho = 1.0 / sigma1;
    coef[0] = 1.0 / theta; coef[1] = 0.0; coef[2] = 0.0; coef[3] = 0.0;
    for (int k = 1; k < kMaxDeg; ++k) {
      const double rho_new = 1.0 / (2.0 * sigma1 - rho);
      coef[4 * k + 0] = 1.0 / theta;
      coef[4 * k + 1] = rho_new * rho;
      coef[4 * k + 2] = 2.0 * rho_new / delta;
      coef[4 * k + 3] = 0.0;
      rho = rho_new;
    }
    coef[4 * kMaxDeg] = lmax;
  }
}
// dense copy [sigma A | I] of the coarsest operator (rows [row0, row0 + nrows) of the N x 2N system; M holds those rows
// only), then Gauss-Jordan without pivoting (SPD) in one block
__global__ void k_dense_fill(int64_t nrows, int64_t row0, int64_t N, double *__restrict__ M) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows * 2 * N; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row0 + i / (2 * N), c = i % (2 * N);
    M[i] = (c == r + N) ? 1.0 : 0.0;
  }
}
// colmap: local block column -> block row of the gathered system (nullptr: identity, ghosts skipped)
__global__ void k_dense_entries(int64_t n, int ndof, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                                const double *__restrict__ vals, double sigma, const int32_t *__restrict__ colmap, int64_t row0, int64_t N,
                                double *__restrict__ M) {
  const int B = ndof * ndof;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) {
      int64_t c = gcol[t];
      if (colmap) c = colmap[c];
      else if (c >= n) continue;
      for (int i = 0; i < ndof; ++i)
        for (int j = 0; j < ndof; ++j) M[(v * ndof + i) * 2 * N + c * ndof + j] = sigma * vals[t * B + i * ndof + j];
    }
}
__global__ void __launch_bounds__(1024) k_gauss_jordan(int N, double *__restrict__ M) {
  __shared__ double col[kCoarsestRows];
  __shared__ double pivot_inv;
  const int W = 2 * N;
  for (int k = 0; k < N; ++k) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) col[i] = M[(int64_t)i * W + k];
    if (threadIdx.x == 0) {
      double p = M[(int64_t)k * W + k];
      if (!(fabs(p) > 1e-300)) p = 1.0;            // singular coarse operator (pure Neumann): keep the row as it is
      pivot_inv = 1.0 / p;
    }
    __syncthreads();
    const double pinv = pivot_inv;
    for (int j = threadIdx.x; j < W; j += blockDim.x) M[(int64_t)k * W + j] *= pinv;
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * W; idx += blockDim.x) {
      const int i = idx / W, j = idx % W;
      if (i != k) M[(int64_t)i * W + j] -= col[i] * M[(int64_t)k * W + j];
    }
    __syncthreads();
  }
}
// x[0, nrows) = rows [row0, row0 + nrows) of the inverse times the full right-hand side b[0, N)
__global__ void __launch_bounds__(256) k_dense_solve(int N, int row0, int nrows, const double *__restrict__ M, const double *__restrict__ b,
                                                     double *__restrict__ x, const int *__restrict__ done) {
  if (done && *done) return;
  __shared__ double bs[kCoarsestRows];
  for (int i = threadIdx.x; i < N; i += blockDim.x) bs[i] = b[i];
  __syncthreads();
  const int W = 2 * N;
  for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < N; ++j) acc += M[(int64_t)(row0 + i) * W + N + j] * bs[j];
    x[i] = acc;
  }
}

// ---- cycle kernels ----------------------------------------------------------------------------------------------------------------
// first Chebyshev direction from a zero initial guess: d0 = dinv b / theta (deg == 1: that is already the smoothed x)
__global__ void __launch_bounds__(256) k_first_direction(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv,
                                                         const double *__restrict__ coef, double *__restrict__ d0,
                                                         const int *__restrict__ done) {
  if (done && *done) return;
  const double it = coef[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d0[i] = dinv[i] * b[i] * it;
}
// b_c[I] = sum of res over the members of I (ascending member order); optionally also the first direction of the coarse level
__global__ void __launch_bounds__(256) k_restrict(int64_t nc, int ndof, const int64_t *__restrict__ mem_ptr, const int32_t *__restrict__ mem,
                                                  const double *__restrict__ res, double *__restrict__ bc,
                                                  const double *__restrict__ dinv_c, const double *__restrict__ coef_c,
                                                  double *__restrict__ d0_c, const int *__restrict__ done) {
  if (done && *done) return;
  const int64_t total = nc * ndof;
  const double it = d0_c ? coef_c[0] : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = i / ndof;
    const int f = (int)(i % ndof);
    double acc = 0.0;
    for (int64_t q = mem_ptr[I]; q < mem_ptr[I + 1]; ++q) acc += res[(int64_t)mem[q] * ndof + f];
    bc[i] = acc;
    if (d0_c) d0_c[i] = dinv_c[i] * acc * it;
  }
}
// x += scale * x_c[aggregate]
__global__ void __launch_bounds__(256) k_prolong(int64_t n, int ndof, const int32_t *__restrict__ agg, const double *__restrict__ xc,
                                                 double scale, double *__restrict__ x, const int *__restrict__ done) {
  if (done && *done) return;
  const int64_t total = n * ndof;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / ndof;
    const int f = (int)(i % ndof);
    x[i] += scale * xc[(int64_t)agg[v] * ndof + f];
  }
}
// b . x of a short vector in one block: partial[0] = sum, partial[1..G) = 0 (the consumers sum G entries)
__global__ void __launch_bounds__(256) k_dot_partial(int64_t n, const double *__restrict__ b, const double *__restrict__ x,
                                                     double *__restrict__ partial, int G, const int *__restrict__ done) {
  __shared__ double red[32];
  if (done && *done) return;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += b[i] * x[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[0] = acc;
  for (int i = 1 + threadIdx.x; i < G; i += blockDim.x) partial[i] = 0.0;
}
__global__ void __launch_bounds__(256) k_axpy1(int64_t n, const double *__restrict__ d, double *__restrict__ x, const int *__restrict__ done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += d[i];
}

// ---- set-up ----------------------------------------------------------------------------------------------------------------------------
static void level_vectors(AmgLevel *L, bool top) {
  const size_t len = (size_t)(L->n + L->ng) * L->sys->ndof;
  L->dinv.alloc(len); L->r.alloc(len); L->d0.alloc(len); L->d1.alloc(len); L->res.alloc(len);
  L->d0.zero(); L->d1.zero(); L->res.zero();
  if (!top) { L->x.alloc(len); L->b.alloc(len); L->x.zero(); L->b.zero(); }
  L->coef.alloc(4 * (kMaxDeg + 1));
}

static void build_transfer(AmgLevel *F, AmgLevel *C) {
  efv_system *fs = F->sys;
  const int64_t n = F->n;
  // 1. aggregates of the owned rows
  F->agg.alloc(n + F->ng);
  int64_t nc = 0;
  if (F->lat[0] > 0) {
    const int nx = F->lat[0], ny = F->lat[1], nz = F->lat[2];
    const int cx = (nx + 1) / 2, cy = (ny + 1) / 2, cz = (nz + 1) / 2;
    EFV_LAUNCH(k_agg_lattice, grid_for(n, 256), 256, 0, n, nx, ny, cx, cy, F->agg.p);
    nc = (int64_t)cx * cy * cz;
    C->lat[0] = cx; C->lat[1] = cy; C->lat[2] = cz;
  } else {
    std::vector<int64_t> ptr(n + 1);
    fs->gptr.download(ptr.data(), n + 1);
    std::vector<int32_t> col(ptr[n]);
    fs->gcol.download(col.data(), ptr[n]);
    const int passes = (fs->mesh && fs->mesh->dim == 2) ? 2 : 3;
    std::vector<int32_t> map = host_aggregate(n, ptr, col, passes, &nc);
    F->agg.upload(map.data(), n);
    EFV_CUDA(cudaStreamSynchronize(stream()));
  }
  F->nc = nc;
  // 1b. partitioned level: aggregates of the ghost rows (they live on the neighbours) and the halo of the coarse level.
  // Every rank sends the aggregate number of its interface rows through the level's halo; the distinct numbers received
  // from neighbour q, ascending, become the ghost rows of the coarse level (same order as q derives for its send list).
  const bool dist = F->ng > 0 || (fs->halo.active && comm_active());
  int64_t ncg = 0;
  std::vector<int> c_nbr;
  std::vector<int64_t> c_send_ptr(1, 0), c_recv_ptr(1, 0);
  std::vector<int32_t> c_send_idx;
  if (dist) {
    Halo &fh = fs->halo;
    EFV_REQUIRE(fh.active && fh.peer && comm_peer_active(), "the multigrid preconditioner of a partitioned system needs the peer-memory halo (EFV_COMM=peer)");
    const int64_t ng = F->ng;
    DBuf<double> aggd(n + ng);
    EFV_LAUNCH(k_i32_to_f64, grid_for(n, 256), 256, 0, n, F->agg.p, aggd.p);
    halo_exchange(fh, aggd.p, 1);
    std::vector<double> gh(ng);
    std::vector<int32_t> own(n);
    if (ng) EFV_CUDA(cudaMemcpyAsync(gh.data(), aggd.p + n, ng * sizeof(double), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaMemcpyAsync(own.data(), F->agg.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    std::vector<int32_t> ghost_agg(ng);
    const int nn = (int)fh.nbr.size();
    for (int i = 0; i < nn; ++i) {
      std::vector<int64_t> ids;
      for (int64_t k = fh.recv_ptr[i]; k < fh.recv_ptr[i + 1]; ++k) ids.push_back((int64_t)gh[k]);
      std::vector<int64_t> u = ids;
      std::sort(u.begin(), u.end());
      u.erase(std::unique(u.begin(), u.end()), u.end());
      for (int64_t k = fh.recv_ptr[i]; k < fh.recv_ptr[i + 1]; ++k)
        ghost_agg[k] = (int32_t)(nc + ncg + (std::lower_bound(u.begin(), u.end(), ids[k - fh.recv_ptr[i]]) - u.begin()));
      std::vector<int32_t> snd;
      for (int64_t k = fh.send_ptr[i]; k < fh.send_ptr[i + 1]; ++k) snd.push_back(own[fh.send_idx_host[k]]);
      std::sort(snd.begin(), snd.end());
      snd.erase(std::unique(snd.begin(), snd.end()), snd.end());
      c_nbr.push_back(fh.nbr[i]);
      c_send_idx.insert(c_send_idx.end(), snd.begin(), snd.end());
      c_send_ptr.push_back((int64_t)c_send_idx.size());
      for (int64_t id : u) { C->ghost_rank.push_back(fh.nbr[i]); C->ghost_owner_local.push_back(id); }
      ncg += (int64_t)u.size();
      c_recv_ptr.push_back(ncg);
    }
    if (ng) EFV_CUDA(cudaMemcpyAsync(F->agg.p + n, ghost_agg.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
  }
  // 2. members of every aggregate
  DBuf<int32_t> cnt(nc);
  cnt.zero();
  EFV_LAUNCH(k_count_members, grid_for(n, 256), 256, 0, n, F->agg.p, cnt.p);
  F->mem_ptr.alloc(nc + 1);
  const int64_t total = exclusive_scan_i32_to_i64(cnt.p, F->mem_ptr.p, nc);
  EFV_REQUIRE(total == n, "aggregation lost rows");
  F->mem.alloc(n);
  cnt.zero();
  EFV_LAUNCH(k_fill_members, grid_for(n, 256), 256, 0, n, F->agg.p, F->mem_ptr.p, cnt.p, F->mem.p);
  EFV_LAUNCH(k_sort_lists, grid_for(nc, 128), 128, 0, nc, F->mem_ptr.p, F->mem.p);
  // 3. coarse pattern
  efv_system *cs = new efv_system();
  C->sys = cs; C->owns_sys = true;
  cs->mesh = nullptr; cs->ndof = fs->ndof; cs->nV = nc + ncg; cs->n_owned = nc; cs->rows = (nc + ncg) * fs->ndof;
  DBuf<int32_t> rowlen(nc);
  DBuf<int> err(1);
  err.zero();
  const int grid = grid_for(nc * 32, kCWarps * 32, 8);
  EFV_LAUNCH(k_coarse_rows<false>, grid, kCWarps * 32, 0, nc, n, F->mem_ptr.p, F->mem.p, fs->gptr.p, fs->gcol.p, F->agg.p, rowlen.p,
             (const int64_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, err.p);
  cs->gptr.alloc(nc + 1);
  cs->nnzg = exclusive_scan_i32_to_i64(rowlen.p, cs->gptr.p, nc);
  EFV_REQUIRE(cs->nnzg < (int64_t(1) << 31), "coarse nnz exceeds int32 range");
  cs->gcol.alloc(cs->nnzg);
  cs->diag_slot.alloc(nc);
  EFV_LAUNCH(k_coarse_rows<true>, grid, kCWarps * 32, 0, nc, n, F->mem_ptr.p, F->mem.p, fs->gptr.p, fs->gcol.p, F->agg.p, (int32_t *)nullptr,
             cs->gptr.p, cs->gcol.p, cs->diag_slot.p, err.p);
  cs->vals.alloc((size_t)cs->nnzg * fs->ndof * fs->ndof);
  // 4. gather lists: coarse entry <- fine entries
  DBuf<int32_t> ccnt(cs->nnzg);
  ccnt.zero();
  EFV_LAUNCH((k_gather_lists<0>), grid_for(n, 128), 128, 0, n, fs->gptr.p, fs->gcol.p, F->agg.p, cs->gptr.p, cs->gcol.p, ccnt.p,
             (const int64_t *)nullptr, (int32_t *)nullptr, err.p);
  F->cl_ptr.alloc(cs->nnzg + 1);
  int64_t fine_nnz_owned = 0;
  {
    int64_t ends[2] = {0, 0};
    EFV_CUDA(cudaMemcpyAsync(&ends[1], fs->gptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    fine_nnz_owned = ends[1];
  }
  const int64_t listed = exclusive_scan_i32_to_i64(ccnt.p, F->cl_ptr.p, cs->nnzg);
  int herr = 0;
  err.download(&herr, 1);
  EFV_REQUIRE(herr == 0, herr == 1 ? "an aggregate has more than 2048 candidate columns" : "coarse pattern misses an entry");
  EFV_REQUIRE(listed == fine_nnz_owned, "gather lists do not cover the fine entries");
  F->cl.alloc(listed);
  ccnt.zero();
  EFV_LAUNCH((k_gather_lists<1>), grid_for(n, 128), 128, 0, n, fs->gptr.p, fs->gcol.p, F->agg.p, cs->gptr.p, cs->gcol.p, ccnt.p,
             F->cl_ptr.p, F->cl.p, err.p);
  EFV_LAUNCH(k_sort_lists, grid_for(cs->nnzg, 128), 128, 0, cs->nnzg, F->cl_ptr.p, F->cl.p);
  C->n = nc; C->ng = ncg;
  if (dist) {                                   // halo of the coarse level + its peer staging buffers
    Halo &h = cs->halo;
    h.nbr = c_nbr; h.send_ptr = c_send_ptr; h.recv_ptr = c_recv_ptr;
    h.n_owned = nc; h.n_ghost = ncg; h.cap_ndof = cs->ndof;
    h.send_idx_host = c_send_idx;
    h.send_idx.upload(c_send_idx.data(), c_send_idx.size());
    h.send_buf.alloc(std::max<size_t>(c_send_idx.size() * cs->ndof, 1));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    h.active = true;
    halo_connect_peer(h);
  }
}

// Replicate the (partitioned) level C on every rank: same operator, rows of rank r after those of ranks < r.
static AmgLevel *build_replicated(AmgHierarchy *H, AmgLevel *C, const std::vector<int64_t> &counts) {
  efv_system *cs = C->sys;
  const int nranks = comm_size(), me = comm_rank();
  const int B = cs->ndof * cs->ndof;
  std::vector<int64_t> off(nranks + 1, 0), nnz(nranks, 0), nnz_off(nranks + 1, 0);
  for (int r = 0; r < nranks; ++r) off[r + 1] = off[r] + counts[r];
  int64_t my_nnz = 0;
  EFV_CUDA(cudaMemcpyAsync(&my_nnz, cs->gptr.p + C->n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  comm_allgather_i64(my_nnz, nnz.data());
  for (int r = 0; r < nranks; ++r) nnz_off[r + 1] = nnz_off[r] + nnz[r];
  const int64_t Ng = off[nranks], NNZ = nnz_off[nranks];
  // lattice of the stacked slabs: same (cx, cy) on every rank, planes added up
  std::vector<int64_t> lat(nranks, 0);
  comm_allgather_i64(C->lat[0] > 0 ? ((int64_t)C->lat[0] << 40) | ((int64_t)C->lat[1] << 20) | (int64_t)C->lat[2] : 0, lat.data());
  bool lattice = true;
  int64_t planes = 0;
  for (int r = 0; r < nranks; ++r) {
    lattice = lattice && lat[r] != 0 && (lat[r] >> 20) == (lat[0] >> 20);
    planes += lat[r] & 0xfffff;
  }
  std::vector<int32_t> cm(C->n + C->ng);
  for (int64_t v = 0; v < C->n; ++v) cm[v] = (int32_t)(off[me] + v);
  for (int64_t g = 0; g < C->ng; ++g) cm[C->n + g] = (int32_t)(off[C->ghost_rank[g]] + C->ghost_owner_local[g]);
  DBuf<int32_t> colmap;
  colmap.upload(cm.data(), cm.size());
  H->rep_stage = gather_stage_create((size_t)std::max<int64_t>(NNZ * B, Ng * cs->ndof));
  DBuf<double> len_d(std::max<int64_t>(C->n, 1)), col_d(std::max<int64_t>(my_nnz, 1)), all_len(Ng), all_col(NNZ);
  EFV_LAUNCH(k_pattern_to_f64, grid_for(std::max<int64_t>(C->n, 1), 128), 128, 0, C->n, cs->gptr.p, cs->gcol.p, colmap.p, len_d.p, col_d.p);
  gather_all(H->rep_stage, off[me], C->n, Ng, len_d.p, all_len.p, nullptr);
  gather_all(H->rep_stage, nnz_off[me], my_nnz, NNZ, col_d.p, all_col.p, nullptr);
  std::vector<double> hl(Ng);
  all_len.download(hl.data(), Ng);
  std::vector<int64_t> ptr(Ng + 1, 0);
  for (int64_t v = 0; v < Ng; ++v) ptr[v + 1] = ptr[v] + (int64_t)hl[v];
  EFV_REQUIRE(ptr[Ng] == NNZ, "replicated pattern is inconsistent");
  AmgLevel *R = new AmgLevel();
  efv_system *rs = new efv_system();
  R->sys = rs; R->owns_sys = true;
  rs->mesh = nullptr; rs->ndof = cs->ndof; rs->nV = rs->n_owned = Ng; rs->rows = Ng * cs->ndof; rs->nnzg = NNZ;
  rs->gptr.upload(ptr.data(), Ng + 1);
  rs->gcol.alloc(NNZ);
  EFV_LAUNCH(k_f64_to_i32, grid_for(NNZ, 256), 256, 0, NNZ, all_col.p, rs->gcol.p);
  rs->diag_slot.alloc(Ng);
  EFV_LAUNCH(k_find_diag_slot, grid_for(Ng, 128), 128, 0, Ng, rs->gptr.p, rs->gcol.p, rs->diag_slot.p);
  rs->vals.alloc((size_t)NNZ * B);
  EFV_CUDA(cudaStreamSynchronize(stream()));
  R->n = Ng; R->ng = 0;
  if (lattice) { R->lat[0] = C->lat[0]; R->lat[1] = C->lat[1]; R->lat[2] = (int)planes; }
  C->gathered = true;
  H->rep_row_off = off[me]; H->rep_rows = C->n; H->rep_nnz_off = nnz_off[me]; H->rep_nnz = my_nnz;
  H->rep_total_rows = Ng; H->rep_total_nnz = NNZ;
  return R;
}

AmgHierarchy *amg_symbolic(efv_system *s) {
  AmgHierarchy *H = new AmgHierarchy();
  try {
    const char *e;
    if ((e = getenv("EFV_AMG_DEGREE"))) H->deg = std::max(1, std::min(kMaxDeg, atoi(e)));
    if ((e = getenv("EFV_AMG_RATIO"))) H->ratio = std::max(1.5, atof(e));
    if ((e = getenv("EFV_AMG_POWER"))) H->power_steps = std::max(0, atoi(e));
    H->scale = s->ndof == 1 ? 1.8 : 1.3;
    if ((e = getenv("EFV_AMG_SCALE"))) H->scale = atof(e);
    AmgLevel *L = new AmgLevel();
    L->sys = s; L->n = s->n_owned; L->ng = s->nV - s->n_owned;
    for (int k = 0; k < 3; ++k) L->lat[k] = s->lattice[k];
    H->lv.push_back(L);
    H->dist = is_dist(s);
    const int nranks = H->dist ? comm_size() : 1;
    std::vector<int64_t> counts(nranks, 0);
    bool replicated = false;                        // from here on every rank holds the whole operator
    if ((e = getenv("EFV_AMG_GATHER_ROWS"))) H->gather_rows = atoll(e);
    auto global_rows = [&](int64_t mine) {          // block rows of a level over all ranks (collective: every rank, every level)
      if (!H->dist || replicated) { counts.assign(counts.size(), 0); counts[0] = mine; return mine; }
      comm_allgather_i64(mine, counts.data());
      int64_t t = 0;
      for (int64_t c : counts) t += c;
      return t;
    };
    int64_t glob = global_rows(L->n);
    while ((int)H->lv.size() < kMaxLevels) {
      AmgLevel *F = H->lv.back();
      if (glob * s->ndof <= kCoarsestRows) break;
      AmgLevel *C = new AmgLevel();
      H->lv.push_back(C);
      build_transfer(F, C);
      const int64_t cglob = global_rows(C->n);
      EFV_REQUIRE(cglob < glob, "aggregation did not coarsen the operator");
      glob = cglob;
      if (H->dist && !replicated && glob <= H->gather_rows && glob * s->ndof > kCoarsestRows) {
        AmgLevel *R = build_replicated(H, C, counts);
        H->lv.push_back(R);
        replicated = true;
      }
    }
    AmgLevel *last = H->lv.back();
    EFV_REQUIRE(glob * s->ndof <= kCoarsestRows, "hierarchy too deep");
    EFV_REQUIRE(!H->dist || H->lv.size() >= 2, "a partitioned system of <= 256 rows needs no multigrid (use the Jacobi preconditioner)");
    H->nd = glob * s->ndof;
    H->dense.alloc((size_t)H->nd * 2 * H->nd);
    H->dense_dist = H->dist && !replicated;
    if (H->dense_dist) {
      // gathered coarsest system: rank r's rows come after those of ranks < r; ghost columns map through (owner, row there)
      std::vector<int64_t> off(nranks + 1, 0);
      for (int r = 0; r < nranks; ++r) off[r + 1] = off[r] + counts[r];
      H->nd_off = off[comm_rank()] * s->ndof;
      H->nd_local = last->n * s->ndof;
      std::vector<int32_t> cm(last->n + last->ng);
      for (int64_t v = 0; v < last->n; ++v) cm[v] = (int32_t)(off[comm_rank()] + v);
      for (int64_t g = 0; g < last->ng; ++g) cm[last->n + g] = (int32_t)(off[last->ghost_rank[g]] + last->ghost_owner_local[g]);
      H->colmap.upload(cm.data(), cm.size());
      H->bfull.alloc(H->nd);
      H->dense_local.alloc((size_t)std::max<int64_t>(H->nd_local, 1) * 2 * H->nd);
      H->stage = gather_stage_create((size_t)H->nd * 2 * H->nd);
    }
    for (size_t l = 0; l < H->lv.size(); ++l) level_vectors(H->lv[l], l == 0);
    EFV_CUDA(cudaStreamSynchronize(stream()));
    return H;
  } catch (...) {
    delete H;
    throw;
  }
}
void amg_free(AmgHierarchy *H) { delete H; }

// numeric set-up for the current values of the system (and sign convention sigma)
void amg_numeric(AmgHierarchy *H, double sigma) {
  efv_system *s = H->lv[0]->sys;
  if (H->numeric_version == s->vals_version && H->numeric_sigma == sigma) return;
  KrylovWork *w0 = work(s);
  EFV_CUDA(cudaEventRecord(w0->e0, stream()));
  for (size_t l = 0; l < H->lv.size(); ++l) {
    AmgLevel *L = H->lv[l];
    efv_system *ls = L->sys;
    KrylovWork *w = work(ls);
    if (l > 0) {
      AmgLevel *F = H->lv[l - 1];
      const int B = ls->ndof * ls->ndof;
      if (F->gathered)          // replicated copy of the level above: every rank's rows, in rank order
        gather_all(H->rep_stage, H->rep_nnz_off * B, H->rep_nnz * B, H->rep_total_nnz * B, F->sys->vals.p, ls->vals.p, nullptr);
      else
        EFV_LAUNCH(k_galerkin, grid_for(ls->nnzg * B, 256), 256, 0, ls->nnzg, B, F->cl_ptr.p, F->cl.p, F->sys->vals.p, ls->vals.p);
      ls->vals_version++;
    }
    if (L->gathered) continue;  // no smoothing on the partitioned copy: the cycle goes through the replicated one
    const bool ldist = is_dist(ls);
    if (l + 1 < H->lv.size()) {
      bool have_rowabs = false;
      sell_prepare(ls, w, L->r.p, &have_rowabs);                       // L->r doubles as the row-sum buffer during set-up
      if (have_rowabs)
        EFV_LAUNCH(k_diag_from_rowabs, w->grid, 256, 0, L->n, ls->ndof, ls->vals.p, ls->diag_slot.p, L->r.p, sigma, L->dinv.p, w->partial.p);
      else
        EFV_LAUNCH(k_diag_gershgorin, w->grid, 256, 0, L->n, ls->ndof, ls->gptr.p, ls->gcol.p, ls->vals.p, ls->diag_slot.p, sigma, L->dinv.p,
                   w->partial.p);
      // largest eigenvalue of D^-1 A by a power iteration that restarts from the previous estimate of the eigenvector
      // (12 steps the first time, 2 afterwards): the Gershgorin bound alone (2.0 on these operators against 1.2-1.9 true)
      // costs ~20 % more CG iterations
      const double *pow_dots = nullptr;
      int Gp = w->grid;
      double *PD = w->partial.p + 2 * (size_t)w->grid;                  // [2G, 4G): (y.y, v.v)
      if (H->power_steps > 0) {
        const int64_t len = L->n * ls->ndof;
        if (!L->ev.n) { L->ev.alloc((size_t)(L->n + L->ng) * ls->ndof); L->ev.zero(); }
        if (!L->ev_ready) EFV_LAUNCH(k_pow_init, grid_for(len, 256, 8), 256, 0, len, L->ev.p);
        const int steps = L->ev_ready ? H->power_steps : 6 * H->power_steps;
        for (int k = 0; k < steps; ++k) {
          spmv_halo(ls, w->grid, L->ev.p, plain_epi(L->res.p, sigma, L->dinv.p, nullptr), nullptr, nullptr);
          EFV_LAUNCH(k_pow_dots, w->grid, 256, 0, len, L->res.p, L->ev.p, PD);
          pow_dots = PD; Gp = w->grid;
          if (ldist) { comm_allreduce_partials(PD, w->grid, 2, w->sc.p + 52, 0); pow_dots = w->sc.p + 52; Gp = 1; }
          EFV_LAUNCH(k_pow_scale, grid_for(len, 256, 8), 256, 0, len, L->res.p, pow_dots, Gp, L->ev.p);
        }
        L->ev_ready = true;
      }
      if (ldist) {            // one interval for all ranks: the smoother must be the same polynomial everywhere
        comm_allreduce_partials(w->partial.p, w->grid, 1, w->sc.p + 50, 1);
        EFV_LAUNCH(k_cheb_coef, 1, 256, 0, w->sc.p + 50, 1, H->ratio, pow_dots, Gp, L->coef.p);
      } else {
        EFV_LAUNCH(k_cheb_coef, 1, 256, 0, w->partial.p, w->grid, H->ratio, pow_dots, Gp, L->coef.p);
      }
    } else {
      const int64_t N = H->nd;
      if (!H->dense_dist) {
        EFV_LAUNCH(k_dense_fill, grid_for(N * 2 * N, 256), 256, 0, N, (int64_t)0, N, H->dense.p);
        EFV_LAUNCH(k_dense_entries, grid_for(L->n, 64), 64, 0, L->n, ls->ndof, ls->gptr.p, ls->gcol.p, ls->vals.p, sigma, (const int32_t *)nullptr,
                   (int64_t)0, N, H->dense.p);
      } else {                // my rows of [sigma A | I] in gathered numbering, then every rank receives every block of rows
        EFV_LAUNCH(k_dense_fill, grid_for(std::max<int64_t>(H->nd_local, 1) * 2 * N, 256), 256, 0, H->nd_local, H->nd_off, N, H->dense_local.p);
        EFV_LAUNCH(k_dense_entries, grid_for(std::max<int64_t>(L->n, 1), 64), 64, 0, L->n, ls->ndof, ls->gptr.p, ls->gcol.p, ls->vals.p, sigma,
                   H->colmap.p, H->nd_off, N, H->dense_local.p);
        gather_all(H->stage, H->nd_off * 2 * N, H->nd_local * 2 * N, N * 2 * N, H->dense_local.p, H->dense.p, nullptr);
      }
      EFV_LAUNCH(k_gauss_jordan, 1, 1024, 0, (int)N, H->dense.p);
    }
  }
  EFV_CUDA(cudaEventRecord(w0->e1, stream()));
  H->numeric_timed = true;
  H->numeric_version = s->vals_version;
  H->numeric_sigma = sigma;
}

// ---- cycle ------------------------------------------------------------------------------------------------------------------------------
// z = M b on level l (b, z: level vectors; on level 0 the caller's).  rz_partial (level 0 only): per-block partials of b . z,
// produced by the epilogue of the last smoothing step.
static void cycle(AmgHierarchy *H, size_t l, const double *b, double *x, double sigma, double *rz_partial, const int *done,
                  bool first_direction_ready) {
  AmgLevel *L = H->lv[l];
  efv_system *ls = L->sys;
  KrylovWork *w = work(ls);
  const int G = w->grid;
  const int64_t len = L->n * ls->ndof;
  if (L->gathered) {             // hand over to the replicated copy: gather the right-hand side, cycle there, keep my rows
    AmgLevel *R = H->lv[l + 1];
    const int nd = ls->ndof;
    gather_all(H->rep_stage, H->rep_row_off * nd, H->rep_rows * nd, H->rep_total_rows * nd, b, R->b.p, done);
    cycle(H, l + 1, R->b.p, R->x.p, sigma, nullptr, done, false);
    EFV_LAUNCH(k_copy_slice, grid_for(std::max<int64_t>(len, 1), 256, 8), 256, 0, len, R->x.p + H->rep_row_off * nd, x, done);
    return;
  }
  if (l + 1 == H->lv.size()) {
    if (H->dense_dist) {
      gather_all(H->stage, H->nd_off, H->nd_local, H->nd, b, H->bfull.p, done);
      EFV_LAUNCH(k_dense_solve, 1, 256, 0, (int)H->nd, (int)H->nd_off, (int)H->nd_local, H->dense.p, H->bfull.p, x, done);
    } else {
      EFV_LAUNCH(k_dense_solve, 1, 256, 0, (int)H->nd, 0, (int)H->nd, H->dense.p, b, x, done);
    }
    if (rz_partial) EFV_LAUNCH(k_dot_partial, 1, 256, 0, len, b, x, rz_partial, G, done);      // single-level hierarchy (<= 256 rows)
    return;
  }
  const int deg = H->deg;
  auto cheb = [&](int step, const double *din, double *dout, int flags, bool last, const double *dotw, double *partial) {
    SpmvEpi e;
    e.mode = last ? kEpiChebLast : kEpiChebMid;
    e.flags = flags; e.sigma = sigma;
    e.b = b; e.dinv = L->dinv.p; e.r_in = L->r.p; e.r_out = L->r.p; e.d_in = din; e.d_out = dout; e.x = x;
    e.coef = L->coef.p + 4 * step; e.w = dotw;
    spmv_halo(ls, G, const_cast<double *>(din), e, partial, done);
  };
  const std::string tag = "L" + std::to_string(l);
  // pre-smoothing from x = 0
  if (!first_direction_ready)     // (coarse levels: written by the restriction kernel)
    EFV_LAUNCH(k_first_direction, grid_for(len, 256, 8), 256, 0, len, b, L->dinv.p, L->coef.p, deg == 1 ? x : L->d0.p, done);
  double *din = L->d0.p, *dout = L->d1.p;
  for (int k = 1; k < deg; ++k) {
    cheb(k, din, dout, k == 1 ? (kEpiXZero | kEpiRFromB) : 0, k == deg - 1, nullptr, nullptr);
    std::swap(din, dout);
  }
  if (g_pt.on) EFV_PT(tag + ".pre");
  // residual, restriction (+ first direction of the next level), coarse solve, over-corrected prolongation
  {
    SpmvEpi e;
    e.mode = kEpiResid; e.sigma = sigma; e.b = b; e.y = L->res.p;
    spmv_halo(ls, G, x, e, nullptr, done);
  }
  if (g_pt.on) EFV_PT(tag + ".resid");
  AmgLevel *C = H->lv[l + 1];
  const bool c_smooths = l + 2 < H->lv.size() && !C->gathered;
  EFV_LAUNCH(k_restrict, grid_for(C->n * ls->ndof, 256, 8), 256, 0, C->n, ls->ndof, L->mem_ptr.p, L->mem.p, L->res.p, C->b.p,
             c_smooths ? C->dinv.p : (const double *)nullptr, c_smooths ? C->coef.p : (const double *)nullptr,
             c_smooths ? (H->deg == 1 ? C->x.p : C->d0.p) : (double *)nullptr, done);
  if (g_pt.on) EFV_PT(tag + ".restrict");
  cycle(H, l + 1, C->b.p, C->x.p, sigma, nullptr, done, true);
  if (g_pt.on) EFV_PT(tag + ".coarse_tail");
  EFV_LAUNCH(k_prolong, grid_for(len, 256, 8), 256, 0, L->n, ls->ndof, L->agg.p, C->x.p, H->scale, x, done);
  // post-smoothing
  {
    SpmvEpi e;
    e.mode = kEpiChebFirst; e.sigma = sigma; e.b = b; e.dinv = L->dinv.p; e.r_out = L->r.p; e.d_out = L->d0.p; e.coef = L->coef.p;
    spmv_halo(ls, G, x, e, nullptr, done);
  }
  if (deg == 1) {
    EFV_LAUNCH(k_axpy1, grid_for(len, 256, 8), 256, 0, len, L->d0.p, x, done);
    EFV_REQUIRE(!rz_partial, "degree-1 smoothing has no fused r.z (use degree >= 2)");
    return;
  }
  din = L->d0.p; dout = L->d1.p;
  for (int k = 1; k < deg; ++k) {
    const bool last = k == deg - 1;
    cheb(k, din, dout, 0, last, last && rz_partial ? b : nullptr, last ? rz_partial : nullptr);
    std::swap(din, dout);
  }
  if (g_pt.on) EFV_PT(tag + ".post");
}

// ---- CG preconditioned by the cycle ----------------------------------------------------------------------------------------------------
// r = sigma b - y; partials [0] r.r [1] b.b
__global__ void __launch_bounds__(256) k_mg_init(int64_t n, const double *__restrict__ b, const double *__restrict__ y, double sigma,
                                                 double *__restrict__ r, double *__restrict__ partial) {
  __shared__ double red[32];
  double rr = 0, bb = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = sigma * b[i], ri = bi - y[i];
    r[i] = ri;
    rr += ri * ri; bb += bi * bi;
  }
  const int G = gridDim.x;
  rr = block_sum(rr, red); if (threadIdx.x == 0) partial[blockIdx.x] = rr;
  bb = block_sum(bb, red); if (threadIdx.x == 0) partial[G + blockIdx.x] = bb;
}
// sc: [0] rz, [2] b.b, [3] r.r, [4] rtol^2; flags: [0] done [1] iterations [2] breakdown
__global__ void k_mg_init_scalars(int G, const double *__restrict__ partial, double *__restrict__ sc, int *__restrict__ flags, double rtol) {
  __shared__ double red[32];
  const double rr = sum_partials(partial, G, red);
  const double bb = sum_partials(partial + G, G, red);
  if (threadIdx.x == 0) {
    sc[0] = 0.0; sc[2] = bb; sc[3] = rr; sc[4] = rtol * rtol;
    flags[0] = (rr <= rtol * rtol * bb) ? 1 : 0;
    flags[1] = 0; flags[2] = 0;
  }
}
// first direction: p = z, rz = r.z
__global__ void __launch_bounds__(256) k_mg_first_p(int64_t n, const double *__restrict__ z, double *__restrict__ p,
                                                    const double *__restrict__ rz_partial, int G, double *__restrict__ sc,
                                                    const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  const double rz = sum_partials(rz_partial, G, red);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) sc[0] = rz;
}
// alpha = rz / p.Ap; x += alpha p; r -= alpha Ap; partial r.r
__global__ void __launch_bounds__(256) k_mg_xr(int64_t n, const double *__restrict__ sc, const double *__restrict__ pAp_partial, int G,
                                               const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ x,
                                               double *__restrict__ r, double *__restrict__ partial, const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  const double pAp = sum_partials(pAp_partial, G, red);
  if (!(pAp > 0.0) || !(pAp < 1e300)) return;                 // k_mg_check raises the breakdown flag
  const double alpha = sc[0] / pAp;
  double rr = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    rr += ri * ri;
  }
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = rr;
}
// one block: convergence / breakdown decision right after the update, so that a converged solve skips the last cycle
__global__ void k_mg_check(const double *__restrict__ pAp_partial, int Gp, const double *__restrict__ rr_partial, int Gr,
                           double *__restrict__ sc, int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  const double pAp = sum_partials(pAp_partial, Gp, red);
  const double rr = sum_partials(rr_partial, Gr, red);
  if (threadIdx.x == 0) {
    if (!(pAp > 0.0) || !(pAp < 1e300) || !(sc[0] > 0.0)) { flags[2] = 1; flags[0] = 1; return; }
    sc[3] = rr;
    flags[1] += 1;                                    // iterations done (device-side: the iteration body is a replayed CUDA graph)
    if (!(rr > sc[4] * sc[2])) flags[0] = 1;
  }
}
// beta = rz_new / rz; p = z + beta p
__global__ void __launch_bounds__(256) k_mg_p(int64_t n, double *__restrict__ sc, const double *__restrict__ rz_partial, int G,
                                              const double *__restrict__ z, double *__restrict__ p, const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  const double rz_new = sum_partials(rz_partial, G, red);
  const double beta = rz_new / sc[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
  // sc[0] is read by every block above and overwritten by the NEXT kernel in the stream (k_mg_store_rz)
}
__global__ void k_mg_store_rz(const double *__restrict__ rz_partial, int G, double *__restrict__ sc, const int *__restrict__ flags) {
  __shared__ double red[32];
  if (flags[0]) return;
  const double rz = sum_partials(rz_partial, G, red);
  if (threadIdx.x == 0) sc[0] = rz;
}

void solve_pcg_amg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  ensure_vectors(s);
  KrylovWork *w = work(s);
  const int64_t n = s->n_owned * s->ndof;
  if (!w->r.n) { w->r.alloc(s->rows); w->p.alloc(s->rows); w->Ap.alloc(s->rows); w->dinv.alloc(s->rows); }
  if (!w->v.n) { w->v.alloc(s->rows); w->v.zero(); }                 // z
  const double sigma = negate ? -1.0 : 1.0;
  if (!s->amg) s->amg = amg_symbolic(s);
  AmgHierarchy *H = s->amg;
  const int G = w->grid;
  double *P = w->partial.p;                                          // [0,G) r.r  [G,2G) b.b / r.z  [3G,4G) p.Ap
  cudaEvent_t t0, t1;
  EFV_CUDA(cudaEventCreate(&t0));
  EFV_CUDA(cudaEventCreate(&t1));
  EFV_CUDA(cudaEventRecord(t0, stream()));
  amg_numeric(H, sigma);          // includes the solver-side SELL copy of every level
  sell_prepare(s, w);             // (no-op unless the hierarchy was current and the copy is not)
  EFV_CUDA(cudaMemsetAsync(w->flags.p, 0, 4 * sizeof(int), stream()));
  int Gc = G, Gp = G, Gr = G, Gz = G;               // partitioned runs: the partials are all-reduced into single scalars
  spmv_halo(s, G, s->x.p, plain_epi(w->Ap.p, sigma, nullptr, nullptr), nullptr, nullptr);
  EFV_LAUNCH(k_mg_init, G, 256, 0, n, s->b.p, w->Ap.p, sigma, w->r.p, P);
  const double *c0 = consolidate(s, P, 2, 0, &Gc);
  EFV_LAUNCH(k_mg_init_scalars, 1, 256, 0, Gc, c0, w->sc.p, w->flags.p, rtol);
  double *z = w->v.p;
  cycle(H, 0, w->r.p, z, sigma, P + G, w->flags.p, false);
  const double *cz = consolidate(s, P + G, 1, 3, &Gz);
  EFV_LAUNCH(k_mg_first_p, G, 256, 0, n, z, w->p.p, cz, Gz, w->sc.p, w->flags.p);
  int it = 0;
  bool done = false;
  const int check_every = (n > 200000) ? 1 : 4;
  { const char *et = getenv("EFV_AMG_TIMING"); g_pt.on = et && atoi(et) == 1; }
  EFV_PT("start");
  auto body = [&]() {
      spmv_halo(s, G, w->p.p, plain_epi(w->Ap.p, sigma, nullptr, w->p.p), P + 3 * G, w->flags.p);
      EFV_PT("cg.spmv");
      const double *cp = consolidate(s, P + 3 * G, 1, 1, &Gp);
      EFV_LAUNCH(k_mg_xr, G, 256, 0, n, w->sc.p, cp, Gp, w->p.p, w->Ap.p, s->x.p, w->r.p, P, w->flags.p);
      const double *cr = consolidate(s, P, 1, 2, &Gr);
      EFV_LAUNCH(k_mg_check, 1, 256, 0, cp, Gp, cr, Gr, w->sc.p, w->flags.p);
      EFV_PT("cg.xr_check");
      cycle(H, 0, w->r.p, z, sigma, P + G, w->flags.p, false);
      const double *cz2 = consolidate(s, P + G, 1, 3, &Gz);
      EFV_LAUNCH(k_mg_p, G, 256, 0, n, w->sc.p, cz2, Gz, z, w->p.p, w->flags.p);
      EFV_LAUNCH(k_mg_store_rz, 1, 256, 0, cz2, Gz, w->sc.p, w->flags.p);
      EFV_PT("cg.p");
  };
  static const int graph_on = [] { const char *e = getenv("EFV_AMG_GRAPH"); return (e && atoi(e) == 0) ? 0 : 1; }();
  const bool use_graph = graph_on && !is_dist(s) && !g_pt.on;
  if (use_graph && (!H->iter_graph || H->graph_sigma != sigma)) {
    if (H->iter_graph) { cudaGraphExecDestroy(H->iter_graph); H->iter_graph = nullptr; }
    cudaGraph_t graph = nullptr;
    const int64_t before = launch_count();
    EFV_CUDA(cudaStreamBeginCapture(stream(), cudaStreamCaptureModeThreadLocal));
    try {
      body();
    } catch (...) {
      cudaStreamEndCapture(stream(), &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    EFV_CUDA(cudaStreamEndCapture(stream(), &graph));
    H->graph_nodes = (int)(launch_count() - before);
    count_launch(-H->graph_nodes);                            // nothing ran yet
    EFV_CUDA(cudaGraphInstantiate(&H->iter_graph, graph, 0));
    cudaGraphDestroy(graph);
    H->graph_sigma = sigma;
  }
  while (!done && it < maxit) {
    const int batch_end = std::min(maxit, it + check_every);
    for (; it < batch_end; ++it) {
      if (use_graph) {
        EFV_CUDA(cudaGraphLaunch(H->iter_graph, stream()));
        count_launch(H->graph_nodes);
      } else {
        body();
      }
    }
    EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    done = w->h_flags[0] != 0;
    EFV_PT("cg.poll");
  }
  g_pt.report(comm_rank(), (long long)n);
  EFV_CUDA(cudaMemcpyAsync(w->h_flags, w->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaMemcpyAsync(w->h_sc, w->sc.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaEventRecord(t1, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  float ms = 0;
  EFV_CUDA(cudaEventElapsedTime(&ms, t0, t1));
  s->t_solve_ms = ms;
  if (H->numeric_timed) {
    EFV_CUDA(cudaEventElapsedTime(&ms, w->e0, w->e1));
    H->t_numeric_ms = ms;
    H->numeric_timed = false;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  if (is_dist(s) && comm_peer_active()) EFV_REQUIRE(comm_peer_error() == 0, "peer-memory collective timed out (a rank did not arrive)");
  if (iters) *iters = w->h_flags[1];
  if (relres) *relres = w->h_flags[2] ? NAN : ((w->h_sc[2] > 0) ? std::sqrt(w->h_sc[3] / w->h_sc[2]) : 0.0);
}

// ---- block preconditioner for the coupled Biot operator (BiCGStab, krylov.cu) ---------------------------------------------------
// The coupled system [A_uu A_up; A_pu A_pp] (d + 1 unknowns per vertex, row-replaced Dirichlet form, non-symmetric) is
// preconditioned block-diagonally: one multigrid V-cycle on the displacement block and one on the pressure block.
// Both blocks are extracted from the assembled values into two systems on the SAME vertex graph, in the form CG-type
// smoothers need: -A_uu (the reference's stiffness rows are negative definite) and A_pp, Dirichlet
// rows AND columns replaced by the identity.  With exact block solves BiCGStab needs 11-15 iterations independently of
// the mesh size (numpy prototype on the oracle's matrices, n = 10...22) against 80-240 (and 660-1 700 at C4) with the
// diagonal alone.
struct BlockPrec {
  efv_system *uu = nullptr, *pp = nullptr;
  uint64_t version = 0;
  DBuf<double> ru, rp, zu, zp;
  ~BlockPrec() {
    for (efv_system *q : {uu, pp})
      if (q) { if (q->amg) amg_free(q->amg); krylov_free(q); delete q; }
  }
};
__global__ void k_bp_extract(int64_t n_owned, int nd, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                             const double *__restrict__ vals, const uint8_t *__restrict__ dflag, double *__restrict__ vuu,
                             double *__restrict__ vpp) {
  const int d = nd - 1, B = nd * nd, Bu = d * d;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_owned; v += (int64_t)gridDim.x * blockDim.x)
    for (int64_t t = gptr[v]; t < gptr[v + 1]; ++t) {
      const int64_t c = gcol[t];
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          double a = -vals[t * B + i * nd + j];
          if (dflag[v * nd + i]) a = (c == v && i == j) ? 1.0 : 0.0;
          else if (dflag[c * nd + j]) a = 0.0;
          vuu[t * Bu + i * d + j] = a;
        }
      double a = vals[t * B + d * nd + d];
      if (dflag[v * nd + d]) a = (c == v) ? 1.0 : 0.0;
      else if (dflag[c * nd + d]) a = 0.0;
      vpp[t] = a;
    }
}
// right-hand sides of the two block solves from a vector of the ROW-SCALED coupled system: rhs = D w (undo the scaling),
// displacement rows negated unless Dirichlet (the block holds -A_uu)
__global__ void __launch_bounds__(256) k_bp_split(int64_t n_owned, int nd, const double *__restrict__ w, const double *__restrict__ dinv,
                                                  const uint8_t *__restrict__ dflag, double *__restrict__ ru, double *__restrict__ rp,
                                                  const int *__restrict__ done) {
  if (done && *done) return;
  const int d = nd - 1;
  const int64_t n = n_owned * nd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / nd;
    const int f = (int)(i % nd);
    const double val = w[i] / dinv[i];
    if (f < d) ru[v * d + f] = dflag[i] ? val : -val;
    else rp[v] = val;
  }
}
__global__ void __launch_bounds__(256) k_bp_merge(int64_t n_owned, int nd, const double *__restrict__ zu, const double *__restrict__ zp,
                                                  double *__restrict__ z, const int *__restrict__ done) {
  if (done && *done) return;
  const int d = nd - 1;
  const int64_t n = n_owned * nd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / nd;
    const int f = (int)(i % nd);
    z[i] = (f < d) ? zu[v * d + f] : zp[v];
  }
}
static efv_system *block_system(efv_system *s, int ndof) {
  efv_system *q = new efv_system();
  q->mesh = nullptr; q->ndof = ndof; q->nV = s->nV; q->n_owned = s->n_owned; q->rows = s->nV * ndof; q->nnzg = s->nnzg;
  q->max_row_len = s->max_row_len;
  q->gptr.alloc(s->nV + 1); q->gcol.alloc(s->nnzg); q->diag_slot.alloc(s->nV);
  EFV_CUDA(cudaMemcpyAsync(q->gptr.p, s->gptr.p, (s->nV + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream()));
  EFV_CUDA(cudaMemcpyAsync(q->gcol.p, s->gcol.p, s->nnzg * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream()));
  EFV_CUDA(cudaMemcpyAsync(q->diag_slot.p, s->diag_slot.p, s->nV * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream()));
  q->vals.alloc((size_t)s->nnzg * ndof * ndof);
  for (int k = 0; k < 3; ++k) q->lattice[k] = s->lattice[k];
  q->precond = 1;
  return q;
}
void block_prec_free(BlockPrec *b) { delete b; }
// (re)build for the current values of the coupled system
void block_prec_prepare(efv_system *s) {
  EFV_REQUIRE(s->ndof >= 3 && s->dir_flag.n, "the block preconditioner is for the coupled (d + 1)-unknown Biot system");
  EFV_REQUIRE(!is_dist(s), "the block preconditioner of a partitioned coupled system is not implemented");
  if (!s->bprec) {
    BlockPrec *b = new BlockPrec();
    s->bprec = b;
    b->uu = block_system(s, s->ndof - 1);
    b->pp = block_system(s, 1);
    const size_t nV = (size_t)s->nV;
    b->ru.alloc(nV * (s->ndof - 1)); b->zu.alloc(nV * (s->ndof - 1)); b->rp.alloc(nV); b->zp.alloc(nV);
    b->ru.zero(); b->zu.zero(); b->rp.zero(); b->zp.zero();
  }
  BlockPrec *b = s->bprec;
  if (b->version == s->vals_version) return;
  EFV_LAUNCH(k_bp_extract, grid_for(s->n_owned, 128), 128, 0, s->n_owned, s->ndof, s->gptr.p, s->gcol.p, s->vals.p, s->dir_flag.p,
             b->uu->vals.p, b->pp->vals.p);
  b->uu->vals_version++;
  b->pp->vals_version++;
  for (efv_system *q : {b->uu, b->pp}) {
    if (!q->amg) q->amg = amg_symbolic(q);
    amg_numeric(q->amg, 1.0);
  }
  b->version = s->vals_version;
}
// z = M^-1 w for vectors of the row-scaled coupled system (dinv = 1 / diagonal of the coupled operator)
void block_prec_apply(efv_system *s, const double *w, const double *dinv, double *z, const int *done) {
  BlockPrec *b = s->bprec;
  const int64_t n = s->n_owned * s->ndof;
  EFV_LAUNCH(k_bp_split, grid_for(n, 256, 8), 256, 0, s->n_owned, s->ndof, w, dinv, s->dir_flag.p, b->ru.p, b->rp.p, done);
  cycle(b->uu->amg, 0, b->ru.p, b->zu.p, 1.0, nullptr, done, false);
  cycle(b->pp->amg, 0, b->rp.p, b->zp.p, 1.0, nullptr, done, false);
  EFV_LAUNCH(k_bp_merge, grid_for(n, 256, 8), 256, 0, s->n_owned, s->ndof, b->zu.p, b->zp.p, z, done);
}

// diagnostics for tests / DESIGN.md: levels, rows and nnz per level, Chebyshev bound per level, last numeric set-up time
int amg_info(efv_system *s, int64_t *rows, int64_t *nnz, double *lmax, int cap, double *numeric_ms) {
  if (!s->amg) return 0;
  AmgHierarchy *H = s->amg;
  const int nl = (int)H->lv.size();
  for (int l = 0; l < nl && l < cap; ++l) {
    rows[l] = H->lv[l]->n * H->lv[l]->sys->ndof;
    nnz[l] = H->lv[l]->sys->nnzg * H->lv[l]->sys->ndof * H->lv[l]->sys->ndof;
    double v = 0.0;
    if (l + 1 < nl) EFV_CUDA(cudaMemcpyAsync(&v, H->lv[l]->coef.p + 4 * kMaxDeg, sizeof(double), cudaMemcpyDeviceToHost, stream()));
    EFV_CUDA(cudaStreamSynchronize(stream()));
    lmax[l] = v;
  }
  if (numeric_ms) *numeric_ms = H->t_numeric_ms;
  return nl;
}

// one level of the hierarchy on the host (tests): pattern, values, aggregate of every owned row (level < last)
void amg_level_export(efv_system *s, int level, int64_t *gptr, int32_t *gcol, double *vals, int32_t *agg) {
  EFV_REQUIRE(s->amg, "no multigrid hierarchy (solve once with the multigrid preconditioner first)");
  AmgHierarchy *H = s->amg;
  EFV_REQUIRE(level >= 0 && level < (int)H->lv.size(), "level out of range");
  AmgLevel *L = H->lv[level];
  efv_system *ls = L->sys;
  if (gptr) ls->gptr.download(gptr, L->n + 1);
  if (gcol) ls->gcol.download(gcol, ls->nnzg);
  if (vals) ls->vals.download(vals, (size_t)ls->nnzg * ls->ndof * ls->ndof);
  if (agg) { EFV_REQUIRE(L->agg.n, "the coarsest level has no aggregates"); L->agg.download(agg, L->n); }
}

}  // namespace efv
