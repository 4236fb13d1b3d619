// csr.cu -- north_star kernel 2: CSR sparsity (union of element cliques, sorted columns) and the
// element-to-CSR slot map.  Structure is bit-exact w.r.t. the oracle (integer work, no atomics on results).
//
// One warp per graph row: gather the candidate columns of all incident elements into shared memory, rank the
// distinct values (O(n^2/32) per row, n <= 64 for hexahedra), write them in ascending order.
// Pass 1 counts, a device scan turns counts into row pointers, pass 2 fills.  The slot map stores, for
// element e and local pair (i, j), the POSITION of column conn[e][j] inside row conn[e][i] as one byte
// (rows are limited to 255 entries); the absolute slot is gptr[row] + position.
#include "views.cuh"

namespace efv {

constexpr int kRowWarps = 8;        // warps per block
constexpr int kMaxCand = 1024;      // candidate columns per row held in shared memory

template <bool FILL>
__global__ void __launch_bounds__(kRowWarps * 32) k_graph_rows(MeshView m, int32_t *__restrict__ rowlen,
                                                               const int64_t *__restrict__ gptr,
                                                               int32_t *__restrict__ gcol,
                                                               int32_t *__restrict__ diag_slot,
                                                               int *__restrict__ error_flag) {
  __shared__ int32_t cand_s[kRowWarps][kMaxCand];
  __shared__ uint8_t first_s[kRowWarps][kMaxCand];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t *cand = cand_s[warp];
  uint8_t *first = first_s[warp];
  for (int64_t v = (int64_t)blockIdx.x * kRowWarps + warp; v < m.nV; v += (int64_t)gridDim.x * kRowWarps) {
    // gather candidates
    int n = 0;
    int64_t p0 = m.inc_ptr[v], p1 = m.inc_ptr[v + 1];
    for (int64_t p = p0; p < p1; ++p) {
      int64_t e = inc_elem(m.inc[p]);
      int mm = shape_m(m.code(e));
      int64_t b = m.begin(e);
      if (n + mm > kMaxCand) { if (lane == 0) atomicExch(error_flag, 1); break; }
      if (lane < mm) cand[n + lane] = m.conn[b + lane];
      n += mm;
    }
    if (p0 == p1) {   // isolated vertex: keep a diagonal entry so that the matrix stays square and solvable
      if (lane == 0) cand[0] = (int32_t)v;
      n = 1;
    }
    __syncwarp();
    if (n <= 64) {
      // fast path (hexahedra: 64 candidates, tetrahedra on structured meshes: <= 96 -> slow path): bitonic sort of
      // 64 keys held two per lane (key i lives in lane i & 31, slot i >> 5), then neighbour comparison for uniqueness
      int k0 = (lane < n) ? cand[lane] : 0x7fffffff;
      int k1 = (lane + 32 < n) ? cand[lane + 32] : 0x7fffffff;
#pragma unroll
      for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
          if (j == 32) {                                  // partner = other slot of the same lane (only when size == 64)
            const int lo = min(k0, k1), hi = max(k0, k1);
            k0 = lo; k1 = hi;
          } else {
            const int o0 = __shfl_xor_sync(0xffffffffu, k0, j), o1 = __shfl_xor_sync(0xffffffffu, k1, j);
            const bool lower = (lane & j) == 0;           // this lane holds the lower index of the pair
            const bool asc0 = (size == 64) ? true : ((lane & size) == 0);
            const bool asc1 = (size == 64) ? true : (((lane + 32) & size) == 0);
            k0 = (lower == asc0) ? min(k0, o0) : max(k0, o0);
            k1 = (lower == asc1) ? min(k1, o1) : max(k1, o1);
          }
        }
      }
      int q0 = __shfl_up_sync(0xffffffffu, k0, 1), q1 = __shfl_up_sync(0xffffffffu, k1, 1);
      const int last0 = __shfl_sync(0xffffffffu, k0, 31);
      if (lane == 0) { q0 = -1; q1 = last0; }
      const bool u0 = (k0 != 0x7fffffff) && (k0 != q0), u1 = (k1 != 0x7fffffff) && (k1 != q1);
      const unsigned b0 = __ballot_sync(0xffffffffu, u0), b1 = __ballot_sync(0xffffffffu, u1);
      if (!FILL) {
        if (lane == 0) rowlen[v] = __popc(b0) + __popc(b1);
      } else {
        const int64_t base = gptr[v];
        const unsigned lt = (1u << lane) - 1u;
        if (u0) {
          const int r = __popc(b0 & lt);
          gcol[base + r] = k0;
          if (k0 == (int32_t)v) diag_slot[v] = (int32_t)(base + r);
        }
        if (u1) {
          const int r = __popc(b0) + __popc(b1 & lt);
          gcol[base + r] = k1;
          if (k1 == (int32_t)v) diag_slot[v] = (int32_t)(base + r);
        }
      }
      __syncwarp();
      continue;
    }
    int count = 0;
    for (int i = lane; i < n; i += 32) {
      int32_t c = cand[i];
      bool is_first = true;
      for (int t = 0; t < i; ++t) is_first &= (cand[t] != c);
      first[i] = is_first;
      count += is_first;
    }
    __syncwarp();
    if (!FILL) {
      for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
      if (lane == 0) rowlen[v] = count;
    } else {
      int64_t base = gptr[v];
      for (int i = lane; i < n; i += 32) {
        if (!first[i]) continue;
        int32_t c = cand[i];
        int rank = 0;
        for (int t = 0; t < n; ++t) rank += (first[t] && cand[t] < c);
        gcol[base + rank] = c;
        if (c == (int32_t)v) diag_slot[v] = (int32_t)(base + rank);
      }
    }
    __syncwarp();
  }
}

// slot map: one thread per (element, local row i)
__global__ void k_slot_map(MeshView m, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                           uint8_t *__restrict__ slotpos, int *__restrict__ error_flag) {
  int64_t total = m.nE * (int64_t)kMaxM;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = t / kMaxM;
    int i = (int)(t % kMaxM);
    int mm = shape_m(m.code(e));
    if (i >= mm) continue;
    int64_t b = m.begin(e);
    int row = m.conn[b + i];
    int64_t r0 = gptr[row];
    int len = (int)(gptr[row + 1] - r0);
    int64_t sb = m.slot_begin(e) + (int64_t)i * mm;
    for (int j = 0; j < mm; ++j) {
      int c = m.conn[b + j];
      int lo = 0, hi = len - 1, pos = -1;
      while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int cv = gcol[r0 + mid];
        if (cv == c) { pos = mid; break; }
        if (cv < c) lo = mid + 1; else hi = mid - 1;
      }
      if (pos < 0 || pos > 255) { atomicExch(error_flag, 2); pos = 0; }
      slotpos[sb + j] = (uint8_t)pos;
    }
  }
}

__global__ void k_max_i32(const int32_t *__restrict__ a, int64_t n, int *__restrict__ out) {
  int mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mx = max(mx, a[i]);
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

void system_build_graph(efv_system *s) {
  efv_mesh *m = s->mesh;
  MeshView mv = view_of(m);
  s->nV = m->nV;
  s->n_owned = m->nV;
  s->rows = m->nV * s->ndof;
  DBuf<int32_t> rowlen(m->nV);
  DBuf<int> err(2);
  err.zero();
  int grid = grid_for(m->nV * 32, kRowWarps * 32, 8);
  EFV_LAUNCH(k_graph_rows<false>, grid, kRowWarps * 32, 0, mv, rowlen.p, (const int64_t *)nullptr, (int32_t *)nullptr,
             (int32_t *)nullptr, err.p);
  EFV_LAUNCH(k_max_i32, grid_for(m->nV, 256), 256, 0, rowlen.p, m->nV, err.p + 1);
  s->gptr.alloc(m->nV + 1);
  s->nnzg = exclusive_scan_i32_to_i64(rowlen.p, s->gptr.p, m->nV);
  int herr[2];
  err.download(herr, 2);
  EFV_REQUIRE(herr[0] == 0, "a vertex has more than 1024 candidate neighbours");
  s->max_row_len = herr[1];
  EFV_REQUIRE(s->max_row_len <= 256, "graph rows longer than 256 entries are not supported by the byte slot map");
  EFV_REQUIRE(s->nnzg < (int64_t(1) << 31), "graph nnz exceeds int32 slot range");
  s->gcol.alloc(s->nnzg);
  s->diag_slot.alloc(m->nV);
  EFV_LAUNCH(k_graph_rows<true>, grid, kRowWarps * 32, 0, mv, (int32_t *)nullptr, s->gptr.p, s->gcol.p, s->diag_slot.p, err.p);
  // slot map
  s->slot_size = 0;
  for (int c = 0; c < kNumShapes; ++c) s->slot_size += m->groups[c].n * shape_m(c) * shape_m(c);
  s->slotpos.alloc(s->slot_size);
  EFV_LAUNCH(k_slot_map, grid_for(m->nE * kMaxM, 256), 256, 0, mv, s->gptr.p, s->gcol.p, s->slotpos.p, err.p);
  err.download(herr, 2);
  EFV_REQUIRE(herr[0] == 0, "slot map construction failed (column not found in its row)");
}

// ---- exports -----------------------------------------------------------------------------------------------------
// DOF-expanded field-major CSR exactly like LinearSystemCSR.initialize (LinearSystem.py:114-128) with sorted stencils
__global__ void k_expand_csr(int64_t nV, int ndof, int64_t nnzg, const int64_t *__restrict__ gptr,
                             const int32_t *__restrict__ gcol, int64_t *__restrict__ indptr,
                             int32_t *__restrict__ indices) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
    int64_t r0 = gptr[v];
    int len = (int)(gptr[v + 1] - r0);
    for (int i = 0; i < ndof; ++i) {
      int64_t start = (int64_t)ndof * ((int64_t)i * nnzg + r0);
      indptr[(int64_t)i * nV + v] = start;
      for (int j = 0; j < ndof; ++j)
        for (int k = 0; k < len; ++k) indices[start + (int64_t)j * len + k] = (int32_t)(gcol[r0 + k] + (int64_t)j * nV);
    }
    if (v == nV - 1) indptr[(int64_t)ndof * nV] = (int64_t)ndof * ndof * nnzg;
  }
}

void export_csr(const efv_system *s, int64_t *indptr, int32_t *indices) {
  int64_t nnz = s->nnzg * s->ndof * s->ndof;
  EFV_REQUIRE((int64_t)s->ndof * s->nV < (int64_t(1) << 31), "DOF count exceeds int32 column range");
  DBuf<int64_t> dp(s->rows + 1);
  DBuf<int32_t> di(nnz);
  EFV_LAUNCH(k_expand_csr, grid_for(s->nV, 128), 128, 0, s->nV, s->ndof, s->nnzg, s->gptr.p, s->gcol.p, dp.p, di.p);
  if (indptr) dp.download(indptr, s->rows + 1);
  if (indices) di.download(indices, nnz);
}

__global__ void k_export_slots(MeshView m, const int64_t *__restrict__ gptr, const uint8_t *__restrict__ slotpos,
                               int64_t *__restrict__ out) {
  int64_t total = m.nE * (int64_t)kMaxM;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = t / kMaxM;
    int i = (int)(t % kMaxM);
    int mm = shape_m(m.code(e));
    if (i >= mm) continue;
    int row = m.conn[m.begin(e) + i];
    int64_t sb = m.slot_begin(e) + (int64_t)i * mm;
    for (int j = 0; j < mm; ++j) out[sb + j] = gptr[row] + slotpos[sb + j];
  }
}

void export_slotmap(const efv_system *s, int64_t *slots) {
  DBuf<int64_t> d(s->slot_size);
  MeshView mv = view_of(s->mesh);
  EFV_LAUNCH(k_export_slots, grid_for(mv.nE * kMaxM, 256), 256, 0, mv, s->gptr.p, s->slotpos.p, d.p);
  d.download(slots, s->slot_size);
}

}  // namespace efv
