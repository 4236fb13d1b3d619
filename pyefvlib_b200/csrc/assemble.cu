// assemble.cu -- north_star kernels 3 and 4: fused FP64 assembly with a deterministic row-wise reduction,
// Dirichlet row replacement / symmetric elimination, Neumann facet terms, right-hand sides.
//
// Assembly is ROW-CENTRIC: one owner per graph row (vertex v) walks the (element e, local vertex l) pairs of v in
// ascending (local vertex, element) order, evaluates the inner faces that touch l from the stored geometry, and adds
// the coefficient of every column conn[e][j] into a shared-memory copy of the row at the position given by the byte
// slot map.  Every CSR entry is therefore summed by exactly one owner in a fixed order: no float atomics,
// bit-reproducible, values written with coalesced stores.  (apps/heat_transfer.py:41-68)
// The owner is a LANE in the default kernels (k_heat_lanes: a warp = 32 rows; k_elas_lanes: a warp = 8 vertices x 4 lanes)
// and an 8-lane group in the round-1 row-group kernels (k_heat_rows, k_elasticity_rows, k_biot_rows), which remain the
// fall-back and the Biot path.
#include "views.cuh"
#include <cstdlib>
#include <cstring>

namespace efv {

struct SysView {
  int ndof;
  int64_t nV, nnzg;
  const int64_t *gptr;
  const int32_t *gcol;
  const uint8_t *slotpos;
  const int32_t *diag_slot;
  double *vals;
};
static SysView sys_view(efv_system *s) {
  SysView v;
  v.ndof = s->ndof; v.nV = s->n_owned; v.nnzg = s->nnzg;     // row kernels only visit owned rows
  v.gptr = s->gptr.p; v.gcol = s->gcol.p; v.slotpos = s->slotpos.p; v.diag_slot = s->diag_slot.p; v.vals = s->vals.p;
  return v;
}

// shared-memory copy of the inner-face derivative tables of all shapes (dynamic face index per lane)
struct SmemTables {
  double ifD_tri[3][3][2], ifD_quad[4][4][2], ifD_tet[6][4][3], ifD_hex[12][8][3], ifD_prism[9][6][3], ifD_pyr[8][5][3];
  int vface_tri[3][2], vface_quad[4][2], vface_tet[4][3], vface_hex[8][3], vface_prism[6][3], vface_pyr[5][4];
};
__device__ __forceinline__ void load_tables(SmemTables &T) {
  for (int i = threadIdx.x; i < 3 * 3 * 2; i += blockDim.x) (&T.ifD_tri[0][0][0])[i] = (&c_Triangle_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 4 * 4 * 2; i += blockDim.x) (&T.ifD_quad[0][0][0])[i] = (&c_Quadrilateral_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 6 * 4 * 3; i += blockDim.x) (&T.ifD_tet[0][0][0])[i] = (&c_Tetrahedron_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 12 * 8 * 3; i += blockDim.x) (&T.ifD_hex[0][0][0])[i] = (&c_Hexahedron_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 3 * 2; i += blockDim.x) (&T.vface_tri[0][0])[i] = (&c_Triangle_vface[0][0])[i];
  for (int i = threadIdx.x; i < 4 * 2; i += blockDim.x) (&T.vface_quad[0][0])[i] = (&c_Quadrilateral_vface[0][0])[i];
  for (int i = threadIdx.x; i < 4 * 3; i += blockDim.x) (&T.vface_tet[0][0])[i] = (&c_Tetrahedron_vface[0][0])[i];
  for (int i = threadIdx.x; i < 8 * 3; i += blockDim.x) (&T.vface_hex[0][0])[i] = (&c_Hexahedron_vface[0][0])[i];
  for (int i = threadIdx.x; i < 9 * 6 * 3; i += blockDim.x) (&T.ifD_prism[0][0][0])[i] = (&c_Prism_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 8 * 5 * 3; i += blockDim.x) (&T.ifD_pyr[0][0][0])[i] = (&c_Pyramid_ifD[0][0][0])[i];
  for (int i = threadIdx.x; i < 6 * 3; i += blockDim.x) (&T.vface_prism[0][0])[i] = (&c_Prism_vface[0][0])[i];
  for (int i = threadIdx.x; i < 5 * 4; i += blockDim.x) (&T.vface_pyr[0][0])[i] = (&c_Pyramid_vface[0][0])[i];
  __syncthreads();
}
__device__ __forceinline__ int tab_vface(const SmemTables &T, int code, int l, int q) {
  switch (code) {
    case 0: return T.vface_tri[l][q];
    case 1: return T.vface_quad[l][q];
    case 2: return T.vface_tet[l][q];
    case 3: return T.vface_hex[l][q];
    case 4: return T.vface_prism[l][q];
    default: return T.vface_pyr[l][q];           // -1: a base vertex of a pyramid touches only 3 faces
  }
}
__device__ __forceinline__ double tab_ifD(const SmemTables &T, int code, int f, int j, int a) {
  switch (code) {
    case 0: return T.ifD_tri[f][j][a];
    case 1: return T.ifD_quad[f][j][a];
    case 2: return T.ifD_tet[f][j][a];
    case 3: return T.ifD_hex[f][j][a];
    case 4: return T.ifD_prism[f][j][a];
    default: return T.ifD_pyr[f][j][a];
  }
}

// per-pair descriptor of the row-group kernels: bit 0 valid | shape code (3 bits) | up to 4 faces (4 bits each) |
// their "this vertex is the forward one" bits.  A missing 4th face (pyramid base vertices) is staged as zeros.
constexpr int kMaxVF = 4;
__device__ __forceinline__ int mt_code(int mt) { return (mt >> 1) & 7; }
__device__ __forceinline__ int mt_face(int mt, int k) { return (mt >> (4 + 4 * k)) & 15; }
__device__ __forceinline__ int mt_fwd(int mt, int k) { return (mt >> (20 + k)) & 1; }
__device__ __forceinline__ int mt_pack_face(int enc, int k) { return enc < 0 ? 0 : (((enc >> 1) << (4 + 4 * k)) | ((enc & 1) << (20 + k))); }

// ---- heat ---------------------------------------------------------------------------------------------------------
// One group of 8 lanes per graph row.  The incidence list of the row is processed in chunks of 8 (element,
// local vertex) pairs:
//   load phase   : lane q fetches everything pair q needs -- the 3 (2 in 2-D) face vectors  J^-T s  that touch the
//                  local vertex (pre-multiplied by +-k), the byte slot positions of the m columns, the sub-control
//                  volume -- and parks it in shared memory.  All loads of a chunk are independent => deep MLP.
//   reduce phase : pairs are consumed in ascending element order; lane j turns the face vectors into the
//                  coefficient of column conn[e][j] (3 FMAs per face against the shape table) and adds it to the
//                  shared-memory row at slotpos[e][l][j].  Fixed order, one owner per entry: deterministic.
// Epilogue: accumulation term on the diagonal, optional Dirichlet treatment, coalesced store of the row.
// props per region: [0] conductivity, [1] rho*cp/dt (0 when steady), [2] heat generation
constexpr int kAsmGroup = 8;
constexpr int kAsmRows = 256 / kAsmGroup;

template <int UCODE>
__global__ void __launch_bounds__(256) k_heat_rows(MeshView m, GeomView g, SysView s, const double *__restrict__ props,
                                                   int transient, int max_row_len, int dir_mode,
                                                   const uint8_t *__restrict__ dflag, const double *__restrict__ dval,
                                                   double *__restrict__ gen_out, double *__restrict__ acc_out,
                                                   double *__restrict__ elim_out) {
  extern __shared__ double smem_dyn[];
  __shared__ SmemTables T;
  __shared__ double tbuf[kAsmRows][kAsmGroup][3 * kMaxVF];
  __shared__ double volq[kAsmRows][kAsmGroup][2];
  __shared__ unsigned long long posb[kAsmRows][kAsmGroup];
  __shared__ int meta[kAsmRows][kAsmGroup];
  load_tables(T);
  const int grp = threadIdx.x / kAsmGroup, gl = threadIdx.x % kAsmGroup;
  double *row = smem_dyn + (size_t)grp * max_row_len;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) / kAsmGroup * kAsmGroup);
  for (int64_t v0 = (int64_t)blockIdx.x * kAsmRows; v0 < s.nV; v0 += (int64_t)gridDim.x * kAsmRows) {
    const int64_t v = v0 + grp;
    if (v >= s.nV) continue;               // whole groups drop out together; syncs below are group-scoped
    const int64_t r0 = s.gptr[v];
    const int len = (int)(s.gptr[v + 1] - r0);
    for (int t = gl; t < len; t += kAsmGroup) row[t] = 0.0;
    double gen = 0.0, accv = 0.0;
    const int64_t p0 = m.inc_ptr[v], p1 = m.inc_ptr[v + 1];
    for (int64_t pc = p0; pc < p1; pc += kAsmGroup) {
      // ---- load phase ------------------------------------------------------------------------------------------
      const int64_t p = pc + gl;
      if (p < p1) {
        const uint32_t pk = m.inc[p];
        const int64_t e = inc_elem(pk);
        const int l = inc_local(pk);
        const int code = (UCODE >= 0) ? UCODE : m.code(e);
        const int mm = shape_m(code), nvf = shape_nvf(code), d = shape_dim(code);
        const int64_t gi = (UCODE >= 0) ? e : m.group_index(e);
        const int64_t n = g.n[code];
        const double *pr = props + 3 * m.region[e];
        const double cond = pr[0];
        const double *St = g.St[code];
        int mt = 1 | (code << 1);
#pragma unroll
        for (int q = 0; q < kMaxVF; ++q) {
          if (q < nvf) {
            const int enc = tab_vface(T, code, l, q);
            const int f = enc < 0 ? 0 : enc >> 1;
            const double sg = enc < 0 ? 0.0 : ((enc & 1) ? cond : -cond);     // forward vertex +flux, backward -flux (heat_transfer.py:52-54)
            mt |= mt_pack_face(enc, q);
#pragma unroll
            for (int a = 0; a < 3; ++a) tbuf[grp][gl][q * 3 + a] = (a < d) ? sg * St[(int64_t)(f * d + a) * n + gi] : 0.0;
          }
        }
        meta[grp][gl] = mt;
        const uint8_t *sp = s.slotpos + ((UCODE >= 0) ? e * (int64_t)(mm * mm) : m.slot_begin(e)) + l * mm;
        unsigned long long pb;
        if (UCODE >= 0 && mm == 8) pb = *reinterpret_cast<const unsigned long long *>(sp);      // aligned: e*64 + l*8
        else if (UCODE >= 0 && mm == 4) pb = *reinterpret_cast<const unsigned int *>(sp);       // aligned: e*16 + l*4
        else {
          pb = 0;
          for (int j = 0; j < mm; ++j) pb |= (unsigned long long)sp[j] << (8 * j);
        }
        posb[grp][gl] = pb;
        const double vol = g.vol[code][(int64_t)l * n + gi];
        volq[grp][gl][0] = vol * pr[2];
        volq[grp][gl][1] = vol * pr[1];
      }
      __syncwarp(gmask);
      // ---- reduce phase (ascending element order) ------------------------------------------------------------------
      const int cnt = (int)((p1 - pc < kAsmGroup) ? (p1 - pc) : kAsmGroup);
      for (int q = 0; q < cnt; ++q) {
        const int mt = meta[grp][q];
        const int code = (UCODE >= 0) ? UCODE : mt_code(mt);
        const int mm = shape_m(code), nvf = shape_nvf(code);
        if (gl < mm) {
          double w = 0.0;
#pragma unroll
          for (int k = 0; k < kMaxVF; ++k) {
            if (k < nvf) {
              const int f = mt_face(mt, k);
              w += tab_ifD(T, code, f, gl, 0) * tbuf[grp][q][k * 3] + tab_ifD(T, code, f, gl, 1) * tbuf[grp][q][k * 3 + 1];
              if (code >= 2) w += tab_ifD(T, code, f, gl, 2) * tbuf[grp][q][k * 3 + 2];
            }
          }
          const int pos = (int)((posb[grp][q] >> (8 * gl)) & 0xff);
          row[pos] += w;
        }
        if (gl == 0) { gen += volq[grp][q][0]; accv += volq[grp][q][1]; }
      }
      __syncwarp(gmask);
    }
    // ---- epilogue ------------------------------------------------------------------------------------------------
    if (gl == 0) {
      if (transient) row[s.diag_slot[v] - r0] += accv;       // heat_transfer.py:62-67
      gen_out[v] = gen;
      acc_out[v] = accv;
    }
    __syncwarp(gmask);
    if (dir_mode < 0) {
      for (int t = gl; t < len; t += kAsmGroup) s.vals[r0 + t] = row[t];
    } else {
      const bool rowd = dflag[v];
      double el = 0.0;
      for (int t = gl; t < len; t += kAsmGroup) {
        const int c = s.gcol[r0 + t];
        double a = row[t];
        if (rowd) a = (c == v) ? 1.0 : 0.0;                    // heat_transfer.py:70-76
        else if (dir_mode == 1 && dflag[c]) { el -= a * dval[c]; a = 0.0; }
        s.vals[r0 + t] = a;
      }
      if (dir_mode == 1) {
#pragma unroll
        for (int o = kAsmGroup / 2; o > 0; o >>= 1) el += __shfl_xor_sync(gmask, el, o);
        if (gl == 0) elim_out[v] = rowd ? 0.0 : el;
      }
    }
    __syncwarp(gmask);
  }
}

template <int UCODE>
static void launch_heat_rows(efv_system *s, const double *dprops, int transient) {
  efv_mesh *m = s->mesh;
  MeshView mv = view_of(m);
  GeomView gv = geom_of(m);
  SysView sv = sys_view(s);
  size_t smem = (size_t)kAsmRows * s->max_row_len * sizeof(double);
  EFV_CUDA(cudaFuncSetAttribute(k_heat_rows<UCODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EFV_LAUNCH(k_heat_rows<UCODE>, grid_for(s->n_owned * kAsmGroup, 256, 6), 256, smem, mv, gv, sv, dprops, transient,
             s->max_row_len, s->fuse_mode, s->dir_flag.p, s->dir_value.p, s->gen.p, s->acc.p, s->elim.p);
}

// rows whose columns contain a Dirichlet DOF: the fused assembly epilogue only inspects the columns of those
// (the mesh graph is structurally symmetric, so the neighbours of row d are the columns of row d)
__global__ void k_mark_scatter(const int64_t *__restrict__ rows, int64_t n, const int64_t *__restrict__ gptr,
                               const int32_t *__restrict__ gcol, uint8_t *__restrict__ mark) {
  const int lane = threadIdx.x & 7;
  for (int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; k < n; k += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int64_t v = rows[k];
    for (int64_t t = gptr[v] + lane; t < gptr[v + 1]; t += 8) mark[gcol[t]] = 1;
  }
}
__global__ void k_mark_rows(int64_t nrows, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                            const uint8_t *__restrict__ flag, uint8_t *__restrict__ mark) {
  const int lane = threadIdx.x & 7;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  const int64_t rounds = (nrows + groups - 1) / groups;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t v = it * groups + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3);
    int any = 0;
    if (v < nrows)
      for (int64_t t = gptr[v] + lane; t < gptr[v + 1]; t += 8) any |= flag[gcol[t]];
    any = __any_sync(gmask, any);
    if (v < nrows && lane == 0) mark[v] = (uint8_t)(any != 0);
  }
}

// ---- lane-per-row kernel (default heat assembly, round 2) --------------------------------------------------------------
// k_heat_rows spends its time in the load/store unit (every lane of a row group re-reads staged operands and table entries
// from shared memory for every (row, element) pair; the round-1 tile kernel that replaced it still needed ~7 500 warp
// instructions per 32 rows for bucketing, slab traffic and three block barriers: 4.25 ms at C2).  k_heat_lanes has no
// block-level phase at all: a WARP owns a tile of 32 consecutive rows and LANE r owns row r.
//  * pair schedule (mesh-level, built once: ensure_pair_schedule): the tile's (element, local vertex) pairs are stored
//    [step][lane]; one step holds, for every lane, its next pair of ONE (shape, local vertex) key or a "none" marker.  The
//    key is warp-uniform, so the pair arithmetic is statically unrolled per local vertex with the table entries as
//    immediates and nothing diverges; descriptors are read with coalesced loads.  Per row the order stays ascending
//    (local vertex, shape, element): every entry is summed by one lane in a fixed order -- no atomics, bit-reproducible,
//    and for single-shape meshes the same order as the row-group kernel and the partitioned runs.
//  * software pipeline over the warp's steps (across tile boundaries): the operands of step i+1 (3 face vectors J^-T s,
//    sub-volume, slot bytes: 11 independent loads per lane) are requested into registers while step i computes -- the
//    scaled face vectors of step i are taken out of the operand registers first, so one set is enough -- and the
//    descriptors run 3 steps ahead through a small cp.async ring in shared memory; the row data of the next tile is
//    requested a whole tile ahead.  Loaded values are never touched before the step that consumes them (any use would
//    wait for the load on the spot: the first cut lost 35 % of its time in one such SEL).
//  * accumulator: a warp-private shared-memory image of the tile's CSR value segment (rows of a tile are contiguous in the
//    CSR array); lane r adds coefficient j at  row_start(r) + slotpos[j].  The write-back is then a flat, fully
//    coalesced streaming copy of that image; rows with a Dirichlet neighbour are fixed up in shared memory first.
// A variant that staged the operands in shared memory with per-lane cp.async (deeper pipeline, fewer registers) measured
// 1.73 ms against 1.62 ms: a cp.async lands one 32-byte sector per shared-memory wavefront, so its 11 copies per step cost
// ~90 wavefronts where 11 register loads cost 22 (LSU data pipe 75 % busy, profiles/r02_assembly_lanes.json).
// descriptor word: local vertex << 29 | last step of its tile << 28 | [mixed meshes: shape code << 25] | element
constexpr int kPdLocalShift = 29, kPdLastBit = 28, kPdCodeShift = 25;
__host__ __device__ constexpr uint32_t pd_elem_mask(bool mixed) { return (1u << (mixed ? kPdCodeShift : kPdLastBit)) - 1u; }
__device__ __forceinline__ bool pd_none(uint32_t pk, bool mixed) { return (pk & pd_elem_mask(mixed)) == pd_elem_mask(mixed); }
__device__ __forceinline__ bool pd_last(uint32_t pk) { return (pk >> kPdLastBit) & 1u; }
__device__ __forceinline__ int pd_key(uint32_t pk, bool mixed) {                    // shape code * 8 + local vertex
  return mixed ? (int)(((pk >> kPdCodeShift) & 7u) * 8u + (pk >> kPdLocalShift)) : (int)(pk >> kPdLocalShift);
}

// ---- pair schedule ---------------------------------------------------------------------------------------------------
// one warp per tile, lane = vertex.  The incidence list of a vertex is ascending in (local vertex, element); for mixed
// meshes the pairs of one local vertex are visited shape by shape.  FILL = false: steps of the tile; true: write them.
template <bool FILL, int R>
__global__ void __launch_bounds__(128) k_sched_build(MeshView m, int64_t ntiles, int32_t *__restrict__ tile_steps,
                                                     const int64_t *__restrict__ off, uint32_t *__restrict__ spk) {
  // R rows per tile (R divides 32): a warp builds 32 / R tiles, one group of R lanes each
  const int lane = threadIdx.x & 31;
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t tile = gt / R;
  const int r = (int)(gt % R);
  if (tile >= ntiles) return;
  const unsigned gmask = (R == 32) ? 0xffffffffu : (((1u << (R & 31)) - 1u) << (lane & ~(R - 1)));
  const int64_t v = tile * R + r;
  int64_t p = 0, p1 = 0;
  if (v < m.nV) { p = m.inc_ptr[v]; p1 = m.inc_ptr[v + 1]; }
  const bool mixed = m.uniform_code < 0;
  const uint32_t emask = pd_elem_mask(mixed);
  int64_t st = 0, st_end = 0;
  if (FILL) { st = off[tile]; st_end = off[tile + 1]; }
  int steps = 0;
  for (int l = 0; l < kMaxM; ++l) {
    int64_t q = p;                                               // [p, q): this vertex's pairs with local vertex l
    while (q < p1 && inc_local(m.inc[q]) == l) ++q;
    for (int code = mixed ? 0 : m.uniform_code; code < (mixed ? kNumShapes : m.uniform_code + 1); ++code) {
      int cnt = 0;
      if (!mixed) cnt = (int)(q - p);
      else for (int64_t t = p; t < q; ++t) cnt += m.code(inc_elem(m.inc[t])) == code;
      const int nst = __reduce_max_sync(gmask, cnt);
      if (FILL) {
        int64_t t = p;
        for (int i = 0; i < nst; ++i) {
          uint32_t pk = emask;                                   // no pair for this row in this step
          if (i < cnt) {
            if (mixed) while (m.code(inc_elem(m.inc[t])) != code) ++t;
            pk = (uint32_t)inc_elem(m.inc[t++]);
          }
          pk |= (uint32_t)l << kPdLocalShift;
          if (mixed) pk |= (uint32_t)code << kPdCodeShift;
          if (st + i + 1 == st_end) pk |= 1u << kPdLastBit;
          spk[(st + i) * R + r] = pk;
        }
        st += nst;
      }
      steps += nst;
    }
    p = q;
  }
  if (steps == 0) {                                              // vertices without elements: one empty step keeps the pipeline uniform
    if (FILL) spk[st * R + r] = emask | (1u << kPdLastBit);
    steps = 1;
  }
  if (!FILL && r == 0) tile_steps[tile] = steps;
}
template <int R>
static bool build_pair_schedule(efv_mesh *m, DBuf<uint32_t> &pk, DBuf<int64_t> &off, int64_t &steps) {
  if (off.n) return true;
  if (m->nE >= (int64_t)pd_elem_mask(m->uniform_code < 0)) return false;     // element ids must fit below the "none" marker
  const int64_t ntiles = div_up(m->nV, R);
  MeshView mv = view_of(m);
  DBuf<int32_t> cnt(ntiles);
  EFV_LAUNCH((k_sched_build<false, R>), (unsigned)div_up(ntiles * R, 128), 128, 0, mv, ntiles, cnt.p, (const int64_t *)nullptr,
             (uint32_t *)nullptr);
  off.alloc(ntiles + 1);
  steps = exclusive_scan_i32_to_i64(cnt.p, off.p, ntiles);
  pk.alloc((size_t)steps * R);
  EFV_LAUNCH((k_sched_build<true, R>), (unsigned)div_up(ntiles * R, 128), 128, 0, mv, ntiles, (int32_t *)nullptr, off.p, pk.p);
  return true;
}
static bool ensure_pair_schedule(efv_mesh *m) { return build_pair_schedule<32>(m, m->sched_pk, m->sched_off, m->sched_steps); }
// tiles of 8 rows: the block (ndof x ndof) assembly gives every row 4 lanes
static bool ensure_pair_schedule8(efv_mesh *m) { return build_pair_schedule<8>(m, m->sched8_pk, m->sched8_off, m->sched8_steps); }

__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// key = shape code * 8 + local vertex (warp-uniform); UCODE >= 0 instantiates the 8 cases of one shape only (key = local vertex)
#define EFV_LANE_KEYS(Tab, B, OP)                                                              \
  case B + 0: if constexpr (Tab::M > 0) { OP(Tab, 0); } break;                                 \
  case B + 1: if constexpr (Tab::M > 1) { OP(Tab, 1); } break;                                 \
  case B + 2: if constexpr (Tab::M > 2) { OP(Tab, 2); } break;                                 \
  case B + 3: if constexpr (Tab::M > 3) { OP(Tab, 3); } break;                                 \
  case B + 4: if constexpr (Tab::M > 4) { OP(Tab, 4); } break;                                 \
  case B + 5: if constexpr (Tab::M > 5) { OP(Tab, 5); } break;                                 \
  case B + 6: if constexpr (Tab::M > 6) { OP(Tab, 6); } break;                                 \
  case B + 7: if constexpr (Tab::M > 7) { OP(Tab, 7); } break;
#define EFV_LANE_DISPATCH(UCODE, key, OP)                                                                      \
  do {                                                                                                         \
    if constexpr (UCODE == 0) { switch (key) { EFV_LANE_KEYS(TriangleTab, 0, OP) default: break; } }           \
    else if constexpr (UCODE == 1) { switch (key) { EFV_LANE_KEYS(QuadrilateralTab, 0, OP) default: break; } } \
    else if constexpr (UCODE == 2) { switch (key) { EFV_LANE_KEYS(TetrahedronTab, 0, OP) default: break; } }   \
    else if constexpr (UCODE == 3) { switch (key) { EFV_LANE_KEYS(HexahedronTab, 0, OP) default: break; } }    \
    else if constexpr (UCODE == 4) { switch (key) { EFV_LANE_KEYS(PrismTab, 0, OP) default: break; } }         \
    else if constexpr (UCODE == 5) { switch (key) { EFV_LANE_KEYS(PyramidTab, 0, OP) default: break; } }       \
    else {                                                                                                     \
      switch (key) {                                                                                           \
        EFV_LANE_KEYS(TriangleTab, 0, OP) EFV_LANE_KEYS(QuadrilateralTab, 8, OP) EFV_LANE_KEYS(TetrahedronTab, 16, OP) \
        EFV_LANE_KEYS(HexahedronTab, 24, OP) EFV_LANE_KEYS(PrismTab, 32, OP) EFV_LANE_KEYS(PyramidTab, 40, OP) \
        default: break;                                                                                        \
      }                                                                                                        \
    }                                                                                                          \
  } while (0)

extern __shared__ __align__(16) double lane_smem[];
constexpr int kLaneSmemBudget = 160 * 1024;   // shared memory per SM the lane kernels may fill: the copies / loads in flight live in L1
                                              // lines until they land, and with the 228 KB carve-out (28 KB of L1) they run 1.7x slower
// operands of one pair step (per lane): loaded for step i+1 while step i computes
struct LaneOps { double t[12]; double vol; unsigned lo, hi; };
template <class Tab, int L, bool U>
__device__ __forceinline__ void heat_reg_load(LaneOps &o, uint32_t pk, const MeshView &m, const GeomView &g, const SysView &s) {
  constexpr int M = Tab::M, D = Tab::DIM, NVF = Tab::NVF, CODE = Tab::CODE;
  if (pd_none(pk, !U)) return;
  const int64_t e = (int64_t)(pk & pd_elem_mask(!U));
  const int64_t gi = U ? e : m.group_index(e);
  unsigned nb = (unsigned)g.n[CODE] * 8u;                         // < 2^28 elements: plane offsets = 32 x 32 -> 64-bit products
  asm volatile("" : "+r"(nb));                                    // one IMAD.WIDE per address instead of hoisted 64-bit offsets
  const char *St = reinterpret_cast<const char *>(g.St[CODE] + gi);
#pragma unroll
  for (int k = 0; k < NVF; ++k) {
    const int f = Tab::vface_c(L, k) >> 1;
#pragma unroll
    for (int a = 0; a < D; ++a)
      o.t[k * D + a] = (Tab::vface_c(L, k) >= 0)
                           ? __ldg(reinterpret_cast<const double *>(St + (uint64_t)(unsigned)(f < 0 ? 0 : f * D + a) * nb)) : 0.0;
  }
  o.vol = __ldg(reinterpret_cast<const double *>(reinterpret_cast<const char *>(g.vol[CODE] + gi) + (uint64_t)(unsigned)L * nb));
  const uint8_t *sp = s.slotpos + (U ? e * (int64_t)(M * M) : m.slot_begin(e)) + L * M;
  if (U && M == 8) {
    const uint2 w = __ldg(reinterpret_cast<const uint2 *>(sp));
    o.lo = w.x; o.hi = w.y;
  } else if (U && M == 4) {
    o.lo = __ldg(reinterpret_cast<const unsigned *>(sp)); o.hi = 0;
  } else {
    unsigned long long pb = 0;
#pragma unroll
    for (int j = 0; j < M; ++j) pb |= (unsigned long long)sp[j] << (8 * j);
    o.lo = (unsigned)pb; o.hi = (unsigned)(pb >> 32);
  }
}
// tc = the step's face vectors already multiplied by the conductivity; the forward / backward sign of a face is folded into
// the table entries (exact: (-D) t = -(D t))
template <class Tab, int L, bool U>
__device__ __forceinline__ void heat_reg_compute(const double (&tc)[12], unsigned lo, unsigned hi, uint32_t pk, double *__restrict__ row) {
  constexpr int M = Tab::M, D = Tab::DIM, NVF = Tab::NVF;
  if (pd_none(pk, !U)) return;
#pragma unroll
  for (int j0 = 0; j0 < M; j0 += 4) {
    double w[4], a_[4];
    int pos[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < M) {
        pos[jj] = (int)(((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xffu);
        a_[jj] = row[pos[jj]];
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < M) {
        double wj = 0.0;
#pragma unroll
        for (int k = 0; k < NVF; ++k) {
          const int f = Tab::vface_c(L, k) >> 1;
#pragma unroll
          for (int a = 0; a < D; ++a)
            if (Tab::vface_c(L, k) >= 0 && Tab::ifD_c(f < 0 ? 0 : f, j, a) != 0.0)
              wj += ((Tab::vface_c(L, k) & 1) ? Tab::ifD_c(f < 0 ? 0 : f, j, a) : -Tab::ifD_c(f < 0 ? 0 : f, j, a)) * tc[k * D + a];
        }
        w[jj] = wj;
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (j0 + jj < M) row[pos[jj]] = a_[jj] + w[jj];
  }
}
constexpr int kLaneWarps = 4;               // warps per block (they never synchronise with each other)
template <int UCODE>
__global__ void __launch_bounds__(kLaneWarps * 32, 5) k_heat_lanes(const __grid_constant__ MeshView m, const __grid_constant__ GeomView g,
                                                                   const __grid_constant__ SysView s, const double *__restrict__ props,
                                                                   const uint32_t *__restrict__ spk, const int64_t *__restrict__ soff,
                                                                   int transient, int acc_cap, int dir_mode, int single_region_i,
                                                                   const uint8_t *__restrict__ dflag, const uint8_t *__restrict__ rmark,
                                                                   const double *__restrict__ dval,
                                                                   double *__restrict__ gen_out, double *__restrict__ acc_out,
                                                                   double *__restrict__ elim_out) {
  constexpr int RR = 4;                                                   // descriptor ring: steps i .. i+3
  constexpr bool U = UCODE >= 0;
  constexpr int NT = U ? shape_nvf(UCODE < 0 ? 0 : UCODE) * shape_dim(UCODE < 0 ? 0 : UCODE) : 12;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double *accw = lane_smem + (size_t)wib * (acc_cap + RR * 16);           // image of a tile's CSR value segment | descriptor ring
  const unsigned pkring_sa = (unsigned)__cvta_generic_to_shared(accw + acc_cap) + (unsigned)lane * 4u;
  const bool single_region = single_region_i != 0;
  const int ntiles = (int)((s.nV + 31) >> 5);
  const int stride = (int)gridDim.x * kLaneWarps;
  int tile = (int)blockIdx.x * kLaneWarps + wib;
  if (tile >= ntiles) return;
  const int tail_rows = (int)(s.nV - ((int64_t)(ntiles - 1) << 5));
  double cond = props[0], pr1 = props[1], pr2 = props[2];
  struct RowRaw { int64_t g0, g1; int dslot; uint8_t mark; };
  auto load_rows = [&](int t) {
    RowRaw rr{0, 0, 0, 0};
    const int64_t v = ((int64_t)t << 5) + lane;
    if (v < s.nV) {
      rr.g0 = s.gptr[v];
      rr.g1 = s.gptr[v + 1];
      rr.dslot = s.diag_slot[v];
      if (dir_mode >= 0) rr.mark = rmark[v];
    } else {
      rr.g0 = rr.g1 = s.gptr[s.nV];
    }
    return rr;
  };
  int tF = tile;
  int64_t f0 = soff[tF], f1 = soff[tF + 1];
  const uint32_t *pF = spk + f0 * 32 + lane;
  int remF = (int)(f1 - f0);
  int64_t nxt0 = 0, nxt1 = 0;
  if (tF + stride < ntiles) { nxt0 = soff[tF + stride]; nxt1 = soff[tF + stride + 1]; }
  bool okF = tF < ntiles - 1 || lane < tail_rows;
  auto fetch_desc = [&](int slot) {
    const unsigned dst = pkring_sa + (unsigned)slot * 128u;
    if (remF > 0 && okF) {
      cp_async4(dst, pF);
    } else {
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(pd_elem_mask(!U) | (remF == 1 ? 1u << kPdLastBit : 0u)) : "memory");
    }
    if (remF > 0) {
      pF += 32;
      if (--remF == 0) {
        tF += stride;
        if (tF < ntiles) {
          pF = spk + nxt0 * 32 + lane;
          remF = (int)(nxt1 - nxt0);
          okF = tF < ntiles - 1 || lane < tail_rows;
          if (tF + stride < ntiles) { nxt0 = soff[tF + stride]; nxt1 = soff[tF + stride + 1]; }
        }
      }
    }
  };
  auto desc = [&](int slot) {
    uint32_t pk;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pk) : "r"(pkring_sa + (unsigned)slot * 128u) : "memory");
    return pk;
  };
#define EFV_RLOAD_OP(Tab, LL) heat_reg_load<Tab, LL, U>(ops, pkN, m, g, s)
#define EFV_RCOMP_OP(Tab, LL) heat_reg_compute<Tab, LL, U>(tc, lo, hi, pkA, row)
  LaneOps ops;
#pragma unroll
  for (int n = 0; n < 12; ++n) ops.t[n] = 0.0;
  ops.vol = 0.0; ops.lo = ops.hi = 0;
  // prologue: descriptors of steps 0 .. 2, operands of step 0
  fetch_desc(0); fetch_desc(1); fetch_desc(2);
  cp_async_commit();
  RowRaw rr = load_rows(tile), rrn = rr;
  cp_async_wait<0>();
  uint32_t pkA = desc(0);
  {
    const uint32_t pkN = pkA;
    EFV_LANE_DISPATCH(UCODE, pd_key(pkN, !U), EFV_RLOAD_OP);
  }
  int dA = 0;
  bool first = true;
  double gen = 0.0, accv = 0.0;
  double *row = accw;
  int64_t base0 = 0;
  int total = 0;
  bool marked = false;
  while (true) {
    if (first) {                                                          // first step of a tile: clear the image
      if (tile + stride < ntiles) rrn = load_rows(tile + stride);         // and request the row data of the next tile
      base0 = __shfl_sync(0xffffffffu, rr.g0, 0);
      total = (int)(__shfl_sync(0xffffffffu, rr.g1, 31) - base0);
      marked = rr.mark != 0;
      for (int i = lane; i < total; i += 32) accw[i] = 0.0;
      row = accw + (int)(rr.g0 - base0);
      gen = 0.0; accv = 0.0;
      first = false;
      __syncwarp();
    }
    // take step i out of the operand registers
    if (!single_region && !pd_none(pkA, !U)) {
      const double *pr = props + 3 * m.region[pkA & pd_elem_mask(!U)];
      cond = pr[0]; pr1 = pr[1]; pr2 = pr[2];
    }
    double tc[12];
#pragma unroll
    for (int n = 0; n < NT; ++n) tc[n] = ops.t[n] * cond;
    if (!pd_none(pkA, !U)) {
      gen += __dmul_rn(ops.vol, pr2);                                     // products rounded first: the same bits as the tile kernel
      accv += __dmul_rn(ops.vol, pr1);
    }
    const unsigned lo = ops.lo, hi = ops.hi;
    // descriptor of step i+3, operands of step i+1
    int dN = dA + 1; if (dN == RR) dN = 0;
    int dF = dA + 3; if (dF >= RR) dF -= RR;
    fetch_desc(dF);
    cp_async_commit();
    cp_async_wait<1>();                                                   // descriptors up to step i+2 have landed
    const uint32_t pkN = desc(dN);
    EFV_LANE_DISPATCH(UCODE, pd_key(pkN, !U), EFV_RLOAD_OP);
    // compute step i
    EFV_LANE_DISPATCH(UCODE, pd_key(pkA, !U), EFV_RCOMP_OP);
    const bool last = pd_last(pkA);
    pkA = pkN;
    dA = dN;
    if (last) {
      // ---- tile epilogue ----------------------------------------------------------------------------------------------
      const int64_t v = ((int64_t)tile << 5) + lane;
      const int len = (int)(rr.g1 - rr.g0);
      if (v < s.nV) {
        if (transient) row[(int)((int64_t)rr.dslot - rr.g0)] += accv;     // heat_transfer.py:62-67
        gen_out[v] = gen;
        acc_out[v] = accv;
        if (dir_mode == 1 && !marked) elim_out[v] = 0.0;
      }
      __syncwarp();
      unsigned mm = __ballot_sync(0xffffffffu, marked);
      while (mm) {                                                        // rows with a Dirichlet DOF among their columns
        const int r = __ffs(mm) - 1;
        mm &= mm - 1;
        const int rrel = (int)(__shfl_sync(0xffffffffu, rr.g0, r) - base0);
        const int rlen = __shfl_sync(0xffffffffu, len, r);
        const int64_t vr = ((int64_t)tile << 5) + r;
        const bool rowd = dflag[vr];
        double el = 0.0;
        for (int t = lane; t < rlen; t += 32) {
          const int c = s.gcol[base0 + rrel + t];
          double a = accw[rrel + t];
          if (rowd) a = (c == vr) ? 1.0 : 0.0;                            // heat_transfer.py:70-76
          else if (dir_mode == 1 && dflag[c]) { el -= a * dval[c]; a = 0.0; }
          accw[rrel + t] = a;
        }
        if (dir_mode == 1) {
          el = warp_sum(el);
          if (lane == 0) elim_out[vr] = rowd ? 0.0 : el;
        }
      }
      __syncwarp();
      double *dst = s.vals + base0;
      for (int i = lane; i < total; i += 32) __stcs(dst + i, accw[i]);    // flat, coalesced, streaming copy of the segment
      __syncwarp();
      if (tile + stride >= ntiles) break;
      tile += stride;
      rr = rrn;
      first = true;
    }
  }
  cp_async_wait<0>();
#undef EFV_RLOAD_OP
#undef EFV_RCOMP_OP
}
template <int UCODE>
static bool launch_heat_lanes(efv_system *s, const double *dprops, int transient) {
  constexpr int W = kLaneWarps;
  const int acc_cap = 32 * s->max_row_len;
  const size_t smem = (size_t)W * (acc_cap + 4 * 16) * sizeof(double);
  if (smem + 1024 > (size_t)kLaneSmemBudget) return false;
  efv_mesh *m = s->mesh;
  if (!ensure_pair_schedule(m)) return false;
  MeshView mv = view_of(m);
  GeomView gv = geom_of(m);
  SysView sv = sys_view(s);
  EFV_CUDA(cudaFuncSetAttribute(k_heat_lanes<UCODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EFV_CUDA(cudaFuncSetAttribute(k_heat_lanes<UCODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 58));   // 132 KB: room for L1
  int per_sm = 0;
  EFV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_heat_lanes<UCODE>, W * 32, smem));
  if (per_sm < 1) return false;
  const int64_t ntiles = div_up(s->n_owned, 32);
  int64_t blocks = std::min<int64_t>(div_up(ntiles, W), (int64_t)num_sms() * per_sm);
  if (const char *e = getenv("EFV_LANE_GRID")) blocks = std::min<int64_t>(blocks, std::max(1, atoi(e)));   // tests: many tiles per warp
  EFV_LAUNCH((k_heat_lanes<UCODE>), (int)std::max<int64_t>(blocks, 1), W * 32, smem, mv, gv, sv, dprops, m->sched_pk.p,
             m->sched_off.p, transient, acc_cap, s->fuse_mode, m->n_regions == 1 ? 1 : 0, s->dir_flag.p,
             s->row_mark.p, s->dir_value.p, s->gen.p, s->acc.p, s->elim.p);
  return true;
}

void heat_assemble(efv_system *s, const double *props_host, double dt, int transient) {
  efv_mesh *m = s->mesh;
  EFV_REQUIRE(s->ndof == 1, "heat needs a 1-DOF system");
  geometry_build(m);
  if (!s->dir_flag.n) bc_clear(s);
  std::vector<double> pr(3 * m->n_regions);
  for (int r = 0; r < m->n_regions; ++r) {
    const double *p = props_host + 4 * r;   // Conductivity, Density, HeatCapacity, HeatGeneration
    pr[3 * r + 0] = p[0];
    pr[3 * r + 1] = transient ? p[1] * p[2] / dt : 0.0;    // heat_transfer.py:61
    pr[3 * r + 2] = p[3];
  }
  if (s->props_dev.n < pr.size()) s->props_dev.alloc(pr.size());
  s->props_dev.upload(pr.data(), pr.size());
  if (!s->gen.n) { s->gen.alloc(s->nV); s->acc.alloc(s->nV); }
  if (!s->ev0) { EFV_CUDA(cudaEventCreate(&s->ev0)); EFV_CUDA(cudaEventCreate(&s->ev1)); }
  if (s->fuse_mode >= 0 && !s->marks_valid) {
    EFV_LAUNCH(k_mark_rows, grid_for(s->n_owned * 8, 256), 256, 0, s->n_owned, s->gptr.p, s->gcol.p, s->dir_flag.p, s->row_mark.p);
    s->marks_valid = true;
  }
  EFV_CUDA(cudaEventRecord(s->ev0, stream()));
  // default: the lane-per-row kernel (k_heat_lanes); EFV_ASSEMBLY=rows selects the row-group kernel (k_heat_rows), which is
  // also the fall-back for rows too long for the shared-memory image and for element ids that do not fit in a descriptor.
  // Both sum every entry in the same order.
  const char *mode = getenv("EFV_ASSEMBLY");
  const bool want_rows = mode && strcmp(mode, "rows") == 0;
  bool done = false;
#define EFV_HEAT_BY_CODE(fn)                                                                  \
  (m->uniform_code == 0   ? fn<0>(s, s->props_dev.p, transient)                               \
   : m->uniform_code == 1 ? fn<1>(s, s->props_dev.p, transient)                               \
   : m->uniform_code == 2 ? fn<2>(s, s->props_dev.p, transient)                               \
   : m->uniform_code == 3 ? fn<3>(s, s->props_dev.p, transient)                               \
   : m->uniform_code == 4 ? fn<4>(s, s->props_dev.p, transient)                               \
   : m->uniform_code == 5 ? fn<5>(s, s->props_dev.p, transient)                               \
                          : fn<-1>(s, s->props_dev.p, transient))
  if (!want_rows) done = EFV_HEAT_BY_CODE(launch_heat_lanes);
#undef EFV_HEAT_BY_CODE
  if (!done) {
    switch (m->uniform_code) {
      case 0: launch_heat_rows<0>(s, s->props_dev.p, transient); break;
      case 1: launch_heat_rows<1>(s, s->props_dev.p, transient); break;
      case 2: launch_heat_rows<2>(s, s->props_dev.p, transient); break;
      case 3: launch_heat_rows<3>(s, s->props_dev.p, transient); break;
      default: launch_heat_rows<-1>(s, s->props_dev.p, transient); break;
    }
  }
  EFV_CUDA(cudaEventRecord(s->ev1, stream()));
  s->assemble_timed = true;
  s->vals_version++;
  s->dirichlet_mode = s->fuse_mode;
  s->physics = 0;
}

// ---- linear elasticity ----------------------------------------------------------------------------------------------
// Same row-centric scheme with an ndof x ndof block per graph entry.  For an inner face with area vector s and
// column vertex c with global shape gradient g = G_f[:, c] = J^-1 D_f[c, :] the reference's
//   M = At . Ce . Bv          (stress_equilibrium.py:50-67, :100;  Ce isotropic, :29-48)
// is, in closed form,  M[i][j] = lambda s_i g_j + mu (s_j g_i + delta_ij s.g).  Row of the backward vertex gets +M,
// row of the forward vertex -M (stress_equilibrium.py:105-117).  props per region: [0] lambda, [1] mu, [2] -rho*g.
constexpr int kElRows = 16;          // rows (8-lane groups) per 128-thread block
template <int UCODE, int NDOF>
__global__ void __launch_bounds__(128) k_elasticity_rows(MeshView m, GeomView g, SysView s, const double *__restrict__ props,
                                                         int gravity, int max_row_len, int dir_mode,
                                                         const uint8_t *__restrict__ dflag, const double *__restrict__ dval,
                                                         double *__restrict__ gen_out, double *__restrict__ elim_out) {
  extern __shared__ double smem_dyn[];
  __shared__ SmemTables T;
  // per pair: up to 4 faces x (J^-1 row-major 9 | s 3); 2-D uses 4 | 2 of each.  Lives behind the row accumulators in
  // dynamic shared memory (16 rows x 8 pairs x 48 doubles exceed the static limit)
  double(*fbuf)[kAsmGroup][12 * kMaxVF] = reinterpret_cast<double(*)[kAsmGroup][12 * kMaxVF]>(smem_dyn + (size_t)kElRows * max_row_len * NDOF * NDOF);
  __shared__ double aux[kElRows][kAsmGroup][3];        // lambda, mu, vol * (-rho g)
  __shared__ unsigned long long posb[kElRows][kAsmGroup];
  __shared__ int meta[kElRows][kAsmGroup];
  load_tables(T);
  constexpr int B = NDOF * NDOF;
  const int grp = threadIdx.x / kAsmGroup, gl = threadIdx.x % kAsmGroup;
  double *row = smem_dyn + (size_t)grp * max_row_len * B;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) / kAsmGroup * kAsmGroup);
  for (int64_t v0 = (int64_t)blockIdx.x * kElRows; v0 < s.nV; v0 += (int64_t)gridDim.x * kElRows) {
    const int64_t v = v0 + grp;
    if (v >= s.nV) continue;
    const int64_t r0 = s.gptr[v];
    const int len = (int)(s.gptr[v + 1] - r0);
    for (int t = gl; t < len * B; t += kAsmGroup) row[t] = 0.0;
    double grav = 0.0;
    const int64_t p0 = m.inc_ptr[v], p1 = m.inc_ptr[v + 1];
    for (int64_t pc = p0; pc < p1; pc += kAsmGroup) {
      const int64_t p = pc + gl;
      if (p < p1) {
        const uint32_t pk = m.inc[p];
        const int64_t e = inc_elem(pk);
        const int l = inc_local(pk);
        const int code = (UCODE >= 0) ? UCODE : m.code(e);
        const int mm = shape_m(code), nvf = shape_nvf(code), d = shape_dim(code);
        const int64_t gi = (UCODE >= 0) ? e : m.group_index(e);
        const int64_t n = g.n[code];
        const double *pr = props + 3 * m.region[e];
        const double *Ji = g.Jinv[code], *Sv = g.S[code];
        int mt = 1 | (code << 1);
        for (int q = 0; q < nvf; ++q) {
          const int enc = tab_vface(T, code, l, q);
          const int f = enc < 0 ? 0 : enc >> 1;
          mt |= mt_pack_face(enc, q);                  // forward bit: this vertex is the forward one -> sign -1
          double *fb = &fbuf[grp][gl][q * 12];
          for (int a = 0; a < d * d; ++a) fb[a] = enc < 0 ? 0.0 : Ji[(int64_t)(f * d * d + a) * n + gi];
          for (int a = 0; a < d; ++a) fb[9 + a] = enc < 0 ? 0.0 : Sv[(int64_t)(f * 3 + a) * n + gi];
        }
        meta[grp][gl] = mt;
        const uint8_t *sp = s.slotpos + ((UCODE >= 0) ? e * (int64_t)(mm * mm) : m.slot_begin(e)) + l * mm;
        unsigned long long pb;
        if (UCODE >= 0 && mm == 8) pb = *reinterpret_cast<const unsigned long long *>(sp);
        else if (UCODE >= 0 && mm == 4) pb = *reinterpret_cast<const unsigned int *>(sp);
        else {
          pb = 0;
          for (int j = 0; j < mm; ++j) pb |= (unsigned long long)sp[j] << (8 * j);
        }
        posb[grp][gl] = pb;
        aux[grp][gl][0] = pr[0];
        aux[grp][gl][1] = pr[1];
        aux[grp][gl][2] = g.vol[code][(int64_t)l * n + gi] * pr[2];
      }
      __syncwarp(gmask);
      const int cnt = (int)((p1 - pc < kAsmGroup) ? (p1 - pc) : kAsmGroup);
      for (int q = 0; q < cnt; ++q) {
        const int mt = meta[grp][q];
        const int code = (UCODE >= 0) ? UCODE : mt_code(mt);
        const int mm = shape_m(code), nvf = shape_nvf(code);
        if (gl < mm) {
          const double lam = aux[grp][q][0], mu = aux[grp][q][1];
          double M[NDOF][NDOF];
#pragma unroll
          for (int i = 0; i < NDOF; ++i)
#pragma unroll
            for (int j = 0; j < NDOF; ++j) M[i][j] = 0.0;
          for (int k = 0; k < nvf; ++k) {
            const int f = mt_face(mt, k);
            const double sgn = mt_fwd(mt, k) ? -1.0 : 1.0;
            const double *fb = &fbuf[grp][q][k * 12];
            double gv[NDOF], sv[NDOF], dl[NDOF];
#pragma unroll
            for (int a = 0; a < NDOF; ++a) { dl[a] = tab_ifD(T, code, f, gl, a); sv[a] = fb[9 + a]; }
            double gs = 0.0;
#pragma unroll
            for (int i = 0; i < NDOF; ++i) {
              double acc = 0.0;
#pragma unroll
              for (int a = 0; a < NDOF; ++a) acc += fb[i * NDOF + a] * dl[a];      // g_i = sum_a Jinv[i][a] D[c][a]
              gv[i] = acc;
              gs += acc * sv[i];
            }
#pragma unroll
            for (int i = 0; i < NDOF; ++i)
#pragma unroll
              for (int j = 0; j < NDOF; ++j)
                M[i][j] += sgn * (lam * sv[i] * gv[j] + mu * (sv[j] * gv[i] + ((i == j) ? gs : 0.0)));
          }
          const int pos = (int)((posb[grp][q] >> (8 * gl)) & 0xff);
          double *blk = row + pos * B;
#pragma unroll
          for (int i = 0; i < NDOF; ++i)
#pragma unroll
            for (int j = 0; j < NDOF; ++j) blk[i * NDOF + j] += M[i][j];
        }
        if (gl == 0) grav += aux[grp][q][2];
      }
      __syncwarp(gmask);
    }
    if (gl == 0) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) gen_out[v * NDOF + i] = (gravity && i == 1) ? grav : 0.0;   // y component only (:81-90)
    }
    // epilogue: Dirichlet treatment + store (block layout [slot][i][j])
    bool rowd[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) rowd[i] = (dir_mode >= 0) ? dflag[v * NDOF + i] : false;
    double el[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) el[i] = 0.0;
    for (int t = gl; t < len; t += kAsmGroup) {
      const int64_t c = s.gcol[r0 + t];
      double *blk = row + t * B;
      double *out = s.vals + (r0 + t) * B;
#pragma unroll
      for (int i = 0; i < NDOF; ++i)
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
          double a = blk[i * NDOF + j];
          if (rowd[i]) a = (c == v && i == j) ? 1.0 : 0.0;
          else if (dir_mode == 1 && dflag[c * NDOF + j]) { el[i] -= a * dval[c * NDOF + j]; a = 0.0; }
          out[i * NDOF + j] = a;
        }
    }
    if (dir_mode == 1) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) {
        double tot = el[i];
#pragma unroll
        for (int o = kAsmGroup / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(gmask, tot, o);
        if (gl == 0) elim_out[v * NDOF + i] = rowd[i] ? 0.0 : tot;
      }
    }
    __syncwarp(gmask);
  }
}

// ---- geometry records of the block (elasticity) lane kernel ------------------------------------------------------------
// The SoA planes of geometry.cu are coalesced when consecutive lanes hold consecutive elements; the block kernel gives a
// vertex 4 lanes that fetch ONE pair together, i.e. a warp touches 8 scattered elements, and every 8-byte read of a plane
// would pull its own 32-byte sector.  The face records put [J^-1 | s] of an (element, face) into 96 contiguous bytes = three
// full sectors; simplices get one compact record per element (below).
__global__ void k_face_records(int64_t n, int NF, int D, const double *__restrict__ Jinv, const double *__restrict__ S,
                               double *__restrict__ rec) {
  const int RF = D * D + D;
  const int64_t total = n * NF * RF;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(t % RF);              // output order: coalesced stores (one-time pass; the strided reads hit L2)
    const int f = (int)((t / RF) % NF);
    const int64_t k = t / ((int64_t)RF * NF);
    rec[t] = (a < D * D) ? Jinv[(int64_t)(f * D * D + a) * n + k] : S[(int64_t)(f * 3 + (a - D * D)) * n + k];
  }
}
// simplex records: [s_f (d doubles) for every inner face | J^-1 of the element (the first face's copy) | pad]
__host__ __device__ constexpr int simplex_rec_doubles(int code) { return code == 0 ? 10 : 28; }
__global__ void k_simplex_records(int64_t n, int NF, int D, int RS, const double *__restrict__ Jinv, const double *__restrict__ S,
                                  double *__restrict__ rec) {
  const int64_t total = n * RS;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(t % RS);
    const int64_t k = t / RS;
    double v = 0.0;
    if (a < NF * D) v = S[(int64_t)((a / D) * 3 + (a % D)) * n + k];
    else if (a < NF * D + D * D) v = Jinv[(int64_t)(a - NF * D) * n + k];
    rec[t] = v;
  }
}
static void ensure_simplex_records(efv_mesh *m) {
  for (int code = 0; code <= 2; code += 2) {
    ShapeGroup &g = m->groups[code];
    if (!g.n || g.Srec.n) continue;
    const int d = shape_dim(code), nf = shape_nf(code), rs = simplex_rec_doubles(code);
    g.Srec.alloc((size_t)g.n * rs);
    EFV_LAUNCH(k_simplex_records, grid_for(g.n * rs, 256), 256, 0, g.n, nf, d, rs, g.Jinv.p, g.S.p, g.Srec.p);
  }
}
static void ensure_face_records(efv_mesh *m, bool skip_simplices = false) {
  for (int code = 0; code < kNumShapes; ++code) {
    ShapeGroup &g = m->groups[code];
    if (!g.n || g.Frec.n) continue;
    if (skip_simplices && (code == 0 || code == 2)) continue;
    const int d = shape_dim(code), nf = shape_nf(code);
    g.Frec.alloc((size_t)g.n * nf * (d * d + d));
    EFV_LAUNCH(k_face_records, grid_for(g.n * nf * (d * d + d), 256), 256, 0, g.n, nf, d, g.Jinv.p, g.S.p, g.Frec.p);
  }
}
// ---- lane kernel for the block (elasticity) assembly ---------------------------------------------------------------------
// The scheme of k_heat_lanes with an ndof x ndof block per graph entry: a warp owns a tile of 8 consecutive vertices, every
// vertex gets 4 lanes, lane q < ndof owns component q of the vertex's block row (the three scalar rows of a vertex never
// touch the same value).  A step of the pair schedule (tiles of 8, ensure_pair_schedule8) holds one (element, local vertex)
// pair per vertex with a warp-uniform (shape, local vertex) key; the 4 lanes of a vertex split the asynchronous copies of
// the pair's face records [J^-1 | s] (16-byte cp.async) into the stage of their vertex and, after the group has landed
// and a __syncwarp, every compute lane reads all of it back (broadcast reads) and evaluates ITS row of
//   M = lambda s g^T + mu (g s^T + (s.g) I)          (stress_equilibrium.py:50-67, :100)
// for every column of the element.  Simplices (J^-1 and the shape derivatives are the same on every inner face) collapse
// their faces into ONE with the signed sum of the area vectors (M is linear in s).  The accumulator is the shared-
// memory image of the tile's BSR value segment; Dirichlet rows / columns are fixed up there and the segment leaves with a
// flat coalesced streaming copy.  Per scalar row the summation order is ascending (local vertex, shape, element).
constexpr int kBlkRows = 8, kBlkStages = 2, kBlkWarps = 2;
// doubles per vertex and stage: NVF face records + sub-volume + slot bytes, padded to an ODD number of 16-byte units so that
// the 8 vertices of a tile read their records from 8 different bank groups
// simplices: the element's whole compact record [s of every face | J^-1 | pad], then sub-volume and slot bytes
__host__ __device__ constexpr int blk_row_stride(int ucode, int dim) {
  const int rf = dim * dim + dim;
  const int slots = (ucode == 0 || ucode == 2) ? simplex_rec_doubles(ucode) + 2
                                               : (ucode < 0 ? (dim == 3 ? 4 : 2) : shape_nvf(ucode)) * rf + 2;
  const int units = (slots + 1) / 2;
  return 2 * (units | 1);
}
__host__ __device__ constexpr int blk_warp_doubles(int ucode, int dim, int acc_cap) {
  return acc_cap + kBlkStages * kBlkRows * blk_row_stride(ucode, dim) + 2 * kBlkStages * 16;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// copies of one pair into the stage row of its vertex (stg = shared address of that row); lane q of the vertex's 4 lanes
// takes every 4th 16-byte chunk, lane 0 also the sub-volume, lane 1 the slot bytes
template <class Tab, int L, bool U>
__device__ __forceinline__ void elas_lane_issue(unsigned stg, uint32_t pk, bool mx, int q, const MeshView &m, const GeomView &g,
                                                const SysView &s) {
  constexpr int M = Tab::M, D = Tab::DIM, NVF = Tab::NVF, CODE = Tab::CODE, RF = D * D + D, NF = Tab::NF;
  constexpr bool SIMPLEX = (CODE == 0 || CODE == 2);
  if (pd_none(pk, mx)) return;
  const int64_t e = (int64_t)(pk & pd_elem_mask(mx));
  const int64_t gi = U ? e : m.group_index(e);
  constexpr int VOFF = SIMPLEX ? simplex_rec_doubles(CODE) : NVF * RF;         // sub-volume, then the slot bytes
  // The 4 lanes of a vertex copy 4 CONSECUTIVE doubles per instruction = one 32-byte sector (records are sector-aligned): a
  // cp.async costs one shared-memory wavefront per sector it brings, whatever the lanes do with it.
  if constexpr (SIMPLEX) {
    constexpr int RS = simplex_rec_doubles(CODE);                              // the element's whole record [s of every face | J^-1 | pad]
    const double *rec = g.Srec[CODE] + (size_t)gi * RS;
#pragma unroll
    for (int i0 = 0; i0 < RS; i0 += 4)
      if (i0 + 3 < RS || q < RS - i0) cp_async8(stg + (i0 + q) * 8, rec + i0 + q);
  } else {
    const double *rec = g.Frec[CODE] + (size_t)gi * NF * RF;
#pragma unroll
    for (int k = 0; k < NVF; ++k) {
      const int enc = Tab::vface_c(L, k);
      const int f = enc < 0 ? 0 : enc >> 1;
#pragma unroll
      for (int i0 = 0; i0 < RF; i0 += 4)
        if (enc >= 0 && (i0 + 3 < RF || q < RF - i0)) cp_async8(stg + (k * RF + i0 + q) * 8, rec + f * RF + i0 + q);
    }
  }
  if (q == 0) {
    unsigned nb = (unsigned)g.n[CODE] * 8u;
    cp_async8(stg + VOFF * 8, reinterpret_cast<const char *>(g.vol[CODE] + gi) + (uint64_t)(unsigned)L * nb);
  }
  if (q == 1) {
    const uint8_t *sp = s.slotpos + (U ? e * (int64_t)(M * M) : m.slot_begin(e)) + L * M;
    if (U && M == 8) cp_async8(stg + (VOFF + 1) * 8, sp);
    else if (U && M == 4) cp_async4(stg + (VOFF + 1) * 8, sp);
    else {
      unsigned long long pb = 0;
#pragma unroll
      for (int j = 0; j < M; ++j) pb |= (unsigned long long)sp[j] << (8 * j);
      asm volatile("st.shared.b64 [%0], %1;" ::"r"(stg + (VOFF + 1) * 8), "l"(pb) : "memory");
    }
  }
}
// row q of every column block of the pair, added into the image (row = image of the vertex's block row, blocks row-major)
template <class Tab, int L, int NDOF, bool U>
__device__ __forceinline__ void elas_lane_compute(const double *__restrict__ stg, uint32_t pk, bool mx, int q, const double *__restrict__ pr,
                                                  double *__restrict__ row, double &grav) {
  constexpr int M = Tab::M, D = Tab::DIM, NVF = Tab::NVF, CODE = Tab::CODE, RF = D * D + D, B = NDOF * NDOF;
  static_assert(D == NDOF, "elasticity: one displacement component per space dimension");
  constexpr bool SIMPLEX = (CODE == 0 || CODE == 2);
  constexpr int NK = SIMPLEX ? 1 : NVF;
  if (pd_none(pk, mx) || q >= NDOF) return;
  const double lam = pr[0], mu = pr[1];
  double J[NK][D][D], ls[NK][D], ms[NK][D];        // lambda * s and mu * s (signed)
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int i = 0; i < D; ++i) { ls[k][i] = 0.0; ms[k][i] = 0.0; }
  constexpr int VOFF = SIMPLEX ? simplex_rec_doubles(CODE) : NVF * RF;
  if constexpr (SIMPLEX) {
    constexpr int JB = Tab::NF * D;                 // record = [s of every face | J^-1 | pad]
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int a = 0; a < D; ++a) J[0][i][a] = stg[JB + i * D + a];
#pragma unroll
    for (int k = 0; k < NVF; ++k) {
      const double sg = (Tab::vface_c(L, k) & 1) ? -1.0 : 1.0;
      const int f = Tab::vface_c(L, k) >> 1;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double sv = stg[f * D + i];
        ls[0][i] += lam * (sg * sv);
        ms[0][i] += mu * (sg * sv);
      }
    }
  } else {
    const double2 *st2 = reinterpret_cast<const double2 *>(stg);
#pragma unroll
    for (int k = 0; k < NVF; ++k) {
      const int enc = Tab::vface_c(L, k);
      const double sg = (enc & 1) ? -1.0 : 1.0;     // this vertex is the forward one of the face: -M (stress_equilibrium.py:105-117)
      double buf[RF];
#pragma unroll
      for (int p = 0; p < RF / 2; ++p) {
        double2 t2 = make_double2(0.0, 0.0);
        if (enc >= 0) t2 = st2[k * (RF / 2) + p];
        buf[2 * p] = t2.x; buf[2 * p + 1] = t2.y;
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int a = 0; a < D; ++a) J[k][i][a] = buf[i * D + a];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        ls[k][i] += lam * (sg * buf[D * D + i]);
        ms[k][i] += mu * (sg * buf[D * D + i]);
      }
    }
  }
  const double vol = stg[VOFF];
  const uint2 sb = *reinterpret_cast<const uint2 *>(stg + VOFF + 1);
  grav += __dmul_rn(vol, pr[2]);
  double lsq[NK], dq[NDOF];                         // lambda s_q of every face; unit vector of this lane's component
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    lsq[k] = ls[k][0];
#pragma unroll
    for (int i = 1; i < D; ++i) lsq[k] = (q == i) ? ls[k][i] : lsq[k];
  }
#pragma unroll
  for (int c = 0; c < NDOF; ++c) dq[c] = (q == c) ? 1.0 : 0.0;
  double *rowq = row + q * NDOF;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double Mq[NDOF];
#pragma unroll
    for (int c = 0; c < NDOF; ++c) Mq[c] = 0.0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int enc = Tab::vface_c(L, k);
      const int f = enc < 0 ? 0 : enc >> 1;
      if (enc >= 0) {
        double gv[NDOF];
        double mgs = 0.0;
#pragma unroll
        for (int i = 0; i < NDOF; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < D; ++a)
            if (Tab::ifD_c(f, j, a) != 0.0) acc += J[k][i][a] * Tab::ifD_c(f, j, a);       // g_i = sum_a Jinv[i][a] D[j][a]
          gv[i] = acc;
          mgs += acc * ms[k][i];                                                            // mu (s . g)
        }
        double gq = gv[0];
#pragma unroll
        for (int i = 1; i < NDOF; ++i) gq = (q == i) ? gv[i] : gq;
#pragma unroll
        for (int c = 0; c < NDOF; ++c) Mq[c] += lsq[k] * gv[c] + ms[k][c] * gq + dq[c] * mgs;   // row q of At.Ce.Bv in closed form
      }
    }
    const int pos = (int)(((j < 4 ? sb.x : sb.y) >> (8 * (j & 3))) & 0xffu);
    double *blk = rowq + pos * B;
#pragma unroll
    for (int c = 0; c < NDOF; ++c) blk[c] += Mq[c];
  }
}
template <int UCODE, int NDOF>
__global__ void __launch_bounds__(kBlkWarps * 32) k_elas_lanes(const __grid_constant__ MeshView m, const __grid_constant__ GeomView g,
                                                               const __grid_constant__ SysView s, const double *__restrict__ props,
                                                               const uint32_t *__restrict__ spk, const int64_t *__restrict__ soff,
                                                               int gravity, int acc_cap, int dir_mode, int single_region_i,
                                                               const uint8_t *__restrict__ dflag, const uint8_t *__restrict__ vmark,
                                                               const double *__restrict__ dval,
                                                               double *__restrict__ gen_out, double *__restrict__ elim_out) {
  constexpr int S = kBlkStages, RR = 2 * kBlkStages, R = kBlkRows, B = NDOF * NDOF;
  constexpr bool U = UCODE >= 0;
  // the generic instantiation also serves single-shape meshes without an instantiation of their own (prisms, pyramids):
  // the descriptor format follows the MESH (shape bits only when it is mixed)
  const bool MX = !U && m.uniform_code < 0;
  auto keyof = [&](uint32_t pk) { return MX ? pd_key(pk, true) : (U ? pd_key(pk, false) : m.uniform_code * 8 + pd_key(pk, false)); };
  constexpr int kRowStride = blk_row_stride(UCODE, NDOF), kStageBytes = R * kRowStride * 8;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;                                  // vertex of the tile, component / copy lane
  const bool single_region = single_region_i != 0;
  double *accw = lane_smem + (size_t)wib * blk_warp_doubles(UCODE, NDOF, acc_cap);   // image of the tile's BSR value segment
  const unsigned accw_sa = (unsigned)__cvta_generic_to_shared(accw);
  const unsigned ring_sa = accw_sa + (unsigned)acc_cap * 8u + (unsigned)(r * kRowStride) * 8u;   // this vertex's row of the stages
  const unsigned pkring_sa = accw_sa + (unsigned)acc_cap * 8u + S * kStageBytes + (unsigned)lane * 4u;
  const int ntiles = (int)((s.nV + R - 1) / R);
  const int stride = (int)gridDim.x * kBlkWarps;
  int tile = (int)blockIdx.x * kBlkWarps + wib;
  if (tile >= ntiles) return;
  const int tail_rows = (int)(s.nV - (int64_t)(ntiles - 1) * R);
  struct RowRaw { int64_t g0, g1; uint8_t mark; };
  auto load_rows = [&](int t) {
    RowRaw rr{0, 0, 0};
    const int64_t v = (int64_t)t * R + r;
    if (v < s.nV) {
      rr.g0 = s.gptr[v];
      rr.g1 = s.gptr[v + 1];
      if (dir_mode >= 0) rr.mark = vmark[v];
    } else {
      rr.g0 = rr.g1 = s.gptr[s.nV];
    }
    return rr;
  };
  int tF = tile;
  int64_t f0 = soff[tF], f1 = soff[tF + 1];
  const uint32_t *pF = spk + f0 * R + r;
  int remF = (int)(f1 - f0);
  int64_t nxt0 = 0, nxt1 = 0;
  if (tF + stride < ntiles) { nxt0 = soff[tF + stride]; nxt1 = soff[tF + stride + 1]; }
  bool okF = tF < ntiles - 1 || r < tail_rows;
  auto fetch_desc = [&](int slot) {
    const unsigned dst = pkring_sa + (unsigned)slot * 128u;
    if (remF > 0 && okF) {
      cp_async4(dst, pF);
    } else {
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(pd_elem_mask(MX) | (remF == 1 ? 1u << kPdLastBit : 0u)) : "memory");
    }
    if (remF > 0) {
      pF += R;
      if (--remF == 0) {
        tF += stride;
        if (tF < ntiles) {
          pF = spk + nxt0 * R + r;
          remF = (int)(nxt1 - nxt0);
          okF = tF < ntiles - 1 || r < tail_rows;
          if (tF + stride < ntiles) { nxt0 = soff[tF + stride]; nxt1 = soff[tF + stride + 1]; }
        }
      }
    }
  };
  // the key of a step is warp-uniform but a vertex without a pair carries only the local vertex of ITS marker: take the key
  // from the word itself (markers keep the step's local vertex and shape bits)
#define EFV_BISSUE_OP(Tab, LL) if constexpr (Tab::DIM == NDOF) elas_lane_issue<Tab, LL, U>(stg_sa, pkI, MX, q, m, g, s)
#define EFV_BCOMP_OP(Tab, LL) if constexpr (Tab::DIM == NDOF) elas_lane_compute<Tab, LL, NDOF, U>(stg, pkA, MX, q, pr, row, grav)
  auto issue = [&](int stage, int drow) {
    uint32_t pkI;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pkI) : "r"(pkring_sa + (unsigned)drow * 128u) : "memory");
    const unsigned stg_sa = ring_sa + (unsigned)stage * kStageBytes;
    EFV_LANE_DISPATCH(UCODE, keyof(pkI), EFV_BISSUE_OP);
  };
#pragma unroll 1
  for (int j = 0; j < 2 * S - 1; ++j) fetch_desc(j);
  cp_async_commit();
  RowRaw rr = load_rows(tile), rrn = rr;
  cp_async_wait<0>();
#pragma unroll 1
  for (int j = 0; j < S - 1; ++j) { issue(j, j); cp_async_commit(); }
  int slotA = 0, dA = 0;
  bool first = true;
  double grav = 0.0;
  double *row = accw;
  int64_t base0 = 0;
  int total = 0, len = 0;
  bool marked = false;
  while (true) {
    {
      const int dF = dA == 0 ? RR - 1 : dA - 1;
      int dI = dA + S - 1;
      if (dI >= RR) dI -= RR;
      const int sI = slotA == 0 ? S - 1 : slotA - 1;
      fetch_desc(dF);
      issue(sI, dI);
      cp_async_commit();
    }
    if (first) {
      if (tile + stride < ntiles) rrn = load_rows(tile + stride);
      base0 = __shfl_sync(0xffffffffu, rr.g0, 0);                         // in blocks
      total = (int)(__shfl_sync(0xffffffffu, rr.g1, 31) - base0) * B;     // scalars of the segment
      len = (int)(rr.g1 - rr.g0);
      marked = rr.mark != 0;
      for (int i = lane; i < total; i += 32) accw[i] = 0.0;
      row = accw + (int)(rr.g0 - base0) * B;
      grav = 0.0;
      first = false;
    }
    cp_async_wait<S - 1>();
    __syncwarp();                                                         // the 4 lanes of a vertex read each other's copies
    uint32_t pkA;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pkA) : "r"(pkring_sa + (unsigned)dA * 128u) : "memory");
    {
      const double *stg = accw + acc_cap + slotA * (kStageBytes / 8) + r * kRowStride;
      const double *pr = props;
      if (!single_region && !pd_none(pkA, MX)) pr += 3 * m.region[pkA & pd_elem_mask(MX)];
      EFV_LANE_DISPATCH(UCODE, keyof(pkA), EFV_BCOMP_OP);
    }
    __syncwarp();                                                         // stage may be overwritten by the next issue
    if (++slotA == S) slotA = 0;
    if (++dA == RR) dA = 0;
    if (pd_last(pkA)) {
      // ---- tile epilogue ----------------------------------------------------------------------------------------------
      const int64_t v = (int64_t)tile * R + r;
      if (v < s.nV && q < NDOF) {
        gen_out[v * NDOF + q] = (gravity && q == 1) ? grav : 0.0;         // y component only (stress_equilibrium.py:81-90)
        if (dir_mode == 1 && !marked) elim_out[v * NDOF + q] = 0.0;
      }
      __syncwarp();
      unsigned mm = __ballot_sync(0xffffffffu, marked && q == 0);
      while (mm) {                                                        // vertices with a Dirichlet DOF among their columns
        const int ln = __ffs(mm) - 1;
        mm &= mm - 1;
        const int rrel = (int)(__shfl_sync(0xffffffffu, rr.g0, ln) - base0);
        const int rlen = __shfl_sync(0xffffffffu, len, ln);
        const int64_t vr = (int64_t)tile * R + (ln >> 2);
        double el[NDOF];
#pragma unroll
        for (int i = 0; i < NDOF; ++i) el[i] = 0.0;
        for (int u = lane; u < rlen * B; u += 32) {
          const int t = u / B, ij = u - t * B, i = ij / NDOF, j = ij - i * NDOF;
          const int64_t c = s.gcol[base0 + rrel + t];
          double a = accw[rrel * B + u];
          if (dflag[vr * NDOF + i]) a = (c == vr && i == j) ? 1.0 : 0.0;
          else if (dir_mode == 1 && dflag[c * NDOF + j]) {
            const double d = a * dval[c * NDOF + j];
#pragma unroll
            for (int ii = 0; ii < NDOF; ++ii) el[ii] -= (ii == i) ? d : 0.0;
            a = 0.0;
          }
          accw[rrel * B + u] = a;
        }
        if (dir_mode == 1) {
#pragma unroll
          for (int i = 0; i < NDOF; ++i) {
            const double tot = warp_sum(el[i]);
            if (lane == 0) elim_out[vr * NDOF + i] = dflag[vr * NDOF + i] ? 0.0 : tot;
          }
        }
      }
      __syncwarp();
      double *dst = s.vals + base0 * B;
      for (int i = lane; i < total; i += 32) __stcs(dst + i, accw[i]);
      __syncwarp();
      if (tile + stride >= ntiles) break;
      tile += stride;
      rr = rrn;
      first = true;
    }
  }
  cp_async_wait<0>();
#undef EFV_BISSUE_OP
#undef EFV_BCOMP_OP
}
// vertices with a Dirichlet DOF (any component) among the columns of their block row, themselves included
__global__ void k_mark_vertices(int64_t nrows, int ndof, const int64_t *__restrict__ gptr, const int32_t *__restrict__ gcol,
                                const uint8_t *__restrict__ flag, uint8_t *__restrict__ mark) {
  const int lane = threadIdx.x & 7;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  const int64_t rounds = (nrows + groups - 1) / groups;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t v = it * groups + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3);
    int any = 0;
    if (v < nrows)
      for (int64_t t = gptr[v] + lane; t < gptr[v + 1]; t += 8) {
        const int64_t c = gcol[t];
        for (int j = 0; j < ndof; ++j) any |= flag[c * ndof + j];
      }
    any = __any_sync(gmask, any);
    if (v < nrows && lane == 0) mark[v] = (uint8_t)(any != 0);
  }
}
template <int UCODE, int NDOF>
static bool launch_elas_lanes(efv_system *s, const double *dprops, int gravity) {
  const char *mode = getenv("EFV_ASSEMBLY");
  if (mode && strcmp(mode, "rows") == 0) return false;
  constexpr int B = NDOF * NDOF;
  const int acc_cap = kBlkRows * s->max_row_len * B;
  const size_t smem = (size_t)kBlkWarps * blk_warp_doubles(UCODE, NDOF, acc_cap) * sizeof(double);
  if (smem + 1024 > (size_t)kLaneSmemBudget) return false;
  efv_mesh *m = s->mesh;
  if (!ensure_pair_schedule8(m)) return false;
  ensure_simplex_records(m);
  ensure_face_records(m, true);
  if (s->fuse_mode >= 0) {
    if (s->row_mark.n < (size_t)s->nV) s->row_mark.alloc(s->nV);
    EFV_LAUNCH(k_mark_vertices, grid_for(s->n_owned * 8, 256), 256, 0, s->n_owned, NDOF, s->gptr.p, s->gcol.p, s->dir_flag.p, s->row_mark.p);
  }
  MeshView mv = view_of(m);
  GeomView gv = geom_of(m);
  SysView sv = sys_view(s);
  EFV_CUDA(cudaFuncSetAttribute((k_elas_lanes<UCODE, NDOF>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EFV_CUDA(cudaFuncSetAttribute((k_elas_lanes<UCODE, NDOF>), cudaFuncAttributePreferredSharedMemoryCarveout, 70));
  int per_sm = 0;
  EFV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_elas_lanes<UCODE, NDOF>), kBlkWarps * 32, smem));
  per_sm = std::min<int>(per_sm, (int)(kLaneSmemBudget / (smem + 1024)));
  if (per_sm < 1) return false;
  const int64_t ntiles = div_up(s->n_owned, kBlkRows);
  int64_t blocks = std::min<int64_t>(div_up(ntiles, kBlkWarps), (int64_t)num_sms() * per_sm);
  if (const char *e = getenv("EFV_LANE_GRID")) blocks = std::min<int64_t>(blocks, std::max(1, atoi(e)));   // tests: many tiles per warp
  EFV_LAUNCH((k_elas_lanes<UCODE, NDOF>), (int)std::max<int64_t>(blocks, 1), kBlkWarps * 32, smem, mv, gv, sv, dprops, m->sched8_pk.p,
             m->sched8_off.p, gravity, acc_cap, s->fuse_mode, m->n_regions == 1 ? 1 : 0, s->dir_flag.p, s->row_mark.p, s->dir_value.p,
             s->gen.p, s->elim.p);
  return true;
}

template <int UCODE, int NDOF>
static void launch_elasticity_rows(efv_system *s, const double *dprops, int gravity) {
  if (launch_elas_lanes<UCODE, NDOF>(s, dprops, gravity)) return;       // default: lane kernel (4 lanes per vertex); EFV_ASSEMBLY=rows: below
  efv_mesh *m = s->mesh;
  size_t smem = ((size_t)kElRows * s->max_row_len * NDOF * NDOF + (size_t)kElRows * kAsmGroup * 12 * kMaxVF) * sizeof(double);
  EFV_CUDA(cudaFuncSetAttribute((k_elasticity_rows<UCODE, NDOF>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EFV_LAUNCH((k_elasticity_rows<UCODE, NDOF>), grid_for(s->n_owned * kAsmGroup, 128, 12), 128, smem, view_of(m), geom_of(m),
             sys_view(s), dprops, gravity, s->max_row_len, s->fuse_mode, s->dir_flag.p, s->dir_value.p, s->gen.p, s->elim.p);
}

void elasticity_assemble(efv_system *s, const double *props_host, int gravity) {
  efv_mesh *m = s->mesh;
  EFV_REQUIRE(s->ndof == m->dim, "elasticity needs ndof == dimension");
  geometry_build(m);
  if (!s->dir_flag.n) bc_clear(s);
  std::vector<double> pr(3 * m->n_regions);
  for (int r = 0; r < m->n_regions; ++r) {
    const double *p = props_host + 4 * r;             // ShearModulus, PoissonsRatio, Density, Gravity
    const double G = p[0], nu = p[1];
    pr[3 * r + 0] = 2 * G * nu / (1 - 2 * nu);        // lameParameter (stress_equilibrium.py:33)
    pr[3 * r + 1] = G;
    pr[3 * r + 2] = -p[2] * p[3];                     // - density * gravity (stress_equilibrium.py:89)
  }
  if (s->props_dev.n < pr.size()) s->props_dev.alloc(pr.size());
  s->props_dev.upload(pr.data(), pr.size());
  if (!s->gen.n) s->gen.alloc(s->rows);
  if (!s->ev0) { EFV_CUDA(cudaEventCreate(&s->ev0)); EFV_CUDA(cudaEventCreate(&s->ev1)); }
  EFV_CUDA(cudaEventRecord(s->ev0, stream()));
  const int uc = m->uniform_code;
  if (m->dim == 2) {
    if (uc == 0) launch_elasticity_rows<0, 2>(s, s->props_dev.p, gravity);
    else if (uc == 1) launch_elasticity_rows<1, 2>(s, s->props_dev.p, gravity);
    else launch_elasticity_rows<-1, 2>(s, s->props_dev.p, gravity);
  } else {
    if (uc == 2) launch_elasticity_rows<2, 3>(s, s->props_dev.p, gravity);
    else if (uc == 3) launch_elasticity_rows<3, 3>(s, s->props_dev.p, gravity);
    else launch_elasticity_rows<-1, 3>(s, s->props_dev.p, gravity);
  }
  EFV_CUDA(cudaEventRecord(s->ev1, stream()));
  s->assemble_timed = true;
  s->vals_version++;
  s->dirichlet_mode = s->fuse_mode;
  s->physics = 1;
}

// ---- Biot poroelasticity (fully coupled) ------------------------------------------------------------------------------
// Unknowns per vertex: u_0..u_{d-1}, p (reference numbering [u_x | u_y | (u_z) | p], geomechanics.py:255-256).
// Per inner face (area s, column vertex c, gradient g = G_f[:,c], shape value N = N_f[c]) the row of the backward vertex
// receives +coef and the row of the forward vertex -coef with  (geomechanics.py:61-112)
//   uu : lambda s_i g_j + mu (s_j g_i + delta_ij s.g)          (:89-100)
//   up : -alpha s_i N                                          (:66-77)
//   pp : -k*mu (s.g)                                           (:80-86; note k*mu, not k/mu)
//   pu : alpha/dt N s_j                                        (:103-112)
// plus S/dt vol_k on the pp diagonal (:62-63) and, on boundary vertices, alpha/dt N_outer s_outer_j in the p row of the
// outer face's own vertex (:113-115).  The pre-Dirichlet p-row/u-column block is also kept (rhs_op) because the RHS of
// a time step is that block applied to u_old (:185-196) plus S/dt vol p_old (:172-173).
// props per region: [0] lambda [1] mu [2] alpha [3] k*mu [4] S/dt [5] alpha/dt [6] rho
struct BndView {
  const int32_t *vtx_to_bv;
  const int64_t *bvo_ptr, *bvo, *outer_facet, *facet_element, *facet_ptr;
  const int32_t *facet_local, *outer_local;
  const double *facet_area;
};
__device__ __forceinline__ int tab_facetV(int code, int lf, int k) {
  switch (code) {
    case 0: return c_Triangle_facetV[lf][k];
    case 1: return c_Quadrilateral_facetV[lf][k];
    case 2: return c_Tetrahedron_facetV[lf][k];
    case 3: return c_Hexahedron_facetV[lf][k];
    case 4: return c_Prism_facetV[lf][k];
    default: return c_Pyramid_facetV[lf][k];
  }
}
__device__ __forceinline__ double tab_ofN(int code, int lf, int k, int j) {
  switch (code) {
    case 0: return c_Triangle_ofN[lf][k][j];
    case 1: return c_Quadrilateral_ofN[lf][k][j];
    case 2: return c_Tetrahedron_ofN[lf][k][j];
    case 3: return c_Hexahedron_ofN[lf][k][j];
    case 4: return c_Prism_ofN[lf][k][j];
    default: return c_Pyramid_ofN[lf][k][j];
  }
}
__device__ __forceinline__ double tab_ifN(int code, int f, int j) {
  switch (code) {
    case 0: return c_Triangle_ifN[f][j];
    case 1: return c_Quadrilateral_ifN[f][j];
    case 2: return c_Tetrahedron_ifN[f][j];
    case 3: return c_Hexahedron_ifN[f][j];
    case 4: return c_Prism_ifN[f][j];
    default: return c_Pyramid_ifN[f][j];
  }
}

constexpr int kBiotRows = 8;          // rows (8-lane groups) per 64-thread block
template <int UCODE, int D>
__global__ void __launch_bounds__(64) k_biot_rows(MeshView m, GeomView g, SysView s, BndView bd, const double *__restrict__ props,
                                                  double gx, double gy, double gz, int max_row_len, int dir_mode,
                                                  const uint8_t *__restrict__ dflag, const double *__restrict__ dval,
                                                  double *__restrict__ gen_out, double *__restrict__ acc_out,
                                                  double *__restrict__ rhs_op, double *__restrict__ elim_out) {
  constexpr int NDOF = D + 1, B = NDOF * NDOF;
  extern __shared__ double smem_dyn[];
  __shared__ SmemTables T;
  __shared__ double fbuf[kBiotRows][kAsmGroup][12 * kMaxVF];
  __shared__ double aux[kBiotRows][kAsmGroup][8];
  __shared__ unsigned long long posb[kBiotRows][kAsmGroup];
  __shared__ int meta[kBiotRows][kAsmGroup];
  load_tables(T);
  const double gvec[3] = {gx, gy, gz};
  const int grp = threadIdx.x / kAsmGroup, gl = threadIdx.x % kAsmGroup;
  double *row = smem_dyn + (size_t)grp * max_row_len * B;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) / kAsmGroup * kAsmGroup);
  for (int64_t v0 = (int64_t)blockIdx.x * kBiotRows; v0 < s.nV; v0 += (int64_t)gridDim.x * kBiotRows) {
    const int64_t v = v0 + grp;
    if (v >= s.nV) continue;
    const int64_t r0 = s.gptr[v];
    const int len = (int)(s.gptr[v + 1] - r0);
    for (int t = gl; t < len * B; t += kAsmGroup) row[t] = 0.0;
    double gen[NDOF], accv = 0.0;
#pragma unroll
    for (int i = 0; i < NDOF; ++i) gen[i] = 0.0;
    const int64_t p0 = m.inc_ptr[v], p1 = m.inc_ptr[v + 1];
    for (int64_t pc = p0; pc < p1; pc += kAsmGroup) {
      const int64_t p = pc + gl;
      if (p < p1) {
        const uint32_t pk = m.inc[p];
        const int64_t e = inc_elem(pk);
        const int l = inc_local(pk);
        const int code = (UCODE >= 0) ? UCODE : m.code(e);
        const int mm = shape_m(code), nvf = shape_nvf(code);
        const int64_t gi = (UCODE >= 0) ? e : m.group_index(e);
        const int64_t n = g.n[code];
        const double *pr = props + 7 * m.region[e];
        const double *Ji = g.Jinv[code], *Sv = g.S[code];
        int mt = 1 | (code << 1);
        for (int q = 0; q < nvf; ++q) {
          const int enc = tab_vface(T, code, l, q);
          const int f = enc < 0 ? 0 : enc >> 1;
          mt |= mt_pack_face(enc, q);
          double *fb = &fbuf[grp][gl][q * 12];
          for (int a = 0; a < D * D; ++a) fb[a] = enc < 0 ? 0.0 : Ji[(int64_t)(f * D * D + a) * n + gi];
          for (int a = 0; a < D; ++a) fb[9 + a] = enc < 0 ? 0.0 : Sv[(int64_t)(f * 3 + a) * n + gi];
        }
        meta[grp][gl] = mt;
        const uint8_t *sp = s.slotpos + ((UCODE >= 0) ? e * (int64_t)(mm * mm) : m.slot_begin(e)) + l * mm;
        unsigned long long pb;
        if (UCODE >= 0 && mm == 8) pb = *reinterpret_cast<const unsigned long long *>(sp);
        else if (UCODE >= 0 && mm == 4) pb = *reinterpret_cast<const unsigned int *>(sp);
        else {
          pb = 0;
          for (int j = 0; j < mm; ++j) pb |= (unsigned long long)sp[j] << (8 * j);
        }
        posb[grp][gl] = pb;
#pragma unroll
        for (int k = 0; k < 7; ++k) aux[grp][gl][k] = pr[k];
        aux[grp][gl][7] = g.vol[code][(int64_t)l * n + gi];
      }
      __syncwarp(gmask);
      const int cnt = (int)((p1 - pc < kAsmGroup) ? (p1 - pc) : kAsmGroup);
      for (int q = 0; q < cnt; ++q) {
        const int mt = meta[grp][q];
        const int code = (UCODE >= 0) ? UCODE : mt_code(mt);
        const int mm = shape_m(code), nvf = shape_nvf(code);
        const double lam = aux[grp][q][0], mu = aux[grp][q][1], alpha = aux[grp][q][2], kmu = aux[grp][q][3];
        const double sdt = aux[grp][q][4], adt = aux[grp][q][5], rho = aux[grp][q][6], vol = aux[grp][q][7];
        if (gl < mm) {
          double M[NDOF][NDOF];
#pragma unroll
          for (int i = 0; i < NDOF; ++i)
#pragma unroll
            for (int j = 0; j < NDOF; ++j) M[i][j] = 0.0;
          for (int k = 0; k < nvf; ++k) {
            const int f = mt_face(mt, k);
            const double sgn = mt_fwd(mt, k) ? -1.0 : 1.0;
            const double *fb = &fbuf[grp][q][k * 12];
            const double N = tab_ifN(code, f, gl);
            double gv[D], sv[D], dl[D];
#pragma unroll
            for (int a = 0; a < D; ++a) { dl[a] = tab_ifD(T, code, f, gl, a); sv[a] = fb[9 + a]; }
            double gs = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              double acc = 0.0;
#pragma unroll
              for (int a = 0; a < D; ++a) acc += fb[i * D + a] * dl[a];
              gv[i] = acc;
              gs += acc * sv[i];
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
              for (int j = 0; j < D; ++j) M[i][j] += sgn * (lam * sv[i] * gv[j] + mu * (sv[j] * gv[i] + ((i == j) ? gs : 0.0)));
              M[i][D] += sgn * (-alpha * sv[i] * N);
              M[D][i] += sgn * (adt * N * sv[i]);
            }
            M[D][D] += sgn * (-kmu * gs);
          }
          const int pos = (int)((posb[grp][q] >> (8 * gl)) & 0xff);
          double *blk = row + pos * B;
#pragma unroll
          for (int i = 0; i < NDOF; ++i)
#pragma unroll
            for (int j = 0; j < NDOF; ++j) blk[i * NDOF + j] += M[i][j];
        }
        if (gl == 0) {
          accv += vol * sdt;                                   // S/dt vol (geomechanics.py:62-63)
#pragma unroll
          for (int i = 0; i < D; ++i) gen[i] += vol * (-rho * gvec[i]);                 // (:166-169)
          for (int k = 0; k < nvf; ++k) {                      // gravity flux (:176-183)
            const double sgn = mt_fwd(mt, k) ? -1.0 : 1.0;
            const double *fb = &fbuf[grp][q][k * 12];
            double sg = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) sg += fb[9 + a] * gvec[a];
            gen[D] += sgn * (-kmu * rho * sg);
          }
        }
      }
      __syncwarp(gmask);
    }
    // outer faces of this vertex: alpha/dt N_outer[c] s_outer_j on (p row of v) x (u_j of every element vertex c)
    const int bvi = bd.vtx_to_bv[v];
    if (bvi >= 0) {
      for (int64_t po = bd.bvo_ptr[bvi]; po < bd.bvo_ptr[bvi + 1]; ++po) {
        const int64_t o = bd.bvo[po];
        const int64_t f = bd.outer_facet[o];
        const int64_t e = bd.facet_element[f];
        const int lf = bd.facet_local[f], ol = bd.outer_local[o];
        const int mf = (int)(bd.facet_ptr[f + 1] - bd.facet_ptr[f]);
        const int code = (UCODE >= 0) ? UCODE : m.code(e);
        const int mm = shape_m(code);
        const int l = tab_facetV(code, lf, ol);
        const double adt = props[7 * m.region[e] + 5];
        if (gl < mm) {
          const double N = tab_ofN(code, lf, ol, gl);
          const int pos = s.slotpos[((UCODE >= 0) ? e * (int64_t)(mm * mm) : m.slot_begin(e)) + l * mm + gl];
          double *blk = row + pos * B;
#pragma unroll
          for (int j = 0; j < D; ++j) blk[D * NDOF + j] += adt * N * (bd.facet_area[f * 3 + j] / mf);    // OuterFace.py:18-19
        }
        __syncwarp(gmask);
      }
    }
    if (gl == 0) {
      row[(s.diag_slot[v] - r0) * B + D * NDOF + D] += accv;
      acc_out[v] = accv;
#pragma unroll
      for (int i = 0; i < NDOF; ++i) gen_out[v * NDOF + i] = gen[i];
    }
    __syncwarp(gmask);
    bool rowd[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) rowd[i] = (dir_mode >= 0) ? dflag[v * NDOF + i] : false;
    double el[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) el[i] = 0.0;
    for (int t = gl; t < len; t += kAsmGroup) {
      const int64_t c = s.gcol[r0 + t];
      double *blk = row + t * B;
      double *out = s.vals + (r0 + t) * B;
#pragma unroll
      for (int j = 0; j < D; ++j) rhs_op[(r0 + t) * D + j] = blk[D * NDOF + j];      // pre-Dirichlet p-row / u-col block
#pragma unroll
      for (int i = 0; i < NDOF; ++i)
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
          double a = blk[i * NDOF + j];
          if (rowd[i]) a = (c == v && i == j) ? 1.0 : 0.0;                          // geomechanics.py:117-137
          else if (dir_mode == 1 && dflag[c * NDOF + j]) { el[i] -= a * dval[c * NDOF + j]; a = 0.0; }
          out[i * NDOF + j] = a;
        }
    }
    if (dir_mode == 1) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) {
        double tot = el[i];
#pragma unroll
        for (int o = kAsmGroup / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(gmask, tot, o);
        if (gl == 0) elim_out[v * NDOF + i] = rowd[i] ? 0.0 : tot;
      }
    }
    __syncwarp(gmask);
  }
}

template <int UCODE, int D>
static void launch_biot_rows(efv_system *s, const double *dprops, const double *gv) {
  efv_mesh *m = s->mesh;
  constexpr int NDOF = D + 1;
  size_t smem = (size_t)kBiotRows * s->max_row_len * NDOF * NDOF * sizeof(double);
  EFV_REQUIRE(smem <= 160 * 1024, "graph rows too long for the Biot assembly kernel");
  const Boundaries &B = m->bnd;
  EFV_REQUIRE(B.vtx_to_bv.n, "set the mesh boundaries before assembling the Biot system");
  BndView bd{B.vtx_to_bv.p, B.bvo_ptr.p, B.bvo.p, B.outer_facet.p, B.facet_element.p, B.facet_ptr.p, B.facet_local.p, B.outer_local.p, B.facet_area.p};
  EFV_CUDA(cudaFuncSetAttribute((k_biot_rows<UCODE, D>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EFV_LAUNCH((k_biot_rows<UCODE, D>), grid_for(s->n_owned * kAsmGroup, 64, 16), 64, smem, view_of(m), geom_of(m), sys_view(s), bd, dprops,
             gv[0], gv[1], gv[2], s->max_row_len, s->fuse_mode, s->dir_flag.p, s->dir_value.p, s->gen.p, s->acc.p, s->rhs_op.p, s->elim.p);
}

void biot_assemble(efv_system *s, const double *props_host, double dt, const double *gvec) {
  efv_mesh *m = s->mesh;
  EFV_REQUIRE(s->ndof == m->dim + 1, "Biot needs ndof == dimension + 1");
  geometry_build(m);
  if (!s->dir_flag.n) bc_clear(s);
  std::vector<double> pr(7 * m->n_regions);
  for (int r = 0; r < m->n_regions; ++r) {
    const double *p = props_host + 9 * r;      // nu, G, cs, phi, k, cf, mu, rhos, rhof  (geomechanics.py:41-49)
    const double nu = p[0], G = p[1], cs = p[2], phi = p[3], k = p[4], cf = p[5], mu = p[6], rhos = p[7], rhof = p[8];
    const double lame = 2 * G * nu / (1 - 2 * nu);
    const double rho = phi * rhof + (1 - phi) * rhos;
    const double K = 2 * G * (1 + nu) / 3 * (1 - 2 * nu);     // precedence quirk of geomechanics.py:55, reproduced
    const double cb = 1 / K;
    const double alpha = 1 - cs / cb;
    const double S = phi * cf + (alpha - phi) * cs;
    pr[7 * r + 0] = lame; pr[7 * r + 1] = G; pr[7 * r + 2] = alpha; pr[7 * r + 3] = k * mu;
    pr[7 * r + 4] = S * (1 / dt); pr[7 * r + 5] = alpha * (1 / dt); pr[7 * r + 6] = rho;
  }
  if (s->props_dev.n < pr.size()) s->props_dev.alloc(pr.size());
  s->props_dev.upload(pr.data(), pr.size());
  if (!s->gen.n) { s->gen.alloc(s->rows); s->acc.alloc(s->nV); s->rhs_op.alloc((size_t)s->nnzg * m->dim); }
  double gv[3] = {gvec ? gvec[0] : 0.0, gvec ? gvec[1] : 0.0, gvec ? gvec[2] : 0.0};
  if (!s->ev0) { EFV_CUDA(cudaEventCreate(&s->ev0)); EFV_CUDA(cudaEventCreate(&s->ev1)); }
  EFV_CUDA(cudaEventRecord(s->ev0, stream()));
  const int uc = m->uniform_code;
  if (m->dim == 2) {
    if (uc == 0) launch_biot_rows<0, 2>(s, s->props_dev.p, gv);
    else if (uc == 1) launch_biot_rows<1, 2>(s, s->props_dev.p, gv);
    else launch_biot_rows<-1, 2>(s, s->props_dev.p, gv);
  } else {
    if (uc == 2) launch_biot_rows<2, 3>(s, s->props_dev.p, gv);
    else if (uc == 3) launch_biot_rows<3, 3>(s, s->props_dev.p, gv);
    else launch_biot_rows<-1, 3>(s, s->props_dev.p, gv);
  }
  EFV_CUDA(cudaEventRecord(s->ev1, stream()));
  s->assemble_timed = true;
  s->vals_version++;
  s->dirichlet_mode = s->fuse_mode;
  s->physics = 2;
}

// ---- boundary conditions -----------------------------------------------------------------------------------------
// field-major DOF index (C-ABI) -> internal vertex-major index
__device__ __forceinline__ int64_t internal_index(int64_t row, int64_t nV, int ndof) {
  int64_t field = row / nV, v = row % nV;
  return v * ndof + field;
}

__global__ void k_dir_last(const int64_t *__restrict__ rows, int64_t n, int base, int64_t nV, int ndof,
                           int32_t *__restrict__ last) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    atomicMax(&last[internal_index(rows[k], nV, ndof)], base + (int)k);
}
__global__ void k_dir_set(const int64_t *__restrict__ rows, const double *__restrict__ vals, int64_t n, int base,
                          int64_t nV, int ndof, const int32_t *__restrict__ last, uint8_t *__restrict__ flag,
                          double *__restrict__ value) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx = internal_index(rows[k], nV, ndof);
    if (last[idx] == base + (int)k) { flag[idx] = 1; value[idx] = vals[k]; }   // the LAST visit wins (heat_transfer.py:122-126)
  }
}



void bc_clear(efv_system *s) {
  if (!s->dir_flag.n) {
    s->dir_flag.alloc(s->rows);
    s->dir_value.alloc(s->rows);
    s->dir_last.alloc(s->rows);
    s->static_rhs.alloc(s->rows);
    s->elim.alloc(s->rows);
  }
  s->dir_flag.zero();
  s->dir_value.zero();
  if (s->ndof == 1) {
    if (!s->row_mark.n) s->row_mark.alloc(s->rows);
    s->row_mark.zero();
    s->marks_valid = true;
  }
  EFV_CUDA(cudaMemsetAsync(s->dir_last.p, 0xff, s->dir_last.bytes(), stream()));   // -1
  s->static_rhs.zero();
  s->elim.zero();
  s->has_dirichlet = false;
  s->dir_counter = 0;
}

void bc_dirichlet(efv_system *s, const int64_t *rows, const double *values, int64_t n) {
  if (!s->dir_flag.n) bc_clear(s);
  if (n == 0) return;
  EFV_REQUIRE(s->dir_counter + n < (int64_t(1) << 31), "too many Dirichlet entries");
  DBuf<int64_t> dr;
  DBuf<double> dv;
  dr.upload(rows, n);
  dv.upload(values, n);
  EFV_LAUNCH(k_dir_last, grid_for(n, 256), 256, 0, dr.p, n, s->dir_counter, s->nV, s->ndof, s->dir_last.p);
  EFV_LAUNCH(k_dir_set, grid_for(n, 256), 256, 0, dr.p, dv.p, n, s->dir_counter, s->nV, s->ndof, s->dir_last.p,
             s->dir_flag.p, s->dir_value.p);
  if (s->ndof == 1 && s->marks_valid) {
    if (s->n_owned == s->nV && s->gptr.n)
      EFV_LAUNCH(k_mark_scatter, grid_for(n * 8, 256), 256, 0, dr.p, n, s->gptr.p, s->gcol.p, s->row_mark.p);
    else
      s->marks_valid = false;            // ghost Dirichlet rows have no graph row here: recomputed before the assembly
  }
  EFV_CUDA(cudaStreamSynchronize(stream()));
  s->dir_counter += (int)n;
  s->has_dirichlet = true;
}

// Neumann: one thread per boundary vertex; its outer faces are visited in ascending outer-face id, i.e. in the
// reference's facet order (heat_transfer.py:115-120) -> deterministic
__global__ void k_neumann(int64_t nbv, const int32_t *__restrict__ bv, const int64_t *__restrict__ bvo_ptr,
                          const int64_t *__restrict__ bvo, const double *__restrict__ norm, int64_t o0, int64_t o1,
                          int ndof, int dof, double coef, const double *__restrict__ per_face, double *__restrict__ rhs,
                          const uint8_t *__restrict__ dflag, double *__restrict__ dvalue) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nbv; k += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    bool any = false;
    for (int64_t p = bvo_ptr[k]; p < bvo_ptr[k + 1]; ++p) {
      int64_t o = bvo[p];
      if (o >= o0 && o < o1) { acc += (per_face ? coef * per_face[o - o0] : coef) * norm[o]; any = true; }
    }
    if (any) {
      const int64_t idx = (int64_t)bv[k] * ndof + dof;
      rhs[idx] += acc;
      // the reference applies boundary conditions in visiting order: a Neumann boundary visited AFTER a Dirichlet one
      // that shares the vertex adds its flux on top of the prescribed value (stress_equilibrium.py:119-163 loops
      // boundary by boundary); a later Dirichlet visit overwrites it again (k_dir_set)
      if (dflag[idx]) dvalue[idx] += acc;
    }
  }
}

// values == nullptr: the constant `value` on every outer face; otherwise one value per outer face of the boundary, in the
// reference's visiting order (facet by facet, outer face by outer face: Boundary.py:62-75)
void bc_neumann(efv_system *s, int boundary, int dof, double value, double sign, const double *values) {
  efv_mesh *m = s->mesh;
  Boundaries &B = m->bnd;
  EFV_REQUIRE(boundary >= 0 && boundary < B.nB, "bad boundary index");
  EFV_REQUIRE(dof >= 0 && dof < s->ndof, "bad dof");
  if (!s->dir_flag.n) bc_clear(s);
  if ((!values && value == 0.0) || B.nbv == 0) return;
  const int64_t o0 = B.outer_ptr_host[boundary], o1 = B.outer_ptr_host[boundary + 1];
  DBuf<double> dv;
  if (values) dv.upload(values, (size_t)(o1 - o0));
  EFV_LAUNCH(k_neumann, grid_for(B.nbv, 128), 128, 0, B.nbv, B.bv.p, B.bvo_ptr.p, B.bvo.p, B.outer_area_norm.p, o0, o1, s->ndof, dof,
             values ? sign : sign * value, values ? dv.p : (const double *)nullptr, s->static_rhs.p, s->dir_flag.p, s->dir_value.p);
  if (values) EFV_CUDA(cudaStreamSynchronize(stream()));      // dv is released on return
}

// Dirichlet on the matrix: GROUP lanes per row-vertex.
//   rows mode      : Dirichlet row <- unit row                                   (heat_transfer.py:70-76)
//   symmetric mode : additionally a_ij g_j moves to elim_i for Dirichlet columns
template <int NDOF>
__global__ void __launch_bounds__(256) k_apply_dirichlet(SysView s, const uint8_t *__restrict__ flag,
                                                         const double *__restrict__ gval, int symmetric,
                                                         double *__restrict__ elim) {
  constexpr int GROUP = 8;
  int gl = threadIdx.x % GROUP;
  const int64_t rows_per_block = blockDim.x / GROUP;
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < s.nV; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t v = base + threadIdx.x / GROUP;      // block-uniform loop: all lanes take part in the shuffles
    const bool valid = v < s.nV;
    int64_t r0 = valid ? s.gptr[v] : 0;
    int len = valid ? (int)(s.gptr[v + 1] - r0) : 0;
    bool rowd[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) rowd[i] = valid ? flag[v * NDOF + i] : false;
    double el[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) el[i] = 0.0;
    for (int t = gl; t < len; t += GROUP) {
      int64_t c = s.gcol[r0 + t];
      double *blk = s.vals + (r0 + t) * (NDOF * NDOF);
      bool cold[NDOF];
      bool anyc = false;
#pragma unroll
      for (int j = 0; j < NDOF; ++j) { cold[j] = symmetric ? flag[c * NDOF + j] : false; anyc |= cold[j]; }
#pragma unroll
      for (int i = 0; i < NDOF; ++i) {
        if (rowd[i]) {
#pragma unroll
          for (int j = 0; j < NDOF; ++j) blk[i * NDOF + j] = (c == v && i == j) ? 1.0 : 0.0;
        } else if (anyc) {
#pragma unroll
          for (int j = 0; j < NDOF; ++j)
            if (cold[j]) { el[i] -= blk[i * NDOF + j] * gval[c * NDOF + j]; blk[i * NDOF + j] = 0.0; }
        }
      }
    }
    if (symmetric) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) {
        double tot = group_sum<GROUP>(el[i]);
        if (valid && gl == 0) elim[v * NDOF + i] = rowd[i] ? 0.0 : tot;
      }
    }
  }
}

void apply_dirichlet(efv_system *s, int mode) {
  EFV_REQUIRE(mode == 0 || mode == 1, "bad Dirichlet mode");
  EFV_REQUIRE(s->vals.n, "assemble before applying Dirichlet conditions");
  if (s->dirichlet_mode == mode) return;      // already applied inside the assembly kernel (efv_assemble_dirichlet_mode)
  EFV_REQUIRE(s->dirichlet_mode < 0, "Dirichlet conditions were already applied to these values; re-assemble first");
  if (!s->dir_flag.n) bc_clear(s);
  SysView sv = sys_view(s);
  int grid = grid_for(s->nV * 8, 256, 8);
  switch (s->ndof) {
    case 1: EFV_LAUNCH(k_apply_dirichlet<1>, grid, 256, 0, sv, s->dir_flag.p, s->dir_value.p, mode, s->elim.p); break;
    case 2: EFV_LAUNCH(k_apply_dirichlet<2>, grid, 256, 0, sv, s->dir_flag.p, s->dir_value.p, mode, s->elim.p); break;
    case 3: EFV_LAUNCH(k_apply_dirichlet<3>, grid, 256, 0, sv, s->dir_flag.p, s->dir_value.p, mode, s->elim.p); break;
    case 4: EFV_LAUNCH(k_apply_dirichlet<4>, grid, 256, 0, sv, s->dir_flag.p, s->dir_value.p, mode, s->elim.p); break;
    default: throw Error("unsupported ndof");
  }
  s->dirichlet_mode = mode;
  s->vals_version++;
}

// ---- right-hand side -----------------------------------------------------------------------------------------------
// heat: b = gen + acc * T_old + neumann (+ eliminated Dirichlet columns); Dirichlet rows <- g  (heat_transfer.py:89-132)
__global__ void k_rhs_heat(int64_t n, const double *__restrict__ gen, const double *__restrict__ acc,
                           const double *__restrict__ xold, const double *__restrict__ stat,
                           const double *__restrict__ elim, const uint8_t *__restrict__ flag,
                           const double *__restrict__ gval, int transient, int symmetric, double *__restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double r = gen[i];
    if (transient) r += acc[i] * xold[i];
    r += stat[i];
    if (symmetric) r += elim[i];
    if (flag[i]) r = gval[i];
    b[i] = r;
  }
}
__global__ void k_rhs_static(int64_t n, const double *__restrict__ base, const double *__restrict__ stat,
                             const double *__restrict__ elim, const uint8_t *__restrict__ flag,
                             const double *__restrict__ gval, int symmetric, double *__restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double r = (base ? base[i] : 0.0) + stat[i];
    if (symmetric) r += elim[i];
    if (flag[i]) r = gval[i];
    b[i] = r;
  }
}

void ensure_vectors(efv_system *s) {
  if (!s->x.n) { s->x.alloc(s->rows); s->x.zero(); }
  if (!s->b.n) { s->b.alloc(s->rows); s->b.zero(); }
  if (!s->xold.n) { s->xold.alloc(s->rows); s->xold.zero(); }
}

// Biot: u rows = body force + neumann; p row = gravity flux + S/dt vol p_old + (alpha/dt N s) . u_old + neumann
// (geomechanics.py:166-221); Dirichlet rows <- g (:224-239).  One 8-lane group per vertex for the sparse product.
template <int D>
__global__ void __launch_bounds__(256) k_rhs_biot(SysView s, const double *__restrict__ gen, const double *__restrict__ acc,
                                                  const double *__restrict__ rhs_op, const double *__restrict__ xold,
                                                  const double *__restrict__ stat, const double *__restrict__ elim,
                                                  const uint8_t *__restrict__ flag, const double *__restrict__ gval,
                                                  int symmetric, double *__restrict__ b) {
  constexpr int NDOF = D + 1, GROUP = 8;
  const int gl = threadIdx.x % GROUP;
  const int64_t rows_per_block = blockDim.x / GROUP;
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < s.nV; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t v = base + threadIdx.x / GROUP;
    const bool valid = v < s.nV;
    const int64_t r0 = valid ? s.gptr[v] : 0, r1 = valid ? s.gptr[v + 1] : 0;
    double sum = 0.0;
    for (int64_t t = r0 + gl; t < r1; t += GROUP) {
      const int64_t c = s.gcol[t];
#pragma unroll
      for (int j = 0; j < D; ++j) sum += rhs_op[t * D + j] * xold[c * NDOF + j];
    }
    sum = group_sum<GROUP>(sum);
    if (valid && gl == 0) {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) {
        const int64_t k = v * NDOF + i;
        double r = gen[k] + stat[k];
        if (i == D) r += acc[v] * xold[k] + sum;
        if (symmetric) r += elim[k];
        if (flag[k]) r = gval[k];
        b[k] = r;
      }
    }
  }
}

void build_rhs(efv_system *s, int transient) {
  ensure_vectors(s);
  if (!s->dir_flag.n) bc_clear(s);
  int sym = s->dirichlet_mode == 1;
  if (s->physics == 2) {
    EFV_REQUIRE(s->rhs_op.n, "efv_biot_assemble first");
    SysView sv = sys_view(s);
    int grid = grid_for(s->nV * 8, 256, 8);
    if (s->mesh->dim == 2)
      EFV_LAUNCH(k_rhs_biot<2>, grid, 256, 0, sv, s->gen.p, s->acc.p, s->rhs_op.p, s->xold.p, s->static_rhs.p, s->elim.p,
                 s->dir_flag.p, s->dir_value.p, sym, s->b.p);
    else
      EFV_LAUNCH(k_rhs_biot<3>, grid, 256, 0, sv, s->gen.p, s->acc.p, s->rhs_op.p, s->xold.p, s->static_rhs.p, s->elim.p,
                 s->dir_flag.p, s->dir_value.p, sym, s->b.p);
    return;
  }
  if (s->physics == 0) {
    EFV_REQUIRE(s->gen.n, "efv_heat_assemble first");
    EFV_LAUNCH(k_rhs_heat, grid_for(s->n_owned, 256), 256, 0, s->n_owned, s->gen.p, s->acc.p, s->xold.p, s->static_rhs.p,
               s->elim.p, s->dir_flag.p, s->dir_value.p, transient, sym, s->b.p);
  } else {
    EFV_LAUNCH(k_rhs_static, grid_for(s->n_owned * s->ndof, 256), 256, 0, s->n_owned * s->ndof, s->gen.n ? s->gen.p : nullptr, s->static_rhs.p,
               s->elim.p, s->dir_flag.p, s->dir_value.p, sym, s->b.p);
  }
}

// ---- fixed-stress split: right-hand sides from the blocks of the coupled operator (apps/fixed_stress.py:119-275) ----
// `biot` holds the UNMODIFIED coupled operator (assembled with Dirichlet mode -1), the current iterate (u_k, p_k) in X and
// the previous time step in XOLD; `mass` (1 DOF) and `geo` (d DOF) are the two systems actually solved.
//   mass RHS = S/dt vol p_prev + cb alpha^2/dt vol p_k + A_pu (u_prev - u_k) + Neumann(p);   Dirichlet rows <- g   (:142-199)
//   geo  RHS = - A_up p_next + Neumann(u);                                                   Dirichlet rows <- g   (:225-275)
template <int D>
__global__ void __launch_bounds__(256) k_fs_mass_rhs(SysView sb, const double *__restrict__ xk, const double *__restrict__ xprev,
                                                     const double *__restrict__ accS, const double *__restrict__ accM,
                                                     const double *__restrict__ stat, const double *__restrict__ elim,
                                                     const uint8_t *__restrict__ flag, const double *__restrict__ gval,
                                                     int symmetric, double *__restrict__ b) {
  constexpr int NDOF = D + 1, B = NDOF * NDOF, GROUP = 8;
  const int gl = threadIdx.x % GROUP;
  const int64_t rows_per_block = blockDim.x / GROUP;
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < sb.nV; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t v = base + threadIdx.x / GROUP;
    const bool valid = v < sb.nV;
    const int64_t r0 = valid ? sb.gptr[v] : 0, r1 = valid ? sb.gptr[v + 1] : 0;
    double sum = 0.0;
    for (int64_t t = r0 + gl; t < r1; t += GROUP) {
      const int64_t c = sb.gcol[t];
#pragma unroll
      for (int j = 0; j < D; ++j) sum += sb.vals[t * B + D * NDOF + j] * (xprev[c * NDOF + j] - xk[c * NDOF + j]);
    }
    sum = group_sum<GROUP>(sum);
    if (valid && gl == 0) {
      double r = accS[v] * xprev[v * NDOF + D] + (accM[v] - accS[v]) * xk[v * NDOF + D] + sum + stat[v];
      if (symmetric) r += elim[v];
      if (flag[v]) r = gval[v];
      b[v] = r;
    }
  }
}
template <int D>
__global__ void __launch_bounds__(256) k_fs_geo_rhs(SysView sb, const double *__restrict__ pnext, const double *__restrict__ stat,
                                                    const double *__restrict__ elim, const uint8_t *__restrict__ flag,
                                                    const double *__restrict__ gval, int symmetric, double *__restrict__ b) {
  constexpr int NDOF = D + 1, B = NDOF * NDOF, GROUP = 8;
  const int gl = threadIdx.x % GROUP;
  const int64_t rows_per_block = blockDim.x / GROUP;
  for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < sb.nV; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t v = base + threadIdx.x / GROUP;
    const bool valid = v < sb.nV;
    const int64_t r0 = valid ? sb.gptr[v] : 0, r1 = valid ? sb.gptr[v + 1] : 0;
    double sum[D];
#pragma unroll
    for (int i = 0; i < D; ++i) sum[i] = 0.0;
    for (int64_t t = r0 + gl; t < r1; t += GROUP) {
      const double pc = pnext[sb.gcol[t]];
#pragma unroll
      for (int i = 0; i < D; ++i) sum[i] += sb.vals[t * B + i * NDOF + D] * pc;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) sum[i] = group_sum<GROUP>(sum[i]);
    if (valid && gl == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int64_t k = v * D + i;
        double r = -sum[i] + stat[k];
        if (symmetric) r += elim[k];
        if (flag[k]) r = gval[k];
        b[k] = r;
      }
    }
  }
}
// difference = max(|p_next - p_k|, |u_next - u_k|) (fixed_stress.py:300), then (u_k, p_k) <- (u_next, p_next)
template <int D>
__global__ void __launch_bounds__(256) k_fs_update(int64_t nV, const double *__restrict__ unext, const double *__restrict__ pnext,
                                                   double *__restrict__ xk, double *__restrict__ partial) {
  constexpr int NDOF = D + 1;
  __shared__ double red[32];
  double mx = 0.0;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double un = unext[v * D + i];
      mx = fmax(mx, fabs(un - xk[v * NDOF + i]));
      xk[v * NDOF + i] = un;
    }
    const double pn = pnext[v];
    mx = fmax(mx, fabs(pn - xk[v * NDOF + D]));
    xk[v * NDOF + D] = pn;
  }
  mx = block_max(mx, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = mx;
}
__global__ void k_fs_max_final(int G, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) m = fmax(m, partial[i]);
  m = block_max(m, red);
  if (threadIdx.x == 0) *out = m;
}

static void fs_check(efv_system *biot, efv_system *mass, efv_system *geo) {
  EFV_REQUIRE(biot && biot->mesh && biot->physics == 2 && biot->dirichlet_mode < 0,
              "fixed-stress: the coupled system must be assembled by efv_biot_assemble with Dirichlet mode -1");
  const int d = biot->mesh->dim;
  if (mass) EFV_REQUIRE(mass->mesh == biot->mesh && mass->ndof == 1 && mass->acc.n, "fixed-stress: mass system must be a 1-DOF heat-type system on the same mesh");
  if (geo) EFV_REQUIRE(geo->mesh == biot->mesh && geo->ndof == d, "fixed-stress: geo system must be a d-DOF elasticity system on the same mesh");
}
void fs_mass_rhs(efv_system *biot, efv_system *mass) {
  fs_check(biot, mass, nullptr);
  ensure_vectors(biot); ensure_vectors(mass);
  SysView sb = sys_view(biot);
  int grid = grid_for(biot->n_owned * 8, 256, 8);
  int sym = mass->dirichlet_mode == 1;
  if (biot->mesh->dim == 2)
    EFV_LAUNCH(k_fs_mass_rhs<2>, grid, 256, 0, sb, biot->x.p, biot->xold.p, biot->acc.p, mass->acc.p, mass->static_rhs.p, mass->elim.p,
               mass->dir_flag.p, mass->dir_value.p, sym, mass->b.p);
  else
    EFV_LAUNCH(k_fs_mass_rhs<3>, grid, 256, 0, sb, biot->x.p, biot->xold.p, biot->acc.p, mass->acc.p, mass->static_rhs.p, mass->elim.p,
               mass->dir_flag.p, mass->dir_value.p, sym, mass->b.p);
}
void fs_geo_rhs(efv_system *biot, efv_system *mass, efv_system *geo) {
  fs_check(biot, mass, geo);
  ensure_vectors(geo);
  SysView sb = sys_view(biot);
  int grid = grid_for(biot->n_owned * 8, 256, 8);
  int sym = geo->dirichlet_mode == 1;
  if (biot->mesh->dim == 2)
    EFV_LAUNCH(k_fs_geo_rhs<2>, grid, 256, 0, sb, mass->x.p, geo->static_rhs.p, geo->elim.p, geo->dir_flag.p, geo->dir_value.p, sym, geo->b.p);
  else
    EFV_LAUNCH(k_fs_geo_rhs<3>, grid, 256, 0, sb, mass->x.p, geo->static_rhs.p, geo->elim.p, geo->dir_flag.p, geo->dir_value.p, sym, geo->b.p);
}
double fs_update(efv_system *biot, efv_system *mass, efv_system *geo) {
  fs_check(biot, mass, geo);
  const int G = num_sms() * 4;
  DBuf<double> partial(G + 1);
  if (biot->mesh->dim == 2) EFV_LAUNCH(k_fs_update<2>, G, 256, 0, biot->n_owned, geo->x.p, mass->x.p, biot->x.p, partial.p);
  else EFV_LAUNCH(k_fs_update<3>, G, 256, 0, biot->n_owned, geo->x.p, mass->x.p, biot->x.p, partial.p);
  EFV_LAUNCH(k_fs_max_final, 1, 256, 0, G, partial.p, partial.p + G);
  double out = 0.0;
  EFV_CUDA(cudaMemcpyAsync(&out, partial.p + G, sizeof(double), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  return out;
}

// ---- exports / vector transfer --------------------------------------------------------------------------------------
__global__ void k_export_values(int64_t nV, int ndof, int64_t nnzg, const int64_t *__restrict__ gptr,
                                const double *__restrict__ vals, double *__restrict__ out) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
    int64_t r0 = gptr[v];
    int len = (int)(gptr[v + 1] - r0);
    for (int i = 0; i < ndof; ++i) {
      int64_t start = (int64_t)ndof * ((int64_t)i * nnzg + r0);
      for (int j = 0; j < ndof; ++j)
        for (int k = 0; k < len; ++k) out[start + (int64_t)j * len + k] = vals[((r0 + k) * ndof + i) * ndof + j];
    }
  }
}
void export_values(const efv_system *s, double *data) {
  EFV_REQUIRE(s->vals.n, "no values assembled");
  int64_t nnz = s->nnzg * s->ndof * s->ndof;
  if (s->ndof == 1) { s->vals.download(data, nnz); return; }
  DBuf<double> d(nnz);
  EFV_LAUNCH(k_export_values, grid_for(s->nV, 128), 128, 0, s->nV, s->ndof, s->nnzg, s->gptr.p, s->vals.p, d.p);
  d.download(data, nnz);
}

__global__ void k_field_to_vertex_major(int64_t nV, int ndof, const double *__restrict__ in, double *__restrict__ out, int reverse) {
  int64_t n = nV * ndof;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i / ndof, f = i % ndof;      // i is the vertex-major index
    if (reverse) out[f * nV + v] = in[i]; else out[i] = in[f * nV + v];
  }
}
void vec_upload(efv_system *s, DBuf<double> &dst, const double *host) {
  if (!dst.n) dst.alloc(s->rows);
  count_field_copy(0, s->rows * sizeof(double));
  if (s->ndof == 1) { dst.upload(host, s->rows); EFV_CUDA(cudaStreamSynchronize(stream())); return; }
  DBuf<double> tmp;
  tmp.upload(host, s->rows);
  EFV_LAUNCH(k_field_to_vertex_major, grid_for(s->rows, 256), 256, 0, s->nV, s->ndof, tmp.p, dst.p, 0);
  EFV_CUDA(cudaStreamSynchronize(stream()));
}
void vec_download(const efv_system *s, const DBuf<double> &src, double *host) {
  EFV_REQUIRE(src.n, "vector not allocated");
  count_field_copy(1, s->rows * sizeof(double));
  if (s->ndof == 1) { src.download(host, s->rows); return; }
  DBuf<double> tmp(s->rows);
  EFV_LAUNCH(k_field_to_vertex_major, grid_for(s->rows, 256), 256, 0, s->nV, s->ndof, src.p, tmp.p, 1);
  tmp.download(host, s->rows);
}

}  // namespace efv

