// assemble.cu -- north_star kernels 3 and 4: fused FP64 assembly with a deterministic row-wise reduction,
// Dirichlet row replacement / symmetric elimination, Neumann facet terms, right-hand sides.
//
// Assembly is ROW-CENTRIC: a group of GROUP lanes owns one graph row (vertex v).  It walks the incidence list
// of v in ascending element order; for each incident (element e, local vertex l) it evaluates the inner faces
// that touch l from the SoA geometry (Jacobian inverse + area vector per face), lane j producing the
// coefficient of column conn[e][j], and adds it into a shared-memory copy of the row at the position given by
// the byte slot map.  Every CSR entry is therefore summed by exactly one lane group in a fixed order: no float
// atomics, bit-reproducible, values written with coalesced stores.  (apps/heat_transfer.py:41-68)
#include "views.cuh"

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
  v.ndof = s->ndof; v.nV = s->nV; v.nnzg = s->nnzg;
  v.gptr = s->gptr.p; v.gcol = s->gcol.p; v.slotpos = s->slotpos.p; v.diag_slot = s->diag_slot.p; v.vals = s->vals.p;
  return v;
}

// shared-memory copy of the inner-face derivative tables of all shapes (dynamic face index per lane)
struct SmemTables {
  double ifD_tri[3][3][2], ifD_quad[4][4][2], ifD_tet[6][4][3], ifD_hex[12][8][3];
  int vface_tri[3][2], vface_quad[4][2], vface_tet[4][3], vface_hex[8][3];
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
  __syncthreads();
}
__device__ __forceinline__ int tab_vface(const SmemTables &T, int code, int l, int q) {
  switch (code) {
    case 0: return T.vface_tri[l][q];
    case 1: return T.vface_quad[l][q];
    case 2: return T.vface_tet[l][q];
    default: return T.vface_hex[l][q];
  }
}
__device__ __forceinline__ double tab_ifD(const SmemTables &T, int code, int f, int j, int a) {
  switch (code) {
    case 0: return T.ifD_tri[f][j][a];
    case 1: return T.ifD_quad[f][j][a];
    case 2: return T.ifD_tet[f][j][a];
    default: return T.ifD_hex[f][j][a];
  }
}

// t[a] = sum_b Jinv[b][a] s[b]  so that  (G^T s)_j = sum_a D[j][a] t[a]   (InnerFace.py:25 + heat_transfer.py:47)
template <int D>
__device__ __forceinline__ void face_t(const double *__restrict__ Jinv, const double *__restrict__ S, int f, int64_t n,
                                       int64_t gi, double (&t)[3]) {
  double R[D][D], s[D];
#pragma unroll
  for (int b = 0; b < D; ++b) {
    s[b] = S[(int64_t)(f * 3 + b) * n + gi];
#pragma unroll
    for (int a = 0; a < D; ++a) R[b][a] = Jinv[(int64_t)(f * D * D + b * D + a) * n + gi];
  }
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < D; ++b) acc += R[b][a] * s[b];
    t[a] = acc;
  }
  if (D == 2) t[2] = 0.0;
}

// props per region: [0] conductivity, [1] rho*cp/dt (0 when steady), [2] heat generation
template <int GROUP>
__global__ void __launch_bounds__(256) k_heat_rows(MeshView m, GeomView g, SysView s, const double *__restrict__ props,
                                                   int transient, int max_row_len, double *__restrict__ gen_out,
                                                   double *__restrict__ acc_out) {
  extern __shared__ double smem_dyn[];
  __shared__ SmemTables T;
  load_tables(T);
  constexpr int ROWS = 256 / GROUP;
  int grp = threadIdx.x / GROUP, gl = threadIdx.x % GROUP;
  double *row = smem_dyn + (size_t)grp * max_row_len;
  unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << ((threadIdx.x & 31) / GROUP * GROUP));
  for (int64_t v0 = (int64_t)blockIdx.x * ROWS; v0 < m.nV; v0 += (int64_t)gridDim.x * ROWS) {
    int64_t v = v0 + grp;
    if (v < m.nV) {
      int64_t r0 = s.gptr[v];
      int len = (int)(s.gptr[v + 1] - r0);
      for (int t = gl; t < len; t += GROUP) row[t] = 0.0;
      __syncwarp(gmask);
      double gen = 0.0, accv = 0.0;
      for (int64_t p = m.inc_ptr[v]; p < m.inc_ptr[v + 1]; ++p) {
        uint32_t pk = m.inc[p];
        int64_t e = pk >> 3;
        int l = pk & 7;
        int code = m.code(e);
        int mm = shape_m(code), nvf = shape_nvf(code);
        int64_t gi = m.group_index(e), n = g.n[code];
        const double *pr = props + 3 * m.region[e];
        double cond = pr[0];
        double sum = 0.0;
        for (int q = 0; q < nvf; ++q) {
          int enc = tab_vface(T, code, l, q);
          int f = enc >> 1;
          double t[3];
          if (code < 2) face_t<2>(g.Jinv[code], g.S[code], f, n, gi, t);
          else face_t<3>(g.Jinv[code], g.S[code], f, n, gi, t);
          if (gl < mm) {
            double w = tab_ifD(T, code, f, gl, 0) * t[0] + tab_ifD(T, code, f, gl, 1) * t[1];
            if (code >= 2) w += tab_ifD(T, code, f, gl, 2) * t[2];
            w *= cond;
            sum += (enc & 1) ? w : -w;          // forward vertex row gets +flux, backward -flux (heat_transfer.py:52-54)
          }
        }
        if (gl < mm) row[s.slotpos[m.slot_begin(e) + (int64_t)l * mm + gl]] += sum;
        if (gl == 0) {
          double vol = g.vol[code][(int64_t)l * n + gi];
          gen += vol * pr[2];
          accv += vol * pr[1];
        }
        __syncwarp(gmask);
      }
      if (gl == 0) {
        if (transient) row[s.diag_slot[v] - r0] += accv;       // heat_transfer.py:62-67
        gen_out[v] = gen;
        acc_out[v] = accv;
      }
      __syncwarp(gmask);
      for (int t = gl; t < len; t += GROUP) s.vals[r0 + t] = row[t];
      __syncwarp(gmask);
    }
  }
}

void heat_assemble(efv_system *s, const double *props_host, double dt, int transient) {
  efv_mesh *m = s->mesh;
  EFV_REQUIRE(s->ndof == 1, "heat needs a 1-DOF system");
  geometry_build(m);
  std::vector<double> pr(3 * m->n_regions);
  for (int r = 0; r < m->n_regions; ++r) {
    const double *p = props_host + 4 * r;   // Conductivity, Density, HeatCapacity, HeatGeneration
    pr[3 * r + 0] = p[0];
    pr[3 * r + 1] = transient ? p[1] * p[2] / dt : 0.0;    // heat_transfer.py:61
    pr[3 * r + 2] = p[3];
  }
  DBuf<double> dprops;
  dprops.upload(pr.data(), pr.size());
  if (!s->gen.n) { s->gen.alloc(s->nV); s->acc.alloc(s->nV); }
  cudaEvent_t e0, e1;
  EFV_CUDA(cudaEventCreate(&e0)); EFV_CUDA(cudaEventCreate(&e1));
  EFV_CUDA(cudaEventRecord(e0, stream()));
  MeshView mv = view_of(m);
  GeomView gv = geom_of(m);
  SysView sv = sys_view(s);
  if (m->max_m <= 4) {
    constexpr int GROUP = 4;
    size_t smem = (size_t)(256 / GROUP) * s->max_row_len * sizeof(double);
    EFV_CUDA(cudaFuncSetAttribute(k_heat_rows<GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EFV_LAUNCH(k_heat_rows<GROUP>, grid_for(s->nV * GROUP, 256, 8), 256, smem, mv, gv, sv, dprops.p, transient, s->max_row_len, s->gen.p, s->acc.p);
  } else {
    constexpr int GROUP = 8;
    size_t smem = (size_t)(256 / GROUP) * s->max_row_len * sizeof(double);
    EFV_CUDA(cudaFuncSetAttribute(k_heat_rows<GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EFV_LAUNCH(k_heat_rows<GROUP>, grid_for(s->nV * GROUP, 256, 8), 256, smem, mv, gv, sv, dprops.p, transient, s->max_row_len, s->gen.p, s->acc.p);
  }
  EFV_CUDA(cudaEventRecord(e1, stream()));
  EFV_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  EFV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  s->t_assemble_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  s->dirichlet_mode = -1;
  s->physics = 0;
}

void elasticity_assemble(efv_system *, const double *, int) { throw Error("elasticity assembly: not implemented yet"); }
void biot_assemble(efv_system *, const double *, double, const double *) { throw Error("Biot assembly: not implemented yet"); }

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

static int g_dir_counter = 0;

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
  EFV_CUDA(cudaMemsetAsync(s->dir_last.p, 0xff, s->dir_last.bytes(), stream()));   // -1
  s->static_rhs.zero();
  s->elim.zero();
  s->has_dirichlet = false;
  g_dir_counter = 0;
}

void bc_dirichlet(efv_system *s, const int64_t *rows, const double *values, int64_t n) {
  if (!s->dir_flag.n) bc_clear(s);
  if (n == 0) return;
  EFV_REQUIRE(g_dir_counter + n < (int64_t(1) << 31), "too many Dirichlet entries");
  DBuf<int64_t> dr;
  DBuf<double> dv;
  dr.upload(rows, n);
  dv.upload(values, n);
  EFV_LAUNCH(k_dir_last, grid_for(n, 256), 256, 0, dr.p, n, g_dir_counter, s->nV, s->ndof, s->dir_last.p);
  EFV_LAUNCH(k_dir_set, grid_for(n, 256), 256, 0, dr.p, dv.p, n, g_dir_counter, s->nV, s->ndof, s->dir_last.p,
             s->dir_flag.p, s->dir_value.p);
  EFV_CUDA(cudaStreamSynchronize(stream()));
  g_dir_counter += (int)n;
  s->has_dirichlet = true;
}

// Neumann: one thread per boundary vertex; its outer faces are visited in ascending outer-face id, i.e. in the
// reference's facet order (heat_transfer.py:115-120) -> deterministic
__global__ void k_neumann(int64_t nbv, const int32_t *__restrict__ bv, const int64_t *__restrict__ bvo_ptr,
                          const int64_t *__restrict__ bvo, const double *__restrict__ norm, int64_t o0, int64_t o1,
                          int ndof, int dof, double coef, double *__restrict__ rhs) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nbv; k += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    bool any = false;
    for (int64_t p = bvo_ptr[k]; p < bvo_ptr[k + 1]; ++p) {
      int64_t o = bvo[p];
      if (o >= o0 && o < o1) { acc += coef * norm[o]; any = true; }
    }
    if (any) rhs[(int64_t)bv[k] * ndof + dof] += acc;
  }
}

void bc_neumann(efv_system *s, int boundary, int dof, double value, double sign) {
  efv_mesh *m = s->mesh;
  Boundaries &B = m->bnd;
  EFV_REQUIRE(boundary >= 0 && boundary < B.nB, "bad boundary index");
  EFV_REQUIRE(dof >= 0 && dof < s->ndof, "bad dof");
  if (!s->dir_flag.n) bc_clear(s);
  if (value == 0.0 || B.nbv == 0) return;
  EFV_LAUNCH(k_neumann, grid_for(B.nbv, 128), 128, 0, B.nbv, B.bv.p, B.bvo_ptr.p, B.bvo.p, B.outer_area_norm.p,
             B.outer_ptr_host[boundary], B.outer_ptr_host[boundary + 1], s->ndof, dof, sign * value, s->static_rhs.p);
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

void build_rhs(efv_system *s, int transient) {
  ensure_vectors(s);
  if (!s->dir_flag.n) bc_clear(s);
  int sym = s->dirichlet_mode == 1;
  if (s->physics == 0) {
    EFV_REQUIRE(s->gen.n, "efv_heat_assemble first");
    EFV_LAUNCH(k_rhs_heat, grid_for(s->rows, 256), 256, 0, s->rows, s->gen.p, s->acc.p, s->xold.p, s->static_rhs.p,
               s->elim.p, s->dir_flag.p, s->dir_value.p, transient, sym, s->b.p);
  } else {
    EFV_LAUNCH(k_rhs_static, grid_for(s->rows, 256), 256, 0, s->rows, s->gen.n ? s->gen.p : nullptr, s->static_rhs.p,
               s->elim.p, s->dir_flag.p, s->dir_value.p, sym, s->b.p);
  }
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
  if (s->ndof == 1) { dst.upload(host, s->rows); EFV_CUDA(cudaStreamSynchronize(stream())); return; }
  DBuf<double> tmp;
  tmp.upload(host, s->rows);
  EFV_LAUNCH(k_field_to_vertex_major, grid_for(s->rows, 256), 256, 0, s->nV, s->ndof, tmp.p, dst.p, 0);
  EFV_CUDA(cudaStreamSynchronize(stream()));
}
void vec_download(const efv_system *s, const DBuf<double> &src, double *host) {
  EFV_REQUIRE(src.n, "vector not allocated");
  if (s->ndof == 1) { src.download(host, s->rows); return; }
  DBuf<double> tmp(s->rows);
  EFV_LAUNCH(k_field_to_vertex_major, grid_for(s->rows, 256), 256, 0, s->nV, s->ndof, src.p, tmp.p, 1);
  tmp.download(host, s->rows);
}

}  // namespace efv
