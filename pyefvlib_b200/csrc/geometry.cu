// geometry.cu -- north_star kernel 1: the one-time geometry pass.  One thread per element, everything
// unrolled over the constant reference-element tables so that the table entries become constant-bank
// operands of the FP64 FMAs and the vertex coordinates stay in registers.  Output is component-major SoA
// (consecutive threads -> consecutive addresses): per inner face the Jacobian inverse and the area vector,
// per local vertex the sub-control volume.
//   J^T = D^T X           Element.py:69-71
//   inv(J^T)              InnerFace.py:23-25  (closed-form adjugate instead of LAPACK; ~1e-16 relative)
//   s_f                   Shape.py:29-33, 57-61 (2-D), 84-96 (tet), 119-139 (hex)
//   vol_k = T_k det(..)   Element.py:39-48
#include "views.cuh"

namespace efv {

template <int D> struct SmallMat;
template <> struct SmallMat<2> {
  __device__ static __forceinline__ double det(const double (&A)[2][2]) { return A[0][0] * A[1][1] - A[0][1] * A[1][0]; }
  __device__ static __forceinline__ void inv(const double (&A)[2][2], double (&R)[2][2]) {
    double id = 1.0 / det(A);
    R[0][0] = A[1][1] * id; R[0][1] = -A[0][1] * id;
    R[1][0] = -A[1][0] * id; R[1][1] = A[0][0] * id;
  }
};
template <> struct SmallMat<3> {
  __device__ static __forceinline__ double det(const double (&A)[3][3]) {
    return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  }
  __device__ static __forceinline__ void inv(const double (&A)[3][3], double (&R)[3][3]) {
    double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    double id = 1.0 / (A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02);
    R[0][0] = c00 * id; R[1][0] = c01 * id; R[2][0] = c02 * id;
    R[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    R[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    R[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    R[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    R[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    R[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
  }
};

template <class Tab>
__global__ void __launch_bounds__(128) k_geometry(MeshView m, const int32_t *__restrict__ ids, int64_t n,
                                                  double *__restrict__ Jinv, double *__restrict__ S,
                                                  double *__restrict__ vol, double *__restrict__ St) {
  constexpr int M = Tab::M, D = Tab::DIM, NF = Tab::NF;
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t e = ids ? (int64_t)ids[k] : k;
  int64_t b = m.begin(e);
  double X[M][3];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    int64_t v = m.conn[b + j];
#pragma unroll
    for (int a = 0; a < 3; ++a) X[j][a] = m.coords[v * 3 + a];
  }
  // element centroid: sum(vertices) / m  (Element.py:59); the pyramid's area vectors use the centre of its base
  // instead (Shape.py:213-216)
  constexpr int MC = (Tab::CODE == 5) ? 4 : M;
  double c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double sum = X[0][a];
#pragma unroll
    for (int j = 1; j < MC; ++j) sum += X[j][a];
    c[a] = sum / MC;
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    double Jt[D][D], R[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int q = 0; q < D; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) acc += Tab::ifD(f, j, a) * X[j][q];
        Jt[a][q] = acc;
      }
    SmallMat<D>::inv(Jt, R);
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int q = 0; q < D; ++q) Jinv[(int64_t)(f * D * D + a * D + q) * n + k] = R[a][q];
    // area vector
    const int vb = Tab::nbr(f, 0), vf = Tab::nbr(f, 1);
    double s[3];
    if constexpr (D == 2) {
      double ax = c[0] - (X[vb][0] + X[vf][0]) / 2.0;
      double ay = c[1] - (X[vb][1] + X[vf][1]) / 2.0;
      s[0] = ay; s[1] = -ax; s[2] = 0.0;
    } else {
      // p1 / p3: centroids of the two element facets that share the edge (triangles or quadrilaterals: Shape.py:84-96
      // tet, :119-139 hex, :162-185 prism, :208-247 pyramid); the neighbour table lists their remaining vertices
      const int NA = Tab::adjn(f, 0), NB = Tab::adjn(f, 1);        // compile-time after unrolling
      double p1[3], p2[3], p3[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        p2[a] = (X[vb][a] + X[vf][a]) / 2.0;
        double sa = X[vb][a] + X[vf][a], sb = sa;
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (q < NA) sa += X[Tab::nbr(f, 2 + q)][a];
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (q < NB) sb += X[Tab::nbr(f, (2 + NA + q) < Tab::NNBR ? (2 + NA + q) : 0)][a];
        p1[a] = sa / (2.0 + NA);
        p3[a] = sb / (2.0 + NB);
      }
      if (NB == 0) {
        // triangular face  centroid -> facet centroid -> edge midpoint  (pyramid base edges, Shape.py:222-231)
        double u[3], w[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { u[a] = p1[a] - c[a]; w[a] = p2[a] - c[a]; }
        s[0] = 0.5 * (u[1] * w[2] - w[1] * u[2]);
        s[1] = 0.5 * (w[0] * u[2] - u[0] * w[2]);
        s[2] = 0.5 * (u[0] * w[1] - w[0] * u[1]);
      } else {
        double u[3], w[3], y[3], z[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { u[a] = p1[a] - c[a]; w[a] = p3[a] - c[a]; y[a] = p3[a] - p2[a]; z[a] = p1[a] - p2[a]; }
        s[0] = 0.5 * (u[1] * w[2] - w[1] * u[2] + y[1] * z[2] - z[1] * y[2]);
        s[1] = 0.5 * (w[0] * u[2] - u[0] * w[2] + z[0] * y[2] - y[0] * z[2]);
        s[2] = 0.5 * (u[0] * w[1] - w[0] * u[1] + y[0] * z[1] - z[0] * y[1]);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) S[(int64_t)(f * 3 + a) * n + k] = s[a];
    // J^-T s : the heat flux coefficients are  k * D_f[j,:] . (J^-T s)   (InnerFace.py:25 + heat_transfer.py:47)
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) acc += R[q][a] * s[q];
      St[(int64_t)(f * D + a) * n + k] = acc;
    }
  }
#pragma unroll
  for (int l = 0; l < M; ++l) {
    double Js[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int q = 0; q < D; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) acc += Tab::subD(l, j, a) * X[j][q];
        Js[a][q] = acc;
      }
    vol[(int64_t)l * n + k] = Tab::subvol(l) * SmallMat<D>::det(Js);
  }
}

// lumped control-volume of every vertex, summed in element order (Element.py:46 / Vertex.py:9)
__global__ void k_vertex_volume(MeshView m, GeomView g, double *__restrict__ vv) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < m.nV; v += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int64_t p = m.inc_ptr[v]; p < m.inc_ptr[v + 1]; ++p) {
      uint32_t pk = m.inc[p];
      int64_t e = inc_elem(pk);
      int l = inc_local(pk);
      int code = m.code(e);
      acc += g.vol[code][(int64_t)l * g.n[code] + m.group_index(e)];
    }
    vv[v] = acc;
  }
}

void geometry_build(efv_mesh *m) {
  if (m->has_geometry) return;
  MeshView mv = view_of(m);
  for (int code = 0; code < kNumShapes; ++code) {
    ShapeGroup &g = m->groups[code];
    if (!g.n) continue;
    int d = shape_dim(code), nf = shape_nf(code), mm = shape_m(code);
    g.Jinv.alloc((size_t)nf * d * d * g.n);
    g.S.alloc((size_t)nf * 3 * g.n);
    g.vol.alloc((size_t)mm * g.n);
    g.St.alloc((size_t)nf * d * g.n);
    const int32_t *ids = (m->uniform_code >= 0) ? nullptr : g.ids.p;
    unsigned grid = (unsigned)div_up(g.n, 128);
    EFV_DISPATCH_SHAPE(code, EFV_LAUNCH(k_geometry<Tab>, grid, 128, 0, mv, ids, g.n, g.Jinv.p, g.S.p, g.vol.p, g.St.p));
  }
  m->vertex_volume.alloc(m->nV);
  EFV_LAUNCH(k_vertex_volume, grid_for(m->nV, 256), 256, 0, mv, geom_of(m), m->vertex_volume.p);
  m->has_geometry = true;
}

// ---- export (tests / parity): AoS per element in index-in-group order -----------------------------------------
template <class Tab>
__global__ void k_export_group(int64_t n, const double *__restrict__ Jinv, const double *__restrict__ S,
                               const double *__restrict__ vol, double *__restrict__ Gout, double *__restrict__ Jout,
                               double *__restrict__ Sout, double *__restrict__ Vout) {
  constexpr int M = Tab::M, D = Tab::DIM, NF = Tab::NF;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * NF) return;
  int64_t k = t / NF;
  int f = (int)(t % NF);
  double R[D][D];
  for (int a = 0; a < D; ++a)
    for (int q = 0; q < D; ++q) R[a][q] = Jinv[(int64_t)(f * D * D + a * D + q) * n + k];
  if (Jout)
    for (int a = 0; a < D; ++a)
      for (int q = 0; q < D; ++q) Jout[(t * D + a) * D + q] = R[a][q];
  if (Gout)   // G = inv(J^T) D^T : G[b][j] = sum_a R[b][a] D[j][a]   (InnerFace.py:25)
    for (int q = 0; q < D; ++q)
      for (int j = 0; j < M; ++j) {
        double acc = 0.0;
        for (int a = 0; a < D; ++a) acc += R[q][a] * Tab::ifD(f, j, a);
        Gout[(t * D + q) * M + j] = acc;
      }
  if (Sout)
    for (int a = 0; a < 3; ++a) Sout[t * 3 + a] = S[(int64_t)(f * 3 + a) * n + k];
  if (Vout && f == 0)
    for (int l = 0; l < M; ++l) Vout[k * M + l] = vol[(int64_t)l * n + k];
}

void geometry_export_group(const efv_mesh *m, int code, double *G, double *Jinv, double *S, double *vol) {
  EFV_REQUIRE(m->has_geometry, "call efv_geometry_build first");
  EFV_REQUIRE(code >= 0 && code < kNumShapes, "bad shape code");
  const ShapeGroup &g = m->groups[code];
  if (!g.n) return;
  int d = shape_dim(code), nf = shape_nf(code), mm = shape_m(code);
  int64_t nfaces = g.n * nf;
  DBuf<double> dG, dJ, dS, dV;
  if (G) dG.alloc((size_t)nfaces * d * mm);
  if (Jinv) dJ.alloc((size_t)nfaces * d * d);
  if (S) dS.alloc((size_t)nfaces * 3);
  if (vol) dV.alloc((size_t)g.n * mm);
  unsigned grid = (unsigned)div_up(nfaces, 128);
  EFV_DISPATCH_SHAPE(code, EFV_LAUNCH(k_export_group<Tab>, grid, 128, 0, g.n, g.Jinv.p, g.S.p, g.vol.p, dG.p, dJ.p, dS.p, dV.p));
  if (G) dG.download(G, dG.n);
  if (Jinv) dJ.download(Jinv, dJ.n);
  if (S) dS.download(S, dS.n);
  if (vol) dV.download(vol, dV.n);
}

void vertex_volume_export(const efv_mesh *m, double *vv) {
  EFV_REQUIRE(m->has_geometry, "call efv_geometry_build first");
  m->vertex_volume.download(vv, m->nV);
}

}  // namespace efv
