// mesh.cu -- device mesh: upload, vertex->element incidence, boundary facets (owner match, outward areas,
// outer faces).  Replaces the object graph built by PyEFVLib/Grid.py:18-65 with flat arrays.
#include "views.cuh"
#include <algorithm>

namespace efv {

// ---- input validation: vertex ids must lie in [0, nV) (fail loudly instead of faulting later) --------------------------
__global__ void k_check_range(const int32_t *__restrict__ a, int64_t n, int64_t nV, int *__restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (a[i] < 0 || a[i] >= nV) atomicExch(bad, 1);
}
static void check_vertex_ids(const int32_t *dev, int64_t n, int64_t nV, const char *what) {
  if (!n) return;
  DBuf<int> bad(1);
  bad.zero();
  EFV_LAUNCH(k_check_range, grid_for(n, 256), 256, 0, dev, n, nV, bad.p);
  int h = 0;
  bad.download(&h, 1);
  EFV_REQUIRE(h == 0, std::string(what) + " refers to a vertex id outside [0, numberOfVertices)");
}

__global__ void k_max_count(const int32_t *__restrict__ a, int64_t n, int *__restrict__ out) {
  int mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mx = max(mx, a[i]);
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// ---- incidence ------------------------------------------------------------------------------------------------
__global__ void k_count_incidence(const int32_t *__restrict__ conn, int64_t n, int32_t *__restrict__ cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[conn[i]], 1);
}

__global__ void k_fill_incidence(MeshView m, int32_t *__restrict__ cursor, uint32_t *__restrict__ inc) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < m.nE; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = m.begin(e);
    int mm = shape_m(m.code(e));
    for (int l = 0; l < mm; ++l) {
      int v = m.conn[b + l];
      int pos = atomicAdd(&cursor[v], 1);
      inc[m.inc_ptr[v] + pos] = inc_pack(e, l);
    }
  }
}

// deterministic order: ascending (local vertex, element) inside every vertex list
__global__ void k_sort_incidence(int64_t nV, const int64_t *__restrict__ inc_ptr, uint32_t *__restrict__ inc) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
    int64_t a = inc_ptr[v], b = inc_ptr[v + 1];
    for (int64_t i = a + 1; i < b; ++i) {
      uint32_t key = inc[i];
      int64_t j = i - 1;
      while (j >= a && inc[j] > key) { inc[j + 1] = inc[j]; --j; }
      inc[j + 1] = key;
    }
  }
}

static int code_for(int mm, int dim) {
  if (mm == 3 && dim == 2) return EFV_TRI;
  if (mm == 4) return (dim == 2) ? EFV_QUAD : EFV_TET;
  if (mm == 8 && dim == 3) return EFV_HEX;
  if (mm == 6 && dim == 3) return EFV_PRISM;
  if (mm == 5 && dim == 3) return EFV_PYRAMID;
  throw Error("This element has not been registered in Grid yet");   // Element.py:27
}

// ep: host element pointer array (nE+1) or nullptr when every element has uniform_m vertices
void mesh_finalize(efv_mesh *m, const int64_t *ep, int uniform_m) {
  // classify shapes on the host (nE integer ops), build groups
  const int64_t nE = m->nE;
  int64_t counts[kNumShapes] = {0, 0, 0, 0, 0, 0};
  std::vector<uint8_t> shape;
  m->max_m = 0;
  m->n_inner_faces = 0;
  if (ep && !uniform_m) {               // uniform stride written out explicitly? (the common case for big meshes)
    const int64_t m0 = ep[1] - ep[0];
    bool same = true;
    for (int64_t e = 0; e < nE && same; ++e) same = (ep[e + 1] - ep[e]) == m0;
    if (same) uniform_m = (int)m0;
  }
  if (uniform_m) {
    const int code = code_for(uniform_m, m->dim);
    counts[code] = nE;
    m->max_m = uniform_m;
    m->n_inner_faces = nE * shape_nf(code);
  } else {
    shape.resize(nE);
    for (int64_t e = 0; e < nE; ++e) {
      int mm = (int)(ep[e + 1] - ep[e]);
      int code = code_for(mm, m->dim);
      shape[e] = (uint8_t)code;
      counts[code]++;
      m->max_m = std::max(m->max_m, mm);
      m->n_inner_faces += shape_nf(code);
    }
  }
  EFV_REQUIRE(nE < (int64_t(1) << kIncShift), "mesh too large for the packed incidence (2^29 elements)");
  int nonempty = 0, last = -1;
  for (int c = 0; c < kNumShapes; ++c)
    if (counts[c]) { nonempty++; last = c; }
  m->uniform_code = (nonempty == 1) ? last : -1;
  for (int c = 0; c < kNumShapes; ++c) { m->groups[c].code = c; m->groups[c].n = counts[c]; }
  if (m->uniform_code < 0) {
    std::vector<int32_t> gidx(nE);
    std::vector<std::vector<int32_t>> ids(kNumShapes);
    std::vector<int64_t> slot_ptr(nE + 1, 0);
    for (int64_t e = 0; e < nE; ++e) {
      gidx[e] = (int32_t)ids[shape[e]].size();
      ids[shape[e]].push_back((int32_t)e);
      int mm = shape_m(shape[e]);
      slot_ptr[e + 1] = slot_ptr[e] + mm * mm;
    }
    m->elem_shape.upload(shape.data(), nE);
    m->elem_gidx.upload(gidx.data(), nE);
    m->slot_ptr.upload(slot_ptr.data(), nE + 1);
    m->elem_ptr.upload(ep, nE + 1);
    for (int c = 0; c < kNumShapes; ++c)
      if (counts[c]) m->groups[c].ids.upload(ids[c].data(), ids[c].size());
    EFV_CUDA(cudaStreamSynchronize(stream()));
  }
  check_vertex_ids(m->conn.p, m->conn_size, m->nV, "element connectivity");
  // incidence
  MeshView mv = view_of(m);
  DBuf<int32_t> cnt(m->nV);
  cnt.zero();
  EFV_LAUNCH(k_count_incidence, grid_for(m->conn_size, 256), 256, 0, m->conn.p, m->conn_size, cnt.p);
  m->inc_ptr.alloc(m->nV + 1);
  int64_t total = exclusive_scan_i32_to_i64(cnt.p, m->inc_ptr.p, m->nV);
  EFV_REQUIRE(total == m->conn_size, "incidence scan mismatch");
  {
    DBuf<int> mx(1);
    mx.zero();
    EFV_LAUNCH(k_max_count, grid_for(m->nV, 256), 256, 0, cnt.p, m->nV, mx.p);
    mx.download(&m->max_incidence, 1);
  }
  m->inc.alloc(total);
  cnt.zero();
  mv = view_of(m);
  EFV_LAUNCH(k_fill_incidence, grid_for(nE, 256), 256, 0, mv, cnt.p, m->inc.p);
  EFV_LAUNCH(k_sort_incidence, grid_for(m->nV, 128), 128, 0, m->nV, m->inc_ptr.p, m->inc.p);
  EFV_CUDA(cudaStreamSynchronize(stream()));
}

// ---- boundaries --------------------------------------------------------------------------------------------------
// one thread per facet: owning element = first element (grid order) containing all facet vertices
// (Boundary.py:37-41), local facet by vertex-set match (Boundary.py:44-47), area vector (Facet.py:22-36),
// outward correction (Boundary.py:56-60), outer faces (Facet.py:41-45, Boundary.py:62-70, OuterFace.py:18-19)
template <class Tab>
__device__ bool match_local_facet(const int *loc, int mf, int &lf_out) {
  for (int lf = 0; lf < Tab::NFACETS; ++lf) {
    if (Tab::facetN(lf) != mf) continue;            // prisms and pyramids have 3- and 4-vertex facets
    bool all = true;
    for (int k = 0; k < mf && all; ++k) {
      bool found = false;
      for (int q = 0; q < mf; ++q) found |= (Tab::facetV(lf, q) == loc[k]);
      all &= found;
    }
    if (all) { lf_out = lf; return true; }
  }
  return false;
}
template <class Tab>
__device__ int local_pos_in_facet(int lf, int local_vertex) {
  for (int q = 0; q < Tab::MF; ++q)
    if (Tab::facetV(lf, q) == local_vertex) return q;
  return -1;
}

__global__ void k_match_facets(MeshView m, int64_t nF, const int64_t *__restrict__ facet_ptr,
                               const int32_t *__restrict__ facet_conn, int64_t *__restrict__ facet_element,
                               int32_t *__restrict__ facet_local, double *__restrict__ facet_area,
                               int32_t *__restrict__ outer_vertex, int32_t *__restrict__ outer_local,
                               int64_t *__restrict__ outer_facet, double *__restrict__ outer_norm,
                               int *__restrict__ error_flag) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x) {
    int64_t fb = facet_ptr[f];
    int mf = (int)(facet_ptr[f + 1] - fb);
    int fv[4];
    for (int k = 0; k < mf; ++k) fv[k] = facet_conn[fb + k];
    // owner search through the incidence list of the first facet vertex (ascending element id)
    int64_t owner = -1;
    int loc[4] = {0, 0, 0, 0};
    for (int64_t p = m.inc_ptr[fv[0]]; p < m.inc_ptr[fv[0] + 1]; ++p) {
      int64_t e = inc_elem(m.inc[p]);
      if (owner >= 0 && e > owner) continue;          // keep the FIRST element in grid order (Boundary.py:37-41)
      int64_t b = m.begin(e);
      int mm = shape_m(m.code(e));
      int cand[4] = {0, 0, 0, 0};
      bool all = true;
      for (int k = 0; k < mf && all; ++k) {
        int where = -1;
        for (int j = 0; j < mm; ++j)
          if (m.conn[b + j] == fv[k]) where = j;
        cand[k] = where;
        all &= (where >= 0);
      }
      if (all) {
        owner = e;
        for (int k = 0; k < 4; ++k) loc[k] = cand[k];
      }
    }
    if (owner < 0) { atomicExch(error_flag, 1); continue; }
    int code = m.code(owner);
    int lf = -1;
    bool ok = false;
    switch (code) {
      case 0: ok = match_local_facet<TriangleTab>(loc, mf, lf); break;
      case 1: ok = match_local_facet<QuadrilateralTab>(loc, mf, lf); break;
      case 2: ok = match_local_facet<TetrahedronTab>(loc, mf, lf); break;
      case 3: ok = match_local_facet<HexahedronTab>(loc, mf, lf); break;
      case 4: ok = match_local_facet<PrismTab>(loc, mf, lf); break;
      default: ok = match_local_facet<PyramidTab>(loc, mf, lf); break;
    }
    if (!ok) { atomicExch(error_flag, 2); continue; }
    facet_element[f] = owner;
    facet_local[f] = lf;
    // area vector
    double P[4][3];
    for (int k = 0; k < mf; ++k)
      for (int a = 0; a < 3; ++a) P[k][a] = m.coords[(int64_t)fv[k] * 3 + a];
    double ax, ay, az;
    if (mf == 2) {                         // Facet.py:23-24
      ax = P[0][1] - P[1][1]; ay = P[1][0] - P[0][0]; az = 0.0;
    } else if (mf == 3) {                  // Facet.py:26-28
      double u[3], w[3];
      for (int a = 0; a < 3; ++a) { u[a] = P[1][a] - P[0][a]; w[a] = P[2][a] - P[0][a]; }
      ax = (u[1] * w[2] - u[2] * w[1]) / 2; ay = (u[2] * w[0] - u[0] * w[2]) / 2; az = (u[0] * w[1] - u[1] * w[0]) / 2;
    } else {                               // Facet.py:30-36
      double cm[3], lr[3];
      for (int a = 0; a < 3; ++a) { cm[a] = P[2][a] - P[0][a]; lr[a] = P[3][a] - P[1][a]; }
      ax = 0.5 * (cm[1] * lr[2] - cm[2] * lr[1]);
      ay = 0.5 * (cm[2] * lr[0] - cm[0] * lr[2]);
      az = 0.5 * (cm[0] * lr[1] - cm[1] * lr[0]);
    }
    // outward correction with the first element vertex that is not on the facet (Boundary.py:56-60)
    double c[3] = {0, 0, 0};
    for (int k = 0; k < mf; ++k)
      for (int a = 0; a < 3; ++a) c[a] += P[k][a];
    for (int a = 0; a < 3; ++a) c[a] /= mf;
    int64_t b = m.begin(owner);
    int mm = shape_m(code);
    int inner = -1;
    for (int j = 0; j < mm && inner < 0; ++j) {
      int v = m.conn[b + j];
      bool onf = false;
      for (int k = 0; k < mf; ++k) onf |= (fv[k] == v);
      if (!onf) inner = v;
    }
    double dot = ax * (c[0] - m.coords[(int64_t)inner * 3]) + ay * (c[1] - m.coords[(int64_t)inner * 3 + 1]) +
                 az * (c[2] - m.coords[(int64_t)inner * 3 + 2]);
    if (dot < 0) { ax = -ax; ay = -ay; az = -az; }
    facet_area[f * 3] = ax; facet_area[f * 3 + 1] = ay; facet_area[f * 3 + 2] = az;
    double ox = ax / mf, oy = ay / mf, oz = az / mf;          // OuterFace.py:18-19
    double nrm = sqrt(ox * ox + oy * oy + oz * oz);
    for (int k = 0; k < mf; ++k) {
      int64_t o = fb + k;
      outer_vertex[o] = fv[k];
      outer_facet[o] = f;
      outer_norm[o] = nrm;
      int pos;
      switch (code) {
        case 0: pos = local_pos_in_facet<TriangleTab>(lf, loc[k]); break;
        case 1: pos = local_pos_in_facet<QuadrilateralTab>(lf, loc[k]); break;
        case 2: pos = local_pos_in_facet<TetrahedronTab>(lf, loc[k]); break;
        case 3: pos = local_pos_in_facet<HexahedronTab>(lf, loc[k]); break;
        case 4: pos = local_pos_in_facet<PrismTab>(lf, loc[k]); break;
        default: pos = local_pos_in_facet<PyramidTab>(lf, loc[k]); break;
      }
      outer_local[o] = pos;
    }
  }
}

__global__ void k_flag_vertices(const int32_t *__restrict__ ov, int64_t n, int32_t *__restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flags[ov[i]] = 1;
}
__global__ void k_compact_vertices(int64_t nV, const int32_t *__restrict__ flags, const int64_t *__restrict__ cidx, int32_t *__restrict__ bv,
                                   int32_t *__restrict__ vtx_to_bv) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nV; v += (int64_t)gridDim.x * blockDim.x) {
    if (flags[v]) bv[cidx[v]] = (int32_t)v;
    vtx_to_bv[v] = flags[v] ? (int32_t)cidx[v] : -1;
  }
}
__global__ void k_count_outer(const int32_t *__restrict__ ov, int64_t n, const int64_t *__restrict__ cidx, int32_t *__restrict__ cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[cidx[ov[i]]], 1);
}
__global__ void k_fill_outer(const int32_t *__restrict__ ov, int64_t n, const int64_t *__restrict__ cidx,
                             const int64_t *__restrict__ ptr, int32_t *__restrict__ cursor, int64_t *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = cidx[ov[i]];
    out[ptr[k] + atomicAdd(&cursor[k], 1)] = i;
  }
}
__global__ void k_sort_lists_i64(int64_t n, const int64_t *__restrict__ ptr, int64_t *__restrict__ list) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t a = ptr[k], b = ptr[k + 1];
    for (int64_t i = a + 1; i < b; ++i) {
      int64_t key = list[i];
      int64_t j = i - 1;
      while (j >= a && list[j] > key) { list[j + 1] = list[j]; --j; }
      list[j + 1] = key;
    }
  }
}

void mesh_set_boundaries(efv_mesh *m, int nB, const int64_t *boundary_ptr, int64_t nF, const int64_t *facet_ptr,
                         const int32_t *facet_conn) {
  Boundaries &B = m->bnd;
  B.nB = nB;
  B.nF = nF;
  B.boundary_ptr.assign(boundary_ptr, boundary_ptr + nB + 1);
  EFV_REQUIRE(B.boundary_ptr[nB] == nF, "boundary_ptr does not cover the facets");
  B.nOuter = facet_ptr[nF];
  B.outer_ptr_host.resize(nB + 1);
  for (int b = 0; b <= nB; ++b) B.outer_ptr_host[b] = facet_ptr[boundary_ptr[b]];
  B.facet_ptr.upload(facet_ptr, nF + 1);
  B.facet_conn.upload(facet_conn, B.nOuter);
  check_vertex_ids(B.facet_conn.p, B.nOuter, m->nV, "boundary connectivity");
  for (int64_t f = 0; f < nF; ++f) {
    const int64_t mf = facet_ptr[f + 1] - facet_ptr[f];
    EFV_REQUIRE(mf >= 2 && mf <= 4 && mf == (m->dim == 2 ? 2 : mf) && (m->dim == 3 ? mf >= 3 : true), "boundary facet with an unsupported vertex count");
  }
  B.facet_element.alloc(nF);
  B.facet_local.alloc(nF);
  B.facet_area.alloc(nF * 3);
  B.outer_vertex.alloc(B.nOuter);
  B.outer_local.alloc(B.nOuter);
  B.outer_facet.alloc(B.nOuter);
  B.outer_area_norm.alloc(B.nOuter);
  if (nF == 0) return;
  DBuf<int> err(1);
  err.zero();
  MeshView mv = view_of(m);
  EFV_LAUNCH(k_match_facets, grid_for(nF, 128), 128, 0, mv, nF, B.facet_ptr.p, B.facet_conn.p, B.facet_element.p,
             B.facet_local.p, B.facet_area.p, B.outer_vertex.p, B.outer_local.p, B.outer_facet.p,
             B.outer_area_norm.p, err.p);
  int herr = 0;
  err.download(&herr, 1);
  EFV_REQUIRE(herr == 0, herr == 1 ? "a boundary facet has no owning element" : "a boundary facet matches no local facet of its element");
  // compact boundary-vertex -> outer faces map
  DBuf<int32_t> flags(m->nV);
  flags.zero();
  EFV_LAUNCH(k_flag_vertices, grid_for(B.nOuter, 256), 256, 0, B.outer_vertex.p, B.nOuter, flags.p);
  DBuf<int64_t> cidx(m->nV + 1);
  B.nbv = exclusive_scan_i32_to_i64(flags.p, cidx.p, m->nV);
  B.bv.alloc(B.nbv);
  B.vtx_to_bv.alloc(m->nV);
  EFV_LAUNCH(k_compact_vertices, grid_for(m->nV, 256), 256, 0, m->nV, flags.p, cidx.p, B.bv.p, B.vtx_to_bv.p);
  DBuf<int32_t> cnt(B.nbv);
  cnt.zero();
  EFV_LAUNCH(k_count_outer, grid_for(B.nOuter, 256), 256, 0, B.outer_vertex.p, B.nOuter, cidx.p, cnt.p);
  B.bvo_ptr.alloc(B.nbv + 1);
  int64_t tot = exclusive_scan_i32_to_i64(cnt.p, B.bvo_ptr.p, B.nbv);
  EFV_REQUIRE(tot == B.nOuter, "outer-face map scan mismatch");
  B.bvo.alloc(B.nOuter);
  cnt.zero();
  EFV_LAUNCH(k_fill_outer, grid_for(B.nOuter, 256), 256, 0, B.outer_vertex.p, B.nOuter, cidx.p, B.bvo_ptr.p, cnt.p, B.bvo.p);
  EFV_LAUNCH(k_sort_lists_i64, grid_for(B.nbv, 128), 128, 0, B.nbv, B.bvo_ptr.p, B.bvo.p);
  EFV_CUDA(cudaStreamSynchronize(stream()));
}

}  // namespace efv
