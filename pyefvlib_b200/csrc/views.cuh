// views.cuh -- POD views of the device mesh passed to kernels by value, and shape dispatch helpers.
#pragma once
#include "efv_internal.cuh"
#include "shape_tables.inc"

namespace efv {

struct MeshView {
  int dim, uniform_code, uniform_m;
  int64_t nV, nE;
  const double *coords;
  const int64_t *elem_ptr;
  const int32_t *conn;
  const uint8_t *shape;
  const int32_t *gidx;
  const int32_t *region;
  const int64_t *inc_ptr;
  const uint32_t *inc;
  const int64_t *slot_ptr;
  __device__ __forceinline__ int64_t begin(int64_t e) const { return uniform_m ? e * uniform_m : elem_ptr[e]; }
  __device__ __forceinline__ int code(int64_t e) const { return uniform_code >= 0 ? uniform_code : (int)shape[e]; }
  __device__ __forceinline__ int64_t group_index(int64_t e) const { return uniform_code >= 0 ? e : (int64_t)gidx[e]; }
  __device__ __forceinline__ int64_t slot_begin(int64_t e) const {
    return uniform_m ? e * (int64_t)(uniform_m * uniform_m) : slot_ptr[e];
  }
};

// packed incidence entry: local vertex in the top 3 bits, element id below => a vertex list sorted ascending is ordered by
// (local vertex, element): the deterministic reduction order of the row kernels (independent of the partition)
constexpr int kIncShift = 29;
__host__ __device__ __forceinline__ uint32_t inc_pack(int64_t e, int l) { return ((uint32_t)l << kIncShift) | (uint32_t)e; }
__host__ __device__ __forceinline__ int64_t inc_elem(uint32_t pk) { return (int64_t)(pk & ((1u << kIncShift) - 1u)); }
__host__ __device__ __forceinline__ int inc_local(uint32_t pk) { return (int)(pk >> kIncShift); }

// shape codes: 0 triangle, 1 quadrilateral, 2 tetrahedron, 3 hexahedron, 4 prism, 5 pyramid
__host__ __device__ constexpr __forceinline__ int shape_m(int code) {
  return code == 0 ? 3 : (code == 3 ? 8 : (code == 4 ? 6 : (code == 5 ? 5 : 4)));
}
__host__ __device__ constexpr __forceinline__ int shape_nf(int code) {
  return code == 0 ? 3 : (code == 1 ? 4 : (code == 2 ? 6 : (code == 3 ? 12 : (code == 4 ? 9 : 8))));
}
__host__ __device__ constexpr __forceinline__ int shape_dim(int code) { return code < 2 ? 2 : 3; }
// inner faces touching a local vertex (the apex of a pyramid has 4, its base vertices 3: padded with -1 in the tables)
__host__ __device__ constexpr __forceinline__ int shape_nvf(int code) { return code < 2 ? 2 : (code == 5 ? 4 : 3); }

inline MeshView view_of(const efv_mesh *m) {
  MeshView v;
  v.dim = m->dim;
  v.uniform_code = m->uniform_code;
  v.uniform_m = m->uniform_code >= 0 ? shape_m(m->uniform_code) : 0;
  v.nV = m->nV;
  v.nE = m->nE;
  v.coords = m->coords.p;
  v.elem_ptr = m->elem_ptr.p;
  v.conn = m->conn.p;
  v.shape = m->elem_shape.p;
  v.gidx = m->elem_gidx.p;
  v.region = m->region.p;
  v.inc_ptr = m->inc_ptr.p;
  v.inc = m->inc.p;
  v.slot_ptr = m->slot_ptr.p;
  return v;
}

// per-group geometry view
struct GeomView {
  const double *Jinv[kNumShapes];
  const double *S[kNumShapes];
  const double *vol[kNumShapes];
  const double *St[kNumShapes];
  const double *Frec[kNumShapes];
  const double *Srec[kNumShapes];
  int64_t n[kNumShapes];
};
inline GeomView geom_of(const efv_mesh *m) {
  GeomView g;
  for (int c = 0; c < kNumShapes; ++c) {
    g.Jinv[c] = m->groups[c].Jinv.p;
    g.S[c] = m->groups[c].S.p;
    g.vol[c] = m->groups[c].vol.p;
    g.St[c] = m->groups[c].St.p;
    g.Frec[c] = m->groups[c].Frec.p;
    g.Srec[c] = m->groups[c].Srec.p;
    g.n[c] = m->groups[c].n;
  }
  return g;
}

// call f(Tab{}) for the table type of `code`
#define EFV_DISPATCH_SHAPE(code, ...)                                         \
  do {                                                                        \
    switch (code) {                                                           \
      case 0: { using Tab = TriangleTab; __VA_ARGS__; } break;                \
      case 1: { using Tab = QuadrilateralTab; __VA_ARGS__; } break;           \
      case 2: { using Tab = TetrahedronTab; __VA_ARGS__; } break;             \
      case 3: { using Tab = HexahedronTab; __VA_ARGS__; } break;              \
      case 4: { using Tab = PrismTab; __VA_ARGS__; } break;                   \
      case 5: { using Tab = PyramidTab; __VA_ARGS__; } break;                 \
      default: throw efv::Error("unknown shape code");                        \
    }                                                                         \
  } while (0)

}  // namespace efv
