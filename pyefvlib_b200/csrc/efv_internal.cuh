// efv_internal.cuh -- device data model of libefv_b200.
//
// HBM layout (all SoA, sized for one B200):
//   mesh      coords f64[nV][3] | elem_ptr i64[nE+1] | conn i32[sum m] | elem_shape u8[nE] | elem_gidx i32[nE]
//             region i32[nE] | incidence: inc_ptr i64[nV+1], inc u32[sum m] = (element << 3 | local vertex),
//             sorted by element id inside each vertex list (deterministic reduction order)
//   geometry  per shape group g (elements of one shape, index-in-group k):
//             Jinv f64[NF*d*d][n_g] | S f64[NF*3][n_g] | vol f64[M][n_g]          (component-major = coalesced)
//   system    graph CSR gptr i64[nV+1], gcol i32[nnzg] (sorted) | slotpos u8[sum m*m] (position inside the row)
//             vals f64[nnzg][ndof][ndof] (BSR over the vertex graph) | vectors f64[nV][ndof] (vertex-major inside,
//             field-major at the C-ABI) | dirichlet flag u8 / value f64 per DOF
#pragma once
#include "common.cuh"

namespace efv {

constexpr int kNumShapes = 6;
constexpr int kMaxM = 8;
enum { EFV_TRI = 0, EFV_QUAD = 1, EFV_TET = 2, EFV_HEX = 3, EFV_PRISM = 4, EFV_PYRAMID = 5 };

struct ShapeGroup {
  int code = -1;
  int64_t n = 0;            // elements of this shape
  DBuf<int32_t> ids;        // global element id of group member k (empty when the mesh is uniform: id = k)
  // SoA geometry (efv_geometry_build)
  DBuf<double> Jinv, S, vol;
  DBuf<double> St;          // per face  J^-T s  (NF*d components): all the heat operator needs
  DBuf<double> Frec;        // AoS copy for the block assembly kernels: per (element, face) one record [J^-1 row-major | s] of
                            // d*d + d doubles (96 B in 3-D = three full sectors), built on first use
  DBuf<double> Srec;        // simplices only (J^-1 is the same on every face): per element [s of every face | J^-1], padded
                            // to an even number of doubles (tetrahedron 28, triangle 10) -- 2.6x smaller than Frec
};

struct Boundaries {
  int nB = 0;
  int64_t nF = 0, nOuter = 0;
  std::vector<int64_t> boundary_ptr;       // host copy, facets of boundary b
  std::vector<int64_t> outer_ptr_host;     // outer faces of boundary b: [outer_ptr[b], outer_ptr[b+1])
  DBuf<int64_t> facet_ptr;                 // nF+1
  DBuf<int32_t> facet_conn;
  DBuf<int64_t> facet_element;             // nF
  DBuf<int32_t> facet_local;               // nF
  DBuf<double> facet_area;                 // nF x 3 (outward)
  // outer faces = one per facet vertex, in facet order (Facet.py:41-45)
  DBuf<int32_t> outer_vertex;              // nOuter
  DBuf<int32_t> outer_local;               // position of the vertex in the element's facet table (Boundary.py:66)
  DBuf<int64_t> outer_facet;               // nOuter
  DBuf<double> outer_area_norm;            // nOuter
  // compact boundary-vertex -> outer-face map (ascending outer id = reference visiting order)
  int64_t nbv = 0;
  DBuf<int32_t> bv;                        // nbv vertex ids (ascending)
  DBuf<int64_t> bvo_ptr, bvo;              // nbv+1, nOuter
  DBuf<int32_t> vtx_to_bv;                 // nV: compact index of a boundary vertex, -1 otherwise
};

}  // namespace efv

struct efv_mesh {
  int dim = 0;
  int64_t nV = 0, nE = 0, conn_size = 0, n_inner_faces = 0;
  int n_regions = 0;
  int uniform_code = -1;     // shape code when every element has the same shape, else -1
  int max_m = 0;
  int max_incidence = 0;     // largest number of elements around a vertex
  efv::DBuf<double> coords;
  efv::DBuf<int64_t> elem_ptr;
  efv::DBuf<int32_t> conn;
  efv::DBuf<uint8_t> elem_shape;
  efv::DBuf<int32_t> elem_gidx;
  efv::DBuf<int32_t> region;
  efv::DBuf<int64_t> inc_ptr;
  efv::DBuf<uint32_t> inc;
  efv::DBuf<int64_t> slot_ptr;   // per element offset into slot-map-sized arrays (sum of m*m); empty if uniform
  efv::DBuf<double> vertex_volume;
  // pair schedule of the lane-per-row assembly (assemble.cu: k_heat_lanes), built on first use: tiles of 32 consecutive
  // vertices; a tile's (element, local vertex) pairs laid out [step][lane] so that a step holds, for every lane (= row),
  // the next pair of ONE (shape, local vertex) key (0xffffffff: none) -- coalesced, warp-uniform key, order per row
  // ascending (local vertex, shape, element)
  efv::DBuf<uint32_t> sched_pk;     // [steps][32]  local vertex | last step of the tile | [shape] | element (assemble.cu: kPd*)
  efv::DBuf<int64_t> sched_off;     // [tiles + 1] first step of a tile
  int64_t sched_steps = 0;
  efv::DBuf<uint32_t> sched8_pk;    // the same with tiles of 8 vertices, [steps][8] (block assembly: 4 lanes per row)
  efv::DBuf<int64_t> sched8_off;
  int64_t sched8_steps = 0;
  efv::ShapeGroup groups[efv::kNumShapes];
  efv::Boundaries bnd;
  bool has_geometry = false;
};

namespace efv {
struct KrylovWork;
struct AmgHierarchy;
struct BlockPrec;
// owned/ghost exchange lists of a vertex-block partition
struct Halo {
  bool active = false;
  std::vector<int> nbr;                 // neighbour ranks
  std::vector<int64_t> send_ptr;        // per neighbour range into send_idx
  std::vector<int64_t> recv_ptr;        // per neighbour range into the ghost block (ghosts are grouped by owner)
  DBuf<int32_t> send_idx;               // local ids of owned vertices to send
  std::vector<int32_t> send_idx_host;   // host copy (coarse-level halos of the multigrid hierarchy are derived from it)
  DBuf<double> send_buf;
  int64_t n_owned = 0, n_ghost = 0;     // vector layout: [owned | ghosts grouped by owner]
  int cap_ndof = 1;                     // widest exchange (values per vertex) the buffers are sized for
  // peer-memory path (CUDA IPC): own staging buffer [flags | ghost values x 2 parities] and the neighbours' mapped buffers
  bool peer = false;
  char *stage = nullptr;
  unsigned *ticket = nullptr;           // block counter of the push kernel (last block raises the flags)
  std::vector<char *> remote;
  std::vector<int64_t> remote_off, remote_ghosts;
  std::vector<int> remote_slot;
  uint32_t seq = 0;
};
}

struct efv_system {
  efv_mesh *mesh = nullptr;
  int ndof = 1;
  int64_t nV = 0, rows = 0, nnzg = 0, slot_size = 0;
  int64_t n_owned = 0;                 // leading vertices owned by this rank (== nV on one GPU); the rest are ghosts
  efv::Halo halo;
  int max_row_len = 0;
  efv::DBuf<int64_t> gptr;
  efv::DBuf<int32_t> gcol;
  efv::DBuf<uint8_t> slotpos;
  efv::DBuf<int32_t> diag_slot;        // graph slot of (v, v)
  efv::DBuf<double> vals;              // nnzg * ndof * ndof
  // vectors (vertex-major interleaved: [v][dof])
  efv::DBuf<double> x, b, xold, static_rhs, elim;
  // heat lumped vectors; Biot reuses acc for S/dt * vol
  efv::DBuf<double> gen, acc;
  // Biot: copy of the pre-Dirichlet p-row operator needed for the RHS (p rows x all cols)
  efv::DBuf<double> rhs_op;
  // Dirichlet set
  efv::DBuf<uint8_t> dir_flag;
  efv::DBuf<uint8_t> row_mark;         // ndof == 1: row has a Dirichlet DOF among its columns (itself included)
  bool marks_valid = false;
  efv::DBuf<double> dir_value;
  efv::DBuf<int32_t> dir_last;
  int dirichlet_mode = -1;
  uint64_t vals_version = 1;           // bumped whenever `vals` changes (the solver's SELL copy is rebuilt lazily)
  int fuse_mode = -1;                  // Dirichlet mode applied inside the next assembly kernel (-1: none)
  bool has_dirichlet = false;
  int dir_counter = 0;                 // running index of the Dirichlet visits since the last bc_clear (last visit wins)
  int physics = 0;
  // Krylov workspace
  efv::KrylovWork *kw = nullptr;
  efv::AmgHierarchy *amg = nullptr;    // aggregation multigrid hierarchy (built at the first preconditioned solve)
  efv::BlockPrec *bprec = nullptr;     // coupled Biot systems: multigrid on the displacement and pressure blocks (BiCGStab)
  int precond = 0;                     // 0: Jacobi, 1: aggregation multigrid (efv_system_set_preconditioner)
  int lattice[3] = {0, 0, 0};          // owned vertices are numbered i + nx (j + ny k) on an nx x ny x nz lattice (0: unknown)
  double t_assemble_ms = 0, t_solve_ms = 0, t_spmv_ms = 0;
  efv::DBuf<double> props_dev;
  // asynchronous field download: XOLD -> pinned staging buffer on a copy stream, overlapped with the next step
  double *dl_host = nullptr;
  cudaStream_t dl_stream = nullptr;
  cudaEvent_t dl_ready = nullptr, dl_done = nullptr;
  bool dl_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // assembly timing (read lazily: no host sync in the hot path)
  bool assemble_timed = false;
};

namespace efv {
// mesh.cu
void mesh_finalize(efv_mesh *m, const int64_t *elem_ptr_host, int uniform_m);
void mesh_set_boundaries(efv_mesh *m, int nB, const int64_t *boundary_ptr, int64_t nF, const int64_t *facet_ptr,
                         const int32_t *facet_conn);
// geometry.cu
void geometry_build(efv_mesh *m);
void geometry_export_group(const efv_mesh *m, int code, double *G, double *Jinv, double *S, double *vol);
void vertex_volume_export(const efv_mesh *m, double *vv);
// csr.cu
void system_build_graph(efv_system *s);
// assemble.cu
void heat_assemble(efv_system *s, const double *props_host, double dt, int transient);
void elasticity_assemble(efv_system *s, const double *props_host, int gravity);
void biot_assemble(efv_system *s, const double *props_host, double dt, const double *gvec);
void bc_clear(efv_system *s);
void bc_dirichlet(efv_system *s, const int64_t *rows, const double *values, int64_t n);
void bc_neumann(efv_system *s, int boundary, int dof, double value, double sign, const double *values = nullptr);
void apply_dirichlet(efv_system *s, int mode);
void build_rhs(efv_system *s, int transient);
void ensure_vectors(efv_system *s);
void export_csr(const efv_system *s, int64_t *indptr, int32_t *indices);
void export_slotmap(const efv_system *s, int64_t *slots);
void export_values(const efv_system *s, double *data);
void fs_mass_rhs(efv_system *biot, efv_system *mass);
void fs_geo_rhs(efv_system *biot, efv_system *mass, efv_system *geo);
double fs_update(efv_system *biot, efv_system *mass, efv_system *geo);
void vec_upload(efv_system *s, DBuf<double> &dst, const double *host);
void vec_download(const efv_system *s, const DBuf<double> &src, double *host);
// krylov.cu
void krylov_free(efv_system *s);
void solve_pcg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres);
void solve_bicgstab(efv_system *s, double rtol, int maxit, int *iters, double *relres);
void spmv(efv_system *s, const double *x, double *y);
void solve_pcg_amg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres);
void amg_free(AmgHierarchy *H);
void block_prec_free(BlockPrec *b);
void block_prec_prepare(efv_system *s);
void block_prec_apply(efv_system *s, const double *w, const double *dinv, double *z, const int *done);
void amg_level_export(efv_system *s, int level, int64_t *gptr, int32_t *gcol, double *vals, int32_t *agg);
int amg_info(efv_system *s, int64_t *rows, int64_t *nnz, double *lmax, int cap, double *numeric_ms);
double max_abs_diff(efv_system *s);
double true_relres(efv_system *s, int negate);
void vec_sums(efv_system *s, const double *vec, double *sum, double *sumsq);
double symmetry_defect(efv_system *s);
// comm.cu
void comm_unique_id(const char *nccl_path, char *out128);
void comm_init(const char *nccl_path, int rank, int nranks, const char *id128);
void comm_destroy();
bool comm_active();
int comm_rank();
int comm_size();
void comm_allreduce_sum(double *dev, int count);
void comm_allreduce_max(double *dev, int count);
void comm_peer_export(char *out64);
void comm_peer_import(const char *handles);
bool comm_peer_active();
void comm_allreduce_partials(const double *partial, int G, int narr, double *out, int op);
void comm_allreduce_partials_inplace(double *partial, int G, int narr);
int comm_peer_error();
void halo_stage_export(Halo &h, char *out64);
void halo_stage_import(Halo &h, const char *handles, const int64_t *remote_off, const int *remote_slot, const int64_t *remote_ghosts);
void halo_release(Halo &h);
void halo_exchange(Halo &h, double *vec, int ndof);
// ghosts read from the staging buffer by the consumer kernel itself (krylov.cu: k_spmv_sell_tma)
struct HaloIn {
  const double *ghost = nullptr;        // staging values of the current exchange (nullptr: ghosts are inside the vector)
  const uint32_t *flags = nullptr;      // one per neighbour
  int n_nbr = 0;
  uint32_t seq = 0;
  int64_t n_owned = 0;
  int *err = nullptr;
};
bool halo_push(Halo &h, const double *vec, int ndof);
HaloIn halo_in(const Halo &h, int ndof);
void halo_wait_copy(Halo &h, double *vec, int ndof);
void comm_peer_release();
void comm_mail_exchange(int npeers, const int *peers, const char *send, char *recv);
void comm_allgather_i64(int64_t mine, int64_t *all);
struct GatherStage;
GatherStage *gather_stage_create(size_t cap_doubles);
void gather_stage_destroy(GatherStage *g);
void gather_all(GatherStage *g, int64_t offset, int64_t count, int64_t total, const double *src, double *dst, const int *done);
void halo_connect_peer(Halo &h);
void halo_set(efv_system *s, int n_nbr, const int *nbr, const int64_t *send_ptr, const int32_t *send_idx, const int64_t *recv_ptr);
void halo_exchange(efv_system *s, double *vec);
}  // namespace efv
