// api.cu -- the extern "C" boundary declared in include/efv_b200.h.
#include "../../include/efv_b200.h"
#include "views.cuh"
#include <cstring>

namespace efv {
const std::string &last_error();
void set_device(int device);
int64_t launch_count();
void field_copy_counters(int64_t *out4);
}  // namespace efv

#define EFV_API_BEGIN try {
#define EFV_API_END                                   \
  return 0;                                           \
  }                                                   \
  catch (const std::exception &ex) {                  \
    efv::set_last_error(ex.what());                   \
    return -1;                                        \
  }
#define EFV_API_END_PTR(ptr_to_free)                  \
  }                                                   \
  catch (const std::exception &ex) {                  \
    efv::set_last_error(ex.what());                   \
    delete ptr_to_free;                               \
    return nullptr;                                   \
  }

extern "C" {

const char *efv_last_error(void) { return efv::last_error().c_str(); }

int efv_set_device(int device) {
  EFV_API_BEGIN
  efv::set_device(device);
  EFV_API_END
}
int efv_device_count(int *count) {
  EFV_API_BEGIN
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { c = 0; cudaGetLastError(); }
  *count = c;
  EFV_API_END
}
void *efv_get_stream(void) {
  try { return (void *)efv::stream(); } catch (const std::exception &ex) { efv::set_last_error(ex.what()); return nullptr; }
}
int efv_synchronize(void) {
  EFV_API_BEGIN
  EFV_CUDA(cudaStreamSynchronize(efv::stream()));
  EFV_API_END
}
int64_t efv_kernel_launch_count(void) { return efv::launch_count(); }
int efv_field_copy_counters(int64_t *uploads, int64_t *upload_bytes, int64_t *downloads, int64_t *download_bytes) {
  EFV_API_BEGIN
  int64_t c[4];
  efv::field_copy_counters(c);
  if (uploads) *uploads = c[0];
  if (upload_bytes) *upload_bytes = c[1];
  if (downloads) *downloads = c[2];
  if (download_bytes) *download_bytes = c[3];
  EFV_API_END
}

// ---- mesh ----------------------------------------------------------------------------------------------------------
static efv_mesh *mesh_create_impl(int dim, int64_t nV, const double *coords, int64_t nE, const int64_t *elem_ptr, int uniform_m,
                                  const int32_t *conn, const int32_t *region_of_elem, int n_regions) {
  efv_mesh *m = nullptr;
  try {
    EFV_REQUIRE(dim == 2 || dim == 3, "dimension must be 2 or 3");
    EFV_REQUIRE(nV > 0 && nE > 0, "empty mesh");
    EFV_REQUIRE(nV < (int64_t(1) << 31), "vertex ids must fit int32");
    EFV_REQUIRE(coords && conn && (elem_ptr || uniform_m > 0), "null mesh array");
    m = new efv_mesh();
    m->dim = dim; m->nV = nV; m->nE = nE; m->n_regions = n_regions > 0 ? n_regions : 1;
    m->conn_size = elem_ptr ? elem_ptr[nE] : nE * (int64_t)uniform_m;
    m->coords.upload(coords, (size_t)nV * 3);
    m->conn.upload(conn, (size_t)m->conn_size);
    if (region_of_elem && m->n_regions > 1) m->region.upload(region_of_elem, (size_t)nE);
    else { m->region.alloc(nE); m->region.zero(); }
    efv::mesh_finalize(m, elem_ptr, uniform_m);
    return m;
  EFV_API_END_PTR(m)
}
efv_mesh *efv_mesh_create(int dim, int64_t nV, const double *coords, int64_t nE, const int64_t *elem_ptr,
                          const int32_t *conn, const int32_t *region_of_elem, int n_regions) {
  if (!elem_ptr) { efv::set_last_error("null elem_ptr (use efv_mesh_create_uniform)"); return nullptr; }
  return mesh_create_impl(dim, nV, coords, nE, elem_ptr, 0, conn, region_of_elem, n_regions);
}
efv_mesh *efv_mesh_create_uniform(int dim, int64_t nV, const double *coords, int64_t nE, int vertices_per_element,
                                  const int32_t *conn, const int32_t *region_of_elem, int n_regions) {
  return mesh_create_impl(dim, nV, coords, nE, nullptr, vertices_per_element, conn, region_of_elem, n_regions);
}

int efv_mesh_set_boundaries(efv_mesh *m, int n_boundaries, const int64_t *boundary_ptr, int64_t n_facets,
                            const int64_t *facet_ptr, const int32_t *facet_conn) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  efv::mesh_set_boundaries(m, n_boundaries, boundary_ptr, n_facets, facet_ptr, facet_conn);
  EFV_API_END
}
void efv_mesh_destroy(efv_mesh *m) { delete m; }
int64_t efv_mesh_num_vertices(const efv_mesh *m) { return m ? m->nV : -1; }
int64_t efv_mesh_num_elements(const efv_mesh *m) { return m ? m->nE : -1; }
int64_t efv_mesh_num_inner_faces(const efv_mesh *m) { return m ? m->n_inner_faces : -1; }
int64_t efv_mesh_num_facets(const efv_mesh *m) { return m ? m->bnd.nF : -1; }
int64_t efv_mesh_num_outer_faces(const efv_mesh *m) { return m ? m->bnd.nOuter : -1; }
int64_t efv_mesh_group_size(const efv_mesh *m, int shape_code) {
  return (m && shape_code >= 0 && shape_code < efv::kNumShapes) ? m->groups[shape_code].n : -1;
}

int efv_geometry_build(efv_mesh *m) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  efv::geometry_build(m);
  EFV_CUDA(cudaStreamSynchronize(efv::stream()));
  EFV_API_END
}
int efv_geometry_release(efv_mesh *m) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  for (int c = 0; c < efv::kNumShapes; ++c) { m->groups[c].Jinv.release(); m->groups[c].S.release(); m->groups[c].vol.release(); m->groups[c].St.release(); m->groups[c].Frec.release(); }
  m->vertex_volume.release();
  m->has_geometry = false;
  EFV_API_END
}
int efv_geometry_export_group(const efv_mesh *m, int shape_code, double *G, double *Jinv, double *S, double *vol) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  efv::geometry_export_group(m, shape_code, G, Jinv, S, vol);
  EFV_API_END
}
int efv_vertex_volume_export(const efv_mesh *m, double *vertex_volume) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  efv::vertex_volume_export(m, vertex_volume);
  EFV_API_END
}
int efv_boundary_export(const efv_mesh *m, int64_t *facet_element, int32_t *facet_local, double *facet_area,
                        int32_t *outer_vertex, double *outer_area_norm) {
  EFV_API_BEGIN
  EFV_REQUIRE(m, "null mesh");
  const efv::Boundaries &B = m->bnd;
  if (facet_element) B.facet_element.download(facet_element, B.nF);
  if (facet_local) B.facet_local.download(facet_local, B.nF);
  if (facet_area) B.facet_area.download(facet_area, B.nF * 3);
  if (outer_vertex) B.outer_vertex.download(outer_vertex, B.nOuter);
  if (outer_area_norm) B.outer_area_norm.download(outer_area_norm, B.nOuter);
  EFV_API_END
}

// ---- system --------------------------------------------------------------------------------------------------------
efv_system *efv_system_create(efv_mesh *m, int ndof) {
  efv_system *s = nullptr;
  try {
    EFV_REQUIRE(m, "null mesh");
    EFV_REQUIRE(ndof >= 1 && ndof <= 4, "ndof must be 1..4");
    s = new efv_system();
    s->mesh = m;
    s->ndof = ndof;
    efv::system_build_graph(s);
    s->vals.alloc((size_t)s->nnzg * ndof * ndof);
    s->vals.zero();
    efv::ensure_vectors(s);
    efv::bc_clear(s);
    EFV_CUDA(cudaStreamSynchronize(efv::stream()));
    return s;
  EFV_API_END_PTR(s)
}
// a system on a caller-supplied CSR pattern (no mesh): the LinearSystemCSR seam (LinearSystem.py:113-183) for matrices
// that were filled on the host
__global__ void k_find_diag(int64_t n, const int64_t *__restrict__ ptr, const int32_t *__restrict__ col,
                            int32_t *__restrict__ diag, int *__restrict__ err) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    int found = -1;
    for (int64_t t = ptr[v]; t < ptr[v + 1]; ++t)
      if (col[t] == v) found = (int)t;
    if (found < 0) atomicExch(err, 1);
    diag[v] = found < 0 ? 0 : found;
  }
}
efv_system *efv_system_create_from_csr(int64_t n, const int64_t *indptr, const int32_t *indices) {
  efv_system *s = nullptr;
  try {
    EFV_REQUIRE(n > 0 && indptr && indices, "bad CSR arguments");
    EFV_REQUIRE(indptr[n] < (int64_t(1) << 31), "nnz exceeds int32 slot range");
    for (int64_t k = 0; k < indptr[n]; ++k) EFV_REQUIRE(indices[k] >= 0 && indices[k] < n, "column index out of range");
    s = new efv_system();
    s->mesh = nullptr; s->ndof = 1; s->nV = s->n_owned = s->rows = n; s->nnzg = indptr[n];
    s->gptr.upload(indptr, n + 1);
    s->gcol.upload(indices, s->nnzg);
    s->diag_slot.alloc(n);
    efv::DBuf<int> err(1);
    err.zero();
    EFV_LAUNCH(k_find_diag, efv::grid_for(n, 256), 256, 0, n, s->gptr.p, s->gcol.p, s->diag_slot.p, err.p);
    int herr = 0;
    err.download(&herr, 1);
    EFV_REQUIRE(herr == 0, "every row of the pattern must contain its diagonal entry");
    s->vals.alloc(s->nnzg);
    s->vals.zero();
    efv::ensure_vectors(s);
    efv::bc_clear(s);
    s->physics = 1;      // generic: RHS = static part (+ eliminated columns), Dirichlet rows <- g
    EFV_CUDA(cudaStreamSynchronize(efv::stream()));
    return s;
  EFV_API_END_PTR(s)
}
int efv_system_set_values(efv_system *s, const double *data) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && data, "null argument");
  EFV_REQUIRE(s->ndof == 1, "efv_system_set_values expects a scalar (ndof = 1) system");
  s->vals.upload(data, s->nnzg);
  s->vals_version++;
  EFV_CUDA(cudaStreamSynchronize(efv::stream()));
  s->dirichlet_mode = -1;
  EFV_API_END
}
void efv_system_destroy(efv_system *s) {
  if (!s) return;
  if (s->amg) efv::amg_free(s->amg);
  if (s->bprec) efv::block_prec_free(s->bprec);
  efv::krylov_free(s);
  efv::halo_release(s->halo);
  if (s->dl_stream) { cudaStreamSynchronize(s->dl_stream); cudaStreamDestroy(s->dl_stream); cudaEventDestroy(s->dl_ready); cudaEventDestroy(s->dl_done); }
  if (s->dl_host) cudaFreeHost(s->dl_host);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
}
int64_t efv_system_rows(const efv_system *s) { return s ? s->rows : -1; }
int64_t efv_system_nnz(const efv_system *s) { return s ? s->nnzg * s->ndof * s->ndof : -1; }
int64_t efv_system_graph_nnz(const efv_system *s) { return s ? s->nnzg : -1; }
int64_t efv_system_slotmap_size(const efv_system *s) { return s ? s->slot_size : -1; }
int efv_system_max_row_length(const efv_system *s) { return s ? s->max_row_len : -1; }

int efv_system_export_csr(const efv_system *s, int64_t *indptr, int32_t *indices) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::export_csr(s, indptr, indices);
  EFV_API_END
}
int efv_system_export_graph(const efv_system *s, int64_t *gptr, int32_t *gcol) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  if (gptr) s->gptr.download(gptr, s->nV + 1);
  if (gcol) s->gcol.download(gcol, s->nnzg);
  EFV_API_END
}
int efv_system_export_slotmap(const efv_system *s, int64_t *slots) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  efv::export_slotmap(s, slots);
  EFV_API_END
}
int efv_system_export_values(const efv_system *s, double *data) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::export_values(s, data);
  EFV_API_END
}

static efv::DBuf<double> *pick_vec(efv_system *s, int which) {
  switch (which) {
    case EFV_VEC_X: return &s->x;
    case EFV_VEC_B: return &s->b;
    case EFV_VEC_XOLD: return &s->xold;
    default: throw efv::Error("unknown vector id");
  }
}
int efv_vec_upload(efv_system *s, int which, const double *host) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && host, "null argument");
  efv::vec_upload(s, *pick_vec(s, which), host);
  EFV_API_END
}
int efv_vec_download(const efv_system *s, int which, double *host) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && host, "null argument");
  efv::vec_download(s, *pick_vec(const_cast<efv_system *>(s), which), host);
  EFV_API_END
}

// ---- assembly / BCs ------------------------------------------------------------------------------------------------
int efv_heat_assemble(efv_system *s, const double *props, double dt, int transient) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && props, "null argument");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  EFV_REQUIRE(!transient || dt > 0, "transient assembly needs a positive time step");
  efv::heat_assemble(s, props, dt, transient);
  EFV_API_END
}
int efv_elasticity_assemble(efv_system *s, const double *props, int gravity) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && props, "null argument");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  efv::elasticity_assemble(s, props, gravity);
  EFV_API_END
}
int efv_biot_assemble(efv_system *s, const double *props, double dt, const double *gvec) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && props, "null argument");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  EFV_REQUIRE(dt > 0, "Biot assembly needs a positive time step");
  efv::biot_assemble(s, props, dt, gvec);
  EFV_API_END
}
int efv_bc_clear(efv_system *s) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::bc_clear(s);
  EFV_API_END
}
int efv_bc_dirichlet(efv_system *s, const int64_t *rows, const double *values, int64_t n) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(n == 0 || (rows && values), "null Dirichlet arrays");
  for (int64_t k = 0; k < n; ++k) EFV_REQUIRE(rows[k] >= 0 && rows[k] < s->rows, "Dirichlet row out of range");
  efv::bc_dirichlet(s, rows, values, n);
  EFV_API_END
}
int efv_bc_neumann(efv_system *s, int boundary, int dof, double value, double sign) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  efv::bc_neumann(s, boundary, dof, value, sign);
  EFV_API_END
}
int64_t efv_boundary_outer_faces(const efv_mesh *m, int boundary, int64_t *first) {
  if (!m || boundary < 0 || boundary >= m->bnd.nB) return -1;
  if (first) *first = m->bnd.outer_ptr_host[boundary];
  return m->bnd.outer_ptr_host[boundary + 1] - m->bnd.outer_ptr_host[boundary];
}
int efv_bc_neumann_values(efv_system *s, int boundary, int dof, const double *values, int64_t n, double sign) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && values, "null argument");
  EFV_REQUIRE(s->mesh, "this system has no mesh (created from a CSR pattern)");
  EFV_REQUIRE(boundary >= 0 && boundary < s->mesh->bnd.nB, "bad boundary index");
  EFV_REQUIRE(n == s->mesh->bnd.outer_ptr_host[boundary + 1] - s->mesh->bnd.outer_ptr_host[boundary], "one value per outer face of the boundary");
  efv::bc_neumann(s, boundary, dof, 0.0, sign, values);
  EFV_API_END
}
int efv_assemble_dirichlet_mode(efv_system *s, int mode) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(mode >= -1 && mode <= 1, "bad Dirichlet mode");
  s->fuse_mode = mode;
  EFV_API_END
}
int efv_apply_dirichlet(efv_system *s, int mode) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::apply_dirichlet(s, mode);
  EFV_API_END
}
int efv_build_rhs(efv_system *s, int transient) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::build_rhs(s, transient);
  EFV_API_END
}

// ---- fixed-stress split ------------------------------------------------------------------------------------------
int efv_fs_mass_rhs(efv_system *biot, efv_system *mass) {
  EFV_API_BEGIN
  efv::fs_mass_rhs(biot, mass);
  EFV_API_END
}
int efv_fs_geo_rhs(efv_system *biot, efv_system *mass, efv_system *geo) {
  EFV_API_BEGIN
  efv::fs_geo_rhs(biot, mass, geo);
  EFV_API_END
}
int efv_fs_update(efv_system *biot, efv_system *mass, efv_system *geo, double *difference) {
  EFV_API_BEGIN
  EFV_REQUIRE(difference, "null argument");
  *difference = efv::fs_update(biot, mass, geo);
  EFV_API_END
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------
int efv_comm_unique_id(const char *nccl_library_path, char *out128) {
  EFV_API_BEGIN
  EFV_REQUIRE(out128, "null argument");
  efv::comm_unique_id(nccl_library_path, out128);
  EFV_API_END
}
int efv_comm_init(const char *nccl_library_path, int rank, int nranks, const char *id128) {
  EFV_API_BEGIN
  EFV_REQUIRE(nranks >= 1 && nranks <= 1024 && rank >= 0 && rank < nranks, "bad communicator arguments");
  efv::comm_init(nccl_library_path, rank, nranks, id128);
  EFV_API_END
}
int efv_comm_destroy(void) {
  EFV_API_BEGIN
  efv::comm_destroy();
  EFV_API_END
}
int efv_comm_peer_export(char *out64) {
  EFV_API_BEGIN
  EFV_REQUIRE(out64, "null argument");
  efv::comm_peer_export(out64);
  EFV_API_END
}
int efv_comm_peer_import(const char *handles) {
  EFV_API_BEGIN
  EFV_REQUIRE(handles, "null argument");
  efv::comm_peer_import(handles);
  EFV_API_END
}
int efv_comm_peer_release(void) {
  EFV_API_BEGIN
  efv::comm_peer_release();
  EFV_API_END
}
int efv_halo_stage_export(efv_system *s, char *out64) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && out64, "null argument");
  efv::halo_stage_export(s->halo, out64);
  EFV_API_END
}
int efv_halo_stage_import(efv_system *s, const char *handles, const int64_t *remote_offsets, const int *remote_slots,
                          const int64_t *remote_ghosts) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && handles && remote_offsets && remote_slots && remote_ghosts, "null argument");
  efv::halo_stage_import(s->halo, handles, remote_offsets, remote_slots, remote_ghosts);
  EFV_API_END
}
int efv_system_set_owned(efv_system *s, int64_t n_owned) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(n_owned >= 0 && n_owned <= s->nV, "owned count out of range");
  s->n_owned = n_owned;
  EFV_API_END
}
int efv_system_set_halo(efv_system *s, int n_neighbours, const int *neighbour_ranks, const int64_t *send_ptr,
                        const int32_t *send_idx, const int64_t *recv_ptr) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && (n_neighbours == 0 || (neighbour_ranks && send_ptr && recv_ptr)), "null argument");
  efv::halo_set(s, n_neighbours, neighbour_ranks, send_ptr, send_idx, recv_ptr);
  EFV_API_END
}
int efv_halo_exchange(efv_system *s, int which) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::halo_exchange(s, pick_vec(s, which)->p);
  EFV_API_END
}

// ---- solvers -------------------------------------------------------------------------------------------------------
int efv_solve_pcg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  if (s->precond == 1) efv::solve_pcg_amg(s, rtol, maxit, negate, iters, relres);
  else efv::solve_pcg(s, rtol, maxit, negate, iters, relres);
  EFV_API_END
}
int efv_system_set_preconditioner(efv_system *s, int kind) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(kind == 0 || kind == 1, "preconditioner: 0 = Jacobi, 1 = aggregation multigrid");
  s->precond = kind;
  EFV_API_END
}
int efv_system_set_lattice(efv_system *s, int nx, int ny, int nz) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(nx > 0 && ny > 0 && nz > 0 && (int64_t)nx * ny * nz == s->n_owned, "lattice extents must multiply to the owned vertex count");
  EFV_REQUIRE(!s->amg, "set the lattice before the first preconditioned solve");
  s->lattice[0] = nx; s->lattice[1] = ny; s->lattice[2] = nz;
  EFV_API_END
}
int efv_amg_level_export(efv_system *s, int level, int64_t *gptr, int32_t *gcol, double *vals, int32_t *aggregates) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::amg_level_export(s, level, gptr, gcol, vals, aggregates);
  EFV_API_END
}
int efv_amg_info(efv_system *s, int capacity, int *levels, int64_t *rows, int64_t *nnz, double *lmax, double *numeric_ms) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && levels && rows && nnz && lmax, "null argument");
  *levels = efv::amg_info(s, rows, nnz, lmax, capacity, numeric_ms);
  EFV_API_END
}
int efv_solve_bicgstab(efv_system *s, double rtol, int maxit, int *iters, double *relres) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::solve_bicgstab(s, rtol, maxit, iters, relres);
  EFV_API_END
}
int efv_spmv(efv_system *s, int which_in, int which_out) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(which_in != which_out, "in-place SpMV is not supported");
  efv::ensure_vectors(s);
  efv::spmv(s, pick_vec(s, which_in)->p, pick_vec(s, which_out)->p);
  EFV_API_END
}
int efv_true_residual(efv_system *s, int negate, double *relres) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && relres, "null argument");
  *relres = efv::true_relres(s, negate);
  EFV_API_END
}
int efv_vec_sums(efv_system *s, int which, double *sum, double *sum_of_squares) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null argument");
  efv::ensure_vectors(s);
  efv::vec_sums(s, pick_vec(s, which)->p, sum, sum_of_squares);
  EFV_API_END
}
int efv_symmetry_defect(efv_system *s, double *defect) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && defect, "null argument");
  *defect = efv::symmetry_defect(s);
  EFV_API_END
}
int efv_max_abs_diff(efv_system *s, double *out) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && out, "null argument");
  *out = efv::max_abs_diff(s);
  EFV_API_END
}
// XOLD is constant from efv_advance until the next efv_advance, so its download can overlap the next time step: the copy
// runs on its own stream behind an event of the library stream, into a pinned buffer owned by the system
int efv_download_xold_begin(efv_system *s) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  EFV_REQUIRE(!s->dl_pending, "a field download is already in flight");
  efv::ensure_vectors(s);
  if (!s->dl_stream) {
    EFV_CUDA(cudaStreamCreateWithFlags(&s->dl_stream, cudaStreamNonBlocking));
    EFV_CUDA(cudaEventCreateWithFlags(&s->dl_ready, cudaEventDisableTiming));
    EFV_CUDA(cudaEventCreateWithFlags(&s->dl_done, cudaEventDisableTiming));
    EFV_CUDA(cudaMallocHost((void **)&s->dl_host, s->rows * sizeof(double)));
  }
  EFV_CUDA(cudaEventRecord(s->dl_ready, efv::stream()));
  EFV_CUDA(cudaStreamWaitEvent(s->dl_stream, s->dl_ready, 0));
  EFV_CUDA(cudaMemcpyAsync(s->dl_host, s->xold.p, s->rows * sizeof(double), cudaMemcpyDeviceToHost, s->dl_stream));
  EFV_CUDA(cudaEventRecord(s->dl_done, s->dl_stream));
  s->dl_pending = true;
  efv::count_field_copy(1, s->rows * sizeof(double));
  EFV_API_END
}
int efv_download_xold_end(efv_system *s, double *host) {
  EFV_API_BEGIN
  EFV_REQUIRE(s && host, "null argument");
  EFV_REQUIRE(s->dl_pending, "no field download in flight");
  EFV_CUDA(cudaEventSynchronize(s->dl_done));
  s->dl_pending = false;
  const int64_t nV = s->nV;
  if (s->ndof == 1) memcpy(host, s->dl_host, s->rows * sizeof(double));
  else
    for (int64_t v = 0; v < nV; ++v)
      for (int f = 0; f < s->ndof; ++f) host[(int64_t)f * nV + v] = s->dl_host[v * s->ndof + f];   // field-major at the ABI
  EFV_API_END
}
int efv_advance(efv_system *s) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  efv::ensure_vectors(s);
  if (s->dl_pending) EFV_CUDA(cudaStreamWaitEvent(efv::stream(), s->dl_done, 0));     // XOLD is still being read by the copy stream
  EFV_CUDA(cudaMemcpyAsync(s->xold.p, s->x.p, s->rows * sizeof(double), cudaMemcpyDeviceToDevice, efv::stream()));
  EFV_API_END
}
int efv_last_timings(efv_system *s, double *assemble_ms, double *solve_ms) {
  EFV_API_BEGIN
  EFV_REQUIRE(s, "null system");
  if (s->assemble_timed) {
    EFV_CUDA(cudaEventSynchronize(s->ev1));
    float ms = 0;
    EFV_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->t_assemble_ms = ms;
    s->assemble_timed = false;
  }
  if (assemble_ms) *assemble_ms = s->t_assemble_ms;
  if (solve_ms) *solve_ms = s->t_solve_ms;
  EFV_API_END
}

}  // extern "C"
