/*
 * efv_b200.h -- C-ABI of the B200-native EbFVM hot path (assembly + sparse solve) of PyEFVLib.
 *
 * The reference (GustavoExel/PyEFVLib) is pure Python and has no FFI; the seams this ABI slots into are cited
 * per entry point as  [ref: file:line]  (paths relative to the reference root).  All entry points are
 * extern "C", take plain pointers and sizes, and return 0 on success / non-zero on failure with a message
 * available from efv_last_error().  HOST pointers are borrowed for the duration of the call only; device
 * memory is owned by the opaque handles.  One process drives one GPU (efv_set_device); handles are not
 * thread-safe.  There is no CPU fallback: every entry point needs a CUDA device.
 *
 * DOF numbering at this boundary is the reference's FIELD-MAJOR one, row = vertex + field * numberOfVertices
 * [ref: apps/stress_equilibrium.py:77-79, apps/geomechanics.py:63, PyEFVLib/LinearSystem.py:119-124].
 */
#ifndef EFV_B200_H
#define EFV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct efv_mesh efv_mesh;       /* device mesh: vertices, elements, boundaries, incidence, geometry   */
typedef struct efv_system efv_system;   /* linear system on a mesh: CSR graph, slot map, values, vectors, Krylov */

/* shape codes [ref: PyEFVLib/Shape.py:7,35,63,98,141,187 ; Grid.py:41]; 3-D vertex counts 4/8/6/5 = tetrahedron/
 * hexahedron/prism/pyramid, 2-D 3/4 = triangle/quadrilateral */
enum { EFV_TRIANGLE = 0, EFV_QUADRILATERAL = 1, EFV_TETRAHEDRON = 2, EFV_HEXAHEDRON = 3, EFV_PRISM = 4, EFV_PYRAMID = 5 };
/* Dirichlet application modes */
enum { EFV_DIRICHLET_ROWS = 0,       /* reference form: row <- e_i^T, b_i = g  [ref: apps/heat_transfer.py:70-76,122-126] */
       EFV_DIRICHLET_SYMMETRIC = 1   /* + column elimination into the RHS (same solution, CG-ready)  */ };
/* vectors owned by a system (efv_vec_upload / efv_vec_download); all have rows = ndof * nV entries, field-major */
enum { EFV_VEC_X = 0, EFV_VEC_B = 1, EFV_VEC_XOLD = 2 };

/* ---- process-level -------------------------------------------------------------------------------------- */
const char *efv_last_error(void);
int efv_set_device(int device);
int efv_device_count(int *count);
void *efv_get_stream(void);            /* the cudaStream_t every kernel of this library is launched on */
int efv_synchronize(void);
int64_t efv_kernel_launch_count(void);
/* whole-field host<->device copies made through efv_vec_upload / efv_vec_download since the process started (a device-resident
 * time loop downloads a field only when a saver asks for it) */
int efv_field_copy_counters(int64_t *uploads, int64_t *upload_bytes, int64_t *downloads, int64_t *download_bytes); /* kernels launched by this library since load (bench: gpu_launches) */

/* ---- mesh  [ref: PyEFVLib/GridData.py:5-24 setters; Grid.py:18-65 build] ---------------------------------
 * coords: nV x 3 doubles; elements: CSR-like (elem_ptr[nE+1] into conn, vertex ids 0-based) in grid order;
 * the shape of an element follows from its vertex count and `dim` [ref: Element.py:22-27];
 * region_of_elem[nE] in [0, n_regions).                                                                     */
efv_mesh *efv_mesh_create(int dim, int64_t nV, const double *coords, int64_t nE, const int64_t *elem_ptr,
                          const int32_t *conn, const int32_t *region_of_elem, int n_regions);
/* same for a mesh whose elements all have `vertices_per_element` vertices (conn is nE x m row-major): no pointer array.
 * region_of_elem may be NULL when there is a single region.                                                        */
efv_mesh *efv_mesh_create_uniform(int dim, int64_t nV, const double *coords, int64_t nE, int vertices_per_element,
                                  const int32_t *conn, const int32_t *region_of_elem, int n_regions);
/* boundary facets [ref: GridData.py:18-21; Grid.py:51-65; Boundary.py:27-75; Facet.py:22-36]:
 * facets of all boundaries concatenated in boundary order; boundary b owns facets
 * [boundary_ptr[b], boundary_ptr[b+1]); facet f has vertices facet_conn[facet_ptr[f] .. facet_ptr[f+1]).
 * Matches every facet to its owning element + local facet, computes outward area vectors and outer faces. */
int efv_mesh_set_boundaries(efv_mesh *m, int n_boundaries, const int64_t *boundary_ptr, int64_t n_facets,
                            const int64_t *facet_ptr, const int32_t *facet_conn);
void efv_mesh_destroy(efv_mesh *m);
int64_t efv_mesh_num_vertices(const efv_mesh *m);
int64_t efv_mesh_num_elements(const efv_mesh *m);
int64_t efv_mesh_num_inner_faces(const efv_mesh *m);
int64_t efv_mesh_num_facets(const efv_mesh *m);
int64_t efv_mesh_num_outer_faces(const efv_mesh *m);                 /* = sum of facet vertex counts */
int64_t efv_mesh_group_size(const efv_mesh *m, int shape_code);      /* elements of one shape */
/* one-time geometry pass (north_star kernel 1): per inner face the global derivatives G = inv(D^T X) D^T
 * [ref: InnerFace.py:23-25, Element.py:69-71], Jacobian inverse, area vector [ref: Shape.py:29-33,57-61,84-96,162-185,208-247,
 * 119-139]; per element the sub-control volumes [ref: Element.py:39-48].  Stored on the device as SoA;
 * exported per shape group (elements of one shape in grid order), element by element, face by face:
 *   G    : per face d x m doubles (row-major),  Jinv : per face d x d,  S : per face 3,  vol : per element m,
 *   vertex_volume : nV  [ref: Vertex.py:9, Element.py:46].  Any output pointer may be NULL.                  */
int efv_geometry_build(efv_mesh *m);
int efv_geometry_release(efv_mesh *m);   /* frees the SoA (it is rebuilt on demand) */
int efv_geometry_export_group(const efv_mesh *m, int shape_code, double *G, double *Jinv, double *S, double *vol);
int efv_vertex_volume_export(const efv_mesh *m, double *vertex_volume);
/* boundary geometry: per facet owning element, local facet index, outward area vector (3);
 * per outer face (facet vertex) the vertex id and |area| [ref: OuterFace.py:18-19].                        */
int efv_boundary_export(const efv_mesh *m, int64_t *facet_element, int32_t *facet_local, double *facet_area,
                        int32_t *outer_vertex, double *outer_area_norm);

/* ---- linear system  [ref: PyEFVLib/LinearSystem.py:113-183 LinearSystemCSR] -------------------------------
 * CSR sparsity = union of element cliques with sorted columns (north_star kernel 2), DOF-expanded field-major
 * exactly like LinearSystemCSR.initialize [ref: LinearSystem.py:114-128] with sorted stencils.              */
efv_system *efv_system_create(efv_mesh *m, int ndof);
/* a scalar system on a caller-supplied CSR pattern (no mesh): fill it with efv_system_set_values, put the RHS into
 * EFV_VEC_B, solve.  Every row must contain its diagonal entry.  [ref: LinearSystem.py:113-183 used as in
 * apps/csr/heat_transfer_csr.py:57-133]                                                                       */
efv_system *efv_system_create_from_csr(int64_t n, const int64_t *indptr, const int32_t *indices);
int efv_system_set_values(efv_system *s, const double *data);
void efv_system_destroy(efv_system *s);
int64_t efv_system_rows(const efv_system *s);
int64_t efv_system_nnz(const efv_system *s);          /* DOF-expanded: ndof^2 * graph nnz */
int64_t efv_system_graph_nnz(const efv_system *s);
int64_t efv_system_slotmap_size(const efv_system *s); /* sum over elements of m^2 */
int efv_system_max_row_length(const efv_system *s);
int efv_system_export_csr(const efv_system *s, int64_t *indptr, int32_t *indices);     /* getSparse() structure */
int efv_system_export_graph(const efv_system *s, int64_t *gptr, int32_t *gcol);
int efv_system_export_slotmap(const efv_system *s, int64_t *slots);  /* graph slot of (e,i,j), element order, i-major */
int efv_system_export_values(const efv_system *s, double *data);     /* in efv_system_export_csr order */
int efv_vec_upload(efv_system *s, int which, const double *host);
int efv_vec_download(const efv_system *s, int which, double *host);

/* ---- assembly (north_star kernels 3, 4) -------------------------------------------------------------------
 * heat [ref: apps/heat_transfer.py:40-68 matrix, :89-113 RHS]: props = n_regions x 4 doubles
 * {Conductivity, Density, HeatCapacity, HeatGeneration}.  Fills the interior operator (diffusion +, if
 * transient, accumulation) and the lumped per-vertex vectors  gen = sum vol_k q,  acc = sum vol_k rho cp / dt.
 * The kernel reads the SoA geometry of efv_geometry_build (built on demand).                                */
int efv_heat_assemble(efv_system *s, const double *props, double dt, int transient);
/* elasticity [ref: apps/stress_equilibrium.py:81-117]: props = n_regions x 4 {ShearModulus, PoissonsRatio,
 * Density, Gravity}; gravity != 0 adds -rho g vol to the y component [ref: :81-90].  sign = +1 keeps the
 * reference's sign convention (negative-definite interior operator).                                       */
int efv_elasticity_assemble(efv_system *s, const double *props, int gravity);
/* Biot [ref: apps/geomechanics.py:39-115 matrix, :143-199 RHS operator parts]: props = n_regions x 9
 * {nu, G, cs, phi, k, cf, mu, rhos, rhof}; gvec = 3 doubles (reference uses 0).                              */
int efv_biot_assemble(efv_system *s, const double *props, double dt, const double *gvec);

/* boundary conditions.  Dirichlet rows are given in the order the reference visits them (boundaries in grid
 * order, vertices in first-appearance order): for duplicates the LAST value wins
 * [ref: apps/heat_transfer.py:122-126].  rows are field-major DOF indices.                                  */
int efv_bc_clear(efv_system *s);
int efv_bc_dirichlet(efv_system *s, const int64_t *rows, const double *values, int64_t n);
/* Neumann: static_rhs[dof(outer vertex)] += sign * value * |outer face area| over the outer faces of
 * `boundary`  [ref: apps/heat_transfer.py:115-120 (sign +1); apps/stress_equilibrium.py:131-141 (sign -1)].
 * Calls are ORDER-SENSITIVE like the reference's loops: a Neumann call made after a Dirichlet call that prescribed the same
 * DOF adds its flux to the prescribed value (the elasticity app visits boundary by boundary, stress_equilibrium.py:119-163);
 * a later Dirichlet call overwrites it.                                                                              */
int efv_bc_neumann(efv_system *s, int boundary, int dof, double value, double sign);
/* variable Neumann condition: one value per OUTER FACE of the boundary, in the reference's visiting order (facet by facet,
 * outer face by outer face); efv_boundary_outer_faces returns their number and the global number of the first one.
 * The host layer evaluates the expression exactly where the reference does [ref: apps/heat_transfer.py:116-120 calls
 * getValue(outerFace.handle), i.e. at the vertex whose NUMBER equals the outer face's global counter, Boundary.py:74]. */
int64_t efv_boundary_outer_faces(const efv_mesh *m, int boundary, int64_t *first);
int efv_bc_neumann_values(efv_system *s, int boundary, int dof, const double *values, int64_t n, double sign);
/* how the next assembly call treats the recorded Dirichlet set: -1 leave the interior operator (apply later with
 * efv_apply_dirichlet), EFV_DIRICHLET_ROWS or EFV_DIRICHLET_SYMMETRIC applied inside the assembly kernel.          */
int efv_assemble_dirichlet_mode(efv_system *s, int mode);
/* apply the recorded Dirichlet set to the matrix values (north_star kernel 4) */
int efv_apply_dirichlet(efv_system *s, int mode);
/* RHS of one step into EFV_VEC_B:  heat: gen + acc * XOLD + neumann, Dirichlet rows <- g
 * [ref: apps/heat_transfer.py:89-132]; elasticity: gravity + neumann, Dirichlet rows <- g
 * [ref: apps/stress_equilibrium.py:119-163]; Biot: + S/dt vol p_old + alpha/dt N.s u_old terms
 * [ref: apps/geomechanics.py:166-239].  In symmetric mode the eliminated columns are folded in.            */
int efv_build_rhs(efv_system *s, int transient);

/* ---- fixed-stress split [ref: apps/fixed_stress.py:39-329] ------------------------------------------------------------
 * Three systems on one mesh: `biot` (d+1 DOF, efv_biot_assemble with Dirichlet mode -1: the unmodified coupled operator;
 * X = current iterate (u_k, p_k), XOLD = previous time step), `mass` (1 DOF: efv_heat_assemble with Conductivity = k*mu,
 * Density*HeatCapacity = S + cb*alpha^2, transient dt [ref: :64-77]) and `geo` (d DOF: efv_elasticity_assemble [ref: :80-92]).
 *   efv_fs_mass_rhs : mass.B = S/dt vol p_prev + cb alpha^2/dt vol p_k + A_pu (u_prev - u_k) + Neumann; Dirichlet <- g [ref: :119-199]
 *   efv_fs_geo_rhs  : geo.B  = - A_up p_next (= mass.X) + Neumann; Dirichlet <- g                                    [ref: :201-275]
 *   efv_fs_update   : difference = max(|p_next - p_k|, |u_next - u_k|); (u_k, p_k) <- (geo.X, mass.X)                [ref: :298-303] */
int efv_fs_mass_rhs(efv_system *biot, efv_system *mass);
int efv_fs_geo_rhs(efv_system *biot, efv_system *mass, efv_system *geo);
int efv_fs_update(efv_system *biot, efv_system *mass, efv_system *geo, double *difference);

/* ---- multi-GPU: vertex-block partition with owned + ghost vertices (the reference has no
 * distributed code).  One process per GPU.  The unique id is created on rank 0 and distributed by the host layer
 * (torch.distributed); nccl_library_path names the NCCL shared object to dlopen (NULL: default search).  A partitioned
 * system is created on the rank's LOCAL mesh whose vertices are numbered [owned | ghosts grouped by owner rank];
 * efv_system_set_owned restricts assembly / SpMV / vector kernels to the owned rows, efv_system_set_halo registers, per
 * neighbour rank, which owned vertices to send and which slice of the ghost block to receive.  PCG then exchanges the
 * halo of the search direction before every SpMV and all-reduces its dot products.                                   */
int efv_comm_unique_id(const char *nccl_library_path, char *out128);
int efv_comm_init(const char *nccl_library_path, int rank, int nranks, const char *id128);   /* id128 NULL: peer-memory only, no NCCL */
int efv_comm_destroy(void);
/* peer-memory collectives over NVLink (CUDA IPC): every rank exports a 64-byte handle of its scalar region, the host
 * layer all-gathers them (rank order) and hands the table back; the Krylov all-reduces then run as one kernel that stores
 * into every peer and sums in rank order.  Same for the halo: export the system's staging buffer, import the neighbours'
 * handles with the ghost offset / flag slot this rank owns inside each of them and the neighbours' ghost counts (the
 * staging buffers are double-buffered by exchange parity).  Without these calls NCCL is used.  The host layer must take
 * the peer path on ALL ranks or on none (pyefvlib_b200/distributed.py: init_comm).                                   */
int efv_comm_peer_export(char *out64);
int efv_comm_peer_import(const char *handles);
int efv_halo_stage_export(efv_system *s, char *out64);
int efv_comm_peer_release(void);   /* unmap the peers' regions: every later collective uses NCCL */
int efv_halo_stage_import(efv_system *s, const char *handles, const int64_t *remote_offsets, const int *remote_slots,
                          const int64_t *remote_ghost_counts);
int efv_system_set_owned(efv_system *s, int64_t n_owned);
int efv_system_set_halo(efv_system *s, int n_neighbours, const int *neighbour_ranks, const int64_t *send_ptr,
                        const int32_t *send_idx, const int64_t *recv_ptr);
int efv_halo_exchange(efv_system *s, int which);

/* ---- Krylov solvers (north_star kernel 5) -----------------------------------------------------------------
 * replace the reference's sparse/dense inverse [ref: apps/heat_transfer.py:78-81,134-135;
 * apps/stress_equilibrium.py:170-173; apps/geomechanics.py:140,254].  Solve A x = B from initial guess X;
 * stop when ||r||_2 <= rtol * ||b||_2.  scale: multiply the system by -1 internally (elasticity) so that CG
 * sees an SPD operator.  Outputs: iterations done and final relative residual.                             */
int efv_solve_pcg(efv_system *s, double rtol, int maxit, int negate, int *iters, double *relres);
/* Preconditioner of efv_solve_pcg: 0 = Jacobi (default), 1 = aggregation multigrid (plain aggregation in ~8-vertex aggregates,
 * Galerkin coarse operators, Chebyshev-smoothed V-cycle built from the SpMV kernels; csrc/amg.cu).  The hierarchy is symbolic
 * per pattern and re-summed after every assembly.  efv_system_set_lattice tells the system that its owned vertices are
 * numbered i + nx (j + ny k) on an nx x ny x nz lattice (structured meshes and their slabs): aggregates are then 2x2x2 blocks
 * computed arithmetically on the device; otherwise they come from three rounds of pairwise matching on the graph.
 * efv_amg_info reports the hierarchy (levels, rows / nnz / Chebyshev bound per level, time of the last numeric set-up). */
int efv_system_set_preconditioner(efv_system *s, int kind);
int efv_system_set_lattice(efv_system *s, int nx, int ny, int nz);
int efv_amg_info(efv_system *s, int capacity, int *levels, int64_t *rows, int64_t *nnz, double *lmax, double *numeric_ms);
/* one level on the host (any pointer may be NULL): block-CSR pattern over the level's rows, values (ndof x ndof blocks, row-major
 * inside a block), aggregate of every row (not on the coarsest level) -- lets the tests rebuild P^T A P with scipy */
int efv_amg_level_export(efv_system *s, int level, int64_t *gptr, int32_t *gcol, double *vals, int32_t *aggregates);
int efv_solve_bicgstab(efv_system *s, double rtol, int maxit, int *iters, double *relres);
int efv_spmv(efv_system *s, int which_in, int which_out);   /* y = A x on two of the system's vectors */
/* Checks that never trust the solver's own recurrences: the TRUE relative residual ||B - A X||_2 / ||B||_2 recomputed with one
 * extra SpMV (all-reduced over the ranks of a partitioned system); sum and sum of squares of the owned entries of a vector
 * (cross-checks of N-GPU against 1-GPU runs); max |a_ij - a_ji| / max |a_ij| of the assembled operator, which tells the apps
 * at assembly time whether CG applies [ref: the reference never needs this, it inverts: apps/heat_transfer.py:78-81]. */
int efv_true_residual(efv_system *s, int negate, double *relres);
int efv_vec_sums(efv_system *s, int which, double *sum, double *sum_of_squares);
int efv_symmetry_defect(efv_system *s, double *defect);
/* max_i |X_i - XOLD_i|  [ref: apps/heat_transfer.py:142] and X -> XOLD copy */
int efv_max_abs_diff(efv_system *s, double *out);
int efv_advance(efv_system *s);
/* device-resident time loop: the field of the step just advanced (XOLD) is downloaded on a copy stream
 * into pinned memory while the next step is assembled and solved; _end waits for the copy and returns the field (field-major).
 * Fields are downloaded only when a saver asks for them [ref: apps/heat_transfer.py:24-34,137-138]. */
int efv_download_xold_begin(efv_system *s);
int efv_download_xold_end(efv_system *s, double *host);
/* per-kernel timing of the last solve / assemble (CUDA events on the library stream), milliseconds */
int efv_last_timings(efv_system *s, double *assemble_ms, double *solve_ms);

#ifdef __cplusplus
}
#endif
#endif /* EFV_B200_H */
