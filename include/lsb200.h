/*
 * lsb200.h -- C ABI of the B200-native discrete implicit-interpolation engine.
 *
 * Drop-in boundary for LoopStructural's FiniteDifferenceInterpolator /
 * P1Interpolator hot path (assemble -> solve -> evaluate).  The reference has
 * no FFI of its own (pure Python); each entry point below names the reference
 * interface it replaces, with the SURVEY.md prefixes
 *   LI/ = /root/reference/packages/loop_interpolation/src/loop_interpolation/
 *   LC/ = /root/reference/packages/loop_common/src/loop_common/
 *
 * Conventions
 *   - plain C types only; every pointer argument is a HOST pointer unless the
 *     name ends in _dev (then it is a device pointer on the handle's GPU);
 *   - all functions return 0 on success, <0 on error; the message is then
 *     available from lsb200_last_error() (thread-local);
 *   - a handle is not re-entrant; independent handles may be used from
 *     different threads;
 *   - coordinates are the support's LOCAL coordinates (the reference projects
 *     world->local before it reaches the support, LI/_geological_interpolator.py:249-286);
 *   - node vectors (c, p, q, diag ...) are float64, length nx*ny*nz, flat index
 *     gi = i + nx*j + nx*ny*k (LC/supports/_3d_base_structured.py:346-350).
 *     On a z-slab handle they cover the owned planes [k0,k1) only.
 *   - there is NO CPU fallback: every compute entry point fails with
 *     LSB200_ERR_CUDA when no CUDA device is usable.
 */
#ifndef LSB200_H
#define LSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSB200_OK 0
#define LSB200_ERR_ARG (-1)
#define LSB200_ERR_CUDA (-2)
#define LSB200_ERR_STATE (-3)
#define LSB200_ERR_COMM (-4)
#define LSB200_ERR_ALLOC (-5)

#define LSB200_ABI_VERSION 1

typedef struct lsb200_ctx* lsb200_handle;

/* index of the nine effective regularisation weights of lsb200_set_regularisation */
enum {
  LSB200_W_DXX = 0, LSB200_W_DYY = 1, LSB200_W_DZZ = 2,
  LSB200_W_DXY = 3, LSB200_W_DXZ = 4, LSB200_W_DYZ = 5,
  LSB200_W_DX = 6, LSB200_W_DY = 7, LSB200_W_DZ = 8
};

/* kinds of point-generated rows, lsb200_add_point_rows */
enum {
  LSB200_ROWS_VALUE = 0,    /* data (n,5): x y z value w        -> 1 row / point  */
  LSB200_ROWS_GRADIENT = 1, /* data (n,7): x y z gx gy gz w     -> 2 rows / point */
  LSB200_ROWS_NORMAL = 2,   /* data (n,7): x y z nx ny nz w     -> 3 rows / point */
  LSB200_ROWS_TANGENT = 3,  /* data (n,7): x y z tx ty tz b     -> 1 row / point  */
  /* tet meshes only: v . grad(phi) = b on the tetrahedron that holds the point (its barycentre, for the callers this
   * exists for); unlike a tangent row v is taken as given and the weight is not multiplied by the element volume.  The
   * fold orientation / fold axis / fold normalisation rows of DiscreteFoldInterpolator.add_fold_constraints
   * (LI/_discrete_fold_interpolator.py:116-190): vol (v . eg) x = vol b, L2-normalised like every row. */
  LSB200_ROWS_ELEMENT_VECTOR = 4 /* data (n,7): x y z vx vy vz b -> 1 row / point */
};

/* support kinds */
enum { LSB200_SUPPORT_GRID = 0, LSB200_SUPPORT_TETMESH = 1 };

int lsb200_abi_version(void);
const char* lsb200_last_error(void);
/* number of usable CUDA devices (0 without a GPU; never fails) */
int lsb200_device_count(void);

/* ---- supports --------------------------------------------------------- */

/* StructuredGrid with nsteps = NODE counts (LC/supports/_3d_base_structured.py:66),
 * owning the global z-planes [k0,k1) (k0=0,k1=nsteps[2] for a single-GPU handle).
 * Replaces the role of StructuredGrid as the support of FiniteDifferenceInterpolator
 * (LI/_finite_difference_interpolator.py:66-130). */
int lsb200_create_grid(const int64_t nsteps[3], const double origin[3], const double step[3],
                       int64_t k0, int64_t k1, int device, lsb200_handle* out);

/* TetMesh on the same lattice: 5 tets per cell, parity masks
 * (LC/supports/_3d_structured_tetra.py:28-34); nsteps = NODE counts.
 * Support of P1Interpolator (LI/_p1interpolator.py:23-68). */
int lsb200_create_tetmesh(const int64_t nsteps[3], const double origin[3], const double step[3],
                          int device, lsb200_handle* out);

/* the same support owning the node planes [k0,k1) only (z-slab of a multi-GPU P1 solve; evaluation needs a handle
 * that owns every plane) */
int lsb200_create_tetmesh_slab(const int64_t nsteps[3], const double origin[3], const double step[3],
                               int64_t k0, int64_t k1, int device, lsb200_handle* out);

int lsb200_destroy(lsb200_handle h);

/* number of node values held by this handle (owned planes) and the support kind */
int64_t lsb200_local_dof(lsb200_handle h);
int lsb200_support_kind(lsb200_handle h);

/* ---- regularisation ---------------------------------------------------- */

/* FDI: the twelve stencil families of FiniteDifferenceInterpolator.setup_interpolator
 * (assemble_inner LI/_finite_difference_interpolator.py:1146-1164, assemble_borders :1052-1144,
 * _assemble_operator :643-692) are never materialised; only their EFFECTIVE row weights
 * are given (mixed terms already divided by 4 as get_operators does,
 * LC/supports/_3d_structured_grid.py:489-510).  Order: LSB200_W_*. */
int lsb200_set_regularisation(lsb200_handle h, const double w[9]);

/* P1: face gradient-jump regularisation, P1Interpolator.minimise_edge_jumps (LI/_p1interpolator.py:115-166), with
 * row weight w (cgw), applied matrix-free as a variable-coefficient node stencil (csrc/p1.cuh).  The stencil
 * coefficients are the rows of R^T R assembled by the caller on a small mesh of m[a] cells per axis with the same
 * step vector: m[a] = cells if cells <= 6, else 5 (odd cell count) or 6 (even).  table is
 * [(m[2]+1)][(m[1]+1)][(m[0]+1)][125] row-major, tap (dz+2)*25 + (dy+2)*5 + (dx+2), for unit weight
 * (loopstructural_b200/p1_tables.py builds it). */
int lsb200_set_p1_regularisation(lsb200_handle h, double w);
int lsb200_set_p1_tables(lsb200_handle h, const int64_t m[3], const double* table);

/* ---- data rows --------------------------------------------------------- */

/* Generic row ingestion = DiscreteInterpolator.add_constraints_to_least_squares
 * (LI/_discrete_interpolator.py:316-389).  vals/cols are (n_rows,k) row-major; cols are GLOBAL
 * node ids.  If normalise!=0 each row (and its b) is divided by the L2 norm of its listed
 * entries (:361-366); w is the per-row weight (rows enter the normal equations with w^2).
 * On a slab handle rows whose columns do not touch planes [k0-?]: rows are kept when every column lies in
 * the owned+halo planes; the caller distributes rows (see loopstructural_b200.partition). */
int lsb200_add_rows(lsb200_handle h, int64_t n_rows, int k, const double* vals, const int64_t* cols,
                    const double* b, const double* w, int normalise);

/* Device-side row generation from constraint points, replacing
 * FDI.add_value_constraints (LI/_finite_difference_interpolator.py:236-269),
 * add_gradient_constraints (:342-401), add_norm_constraints (:403-476),
 * add_gradient_orthogonal_constraints (:478-539), and on a tet mesh
 * P1Interpolator.add_value_constraints / add_norm_constraints /
 * add_gradient_orthogonal_constraints (LI/_p1interpolator.py:73-113,248-296).
 * data is (n,ncols) row-major as in GeologicalInterpolator.data (LI/_geological_interpolator.py:872-880);
 * weight is the family weight (cpw / gpw / npw / tpw).  Points outside the support are dropped.
 * n_rows_out (may be NULL) receives the number of rows created. */
int lsb200_add_point_rows(lsb200_handle h, int kind, int64_t n, const double* data, int ncols,
                          double weight, int64_t* n_rows_out);

/* Direction-weighted face-jump rows of a tet mesh: v_f . (grad phi_a - grad phi_b) = 0 for n pairs of face-adjacent
 * tetrahedra, P1Interpolator.minimise_edge_jumps with a vector per shared face (LI/_p1interpolator.py:115-166) -- what
 * DiscreteInterpolator.add_directional_regularisation (LI/_discrete_interpolator.py:413-451) and the 'fold regularisation
 * n' families of DiscreteFoldInterpolator.add_fold_constraints (LI/_discrete_fold_interpolator.py:204-242) put into the
 * system.  nodes_a, nodes_b are (n,4) row-major GLOBAL node ids of the two tetrahedra (support.elements[tri]), vectors
 * (n,3) the direction at the shared face; 8 listed entries per row, L2-normalised over them as the reference does before
 * its duplicate merge, right-hand side 0, weight the family weight.  Rows holding a NaN are dropped (the reference drops
 * the whole family then: the caller checks).  n_rows_out (may be NULL) receives the number of rows kept. */
int lsb200_add_face_vector_rows(lsb200_handle h, int64_t n, const int64_t* nodes_a, const int64_t* nodes_b,
                                const double* vectors, double weight, int64_t* n_rows_out);

/* Directional regularisation of a grid, FiniteDifferenceInterpolator.minimise_directional_gradient_change
 * (LI/_finite_difference_interpolator.py:1207-1340; reached through setup_interpolator(directional_regularisation=...),
 * LI/_discrete_interpolator.py:413-451): the six second-derivative masks at every node interior on all axes as rows
 * weighted by weight * (vx^2, vy^2, vz^2, 2 vx vy, 2 vx vz, 2 vy vz) of the direction at the node.  vectors is the
 * (nx*ny*nz, 3) row-major direction field in GLOBAL node order (support.nodes order); a slab handle reads the planes it
 * needs.  Rows of weight 0 (or NaN) are dropped as in the reference.  n_rows_out (may be NULL): rows kept. */
int lsb200_add_directional_rows(lsb200_handle h, const double* vectors, double weight, int64_t* n_rows_out);

/* Interface rows N(x_i) - N(x_j) = 0 for n point pairs (FDI.add_interface_constraints,
 * LI/_finite_difference_interpolator.py:297-340): 16 listed entries per row, weight ipw.  The pairing per
 * interface id (all / star / chain, :271-295) is done by the caller; pairs with a point outside the grid are
 * dropped.  xyz_i, xyz_j are (n,3) row-major. */
int lsb200_add_pair_rows(lsb200_handle h, int64_t n, const double* xyz_i, const double* xyz_j, double weight,
                         int64_t* n_rows_out);

int lsb200_clear_rows(lsb200_handle h);
int64_t lsb200_num_rows(lsb200_handle h);

/* ---- solver ------------------------------------------------------------ */

/* ridge_factor / add_ridge / apply_scaling_matrix of DiscreteInterpolator
 * (LI/_discrete_interpolator.py:76-78; LI/_solver_pipeline.py:86-124).
 * deterministic!=0 selects fixed-order reductions for the data rows. */
int lsb200_set_options(lsb200_handle h, double ridge_factor, int add_ridge, int apply_scaling, int deterministic);

/* Build rhs = A^T W b, diag(A^T W A + ridge^2 I) and the column scaling
 * (preprocess_main_system, LI/_solver_pipeline.py:86-124; compute_column_scaling_matrix,
 * LI/_discrete_interpolator.py:815-833) on the device.  Called implicitly by solve. */
int lsb200_finalize_system(lsb200_handle h);

/* Join an NCCL communicator (z-slab solve): nccl_id is the 128-byte ncclUniqueId
 * created on rank 0 by lsb200_comm_unique_id and broadcast by the host (torch.distributed). */
int lsb200_comm_unique_id(unsigned char id_out[128]);
int lsb200_comm_init(lsb200_handle h, const unsigned char nccl_id[128], int rank, int nranks);

typedef struct {
  int32_t info;            /* scipy convention: 0 converged, >0 = maxiter reached */
  int64_t iterations;      /* completed CG iterations (== number of scipy callbacks) */
  double residual_norm;    /* ||r|| in the (column-scaled) variables the test is made in */
  double threshold;        /* max(atol, rtol*||rhs||) in the same variables */
  double rhs_norm;
  double setup_seconds;    /* finalize_system: rhs, diagonal, scaling */
  double solve_seconds;    /* CG loop, device time (CUDA events) */
  double matvec_seconds;   /* device time inside the normal-operator kernels (profiled solves only) */
  int64_t kernel_launches; /* kernels launched by this call */
} lsb200_solve_info;

/* Device-resident CG on (A^T W A + ridge^2 I) c = A^T W b, = solve_with_cg
 * (LI/_solver_strategy.py:48-71) on the column-scaled system, i.e. scipy.sparse.linalg.cg
 * with M=None on S N S, stopping when ||r|| < max(atol, rtol*||S A^T b||) tested at the top
 * of every iteration; maxiter<=0 means 10*dof.  x0 (may be NULL) is in c-space (unscaled).
 * c_out (may be NULL) receives the owned node values. */
int lsb200_solve_cg(lsb200_handle h, double rtol, double atol, int64_t maxiter, const double* x0,
                    double* c_out, lsb200_solve_info* info);

/* q = (A^T W A + ridge^2 I) p (unscaled), for operator-level parity tests against
 * LI/_finite_difference_interpolator.py:884-1050 and the explicit Gram matrix.
 * parts: bit0 stencil/face regularisation, bit1 data rows, bit2 ridge. */
int lsb200_apply_normal(lsb200_handle h, const double* p, double* q, int parts);

int lsb200_get_rhs(lsb200_handle h, double* out);
int lsb200_get_diagonal(lsb200_handle h, double* out);
int lsb200_get_scaling(lsb200_handle h, double* out); /* S = diag^-1/2 (ones when scaling is off) */
/* Replace the right-hand side A^T W b of the normal equations by a vector of the caller (owned planes, nx*ny*nzl doubles)
 * until the rows change again: the inner solves of the reference's ADMM solver keep the operator and change only the
 * right-hand side (LI/loopsolver/admm_solver.py:243-316: [A; rho Q]^T [b - A m; rho (z - u - Q m)]). */
int lsb200_set_rhs(lsb200_handle h, const double* rhs);

/* 1 when the two scalar reductions of a CG iteration (p.q, r.S^2 r) go through peer-mapped mailboxes (cudaIpc, stores
 * over NVLink) instead of ncclAllReduce; decided in lsb200_comm_init (all ranks agree).  0 on one rank. */
int lsb200_comm_uses_peer_memory(lsb200_handle h);

/* Residual of the current solution, recomputed through the operator (not the recurrence the solver carried):
 * r = A^T W b - N c with N applied by the same kernels as inside CG (halo exchange on z slabs);
 * *scaled_norm = sqrt(r . S^2 r) -- the quantity scipy's stopping test sees on the column-scaled system
 * (LI/_solver_strategy.py:48-71, LI/_solver_pipeline.py:86-124) -- and *plain_norm = ||r||_2, both over ALL ranks.
 * Collective on slab handles.  Uses the search-direction vector as scratch. */
int lsb200_residual_norm(lsb200_handle h, double* scaled_norm, double* plain_norm);

/* solved node values: DiscreteInterpolator.c (LI/_discrete_interpolator.py:1312-1318) */
int lsb200_get_c(lsb200_handle h, double* out);
int lsb200_set_c(lsb200_handle h, const double* c);

/* the same with a device buffer on the handle's GPU holding the owned planes contiguously (nx*ny*nzl doubles);
 * get is asynchronous on `stream`, set returns when the copy has completed.  For replicating the solution of a
 * z-slab solve on every GPU (all-gather of the slabs, then lsb200_set_c_dev on a handle that owns every plane)
 * before evaluation. */
int lsb200_get_c_dev(lsb200_handle h, double* c_dev, void* stream);
int lsb200_set_c_dev(lsb200_handle h, const double* c_dev, void* stream);

/* ---- evaluation -------------------------------------------------------- */

/* StructuredGrid.evaluate_value (LC/supports/_3d_structured_grid.py:295-322) +
 * DiscreteInterpolator._evaluate_value_local (LI/_discrete_interpolator.py:1342-1362):
 * NaN outside the support, 0.0 for rows holding NaN coordinates.  TetMesh.evaluate_value
 * (LC/supports/_3d_structured_tetra.py:208-229) on a tet-mesh handle.  xyz is (n,3) row-major. */
int lsb200_eval_value(lsb200_handle h, int64_t n, const double* xyz, double* out);

/* StructuredGrid.evaluate_gradient (LC/supports/_3d_structured_grid.py:324-391): 0 outside;
 * TetMesh.evaluate_gradient (LC/supports/_3d_structured_tetra.py:231-262): NaN outside.
 * out is (n,3) row-major, local axes. */
int lsb200_eval_gradient(lsb200_handle h, int64_t n, const double* xyz, double* out);

/* same, device pointers, asynchronous on `stream` (a cudaStream_t passed as void*) */
int lsb200_eval_value_dev(lsb200_handle h, int64_t n, const double* xyz_dev, double* out_dev, void* stream);
int lsb200_eval_gradient_dev(lsb200_handle h, int64_t n, const double* xyz_dev, double* out_dev, void* stream);

/* ---- fault-frame evaluation (the consumers of evaluate_value / evaluate_gradient) ------------------------- */

/* One 1-D fault profile: the closed family of LoopStructural/modelling/features/fault/_fault_function.py --
 * Zeros / Ones (:256-305), CubicFunction (:40-140: w0 v^3 + w1 v^2 + w2 v + w3, clamped to its value at lim[]
 * outside the limits, then to [min_v, max_v]) and smooth_peak (:12-16). */
enum { LSB200_PROFILE_ZEROS = 0, LSB200_PROFILE_ONES = 1, LSB200_PROFILE_CUBIC = 2, LSB200_PROFILE_SMOOTH_PEAK = 3 };
typedef struct lsb200_profile1d {
  int32_t kind;
  int32_t reserved;
  double w[4];
  double lim[2];
  double min_v, max_v;
} lsb200_profile1d;

/* FaultDisplacement (:308-369): d = scale * gx(v0) * gy(v1) * gz(v2) over the three fault-frame coordinates.
 * composite[a] != 0: axis a is Composite(pos, neg) (:198-226; pos for v > 0, neg for v < 0, 0 at 0); otherwise
 * pos[a] is used for every v.  has_function == 0: the segment has no fault function, d = [v0 > 0]. */
typedef struct lsb200_fault_profile {
  int32_t has_function;
  int32_t composite[3];
  double scale;
  lsb200_profile1d pos[3];
  lsb200_profile1d neg[3];
} lsb200_fault_profile;

/* The three solved fields of a fault frame (StructuralFrame, coordinates 0 / 1 / 2) are three grid handles on one
 * device, each owning every plane and holding a solution.  xyz are WORLD coordinates; transforms (may be NULL =
 * identity, i.e. xyz already local to all three supports) holds per field 21 doubles: the upper 3 x 4 of
 * BoundingBox.world_to_local_matrix row-major (project, LC/geometry/_bounding_box.py:118-176,702-717) followed by the
 * 3 x 3 rotation of local_to_world_matrix row-major (reproject_vectors, :750-761).
 *
 * lsb200_frame_eval: the gathers one integration step of a fault needs, in ONE pass over the points: values
 * (n,3) = GeologicalInterpolator.evaluate_value of the three fields (LI/_geological_interpolator.py:723-730; NaN
 * outside a support, 0.0 for NaN coordinates) and, when gradient != NULL, gradient (n,3) = evaluate_gradient of field
 * `gradient_of` in world axes (:732-741; 0 outside). */
int lsb200_frame_eval(const lsb200_handle fields[3], const double* transforms, int gradient_of, int64_t n,
                      const double* xyz, double* values, double* gradient);

/* FaultSegment.apply_to_points (LoopStructural/modelling/features/fault/_fault_segment.py:315-423): every point
 * with a non-zero displacement is carried through `steps` steps of  x += (1/steps) d(x) g1(x)/|g1(x)|, d = profile of
 * the three frame values * displacement, g1 = gradient of field 1 -- one thread per point, all steps on the device
 * (the reference runs 4 gathers per step as separate numpy passes).  The reference's `reverse` flag does not
 * change its result (the reversed d is overwritten inside the loop, :360-363 vs :403), so it has no counterpart
 * here.  moved is (n,3) and must not alias xyz. */
int lsb200_fault_displace(const lsb200_handle fields[3], const double* transforms, const lsb200_fault_profile* profile,
                          double displacement, int steps, int64_t n, const double* xyz, double* moved);

/* n_faults faults applied one after the other to the same points (BaseFeature._apply_faults,
 * LoopStructural/modelling/features/_base_geological_feature.py:252-270): fields holds 3 handles per fault, transforms
 * 63 doubles per fault (or NULL for all), one profile / displacement / step count per fault.  The points cross the
 * bus once in each direction, whatever the number of faults. */
int lsb200_fault_displace_chain(int n_faults, const lsb200_handle* fields, const double* transforms,
                                const lsb200_fault_profile* profiles, const double* displacements, const int* steps,
                                int64_t n, const double* xyz, double* moved);

/* Evaluation of a faulted feature in one call: the chain above, then the value (gradient == 0: out is (n)) or the
 * world-axis gradient (gradient != 0: out is (n,3)) of `field` at the moved points, which never leave the device
 * (GeologicalFeature.evaluate_value: _apply_faults then interpolator.evaluate_value,
 * LoopStructural/modelling/features/_geological_feature.py:116-157; evaluate_gradient of an unfaulted interpolator at
 * faulted points, :231-236).  field_transform: 21 doubles as in lsb200_frame_eval or NULL.  n_faults may be 0. */
int lsb200_faulted_eval(int n_faults, const lsb200_handle* fields, const double* transforms,
                        const lsb200_fault_profile* profiles, const double* displacements, const int* steps,
                        lsb200_handle field, const double* field_transform, int gradient, int64_t n, const double* xyz,
                        double* out);

/* FaultSegment.evaluate_displacement (:283-313): d (n) = profile(v0 + fault_offset, v1, v2) * displacement. */
int lsb200_fault_displacement(const lsb200_handle fields[3], const double* transforms,
                              const lsb200_fault_profile* profile, double displacement, double fault_offset, int64_t n,
                              const double* xyz, double* d);

/* same three, device pointers on the handles' GPU, asynchronous on `stream` */
int lsb200_frame_eval_dev(const lsb200_handle fields[3], const double* transforms, int gradient_of, int64_t n,
                          const double* xyz_dev, double* values_dev, double* gradient_dev, void* stream);
int lsb200_fault_displace_dev(const lsb200_handle fields[3], const double* transforms,
                              const lsb200_fault_profile* profile, double displacement, int steps, int64_t n,
                              const double* xyz_dev, double* moved_dev, void* stream);
int lsb200_fault_displacement_dev(const lsb200_handle fields[3], const double* transforms,
                                  const lsb200_fault_profile* profile, double displacement, double fault_offset,
                                  int64_t n, const double* xyz_dev, double* d_dev, void* stream);

/* Frame-orthogonality rows from a solved field (GeologicalFeatureBuilder.install_gradient_constraint,
 * LoopStructural/modelling/features/builders/_geological_feature_builder.py:293-318 -> FDI
 * add_gradient_orthogonal_constraints, LI/_finite_difference_interpolator.py:478-539): for every element_step-th cell e
 * of h's support one row  v . grad(phi) = b  at the cell's barycentre, v = the normalised gradient of `other`
 * (a solved grid handle on the same device owning every plane) there, row-normalised, weight `weight`.  The
 * barycentres (h's LOCAL coordinates) are handed to `other` through other_transform (21 doubles as in
 * lsb200_frame_eval, NULL = identity) exactly as the reference hands them to feature.evaluate_gradient.
 * n_rows_out (may be NULL): rows this handle keeps (all ceil(n_cells / element_step) on a full handle). */
int lsb200_add_field_orthogonal_rows(lsb200_handle h, lsb200_handle other, const double* other_transform,
                                     int64_t element_step, double weight, double b, int64_t* n_rows_out);

/* Evaluation on a regular lattice WITHOUT an xyz input: the points BoundingBox.regular_grid builds
 * (LC/geometry/_bounding_box.py:532-573: np.linspace(start[a], stop[a], n[a]) per axis, meshgrid 'ij', flattened in
 * order 'F' (order_f != 0, first axis fastest) or 'C') for BaseFeature.scalar_field / gradient_norm_scalar_field /
 * GeologicalModel.evaluate_model (LoopStructural/modelling/features/_base_geological_feature.py:359-449), generated
 * inside the kernel.  start/stop are WORLD coordinates, mapped through `transform` (21 doubles as in
 * lsb200_frame_eval, NULL = identity).  gradient == 0: out is (N) values with the semantics of lsb200_eval_value;
 * gradient != 0: out is (N,3) world-axis gradients (0 outside).  Grid handles owning every plane. */
int lsb200_eval_lattice(lsb200_handle h, const double* transform, const double start[3], const double stop[3],
                        const int64_t n[3], int order_f, int gradient, double* out);
int lsb200_eval_lattice_dev(lsb200_handle h, const double* transform, const double start[3], const double stop[3],
                            const int64_t n[3], int order_f, int gradient, double* out_dev, void* stream);

/* ---- isosurface of a lattice volume (SURVEY.md 8 f rank 4) --------------- */
/* What LoopIsosurfacer.fit (LoopStructural/utils/_surface.py:66-186, behind BaseFeature.surfaces,
 * LoopStructural/modelling/features/_base_geological_feature.py:317-357) asks of
 * skimage.measure.marching_cubes(volume, level, spacing=..., mask=~isnan(volume)): vertices (index coordinates times
 * spacing), triangles, vertex normals (towards decreasing values).  volume: dims[0] x dims[1] x dims[2] samples in C order
 * (host); NaN samples are masked out.  The triangulation is marching tetrahedra on the lattice (six Kuhn tetrahedra per
 * cell, one vertex per crossed lattice edge, watertight), not scikit-image's Lewiner cubes: same level set of the same
 * samples, more triangles.  Two steps: the first returns the counts and an object that holds the result on the device,
 * the second copies into caller-owned arrays (vertices, normals: 3 doubles per vertex; triangles: 3 int32 each). */
typedef struct lsb200_iso_result* lsb200_iso;
int lsb200_isosurface(int device, const int64_t dims[3], const double* volume, double level, const double spacing[3],
                      lsb200_iso* out, int64_t* n_vertices, int64_t* n_triangles);
int lsb200_isosurface_fetch(lsb200_iso iso, double* vertices, int32_t* triangles, double* normals);
int lsb200_isosurface_free(lsb200_iso iso);

/* Page-locked host memory for result buffers (c of a large grid is 134 MB - 1 GB: a pageable destination halves
 * the device->host rate and pays its page faults inside every solve). */
int lsb200_host_alloc(int64_t bytes, void** out);
int lsb200_host_free(void* p);

/* ---- measurement hooks (bench.py) -------------------------------------- */

/* run `reps` applications of the dominant kernel on the device-resident vectors and return
 * the average device time per launch in seconds (CUDA events on the launching stream).
 * which: 0 = regularisation normal matvec, 1 = whole CG iteration, 2 = data-row matvec, 3 = x/r update with the
 * fused r.S^2 r reduction, 4 = p update */
int lsb200_time_kernel(lsb200_handle h, int which, int reps, double* seconds_per_launch);

/* the launch plan of the regularisation operator as text (tile shape, chunks, CTAs; after lsb200_finalize_system the
 * plan chosen by timing the planner's candidates on this handle) */
int lsb200_describe_plan(lsb200_handle h, char* out, int64_t capacity);

/* one launch of the regularisation operator with the per-CTA trace on: out[3 b .. 3 b + 2] = {start ns, end ns, SM id}
 * of CTA b in block-index order (the single-launch plans only: k_fdi_fused, k_p1_tiles); *n_ctas = CTAs of the launch.
 * Shows which role of a launch (box chunk, z-face chunk, strip; tile, boundary) ends last. */
int lsb200_debug_cta_trace(lsb200_handle h, int64_t max_ctas, uint64_t* out, int64_t* n_ctas);
/* Diagnostics of the persistent CG loop of small systems (csrc/cg_small.cuh): mean microseconds per iteration that CTA 0
 * spends in each of the eight phases -- p update, barrier, operator, barrier, data rows, barrier + p.q, x / r update,
 * barrier + rho -- over `iters` iterations.  LSB200_ERR_STATE when the handle iterates launch by launch. */
int lsb200_debug_small_phases(lsb200_handle h, int iters, double* out_us);

#ifdef __cplusplus
}
#endif
#endif /* LSB200_H */
