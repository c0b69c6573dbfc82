/* adaflo_b200.h -- C ABI of the B200-native matrix-free operator library.
 *
 * This is the drop-in boundary for ONE hot path of kronbichler/adaflo: the
 * matrix-free operator application.  The reference has no
 * FFI; its boundary is a set of duck-typed C++ operator classes handed to
 * deal.II solver templates.  Every entry point below cites the reference
 * interface (file:line under the adaflo source tree) it replaces.  A thin C++
 * class with the reference's own method names sits on top of this ABI in
 * include/adaflo_b200/navier_stokes_matrix.h; INTEGRATION.md shows the deal.II
 * side adaptor.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    adaflo_b200_last_error() returns a thread-local message.
 *  - "host" pointers are read during the call and may be freed afterwards;
 *    "dev" pointers are CUDA device pointers owned by the caller.  The library
 *    never frees caller memory and never retains vector pointers across calls
 *    (exception: adaflo_b200_ns_set_old_solutions, which mirrors the reference
 *    constructor holding references to solution_old / solution_old_old).
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with ADAFLO_B200_ERR_CUDA.
 *
 * Vector layout (device): a DoF space with C components stores entry
 * (node, c) at index node*C + c; local vectors hold the owned nodes first and
 * the ghost nodes after them (deal.II LinearAlgebra::distributed::Vector
 * layout).  Velocity: C = dim; pressure and level set: C = 1.
 */
#ifndef ADAFLO_B200_H
#define ADAFLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADAFLO_B200_OK 0
#define ADAFLO_B200_ERR_ARG 1
#define ADAFLO_B200_ERR_CUDA 2
#define ADAFLO_B200_ERR_UNSUPPORTED 3 /* reference: AssertThrow(.., ExcNotImplemented()) navier_stokes_matrix.cc:68,82 */
#define ADAFLO_B200_ERR_STATE 4       /* reference: Assert(matrix_free != 0, ExcNotInitialized()) navier_stokes_matrix.cc:228 */
#define ADAFLO_B200_ERR_COMM 5
#define ADAFLO_B200_ERR_NO_CONVERGENCE 6 /* reference: SolverControl::NoConvergence */

/* per-component Dirichlet flags live in the top bits of each cell->node index */
#define ADAFLO_B200_FLAG_SHIFT 29u
#define ADAFLO_B200_NODE_MASK 0x1FFFFFFFu

typedef struct adaflo_b200_ctx_s *adaflo_b200_ctx_t;
typedef struct adaflo_b200_mf_s *adaflo_b200_mf_t;
typedef struct adaflo_b200_ns_s *adaflo_b200_ns_t;

const char *adaflo_b200_last_error(void);
const char *adaflo_b200_version(void);

/* ------------------------------------------------------------------ context
 * One context per rank/GPU.  `stream` is a cudaStream_t (NULL = default
 * stream).  Replaces the MPI communicator + TBB scheduler that deal.II's
 * MatrixFree owns in the reference (source/navier_stokes.cc:414-451). */
int adaflo_b200_ctx_create(adaflo_b200_ctx_t *out, int device, void *stream);
int adaflo_b200_ctx_destroy(adaflo_b200_ctx_t ctx);
int adaflo_b200_ctx_set_stream(adaflo_b200_ctx_t ctx, void *stream);
int adaflo_b200_ctx_synchronize(adaflo_b200_ctx_t ctx);
/* Multi-GPU: NCCL communicator owned by the library (replaces deal.II's
 * Utilities::MPI::Partitioner point-to-point exchange).  Rank 0
 * obtains the 128-byte id, the caller broadcasts it out of band. */
int adaflo_b200_comm_unique_id(unsigned char id[128]);
int adaflo_b200_ctx_init_comm(adaflo_b200_ctx_t ctx, const unsigned char id[128], int rank, int nranks);

/* -------------------------------------------------- matrix-free data mirror
 * Device mirror of dealii::MatrixFree<dim,double> as adaflo configures it
 * (source/navier_stokes.cc:414-451, source/two_phase_base.cc:239-273): a list
 * of DoF spaces (dof handler + constraints) and a list of quadratures with the
 * cell geometry evaluated on them. */
enum adaflo_b200_fe_kind {
  ADAFLO_B200_FE_Q = 0,        /* FE_Q(QGaussLobatto(degree+1)), possibly vector valued */
  ADAFLO_B200_FE_Q_DG0 = 1,    /* FE_Q_DG0(degree): +1 element-constant dof, last in the cell list */
  ADAFLO_B200_FE_Q_ISO_Q1 = 2  /* FE_Q_iso_Q1(degree subdivisions) (level set) */
};

typedef struct {
  int32_t peer;           /* rank */
  uint32_t n_send;        /* owned nodes this peer ghosts */
  const uint32_t *send;   /* host: local node indices, [n_send] */
  uint32_t recv_first;    /* first local ghost node filled by this peer */
  uint32_t n_recv;        /* contiguous range */
} adaflo_b200_exchange;

typedef struct {
  int32_t fe_kind;        /* enum adaflo_b200_fe_kind */
  int32_t degree;
  int32_t n_components;   /* dim for velocity, 1 otherwise */
  uint32_t n_owned_nodes;
  uint32_t n_ghost_nodes;
  uint32_t nodes_per_cell;
  /* host [n_cells][nodes_per_cell]: local node index, lexicographic within the
   * cell (x fastest), bit (29+c) set if component c of that node is constrained
   * to zero by a homogeneous Dirichlet constraint (read as 0, never written):
   * the semantics of FEEvaluation::read_dof_values / distribute_local_to_global
   * on constrained dofs. */
  const uint32_t *cell_nodes;
  /* host [n_cells] or NULL: refinement configuration of cells with hanging
   * faces/edges (bit layout documented in DESIGN.md); on such faces the cell
   * list holds the coarse neighbour's nodes */
  const uint32_t *hanging_masks;
  /* locally owned constrained vector entries (MatrixFree::get_constrained_dofs,
   * used at navier_stokes_matrix.cc:250-255) */
  uint32_t n_constrained;
  const uint32_t *constrained_dofs;
  int32_t n_exchange;
  const adaflo_b200_exchange *exchange;
  /* General homogeneous constraint rows (dealii::AffineConstraints as MatrixFree resolves them inside
   * read_dof_values / distribute_local_to_global; hanging nodes of any degree), used by the degree / dimension dispatch
   * (csrc/ns_generic.cu): the LAST n_hanging_nodes owned nodes are constrained nodes; row r belongs to node
   * n_owned_nodes - n_hanging_nodes + r and reads  value = sum_k hanging_weights[k] * value(hanging_cols[k]),
   * k in [hanging_row_ptr[r], hanging_row_ptr[r + 1]).  hanging_cols are node indices with the Dirichlet flag bits of
   * cell_nodes (a flagged component of a parent is read as 0 and never written).  Cell lists name the constrained node
   * itself; read_dof_values_plain (residual, navier_stokes_matrix.cc:663-664) reads its vector entry.  All zero / NULL on
   * conforming meshes.  The tuned 3-D Q2/Q1 kernels use hanging_masks instead. */
  uint32_t n_hanging_nodes;
  const uint32_t *hanging_row_ptr; /* host [n_hanging_nodes + 1] */
  const uint32_t *hanging_cols;    /* host [hanging_row_ptr[n_hanging_nodes]] */
  const double *hanging_weights;   /* host, same length */
} adaflo_b200_space_desc;

enum adaflo_b200_geometry_type {
  ADAFLO_B200_GEOM_CARTESIAN = 0, /* table rows of 4: 1/hx, 1/hy, 1/hz, |J| */
  ADAFLO_B200_GEOM_AFFINE = 1,    /* table rows of 10: K[f][e] = d xi_f / d x_e (9), |J| */
  ADAFLO_B200_GEOM_GENERAL = 2    /* per quadrature point */
};

typedef struct {
  int32_t n_q_1d;            /* QGauss(n_q_1d) per direction */
  int32_t geometry_type;
  uint32_t n_table_rows;     /* cartesian / affine */
  const double *table;       /* host */
  const uint32_t *cell_row;  /* host [n_cells] row of each cell (deal.II compresses equal Jacobians) */
  const double *inv_jacobian;/* general, host [n_cells][9][n_q], entry f*3+e = d xi_f/d x_e */
  const double *jxw;         /* general, host [n_cells][n_q] */
} adaflo_b200_quad_desc;

typedef struct {
  int32_t dim;               /* 3 (2 is not implemented yet) */
  uint32_t n_cells;          /* locally owned cells */
  int32_t n_spaces;
  const adaflo_b200_space_desc *spaces;
  int32_t n_quads;
  const adaflo_b200_quad_desc *quads;
} adaflo_b200_mf_desc;

int adaflo_b200_mf_create(adaflo_b200_ctx_t ctx, const adaflo_b200_mf_desc *desc, adaflo_b200_mf_t *out);
int adaflo_b200_mf_destroy(adaflo_b200_mf_t mf);
/* vector sizes: MatrixFree::initialize_dof_vector */
int adaflo_b200_mf_vector_size(adaflo_b200_mf_t mf, int space, uint64_t *n_owned_entries, uint64_t *n_local_entries);
/* Dof numbering of the host application.  The library's vectors are node-major with interleaved components
 * (entry = n_components * node + component).  deal.II numbers the dofs of FESystem(FE_Q(2), dim) the same way only until
 * DoFRenumbering::Cuthill_McKee is applied to the velocity DoFHandler, which adaflo does for its ILU (source/navier_stokes.cc:
 * 208-209).  host_position[i] = position of library entry i in the host's locally owned vector (a permutation, [n_owned
 * entries]; NULL restores the default).  Used by every host-pointer entry (ns_*_host, mf_upload_vector / mf_download_vector);
 * device-resident Krylov vectors stay in the library's layout and never see the permutation. */
int adaflo_b200_mf_set_dof_numbering(adaflo_b200_mf_t mf, int space, const uint32_t *host_position);
int adaflo_b200_mf_upload_vector(adaflo_b200_mf_t mf, int space, double *vec_dev, const double *vec_host);   /* owned entries, synchronous */
int adaflo_b200_mf_download_vector(adaflo_b200_mf_t mf, int space, double *vec_host, const double *vec_dev);
/* MatrixFree::get_dof_handler(space).get_fe(): kind (ADAFLO_B200_FE_*), degree, components */
int adaflo_b200_mf_space_info(adaflo_b200_mf_t mf, int space, int *fe_kind, int *degree, int *n_components);
int adaflo_b200_mf_n_cells(adaflo_b200_mf_t mf, uint32_t *n_cells);   /* locally owned cells */
/* LinearAlgebra::distributed::Vector::update_ghost_values / compress(add) */
int adaflo_b200_mf_update_ghost_values(adaflo_b200_mf_t mf, int space, double *vec_dev);
int adaflo_b200_mf_compress_add(adaflo_b200_mf_t mf, int space, double *vec_dev);
/* AffineConstraints::distribute for the general constraint rows of the space (hanging_row_ptr / cols / weights): every
 * constrained node gets sum_k weight_k * value(parent_k), per component -- what the reference calls after every update of a
 * solution vector (hanging_node_constraints_u.distribute, source/navier_stokes.cc:1252; constraints_u.distribute on the Newton
 * update, :643-644, where the Dirichlet entries of the update are zero: flagged parent components are read as they stand in
 * the vector).  No-op for spaces without rows. */
int adaflo_b200_mf_distribute_constraints(adaflo_b200_mf_t mf, int space, double *vec_dev);

/* ------------------------------------------------------- NavierStokesMatrix
 * Enumerations and PODs mirror adaflo::FlowParameters
 * (include/adaflo/parameters.h:54-86) and adaflo::TimeStepping
 * (include/adaflo/time_stepping.h:240-295). */
enum adaflo_b200_physical_type {
  ADAFLO_B200_INCOMPRESSIBLE = 0,
  ADAFLO_B200_INCOMPRESSIBLE_STATIONARY = 1,
  ADAFLO_B200_STOKES = 2
};
enum adaflo_b200_linearization {
  ADAFLO_B200_COUPLED_IMPLICIT_NEWTON = 0,
  ADAFLO_B200_COUPLED_IMPLICIT_PICARD = 1,
  ADAFLO_B200_COUPLED_VELOCITY_SEMI_IMPLICIT = 2,
  ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT = 3,
  ADAFLO_B200_PROJECTION = 4
};

typedef struct {
  int32_t physical_type;
  int32_t linearization;
  int32_t augmented_taylor_hood;
  int32_t reserved;
  double density;
  double viscosity;
  double density_diff;           /* variable coefficients are used iff density_diff != 0 || viscosity_diff != 0, */
  double viscosity_diff;         /* navier_stokes_matrix.cc:97-98 */
  double damping;                /* already sign-flipped as parameters.cc:466-467 does */
  double tau_grad_div;
  double beta_convective_term_momentum_balance;
} adaflo_b200_flow_params;

typedef struct {
  double weight;          /* TimeStepping::weight()         */
  double weight_old;      /* TimeStepping::weight_old()     */
  double weight_old_old;  /* TimeStepping::weight_old_old() */
  double tau1;            /* TimeStepping::tau1(); tau2 must be 0 (navier_stokes_matrix.cc:629) */
  double extrapolate_old;     /* factors of TimeStepping::extrapolate, time_stepping.h:290-295 */
  double extrapolate_old_old;
} adaflo_b200_time_weights;

/* NavierStokesMatrix(parameters, dof_index_u, dof_index_p, quad_index_u,
 * quad_index_p, ..) + initialize(matrix_free, time_stepping,
 * pressure_average_fix): navier_stokes_matrix.cc:43-60,85-169.  Allocates the
 * per-quadrature-point coefficient arrays and builds the pressure null-space
 * vectors (mode 0: constants; mode 1: DG0 constants). */
int adaflo_b200_ns_create(adaflo_b200_ctx_t ctx, adaflo_b200_mf_t mf, const adaflo_b200_flow_params *prm,
                          int dof_index_u, int dof_index_p, int quad_index_u, int quad_index_p,
                          int pressure_average_fix, adaflo_b200_ns_t *out);
int adaflo_b200_ns_destroy(adaflo_b200_ns_t ns); /* NavierStokesMatrix::clear, :173-187 */
int adaflo_b200_ns_set_flow_params(adaflo_b200_ns_t ns, const adaflo_b200_flow_params *prm);
int adaflo_b200_ns_set_time_weights(adaflo_b200_ns_t ns, const adaflo_b200_time_weights *w);
/* the references to solution_old / solution_old_old held by the constructor
 * (velocity blocks only are read, navier_stokes_matrix.cc:677,682) */
int adaflo_b200_ns_set_old_solutions(adaflo_b200_ns_t ns, const double *old_u_dev, const double *old_old_u_dev);

/* vmult(dst, src): navier_stokes_matrix.cc:221-262.  dst = A src including the
 * constrained-row fix-up and the pressure mean projection.  Ghost exchange is
 * done inside when a communicator was set. */
int adaflo_b200_ns_vmult(adaflo_b200_ns_t ns, double *dst_u, double *dst_p, const double *src_u, const double *src_p);
/* the algebraic transpose of vmult (no counterpart in the reference) */
int adaflo_b200_ns_Tvmult(adaflo_b200_ns_t ns, double *dst_u, double *dst_p, const double *src_u, const double *src_p);
/* residual(system_rhs, src, user_rhs): navier_stokes_matrix.cc:266-293.
 * system_rhs is read-modify-write (the caller preloads const_rhs, navier_stokes.cc:784);
 * writes the linearization point consumed by vmult. */
int adaflo_b200_ns_residual(adaflo_b200_ns_t ns, double *rhs_u, double *rhs_p, const double *src_u, const double *src_p,
                            const double *user_rhs_u, const double *user_rhs_p);
/* velocity_vmult: :337-382 ; divergence_vmult_add: :300-332 (adds, no zeroing);
 * pressure_{poisson,mass,convdiff}_vmult: :386-483 */
int adaflo_b200_ns_velocity_vmult(adaflo_b200_ns_t ns, double *dst_u, const double *src_u);
int adaflo_b200_ns_divergence_vmult_add(adaflo_b200_ns_t ns, double *dst_p, const double *src_u, int weight_by_viscosity);
int adaflo_b200_ns_pressure_poisson_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p);
int adaflo_b200_ns_pressure_mass_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p);
int adaflo_b200_ns_pressure_convdiff_vmult(adaflo_b200_ns_t ns, double *dst_p, const double *src_p);
/* diagonal of the velocity block / of the coupled operator ("diagonal
 * computation" has no reference routine beyond the pressure-mass row sum; this is
 * the exact diagonal of vmult, tested against unit-vector products) */
int adaflo_b200_ns_compute_diagonal(adaflo_b200_ns_t ns, double *diag_u, double *diag_p);
/* fix_linearization_point: :1144-1152 (snapshot used by velocity_vmult & co.) */
int adaflo_b200_ns_fix_linearization_point(adaflo_b200_ns_t ns);
/* apply_pressure_average_projection / apply_pressure_shift: :191-217 */
int adaflo_b200_ns_apply_pressure_average_projection(adaflo_b200_ns_t ns, double *p_dev);
int adaflo_b200_ns_apply_pressure_shift(adaflo_b200_ns_t ns, double shift, double *p_dev);
/* begin_densities / begin_viscosities / begin_damping_coeff /
 * begin_linearized_velocities (navier_stokes_matrix.h:162-178,293-395): raw
 * device pointers into the coefficient arrays other components write.
 * Layouts: rho/mu/damping [n_cells][n_q]; linearized [n_cells][12][n_q] with
 * rows 0..2 = u_d, row 3+3d+e = d u_d / d x_e. */
enum adaflo_b200_coefficient {
  ADAFLO_B200_COEF_DENSITY = 0,
  ADAFLO_B200_COEF_VISCOSITY = 1,
  ADAFLO_B200_COEF_DAMPING = 2,
  ADAFLO_B200_COEF_LINEARIZED_VELOCITY = 3
};
/* raw device storage in the decade-blocked order of csrc/common.cuh (coef_index; lin_index for the 12-row
 * linearisation array, whose rows are interleaved in pairs) */
int adaflo_b200_ns_coefficient_ptr(adaflo_b200_ns_t ns, int which, double **dev_ptr, uint64_t *n_doubles);
/* The same arrays in the reference's own storage order (navier_stokes_matrix.h:293-395: AlignedVector<
 * VectorizedArray<double>> indexed [cell batch][q], the linearisation point 12 consecutive VectorizedArrays per point):
 *   canonical[((batch * 27 + q) * n_rows + row) * lanes + lane],  q = qx + 3 qy + 9 qz,  n_rows = 1 or 12,
 * where slot batch * lanes + lane holds device cell slot_cell_host[slot] (identity if NULL; 0xffffffff = empty lane of a
 * partially filled batch).  `canonical` is a device pointer, or a host pointer if on_host != 0 (synchronous copy).  This is
 * what the external writers of the reference (level_set_okz.cc:351-364, navier_stokes.cc:1178-1196) and the C++ mirror's
 * begin_densities(macro_cell) & co. go through; the decade-blocked device order stays private to the library. */
int adaflo_b200_ns_set_coefficient_canonical(adaflo_b200_ns_t ns, int which, const double *canonical, int on_host, uint32_t n_batches,
                                             uint32_t lanes, const uint32_t *slot_cell_host);
int adaflo_b200_ns_get_coefficient_canonical(adaflo_b200_ns_t ns, int which, double *canonical, int on_host, uint32_t n_batches,
                                             uint32_t lanes, const uint32_t *slot_cell_host);
/* pressure null-space data built by initialize (:110-168): device pointers */
int adaflo_b200_ns_nullspace(adaflo_b200_ns_t ns, int mode, const double **modes, const double **weights, double *inverse_weight);
/* get_matvec_statistics: :1194-1206 (count, accumulated seconds; resets) */
int adaflo_b200_ns_get_matvec_statistics(adaflo_b200_ns_t ns, uint64_t *count, double *seconds);

/* Measurement hooks (bench.py): CUDA-event timing of the dominant cell kernel inside
 * vmult on the launching stream, and a process-wide count of kernels this library launched. */
int adaflo_b200_ns_set_kernel_timing(adaflo_b200_ns_t ns, int enable);
int adaflo_b200_ns_get_kernel_timing(adaflo_b200_ns_t ns, uint64_t *launches, double *seconds);
uint64_t adaflo_b200_kernel_launch_count(void);

/* End-to-end variant with HOST vectors (what a deal.II adaptor that keeps host
 * vectors authoritative calls, SURVEY 8(b) "mirror per call"): H2D of src, vmult,
 * D2H of dst on the context stream, synchronous on return. */
int adaflo_b200_ns_vmult_host(adaflo_b200_ns_t ns, double *dst_u_host, double *dst_p_host,
                              const double *src_u_host, const double *src_p_host);
/* Host-pointer entries of the sub-block operators, for the wrappers that hand them to Trilinos (ML / ILU) with Epetra
 * vectors: VelocityMatrix::Multiply (source/navier_stokes_preconditioner.cc:332-343), PressurePoissonMatrix::Multiply
 * (:421-433), PressureMassMatrix (:444-461).  Owned entries, upload + operator + download, synchronous. */
int adaflo_b200_ns_velocity_vmult_host(adaflo_b200_ns_t ns, double *dst_u_host, const double *src_u_host);
int adaflo_b200_ns_pressure_poisson_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host);
int adaflo_b200_ns_pressure_mass_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host);
int adaflo_b200_ns_pressure_convdiff_vmult_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_p_host);
int adaflo_b200_ns_divergence_vmult_add_host(adaflo_b200_ns_t ns, double *dst_p_host, const double *src_u_host, int weight_by_viscosity);

/* ------------------------------------------------------ level-set operators
 * The matrix-free operators of LevelSetOKZSolver{AdvanceConcentration, ComputeNormal,
 * ComputeCurvature, Reinitialization} on FE_Q_iso_Q1(n_sub)
 * with QIterated(QGauss(2), n_sub); all vectors are device vectors of the level-set space
 * (normal fields: three block pointers).  Every *_vmult zeroes dst, runs the cell loop and sets
 * the constrained rows to diag * src (e.g. level_set_okz_reinitialization.cc:227-230). */
typedef struct adaflo_b200_ls_s *adaflo_b200_ls_t;
typedef struct {
  int32_t dim;
  int32_t concentration_subdivisions; /* n_sub = FE_Q_iso_Q1 degree */
  int32_t convection_stabilization;   /* must be 0 (the reference default, parameters.cc:360-364) */
  int32_t reserved;
  double epsilon;                     /* parameters.epsilon */
  double epsilon_used;                /* epsilon scaled with the mesh size, as the solver computes it */
  double damping_scale_factor;        /* LevelSetOKZSolverComputeNormalParameter::damping_scale_factor (4) */
  double minimal_edge_length;
} adaflo_b200_ls_params;

int adaflo_b200_ls_create(adaflo_b200_ctx_t ctx, adaflo_b200_mf_t mf, const adaflo_b200_ls_params *prm, int dof_index_ls,
                          int dof_index_vel, int quad_index_ls, const double *cell_diameters_host, adaflo_b200_ls_t *out);
int adaflo_b200_ls_destroy(adaflo_b200_ls_t ls);
/* TimeStepping scalars: weight(), weight_old(), weight_old_old() (pass 0 unless BDF-2 and step_no > 1,
 * advance_concentration.cc:377-380), step_size() */
int adaflo_b200_ls_set_time(adaflo_b200_ls_t ls, double weight, double weight_old, double weight_old_old, double step_size);
/* DiagonalPreconditioner::get_vector() used on the constrained rows */
int adaflo_b200_ls_set_diagonal(adaflo_b200_ls_t ls, const double *diag_dev);
/* initialize_mass_matrix_diagonal, level_set_okz_preconditioner.h:31-76 */
int adaflo_b200_ls_mass_diagonal(adaflo_b200_ls_t ls, double *diag_dev);
/* advance_concentration_vmult :401-480 ; local_advance_concentration_rhs :288-397 (writes evaluated_convection) */
int adaflo_b200_ls_advance_concentration_vmult(adaflo_b200_ls_t ls, double *dst, const double *src);
int adaflo_b200_ls_advance_concentration_rhs(adaflo_b200_ls_t ls, double *rhs, const double *vel, const double *c, const double *c_old,
                                             const double *c_old_old);
/* compute_normal_vmult :160-183 ; local_compute_normal_rhs :123-156 */
int adaflo_b200_ls_compute_normal_vmult(adaflo_b200_ls_t ls, double *const dst[3], const double *const src[3]);
int adaflo_b200_ls_compute_normal_rhs(adaflo_b200_ls_t ls, double *const rhs[3], const double *level_set);
/* one block of compute_normal_vmult: mass + damping_scale_factor delta^2 Laplace, the operator of the reference's assembled
 * projection_matrix (level_set_okz_preconditioner.cc:61-180) which its curvature projection solves with as well */
int adaflo_b200_ls_projection_vmult(adaflo_b200_ls_t ls, double *dst, const double *src);
/* compute_curvature_vmult :263-304 ; local_compute_curvature_rhs :212-259 ; correction :362-377 */
int adaflo_b200_ls_compute_curvature_vmult(adaflo_b200_ls_t ls, double *dst, const double *src, int apply_diffusion);
int adaflo_b200_ls_compute_curvature_rhs(adaflo_b200_ls_t ls, double *rhs, const double *const normal[3]);
int adaflo_b200_ls_curvature_correction(adaflo_b200_ls_t ls, double *curvature, const double *level_set);
/* reinitialization_vmult :193-231 ; local_reinitialize_rhs :128-189 (freezes the normals when first_reinit_step) */
int adaflo_b200_ls_reinitialization_vmult(adaflo_b200_ls_t ls, double *dst, const double *src, int diffuse_only);
int adaflo_b200_ls_reinitialization_rhs(adaflo_b200_ls_t ls, double *rhs, const double *solution, const double *const normal[3],
                                        int diffuse_only, int first_reinit_step);
/* LevelSetOKZSolver::compute_heaviside, level_set_okz.cc:479-540 (dof-wise; a dof shared by several cells takes
 * the value of the last cell in the cell order of the uploaded lists, as the reference's cell loop does) */
/* On meshes with hanging nodes the reference stores a Heaviside value for every hanging dof as well (plain reads,
 * level_set_okz.cc:341,497-536); compute_heaviside keeps those inside the handle and compute_force uses them, so
 * compute_force must follow a compute_heaviside call there. */
int adaflo_b200_ls_compute_heaviside(adaflo_b200_ls_t ls, double *heaviside, const double *level_set);
/* LevelSetOKZSolver::local_compute_force, level_set_okz.cc:317-409: dst_u += (v, sigma kappa grad(H) - g rho e_z);
 * writes rho_q, mu_q of the operator when it uses variable coefficients.  interpolation_host: the
 * [8][n_ls_dofs_per_cell] matrix interpolation_concentration_pressure (level_set_base.cc:65-137) if
 * interpolate_grad_onto_pressure, else NULL */
int adaflo_b200_ls_compute_force(adaflo_b200_ls_t ls, adaflo_b200_ns_t ns, double *dst_u, const double *heaviside, const double *curvature,
                                 double surface_tension, double gravity, const double *interpolation_host);
/* evaluated_convection (which = 0) / evaluated_normal (which = 1): device arrays [cell][3 components][8 points of a sub-cell][n_sub^3 sub-cells] */
int adaflo_b200_ls_pointdata_ptr(adaflo_b200_ls_t ls, int which, double **dev_ptr, uint64_t *n_doubles);

/* Solver-level routines of the level-set classes: right-hand side + Krylov solve with the diagonal
 * preconditioner (ls_set_diagonal) + update, matrix-free in place of the reference's assembled projection matrix + ILU.
 * LevelSetOKZSolverComputeNormal::compute_normal (level_set_okz_compute_normal.cc:211-289): CG(4000, 1e-50, fast ? 1e-5 : 1e-7)
 * on compute_normal_vmult, initial guess = the normal field passed in */
int adaflo_b200_ls_compute_normal(adaflo_b200_ls_t ls, double *const normal[3], const double *level_set, int fast_computation, int *iterations);
/* LevelSetOKZSolverComputeCurvature::compute_curvature (level_set_okz_compute_curvature.cc:305-392): CG(2000, 1e-50, 1e-8);
 * projection_operator 0: compute_curvature_vmult(apply_diffusion), 1: the reference's projection_matrix operator */
int adaflo_b200_ls_compute_curvature(adaflo_b200_ls_t ls, double *curvature, const double *level_set, const double *const normal[3],
                                     int curvature_correction, int projection_operator, int *iterations);
/* LevelSetOKZSolverReinitialization::reinitialize (level_set_okz_reinitialization.cc:256-367): diff_steps (+3 if the
 * concentration range left [-1.02, 1.02]; range may be NULL) diffusive and stab_steps compressive pseudo time steps, each a
 * CG(2000, 1e-50, 1e-6) solve; the normal is recomputed (fast) at the first compressive step.  iterations: one count per step */
int adaflo_b200_ls_reinitialize(adaflo_b200_ls_t ls, double *solution, double *const normal[3], double dt, unsigned stab_steps, unsigned diff_steps,
                                const double *last_concentration_range, int *iterations, int *n_steps_done);
/* LevelSetOKZSolverAdvanceConcentration::advance_concentration (level_set_okz_advance_concentration.cc:503-693) after the
 * caller has advanced the time weights (ls_set_time) and applied Dirichlet values: rhs, BiCGStab(30, 0.05 tol_nl, 1e-8) with
 * the GMRES(3000) fallback, solution += increment.  Outputs: the numbers of the log line "advect [r0/its]" */
int adaflo_b200_ls_advance_concentration(adaflo_b200_ls_t ls, double *solution, const double *solution_old, const double *solution_old_old,
                                         const double *vel_solution, double tol_nl_iteration, double *initial_residual, int *iterations);

/* ------------------------------------------------- Krylov-step vector kernels
 * The deal.II vector operations the solvers call between two vmults
 *  and adaflo's DiagonalPreconditioner
 * (source/diagonal_preconditioner.cc:38-124), as bandwidth-bound fused kernels.
 * n counts doubles; results of reductions land in result_dev (device) and are
 * all-reduced over the communicator if one is set. */
/* device vector storage for host languages without a CUDA runtime binding
 * (LinearAlgebra::distributed::Vector::reinit / host<->device copies of the adaptor) */
int adaflo_b200_vec_alloc(adaflo_b200_ctx_t ctx, uint64_t n, double **dev_out);   /* zero initialised */
int adaflo_b200_vec_free(adaflo_b200_ctx_t ctx, double *dev);
int adaflo_b200_vec_upload(adaflo_b200_ctx_t ctx, double *dev, const double *host, uint64_t n);   /* synchronous */
int adaflo_b200_vec_download(adaflo_b200_ctx_t ctx, double *host, const double *dev, uint64_t n); /* synchronous */
int adaflo_b200_vec_set(adaflo_b200_ctx_t ctx, double *x, double value, uint64_t n);
int adaflo_b200_vec_axpy(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, uint64_t n);            /* y += a x */
int adaflo_b200_vec_sadd(adaflo_b200_ctx_t ctx, double *y, double s, double a, const double *x, uint64_t n);  /* y = s y + a x */
int adaflo_b200_vec_scale(adaflo_b200_ctx_t ctx, double *y, const double *scaling, uint64_t n);               /* y *= scaling */
int adaflo_b200_vec_dot(adaflo_b200_ctx_t ctx, const double *x, const double *y, uint64_t n, double *result_host);
/* k dots of one vector against k others in one pass (GMRES Gram-Schmidt) */
int adaflo_b200_vec_multi_dot(adaflo_b200_ctx_t ctx, const double *x, const double *const *ys, int k, uint64_t n, double *results_host);
/* y += sum_i a[i] xs[i] in one pass over y (GMRES: Gram-Schmidt update and solution update; deal.II does k axpys,
 * lac/solver_gmres.h) */
int adaflo_b200_vec_multi_axpy(adaflo_b200_ctx_t ctx, double *y, const double *a_host, const double *const *xs, int k, uint64_t n);
/* add_and_dot: y += a x ; returns y . z   (deal.II Vector::add_and_dot, used by SolverCG) */
int adaflo_b200_vec_add_and_dot(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, const double *z, uint64_t n, double *result_host);
/* DiagonalPreconditioner::reinit (guarded inverse, :38-60) and ::vmult (:99-124) */
int adaflo_b200_jacobi_reinit(adaflo_b200_ctx_t ctx, double *inv_diag, const double *diag, uint64_t n);
int adaflo_b200_jacobi_vmult(adaflo_b200_ctx_t ctx, double *dst, const double *src, const double *inv_diag, uint64_t n, int n_blocks);

/* ------------------------------------------------------------------ Krylov drivers on device vectors
 * The loops deal.II's SolverCG / SolverBicgstab / SolverFGMRES run around the operators:
 *   source/level_set_okz_compute_normal.cc:187-217, level_set_okz_compute_curvature.cc:300-330 (CG on the projections),
 *   source/level_set_okz_reinitialization.cc:334-336 (CG), source/level_set_okz_advance_concentration.cc:626-644 (BiCGStab),
 *   source/navier_stokes.cc:571-609 (FGMRES with the block preconditioner).
 * Operators and preconditioners are callbacks on device block vectors (dst/src: n_blocks device pointers); P may be
 * NULL (identity).  adaflo_b200_apply_operator is a ready-made callback for the library's own operators.
 * Stopping rule = ReductionControl: ||r|| <= max(abs_tol, rel_tol * ||r_0||); ERR_NO_CONVERGENCE after max_iter steps
 * (x holds the last iterate, as with deal.II's exception). */
typedef int (*adaflo_b200_apply_fn)(void *user, double *const *dst, const double *const *src);
typedef struct {
  int max_iter;
  double abs_tol, rel_tol;
  int last_step;               /* out */
  double initial_value;        /* out: ||r_0|| */
  double last_value;           /* out */
  double *history;             /* host array or NULL: ||r_0||, then the residual (GMRES: its estimate) after every step */
  int history_capacity;
  int history_size;            /* out */
} adaflo_b200_solver_control;
int adaflo_b200_solve_cg(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A, void *A_user,
                         adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b, adaflo_b200_solver_control *control);
int adaflo_b200_solve_bicgstab(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A,
                               void *A_user, adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b,
                               adaflo_b200_solver_control *control);
/* restarted, right-preconditioned, flexible (the preconditioned basis is stored) */
int adaflo_b200_solve_gmres(adaflo_b200_ctx_t ctx, int n_blocks, const uint64_t *n_owned, const uint64_t *n_local, adaflo_b200_apply_fn A, void *A_user,
                            adaflo_b200_apply_fn P, void *P_user, double *const *x, const double *const *b, int max_basis,
                            adaflo_b200_solver_control *control);

/* Block preconditioner of the linearised Navier-Stokes system on the device: NavierStokesPreconditioner::vmult
 * with inner solves (source/navier_stokes_preconditioner.cc:593-737), lower block-triangular, Cahouet-Chabard Schur complement
 * from pressure_mass_vmult / pressure_poisson_vmult; inner solves matrix-free with Jacobi preconditioners instead of the
 * reference's Trilinos ILU / AMG.  Build it after the linearisation point is set (it takes the diagonal of the velocity
 * block); use adaflo_b200_ns_preconditioner_apply as the P callback of adaflo_b200_solve_gmres. */
typedef struct adaflo_b200_precond_s *adaflo_b200_precond_t;
int adaflo_b200_ns_preconditioner_create(adaflo_b200_ns_t ns, adaflo_b200_precond_t *out);
int adaflo_b200_ns_preconditioner_destroy(adaflo_b200_precond_t precond);
int adaflo_b200_ns_preconditioner_vmult(adaflo_b200_precond_t precond, double *dst_u, double *dst_p, const double *src_u, const double *src_p);
int adaflo_b200_ns_preconditioner_apply(void *precond, double *const *dst, const double *const *src);
int adaflo_b200_ns_preconditioner_statistics(adaflo_b200_precond_t precond, int *n_applications, int inner_iterations[3]);

#define ADAFLO_B200_OP_NS_VMULT 0            /* handle: adaflo_b200_ns_t, 2 blocks (u, p) */
#define ADAFLO_B200_OP_NS_VELOCITY 1
#define ADAFLO_B200_OP_NS_PRESSURE_POISSON 2
#define ADAFLO_B200_OP_NS_PRESSURE_MASS 3
#define ADAFLO_B200_OP_LS_ADVECT 4           /* handle: adaflo_b200_ls_t */
#define ADAFLO_B200_OP_LS_NORMAL 5           /* 3 blocks */
#define ADAFLO_B200_OP_LS_CURVATURE 6        /* flag: apply_diffusion */
#define ADAFLO_B200_OP_LS_REINIT 7           /* flag: diffuse_only */
#define ADAFLO_B200_OP_LS_NORMAL_SCALAR 9    /* one block of compute_normal_vmult (the assembled projection_matrix of the reference) */
#define ADAFLO_B200_OP_JACOBI 8              /* handle: adaflo_b200_ctx_t; inv_diag (n entries) applied to each of n_blocks blocks */
typedef struct {
  int kind;
  void *handle;
  int flag;
  const double *inv_diag;
  uint64_t n;
  int n_blocks;
} adaflo_b200_operator_ref;
int adaflo_b200_apply_operator(void *operator_ref, double *const *dst, const double *const *src);

#ifdef __cplusplus
}
#endif
#endif /* ADAFLO_B200_H */
