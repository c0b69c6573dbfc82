// common.cuh -- internal declarations shared by the translation units of
// libadaflo_b200.so.  Not part of the public ABI (include/adaflo_b200.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/adaflo_b200.h"

namespace ab200 {

// Axis roles of the vmult cell kernel (ns_r1.cuh).  Three lanes own one cell: in phases A/C lane t owns the node plane
// t of kernel axis Z, in phase B it owns the lines (X = j, Y = t) and contracts along Z.  The kernel axes are a cyclic
// rotation of the mesh axes, kernel (X, Y, Z) = mesh (y, z, x): the 30 lanes of a decade of x-neighbours then gather
// (and scatter) 21 consecutive nodes of one mesh line per instruction -- 16 instead of 30 sectors, 4.5 instead of 13
// L1 tag look-ups (profiles/r1_g: the gather and the scatter were two thirds of the global LSU wavefronts).  Any cell
// numbering is valid; only the coalescing depends on it.  AB_R1_ROT=0 restores kernel axes = mesh axes (A/B builds).
#ifndef AB_R1_ROT
#define AB_R1_ROT 0
#endif
constexpr int kAxX = AB_R1_ROT ? 1 : 0, kAxY = AB_R1_ROT ? 2 : 1, kAxZ = AB_R1_ROT ? 0 : 2;  // mesh direction of a kernel axis

// Per-quadrature-point arrays on QGauss(3) and the cell->node lists used by the vmult kernel
// are stored "decade blocked": cells are grouped in tens (one warp of the vmult kernel owns a
// decade, 3 lanes per cell) and the point with kernel indices (kx,ky,kz) of cell 10 g + cw, row r of n_rows, lives at
//   ((g * n_rows + r) * 9 + 3 kx + kz) * 30 + 3 cw + ky
// so that the 30 lanes of a warp read 240 contiguous bytes per (row, kx, kz).  All code outside ns_r1.cuh addresses
// these arrays through coef_index / lin_index with MESH point indices (qx,qy,qz).
constexpr int kGroup = 10;
__host__ __device__ inline size_t coef_slot(int qx, int qy, int qz, int cw)
{
  const int q[3] = {qx, qy, qz};
  return (size_t)(3 * q[kAxX] + q[kAxZ]) * 30 + 3 * cw + q[kAxY];
}
__host__ __device__ inline size_t coef_index(uint32_t cell, int row, int n_rows, int qx, int qy, int qz)
{
  const size_t g = cell / kGroup;
  const int cw = (int)(cell % kGroup);
  return (g * n_rows + row) * 270 + coef_slot(qx, qy, qz, cw);
}
// The 12-row linearisation array (u, grad u at every point) interleaves its rows in pairs, so that a lane of the
// vmult kernel fetches two rows with one 16-byte load (480 contiguous bytes per warp), and is POINT-major inside a
// decade: [g][point n = 3 kx + kz][row pair][30 lanes][2].  The six row pairs of one point are 2880 contiguous bytes:
// one bulk copy of the TMA engine stages the rows of one point of all 10 cells (ns_r2.cuh).  Row r of the point lives at
//   g * 3240 + n * 360 + (r / 2) * 60 + (3 cw + ky) * 2 + (r & 1)  =  lin_index(cell, qx, qy, qz) + lin_row(r)
constexpr int kLinPairStride = 60, kLinPointStride = 360, kLinDecade = 3240;  // doubles
__host__ __device__ inline size_t lin_index(uint32_t cell, int qx, int qy, int qz)
{
  const size_t g = cell / kGroup;
  const int cw = (int)(cell % kGroup);
  const size_t slot = coef_slot(qx, qy, qz, cw);  // n * 30 + lane
  return g * kLinDecade + (slot / 30) * kLinPointStride + (slot % 30) * 2;
}
__host__ __device__ constexpr int lin_row(int r) { return (r >> 1) * kLinPairStride + (r & 1); }
// cell-local (lexicographic, x fastest) index of the node with kernel indices (ix, iy, iz), n1 nodes per direction
__host__ __device__ inline int kernel_node(int ix, int iy, int iz, int n1)
{
  int m[3];
  m[kAxX] = ix;
  m[kAxY] = iy;
  m[kAxZ] = iz;
  return m[0] + n1 * (m[1] + n1 * m[2]);
}
inline size_t n_groups(uint32_t n_cells) { return ((size_t)n_cells + kGroup - 1) / kGroup; }

void set_error(const std::string &msg);
extern unsigned long long g_launches;  // kernels launched by this library (bench.py: gpu_launches)
#define AB_COUNT_LAUNCH() (++::ab200::g_launches)
int fail(int code, const std::string &msg);

#define AB_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return ::ab200::fail(ADAFLO_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define AB_CHECK(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != 0) return rc__; \
  } while (0)

// ------------------------------------------------------------------ handles
struct Exchange {
  int peer;
  uint32_t n_send;
  uint32_t *send_dev;  // local node indices
  uint32_t recv_first;
  uint32_t n_recv;
  double *send_buf;    // send_slots * n_send * n_comp (one slot per vector of this space in one exchange group)
  int send_slots;
  double *recv_buf;    // n_recv * n_comp (compress direction)
};

struct Space {
  int fe_kind, degree, n_comp;
  uint32_t n_owned_nodes, n_ghost_nodes, nodes_per_cell;
  uint32_t *cell_nodes;      // device [n_cells][nodes_per_cell]
  uint32_t *cell_nodes_blk;  // device, decade blocked, KERNEL node indices (kernel_node): u: [g][9 (ix+3iy)][10 cw][3 iz]; p: [g][4 (ix+2iy)][10 cw][3 iz (2 used)]
  uint32_t *hanging_masks;   // device [n_cells] or null
  uint32_t n_constrained;
  uint32_t *constrained;     // device
  // general constraint rows of the generic-degree path (adaflo_b200_space_desc::hanging_*): nodes >= hn_first are constrained
  uint32_t hn_first = 0xffffffffu;
  uint32_t *hn_row_ptr = nullptr, *hn_cols = nullptr;  // device
  double *hn_weights = nullptr;                        // device
  std::vector<Exchange> exchange;
  std::vector<uint32_t> dg0_dofs;  // host: element-constant dofs (FE_Q_DG0)
  // host-vector pipeline (vmult_host): per cell chunk, the node prefix the chunk needs as source, the node
  // prefix that is final after the chunk, and the number of constrained entries inside the final prefix
  std::vector<uint32_t> chunk_need_hi, chunk_final_lo, chunk_constrained_end;
  uint32_t send_lo = 0, send_hi = 0;  // node range covering all nodes this rank exports to its neighbours
  // dof numbering of the host application where it is not node-major (Cuthill-McKee on the velocity DoFHandler,
  // source/navier_stokes.cc:208-209): position of library entry i in the host's vector, [n_owned_entries], or null
  uint32_t *numbering = nullptr;
  double *numbering_stage = nullptr;  // device staging vector in host order
  uint64_t n_owned_entries() const { return (uint64_t)n_owned_nodes * n_comp; }
  uint64_t n_local_entries() const { return (uint64_t)(n_owned_nodes + n_ghost_nodes) * n_comp; }
};

struct Quad {
  int n_q_1d, geometry_type;
  uint32_t n_rows;
  double *table;       // device
  uint32_t *cell_row;  // device
  double *inv_jac;     // device; n_q_1d == 3: decade blocked with 9 rows, else [n_cells][9][n_q]
  double *jxw;         // device; n_q_1d == 3: decade blocked with 1 row, else [n_cells][n_q]
};

}  // namespace ab200

struct adaflo_b200_ctx_s {
  int device;
  cudaStream_t stream;
  cudaStream_t comm_stream;
  cudaEvent_t ev_a, ev_b;
  unsigned *work_counter;
  int sm_count;
  // reductions
  double *red_dev;     // scratch for block partials / results
  double *red_host;    // pinned
  size_t red_capacity; // doubles
  // communicator
  void *nccl_comm;
  int rank, nranks;
  // stall watchdog of the host-side waits (comm.cu stream_wait): seconds, 0 = plain cudaStreamSynchronize
  double comm_timeout_s;
  cudaEvent_t ev_watch;
};

struct adaflo_b200_mf_s {
  adaflo_b200_ctx_t ctx;
  int dim;
  uint32_t n_cells;
  std::vector<ab200::Space> spaces;
  std::vector<ab200::Quad> quads;
  // cells touching ghost nodes of any space, if they form one contiguous decade range [begin, end)
  // (true for contiguous partitions with lower-rank ownership); used to overlap the halo import
  size_t ghost_decade_begin = 0, ghost_decade_end = 0;
  bool ghost_range_valid = false;
  // host-vector pipeline: chunk k covers the decades [chunk_decade_end[k-1], chunk_decade_end[k])
  std::vector<size_t> chunk_decade_end;
  bool pipeline_ok = false;
};

struct adaflo_b200_ns_s {
  adaflo_b200_ctx_t ctx;
  adaflo_b200_mf_t mf;
  adaflo_b200_flow_params prm;
  adaflo_b200_time_weights tw;
  int dof_u, dof_p, quad_u, quad_p;
  int n_q;  // points of quad_u
  // degree / dimension dispatch (ns_generic.cu): anything but 3-D Q2/Q1 runs through the generic kernel with plain
  // per-point arrays [cell][row][q]
  bool generic;
  int gen_dim, gen_degree_u;
  // per-quadrature-point coefficient arrays (owned)
  double *lin, *rho, *mu, *damp;
  double *lin_prec, *rho_prec, *mu_prec, *damp_prec;  // fix_linearization_point snapshots
  const double *old_u, *old_old_u;
  // pressure null space
  double *modes[2], *mode_weights[2];
  double inv_weights[2];
  bool has_mode[2];
  // host staging for the *_host entry points
  double *stage_u[2], *stage_p[2];
  cudaStream_t up_stream, down_stream;      // host-vector pipeline (created on first use)
  std::vector<cudaEvent_t> *pipe_events;    // 2 per chunk: uploaded, computed
  double *tmp_p;  // scratch pressure vector (Tvmult)
  bool var;       // variable coefficients in use
  // statistics
  uint64_t matvec_count;
  double matvec_seconds;
  // optional CUDA-event timing of the cell kernel inside vmult
  bool time_kernel;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *kernel_events;
  size_t kernel_events_used;
  uint64_t kernel_launches;
  double kernel_seconds;
};

namespace ab200 {

// --------------------------------------------------------- kernel arguments
struct NSArgs {
  uint32_t n_cells;
  int np_cell;  // pressure dofs per cell: 8, or 9 with DG0
  const uint32_t *cell_u;
  const uint32_t *cell_p;
  const uint32_t *cell_u_blk;
  const uint32_t *cell_p_blk;
  const uint32_t *hn_u;  // per-cell refinement masks (null on conforming meshes)
  const uint32_t *hn_p;
  // generic-degree path: constraint rows (Space::hn_*); first = 0xffffffff on conforming meshes
  uint32_t gc_first_u, gc_first_p;
  const uint32_t *gc_ptr_u, *gc_cols_u, *gc_ptr_p, *gc_cols_p;
  const double *gc_w_u, *gc_w_p;
  const double *src_u;
  const double *src_p;
  const double *old_u;
  const double *old_old_u;
  double *dst_u;
  double *dst_p;
  double *lin;
  const double *rho;
  const double *mu;
  const double *damp;
  const uint32_t *geom_row;
  const double *geom_table;
  const double *inv_jac;
  const double *jxw;
  double gamma, w_old, w_oo, tau1, beta, rho0, mu0, damp0, tau_gd, ex_old, ex_oo;
  int linearization;  // runtime copy (residual kernel)
  int physical_type;
  int transpose;      // Tvmult
  unsigned *work_counter;  // ticket counter of the persistent cell kernel (per context)
  size_t decade_begin, decade_end;  // range of cell decades of this launch (end == 0: all); halo overlap
  int diag;           // compute_diagonal: unit sources per cell-local slot (velocity block)
  int skip_p_test;    // linearization == projection: pressure block not integrated (:902-907)
  double div_weight;  // generic divergence_vmult_add: -1 or -viscosity (:948-953)
  double scalar_c0;   // generic pressure mass / Poisson: constant coefficient
  // paced L2 prefetch of the source vector by the persistent vmult kernel (doubles per decade, length of src_u; 0: off)
  uint64_t pf_stride, pf_u_entries;
};

// mode of the convective block in a matrix-vector product
enum { MODE_STOKES = 0, MODE_NEWTON = 1, MODE_PICARD = 2, MODE_NOCONV = 3 };

int upload_shape_tables(cudaStream_t s);
int upload_owned(adaflo_b200_mf_t mf, int space, double *dev, const double *host, cudaStream_t s);   // host numbering -> library layout
int download_owned(adaflo_b200_mf_t mf, int space, double *host, const double *dev, cudaStream_t s);
void solver_pool_release(adaflo_b200_ctx_t ctx);  // solvers.cu: work vectors of the Krylov drivers
double *pool_get(adaflo_b200_ctx_t ctx, size_t doubles);  // per-context pool of device work vectors (null: out of memory)
void pool_put(adaflo_b200_ctx_t ctx, double *p);

// launchers (one translation unit per geometry type keeps compile times down)
int launch_ns_vmult_cartesian(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s);
int launch_ns_vmult_affine(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s);
int launch_ns_vmult_general(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s);
int launch_ns_residual(const NSArgs &a, int geom, bool var, cudaStream_t s);
int launch_ns_residual_r1(const NSArgs &a, int geom, bool var, cudaStream_t s);
int launch_ns_generic_scalar(const NSArgs &a, int dim, int degree_p, int kind, int geom, cudaStream_t s);  // ns_generic.cu: 0 mass, 1 Laplace
int launch_ns_generic(const NSArgs &a, int dim, int degree_u, int kind, int with_p, int mode, int geom, cudaStream_t s);  // ns_generic.cu  // Newton / Picard / Stokes, vmult mapping

struct PArgs {
  uint32_t n_cells;
  int np_cell;
  const uint32_t *cell_u;
  const uint32_t *cell_p;
  const uint32_t *hn_u;
  const uint32_t *hn_p;
  const double *src;
  double *dst;
  const double *coeff_q;   // per quadrature point [n_cells][n_q] or null
  int resolve;             // divergence: read_dof_values (1) or read_dof_values_plain (0)
  int coeff_mode;          // see pressure_ops.cu
  double c0, c1;
  const uint32_t *geom_row;
  const double *geom_table;
  const double *inv_jac;
  const double *jxw;
  int geom;
};
int launch_divergence(const PArgs &a, bool weight_by_viscosity, cudaStream_t s);
int launch_pressure_op(const PArgs &a, int kind, int n_q_1d, cudaStream_t s);

// vector kernels
int vec_set(cudaStream_t s, double *x, double v, uint64_t n);
int vec_axpy(cudaStream_t s, double *y, double a, const double *x, uint64_t n);
int vec_sadd(cudaStream_t s, double *y, double sc, double a, const double *x, uint64_t n);
int vec_scale(cudaStream_t s, double *y, const double *sc, uint64_t n);
int vec_copy_indexed(cudaStream_t s, double *dst, const double *src, const uint32_t *idx, uint32_t n, double sign);
int vec_set_indexed(cudaStream_t s, double *dst, const uint32_t *idx, uint32_t n, double value);
int vec_multi_axpy(adaflo_b200_ctx_t ctx, double *y, const double *a_host, const double *const *xs, int k, uint64_t n);
int vec_multi_dot(adaflo_b200_ctx_t ctx, const double *x, const double *const *ys, int k, uint64_t n, double *results_host,
                  bool to_host);
int vec_project(adaflo_b200_ctx_t ctx, double *p, const double *w, const double *mode, double inv_weight, uint64_t n_owned,
                uint64_t n_local);
int vec_pack(cudaStream_t s, double *buf, const double *vec, const uint32_t *idx, uint32_t n_nodes, int n_comp);
int vec_unpack_add(cudaStream_t s, double *vec, const double *buf, const uint32_t *idx, uint32_t n_nodes, int n_comp);
int vec_add_and_dot(adaflo_b200_ctx_t ctx, double *y, double a, const double *x, const double *z, uint64_t n, double *result_host);
int jacobi_reinit(adaflo_b200_ctx_t ctx, double *inv, const double *diag, uint64_t n);
int jacobi_vmult(cudaStream_t s, double *dst, const double *src, const double *inv, uint64_t n, int n_blocks);
int allreduce_sum_host(adaflo_b200_ctx_t ctx, double *vals, int k);  // via device buffer + NCCL when nranks > 1

// communication (comm.cu)
int comm_update_ghosts(adaflo_b200_mf_t mf, int space, double *vec);
int comm_compress_add(adaflo_b200_mf_t mf, int space, double *vec);
int comm_update_ghosts_n(adaflo_b200_mf_t mf, int n, const int *spaces, double *const *vecs);
int comm_compress_add_n(adaflo_b200_mf_t mf, int n, const int *spaces, double *const *vecs);
int comm_allreduce(adaflo_b200_ctx_t ctx, double *dev, int k, int op_max);
int comm_unique_id(unsigned char id[128]);
int comm_init(adaflo_b200_ctx_t ctx, const unsigned char id[128], int rank, int nranks);
int comm_destroy(adaflo_b200_ctx_t ctx);
// Host wait for everything queued on `s` (a stream of this context).  With a communicator (or ADAFLO_B200_COMM_TIMEOUT set) the wait
// is a watchdog: it polls an event, checks the communicator's asynchronous error state and gives up after comm_timeout_s seconds
// with ADAFLO_B200_ERR_COMM (the communicator is aborted) -- a rank that issued a different number of collective-bearing calls
// than its peers then fails loudly instead of hanging (MatrixFree::cell_loop is collective, navier_stokes_matrix.cc:236-242).
int stream_wait(adaflo_b200_ctx_t ctx, cudaStream_t s);

}  // namespace ab200
