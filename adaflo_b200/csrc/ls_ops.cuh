// ls_ops.cuh -- argument block and launchers of the level-set kernels (ls_ops.cu)
#pragma once

#include "common.cuh"
#include "hanging.cuh"

namespace ab200 {

enum { LS_ADVECT = 0, LS_DAMPED = 1, LS_REINIT = 2, LS_REINIT_DIFFUSE = 3, LS_NORMAL_RHS = 4, LS_CURV_RHS = 5, LS_REINIT_RHS = 6,
       LS_REINIT_RHS_DIFFUSE = 7, LS_ADVECT_RHS = 8, LS_MASS_DIAG = 9 };

struct LSArgs {
  uint32_t n_cells;
  int n_sub;
  const uint32_t *cell_ls;   // [n_cells][(n_sub+1)^3], flag bit 29 = constrained
  const uint32_t *cell_u;    // [n_cells][27] (advect rhs)
  const uint32_t *hn_u;      // [n_cells] hanging-node masks of the velocity space or null
  const uint32_t *hn;        // per-cell refinement masks of the level-set space or null
  const uint32_t *geom_row;
  const double *geom_table;
  int geom;                  // cartesian / affine / general
  const double *inv_jac;     // general: [n_cells][9][NQ], q = qx + nq1 (qy + nq1 qz) on QIterated(QGauss(2), n_sub)
  const double *jxw_q;       // general: [n_cells][NQ]
  const double *cell_diameters;
  const double *src[3];
  double *dst[3];
  const double *aux[3];      // additional input vectors (old level sets / normal field)
  const double *vel;         // velocity vector (advect rhs)
  double *pointdata;         // conv or frozen normals [cell][3][8 points][sub-cell]
  double gamma, w_old, w_oo; // time weights
  double delta_scale, eps_ratio, epsilon_used, dtau_inv;
  int first_step;
  int n_comp;
};

// Hanging nodes of FE_Q_iso_Q1(S): a child node m in direction d sits at the parent coordinate (sub S + m) / 2;
// if that is an integer it coincides with a parent node, otherwise it is the mean of its two neighbours (hats).
struct LsStencil {
  int n;
  uint32_t slot[8];
  double w[8];
};
__host__ __device__ __forceinline__ void ls_stencil(const HnInfo hn, const uint32_t *cl, int S, int ix, int iy, int iz, LsStencil &st)
{
  const int n1 = S + 1;
  const int idx[3] = {ix, iy, iz};
  int i0[3], cnt[3];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int d = 0; d < 3; ++d) {
    const int e = (d + 1) % 3, f = (d + 2) % 3;
    i0[d] = idx[d];
    cnt[d] = 1;
    if (hn.line(d, idx[e], idx[f], n1)) {
      const int c2 = hn.sub(d) * S + idx[d];
      i0[d] = c2 >> 1;
      cnt[d] = (c2 & 1) ? 2 : 1;
    }
  }
  st.n = 0;
  for (int c = 0; c < cnt[2]; ++c)
    for (int b = 0; b < cnt[1]; ++b)
      for (int a = 0; a < cnt[0]; ++a) {
        st.slot[st.n] = cl[(i0[0] + a) + n1 * ((i0[1] + b) + n1 * (i0[2] + c))];
        st.w[st.n] = 1.0 / (cnt[0] * cnt[1] * cnt[2]);
        ++st.n;
      }
}

// Level-set spaces with hanging nodes, operations that read dofs PLAINLY (compute_heaviside / local_compute_force,
// level_set_okz.cc:341,497): deal.II stores a value for every hanging dof; here such dofs get ids n_real + h and
// every (cell, slot) an "extended" index (the real dof it coincides with, or the hanging id), built in ls_create.
struct LsHanging {
  uint32_t n_real, n_hang;
  const uint32_t *cell_ext;   // [n_cells][N] extended index of every slot, or null on conforming meshes
  const uint32_t *hang_idx;   // [n_hang][8] real dofs a hanging dof is the mean of
  const uint32_t *hang_n;     // [n_hang] number of those (2, 4 or 8)
};

struct ForceArgs {
  NSArgs ns;                 // geometry, cell_u, hanging masks, dst_u
  const uint32_t *cell_ls;   // [n_cells][n1^3]
  int n_sub;
  const double *heaviside, *curvature;
  LsHanging hang;                         // hanging-dof extension (cell_ext == null: conforming)
  const double *heaviside_hang, *curvature_hang;  // [n_hang] values of the hanging dofs
  double *rho_q, *mu_q;      // decade blocked, may be null (constant coefficients)
  const double *interp;      // [8][n1^3] level set -> pressure interpolation or null
  double hat_v[3][5], hat_d[3][5];  // FE_Q_iso_Q1 hats and derivatives at the 3 Gauss points
  double surface_tension, gravity, density, density_diff, viscosity, viscosity_diff;
  int gravity_dir;           // dim - 1 (2-D problems on the extrusion: 1)
};

int launch_ls(const LSArgs &a, int op, cudaStream_t s);
int launch_ls_force(const ForceArgs &f, int geom, cudaStream_t s);
int launch_heaviside(const uint32_t *cell_ls, const uint32_t *last_cell, signed char *cell_state, const double *c, double *heaviside, int N,
                     uint32_t n_cells, uint64_t n_dofs, double scale, const LsHanging &hang, double *c_hang, double *heaviside_hang,
                     cudaStream_t s);
int launch_hanging_mean(const LsHanging &hang, const double *vec, double *out, cudaStream_t s);
int launch_curvature_correction(double *curv, const double *ls, double eps_used, int dim, uint64_t n, cudaStream_t s);
int launch_scaled_copy_indexed(double *dst, const double *src, const double *diag, const uint32_t *idx, uint32_t n, cudaStream_t s);

}  // namespace ab200
