// ls_ops.cuh -- argument block and launchers of the level-set kernels (ls_ops.cu)
#pragma once

#include "common.cuh"

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
  int geom;                  // cartesian / affine
  const double *cell_diameters;
  const double *src[3];
  double *dst[3];
  const double *aux[3];      // additional input vectors (old level sets / normal field)
  const double *vel;         // velocity vector (advect rhs)
  double *pointdata;         // conv or frozen normals [cell][sub][3][8]
  double gamma, w_old, w_oo; // time weights
  double delta_scale, eps_ratio, epsilon_used, dtau_inv;
  int first_step;
  int n_comp;
};

struct ForceArgs {
  NSArgs ns;                 // geometry, cell_u, hanging masks, dst_u
  const uint32_t *cell_ls;   // [n_cells][n1^3]
  int n_sub;
  const double *heaviside, *curvature;
  double *rho_q, *mu_q;      // decade blocked, may be null (constant coefficients)
  const double *interp;      // [8][n1^3] level set -> pressure interpolation or null
  double hat_v[3][5], hat_d[3][5];  // FE_Q_iso_Q1 hats and derivatives at the 3 Gauss points
  double surface_tension, gravity, density, density_diff, viscosity, viscosity_diff;
};

int launch_ls(const LSArgs &a, int op, cudaStream_t s);
int launch_ls_force(const ForceArgs &f, int geom, cudaStream_t s);
int launch_heaviside(const uint32_t *cell_ls, const uint32_t *last_cell, signed char *cell_state, const double *c, double *heaviside, int N,
                     uint32_t n_cells, uint64_t n_dofs, double scale, cudaStream_t s);
int launch_curvature_correction(double *curv, const double *ls, double eps_used, int dim, uint64_t n, cudaStream_t s);
int launch_scaled_copy_indexed(double *dst, const double *src, const double *diag, const uint32_t *idx, uint32_t n, cudaStream_t s);

}  // namespace ab200
