/* ls_oracle.c -- CPU restatement of adaflo's matrix-free level-set operators (dense loops).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as ns_oracle.c).  Restates, per quadrature point, the
 * weak forms of
 *   source/level_set_okz_advance_concentration.cc:217-258 (vmult), :288-397 (rhs)
 *   source/level_set_okz_compute_normal.cc:82-119 (vmult), :123-156 (rhs)
 *   source/level_set_okz_compute_curvature.cc:86-133 (vmult), :212-259 (rhs), :362-377 (correction)
 *   source/level_set_okz_reinitialization.cc:53-106 (vmult), :128-189 (rhs)
 *   include/adaflo/level_set_okz_preconditioner.h:31-76 (mass diagonal)
 * on FE_Q_iso_Q1(n_sub) with QIterated(QGauss(2), n_sub) (source/two_phase_base.cc:256-271).
 * Parity status: the reference holds no operator-level vectors for these operators, but its golden log
 * tests/rising_bubble_ls.output prints the CG iteration counts of the reinitialisation ("reinitialize (8 + 8)",
 * "(7 + 7)"), which are reproduced by the 2-D restatement oracle/ls_oracle_2d.py (normal projection, reinitialisation
 * operator and right-hand side, mass diagonal; tests/test_2d_cpu.py).  This file is tied to that restatement by the
 * extrusion identity A_2d = R A_3d E, checked operator by operator in the same test file (1e-13): pinned through the
 * 2-D logs.  The curvature, Heaviside and force terms are pinned the same way by the first Navier-Stokes residuals of
 * tests/rising_bubble_ls.output:14 (0.0198) and tests/spurious_currents_ls.output:10 (0.365).  The advection terms
 * have no first-party numbers; they are cross-checked against the independent dense loops of
 * source/level_set_okz_matrix.cc (same terms).
 * Accumulation per cell in long double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef long double real;
#define MAXSUB 4
#define MAXN1 (MAXSUB + 1)
#define MAXQ1 (2 * MAXSUB)

typedef struct {
  int32_t n_cells, n_sub;
  const int64_t *cell_nodes;    /* [n_cells][(n_sub+1)^3] */
  const uint8_t *constrained;   /* per dof: read as 0 (resolved reads), never written */
  const double *inv_jac;        /* [n_cells][NQ][3][3]  K[f][e] = d xi_f / d x_e */
  const double *jxw;            /* [n_cells][NQ] */
  const double *cell_diameters; /* [n_cells] */
  /* optional plain constraint expansion (hanging nodes): local value = sum_k w[k] vec[idx[k]], rows = cell * N + n */
  const int64_t *ptr, *idx;
  const double *w;
} ls_mesh;

typedef struct {
  double epsilon, epsilon_used, damping_scale_factor, minimal_edge_length, step_size;
  int32_t dim, pad_;
} ls_params;

typedef struct {
  int n1, nq1, N, NQ;
  real V[MAXQ1][MAXN1], G[MAXQ1][MAXN1];
} ls_table;

/* FE_Q_iso_Q1(n_sub): hats on n_sub sub-intervals; QIterated(QGauss(2), n_sub) */
static void build(ls_table *t, int n_sub)
{
  const real g0 = 0.5L - 0.5L / sqrtl(3.0L), g1 = 0.5L + 0.5L / sqrtl(3.0L);
  t->n1 = n_sub + 1; t->nq1 = 2 * n_sub; t->N = t->n1 * t->n1 * t->n1; t->NQ = t->nq1 * t->nq1 * t->nq1;
  for (int q = 0; q < t->nq1; ++q) {
    for (int i = 0; i < t->n1; ++i) t->V[q][i] = t->G[q][i] = 0;
    const int s = q / 2;
    const real xi = (q % 2) ? g1 : g0;
    t->V[q][s] = 1 - xi; t->V[q][s + 1] = xi;
    t->G[q][s] = -(real)n_sub; t->G[q][s + 1] = (real)n_sub;
  }
}

static inline real phi(const ls_table *t, int q, int n) { const int a = t->n1, b = t->nq1;
  return t->V[q % b][n % a] * t->V[(q / b) % b][(n / a) % a] * t->V[q / (b * b)][n / (a * a)]; }
static inline void gphi(const ls_table *t, int q, int n, real g[3]) { const int a = t->n1, b = t->nq1;
  const int qx = q % b, qy = (q / b) % b, qz = q / (b * b), ix = n % a, iy = (n / a) % a, iz = n / (a * a);
  g[0] = t->G[qx][ix] * t->V[qy][iy] * t->V[qz][iz];
  g[1] = t->V[qx][ix] * t->G[qy][iy] * t->V[qz][iz];
  g[2] = t->V[qx][ix] * t->V[qy][iy] * t->G[qz][iz]; }

static void gather(const ls_mesh *m, const ls_table *t, int cell, const double *vec, int resolve, real *loc)
{
  for (int n = 0; n < t->N; ++n) {
    if (m->ptr) {
      real s = 0;
      for (int64_t k = m->ptr[(int64_t)cell * t->N + n]; k < m->ptr[(int64_t)cell * t->N + n + 1]; ++k)
        if (!(resolve && m->constrained && m->constrained[m->idx[k]])) s += (real)m->w[k] * (real)vec[m->idx[k]];
      loc[n] = s;
      continue;
    }
    const int64_t d = m->cell_nodes[(int64_t)cell * t->N + n];
    loc[n] = (resolve && m->constrained && m->constrained[d]) ? 0 : (real)vec[d];
  }
}
static void scatter(const ls_mesh *m, const ls_table *t, int cell, const real *out, double *vec)
{
  for (int n = 0; n < t->N; ++n) {
    if (m->ptr) {
      for (int64_t k = m->ptr[(int64_t)cell * t->N + n]; k < m->ptr[(int64_t)cell * t->N + n + 1]; ++k)
        if (!(m->constrained && m->constrained[m->idx[k]])) vec[m->idx[k]] += (double)((real)m->w[k] * out[n]);
      continue;
    }
    const int64_t d = m->cell_nodes[(int64_t)cell * t->N + n];
    if (m->constrained && m->constrained[d]) continue;
    vec[d] += (double)out[n];
  }
}
/* value and real-space gradient of a cell-local scalar field at point q */
static void eval(const ls_mesh *m, const ls_table *t, int cell, int q, const real *loc, real *val, real grad[3])
{
  const double *K = m->inv_jac + ((int64_t)cell * t->NQ + q) * 9;
  real v = 0, gr[3] = {0, 0, 0};
  for (int n = 0; n < t->N; ++n) {
    real g[3];
    v += phi(t, q, n) * loc[n];
    gphi(t, q, n, g);
    for (int f = 0; f < 3; ++f) gr[f] += g[f] * loc[n];
  }
  *val = v;
  for (int e = 0; e < 3; ++e) grad[e] = gr[0] * (real)K[e] + gr[1] * (real)K[3 + e] + gr[2] * (real)K[6 + e];
}
/* out[n] += phi_n tv JxW + grad phi_n . tg JxW */
static void integ(const ls_mesh *m, const ls_table *t, int cell, int q, real tv, const real tg[3], real *out)
{
  const double *K = m->inv_jac + ((int64_t)cell * t->NQ + q) * 9;
  const real JxW = m->jxw[(int64_t)cell * t->NQ + q];
  real r[3];
  for (int f = 0; f < 3; ++f) r[f] = ((real)K[3 * f] * tg[0] + (real)K[3 * f + 1] * tg[1] + (real)K[3 * f + 2] * tg[2]) * JxW;
  for (int n = 0; n < t->N; ++n) {
    real g[3];
    gphi(t, q, n, g);
    out[n] += phi(t, q, n) * tv * JxW + g[0] * r[0] + g[1] * r[1] + g[2] * r[2];
  }
}
static real rmax(real a, real b) { return a > b ? a : b; }

/* (phi, gamma c + u_q . grad c): advance_concentration.cc:242-256 (convection_stabilization off) */
int ls_oracle_advect(const ls_mesh *m, double gamma, const double *conv /*[n_cells][NQ][3]*/, const double *src, double *dst)
{
  ls_table t; build(&t, m->n_sub);
  real *loc = malloc(sizeof(real) * t.N), *out = malloc(sizeof(real) * t.N);
  for (int c = 0; c < m->n_cells; ++c) {
    gather(m, &t, c, src, 1, loc);
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    for (int q = 0; q < t.NQ; ++q) {
      real v, g[3], z[3] = {0, 0, 0};
      eval(m, &t, c, q, loc, &v, g);
      const double *u = conv + ((int64_t)c * t.NQ + q) * 3;
      integ(m, &t, c, q, v * gamma + g[0] * u[0] + g[1] * u[1] + g[2] * u[2], z, out);
    }
    scatter(m, &t, c, out, dst);
  }
  free(loc); free(out);
  return 0;
}

/* Q2 velocity at the level-set quadrature points + BDF residual: advance_concentration.cc:374-395.
 * cell_u: [n_cells][27] velocity node ids (vector entry = 3 node + comp); w_oo = 0 unless BDF-2 and step > 1 */
int ls_oracle_advect_rhs(const ls_mesh *m, const int64_t *cell_u, const double *vel, const double *c, const double *c_old,
                         const double *c_oo, double gamma, double w_old, double w_oo, double *dst, double *conv_out)
{
  ls_table t; build(&t, m->n_sub);
  const real g0 = 0.5L - 0.5L / sqrtl(3.0L), g1 = 0.5L + 0.5L / sqrtl(3.0L);
  real L2[MAXQ1][3]; /* Q2 Lagrange basis on {0,1/2,1} at the iterated Gauss points */
  for (int q = 0; q < t.nq1; ++q) {
    const real x = ((q / 2) + ((q % 2) ? g1 : g0)) / m->n_sub;
    L2[q][0] = 2 * (x - 0.5L) * (x - 1); L2[q][1] = 4 * x * (1 - x); L2[q][2] = 2 * x * (x - 0.5L);
  }
  real *lc = malloc(sizeof(real) * t.N), *lo = malloc(sizeof(real) * t.N), *loo = malloc(sizeof(real) * t.N), *out = malloc(sizeof(real) * t.N);
  for (int cell = 0; cell < m->n_cells; ++cell) {
    gather(m, &t, cell, c, 0, lc); gather(m, &t, cell, c_old, 0, lo); gather(m, &t, cell, c_oo, 0, loo);
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    for (int q = 0; q < t.NQ; ++q) {
      const int qx = q % t.nq1, qy = (q / t.nq1) % t.nq1, qz = q / (t.nq1 * t.nq1);
      real u[3] = {0, 0, 0};
      for (int n = 0; n < 27; ++n) {
        const real w = L2[qx][n % 3] * L2[qy][(n / 3) % 3] * L2[qz][n / 9];
        for (int d = 0; d < 3; ++d) u[d] += w * (real)vel[3 * cell_u[(int64_t)cell * 27 + n] + d];
      }
      real v, g[3], vo, go[3], voo, goo[3], z[3] = {0, 0, 0};
      eval(m, &t, cell, q, lc, &v, g); eval(m, &t, cell, q, lo, &vo, go); eval(m, &t, cell, q, loo, &voo, goo);
      const real res = -(v * gamma + u[0] * g[0] + u[1] * g[1] + u[2] * g[2] + w_old * vo + w_oo * voo);
      integ(m, &t, cell, q, res, z, out);
      for (int d = 0; d < 3; ++d) conv_out[((int64_t)cell * t.NQ + q) * 3 + d] = (double)u[d];
    }
    scatter(m, &t, cell, out, dst);
  }
  free(lc); free(lo); free(loo); free(out);
  return 0;
}

/* generic damped projection / reinitialisation operator on a scalar field:
 * kind 0: (phi, s) + delta (grad phi, grad s)                         normal (per component) / curvature vmult
 * kind 1: (phi, dtau_inv s) + (grad phi, D (n.grad s) n)               reinit vmult
 * kind 2: (phi, dtau_inv s) + (grad phi, D grad s)                     reinit vmult, diffuse_only  */
static real delta_of(const ls_mesh *m, const ls_params *p, int cell, real scale)
{
  const real d = rmax((real)p->epsilon_used / (real)p->epsilon, (real)m->cell_diameters[cell] / m->n_sub);
  return scale * d * d;
}
int ls_oracle_scalar_op(const ls_mesh *m, const ls_params *p, int kind, double scale_or_flag, const double *normal_q, const double *src, double *dst)
{
  ls_table t; build(&t, m->n_sub);
  real *loc = malloc(sizeof(real) * t.N), *out = malloc(sizeof(real) * t.N);
  const real dtau_inv = rmax(0.95L / (1.0L / (p->dim * p->dim) * (real)p->minimal_edge_length / m->n_sub), 1.0L / (5.0L * (real)p->step_size));
  for (int c = 0; c < m->n_cells; ++c) {
    gather(m, &t, c, src, 1, loc);
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    const real delta = delta_of(m, p, c, (real)scale_or_flag);
    const real D = rmax((real)p->epsilon_used, (real)m->cell_diameters[c] / m->n_sub);
    for (int q = 0; q < t.NQ; ++q) {
      real v, g[3], tg[3];
      eval(m, &t, c, q, loc, &v, g);
      if (kind == 0) {
        for (int e = 0; e < 3; ++e) tg[e] = delta * g[e];
        integ(m, &t, c, q, v, tg, out);
      } else if (kind == 1) {
        const double *n = normal_q + ((int64_t)c * t.NQ + q) * 3;
        const real ng = n[0] * g[0] + n[1] * g[1] + n[2] * g[2];
        for (int e = 0; e < 3; ++e) tg[e] = D * ng * (real)n[e];
        integ(m, &t, c, q, dtau_inv * v, tg, out);
      } else {
        for (int e = 0; e < 3; ++e) tg[e] = D * g[e];
        integ(m, &t, c, q, dtau_inv * v, tg, out);
      }
    }
    scatter(m, &t, c, out, dst);
  }
  free(loc); free(out);
  return 0;
}

/* normal rhs: (phi, grad c) per component, compute_normal.cc:142-155; dst = 3 blocks of n_dofs */
int ls_oracle_normal_rhs(const ls_mesh *m, const double *c, double *dst0, double *dst1, double *dst2)
{
  ls_table t; build(&t, m->n_sub);
  real *loc = malloc(sizeof(real) * t.N), *out = malloc(sizeof(real) * t.N * 3);
  double *dst[3] = {dst0, dst1, dst2};
  for (int cell = 0; cell < m->n_cells; ++cell) {
    gather(m, &t, cell, c, 0, loc);
    for (int n = 0; n < 3 * t.N; ++n) out[n] = 0;
    for (int q = 0; q < t.NQ; ++q) {
      real v, g[3], z[3] = {0, 0, 0};
      eval(m, &t, cell, q, loc, &v, g);
      for (int d = 0; d < 3; ++d) integ(m, &t, cell, q, g[d], z, out + d * t.N);
    }
    for (int d = 0; d < 3; ++d) scatter(m, &t, cell, out + d * t.N, dst[d]);
  }
  free(loc); free(out);
  return 0;
}

/* curvature rhs: (phi, -div(n/|n|)) with dof-level normalisation, compute_curvature.cc:231-258 */
int ls_oracle_curvature_rhs(const ls_mesh *m, const double *n0, const double *n1, const double *n2, double *dst)
{
  ls_table t; build(&t, m->n_sub);
  const double *nb[3] = {n0, n1, n2};
  real *loc = malloc(sizeof(real) * t.N * 3), *out = malloc(sizeof(real) * t.N);
  for (int cell = 0; cell < m->n_cells; ++cell) {
    for (int d = 0; d < 3; ++d) gather(m, &t, cell, nb[d], 0, loc + d * t.N);
    int all_zero = 1;
    for (int n = 0; n < t.N; ++n) {
      const real nn = sqrtl(loc[n] * loc[n] + loc[t.N + n] * loc[t.N + n] + loc[2 * t.N + n] * loc[2 * t.N + n]);
      if (nn > 1e-2L) { all_zero = 0; for (int d = 0; d < 3; ++d) loc[d * t.N + n] /= nn; }
      else for (int d = 0; d < 3; ++d) loc[d * t.N + n] = 0;
    }
    if (all_zero) continue;
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    for (int q = 0; q < t.NQ; ++q) {
      real div = 0, z[3] = {0, 0, 0};
      for (int d = 0; d < 3; ++d) { real v, g[3]; eval(m, &t, cell, q, loc + d * t.N, &v, g); div += g[d]; }
      integ(m, &t, cell, q, -div, z, out);
    }
    scatter(m, &t, cell, out, dst);
  }
  free(loc); free(out);
  return 0;
}

/* reinit rhs, reinitialization.cc:163-184.  normal_q [n_cells][NQ][3] is written when first_step, read otherwise */
int ls_oracle_reinit_rhs(const ls_mesh *m, const ls_params *p, int diffuse_only, int first_step, const double *c, const double *n0,
                         const double *n1, const double *n2, double *normal_q, double *dst)
{
  ls_table t; build(&t, m->n_sub);
  const double *nb[3] = {n0, n1, n2};
  real *loc = malloc(sizeof(real) * t.N), *ln = malloc(sizeof(real) * t.N * 3), *out = malloc(sizeof(real) * t.N);
  for (int cell = 0; cell < m->n_cells; ++cell) {
    gather(m, &t, cell, c, 0, loc);
    if (!diffuse_only) for (int d = 0; d < 3; ++d) gather(m, &t, cell, nb[d], 0, ln + d * t.N);
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    const real D = rmax((real)p->epsilon_used, (real)m->cell_diameters[cell] / m->n_sub);
    for (int q = 0; q < t.NQ; ++q) {
      real v, g[3], tg[3];
      eval(m, &t, cell, q, loc, &v, g);
      if (!diffuse_only) {
        double *nq = normal_q + ((int64_t)cell * t.NQ + q) * 3;
        if (first_step) {
          real nv[3], dummy[3];
          for (int d = 0; d < 3; ++d) eval(m, &t, cell, q, ln + d * t.N, &nv[d], dummy);
          const real nn = rmax(1e-4L, sqrtl(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]));
          for (int d = 0; d < 3; ++d) nq[d] = (double)(nv[d] / nn);
        }
        const real ng = nq[0] * g[0] + nq[1] * g[1] + nq[2] * g[2];
        const real f = 0.5L * (1 - v * v) - ng * D;
        for (int e = 0; e < 3; ++e) tg[e] = (real)nq[e] * f;
      } else
        for (int e = 0; e < 3; ++e) tg[e] = -D * g[e];
      integ(m, &t, cell, q, 0, tg, out);
    }
    scatter(m, &t, cell, out, dst);
  }
  free(loc); free(ln); free(out);
  return 0;
}

/* d_i = sum_K sum_q phi_i^2 JxW, level_set_okz_preconditioner.h:31-76 */
int ls_oracle_mass_diagonal(const ls_mesh *m, double *dst)
{
  ls_table t; build(&t, m->n_sub);
  real *out = malloc(sizeof(real) * t.N);
  for (int c = 0; c < m->n_cells; ++c) {
    for (int n = 0; n < t.N; ++n) out[n] = 0;
    for (int q = 0; q < t.NQ; ++q)
      for (int n = 0; n < t.N; ++n) out[n] += phi(&t, q, n) * phi(&t, q, n) * (real)m->jxw[(int64_t)c * t.NQ + q];
    {  /* hanging-node constraints only (no Dirichlet skipping) */
      ls_mesh m2 = *m;
      m2.constrained = 0;
      scatter(&m2, &t, c, out, dst);
    }
  }
  free(out);
  return 0;
}

/* curvature correction, compute_curvature.cc:362-377 */
void ls_oracle_curvature_correction(int64_t n, double epsilon_used, int dim, const double *level_set, double *curvature)
{
  for (int64_t i = 0; i < n; ++i)
    if (curvature[i] > 1e-4) {
      const double c = level_set[i];
      const double distance = (1 - c * c) > 1e-2 ? epsilon_used * log((1. + c) / (1. - c)) : 0;
      curvature[i] = 1. / (1. / curvature[i] + distance / (dim - 1));
    }
}

/* LevelSetOKZSolver::local_compute_force, level_set_okz.cc:317-409, dense.  Geometry arrays on QGauss(3):
 * inv_jac [n_cells][27][9], jxw [n_cells][27]; cell_u [n_cells][27] velocity nodes (entry = 3 node + comp);
 * interp: [8][N] level set -> pressure interpolation or NULL. rho_q / mu_q [n_cells][27] written if variable. */
int ls_oracle_force(int n_cells, int n_sub, const int64_t *cell_ls, const int64_t *cell_u, const uint8_t *u_constrained,
                    const double *inv_jac, const double *jxw, const double *heaviside, const double *curvature, double sigma,
                    double gravity, double rho0, double drho, double mu0, double dmu, int variable, const double *interp,
                    double *dst_u, double *rho_q, double *mu_q)
{
  const int n1 = n_sub + 1, N = n1 * n1 * n1;
  const real s35 = sqrtl(0.6L);
  const real gp[3] = {0.5L - 0.5L * s35, 0.5L, 0.5L + 0.5L * s35};
  real hv[3][MAXN1], hd[3][MAXN1], L2[3][3];
  for (int p = 0; p < 3; ++p) {
    for (int i = 0; i < n1; ++i) hv[p][i] = hd[p][i] = 0;
    int s = (int)(gp[p] * n_sub);
    if (s > n_sub - 1) s = n_sub - 1;
    const real xi = gp[p] * n_sub - s;
    hv[p][s] = 1 - xi; hv[p][s + 1] = xi; hd[p][s] = -(real)n_sub; hd[p][s + 1] = (real)n_sub;
    const real x = gp[p];
    L2[p][0] = 2 * (x - 0.5L) * (x - 1); L2[p][1] = 4 * x * (1 - x); L2[p][2] = 2 * x * (x - 0.5L);
  }
  real *out = malloc(sizeof(real) * 81);
  for (int cell = 0; cell < n_cells; ++cell) {
    const int64_t *cl = cell_ls + (int64_t)cell * N;
    real hp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (interp)
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < N; ++j) hp[i] += (real)interp[i * N + j] * (real)heaviside[cl[j]];
    for (int i = 0; i < 81; ++i) out[i] = 0;
    for (int q = 0; q < 27; ++q) {
      const int qx = q % 3, qy = (q / 3) % 3, qz = q / 9;
      real H = 0, kap = 0, gr[3] = {0, 0, 0};
      for (int k = 0; k < n1; ++k)
        for (int j = 0; j < n1; ++j)
          for (int i = 0; i < n1; ++i) {
            const int64_t d = cl[i + n1 * (j + n1 * k)];
            const real w = hv[qx][i] * hv[qy][j] * hv[qz][k];
            H += w * (real)heaviside[d];
            kap += w * (real)curvature[d];
            gr[0] += hd[qx][i] * hv[qy][j] * hv[qz][k] * (real)heaviside[d];
            gr[1] += hv[qx][i] * hd[qy][j] * hv[qz][k] * (real)heaviside[d];
            gr[2] += hv[qx][i] * hv[qy][j] * hd[qz][k] * (real)heaviside[d];
          }
      if (interp) {
        gr[0] = gr[1] = gr[2] = 0;
        for (int i = 0; i < 8; ++i) {
          const real vx = (i & 1) ? gp[qx] : 1 - gp[qx], vy = (i & 2) ? gp[qy] : 1 - gp[qy], vz = (i & 4) ? gp[qz] : 1 - gp[qz];
          gr[0] += ((i & 1) ? 1 : -1) * vy * vz * hp[i];
          gr[1] += vx * ((i & 2) ? 1 : -1) * vz * hp[i];
          gr[2] += vx * vy * ((i & 4) ? 1 : -1) * hp[i];
        }
      }
      const double *K = inv_jac + ((int64_t)cell * 27 + q) * 9;
      const real JxW = jxw[(int64_t)cell * 27 + q];
      real g[3];
      for (int e = 0; e < 3; ++e) g[e] = gr[0] * (real)K[e] + gr[1] * (real)K[3 + e] + gr[2] * (real)K[6 + e];
      real rho = rho0;
      if (variable) {
        rho = rho0 + drho * H;
        rho_q[(int64_t)cell * 27 + q] = (double)rho;
        mu_q[(int64_t)cell * 27 + q] = (double)(mu0 + dmu * H);
      }
      real f[3];
      for (int d = 0; d < 3; ++d) f[d] = sigma * kap * g[d];
      f[2] -= gravity * rho;
      for (int n = 0; n < 27; ++n) {
        const real phiu = L2[qx][n % 3] * L2[qy][(n / 3) % 3] * L2[qz][n / 9];
        for (int d = 0; d < 3; ++d) out[d * 27 + n] += phiu * f[d] * JxW;
      }
    }
    for (int n = 0; n < 27; ++n)
      for (int d = 0; d < 3; ++d) {
        const int64_t e = 3 * cell_u[(int64_t)cell * 27 + n] + d;
        if (u_constrained && u_constrained[e]) continue;
        dst_u[e] += (double)out[d * 27 + n];
      }
  }
  free(out);
  return 0;
}
