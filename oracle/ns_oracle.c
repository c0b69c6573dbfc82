/* ns_oracle.c -- CPU restatement of adaflo's matrix-free Navier-Stokes operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under adaflo_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: the reference's arithmetic for this path lives in deal.II
 * (un-vendored, unpinned: FIND_PACKAGE(deal.II 9.3), CMakeLists.txt:26) and
 * cannot be built in this image.  This file restates the
 * published algorithm (FEEvaluation read/evaluate/submit/integrate/distribute,
 * SURVEY 8(c) items 1-6) as DENSE loops -- deliberately not sum-factorised, so
 * that it shares no structure with the CUDA kernels -- and the per-quadrature
 * point physics of source/navier_stokes_matrix.cc.  It is pinned against the
 * reference's own golden log tests/beltrami_3d.output:5-7,13 (interpolation
 * error and first nonlinear residual norms) and the exact-in-space Poiseuille
 * solution of tests/poiseuille.cc:74-112 by tests/test_oracle_pins.py.
 *
 * Conventions: cell-local velocity dof i = c*N + n (component-blocked,
 * lexicographic nodes, x fastest); vector entry of (node, c) = node*3 + c.
 * Accumulation inside a cell is done in long double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long double real;

#define MAXN1D 6
#define DIM 3

typedef struct {
  int32_t n_cells;
  int32_t n_q_1d;     /* Gauss points per direction the geometry arrays are given on */
  int32_t u_degree;   /* FE_Q degree of the velocity (2) */
  int32_t p_degree;   /* FE_Q degree of the pressure (1) */
  int32_t p_dg0;      /* 1: FE_Q_DG0 (one extra constant dof per cell, last) */
  int32_t pad_;
  /* gather lists in CSR form: local value = sum_k w[k] * vec[idx[k]] */
  const int64_t *u_ptr;
  const int64_t *u_idx;
  const double *u_w;
  const int64_t *p_ptr;
  const int64_t *p_idx;
  const double *p_w;
  const uint8_t *u_constrained; /* per vector entry */
  const uint8_t *p_constrained;
  const double *inv_jac;        /* [n_cells][nq][3][3], K[f][e] = d xi_f / d x_e */
  const double *jxw;            /* [n_cells][nq] */
} oracle_mesh;

/* mirrors adaflo_b200_flow_params / adaflo_b200_time_weights of include/adaflo_b200.h */
typedef struct {
  int32_t physical_type;  /* 0 incompressible, 1 stationary, 2 stokes  (parameters.h:54-59) */
  int32_t linearization;  /* 0 newton 1 picard 2 semi-implicit 3 explicit 4 projection (parameters.h:76-83) */
  int32_t augmented_taylor_hood;
  int32_t variable_coefficients;
  double density, viscosity, damping, tau_grad_div, beta;
} oracle_flow_params;

typedef struct {
  double weight, weight_old, weight_old_old, tau1, extrapolate_old, extrapolate_old_old;
} oracle_time_weights;

enum { OP_VMULT = 0, OP_RESIDUAL = 1, OP_VMULT_VELOCITY = 2 };

/* ----------------------------------------------------------------- 1-D tools */
static void legendre(int n, real x, real *p, real *dp)
{
  real p0 = 1, p1 = x;
  if (n == 0) { *p = 1; *dp = 0; return; }
  for (int k = 2; k <= n; ++k) {
    real pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1; p1 = pk;
  }
  *p = p1;
  *dp = n * (x * p1 - p0) / (x * x - 1);
}

/* QGauss<1>(n) on [0,1] */
static void gauss01(int n, real *x, real *w)
{
  const real pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < n; ++i) {
    real z = -cosl(pi * (i + 0.75L) / (n + 0.5L));
    for (int it = 0; it < 100; ++it) {
      real p, dp;
      legendre(n, z, &p, &dp);
      real dz = p / dp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    real p, dp;
    legendre(n, z, &p, &dp);
    x[i] = 0.5L * (z + 1);
    w[i] = 1.0L / ((1 - z * z) * dp * dp);
  }
}

/* QGaussLobatto<1>(n) points on [0,1]: support points of FE_Q(n-1) */
static void lobatto01(int n, real *x)
{
  const real pi = 3.14159265358979323846264338327950288L;
  x[0] = 0; x[n - 1] = 1;
  int m = n - 1; /* interior points: roots of P'_m */
  for (int i = 1; i < n - 1; ++i) {
    real z = -cosl(pi * i / m);
    for (int it = 0; it < 100; ++it) {
      real p, dp;
      legendre(m, z, &p, &dp);
      /* P''_m from the Legendre ODE: (1-z^2) P'' = 2 z P' - m(m+1) P */
      real ddp = (2 * z * dp - m * (m + 1) * p) / (1 - z * z);
      real dz = dp / ddp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[i] = 0.5L * (z + 1);
  }
}

static void lagrange(int n, const real *nodes, real x, real *v, real *d)
{
  for (int i = 0; i < n; ++i) {
    real denom = 1, val = 1, der = 0;
    for (int j = 0; j < n; ++j)
      if (j != i) { denom *= nodes[i] - nodes[j]; val *= x - nodes[j]; }
    for (int m = 0; m < n; ++m) {
      if (m == i) continue;
      real t = 1;
      for (int j = 0; j < n; ++j)
        if (j != i && j != m) t *= x - nodes[j];
      der += t;
    }
    v[i] = val / denom;
    d[i] = der / denom;
  }
}

/* dense shape tables: val[q][n], grad[q][n][f] on the reference cell [0,1]^3 */
typedef struct {
  int n1, nq1, N, NQ;
  real *val;
  real *grad;
} shape_table;

static void build_table(shape_table *t, int degree, int nq1)
{
  real nodes[MAXN1D], qx[MAXN1D], qw[MAXN1D];
  real v1[MAXN1D][MAXN1D], d1[MAXN1D][MAXN1D];
  int n1 = degree + 1;
  lobatto01(n1, nodes);
  gauss01(nq1, qx, qw);
  for (int q = 0; q < nq1; ++q) lagrange(n1, nodes, qx[q], v1[q], d1[q]);
  t->n1 = n1; t->nq1 = nq1; t->N = n1 * n1 * n1; t->NQ = nq1 * nq1 * nq1;
  t->val = (real *)malloc(sizeof(real) * t->NQ * t->N);
  t->grad = (real *)malloc(sizeof(real) * t->NQ * t->N * 3);
  for (int qz = 0; qz < nq1; ++qz)
    for (int qy = 0; qy < nq1; ++qy)
      for (int qxx = 0; qxx < nq1; ++qxx) {
        int q = qxx + nq1 * (qy + nq1 * qz);
        for (int iz = 0; iz < n1; ++iz)
          for (int iy = 0; iy < n1; ++iy)
            for (int ix = 0; ix < n1; ++ix) {
              int n = ix + n1 * (iy + n1 * iz);
              t->val[q * t->N + n] = v1[qxx][ix] * v1[qy][iy] * v1[qz][iz];
              t->grad[(q * t->N + n) * 3 + 0] = d1[qxx][ix] * v1[qy][iy] * v1[qz][iz];
              t->grad[(q * t->N + n) * 3 + 1] = v1[qxx][ix] * d1[qy][iy] * v1[qz][iz];
              t->grad[(q * t->N + n) * 3 + 2] = v1[qxx][ix] * v1[qy][iy] * d1[qz][iz];
            }
      }
}

static void free_table(shape_table *t) { free(t->val); free(t->grad); }

/* exported for tests: 1-D quadrature and shape data in double */
void oracle_gauss01(int n, double *x, double *w)
{
  real xx[MAXN1D], ww[MAXN1D];
  gauss01(n, xx, ww); /* ww are already the [0,1] weights (half of the [-1,1] ones) */
  for (int i = 0; i < n; ++i) { x[i] = (double)xx[i]; w[i] = (double)ww[i]; }
}

void oracle_shape_1d(int degree, int nq, double *val, double *der)
{
  real nodes[MAXN1D], qx[MAXN1D], qw[MAXN1D], v[MAXN1D], d[MAXN1D];
  lobatto01(degree + 1, nodes);
  gauss01(nq, qx, qw);
  for (int q = 0; q < nq; ++q) {
    lagrange(degree + 1, nodes, qx[q], v, d);
    for (int i = 0; i <= degree; ++i) { val[q * (degree + 1) + i] = (double)v[i]; der[q * (degree + 1) + i] = (double)d[i]; }
  }
}

/* ------------------------------------------------------------- gather/scatter
 * read_dof_values (resolve=1): constrained entries read as zero, hanging-node
 * rows expand through their weights; read_dof_values_plain (resolve=0): raw
 * entries. */
static void gather(const int64_t *ptr, const int64_t *idx, const double *w, const uint8_t *constrained,
                   const double *vec, int64_t row0, int n, int resolve, real *out)
{
  for (int i = 0; i < n; ++i) {
    real s = 0;
    for (int64_t k = ptr[row0 + i]; k < ptr[row0 + i + 1]; ++k) {
      if (resolve && constrained && constrained[idx[k]]) continue;
      s += (real)w[k] * (real)vec[idx[k]];
    }
    out[i] = s;
  }
}

/* distribute_local_to_global: transposed weights, constrained rows skipped */
static void scatter_add(const int64_t *ptr, const int64_t *idx, const double *w, const uint8_t *constrained,
                        double *vec, int64_t row0, int n, const real *in)
{
  for (int i = 0; i < n; ++i)
    for (int64_t k = ptr[row0 + i]; k < ptr[row0 + i + 1]; ++k) {
      if (constrained && constrained[idx[k]]) continue;
      vec[idx[k]] += (double)((real)w[k] * in[i]);
    }
}

/* --------------------------------------------------------------- cell loop
 * local_operation<degree_p, VectorType, LocalOps>, navier_stokes_matrix.cc:601-916.
 * Adds into dst (cell_loop does not zero dst).  lin: [n_cells][12][NQ]
 * (rows 0..2 u, 3+3d+e grad), read by vmult, written by residual.
 * rho, mu, damp: [n_cells][NQ] or NULL. */
int oracle_ns_cell_loop(const oracle_mesh *m, const oracle_flow_params *prm, const oracle_time_weights *tw, int op,
                        const double *src_u, const double *src_p, const double *old_u, const double *old_old_u,
                        double *dst_u, double *dst_p, double *lin, const double *rho_q, const double *mu_q,
                        const double *damp_q)
{
  shape_table tu, tp;
  build_table(&tu, m->u_degree, m->n_q_1d);
  build_table(&tp, m->p_degree, m->n_q_1d);
  const int N = tu.N, NP = tp.N + (m->p_dg0 ? 1 : 0), NQ = tu.NQ;
  const int stokes = prm->physical_type == 2;
  const int stationary = prm->physical_type == 1;
  const real gamma = prm->physical_type == 0 ? tw->weight : 0.0; /* :621-623 */
  const real w_old = tw->weight_old, w_oo = tw->weight_old_old, tau1 = tw->tau1;
  const real beta = prm->beta;
  const int lz = prm->linearization;
  const int need_extrap = (lz == 4 || lz == 2 || lz == 3);       /* :644-647 */
  const int var = prm->variable_coefficients;

  real *ul = malloc(sizeof(real) * 3 * N), *uo = malloc(sizeof(real) * 3 * N), *uoo = malloc(sizeof(real) * 3 * N);
  real *pl = malloc(sizeof(real) * NP), *out_u = malloc(sizeof(real) * 3 * N), *out_p = malloc(sizeof(real) * NP);

  for (int cell = 0; cell < m->n_cells; ++cell) {
    const int resolve = (op != OP_RESIDUAL); /* :663-666 */
    gather(m->u_ptr, m->u_idx, m->u_w, m->u_constrained, src_u, (int64_t)cell * 3 * N, 3 * N, resolve, ul);
    if (op != OP_VMULT_VELOCITY)
      gather(m->p_ptr, m->p_idx, m->p_w, m->p_constrained, src_p, (int64_t)cell * NP, NP, resolve, pl);
    const int use_old = (op == OP_RESIDUAL && prm->physical_type == 0); /* :673-674 */
    if (use_old) {
      gather(m->u_ptr, m->u_idx, m->u_w, 0, old_u, (int64_t)cell * 3 * N, 3 * N, 0, uo);
      gather(m->u_ptr, m->u_idx, m->u_w, 0, old_old_u, (int64_t)cell * 3 * N, 3 * N, 0, uoo);
    }
    for (int i = 0; i < 3 * N; ++i) out_u[i] = 0;
    for (int i = 0; i < NP; ++i) out_p[i] = 0;

    for (int q = 0; q < NQ; ++q) {
      const double *K = m->inv_jac + ((int64_t)cell * NQ + q) * 9;
      const real JxW = m->jxw[(int64_t)cell * NQ + q];
      real val_u[3] = {0, 0, 0}, gref[3][3] = {{0}}, grad[3][3];
      real val_o[3] = {0, 0, 0}, val_oo[3] = {0, 0, 0}, gro[3][3] = {{0}}, groo[3][3] = {{0}};
      for (int d = 0; d < 3; ++d)
        for (int n = 0; n < N; ++n) {
          real phi = tu.val[q * N + n];
          val_u[d] += phi * ul[d * N + n];
          for (int f = 0; f < 3; ++f) gref[d][f] += tu.grad[(q * N + n) * 3 + f] * ul[d * N + n];
          if (use_old) {
            val_o[d] += phi * uo[d * N + n];
            val_oo[d] += phi * uoo[d * N + n];
            for (int f = 0; f < 3; ++f) {
              gro[d][f] += tu.grad[(q * N + n) * 3 + f] * uo[d * N + n];
              groo[d][f] += tu.grad[(q * N + n) * 3 + f] * uoo[d * N + n];
            }
          }
        }
      /* get_gradient: real-space gradient */
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          real s = 0;
          for (int f = 0; f < 3; ++f) s += gref[d][f] * (real)K[f * 3 + e];
          grad[d][e] = s;
        }
      real pres = 0;
      if (op != OP_VMULT_VELOCITY) {
        for (int n = 0; n < tp.N; ++n) pres += tp.val[q * tp.N + n] * pl[n];
        if (m->p_dg0) pres += pl[tp.N]; /* FE_Q_DG0: last shape function == 1 */
      }
      const real div = grad[0][0] + grad[1][1] + grad[2][2];  /* :706 */
      double *L = lin ? lin + (int64_t)cell * 12 * NQ : 0;
#define LIN(k) L[(k) * NQ + q]
      real conv[3] = {0, 0, 0};
      int have_conv = 0;
      if (!stokes) { /* :708 */
        have_conv = 1;
        const real rho = var ? (real)rho_q[(int64_t)cell * NQ + q] : (real)prm->density;
        for (int d = 0; d < 3; ++d) conv[d] = val_u[d] * gamma; /* :717 */
        if (op == OP_RESIDUAL) {
          if (!stationary) /* :727-732 */
            for (int d = 0; d < 3; ++d) conv[d] += val_o[d] * w_old + val_oo[d] * w_oo;
          if (need_extrap) { /* :740-782 */
            real og[3][3], ov[3];
            for (int d = 0; d < 3; ++d) {
              real a[3], b[3];
              for (int f = 0; f < 3; ++f) { a[f] = gro[d][f]; b[f] = groo[d][f]; }
              for (int e = 0; e < 3; ++e) {
                real s1 = 0, s2 = 0;
                for (int f = 0; f < 3; ++f) { s1 += a[f] * (real)K[f * 3 + e]; s2 += b[f] * (real)K[f * 3 + e]; }
                og[d][e] = s1 * tw->extrapolate_old + s2 * tw->extrapolate_old_old;
              }
              ov[d] = val_o[d] * tw->extrapolate_old + val_oo[d] * tw->extrapolate_old_old;
            }
            const real ediv = og[0][0] + og[1][1] + og[2][2];
            if (lz == 3) { /* explicit, :758-767 */
              for (int d = 0; d < 3; ++d) {
                real res = beta * ediv * ov[d];
                for (int e = 0; e < 3; ++e) res += ov[e] * og[d][e];
                conv[d] += tau1 * res;
              }
            } else { /* :768-781 */
              for (int d = 0; d < 3; ++d) {
                real res = beta * ediv * val_u[d];
                for (int e = 0; e < 3; ++e) res += ov[e] * grad[d][e];
                conv[d] += tau1 * res;
                LIN(d) = (double)ov[d];
              }
              LIN(3) = (double)ediv;
            }
          } else { /* Newton / Picard residual, :783-799 */
            for (int d = 0; d < 3; ++d) {
              real res = beta * div * val_u[d];
              for (int e = 0; e < 3; ++e) res += val_u[e] * grad[d][e];
              conv[d] += tau1 * res;
              LIN(d) = (double)val_u[d];
            }
            if (lz == 0) {
              for (int d = 0; d < 3; ++d)
                for (int e = 0; e < 3; ++e) LIN(3 + 3 * d + e) = (double)grad[d][e];
            } else
              LIN(3) = (double)div;
          }
        } else if (lz == 0) { /* Newton mat-vec, :802-816 */
          real lu[3], lg[3][3];
          for (int d = 0; d < 3; ++d) {
            lu[d] = LIN(d);
            for (int e = 0; e < 3; ++e) lg[d][e] = LIN(3 + 3 * d + e);
          }
          const real f1 = beta * div, f2 = beta * (lg[0][0] + lg[1][1] + lg[2][2]);
          for (int d = 0; d < 3; ++d) {
            real res = f1 * lu[d] + f2 * val_u[d];
            for (int e = 0; e < 3; ++e) res += lu[e] * grad[d][e] + val_u[e] * lg[d][e];
            conv[d] += tau1 * res;
          }
        } else if (lz != 3) { /* Picard / semi-implicit / projection mat-vec, :817-826 */
          for (int d = 0; d < 3; ++d) {
            real res = beta * (real)LIN(3) * val_u[d];
            for (int e = 0; e < 3; ++e) res += (real)LIN(e) * grad[d][e];
            conv[d] += tau1 * res;
          }
        }
        for (int d = 0; d < 3; ++d) conv[d] *= rho; /* :827 */
        const real damping = var ? (real)damp_q[(int64_t)cell * NQ + q] : (real)prm->damping;
        for (int d = 0; d < 3; ++d) conv[d] -= damping * val_u[d]; /* :831-835 */
      }
#undef LIN
      /* viscous stress, :841-892 */
      const real tmu = (var ? (real)mu_q[(int64_t)cell * NQ + q] : (real)prm->viscosity) * tau1;
      real S[3][3];
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e)
          if (d != e) S[d][e] = tmu * (grad[d][e] + grad[e][d]);
      for (int d = 0; d < 3; ++d) {
        S[d][d] = 2 * tmu * grad[d][d] + (real)prm->tau_grad_div * div;
        if (op != OP_VMULT_VELOCITY) S[d][d] -= pres;
      }
      /* submit_value / submit_gradient / integrate */
      for (int d = 0; d < 3; ++d) {
        real Sref[3];
        for (int f = 0; f < 3; ++f) {
          real s = 0;
          for (int e = 0; e < 3; ++e) s += (real)K[f * 3 + e] * S[d][e];
          Sref[f] = s * JxW;
        }
        for (int n = 0; n < N; ++n) {
          real acc = 0;
          if (have_conv) acc += tu.val[q * N + n] * conv[d] * JxW;
          for (int f = 0; f < 3; ++f) acc += tu.grad[(q * N + n) * 3 + f] * Sref[f];
          out_u[d * N + n] += acc;
        }
      }
      if (op != OP_VMULT_VELOCITY && lz != 4) { /* :856, :902-907 */
        for (int n = 0; n < tp.N; ++n) out_p[n] += tp.val[q * tp.N + n] * (-div) * JxW;
        if (m->p_dg0) out_p[tp.N] += (-div) * JxW;
      }
    }
    scatter_add(m->u_ptr, m->u_idx, m->u_w, m->u_constrained, dst_u, (int64_t)cell * 3 * N, 3 * N, out_u);
    if (op != OP_VMULT_VELOCITY && lz != 4)
      scatter_add(m->p_ptr, m->p_idx, m->p_w, m->p_constrained, dst_p, (int64_t)cell * NP, NP, out_p);
  }
  free(ul); free(uo); free(uoo); free(pl); free(out_u); free(out_p);
  free_table(&tu); free_table(&tp);
  return 0;
}

/* local_divergence<degree_p, weight_by_viscosity>, navier_stokes_matrix.cc:920-961.
 * dst_p += -(q, [mu] div u).  Geometry arrays on the velocity quadrature. */
int oracle_ns_divergence(const oracle_mesh *m, const oracle_flow_params *prm, int weight_by_viscosity,
                         const double *src_u, double *dst_p, const double *mu_q)
{
  shape_table tu, tp;
  build_table(&tu, m->u_degree, m->n_q_1d);
  build_table(&tp, m->p_degree, m->n_q_1d);
  const int N = tu.N, NP = tp.N + (m->p_dg0 ? 1 : 0), NQ = tu.NQ;
  real *ul = malloc(sizeof(real) * 3 * N), *out_p = malloc(sizeof(real) * NP);
  for (int cell = 0; cell < m->n_cells; ++cell) {
    gather(m->u_ptr, m->u_idx, m->u_w, m->u_constrained, src_u, (int64_t)cell * 3 * N, 3 * N, prm->linearization != 4, ul);
    for (int i = 0; i < NP; ++i) out_p[i] = 0;
    for (int q = 0; q < NQ; ++q) {
      const double *K = m->inv_jac + ((int64_t)cell * NQ + q) * 9;
      const real JxW = m->jxw[(int64_t)cell * NQ + q];
      real div = 0;
      for (int d = 0; d < 3; ++d)
        for (int n = 0; n < N; ++n)
          for (int f = 0; f < 3; ++f) div += tu.grad[(q * N + n) * 3 + f] * ul[d * N + n] * (real)K[f * 3 + d];
      real weight = -1;
      if (weight_by_viscosity) weight = mu_q ? -(real)mu_q[(int64_t)cell * NQ + q] : -(real)prm->viscosity; /* :948-953 */
      for (int n = 0; n < tp.N; ++n) out_p[n] += tp.val[q * tp.N + n] * weight * div * JxW;
      if (m->p_dg0) out_p[tp.N] += weight * div * JxW;
    }
    scatter_add(m->p_ptr, m->p_idx, m->p_w, m->p_constrained, dst_p, (int64_t)cell * NP, NP, out_p);
  }
  free(ul); free(out_p); free_table(&tu); free_table(&tp);
  return 0;
}

/* Scalar pressure-space operators, navier_stokes_matrix.cc:965-1140.
 * kind 0: mass       q <- p * coeff_cell            (local_pressure_mass, :1036-1071)
 * kind 1: laplace    grad q <- grad p * coeff       (local_pressure_poisson :965-1032, local_pressure_convdiff :1099-1140)
 * kind 2: mass weight q <- 1                        (local_pressure_mass_weight :1075-1095)
 * coeff: per cell [n_cells] (cell-constant) or per q [n_cells][NQ] if coeff_per_q.
 * The caller evaluates the coefficient rule (that is host logic of the operator class). */
int oracle_pressure_op(const oracle_mesh *m, int kind, const double *coeff, int coeff_per_q, const double *src_p, double *dst_p)
{
  shape_table tp;
  build_table(&tp, m->p_degree, m->n_q_1d);
  const int NP = tp.N + (m->p_dg0 ? 1 : 0), NQ = tp.NQ;
  real *pl = malloc(sizeof(real) * NP), *out_p = malloc(sizeof(real) * NP);
  for (int cell = 0; cell < m->n_cells; ++cell) {
    if (kind != 2) gather(m->p_ptr, m->p_idx, m->p_w, m->p_constrained, src_p, (int64_t)cell * NP, NP, 1, pl);
    for (int i = 0; i < NP; ++i) out_p[i] = 0;
    for (int q = 0; q < NQ; ++q) {
      const double *K = m->inv_jac + ((int64_t)cell * NQ + q) * 9;
      const real JxW = m->jxw[(int64_t)cell * NQ + q];
      const real c = coeff ? (real)(coeff_per_q ? coeff[(int64_t)cell * NQ + q] : coeff[cell]) : 1;
      if (kind == 0 || kind == 2) {
        real v = 1;
        if (kind == 0) {
          v = 0;
          for (int n = 0; n < tp.N; ++n) v += tp.val[q * tp.N + n] * pl[n];
          if (m->p_dg0) v += pl[tp.N];
          v *= c;
        }
        for (int n = 0; n < tp.N; ++n) out_p[n] += tp.val[q * tp.N + n] * v * JxW;
        if (m->p_dg0) out_p[tp.N] += v * JxW;
      } else {
        real gref[3] = {0, 0, 0}, g[3], gr[3];
        for (int n = 0; n < tp.N; ++n)
          for (int f = 0; f < 3; ++f) gref[f] += tp.grad[(q * tp.N + n) * 3 + f] * pl[n];
        for (int e = 0; e < 3; ++e) {
          g[e] = 0;
          for (int f = 0; f < 3; ++f) g[e] += gref[f] * (real)K[f * 3 + e];
          g[e] *= c;
        }
        for (int f = 0; f < 3; ++f) {
          gr[f] = 0;
          for (int e = 0; e < 3; ++e) gr[f] += (real)K[f * 3 + e] * g[e];
          gr[f] *= JxW;
        }
        for (int n = 0; n < tp.N; ++n)
          for (int f = 0; f < 3; ++f) out_p[n] += tp.grad[(q * tp.N + n) * 3 + f] * gr[f];
        /* DG0 shape function has zero gradient */
      }
    }
    scatter_add(m->p_ptr, m->p_idx, m->p_w, m->p_constrained, dst_p, (int64_t)cell * NP, NP, out_p);
  }
  free(pl); free(out_p); free_table(&tp);
  return 0;
}

/* L2 norm of (u_h - u_exact) with u_exact given at the quadrature points:
 * VectorTools::integrate_difference(.., L2_norm) as used by tests/beltrami.cc:262-282 */
double oracle_l2_error(const oracle_mesh *m, int n_comp, const double *vec, const double *exact_q /* [n_cells][NQ][n_comp] */)
{
  shape_table t;
  build_table(&t, n_comp == 3 ? m->u_degree : m->p_degree, m->n_q_1d);
  const int N = t.N, NQ = t.NQ;
  const int64_t *ptr = n_comp == 3 ? m->u_ptr : m->p_ptr, *idx = n_comp == 3 ? m->u_idx : m->p_idx;
  const double *w = n_comp == 3 ? m->u_w : m->p_w;
  const int nloc = n_comp * N + ((n_comp == 1 && m->p_dg0) ? 1 : 0);
  real *loc = malloc(sizeof(real) * nloc);
  real sum = 0;
  for (int cell = 0; cell < m->n_cells; ++cell) {
    gather(ptr, idx, w, 0, vec, (int64_t)cell * nloc, nloc, 0, loc);
    for (int q = 0; q < NQ; ++q) {
      const real JxW = m->jxw[(int64_t)cell * NQ + q];
      for (int d = 0; d < n_comp; ++d) {
        real v = 0;
        for (int n = 0; n < N; ++n) v += t.val[q * N + n] * loc[d * N + n];
        if (n_comp == 1 && m->p_dg0) v += loc[N]; /* FE_Q_DG0: the element-constant part */
        real e = v - (real)exact_q[((int64_t)cell * NQ + q) * n_comp + d];
        sum += e * e * JxW;
      }
    }
  }
  free(loc); free_table(&t);
  return (double)sqrtl(sum);
}
