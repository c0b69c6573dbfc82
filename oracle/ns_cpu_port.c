/* ns_cpu_port.c -- optimised CPU port of the Navier-Stokes vmult cell loop, used ONLY as
 * the timed CPU baseline of bench.py (cpu_baseline.kind = "port", --impl reference).
 *
 * TEST/BENCH INFRASTRUCTURE: a restatement, NOT adaflo+deal.II (which cannot be built in
 * this image).  It follows the structure of the reference's CPU path
 * (source/navier_stokes_matrix.cc:601-916 on top of deal.II FEEvaluation): sum-factorised
 * evaluation/integration per cell, threads over contiguous cell ranges (standing in for
 * the MPI ranks of the p4est partition).  Restricted to the benchmark variants: Cartesian
 * cells, constant coefficients, vmult with Stokes / Newton / Picard convection.
 * Validated against the dense oracle (ns_oracle.c) by tests/test_cpu_port.py.
 *
 * SIMD: like deal.II's VectorizedArray, the arithmetic is vectorised ACROSS CELLS: a batch of 4 x-adjacent cells (AVX2, 4 doubles)
 * is gathered lane by lane, evaluated / integrated with 4-wide operations (GCC vector extensions) and scattered lane by lane; the
 * linearisation point of a batch is stored [row][27 points][4 lanes] as MatrixFree stores its per-point data.  Cells of a row that do
 * not fill a batch take the scalar routine.
 *
 * Threading: cells are lexicographic (rows of nx cells along x, ny rows per z-layer); a row
 * (iy, iz) shares dofs only with the rows (iy+-1, iz+-1), so the four colours (iy%2, iz%2) are
 * processed in four conflict-free parallel sweeps and no atomics are needed.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NODE_MASK 0x1FFFFFFFu
#define FLAG_SHIFT 29

typedef struct {
  int32_t n_cells, nx, ny; /* cells are lexicographic: rows of nx cells, ny rows per z-layer */
  const uint32_t *cell_u; /* [n_cells][27] with Dirichlet flag bits (same encoding as the C ABI) */
  const uint32_t *cell_p; /* [n_cells][8] */
  double inv_h[3], det;
  int32_t mode;           /* 0 stokes, 1 newton, 2 picard, 3 no convection */
  double gamma, tau1, beta, rho, mu, damping, tau_gd;
  const double *lin;      /* [n_cells][12][27] */
  const double *lin_b;    /* batches of 4 x-adjacent cells: [row][nx / 4][12][27][4], or null (scalar path only) */
} port_args;

static const double N[3][3] = {{0.6872983346207416, 0.4, -0.08729833462074169}, {0.0, 1.0, 0.0}, {-0.08729833462074169, 0.4, 0.6872983346207416}};
static const double D[3][3] = {{-2.549193338482967, 3.0983866769659336, -0.5491933384829668}, {-1.0, 0.0, 1.0}, {0.5491933384829668, -3.0983866769659336, 2.549193338482967}};
static const double NP[3][2] = {{0.8872983346207417, 0.11270166537925831}, {0.5, 0.5}, {0.11270166537925831, 0.8872983346207417}};
static const double W[3] = {0.2777777777777778, 0.4444444444444444, 0.2777777777777778};

/* apply a 3x3 1-D matrix along direction dir of a 3x3x3 array (transpose: M^T) */
static inline __attribute__((always_inline)) void apply3(const double M[3][3], int transpose, int dir, const double *in, double *out, int add)
{
  const int stride = dir == 0 ? 1 : (dir == 1 ? 3 : 9);
  for (int o = 0; o < 9; ++o) {
    const int base = dir == 0 ? 3 * o : (dir == 1 ? (o % 3) + 9 * (o / 3) : o);
    const double a0 = in[base], a1 = in[base + stride], a2 = in[base + 2 * stride];
    for (int q = 0; q < 3; ++q) {
      const double v = transpose ? M[0][q] * a0 + M[1][q] * a1 + M[2][q] * a2 : M[q][0] * a0 + M[q][1] * a1 + M[q][2] * a2;
      if (add) out[base + q * stride] += v; else out[base + q * stride] = v;
    }
  }
}

static void cell_vmult(const port_args *A, int cell, const double *src_u, const double *src_p, double *dst_u, double *dst_p)
{
  double u[3][27], val[3][27], g[3][3][27], t0[27], t1[27], pl[8], pv[27];
  const uint32_t *cu = A->cell_u + (size_t)cell * 27, *cp = A->cell_p + (size_t)cell * 8;
  for (int n = 0; n < 27; ++n) {
    const uint32_t idx = cu[n] & NODE_MASK, fl = cu[n] >> FLAG_SHIFT;
    for (int c = 0; c < 3; ++c) u[c][n] = ((fl >> c) & 1u) ? 0.0 : src_u[3 * (size_t)idx + c];
  }
  for (int n = 0; n < 8; ++n) pl[n] = (cp[n] >> FLAG_SHIFT) ? 0.0 : src_p[cp[n] & NODE_MASK];
  /* evaluate: interpolate to the Gauss points, then differentiate per direction */
  for (int c = 0; c < 3; ++c) {
    apply3(N, 0, 0, u[c], t0, 0);       /* x */
    apply3(N, 0, 1, t0, t1, 0);         /* xy */
    apply3(N, 0, 2, t1, val[c], 0);     /* value */
    apply3(D, 0, 2, t1, g[c][2], 0);    /* d/dz */
    apply3(D, 0, 1, t0, t1, 0);
    apply3(N, 0, 2, t1, g[c][1], 0);    /* d/dy */
    apply3(D, 0, 0, u[c], t0, 0);
    apply3(N, 0, 1, t0, t1, 0);
    apply3(N, 0, 2, t1, g[c][0], 0);    /* d/dx */
  }
  /* pressure values */
  for (int qz = 0; qz < 3; ++qz)
    for (int qy = 0; qy < 3; ++qy)
      for (int qx = 0; qx < 3; ++qx) {
        double v = 0;
        for (int i = 0; i < 8; ++i) v += NP[qx][i & 1] * NP[qy][(i >> 1) & 1] * NP[qz][i >> 2] * pl[i];
        pv[qx + 3 * qy + 9 * qz] = v;
      }
  double tp[27];
  const double *lin = A->lin ? A->lin + (size_t)cell * 12 * 27 : 0;
  for (int q = 0; q < 27; ++q) {
    const double jxw = A->det * W[q % 3] * W[(q / 3) % 3] * W[q / 9];
    double G[3][3];
    for (int d = 0; d < 3; ++d)
      for (int e = 0; e < 3; ++e) G[d][e] = g[d][e][q] * A->inv_h[e];
    const double div = G[0][0] + G[1][1] + G[2][2];
    double conv[3] = {0, 0, 0};
    if (A->mode != 0) {
      const double uu[3] = {val[0][q], val[1][q], val[2][q]};
      for (int d = 0; d < 3; ++d) conv[d] = A->gamma * uu[d];
      if (A->mode == 1) {
        double lu[3], lg[3][3];
        for (int d = 0; d < 3; ++d) lu[d] = lin[d * 27 + q];
        for (int d = 0; d < 3; ++d)
          for (int e = 0; e < 3; ++e) lg[d][e] = lin[(3 + 3 * d + e) * 27 + q];
        const double f1 = A->beta * div, f2 = A->beta * (lg[0][0] + lg[1][1] + lg[2][2]);
        for (int d = 0; d < 3; ++d) {
          double res = f1 * lu[d] + f2 * uu[d];
          for (int e = 0; e < 3; ++e) res += lu[e] * G[d][e] + uu[e] * lg[d][e];
          conv[d] += A->tau1 * res;
        }
      } else if (A->mode == 2) {
        for (int d = 0; d < 3; ++d) {
          double res = A->beta * lin[3 * 27 + q] * uu[d];
          for (int e = 0; e < 3; ++e) res += lin[e * 27 + q] * G[d][e];
          conv[d] += A->tau1 * res;
        }
      }
      for (int d = 0; d < 3; ++d) conv[d] = (conv[d] * A->rho - A->damping * uu[d]) * jxw;
    }
    const double tmu = A->mu * A->tau1;
    double S[3][3];
    S[0][1] = S[1][0] = tmu * (G[0][1] + G[1][0]);
    S[0][2] = S[2][0] = tmu * (G[0][2] + G[2][0]);
    S[1][2] = S[2][1] = tmu * (G[1][2] + G[2][1]);
    for (int d = 0; d < 3; ++d) S[d][d] = 2 * tmu * G[d][d] + A->tau_gd * div - pv[q];
    tp[q] = -div * jxw;
    for (int d = 0; d < 3; ++d) {
      val[d][q] = conv[d];
      for (int e = 0; e < 3; ++e) g[d][e][q] = A->inv_h[e] * S[d][e] * jxw;
    }
  }
  /* integrate */
  for (int c = 0; c < 3; ++c) {
    double r[27];
    apply3(N, 1, 2, val[c], t0, 0);
    apply3(D, 1, 2, g[c][2], t0, 1);    /* z pass of value + d/dz */
    apply3(N, 1, 1, t0, t1, 0);
    apply3(N, 1, 2, g[c][1], t0, 0);
    apply3(D, 1, 1, t0, t1, 1);         /* + d/dy */
    apply3(N, 1, 0, t1, r, 0);
    apply3(N, 1, 2, g[c][0], t0, 0);
    apply3(N, 1, 1, t0, t1, 0);
    apply3(D, 1, 0, t1, r, 1);          /* + d/dx */
    for (int n = 0; n < 27; ++n) {
      const uint32_t idx = cu[n] & NODE_MASK, fl = cu[n] >> FLAG_SHIFT;
      if (!((fl >> c) & 1u)) dst_u[3 * (size_t)idx + c] += r[n];
    }
  }
  for (int i = 0; i < 8; ++i) {
    if (cp[i] >> FLAG_SHIFT) continue;
    double v = 0;
    for (int qz = 0; qz < 3; ++qz)
      for (int qy = 0; qy < 3; ++qy)
        for (int qx = 0; qx < 3; ++qx) v += NP[qx][i & 1] * NP[qy][(i >> 1) & 1] * NP[qz][i >> 2] * tp[qx + 3 * qy + 9 * qz];
    dst_p[cp[i] & NODE_MASK] += v;
  }
}

/* ------------------------------------------------------------------ 4 cells at a time (SIMD over cells) */
typedef double v4 __attribute__((vector_size(32)));
typedef double v4u __attribute__((vector_size(32), aligned(8)));
#define LANE(v, l) (((double *)&(v))[l])

static inline __attribute__((always_inline)) void apply3v(const double M[3][3], int transpose, int dir, const v4 *in, v4 *out, int add)
{
  const int stride = dir == 0 ? 1 : (dir == 1 ? 3 : 9);
  for (int o = 0; o < 9; ++o) {
    const int base = dir == 0 ? 3 * o : (dir == 1 ? (o % 3) + 9 * (o / 3) : o);
    const v4 a0 = in[base], a1 = in[base + stride], a2 = in[base + 2 * stride];
    for (int q = 0; q < 3; ++q) {
      const v4 v = transpose ? M[0][q] * a0 + M[1][q] * a1 + M[2][q] * a2 : M[q][0] * a0 + M[q][1] * a1 + M[q][2] * a2;
      if (add) out[base + q * stride] += v; else out[base + q * stride] = v;
    }
  }
}

static void batch_vmult(const port_args *A, int cell0, const double *lin_batch, const double *src_u, const double *src_p, double *dst_u,
                        double *dst_p)
{
  v4 u[3][27], val[3][27], g[3][3][27], t0[27], t1[27], pl[8], pv[27], tp[27];
  const uint32_t *cu = A->cell_u + (size_t)cell0 * 27, *cp = A->cell_p + (size_t)cell0 * 8;
  for (int l = 0; l < 4; ++l) {  /* read_dof_values: lane by lane */
    for (int n = 0; n < 27; ++n) {
      const uint32_t e = cu[27 * l + n], idx = e & NODE_MASK, fl = e >> FLAG_SHIFT;
      for (int c = 0; c < 3; ++c) LANE(u[c][n], l) = ((fl >> c) & 1u) ? 0.0 : src_u[3 * (size_t)idx + c];
    }
    for (int n = 0; n < 8; ++n) LANE(pl[n], l) = (cp[8 * l + n] >> FLAG_SHIFT) ? 0.0 : src_p[cp[8 * l + n] & NODE_MASK];
  }
  for (int c = 0; c < 3; ++c) {
    apply3v(N, 0, 0, u[c], t0, 0);
    apply3v(N, 0, 1, t0, t1, 0);
    apply3v(N, 0, 2, t1, val[c], 0);
    apply3v(D, 0, 2, t1, g[c][2], 0);
    apply3v(D, 0, 1, t0, t1, 0);
    apply3v(N, 0, 2, t1, g[c][1], 0);
    apply3v(D, 0, 0, u[c], t0, 0);
    apply3v(N, 0, 1, t0, t1, 0);
    apply3v(N, 0, 2, t1, g[c][0], 0);
  }
  for (int qz = 0; qz < 3; ++qz)
    for (int qy = 0; qy < 3; ++qy)
      for (int qx = 0; qx < 3; ++qx) {
        v4 v = {0, 0, 0, 0};
        for (int i = 0; i < 8; ++i) v += (NP[qx][i & 1] * NP[qy][(i >> 1) & 1] * NP[qz][i >> 2]) * pl[i];
        pv[qx + 3 * qy + 9 * qz] = v;
      }
  const v4u *lin = (const v4u *)lin_batch;  /* row r of point q at lin[r * 27 + q] */
  for (int q = 0; q < 27; ++q) {
    const double jxw = A->det * W[q % 3] * W[(q / 3) % 3] * W[q / 9];
    v4 G[3][3];
    for (int d = 0; d < 3; ++d)
      for (int e = 0; e < 3; ++e) G[d][e] = g[d][e][q] * A->inv_h[e];
    const v4 div = G[0][0] + G[1][1] + G[2][2];
    v4 conv[3] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    if (A->mode != 0) {
      const v4 uu[3] = {val[0][q], val[1][q], val[2][q]};
      for (int d = 0; d < 3; ++d) conv[d] = A->gamma * uu[d];
      if (A->mode == 1) {
        v4 lu[3], lg[3][3];
        for (int d = 0; d < 3; ++d) lu[d] = lin[d * 27 + q];
        for (int d = 0; d < 3; ++d)
          for (int e = 0; e < 3; ++e) lg[d][e] = lin[(3 + 3 * d + e) * 27 + q];
        const v4 f1 = A->beta * div, f2 = A->beta * (lg[0][0] + lg[1][1] + lg[2][2]);
        for (int d = 0; d < 3; ++d) {
          v4 res = f1 * lu[d] + f2 * uu[d];
          for (int e = 0; e < 3; ++e) res += lu[e] * G[d][e] + uu[e] * lg[d][e];
          conv[d] += A->tau1 * res;
        }
      } else if (A->mode == 2) {
        const v4 ldiv = lin[3 * 27 + q];
        for (int d = 0; d < 3; ++d) {
          v4 res = A->beta * ldiv * uu[d];
          for (int e = 0; e < 3; ++e) res += lin[e * 27 + q] * G[d][e];
          conv[d] += A->tau1 * res;
        }
      }
      for (int d = 0; d < 3; ++d) conv[d] = (conv[d] * A->rho - A->damping * uu[d]) * jxw;
    }
    const double tmu = A->mu * A->tau1;
    v4 S[3][3];
    S[0][1] = S[1][0] = tmu * (G[0][1] + G[1][0]);
    S[0][2] = S[2][0] = tmu * (G[0][2] + G[2][0]);
    S[1][2] = S[2][1] = tmu * (G[1][2] + G[2][1]);
    for (int d = 0; d < 3; ++d) S[d][d] = 2 * tmu * G[d][d] + A->tau_gd * div - pv[q];
    tp[q] = -div * jxw;
    for (int d = 0; d < 3; ++d) {
      val[d][q] = conv[d];
      for (int e = 0; e < 3; ++e) g[d][e][q] = A->inv_h[e] * S[d][e] * jxw;
    }
  }
  for (int c = 0; c < 3; ++c) {
    v4 r[27];
    apply3v(N, 1, 2, val[c], t0, 0);
    apply3v(D, 1, 2, g[c][2], t0, 1);
    apply3v(N, 1, 1, t0, t1, 0);
    apply3v(N, 1, 2, g[c][1], t0, 0);
    apply3v(D, 1, 1, t0, t1, 1);
    apply3v(N, 1, 0, t1, r, 0);
    apply3v(N, 1, 2, g[c][0], t0, 0);
    apply3v(N, 1, 1, t0, t1, 0);
    apply3v(D, 1, 0, t1, r, 1);
    for (int l = 0; l < 4; ++l)  /* distribute_local_to_global: lane by lane (x-neighbours of a batch share nodes) */
      for (int n = 0; n < 27; ++n) {
        const uint32_t e = cu[27 * l + n], idx = e & NODE_MASK, fl = e >> FLAG_SHIFT;
        if (!((fl >> c) & 1u)) dst_u[3 * (size_t)idx + c] += LANE(r[n], l);
      }
  }
  for (int i = 0; i < 8; ++i) {
    v4 v = {0, 0, 0, 0};
    for (int qz = 0; qz < 3; ++qz)
      for (int qy = 0; qy < 3; ++qy)
        for (int qx = 0; qx < 3; ++qx) v += (NP[qx][i & 1] * NP[qy][(i >> 1) & 1] * NP[qz][i >> 2]) * tp[qx + 3 * qy + 9 * qz];
    for (int l = 0; l < 4; ++l)
      if (!(cp[8 * l + i] >> FLAG_SHIFT)) dst_p[cp[8 * l + i] & NODE_MASK] += LANE(v, l);
  }
}

/* dst += cell contributions (dst must be zeroed by the caller); returns the thread count used */
int port_ns_vmult_cells(const port_args *A, const double *src_u, const double *src_p, double *dst_u, double *dst_p)
{
  const int rows = A->n_cells / A->nx, nz = rows / A->ny;
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  for (int colour = 0; colour < 4; ++colour) {
    const int cy = colour & 1, cz = colour >> 1;
    const int my = (A->ny - cy + 1) / 2, mz = (nz - cz + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < my * mz; ++r) {
      const int iy = cy + 2 * (r % my), iz = cz + 2 * (r / my);
      const int row = iz * A->ny + iy, c0 = row * A->nx, nb = (A->lin_b || A->mode == 0 || A->mode == 3) ? A->nx / 4 : 0;
      for (int b = 0; b < nb; ++b)
        batch_vmult(A, c0 + 4 * b, A->lin_b ? A->lin_b + ((size_t)row * (A->nx / 4) + b) * (12 * 27 * 4) : 0, src_u, src_p, dst_u, dst_p);
      for (int cell = c0 + 4 * nb; cell < c0 + A->nx; ++cell) cell_vmult(A, cell, src_u, src_p, dst_u, dst_p);
    }
  }
  return threads;
}

/* explicit thread count: launchers such as torchrun export OMP_NUM_THREADS=1 to their workers */
void port_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int port_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
