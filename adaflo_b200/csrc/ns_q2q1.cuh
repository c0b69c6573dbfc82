// ns_q2q1.cuh -- sum-factorised cell kernels of the 3-D Taylor-Hood Q2/Q1
// Navier-Stokes operator (hand-written for sm_100a, fp64).
//
// What it computes: local_operation<1, ., LocalOps> of the reference
// (source/navier_stokes_matrix.cc:601-916): gather, evaluation at the 27 Gauss
// points, per-point physics, integration, scatter-add.
//
// How (B200 mapping, see DESIGN.md "Kernel K1"):
//  * 9 threads own one cell; a warp owns 3 consecutive cells (27 active lanes) and
//    synchronises with __syncwarp only -- warps never wait on each other, so the
//    gather latency of one warp overlaps the fp64 work of the others.
//  * thread (a,b) of a cell owns a LINE of three values; every 1-D contraction
//    along the owned direction runs in registers (3x3 tables folded into the
//    instruction stream as immediates); between directions the lines are
//    transposed through a per-warp shared-memory tile with a plane stride of 12
//    doubles (bank-conflict free for all three access patterns).
//  * per-quadrature-point data (linearisation point, rho/mu/damping, general
//    Jacobians) are streamed once with coalesced 8-byte loads in the row layout
//    [cell][row][q]; Cartesian/affine geometry is one table row per cell.
//  * scatter is colour-free: red.global.add.f64 into the zeroed destination.
#pragma once

#include "common.cuh"
#include "hanging.cuh"

namespace ab200 {

constexpr int kWarpsPerBlock = 4;
constexpr int kCellsPerWarp = 3;
constexpr int kPS = 9;               // plane stride (doubles): dense 3x3x3 tile
constexpr int kAS = 3 * kPS;         // one scalar tile
constexpr int kCellSmem = 10 * kAS + 13;  // 283 doubles: cell stride = 11 (mod 16) minimises bank conflicts between the 3 cells of a warp
constexpr int kWarpSmem = kCellsPerWarp * kCellSmem;

// 1-D shape tables (values from the oracle's long-double evaluation, rounded once):
// Q2 Lagrange basis on {0,1/2,1} and its derivative at the 3 Gauss points, Q1 basis
// at the same points, Gauss weights on [0,1].
#define AB_TABLES_Q2G3                                                                                     \
  constexpr double N[3][3] = {{0.6872983346207416, 0.4, -0.08729833462074169},                            \
                              {0.0, 1.0, 0.0},                                                            \
                              {-0.08729833462074169, 0.4, 0.6872983346207416}};                           \
  constexpr double D[3][3] = {{-2.549193338482967, 3.0983866769659336, -0.5491933384829668},              \
                              {-1.0, 0.0, 1.0},                                                           \
                              {0.5491933384829668, -3.0983866769659336, 2.549193338482967}};              \
  constexpr double NP[3][2] = {{0.8872983346207417, 0.11270166537925831},                                 \
                               {0.5, 0.5},                                                                \
                               {0.11270166537925831, 0.8872983346207417}};                                \
  constexpr double W[3] = {0.2777777777777778, 0.4444444444444444, 0.2777777777777778};                   \
  (void)N; (void)D; (void)NP; (void)W;

__device__ __forceinline__ int tidx(int z, int y, int x) { return z * kPS + y * 3 + x; }

__device__ __forceinline__ void red_add(double *addr, double v)
{
  // no "memory" clobber: the destination is never read back by the kernel, and a clobber would
  // serialise the contraction chains around every scatter
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v));
}

// ---------------------------------------------------------------- evaluation
// in : col[c][k]  = dof values of component c on the column (ix=a, iy=b), k = iz
// out: val/dx/dy/dz[c][qy] at the points (qx=a, qy, qz=b); derivatives w.r.t. the
//      reference coordinates.
template <bool VALUES>
__device__ __forceinline__ void evaluate_u(double *s, int a, int b, unsigned mask, const double (&col)[3][3],
                                           double (&val)[3][3], double (&dx)[3][3], double (&dy)[3][3],
                                           double (&dz)[3][3])
{
  AB_TABLES_Q2G3
  __syncwarp(mask);  // tile may still be read by a previous phase
  // Z: thread = (ix=a, iy=b)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double t = 0, td = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        t = fma(N[q][k], col[c][k], t);
        td = fma(D[q][k], col[c][k], td);
      }
      s[(0 * 3 + c) * kAS + tidx(q, b, a)] = t;
      s[(1 * 3 + c) * kAS + tidx(q, b, a)] = td;
    }
  __syncwarp(mask);
  // X: thread = (iy=a, qz=b)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double A[3], B[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      A[i] = s[(0 * 3 + c) * kAS + tidx(b, a, i)];
      B[i] = s[(1 * 3 + c) * kAS + tidx(b, a, i)];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double t = 0, tx = 0, tz = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        t = fma(N[q][i], A[i], t);
        tx = fma(D[q][i], A[i], tx);
        tz = fma(N[q][i], B[i], tz);
      }
      s[(0 * 3 + c) * kAS + tidx(b, a, q)] = t;
      s[(1 * 3 + c) * kAS + tidx(b, a, q)] = tz;
      s[(2 * 3 + c) * kAS + tidx(b, a, q)] = tx;
    }
  }
  __syncwarp(mask);
  // Y: thread = (qx=a, qz=b)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double A[3], B[3], C[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      A[i] = s[(0 * 3 + c) * kAS + tidx(b, i, a)];
      B[i] = s[(1 * 3 + c) * kAS + tidx(b, i, a)];
      C[i] = s[(2 * 3 + c) * kAS + tidx(b, i, a)];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double v = 0, gy = 0, gz = 0, gx = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (VALUES) v = fma(N[q][i], A[i], v);
        gy = fma(D[q][i], A[i], gy);
        gz = fma(N[q][i], B[i], gz);
        gx = fma(N[q][i], C[i], gx);
      }
      val[c][q] = v;
      dx[c][q] = gx;
      dy[c][q] = gy;
      dz[c][q] = gz;
    }
  }
}

// in : tv/tx/ty/tz[c][qy] = integrands tested by the value / the reference gradient
//      at (qx=a, qy, qz=b), JxW included
// out: out[c][k] = cell-local result on the column (ix=a, iy=b)
template <bool VALUES>
__device__ __forceinline__ void integrate_u(double *s, int a, int b, unsigned mask, const double (&tv)[3][3],
                                            const double (&tx)[3][3], const double (&ty)[3][3],
                                            const double (&tz)[3][3], double (&out)[3][3])
{
  AB_TABLES_Q2G3
  // Y^T: thread = (qx=a, qz=b) writes the line it read in evaluate_u (no hazard)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double r = 0, rz = 0, rx = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (VALUES) r = fma(N[q][i], tv[c][q], r);
        r = fma(D[q][i], ty[c][q], r);
        rz = fma(N[q][i], tz[c][q], rz);
        rx = fma(N[q][i], tx[c][q], rx);
      }
      s[(0 * 3 + c) * kAS + tidx(b, i, a)] = r;
      s[(1 * 3 + c) * kAS + tidx(b, i, a)] = rz;
      s[(2 * 3 + c) * kAS + tidx(b, i, a)] = rx;
    }
  __syncwarp(mask);
  // X^T: thread = (iy=a, qz=b)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double R[3], RZ[3], RX[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      R[q] = s[(0 * 3 + c) * kAS + tidx(b, a, q)];
      RZ[q] = s[(1 * 3 + c) * kAS + tidx(b, a, q)];
      RX[q] = s[(2 * 3 + c) * kAS + tidx(b, a, q)];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double A = 0, B = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        A = fma(N[q][i], R[q], A);
        A = fma(D[q][i], RX[q], A);
        B = fma(N[q][i], RZ[q], B);
      }
      s[(0 * 3 + c) * kAS + tidx(b, a, i)] = A;
      s[(1 * 3 + c) * kAS + tidx(b, a, i)] = B;
    }
  }
  __syncwarp(mask);
  // Z^T: thread = (ix=a, iy=b)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double A[3], B[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      A[q] = s[(0 * 3 + c) * kAS + tidx(q, b, a)];
      B[q] = s[(1 * 3 + c) * kAS + tidx(q, b, a)];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double o = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        o = fma(N[q][k], A[q], o);
        o = fma(D[q][k], B[q], o);
      }
      out[c][k] = o;
    }
  }
}

// Q1 pressure: values at the 27 points.  pc[k] = dof values on the column (ix=a, iy=b), a,b < 2
__device__ __forceinline__ void evaluate_p(double *sp, int a, int b, unsigned mask, const double (&pc)[2], double (&pv)[3])
{
  AB_TABLES_Q2G3
  __syncwarp(mask);
  if (a < 2 && b < 2) {
#pragma unroll
    for (int q = 0; q < 3; ++q) sp[tidx(q, b, a)] = NP[q][0] * pc[0] + NP[q][1] * pc[1];
  }
  __syncwarp(mask);
  if (a < 2) {  // thread = (iy=a, qz=b)
    const double A0 = sp[tidx(b, a, 0)], A1 = sp[tidx(b, a, 1)];
#pragma unroll
    for (int q = 0; q < 3; ++q) sp[tidx(b, a, q)] = NP[q][0] * A0 + NP[q][1] * A1;
  }
  __syncwarp(mask);
  {  // thread = (qx=a, qz=b)
    const double A0 = sp[tidx(b, 0, a)], A1 = sp[tidx(b, 1, a)];
#pragma unroll
    for (int q = 0; q < 3; ++q) pv[q] = NP[q][0] * A0 + NP[q][1] * A1;
  }
}

__device__ __forceinline__ void integrate_p(double *sp, int a, int b, unsigned mask, const double (&tp)[3], double (&out)[2])
{
  AB_TABLES_Q2G3
#pragma unroll
  for (int i = 0; i < 2; ++i) sp[tidx(b, i, a)] = NP[0][i] * tp[0] + NP[1][i] * tp[1] + NP[2][i] * tp[2];
  __syncwarp(mask);
  if (a < 2) {
    const double R0 = sp[tidx(b, a, 0)], R1 = sp[tidx(b, a, 1)], R2 = sp[tidx(b, a, 2)];
#pragma unroll
    for (int i = 0; i < 2; ++i) sp[tidx(b, a, i)] = NP[0][i] * R0 + NP[1][i] * R1 + NP[2][i] * R2;
  }
  __syncwarp(mask);
  out[0] = out[1] = 0;
  if (a < 2 && b < 2) {
    const double A0 = sp[tidx(0, b, a)], A1 = sp[tidx(1, b, a)], A2 = sp[tidx(2, b, a)];
#pragma unroll
    for (int k = 0; k < 2; ++k) out[k] = NP[0][k] * A0 + NP[1][k] * A1 + NP[2][k] * A2;
  }
}

// sum over the 9 lanes of a cell (used for the element-constant pressure dof)
__device__ __forceinline__ double cell_sum9(double v, int lane_base, int t9, unsigned mask)
{
  // butterfly over 16 is not aligned with groups of 9: gather through shuffles from lane_base..lane_base+8
  double s = 0;
#pragma unroll
  for (int j = 0; j < 9; ++j) s += __shfl_sync(mask, v, lane_base + j);
  (void)t9;
  return s;
}

// ------------------------------------------------------------------ geometry
// K[f][e] = d xi_f / d x_e at a point, and JxW
template <int GEOM>
struct Geometry {
  double K[3][3];   // affine/general: full; cartesian: diagonal in K[f][f]
  double det;
  const double *kq; // general: [9][27] block of the cell
  const double *jq;

  __device__ __forceinline__ void init(const NSArgs &A, uint32_t cell)
  {
    if (GEOM == ADAFLO_B200_GEOM_CARTESIAN) {
      const double *row = A.geom_table + 4 * (size_t)A.geom_row[cell];
      K[0][0] = row[0]; K[1][1] = row[1]; K[2][2] = row[2]; det = row[3];
    } else if (GEOM == ADAFLO_B200_GEOM_AFFINE) {
      const double *row = A.geom_table + 10 * (size_t)A.geom_row[cell];
#pragma unroll
      for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int e = 0; e < 3; ++e) K[f][e] = row[3 * f + e];
      det = row[9];
    } else {
      kq = A.inv_jac;
      jq = A.jxw;
    }
    cell_ = cell;
  }
  uint32_t cell_;
  // load the point data (general geometry, decade-blocked arrays) and return JxW
  __device__ __forceinline__ double at(int qx, int qy, int qz, double w)
  {
    if (GEOM == ADAFLO_B200_GEOM_GENERAL) {
#pragma unroll
      for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int e = 0; e < 3; ++e) K[f][e] = __ldg(kq + coef_index(cell_, 3 * f + e, 9, qx, qy, qz));
      return __ldg(jq + coef_index(cell_, 0, 1, qx, qy, qz));
    }
    return det * w;
  }
  // same for the vmult kernel whose lanes are laid out like the blocks: slot = (3 qx + qz) * 30 + lane
  __device__ __forceinline__ void init_blk(const NSArgs &A, uint32_t cell) { init(A, cell); }
  __device__ __forceinline__ double at_blk(const NSArgs &A, size_t g, size_t slot, double w)
  {
    if (GEOM == ADAFLO_B200_GEOM_GENERAL) {
#pragma unroll
      for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int e = 0; e < 3; ++e) K[f][e] = __ldg(A.inv_jac + (g * 9 + (3 * f + e)) * 270 + slot);
      return __ldg(A.jxw + g * 270 + slot);
    }
    return det * w;
  }
  // real-space gradient row from reference-space derivatives: g[e] = sum_f r[f] K[f][e]
  __device__ __forceinline__ void to_real(const double (&r)[3], double (&g)[3]) const
  {
    if (GEOM == ADAFLO_B200_GEOM_CARTESIAN) {
#pragma unroll
      for (int e = 0; e < 3; ++e) g[e] = r[e] * K[e][e];
    } else {
#pragma unroll
      for (int e = 0; e < 3; ++e) g[e] = r[0] * K[0][e] + r[1] * K[1][e] + r[2] * K[2][e];
    }
  }
  // reference-space test integrand from a real-space flux row: t[f] = sum_e K[f][e] S[e] * JxW
  __device__ __forceinline__ void to_ref(const double (&S)[3], double jxw, double (&t)[3]) const
  {
    if (GEOM == ADAFLO_B200_GEOM_CARTESIAN) {
#pragma unroll
      for (int f = 0; f < 3; ++f) t[f] = K[f][f] * S[f] * jxw;
    } else {
#pragma unroll
      for (int f = 0; f < 3; ++f) t[f] = (K[f][0] * S[0] + K[f][1] * S[1] + K[f][2] * S[2]) * jxw;
    }
  }
};

// ------------------------------------------------------------- gather/scatter
__device__ __forceinline__ void gather_u(const uint32_t *cell_u, const double *__restrict__ vec, uint32_t cell, int t9,
                                         bool resolve, double (&col)[3][3], uint32_t (&node)[3])
{
#pragma unroll
  for (int k = 0; k < 3; ++k) node[k] = __ldg(cell_u + (size_t)cell * 27 + 9 * k + t9);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const size_t n = node[k] & ADAFLO_B200_NODE_MASK;
    const uint32_t flags = resolve ? (node[k] >> ADAFLO_B200_FLAG_SHIFT) : 0u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = vec[3 * n + c];
      col[c][k] = ((flags >> c) & 1u) ? 0.0 : v;
    }
  }
}

__device__ __forceinline__ void scatter_u(double *vec, const uint32_t (&node)[3], const double (&out)[3][3])
{
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const size_t n = node[k] & ADAFLO_B200_NODE_MASK;
    const uint32_t flags = node[k] >> ADAFLO_B200_FLAG_SHIFT;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (!((flags >> c) & 1u)) red_add(vec + 3 * n + c, out[c][k]);
  }
}

// ------------------------------------------------------------ vmult kernels
// LocalOps::vmult (WITH_P) and LocalOps::vmult_velocity (!WITH_P),
// navier_stokes_matrix.cc:702-907.
template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3) ns_vmult_kernel(const NSArgs A)
{
  AB_TABLES_Q2G3
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = lane / 9;
  const int t9 = lane - 9 * cw;
  const int a = t9 % 3, b = t9 / 3;
  const size_t triple = (size_t)blockIdx.x * kWarpsPerBlock + warp;
  const size_t cell_l = triple * kCellsPerWarp + cw;
  const bool active = (cw < kCellsPerWarp) && (cell_l < A.n_cells);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t cell = (uint32_t)cell_l;
  double *s = smem + warp * kWarpSmem + cw * kCellSmem;
  double *sp = s + 9 * kAS;

  // ---- gather (read_dof_values: constrained entries read as zero)
  double col[3][3];
  uint32_t node[3];
  gather_u(A.cell_u, A.src_u, cell, t9, true, col, node);
  double pc[2] = {0, 0};
  uint32_t pnode[2] = {0, 0};
  double p_dg0 = 0;
  uint32_t dg0_node = 0;
  if (WITH_P) {
    if (a < 2 && b < 2) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        pnode[k] = __ldg(A.cell_p + (size_t)cell * A.np_cell + 4 * k + 2 * b + a);
        const double v = A.src_p[pnode[k] & ADAFLO_B200_NODE_MASK];
        pc[k] = (pnode[k] >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : v;
      }
    }
    if (A.np_cell == 9) {
      dg0_node = __ldg(A.cell_p + (size_t)cell * 9 + 8);
      p_dg0 = A.src_p[dg0_node & ADAFLO_B200_NODE_MASK];
    }
  }
  Geometry<GEOM> G;
  G.init(A, cell);

  // ---- evaluate
  double val[3][3], dx[3][3], dy[3][3], dz[3][3];
  evaluate_u<MODE != MODE_STOKES>(s, a, b, mask, col, val, dx, dy, dz);
  double pv[3] = {0, 0, 0};
  if (WITH_P) {
    evaluate_p(sp, a, b, mask, pc, pv);
#pragma unroll
    for (int q = 0; q < 3; ++q) pv[q] += p_dg0;
  }

  // ---- quadrature point loop: points (qx=a, qy, qz=b)
  double tp[3];
#pragma unroll
  for (int qy = 0; qy < 3; ++qy) {
    const double *lin = A.lin + lin_index(cell, a, qy, b);  // row r at lin_row(r)
    const double jxw = G.at(a, qy, b, W[a] * W[qy] * W[b]);
    const size_t q1 = coef_index(cell, 0, 1, a, qy, b);  // rho / mu / damping
    double g[3][3];  // g[d][e] = d u_d / d x_e
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double r[3] = {dx[d][qy], dy[d][qy], dz[d][qy]};
      G.to_real(r, g[d]);
    }
    const double div = g[0][0] + g[1][1] + g[2][2];
    double conv[3] = {0, 0, 0};
    double Sadd[3][3];  // transposed convection tests the gradient (Tvmult only)
    if (MODE != MODE_STOKES) {
      const double u[3] = {val[0][qy], val[1][qy], val[2][qy]};
      const double rho = VAR ? __ldg(A.rho + q1) : A.rho0;
      const double damping = VAR ? __ldg(A.damp + q1) : A.damp0;
#pragma unroll
      for (int d = 0; d < 3; ++d) conv[d] = A.gamma * u[d];
      if (MODE == MODE_NEWTON) {
        double lu[3], lg[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) lu[d] = __ldg(lin + lin_row(d));
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int e = 0; e < 3; ++e) lg[d][e] = __ldg(lin + lin_row(3 + 3 * d + e));
        const double ltr = lg[0][0] + lg[1][1] + lg[2][2];
        if (!TRANSPOSE) {
          const double f1 = A.beta * div, f2 = A.beta * ltr;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            double res = f1 * lu[d] + f2 * u[d];
#pragma unroll
            for (int e = 0; e < 3; ++e) res += lu[e] * g[d][e] + u[e] * lg[d][e];
            conv[d] += A.tau1 * res;
          }
        } else {
          const double udl = u[0] * lu[0] + u[1] * lu[1] + u[2] * lu[2];
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            double res = A.beta * ltr * u[e];
#pragma unroll
            for (int d = 0; d < 3; ++d) res += u[d] * lg[d][e];
            conv[e] += A.tau1 * res;
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int e = 0; e < 3; ++e) Sadd[d][e] = A.tau1 * rho * u[d] * lu[e];
            Sadd[d][d] += A.tau1 * rho * A.beta * udl;
          }
        }
      } else if (MODE == MODE_PICARD) {
        double lu[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) lu[d] = __ldg(lin + lin_row(d));
        const double ldiv = __ldg(lin + lin_row(3));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double res = A.beta * ldiv * u[d];
          if (!TRANSPOSE) {
#pragma unroll
            for (int e = 0; e < 3; ++e) res += lu[e] * g[d][e];
          } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) Sadd[d][e] = A.tau1 * rho * u[d] * lu[e];
          }
          conv[d] += A.tau1 * res;
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) conv[d] = (conv[d] * rho - damping * u[d]) * jxw;
    }
    const double tmu = (VAR ? __ldg(A.mu + q1) : A.mu0) * A.tau1;
    double S[3][3];
    S[0][1] = S[1][0] = tmu * (g[0][1] + g[1][0]);
    S[0][2] = S[2][0] = tmu * (g[0][2] + g[2][0]);
    S[1][2] = S[2][1] = tmu * (g[1][2] + g[2][1]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      S[d][d] = 2. * tmu * g[d][d] + A.tau_gd * div;
      if (WITH_P) S[d][d] -= pv[qy];
    }
    if (TRANSPOSE && (MODE == MODE_NEWTON || MODE == MODE_PICARD)) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int e = 0; e < 3; ++e) S[d][e] += Sadd[d][e];
    }
    tp[qy] = -div * jxw;
    // overwrite the evaluation registers with the integrands
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double t[3];
      G.to_ref(S[d], jxw, t);
      val[d][qy] = conv[d];
      dx[d][qy] = t[0];
      dy[d][qy] = t[1];
      dz[d][qy] = t[2];
    }
  }

  // ---- integrate + scatter
  double out[3][3];
  integrate_u<MODE != MODE_STOKES>(s, a, b, mask, val, dx, dy, dz, out);
  scatter_u(A.dst_u, node, out);
  if (WITH_P && !A.skip_p_test) {
    double outp[2];
    integrate_p(sp, a, b, mask, tp, outp);
    if (a < 2 && b < 2) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (!(pnode[k] >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (pnode[k] & ADAFLO_B200_NODE_MASK), outp[k]);
    }
    if (A.np_cell == 9) {
      const double sum = cell_sum9(tp[0] + tp[1] + tp[2], 9 * cw, t9, mask);
      if (t9 == 0 && !(dg0_node >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (dg0_node & ADAFLO_B200_NODE_MASK), sum);
    }
  }
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE>
int launch_vmult_inst(const NSArgs &a, cudaStream_t s)
{
  auto kern = ns_vmult_kernel<WITH_P, MODE, GEOM, VAR, TRANSPOSE>;
  const size_t smem = sizeof(double) * kWarpsPerBlock * kWarpSmem;
  {  // per device and cheap: set on every launch (a process may drive several GPUs)
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const size_t cells_per_block = (size_t)kWarpsPerBlock * kCellsPerWarp;
  const unsigned grid = (unsigned)((a.n_cells + cells_per_block - 1) / cells_per_block);
  if (grid == 0) return 0;
  kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(a);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

template <int GEOM>
int launch_vmult_geom(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  if (a.hn_u) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "the v1 vmult kernel has no hanging-node support");
  const bool tr = a.transpose && (mode == MODE_NEWTON || mode == MODE_PICARD);  // other modes are symmetric
  if (tr) {
    if (!with_p) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "transposed velocity block not instantiated");
    if (mode == MODE_NEWTON && !var) return launch_vmult_inst<true, MODE_NEWTON, GEOM, false, true>(a, s);
    if (mode == MODE_NEWTON && var) return launch_vmult_inst<true, MODE_NEWTON, GEOM, true, true>(a, s);
    if (mode == MODE_PICARD && !var) return launch_vmult_inst<true, MODE_PICARD, GEOM, false, true>(a, s);
    return launch_vmult_inst<true, MODE_PICARD, GEOM, true, true>(a, s);
  }
#define AB_CASE(P, M, V) \
  if (with_p == P && mode == M && var == V) return launch_vmult_inst<P, M, GEOM, V, false>(a, s);
  AB_CASE(true, MODE_STOKES, false)
  AB_CASE(true, MODE_NEWTON, false)
  AB_CASE(true, MODE_PICARD, false)
  AB_CASE(true, MODE_NOCONV, false)
  AB_CASE(true, MODE_STOKES, true)
  AB_CASE(true, MODE_NEWTON, true)
  AB_CASE(true, MODE_PICARD, true)
  AB_CASE(true, MODE_NOCONV, true)
  AB_CASE(false, MODE_STOKES, false)
  AB_CASE(false, MODE_NEWTON, false)
  AB_CASE(false, MODE_PICARD, false)
  AB_CASE(false, MODE_NOCONV, false)
  AB_CASE(false, MODE_STOKES, true)
  AB_CASE(false, MODE_NEWTON, true)
  AB_CASE(false, MODE_PICARD, true)
  AB_CASE(false, MODE_NOCONV, true)
#undef AB_CASE
  return fail(ADAFLO_B200_ERR_ARG, "launch_vmult_geom: bad variant");
}

}  // namespace ab200
