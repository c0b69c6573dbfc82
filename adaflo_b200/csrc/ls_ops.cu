// ls_ops.cu -- matrix-free level-set operators on FE_Q_iso_Q1(n_sub) with
// QIterated(QGauss(2), n_sub):
//   advection            source/level_set_okz_advance_concentration.cc:217-258, 288-397
//   normal projection    source/level_set_okz_compute_normal.cc:82-156
//   curvature projection source/level_set_okz_compute_curvature.cc:86-133, 212-259
//   reinitialisation     source/level_set_okz_reinitialization.cc:53-106, 128-189
//   mass diagonal        include/adaflo/level_set_okz_preconditioner.h:31-76
//
// FE_Q_iso_Q1 hats with the composite Gauss rule are exactly trilinear elements on the n_sub^3
// sub-cells with 2x2x2 Gauss points each, so one THREAD owns one SUB-CELL: 8 nodal values in
// registers, 8 quadrature points, no shared memory, no synchronisation.  Per-point streams
// (convection velocity, frozen normals) are stored [cell][sub-cell][component][8 points], i.e. 64
// contiguous bytes per thread and component; scatter is red.global.add.f64.
#include <cstdlib>

#include "ls_ops.cuh"
#include "ns_q2q1.cuh"

namespace ab200 {

__device__ __forceinline__ double g2(int p) { return p ? 0.7886751345948129 : 0.2113248654051871; }

template <int OP>
__global__ void __launch_bounds__(128) ls_subcell_kernel(const LSArgs A)
{
  const int S = A.n_sub, S3 = S * S * S, n1 = S + 1, N = n1 * n1 * n1;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)A.n_cells * S3) return;
  const uint32_t cell = (uint32_t)(gid / S3);
  const int sub = (int)(gid - (size_t)cell * S3);
  const int sx = sub % S, sy = (sub / S) % S, sz = sub / (S * S);

  // geometry of the cell (Cartesian or affine); the sub-cell is the cell scaled by 1 / n_sub
  double K[3][3] = {{0}};
  double det;
  const bool general = A.geom == ADAFLO_B200_GEOM_GENERAL;  // K and JxW per point, loaded in the point loop
  if (general) {
    det = 0;
  } else if (A.geom == ADAFLO_B200_GEOM_CARTESIAN) {
    const double *row = A.geom_table + 4 * (size_t)A.geom_row[cell];
    K[0][0] = row[0]; K[1][1] = row[1]; K[2][2] = row[2]; det = row[3];
  } else {
    const double *row = A.geom_table + 10 * (size_t)A.geom_row[cell];
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int e = 0; e < 3; ++e) K[f][e] = row[3 * f + e];
    det = row[9];
  }
  double jxw = det / (8.0 * S3);  // Gauss(2) weights: 1/2 per direction on the sub-cell
  const double dS = (double)S;          // slope of the hat functions in cell reference coordinates

  uint32_t node[8];
#pragma unroll
  for (int n = 0; n < 8; ++n)
    node[n] = __ldg(A.cell_ls + (size_t)cell * N + (sx + (n & 1)) + n1 * ((sy + ((n >> 1) & 1)) + n1 * (sz + (n >> 2))));

  const HnInfo hn(A.hn ? __ldg(A.hn + cell) : 0u);
  const uint32_t *cl = A.cell_ls + (size_t)cell * N;
  // nodal read / scatter-add of this thread's node n (resolves hanging-node constraints on the fly)
  auto rd = [&](const double *vec, int n, bool res) -> double {
    if (!hn.any()) {
      const double v = vec[node[n] & ADAFLO_B200_NODE_MASK];
      return (res && (node[n] >> ADAFLO_B200_FLAG_SHIFT)) ? 0.0 : v;
    }
    LsStencil st;
    ls_stencil(hn, cl, S, sx + (n & 1), sy + ((n >> 1) & 1), sz + (n >> 2), st);
    double v = 0;
    for (int i = 0; i < st.n; ++i)
      if (!(res && (st.slot[i] >> ADAFLO_B200_FLAG_SHIFT))) v = fma(st.w[i], vec[st.slot[i] & ADAFLO_B200_NODE_MASK], v);
    return v;
  };
  auto add = [&](double *vec, int n, double v, bool skip_flagged) {
    if (!hn.any()) {
      if (!(skip_flagged && (node[n] >> ADAFLO_B200_FLAG_SHIFT))) red_add(vec + (node[n] & ADAFLO_B200_NODE_MASK), v);
      return;
    }
    LsStencil st;
    ls_stencil(hn, cl, S, sx + (n & 1), sy + ((n >> 1) & 1), sz + (n >> 2), st);
    for (int i = 0; i < st.n; ++i)
      if (!(skip_flagged && (st.slot[i] >> ADAFLO_B200_FLAG_SHIFT))) red_add(vec + (st.slot[i] & ADAFLO_B200_NODE_MASK), st.w[i] * v);
  };
  const bool resolve = (OP == LS_ADVECT || OP == LS_DAMPED || OP == LS_REINIT || OP == LS_REINIT_DIFFUSE);
  const int n_fields = (OP == LS_DAMPED) ? A.n_comp : 1;
  // per-point streams: [cell][component][point p][sub-cell] (the layout of the cell kernel below: a warp reads S3 contiguous doubles)
  double *pd = A.pointdata ? A.pointdata + (size_t)cell * S3 * 24 + sub : nullptr;
  const int ps = S3;  // stride between the points of one sub-cell
  const double h_rel = A.cell_diameters ? A.cell_diameters[cell] / dS : 0.0;
  const double dd = fmax(A.eps_ratio, h_rel);
  const double delta = A.delta_scale * dd * dd;          // compute_normal.cc:105-108, compute_curvature.cc:112-117
  const double D = fmax(A.epsilon_used, h_rel);           // reinitialization.cc:77-79

  // values of auxiliary nodal fields
  double c0[8], aux0[8], aux1[8], aux2[8];
  if (OP == LS_CURV_RHS || OP == LS_REINIT_RHS || OP == LS_ADVECT_RHS) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      aux0[n] = A.aux[0] ? rd(A.aux[0], n, false) : 0.0;
      aux1[n] = A.aux[1] ? rd(A.aux[1], n, false) : 0.0;
      aux2[n] = A.aux[2] ? rd(A.aux[2], n, false) : 0.0;
    }
  }
  if (OP == LS_CURV_RHS) {  // dof-level normalisation, compute_curvature.cc:231-240
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double nn = sqrt(aux0[n] * aux0[n] + aux1[n] * aux1[n] + aux2[n] * aux2[n]);
      const double inv = nn > 1e-2 ? 1.0 / nn : 0.0;
      aux0[n] = nn > 1e-2 ? aux0[n] / nn : 0.0;
      aux1[n] = nn > 1e-2 ? aux1[n] / nn : 0.0;
      aux2[n] = nn > 1e-2 ? aux2[n] / nn : 0.0;
      (void)inv;
    }
  }
  // Q2 velocity at the 8 points (advect rhs): tensor-product Lagrange basis on {0,1/2,1}
  double uq[8][3];
  if (OP == LS_ADVECT_RHS) {
    double L[3][2][3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const double x = ((d == 0 ? sx : (d == 1 ? sy : sz)) + g2(p)) / dS;
        L[d][p][0] = 2 * (x - 0.5) * (x - 1);
        L[d][p][1] = 4 * x * (1 - x);
        L[d][p][2] = 2 * x * (x - 0.5);
      }
#pragma unroll
    for (int p = 0; p < 8; ++p) uq[p][0] = uq[p][1] = uq[p][2] = 0;
    // velocity space with hanging nodes (read_dof_values_plain of a distributed vector = the interpolated
    // values): the cell list holds the parent's nodes on hanging faces / edges, interpolate line by line, all
    // lines of one direction before the next (same order as the Navier-Stokes cell kernels)
    const HnInfo hnu(A.hn_u ? __ldg(A.hn_u + cell) : 0u);
    double U[3][27];
    if (hnu.any()) {
      for (int n = 0; n < 27; ++n) {
        const size_t un = __ldg(A.cell_u + (size_t)cell * 27 + n) & ADAFLO_B200_NODE_MASK;
        for (int c = 0; c < 3; ++c) U[c][n] = A.vel[3 * un + c];
      }
      const int stride[3] = {1, 3, 9};
      for (int d = 0; d < 3; ++d) {
        const int e = (d + 1) % 3, f = (d + 2) % 3;
        for (int ie = 0; ie < 3; ++ie)
          for (int jf = 0; jf < 3; ++jf)
            if (hnu.line(d, ie, jf, 3)) {
              const int n0 = ie * stride[e] + jf * stride[f];
              for (int c = 0; c < 3; ++c) hn_q2(hnu.sub(d), U[c][n0], U[c][n0 + stride[d]], U[c][n0 + 2 * stride[d]]);
            }
      }
    }
    for (int n = 0; n < 27; ++n) {
      double u0, u1, u2;
      if (hnu.any()) {
        u0 = U[0][n]; u1 = U[1][n]; u2 = U[2][n];
      } else {
        const size_t un = __ldg(A.cell_u + (size_t)cell * 27 + n) & ADAFLO_B200_NODE_MASK;
        u0 = A.vel[3 * un]; u1 = A.vel[3 * un + 1]; u2 = A.vel[3 * un + 2];
      }
      const int i = n % 3, j = (n / 3) % 3, k = n / 9;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const double w = L[0][p & 1][i] * L[1][(p >> 1) & 1][j] * L[2][p >> 2][k];
        uq[p][0] = fma(w, u0, uq[p][0]);
        uq[p][1] = fma(w, u1, uq[p][1]);
        uq[p][2] = fma(w, u2, uq[p][2]);
      }
    }
  }

  for (int f = 0; f < n_fields; ++f) {
    double out0[8], out1[8], out2[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) out0[n] = out1[n] = out2[n] = 0;
    if (OP != LS_MASS_DIAG && OP != LS_CURV_RHS) {
#pragma unroll
      for (int n = 0; n < 8; ++n) c0[n] = rd(A.src[f], n, resolve);
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const double vx[2] = {1 - g2(p & 1), g2(p & 1)}, vy[2] = {1 - g2((p >> 1) & 1), g2((p >> 1) & 1)}, vz[2] = {1 - g2(p >> 2), g2(p >> 2)};
      if (general) {  // MappingQ cells: inverse Jacobian and JxW of this point
        const int nq1 = 2 * S;
        const size_t NQ = (size_t)nq1 * nq1 * nq1;
        const size_t q = (2 * sx + (p & 1)) + (size_t)nq1 * ((2 * sy + ((p >> 1) & 1)) + (size_t)nq1 * (2 * sz + (p >> 2)));
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
          for (int e = 0; e < 3; ++e) K[f][e] = __ldg(A.inv_jac + ((size_t)cell * 9 + 3 * f + e) * NQ + q);
        jxw = __ldg(A.jxw_q + (size_t)cell * NQ + q);
      }
      double psi[8], gp[8][3];  // shape values and real-space gradients
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int a = n & 1, b = (n >> 1) & 1, c = n >> 2;
        psi[n] = vx[a] * vy[b] * vz[c];
        const double r0 = (a ? dS : -dS) * vy[b] * vz[c], r1 = vx[a] * (b ? dS : -dS) * vz[c], r2 = vx[a] * vy[b] * (c ? dS : -dS);
        if (A.geom == ADAFLO_B200_GEOM_CARTESIAN) {
          gp[n][0] = r0 * K[0][0]; gp[n][1] = r1 * K[1][1]; gp[n][2] = r2 * K[2][2];
        } else {
#pragma unroll
          for (int e = 0; e < 3; ++e) gp[n][e] = r0 * K[0][e] + r1 * K[1][e] + r2 * K[2][e];
        }
      }
      double v = 0, g[3] = {0, 0, 0};
      if (OP != LS_MASS_DIAG && OP != LS_CURV_RHS) {
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          v = fma(psi[n], c0[n], v);
#pragma unroll
          for (int e = 0; e < 3; ++e) g[e] = fma(gp[n][e], c0[n], g[e]);
        }
      }
      double tv = 0, tg[3] = {0, 0, 0}, tv3[3] = {0, 0, 0};
      if (OP == LS_ADVECT) {
        tv = A.gamma * v + g[0] * pd[(0 + p) * ps] + g[1] * pd[(8 + p) * ps] + g[2] * pd[(16 + p) * ps];
      } else if (OP == LS_DAMPED) {
        tv = v;
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = delta * g[e];
      } else if (OP == LS_REINIT) {
        const double nq[3] = {pd[(0 + p) * ps], pd[(8 + p) * ps], pd[(16 + p) * ps]};
        const double ng = nq[0] * g[0] + nq[1] * g[1] + nq[2] * g[2];
        tv = A.dtau_inv * v;
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = D * ng * nq[e];
      } else if (OP == LS_REINIT_DIFFUSE) {
        tv = A.dtau_inv * v;
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = D * g[e];
      } else if (OP == LS_NORMAL_RHS) {
#pragma unroll
        for (int e = 0; e < 3; ++e) tv3[e] = g[e];
      } else if (OP == LS_CURV_RHS) {
        double div = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) div += gp[n][0] * aux0[n] + gp[n][1] * aux1[n] + gp[n][2] * aux2[n];
        tv = -div;
      } else if (OP == LS_REINIT_RHS) {
        double nq[3];
        if (A.first_step) {  // freeze the normal of the first pseudo time step, reinitialization.cc:167-172
          double nv[3] = {0, 0, 0};
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            nv[0] = fma(psi[n], aux0[n], nv[0]);
            nv[1] = fma(psi[n], aux1[n], nv[1]);
            nv[2] = fma(psi[n], aux2[n], nv[2]);
          }
          const double nn = fmax(1e-4, sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]));
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            nq[e] = nv[e] / nn;
            pd[(8 * e + p) * ps] = nq[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 3; ++e) nq[e] = pd[(8 * e + p) * ps];
        }
        const double ng = nq[0] * g[0] + nq[1] * g[1] + nq[2] * g[2];
        const double fl = 0.5 * (1. - v * v) - ng * D;
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = nq[e] * fl;
      } else if (OP == LS_REINIT_RHS_DIFFUSE) {
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = -D * g[e];
      } else if (OP == LS_ADVECT_RHS) {
        double vo = 0, voo = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          vo = fma(psi[n], aux0[n], vo);
          voo = fma(psi[n], aux1[n], voo);
        }
        tv = -(A.gamma * v + uq[p][0] * g[0] + uq[p][1] * g[1] + uq[p][2] * g[2] + A.w_old * vo + A.w_oo * voo);
#pragma unroll
        for (int e = 0; e < 3; ++e) pd[(8 * e + p) * ps] = uq[p][e];  // evaluated_convection, :389
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (OP == LS_MASS_DIAG)
          out0[n] = fma(psi[n] * psi[n], jxw, out0[n]);
        else if (OP == LS_NORMAL_RHS) {
          out0[n] = fma(psi[n] * jxw, tv3[0], out0[n]);
          out1[n] = fma(psi[n] * jxw, tv3[1], out1[n]);
          out2[n] = fma(psi[n] * jxw, tv3[2], out2[n]);
        } else
          out0[n] += (psi[n] * tv + gp[n][0] * tg[0] + gp[n][1] * tg[1] + gp[n][2] * tg[2]) * jxw;
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const bool skip = OP != LS_MASS_DIAG;
      if (OP == LS_NORMAL_RHS) {
        add(A.dst[0], n, out0[n], skip);
        add(A.dst[1], n, out1[n], skip);
        add(A.dst[2], n, out2[n], skip);
      } else
        add(A.dst[f], n, out0[n], skip);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Cell kernel (n_sub >= 3): one WARP owns one cell.  profiles/r2_ls_a: with one thread per sub-cell every dof is
// gathered and scattered 8 times (453 M red.add per operator at 128^3 x 27 sub-cells) and the trilinear shape data is
// recomputed per point; the kernels ran at 2-23 % of the HBM peak.  Here
//  * the (n_sub+1)^3 nodal values of the cell are gathered once (coalesced list reads) into shared memory, and the
//    cell's result is accumulated in shared memory and scattered once: (n_sub+1)^3 instead of 8 n_sub^3 atomics per cell;
//  * lane = sub-cell: the trilinear element on the sub-cell is evaluated and integrated by sum factorisation (values at
//    the 8 Gauss points plus the 12 distinct gradient entries: 54 flops instead of 256), the tables are immediates;
//  * the per-point streams are stored [cell][component][point][sub-cell]: a warp-level load is n_sub^3 contiguous doubles;
//  * the sub-cell results are added to the cell's shared-memory vector corner by corner: for a fixed corner the nodes of
//    different sub-cells are distinct, so 8 plain read-add-write passes replace shared-memory atomics.
// Hanging nodes are resolved per cell slot at gather / scatter time with the same stencils as above.
namespace lsq {
constexpr double g0 = 0.2113248654051871, g1 = 0.7886751345948129;

// cv[a + 2b + 4c] -> v[p + 2q + 4r] and the gradient w.r.t. the coordinates of the unit sub-cell:
// d/dx = dx[q + 2r], d/dy = dy[p + 2r], d/dz = dz[p + 2q] (constant along their own direction)
__device__ __forceinline__ void eval(const double (&cv)[8], double (&v)[8], double (&dx)[4], double (&dy)[4], double (&dz)[4])
{
  double vx[2][4], ddx[4];
#pragma unroll
  for (int bc = 0; bc < 4; ++bc) {
    const double lo = cv[2 * bc], d = cv[2 * bc + 1] - lo;
    ddx[bc] = d;
    vx[0][bc] = fma(g0, d, lo);
    vx[1][bc] = fma(g1, d, lo);
  }
  double vxy[2][2][2], ddy[2][2], dxy[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double lo = vx[p][2 * c], d = vx[p][2 * c + 1] - lo;
      ddy[p][c] = d;
      vxy[p][0][c] = fma(g0, d, lo);
      vxy[p][1][c] = fma(g1, d, lo);
    }
    const double lo = ddx[2 * c], d = ddx[2 * c + 1] - lo;
    dxy[0][c] = fma(g0, d, lo);
    dxy[1][c] = fma(g1, d, lo);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double lo = vxy[p][q][0], d = vxy[p][q][1] - lo;
      dz[p + 2 * q] = d;
      v[p + 2 * q] = fma(g0, d, lo);
      v[p + 2 * q + 4] = fma(g1, d, lo);
    }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const double lo = ddy[p][0], d = ddy[p][1] - lo;
    dy[p] = fma(g0, d, lo);
    dy[p + 2] = fma(g1, d, lo);
    const double lo2 = dxy[p][0], d2 = dxy[p][1] - lo2;
    dx[p] = fma(g0, d2, lo2);
    dx[p + 2] = fma(g1, d2, lo2);
  }
}

// out[a + 2b + 4c] = sum_points psi_a psi_b psi_c tv + psi_a' psi_b psi_c tx + psi_a psi_b' psi_c ty + psi_a psi_b psi_c' tz
// (hats on the unit sub-cell: psi_0 = 1 - x, psi_1 = x, psi' = -1 / +1); V / G select the value and the gradient part
template <bool V, bool G>
__device__ __forceinline__ void integrate(const double (&tv)[8], const double (&tx)[8], const double (&ty)[8], const double (&tz)[8],
                                          double (&out)[8])
{
  // sum over the two points of a direction against psi_0 / psi_1: g1 lo + g0 hi = lo + g0 (hi - lo), g0 lo + g1 hi = lo + g1 (hi - lo)
  double A[2][4], X[2][4], Y[2][4];  // [c][p + 2q]
#pragma unroll
  for (int pq = 0; pq < 4; ++pq) {
    double a0 = 0, a1 = 0;
    if (V) {
      const double d = tv[pq + 4] - tv[pq];
      a0 = fma(g0, d, tv[pq]);
      a1 = fma(g1, d, tv[pq]);
    }
    if (G) {
      const double zs = tz[pq] + tz[pq + 4];
      a0 -= zs;
      a1 += zs;
      const double dX = tx[pq + 4] - tx[pq], dY = ty[pq + 4] - ty[pq];
      X[0][pq] = fma(g0, dX, tx[pq]);
      X[1][pq] = fma(g1, dX, tx[pq]);
      Y[0][pq] = fma(g0, dY, ty[pq]);
      Y[1][pq] = fma(g1, dY, ty[pq]);
    }
    A[0][pq] = a0;
    A[1][pq] = a1;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    double B[2][2], XX[2][2];  // [b][p]
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double d = A[c][p + 2] - A[c][p];
      double b0 = fma(g0, d, A[c][p]), b1 = fma(g1, d, A[c][p]);
      if (G) {
        const double ys = Y[c][p] + Y[c][p + 2];
        b0 -= ys;
        b1 += ys;
        const double dX = X[c][p + 2] - X[c][p];
        XX[0][p] = fma(g0, dX, X[c][p]);
        XX[1][p] = fma(g1, dX, X[c][p]);
      }
      B[0][p] = b0;
      B[1][p] = b1;
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const double d = B[b][1] - B[b][0];
      double o0 = fma(g0, d, B[b][0]), o1 = fma(g1, d, B[b][0]);
      if (G) {
        const double xs = XX[b][0] + XX[b][1];
        o0 -= xs;
        o1 += xs;
      }
      out[2 * b + 4 * c] = o0;
      out[1 + 2 * b + 4 * c] = o1;
    }
  }
}

// Mass + scaled Laplace operator of the trilinear element on a Cartesian sub-cell, applied as a tensor product of the
// 2x2 matrices M = [[2, 1], [1, 2]] / 6 and K = [[1, -1], [-1, 1]] (exactly what the 2-point Gauss rule integrates):
//   out = vol (M x M x M + kx K x M x M + ky M x K x M + kz M x M x K) cv
// with c0 = vol / 216 and ck[e] = vol k_e / 36.  98 additions / fmas instead of evaluate -> 8 points -> integrate.
__device__ __forceinline__ void mass_laplace(const double (&cv)[8], const double c0, const double (&ck)[3], double (&out)[8])
{
  double mz[8], kz[4];  // M_z cv (times 6), K_z cv (entry c = 1; c = 0 is its negative), index a + 2b
#pragma unroll
  for (int ab = 0; ab < 4; ++ab) {
    const double lo = cv[ab], hi = cv[ab + 4], s = lo + hi;
    mz[ab] = s + lo;
    mz[ab + 4] = s + hi;
    kz[ab] = hi - lo;
  }
  double mm[8], km[4], mk[4];  // M_y M_z (x36) [a + 2b + 4c], K_y M_z (x6) [a + 2c] (b = 1), M_y K_z (x6) [a + 2b] (c = 1)
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double lo = mz[a + 4 * c], hi = mz[a + 2 + 4 * c], s = lo + hi;
      mm[a + 4 * c] = s + lo;
      mm[a + 2 + 4 * c] = s + hi;
      km[a + 2 * c] = hi - lo;
    }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const double lo = kz[a], hi = kz[a + 2], s = lo + hi;
    mk[a] = s + lo;
    mk[a + 2] = s + hi;
  }
  // x direction
  double mmm[8], kmm[4], mkm[4], mmk[4];  // kmm [b + 2c] (a = 1), mkm [a + 2c] (b = 1), mmk [a + 2b] (c = 1)
#pragma unroll
  for (int bc = 0; bc < 4; ++bc) {
    const double lo = mm[2 * bc], hi = mm[2 * bc + 1], s = lo + hi;
    mmm[2 * bc] = s + lo;
    mmm[2 * bc + 1] = s + hi;
    kmm[bc] = hi - lo;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    {
      const double lo = km[2 * i], hi = km[2 * i + 1], s = lo + hi;  // c = i
      mkm[2 * i] = s + lo;
      mkm[2 * i + 1] = s + hi;
    }
    {
      const double lo = mk[2 * i], hi = mk[2 * i + 1], s = lo + hi;  // b = i
      mmk[2 * i] = s + lo;
      mmk[2 * i + 1] = s + hi;
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double r = c0 * mmm[a + 2 * b + 4 * c];
        r = fma(a ? ck[0] : -ck[0], kmm[b + 2 * c], r);
        r = fma(b ? ck[1] : -ck[1], mkm[a + 2 * c], r);
        r = fma(c ? ck[2] : -ck[2], mmk[a + 2 * b], r);
        out[a + 2 * b + 4 * c] = r;
      }
}
}  // namespace lsq

template <int OP>
struct LsCellTraits {
  static constexpr bool resolve = (OP == LS_ADVECT || OP == LS_DAMPED || OP == LS_REINIT || OP == LS_REINIT_DIFFUSE);
  static constexpr bool has_src = (OP != LS_MASS_DIAG && OP != LS_CURV_RHS);
  static constexpr bool has_aux = (OP == LS_CURV_RHS || OP == LS_REINIT_RHS || OP == LS_ADVECT_RHS);
  static constexpr int n_src = OP == LS_DAMPED ? 3 : 1;    // source fields staged per cell (compute_normal: 3 components)
  static constexpr int n_in = n_src + (has_aux ? 3 : 0);   // nodal input arrays in shared memory (src fields, aux 0..2)
  static constexpr int n_out = OP == LS_NORMAL_RHS ? 3 : 1;
  static constexpr bool has_u = OP == LS_ADVECT_RHS;
  static constexpr bool tma_pd = (OP == LS_ADVECT || OP == LS_REINIT);  // per-point stream read-only: staged by the TMA engine
  static constexpr bool needs_diam = (OP == LS_DAMPED || OP == LS_REINIT || OP == LS_REINIT_DIFFUSE || OP == LS_REINIT_RHS ||
                                      OP == LS_REINIT_RHS_DIFFUSE);
};

template <int OP, int S>
__host__ __device__ constexpr int ls_cell_warps()  // warps per block: the double-buffered per-point stream of n_sub = 4 is 24 KB per warp
{
  // n_sub = 3: 15 KB per warp with the stream buffers -> 7 warps per block let two blocks (14 warps) share an SM
  return LsCellTraits<OP>::tma_pd ? (S == 4 ? 4 : 7) : 8;
}
template <int OP, int S>
__host__ __device__ constexpr int ls_cell_warp_doubles()
{
  constexpr int NP = ((S + 1) * (S + 1) * (S + 1) + (S + 1) + 7) / 8 * 8;  // node planes padded by one entry (bank conflicts)
  return LsCellTraits<OP>::n_in * NP + LsCellTraits<OP>::n_out * 8 * NP + (LsCellTraits<OP>::has_u ? 88 : 0) +
         (LsCellTraits<OP>::tma_pd ? 2 * S * S * S * 24 + 2 : 0);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bulk copy global -> shared memory by the TMA engine, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Persistent: warp w of the grid owns the cells w, w + W, w + 2 W, ... (W = warps of the grid), so that the cells in
// flight are neighbours in memory.  The dependent global round trips of a cell (node list -> nodal values, geometry row
// -> table) are taken off the critical path by a register pipeline: node list, geometry row and cell diameter are loaded
// two cells ahead, the source values one cell ahead (conforming cells; cells with hanging nodes resolve their stencils
// when they are processed).  The read-only per-point stream of advection / reinitialisation (convection velocity,
// frozen normal: 24 B per point) is staged one cell ahead into a double-buffered shared-memory slot of the warp by the
// TMA engine (cp.async.bulk + mbarrier: no registers, no LSU global wavefronts, the latency is off the critical path).
template <int OP, int S, bool CART>
__global__ void __launch_bounds__(ls_cell_warps<OP, S>() * 32, (OP == LS_ADVECT_RHS || OP == LS_REINIT_RHS) ? 1 : (ls_cell_warps<OP, S>() == 4 ? 4 : 2))
    ls_cell_kernel(const LSArgs A)
{
  using T = LsCellTraits<OP>;
  // shared-memory arrays of nodal values: node (ix, iy, iz) at ix + n1 iy + PS iz with the plane stride PS = n1^2 + 1 -- with
  // PS = n1^2 = 16 (n_sub = 3) the sub-cells (sx, sy, 0..2) of a warp hit one bank (3-way conflicts on every corner access)
  constexpr int S3 = S * S * S, n1 = S + 1, N = n1 * n1 * n1, PS = n1 * n1 + 1, NP = (N + n1 + 7) / 8 * 8;
  constexpr int NIT = (S3 + 31) / 32, NSL = (N + 31) / 32, WARPS = ls_cell_warps<OP, S>();
  auto sidx = [](const int slot) { return slot + slot / (n1 * n1); };  // cell slot (list order) -> shared-memory index
  constexpr unsigned kPdBytes = S3 * 24 * 8;
  extern __shared__ __align__(16) double ls_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * WARPS;
  size_t cell_l = (size_t)blockIdx.x * WARPS + warp;
  if (cell_l >= A.n_cells) return;  // warp-uniform
  double *cin = ls_smem + (size_t)warp * ls_cell_warp_doubles<OP, S>();  // [n_in][NP]
  double *cbuf = cin + T::n_in * NP;                                      // [n_out][8 corners][NP nodes]
  double *us = cbuf + T::n_out * 8 * NP;                                  // [3][27] Q2 velocity of the cell
  double *pdbuf = us + (T::has_u ? 88 : 0);                               // [2 stages][3][8][S3]
  uint64_t *bar = reinterpret_cast<uint64_t *>(pdbuf + 2 * S3 * 24);      // [2]
  const bool general = !CART && A.geom == ADAFLO_B200_GEOM_GENERAL;
  constexpr double dS = (double)S;
  const int n_fields = (OP == LS_DAMPED) ? A.n_comp : 1;

  // corner buffers: entry (corner n, node) is written by the sub-cell that has the node as its corner n; the entries
  // without such a sub-cell (nodes on the cell boundary) are never written and stay zero
  for (int i = lane; i < T::n_out * 8 * NP; i += 32) cbuf[i] = 0.0;
  if (T::tma_pd) {
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_init(bar + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncwarp();

  // register pipeline: node lists of this cell, the next and the one after; source values of this cell
  uint32_t nd[NSL], nd1[NSL], nd2[NSL];
  uint32_t hm = 0, hm1 = 0, hm2 = 0, row = 0, row1 = 0, row2 = 0;
  double diam = 0, diam1 = 0, diam2 = 0;
  double val[T::n_src][NSL];
  auto load_list = [&](uint32_t (&list)[NSL], uint32_t &mask, uint32_t &grow, double &gdiam, const size_t c) {
    const uint32_t *cl = A.cell_ls + c * N;
#pragma unroll
    for (int i = 0; i < NSL; ++i) list[i] = (lane + 32 * i < N) ? __ldg(cl + lane + 32 * i) : 0u;
    mask = A.hn ? __ldg(A.hn + c) : 0u;
    grow = general ? 0u : __ldg(A.geom_row + c);
    gdiam = (T::needs_diam && A.cell_diameters) ? __ldg(A.cell_diameters + c) : 0.0;
  };
  auto gather = [&](const uint32_t (&list)[NSL]) {
    if (T::has_src) {
#pragma unroll
      for (int f = 0; f < T::n_src; ++f)
        if (f < n_fields) {
#pragma unroll
          for (int i = 0; i < NSL; ++i) val[f][i] = A.src[f][list[i] & ADAFLO_B200_NODE_MASK];
        }
    }
  };
  load_list(nd, hm, row, diam, cell_l);
  if (cell_l + stride < A.n_cells) load_list(nd1, hm1, row1, diam1, cell_l + stride);
  if (T::tma_pd && lane == 0) tma_load_1d(pdbuf, A.pointdata + cell_l * (S3 * 24), kPdBytes, bar);
  if (!HnInfo(hm).any()) gather(nd);

  // sub-cell of this lane (loop invariant)
  unsigned it_cell = 0;
  for (; cell_l < A.n_cells; cell_l += stride, ++it_cell) {
    const uint32_t cell = (uint32_t)cell_l;
    const uint32_t *cl = A.cell_ls + (size_t)cell * N;
    const HnInfo hn(hm);
    // nodal read / scatter-add of cell slot lane + 32 i through the hanging-node stencil
    auto rd = [&](const double *vec, int i, bool res) -> double {
      const int slot = lane + 32 * i;
      if (!hn.any()) {
        const double v = vec[nd[i] & ADAFLO_B200_NODE_MASK];
        return (res && (nd[i] >> ADAFLO_B200_FLAG_SHIFT)) ? 0.0 : v;
      }
      LsStencil st;
      ls_stencil(hn, cl, S, slot % n1, (slot / n1) % n1, slot / (n1 * n1), st);
      double v = 0;
      for (int k = 0; k < st.n; ++k)
        if (!(res && (st.slot[k] >> ADAFLO_B200_FLAG_SHIFT))) v = fma(st.w[k], vec[st.slot[k] & ADAFLO_B200_NODE_MASK], v);
      return v;
    };
    auto add = [&](double *vec, int i, double v, bool skip_flagged) {
      const int slot = lane + 32 * i;
      if (!hn.any()) {
        if (!(skip_flagged && (nd[i] >> ADAFLO_B200_FLAG_SHIFT))) red_add(vec + (nd[i] & ADAFLO_B200_NODE_MASK), v);
        return;
      }
      LsStencil st;
      ls_stencil(hn, cl, S, slot % n1, (slot / n1) % n1, slot / (n1 * n1), st);
      for (int k = 0; k < st.n; ++k)
        if (!(skip_flagged && (st.slot[k] >> ADAFLO_B200_FLAG_SHIFT))) red_add(vec + (st.slot[k] & ADAFLO_B200_NODE_MASK), st.w[k] * v);
    };

    // stage the source values of this cell, then refill the pipeline
    if (T::has_src) {
#pragma unroll
      for (int f = 0; f < T::n_src; ++f)
        if (f < n_fields) {
#pragma unroll
          for (int i = 0; i < NSL; ++i)
            if (lane + 32 * i < N) {
              double v;
              if (hn.any())
                v = rd(A.src[f], i, T::resolve);
              else
                v = (T::resolve && (nd[i] >> ADAFLO_B200_FLAG_SHIFT)) ? 0.0 : val[f][i];
              cin[f * NP + sidx(lane + 32 * i)] = v;
            }
        }
    }
    const bool has1 = cell_l + stride < A.n_cells, has2 = cell_l + 2 * stride < A.n_cells;
    if (has2) load_list(nd2, hm2, row2, diam2, cell_l + 2 * stride);
    if (has1 && !HnInfo(hm1).any()) gather(nd1);
    // the other stage of the per-point buffer was read during the previous cell (warp-synchronised at its end)
    if (T::tma_pd && has1 && lane == 0)
      tma_load_1d(pdbuf + ((it_cell + 1) & 1) * (S3 * 24), A.pointdata + (cell_l + stride) * (S3 * 24), kPdBytes, bar + ((it_cell + 1) & 1));

    // geometry of the cell; the sub-cell is the cell scaled by 1 / n_sub
    double K[3][3] = {{0}};
    double det = 0;
    if (CART) {
      const double *tr = A.geom_table + 4 * (size_t)row;
      K[0][0] = __ldg(tr); K[1][1] = __ldg(tr + 1); K[2][2] = __ldg(tr + 2); det = __ldg(tr + 3);
    } else if (!general) {
      const double *tr = A.geom_table + 10 * (size_t)row;
#pragma unroll
      for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int e = 0; e < 3; ++e) K[f][e] = __ldg(tr + 3 * f + e);
      det = __ldg(tr + 9);
    }
    double jxw = det / (8.0 * S3);
    const double gs[3] = {dS * K[0][0], dS * K[1][1], dS * K[2][2]};   // Cartesian: gradient scale per direction ...
    const double fs[3] = {gs[0] * jxw, gs[1] * jxw, gs[2] * jxw};     // ... and flux scale including JxW
    const double h_rel = diam / dS;
    const double dd = fmax(A.eps_ratio, h_rel);
    const double delta = A.delta_scale * dd * dd;   // compute_normal.cc:105-108, compute_curvature.cc:112-117
    const double D = fmax(A.epsilon_used, h_rel);   // reinitialization.cc:77-79
    double *pdc = A.pointdata ? A.pointdata + (size_t)cell * S3 * 24 : nullptr;

    // auxiliary nodal fields
    if (T::has_aux) {
#pragma unroll
      for (int i = 0; i < NSL; ++i)
        if (lane + 32 * i < N) {
          double a0 = A.aux[0] ? rd(A.aux[0], i, false) : 0.0;
          double a1 = A.aux[1] ? rd(A.aux[1], i, false) : 0.0;
          double a2 = A.aux[2] ? rd(A.aux[2], i, false) : 0.0;
          if (OP == LS_CURV_RHS) {  // dof-level normalisation, compute_curvature.cc:231-240
            const double nn = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
            a0 = nn > 1e-2 ? a0 / nn : 0.0;
            a1 = nn > 1e-2 ? a1 / nn : 0.0;
            a2 = nn > 1e-2 ? a2 / nn : 0.0;
          }
          cin[(T::n_src + 0) * NP + sidx(lane + 32 * i)] = a0;
          cin[(T::n_src + 1) * NP + sidx(lane + 32 * i)] = a1;
          cin[(T::n_src + 2) * NP + sidx(lane + 32 * i)] = a2;
        }
    }
    if (T::has_u) {  // Q2 velocity of the cell (read_dof_values_plain of a distributed vector = the interpolated values)
      if (lane < 27) {
        const size_t un = __ldg(A.cell_u + (size_t)cell * 27 + lane) & ADAFLO_B200_NODE_MASK;
#pragma unroll
        for (int c = 0; c < 3; ++c) us[27 * c + lane] = A.vel[3 * un + c];
      }
      const HnInfo hnu(A.hn_u ? __ldg(A.hn_u + cell) : 0u);
      if (hnu.any()) {  // rare: lane 0 interpolates line by line, all lines of one direction before the next
        __syncwarp();
        if (lane == 0) {
          const int str[3] = {1, 3, 9};
          for (int d = 0; d < 3; ++d) {
            const int e = (d + 1) % 3, f = (d + 2) % 3;
            for (int ie = 0; ie < 3; ++ie)
              for (int jf = 0; jf < 3; ++jf)
                if (hnu.line(d, ie, jf, 3)) {
                  const int n0 = ie * str[e] + jf * str[f];
                  for (int c = 0; c < 3; ++c) hn_q2(hnu.sub(d), us[27 * c + n0], us[27 * c + n0 + str[d]], us[27 * c + n0 + 2 * str[d]]);
                }
          }
        }
      }
    }
    __syncwarp();
    if (T::tma_pd) mbar_wait(bar + (it_cell & 1), (it_cell >> 1) & 1);  // this cell's per-point stream has landed
    const double *pds = pdbuf + (it_cell & 1) * (S3 * 24);

    for (int f = 0; f < n_fields; ++f) {
      if (f > 0) __syncwarp();  // the corner buffers of the previous field have been summed
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int sub = lane + 32 * it;
        if (sub >= S3) continue;
        const int sx = sub % S, sy = (sub / S) % S, sz = sub / (S * S);
        const int nb = sx + n1 * sy + PS * sz;  // corner (0,0,0) of the sub-cell in the shared-memory node numbering
        double out[T::n_out][8];
        double cv[8], v[8], dx[4], dy[4], dz[4];
        if (T::has_src) {
          const double *cf = cin + (OP == LS_DAMPED ? f : 0) * NP;
#pragma unroll
          for (int n = 0; n < 8; ++n) cv[n] = cf[nb + (n & 1) + n1 * ((n >> 1) & 1) + PS * (n >> 2)];
        }
        if (OP == LS_DAMPED && CART) {  // constant coefficients: tensor product of 2x2 matrices, no quadrature loop
          const double vol = 8.0 * jxw;
          const double ck[3] = {vol * delta * gs[0] * gs[0] / 36.0, vol * delta * gs[1] * gs[1] / 36.0, vol * delta * gs[2] * gs[2] / 36.0};
          lsq::mass_laplace(cv, vol / 216.0, ck, out[0]);
        } else {
        if (T::has_src) lsq::eval(cv, v, dx, dy, dz);
        double a0[8], a1[8], a2[8];  // auxiliary fields at the corners
        if (T::has_aux) {
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const int nn = nb + (n & 1) + n1 * ((n >> 1) & 1) + PS * (n >> 2);
            a0[n] = cin[(T::n_src + 0) * NP + nn];
            a1[n] = cin[(T::n_src + 1) * NP + nn];
            a2[n] = cin[(T::n_src + 2) * NP + nn];
          }
        }
        double uq[8][3];
        if (T::has_u) {  // Q2 velocity at the 8 points: tensor-product Lagrange basis on {0, 1/2, 1}
          double L[3][2][3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const double x = ((d == 0 ? sx : (d == 1 ? sy : sz)) + (p ? lsq::g1 : lsq::g0)) / dS;
              L[d][p][0] = 2 * (x - 0.5) * (x - 1);
              L[d][p][1] = 4 * x * (1 - x);
              L[d][p][2] = 2 * x * (x - 0.5);
            }
#pragma unroll
          for (int p = 0; p < 8; ++p) uq[p][0] = uq[p][1] = uq[p][2] = 0;
#pragma unroll 3
          for (int n = 0; n < 27; ++n) {
            const double u0 = us[n], u1 = us[27 + n], u2 = us[54 + n];
            const int i = n % 3, j = (n / 3) % 3, k = n / 9;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const double w = L[0][p & 1][i] * L[1][(p >> 1) & 1][j] * L[2][p >> 2][k];
              uq[p][0] = fma(w, u0, uq[p][0]);
              uq[p][1] = fma(w, u1, uq[p][1]);
              uq[p][2] = fma(w, u2, uq[p][2]);
            }
          }
        }
        double tv[T::n_out][8], tx[8], ty[8], tz[8];
        constexpr bool kGrad = (OP == LS_DAMPED || OP == LS_REINIT || OP == LS_REINIT_DIFFUSE || OP == LS_REINIT_RHS ||
                                OP == LS_REINIT_RHS_DIFFUSE);
        constexpr bool kVal = (OP != LS_REINIT_RHS && OP != LS_REINIT_RHS_DIFFUSE);
        double vold[8], voo[8];  // old level sets at the points (advection right-hand side)
        if (OP == LS_ADVECT_RHS) {
          double d0[4], d1[4], d2[4];
          lsq::eval(a0, vold, d0, d1, d2);
          lsq::eval(a1, voo, d0, d1, d2);
        }
        double *pd = pdc ? pdc + sub : nullptr;
        const double *pq = pds + sub;  // staged copy (advection / reinitialisation vmult)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int px = p & 1, py = (p >> 1) & 1, pz = p >> 2;
          if (!CART && general) {  // MappingQ cells: inverse Jacobian and JxW of this point
            constexpr int nq1 = 2 * S;
            constexpr size_t NQ = (size_t)nq1 * nq1 * nq1;
            const size_t q = (2 * sx + px) + (size_t)nq1 * ((2 * sy + py) + (size_t)nq1 * (2 * sz + pz));
#pragma unroll
            for (int ff = 0; ff < 3; ++ff)
#pragma unroll
              for (int e = 0; e < 3; ++e) K[ff][e] = __ldg(A.inv_jac + ((size_t)cell * 9 + 3 * ff + e) * NQ + q);
            jxw = __ldg(A.jxw_q + (size_t)cell * NQ + q);
          }
          // real-space gradient of a field from its sub-cell reference gradient r (slope of the hats: n_sub)
          auto to_real = [&](const double r0, const double r1, const double r2, double (&g)[3]) {
            if (CART) {
              g[0] = gs[0] * r0; g[1] = gs[1] * r1; g[2] = gs[2] * r2;
            } else {
#pragma unroll
              for (int e = 0; e < 3; ++e) g[e] = dS * (K[0][e] * r0 + K[1][e] * r1 + K[2][e] * r2);
            }
          };
          double g[3] = {0, 0, 0};
          const double vp = T::has_src ? v[p] : 0.0;
          if (T::has_src) to_real(dx[py + 2 * pz], dy[px + 2 * pz], dz[px + 2 * py], g);
          double t = 0, tg[3] = {0, 0, 0}, t3[3] = {0, 0, 0};
          // hats of the sub-cell at this point (for interpolating auxiliary corner values)
          const double wx[2] = {px ? lsq::g0 : lsq::g1, px ? lsq::g1 : lsq::g0}, wy[2] = {py ? lsq::g0 : lsq::g1, py ? lsq::g1 : lsq::g0},
                       wz[2] = {pz ? lsq::g0 : lsq::g1, pz ? lsq::g1 : lsq::g0};
          auto interp = [&](const double (&a)[8]) {
            double r = 0;
#pragma unroll
            for (int n = 0; n < 8; ++n) r = fma(wx[n & 1] * wy[(n >> 1) & 1] * wz[n >> 2], a[n], r);
            return r;
          };
          if (OP == LS_ADVECT) {
            t = A.gamma * vp + g[0] * pq[(0 + p) * S3] + g[1] * pq[(8 + p) * S3] + g[2] * pq[(16 + p) * S3];
          } else if (OP == LS_DAMPED) {
            t = vp;
#pragma unroll
            for (int e = 0; e < 3; ++e) tg[e] = delta * g[e];
          } else if (OP == LS_REINIT) {
            const double nq[3] = {pq[(0 + p) * S3], pq[(8 + p) * S3], pq[(16 + p) * S3]};
            const double ng = nq[0] * g[0] + nq[1] * g[1] + nq[2] * g[2];
            t = A.dtau_inv * vp;
#pragma unroll
            for (int e = 0; e < 3; ++e) tg[e] = D * ng * nq[e];
          } else if (OP == LS_REINIT_DIFFUSE) {
            t = A.dtau_inv * vp;
#pragma unroll
            for (int e = 0; e < 3; ++e) tg[e] = D * g[e];
          } else if (OP == LS_NORMAL_RHS) {
#pragma unroll
            for (int e = 0; e < 3; ++e) t3[e] = g[e];
          } else if (OP == LS_CURV_RHS) {
            // divergence of the normalised nodal normal field: gradient of each component from its corner values
            double div = 0;
            const double(*aa[3])[8] = {&a0, &a1, &a2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const double(&a)[8] = *aa[c];
              double r0 = 0, r1 = 0, r2 = 0;  // reference gradient of the trilinear interpolant of a at this point
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const int ia = n & 1, ib = (n >> 1) & 1, ic = n >> 2;
                r0 = fma((ia ? 1.0 : -1.0) * wy[ib] * wz[ic], a[n], r0);
                r1 = fma(wx[ia] * (ib ? 1.0 : -1.0) * wz[ic], a[n], r1);
                r2 = fma(wx[ia] * wy[ib] * (ic ? 1.0 : -1.0), a[n], r2);
              }
              double ga[3];
              to_real(r0, r1, r2, ga);
              div += ga[c];
            }
            t = -div;
          } else if (OP == LS_REINIT_RHS) {
            double nq[3];
            if (A.first_step) {  // freeze the normal of the first pseudo time step, reinitialization.cc:167-172
              const double nv[3] = {interp(a0), interp(a1), interp(a2)};
              const double nn = fmax(1e-4, sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]));
#pragma unroll
              for (int e = 0; e < 3; ++e) {
                nq[e] = nv[e] / nn;
                pd[(8 * e + p) * S3] = nq[e];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 3; ++e) nq[e] = pd[(8 * e + p) * S3];
            }
            const double ng = nq[0] * g[0] + nq[1] * g[1] + nq[2] * g[2];
            const double fl = 0.5 * (1. - vp * vp) - ng * D;
#pragma unroll
            for (int e = 0; e < 3; ++e) tg[e] = nq[e] * fl;
          } else if (OP == LS_REINIT_RHS_DIFFUSE) {
#pragma unroll
            for (int e = 0; e < 3; ++e) tg[e] = -D * g[e];
          } else if (OP == LS_ADVECT_RHS) {
            // BDF sum of the three level sets: the old values are evaluated with the same arithmetic as the new one and
            // the products are rounded individually (no fma contraction), so that equal fields cancel exactly like in the
            // reference -- the log's "advect [0/0]" of a step with zero velocity is an exact zero
            const double bdf = __dadd_rn(__dadd_rn(__dmul_rn(A.gamma, vp), __dmul_rn(A.w_old, vold[p])), __dmul_rn(A.w_oo, voo[p]));
            t = -(bdf + (uq[p][0] * g[0] + uq[p][1] * g[1] + uq[p][2] * g[2]));
#pragma unroll
            for (int e = 0; e < 3; ++e) pd[(8 * e + p) * S3] = uq[p][e];  // evaluated_convection, :389
          }
          if (OP == LS_NORMAL_RHS) {
#pragma unroll
            for (int e = 0; e < 3; ++e) tv[e][p] = t3[e] * jxw;
          } else
            tv[0][p] = t * jxw;
          if (kGrad) {  // real-space flux -> sub-cell reference flux
            if (CART) {
              tx[p] = fs[0] * tg[0]; ty[p] = fs[1] * tg[1]; tz[p] = fs[2] * tg[2];
            } else {
              tx[p] = dS * (K[0][0] * tg[0] + K[0][1] * tg[1] + K[0][2] * tg[2]) * jxw;
              ty[p] = dS * (K[1][0] * tg[0] + K[1][1] * tg[1] + K[1][2] * tg[2]) * jxw;
              tz[p] = dS * (K[2][0] * tg[0] + K[2][1] * tg[1] + K[2][2] * tg[2]) * jxw;
            }
          }
          if (OP == LS_MASS_DIAG) {  // diagonal of the mass matrix: sum_q psi_n^2 JxW
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const double psi = wx[n & 1] * wy[(n >> 1) & 1] * wz[n >> 2];
              out[0][n] = (p == 0 ? 0.0 : out[0][n]) + psi * psi * jxw;
            }
          }
        }
        if (OP != LS_MASS_DIAG) {
#pragma unroll
          for (int o = 0; o < T::n_out; ++o) lsq::integrate<kVal, kGrad>(tv[o], tx, ty, tz, out[o]);
        }
        }
        // publish the sub-cell vector: one buffer per corner, no two sub-cells share an entry
#pragma unroll
        for (int o = 0; o < T::n_out; ++o)
#pragma unroll
          for (int n = 0; n < 8; ++n) cbuf[(o * 8 + n) * NP + nb + (n & 1) + n1 * ((n >> 1) & 1) + PS * (n >> 2)] = out[o][n];
      }
      __syncwarp();
      // cell vector: sum of the 8 corner buffers; scatter it
#pragma unroll
      for (int i = 0; i < NSL; ++i) {
        const int slot = lane + 32 * i;
        if (slot < N) {
          double sum[T::n_out];
#pragma unroll
          for (int o = 0; o < T::n_out; ++o) {
            sum[o] = cbuf[(o * 8) * NP + sidx(slot)];
#pragma unroll
            for (int n = 1; n < 8; ++n) sum[o] += cbuf[(o * 8 + n) * NP + sidx(slot)];
          }
          constexpr bool skip = OP != LS_MASS_DIAG;
          if (OP == LS_NORMAL_RHS) {
#pragma unroll
            for (int o = 0; o < T::n_out; ++o) add(A.dst[o], i, sum[o], skip);
          } else
            add(A.dst[f], i, sum[0], skip);
        }
      }
    }
    // advance the pipeline
#pragma unroll
    for (int i = 0; i < NSL; ++i) {
      nd[i] = nd1[i];
      nd1[i] = nd2[i];
    }
    hm = hm1; hm1 = hm2;
    row = row1; row1 = row2;
    diam = diam1; diam1 = diam2;
    __syncwarp();  // the shared-memory vectors of this cell have been consumed
  }
}

template <int OP, int S, bool CART>
int launch_ls_cell_sc(const LSArgs &a, cudaStream_t s)
{
  auto kern = ls_cell_kernel<OP, S, CART>;
  constexpr int warps = ls_cell_warps<OP, S>();
  const size_t smem = sizeof(double) * warps * ls_cell_warp_doubles<OP, S>();
  static int blocks_per_sm[64] = {};
  static int n_sm[64] = {};
  int dev = 0;
  AB_CUDA(cudaGetDevice(&dev));
  if (dev >= 64) return fail(ADAFLO_B200_ERR_ARG, "launch_ls_cell: device ordinal out of range");
  if (!blocks_per_sm[dev]) {
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int b = 0;
    AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, warps * 32, smem));
    AB_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    blocks_per_sm[dev] = b > 0 ? b : 1;
  }
  if (LsCellTraits<OP>::tma_pd && !a.pointdata) return fail(ADAFLO_B200_ERR_STATE, "launch_ls_cell: per-point data missing");
  const size_t needed = ((size_t)a.n_cells + warps - 1) / warps;
  const size_t resident = (size_t)n_sm[dev] * blocks_per_sm[dev];
  kern<<<(unsigned)(needed < resident ? needed : resident), warps * 32, smem, s>>>(a);
  return 0;
}
template <int OP>
int launch_ls_cell(const LSArgs &a, cudaStream_t s)
{
  const bool cart = a.geom == ADAFLO_B200_GEOM_CARTESIAN;
  if (a.n_sub == 3) return cart ? launch_ls_cell_sc<OP, 3, true>(a, s) : launch_ls_cell_sc<OP, 3, false>(a, s);
  return cart ? launch_ls_cell_sc<OP, 4, true>(a, s) : launch_ls_cell_sc<OP, 4, false>(a, s);
}

// ---------------------------------------------------------------------------------------------------------------
// Projection operator (mass + delta Laplace: compute_normal_vmult / compute_curvature_vmult, LS_DAMPED) on Cartesian cells
// without hanging nodes, applied at CELL level: with constant coefficients the cell matrix of FE_Q_iso_Q1(S) is
//   alpha [ M (x) M (x) M + kx K (x) M (x) M + ky M (x) K (x) M + kz M (x) M (x) K ],      alpha = |sub-cell| / 216, k_d = 6 delta / h_d^2
// with the assembled 1-D matrices of the S sub-intervals, M = tridiag(1, [2, 4, .., 4, 2], 1), K = tridiag(-1, [1, 2, .., 2, 1], -1)
// (what the 2-point Gauss rule integrates exactly on every sub-interval, compute_normal.cc:82-119).  Sum factorisation on the
// (S + 1)^3 nodal array: one lane per grid line, three register sweeps
//   x: a = M u, b = kx K u          y: c = M a, e = c + M b + ky K a          z: out = alpha (M e + kz K c)
// with two transposes through shared memory (for S = 3: two cells per warp, a 4 x 4 x 4 array per cell stored bank-swizzled so
// that every phase is conflict free).  Per cell and field: 2 (S + 1)^3 shared-memory stores and loads and about 70 flops per line,
// instead of 8 loads + 8 stores per sub-cell plus the 8-fold corner sum per node, and 98 flops per sub-cell.
template <int S>
struct LsLineCfg {
  static constexpr int n1 = S + 1, lines = n1 * n1, N = n1 * n1 * n1;
  static constexpr int cpw = 32 / lines;          // cells per warp
  static constexpr int group = S == 3 ? 16 : 32;  // lanes per cell
  static constexpr int PN = S == 3 ? 64 : (n1 * n1 + 1) * n1;  // doubles per array
  __device__ static __forceinline__ int addr(int ix, int iy, int iz)
  {
    if (S == 3) return ((ix + iz) & 3) + 4 * ((iy + iz) & 3) + 16 * iz;  // every plane of the 4^3 array covers the 16 8-byte banks once
    return ix + n1 * iy + (n1 * n1 + 1) * iz;
  }
};
template <int n1>
__device__ __forceinline__ void tri_m(const double (&x)[n1], double (&y)[n1])
{
  y[0] = 2.0 * x[0] + x[1];
#pragma unroll
  for (int i = 1; i < n1 - 1; ++i) y[i] = fma(4.0, x[i], x[i - 1] + x[i + 1]);
  y[n1 - 1] = 2.0 * x[n1 - 1] + x[n1 - 2];
}
template <int n1>
__device__ __forceinline__ void tri_k(const double (&x)[n1], double (&y)[n1])
{
  y[0] = x[0] - x[1];
#pragma unroll
  for (int i = 1; i < n1 - 1; ++i) y[i] = fma(2.0, x[i], -(x[i - 1] + x[i + 1]));
  y[n1 - 1] = x[n1 - 1] - x[n1 - 2];
}

// 4 x 4 transpose of r[0..3] across the four lanes {base, base + d, base + 2 d, base + 3 d} of a warp (d = 1 or 4; t = the lane's
// index in that group): afterwards r[j] holds what lane j of the group had in r[t].  Two exchange stages, two 64-bit shuffles each:
// half the L1 data-pipe wavefronts of the same transpose through shared memory (4 STS.64 + 4 LDS.64).
__device__ __forceinline__ void transpose4(double (&r)[4], const int t, const int d)
{
  {
    const bool up = t & 2;
    const double x0 = up ? r[0] : r[2], x1 = up ? r[1] : r[3];
    const double y0 = __shfl_xor_sync(0xffffffffu, x0, 2 * d), y1 = __shfl_xor_sync(0xffffffffu, x1, 2 * d);
    if (up) { r[0] = y0; r[1] = y1; } else { r[2] = y0; r[3] = y1; }
  }
  {
    const bool up = t & 1;
    const double x0 = up ? r[0] : r[1], x1 = up ? r[2] : r[3];
    const double y0 = __shfl_xor_sync(0xffffffffu, x0, d), y1 = __shfl_xor_sync(0xffffffffu, x1, d);
    if (up) { r[0] = y0; r[2] = y1; } else { r[1] = y0; r[3] = y1; }
  }
}

template <int S>
__global__ void __launch_bounds__(256, 4) ls_damped_lines_kernel(const LSArgs A)
{
  using C = LsLineCfg<S>;
  constexpr int n1 = C::n1, WARPS = 8;
  __shared__ double sm[WARPS][C::cpw][2][C::PN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / C::group, l = lane % C::group;  // cell of the warp's bundle, line within the cell
  const bool line_ok = sub < C::cpw && l < C::lines;
  const int p = l % n1, q = l / n1;
  const size_t n_bundles = ((size_t)A.n_cells + C::cpw - 1) / C::cpw;
  for (size_t bundle = (size_t)blockIdx.x * WARPS + warp; bundle < n_bundles; bundle += (size_t)gridDim.x * WARPS) {
    const size_t cell = bundle * C::cpw + sub;
    const bool ok = line_ok && cell < A.n_cells;
    double *sa = sm[warp][ok ? sub : 0][0], *sb = sm[warp][ok ? sub : 0][1];
    uint32_t nx[n1], nz[n1];  // node entries of the x line (iy, iz) = (p, q) and of the z line (ix, iy) = (p, q)
    double alpha = 0, kx = 0, ky = 0, kz = 0;
    if (ok) {
      const uint32_t *cl = A.cell_ls + cell * C::N;
      if (S == 3) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(cl) + l);
        nx[0] = v.x; nx[1] = v.y; nx[2] = v.z; nx[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < n1; ++i) nx[i] = __ldg(cl + i + n1 * l);
      }
#pragma unroll
      for (int i = 0; i < n1; ++i) nz[i] = __ldg(cl + l + n1 * n1 * i);
      const double *tr = A.geom_table + 4 * (size_t)__ldg(A.geom_row + cell);
      const double gx = S * __ldg(tr), gy = S * __ldg(tr + 1), gz = S * __ldg(tr + 2), vol = __ldg(tr + 3) / (S * S * S);
      const double h_rel = (A.cell_diameters ? __ldg(A.cell_diameters + cell) : 0.0) / S;
      const double dd = fmax(A.eps_ratio, h_rel);
      const double delta = A.delta_scale * dd * dd;  // compute_normal.cc:105-108, compute_curvature.cc:112-117
      alpha = vol / 216.0;
      kx = 6.0 * delta * gx * gx;
      ky = 6.0 * delta * gy * gy;
      kz = 6.0 * delta * gz * gz;
    }
    if constexpr (S == 3) {  // transposes by warp shuffles: no shared memory, no barriers (idle lanes take part with zeros)
      for (int f = 0; f < A.n_comp; ++f) {
        double u[4] = {0, 0, 0, 0}, a[4], b[4], c[4], e[4], t1[4], t2[4];
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) u[i] = (nx[i] >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : A.src[f][nx[i] & ADAFLO_B200_NODE_MASK];
        }
        tri_m<4>(u, a);
        tri_k<4>(u, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] *= kx;
        transpose4(a, p, 1);  // lane (iy, iz) = (p, q), index ix  ->  lane (ix, iz) = (p, q), index iy
        transpose4(b, p, 1);
        tri_m<4>(a, c);
        tri_m<4>(b, t1);
        tri_k<4>(a, t2);
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = c[i] + fma(ky, t2[i], t1[i]);
        transpose4(c, q, 4);  // lane (ix, iz) = (p, q), index iy  ->  lane (ix, iy) = (p, q), index iz
        transpose4(e, q, 4);
        tri_m<4>(e, t1);
        tri_k<4>(c, t2);
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (!(nz[i] >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst[f] + (nz[i] & ADAFLO_B200_NODE_MASK), alpha * fma(kz, t2[i], t1[i]));
        }
      }
    } else {
    for (int f = 0; f < A.n_comp; ++f) {
      if (ok) {
        double u[n1], a[n1], b[n1];
#pragma unroll
        for (int i = 0; i < n1; ++i) u[i] = (nx[i] >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : A.src[f][nx[i] & ADAFLO_B200_NODE_MASK];
        tri_m<n1>(u, a);
        tri_k<n1>(u, b);
#pragma unroll
        for (int i = 0; i < n1; ++i) {
          sa[C::addr(i, p, q)] = a[i];
          sb[C::addr(i, p, q)] = kx * b[i];
        }
      }
      __syncwarp();
      if (ok) {  // y lines: lane (ix, iz) = (p, q); the lane rewrites exactly the entries it has read
        double a[n1], b[n1], c[n1], mb[n1], ka[n1];
#pragma unroll
        for (int i = 0; i < n1; ++i) {
          a[i] = sa[C::addr(p, i, q)];
          b[i] = sb[C::addr(p, i, q)];
        }
        tri_m<n1>(a, c);
        tri_m<n1>(b, mb);
        tri_k<n1>(a, ka);
#pragma unroll
        for (int i = 0; i < n1; ++i) {
          sa[C::addr(p, i, q)] = c[i];
          sb[C::addr(p, i, q)] = c[i] + fma(ky, ka[i], mb[i]);
        }
      }
      __syncwarp();
      if (ok) {  // z lines: lane (ix, iy) = (p, q)
        double c[n1], e[n1], me[n1], kc[n1];
#pragma unroll
        for (int i = 0; i < n1; ++i) {
          c[i] = sa[C::addr(p, q, i)];
          e[i] = sb[C::addr(p, q, i)];
        }
        tri_m<n1>(e, me);
        tri_k<n1>(c, kc);
#pragma unroll
        for (int i = 0; i < n1; ++i)
          if (!(nz[i] >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst[f] + (nz[i] & ADAFLO_B200_NODE_MASK), alpha * fma(kz, kc[i], me[i]));
      }
      __syncwarp();
    }
    }  // S != 3
  }
}

template <int S>
int launch_ls_damped_lines(const LSArgs &a, cudaStream_t s)
{
  constexpr int cpw = LsLineCfg<S>::cpw;
  int dev = 0, n_sm = 0;
  AB_CUDA(cudaGetDevice(&dev));
  AB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const size_t needed = (((size_t)a.n_cells + cpw - 1) / cpw + 7) / 8;
  const size_t resident = (size_t)n_sm * 8;  // up to 8 blocks of 8 warps per SM
  ls_damped_lines_kernel<S><<<(unsigned)(needed < resident ? needed : resident), 256, 0, s>>>(a);
  return 0;
}

static bool ls_use_subcell_kernel()
{
  static const bool v1 = [] {
    const char *e = getenv("ADAFLO_B200_LS_KERNEL");
    return e && e[0] == 'v' && e[1] == '1';
  }();
  return v1;
}

int launch_ls(const LSArgs &a, int op, cudaStream_t s)
{
  const size_t threads = (size_t)a.n_cells * a.n_sub * a.n_sub * a.n_sub;
  if (threads == 0) return 0;
  const unsigned grid = (unsigned)((threads + 127) / 128);
  const bool cell_kernel = a.n_sub >= 3 && !ls_use_subcell_kernel();  // n_sub <= 2: too few sub-cells for a warp per cell
  static const bool no_lines = [] { const char *e = getenv("ADAFLO_B200_LS_DAMPED_LINES"); return e && e[0] == '0'; }();  // A/B switch
  if (op == LS_DAMPED && cell_kernel && !no_lines && a.geom == ADAFLO_B200_GEOM_CARTESIAN && !a.hn && (a.n_sub == 3 || a.n_sub == 4)) {
    if (a.n_sub == 3)
      AB_CHECK(launch_ls_damped_lines<3>(a, s));
    else
      AB_CHECK(launch_ls_damped_lines<4>(a, s));
    AB_COUNT_LAUNCH();
    AB_CUDA(cudaGetLastError());
    return 0;
  }
#define AB_L(O) case O: if (cell_kernel) { AB_CHECK(launch_ls_cell<O>(a, s)); } else ls_subcell_kernel<O><<<grid, 128, 0, s>>>(a); break;
  switch (op) {
    AB_L(LS_ADVECT) AB_L(LS_DAMPED) AB_L(LS_REINIT) AB_L(LS_REINIT_DIFFUSE) AB_L(LS_NORMAL_RHS) AB_L(LS_CURV_RHS)
    AB_L(LS_REINIT_RHS) AB_L(LS_REINIT_RHS_DIFFUSE) AB_L(LS_ADVECT_RHS) AB_L(LS_MASS_DIAG)
    default: return fail(ADAFLO_B200_ERR_ARG, "launch_ls: bad operator");
  }
#undef AB_L
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// pointwise kernels
__global__ void k_curvature_correction(double *__restrict__ curv, const double *__restrict__ ls, double eps_used, int dim, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (curv[i] > 1e-4) {  // compute_curvature.cc:364-376
      const double c = ls[i];
      const double distance = (1 - c * c) > 1e-2 ? eps_used * log((1. + c) / (1. - c)) : 0;
      curv[i] = 1. / (1. / curv[i] + distance / (dim - 1));
    }
}
__global__ void k_scaled_copy_indexed(double *__restrict__ dst, const double *__restrict__ src, const double *__restrict__ diag,
                                      const uint32_t *__restrict__ idx, uint32_t n)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t j = idx[i];
    dst[j] = diag[j] * src[j];  // constrained rows of every level-set vmult, e.g. reinitialization.cc:227-230
  }
}
int launch_curvature_correction(double *curv, const double *ls, double eps_used, int dim, uint64_t n, cudaStream_t s)
{
  if (!n) return 0;
  k_curvature_correction<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(curv, ls, eps_used, dim, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}
int launch_scaled_copy_indexed(double *dst, const double *src, const double *diag, const uint32_t *idx, uint32_t n, cudaStream_t s)
{
  if (!n) return 0;
  k_scaled_copy_indexed<<<(n + 255) / 256, 256, 0, s>>>(dst, src, diag, idx, n);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
