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
  double *pd = A.pointdata ? A.pointdata + ((size_t)cell * S3 + sub) * 24 : nullptr;
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
        tv = A.gamma * v + g[0] * pd[0 + p] + g[1] * pd[8 + p] + g[2] * pd[16 + p];
      } else if (OP == LS_DAMPED) {
        tv = v;
#pragma unroll
        for (int e = 0; e < 3; ++e) tg[e] = delta * g[e];
      } else if (OP == LS_REINIT) {
        const double nq[3] = {pd[0 + p], pd[8 + p], pd[16 + p]};
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
            pd[8 * e + p] = nq[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 3; ++e) nq[e] = pd[8 * e + p];
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
        for (int e = 0; e < 3; ++e) pd[8 * e + p] = uq[p][e];  // evaluated_convection, :389
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

int launch_ls(const LSArgs &a, int op, cudaStream_t s)
{
  const size_t threads = (size_t)a.n_cells * a.n_sub * a.n_sub * a.n_sub;
  if (threads == 0) return 0;
  const unsigned grid = (unsigned)((threads + 127) / 128);
#define AB_L(O) case O: ls_subcell_kernel<O><<<grid, 128, 0, s>>>(a); break;
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
