// ns_residual_r1.cu -- nonlinear residual (LocalOps::residual, source/navier_stokes_matrix.cc:266-293, q-point block
// :725-800) on the thread mapping of the vmult kernel (ns_r1.cuh): 3 lanes per cell, a warp owns a decade of cells,
// persistent blocks.  Round 1's residual kernel (ns_residual.cu, 9 lanes per cell) wrote the
// linearisation point with scattered 16-byte stores (0.2 of the HBM peak); here every store of a warp covers 480
// contiguous bytes of the decade-blocked, pair-interleaved array the vmult kernel reads (lin_index in common.cuh).
//
// Differences from the vmult cell loop:
//  * read_dof_values_plain (:663-664, 692-693): Dirichlet flags are ignored on the gather (hanging nodes are still
//    interpolated -- the hanging dofs themselves are not stored), and respected on the scatter;
//  * the BDF history enters linearly: w_old u_old + w_oldold u_oldold is combined at dof level and interpolated once
//    (values only) through a second, smaller tile;
//  * the physics is the residual's (:725-800, Newton / Picard branch) and the linearisation point is WRITTEN.
// The extrapolating linearisations (semi-implicit, explicit, projection) need a third field with gradients; they are
// not on the hot path and stay on the round-1 kernel (launch_ns_residual dispatches).
#include "ns_r1.cuh"
#ifndef AB_RES_ROLLJ  // line loop of phase B rolled (2280 instead of 3152 instructions: the unrolled kernel loses 6 % to instruction fetch)
#define AB_RES_ROLLJ 0
#endif

namespace ab200 {

constexpr int kResWarps = 8;            // workers per block: 8 x 28.4 KB of shared memory
constexpr int kResHist = 243;           // offset of the history tile: entry (c, qx, qy, k) at 243 + k + 3 qx + 9 c + 33 qy
constexpr int kResHistQy = 33;          // = 1 (mod 16): lane (cell, t) on bank 3 cell + t in both ownership patterns
constexpr int kResP = kResHist + 3 * kResHistQy;  // 342: 12 pressure slots
constexpr int kResCellSmem = 355;       // = 3 (mod 16)
constexpr int kResWarpSmem = kGroup * kResCellSmem;

__device__ __forceinline__ void st_pair(double *p, double a, double b)
{
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

template <int GEOM, bool VAR, bool HANGING>
__device__ __forceinline__ void r1_residual_decade(const NSArgs &A, const size_t g, double *warp_smem, const int lane)
{
  AB_TABLES_Q2G3
  const int cw = lane / 3;
  const int t = lane - 3 * cw;
  const size_t cell_l = g * kGroup + cw;
  const bool active = (lane < 3 * kGroup) && (cell_l < A.n_cells);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t cell = (uint32_t)cell_l;
  double *s = warp_smem + cw * kResCellSmem;
  double *sh = s + kResHist;
  double *sp = s + kResP;

  const bool stokes = A.physical_type == ADAFLO_B200_STOKES;
  const bool use_old = A.physical_type == ADAFLO_B200_INCOMPRESSIBLE;
  const bool newton = A.linearization == ADAFLO_B200_COUPLED_IMPLICIT_NEWTON;

  Geometry<GEOM> G;
  G.init_blk(A, cell);
  const HnInfo hn(hn_to_kernel_axes((HANGING && A.hn_u) ? __ldg(A.hn_u + cell) : 0u));
  const HnInfo hnp(hn_to_kernel_axes((HANGING && A.hn_p) ? __ldg(A.hn_p + cell) : 0u));
  const bool hn_z_warp = HANGING && __any_sync(mask, hn.any() && (hn.face(0) || hn.face(1) || hn.edge(2)));
  const bool hnp_z_warp = HANGING && __any_sync(mask, hnp.any() && (hnp.face(0) || hnp.face(1) || hnp.edge(2)));

  // ---------------------------------------------------------------- phase A
  uint32_t node[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) node[j] = __ldg(A.cell_u_blk + (g * 9 + j) * 30 + lane);
  uint32_t pnode[4] = {0, 0, 0, 0};
  double p_dg0 = 0;
  uint32_t dg0_node = 0;
  if (t < 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) pnode[j] = __ldg(A.cell_p_blk + (g * 4 + j) * 30 + lane);
  }
  if (A.np_cell == 9) {
    dg0_node = __ldg(A.cell_p + (size_t)cell * 9 + 8);
    p_dg0 = A.src_p[dg0_node & ADAFLO_B200_NODE_MASK];
  }
  // read_dof_values for hanging nodes: parent values interpolated onto this cell's nodes (see ns_r1.cuh)
  auto resolve = [&](double(&P)[3][9]) {
    if (!HANGING) return;
    if (hn.any()) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (hn.line(0, i, t, 3)) {
#pragma unroll
          for (int c = 0; c < 3; ++c) hn_q2(hn.sub(0), P[c][3 * i], P[c][3 * i + 1], P[c][3 * i + 2]);
        }
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (hn.line(1, t, i, 3)) {
#pragma unroll
          for (int c = 0; c < 3; ++c) hn_q2(hn.sub(1), P[c][i], P[c][3 + i], P[c][6 + i]);
        }
    }
    if (hn_z_warp) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const bool need = hn.any() && hn.line(2, j % 3, j / 3, 3);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double v0 = __shfl_sync(mask, P[c][j], 3 * cw), v1 = __shfl_sync(mask, P[c][j], 3 * cw + 1),
                 v2 = __shfl_sync(mask, P[c][j], 3 * cw + 2);
          hn_q2(hn.sub(2), v0, v1, v2);
          if (need) P[c][j] = t == 0 ? v0 : (t == 1 ? v1 : v2);
        }
      }
    }
  };
  {
    double P[3][9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const size_t n = node[j] & ADAFLO_B200_NODE_MASK;
#pragma unroll
      for (int c = 0; c < 3; ++c) P[c][j] = A.src_u[3 * n + c];  // plain: constrained entries are read as stored
    }
    resolve(P);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double Y[3][3], Yd[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        q2::val(P[c][i], P[c][3 + i], P[c][6 + i], Y[0][i], Y[1][i], Y[2][i]);
        q2::der(P[c][i], P[c][3 + i], P[c][6 + i], Yd[0][i], Yd[1][i], Yd[2][i]);
      }
      double *sc = s + c * kR1CompSmem;
#pragma unroll
      for (int qy = 0; qy < 3; ++qy) {
        double tv[3], tx[3], ty[3];
        q2::val(Y[qy][0], Y[qy][1], Y[qy][2], tv[0], tv[1], tv[2]);
        q2::der(Y[qy][0], Y[qy][1], Y[qy][2], tx[0], tx[1], tx[2]);
        q2::val(Yd[qy][0], Yd[qy][1], Yd[qy][2], ty[0], ty[1], ty[2]);
#pragma unroll
        for (int qx = 0; qx < 3; ++qx) {
          sc[0 + r1_idx(qx, qy, t)] = tv[qx];
          sc[9 + r1_idx(qx, qy, t)] = tx[qx];
          sc[18 + r1_idx(qx, qy, t)] = ty[qx];
        }
      }
    }
    if (use_old && !stokes) {  // BDF history, combined at dof level: w_old u_old + w_oldold u_oldold (:727-732)
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const size_t n = node[j] & ADAFLO_B200_NODE_MASK;
#pragma unroll
        for (int c = 0; c < 3; ++c) P[c][j] = A.w_old * A.old_u[3 * n + c] + A.w_oo * A.old_old_u[3 * n + c];
      }
      resolve(P);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double Y[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) q2::val(P[c][i], P[c][3 + i], P[c][6 + i], Y[0][i], Y[1][i], Y[2][i]);
#pragma unroll
        for (int qy = 0; qy < 3; ++qy) {
          double tv[3];
          q2::val(Y[qy][0], Y[qy][1], Y[qy][2], tv[0], tv[1], tv[2]);
#pragma unroll
          for (int qx = 0; qx < 3; ++qx) sh[t + 3 * qx + 9 * c + kResHistQy * qy] = tv[qx];
        }
      }
    }
    {
      double pl[4] = {0, 0, 0, 0};
      if (t < 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pl[j] = A.src_p[pnode[j] & ADAFLO_B200_NODE_MASK];  // plain
        if (HANGING && hnp.any()) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (hnp.line(0, i, t, 2)) hn_q1(hnp.sub(0), pl[2 * i], pl[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (hnp.line(1, t, i, 2)) hn_q1(hnp.sub(1), pl[i], pl[2 + i]);
        }
      }
      if (hnp_z_warp) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double v0 = __shfl_sync(mask, pl[j], 3 * cw), v1 = __shfl_sync(mask, pl[j], 3 * cw + 1);
          hn_q1(hnp.sub(2), v0, v1);
          if (hnp.any() && hnp.line(2, j % 2, j / 2, 2) && t < 2) pl[j] = t == 0 ? v0 : v1;
        }
      }
      if (t < 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[t + 2 * j] = pl[j];
      }
    }
  }
  __syncwarp(mask);
  double py[2][2], pacc[2][2] = {{0, 0}, {0, 0}};
  const double npt0 = t == 0 ? NP[0][0] : (t == 1 ? 0.5 : NP[2][0]), npt1 = 1.0 - npt0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    py[0][k] = npt0 * sp[k] + npt1 * sp[k + 4];
    py[1][k] = npt0 * sp[k + 2] + npt1 * sp[k + 6];
  }
  __syncwarp(mask);

  // ---------------------------------------------------------------- phase B
  const double wt = (t == 1) ? W[1] : W[0];
  double dg0_sum = 0;
  double *lin_lane = A.lin ? A.lin + g * kLinDecade + 2 * lane : nullptr;
#if AB_RES_ROLLJ
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int j = 0; j < 3; ++j) {  // unrolled: no loop-carried register constraints (ns_r1.cuh)
    const double wj = (j == 1) ? W[1] : W[0];
    double val[3][3], dx[3][3], dy[3][3], dz[3][3], hist[3][3];  // [c][qz]
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double *sc = s + c * kR1CompSmem + 3 * j + kR1QyStride * t;
      double T[3], Tx[3], Ty[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        T[k] = sc[k];
        Tx[k] = sc[9 + k];
        Ty[k] = sc[18 + k];
      }
      q2::val(T[0], T[1], T[2], val[c][0], val[c][1], val[c][2]);
      q2::der(T[0], T[1], T[2], dz[c][0], dz[c][1], dz[c][2]);
      q2::val(Tx[0], Tx[1], Tx[2], dx[c][0], dx[c][1], dx[c][2]);
      q2::val(Ty[0], Ty[1], Ty[2], dy[c][0], dy[c][1], dy[c][2]);
      hist[c][0] = hist[c][1] = hist[c][2] = 0;
      if (use_old && !stokes) {
        const double *hc = sh + 3 * j + 9 * c + kResHistQy * t;
        q2::val(hc[0], hc[1], hc[2], hist[c][0], hist[c][1], hist[c][2]);
      }
    }
    double pv[3];
    {
      const double npj0 = j == 0 ? NP[0][0] : (j == 1 ? 0.5 : NP[2][0]), npj1 = 1.0 - npj0;
      const double p0 = npj0 * py[0][0] + npj1 * py[1][0], p1 = npj0 * py[0][1] + npj1 * py[1][1];
#pragma unroll
      for (int q = 0; q < 3; ++q) pv[q] = NP[q][0] * p0 + NP[q][1] * p1 + p_dg0;
    }
    double tp[3];
    const size_t cbase = (size_t)(3 * j) * 30 + lane;
#pragma unroll
    for (int qz = 0; qz < 3; ++qz) {
      const size_t slot = cbase + (size_t)qz * 30;
#if AB_R1_LEAN
      if constexpr (GEOM == ADAFLO_B200_GEOM_CARTESIAN) {
        // lean formulation on axis-aligned cells, constant factors folded (see ns_r1.cuh, same terms as the general branch)
        const double w = G.det * (wj * wt) * W[qz];
        const double g0 = G.K[0][0] * w, g1 = G.K[1][1] * w, g2 = G.K[2][2] * w;
        double gr[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          gr[d][kAxX] = dx[d][qz] * G.K[kAxX][kAxX];
          gr[d][kAxY] = dy[d][qz] * G.K[kAxY][kAxY];
          gr[d][kAxZ] = dz[d][qz] * G.K[kAxZ][kAxZ];
        }
        const double div = gr[0][0] + gr[1][1] + gr[2][2];
        if (!stokes) {
          const double u[3] = {val[0][qz], val[1][qz], val[2][qz]};
          if (lin_lane) {  // linearized_velocities (:791-798)
            double *l = lin_lane + (3 * j + qz) * kLinPointStride;
            if (newton) {
              st_pair(l, u[0], u[1]);
              st_pair(l + kLinPairStride, u[2], gr[0][0]);
              st_pair(l + 2 * kLinPairStride, gr[0][1], gr[0][2]);
              st_pair(l + 3 * kLinPairStride, gr[1][0], gr[1][1]);
              st_pair(l + 4 * kLinPairStride, gr[1][2], gr[2][0]);
              st_pair(l + 5 * kLinPairStride, gr[2][1], gr[2][2]);
            } else {
              st_pair(l, u[0], u[1]);
              st_pair(l + kLinPairStride, u[2], div);
            }
          }
          const double rho = VAR ? __ldg(A.rho + g * 270 + slot) : A.rho0;
          const double damping = VAR ? __ldg(A.damp + g * 270 + slot) : A.damp0;
          const double wR = w * rho, wA = fma(A.gamma, wR, -damping * w), wB = A.tau1 * wR, f1 = A.beta * div;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            double res = f1 * u[d];
#pragma unroll
            for (int e = 0; e < 3; ++e) res = fma(u[e], gr[d][e], res);
            val[d][qz] = fma(wB, res, fma(wR, hist[d][qz], wA * u[d]));  // :727-732, :783-790
          }
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) val[d][qz] = 0.0;
        }
        const double tmu = (VAR ? __ldg(A.mu + g * 270 + slot) : A.mu0) * A.tau1;
        const double tmu2 = tmu + tmu;
        const double q = fma(A.tau_gd, div, -pv[qz]);
        const double s01 = tmu * (gr[0][1] + gr[1][0]), s02 = tmu * (gr[0][2] + gr[2][0]), s12 = tmu * (gr[1][2] + gr[2][1]);
        double tr[3][3];  // [d][mesh axis]
        tr[0][0] = g0 * fma(tmu2, gr[0][0], q);
        tr[1][1] = g1 * fma(tmu2, gr[1][1], q);
        tr[2][2] = g2 * fma(tmu2, gr[2][2], q);
        tr[0][1] = g1 * s01;
        tr[1][0] = g0 * s01;
        tr[0][2] = g2 * s02;
        tr[2][0] = g0 * s02;
        tr[1][2] = g2 * s12;
        tr[2][1] = g1 * s12;
        tp[qz] = -div * w;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          dx[d][qz] = tr[d][kAxX];
          dy[d][qz] = tr[d][kAxY];
          dz[d][qz] = tr[d][kAxZ];
        }
        continue;
      }
#endif
      const double jxw = G.at_blk(A, g, slot, wj * W[qz] * wt);
      double gr[3][3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double r[3];
        r[kAxX] = dx[d][qz];
        r[kAxY] = dy[d][qz];
        r[kAxZ] = dz[d][qz];
        G.to_real(r, gr[d]);
      }
      const double div = gr[0][0] + gr[1][1] + gr[2][2];
      double conv[3] = {0, 0, 0};
      if (!stokes) {
        const double u[3] = {val[0][qz], val[1][qz], val[2][qz]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double res = A.beta * div * u[d];
#pragma unroll
          for (int e = 0; e < 3; ++e) res += u[e] * gr[d][e];
          conv[d] = A.gamma * u[d] + hist[d][qz] + A.tau1 * res;  // :727-732, :783-790
        }
        if (lin_lane) {  // linearized_velocities (:791-798): rows u_d, then d u_d / d x_e (Newton) or the divergence (Picard)
          double *l = lin_lane + (3 * j + qz) * kLinPointStride;
          if (newton) {
            st_pair(l, u[0], u[1]);
            st_pair(l + kLinPairStride, u[2], gr[0][0]);
            st_pair(l + 2 * kLinPairStride, gr[0][1], gr[0][2]);
            st_pair(l + 3 * kLinPairStride, gr[1][0], gr[1][1]);
            st_pair(l + 4 * kLinPairStride, gr[1][2], gr[2][0]);
            st_pair(l + 5 * kLinPairStride, gr[2][1], gr[2][2]);
          } else {
            st_pair(l, u[0], u[1]);
            st_pair(l + kLinPairStride, u[2], div);
          }
        }
        const double rho = VAR ? __ldg(A.rho + g * 270 + slot) : A.rho0;
        const double damping = VAR ? __ldg(A.damp + g * 270 + slot) : A.damp0;
#pragma unroll
        for (int d = 0; d < 3; ++d) conv[d] = (conv[d] * rho - damping * u[d]) * jxw;
      }
      const double tmu = (VAR ? __ldg(A.mu + g * 270 + slot) : A.mu0) * A.tau1;
      double S[3][3];
      S[0][1] = S[1][0] = tmu * (gr[0][1] + gr[1][0]);
      S[0][2] = S[2][0] = tmu * (gr[0][2] + gr[2][0]);
      S[1][2] = S[2][1] = tmu * (gr[1][2] + gr[2][1]);
#pragma unroll
      for (int d = 0; d < 3; ++d) S[d][d] = 2. * tmu * gr[d][d] + A.tau_gd * div - pv[qz];
      tp[qz] = -div * jxw;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double tr[3];
        G.to_ref(S[d], jxw, tr);
        val[d][qz] = conv[d];
        dx[d][qz] = tr[kAxX];
        dy[d][qz] = tr[kAxY];
        dz[d][qz] = tr[kAxZ];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double *sc = s + c * kR1CompSmem + 3 * j + kR1QyStride * t;
      double r[3], rx[3], ry[3];
      q2::val_t(val[c][0], val[c][1], val[c][2], r[0], r[1], r[2]);
      q2::der_t_add(dz[c][0], dz[c][1], dz[c][2], r[0], r[1], r[2]);
      q2::val_t(dx[c][0], dx[c][1], dx[c][2], rx[0], rx[1], rx[2]);
      q2::val_t(dy[c][0], dy[c][1], dy[c][2], ry[0], ry[1], ry[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        sc[k] = r[k];
        sc[9 + k] = rx[k];
        sc[18 + k] = ry[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double pz = NP[0][k] * tp[0] + NP[1][k] * tp[1] + NP[2][k] * tp[2];
      const double npj0 = j == 0 ? NP[0][0] : (j == 1 ? 0.5 : NP[2][0]);
      pacc[0][k] = fma(npj0, pz, pacc[0][k]);
      pacc[1][k] = fma(1.0 - npj0, pz, pacc[1][k]);
    }
    dg0_sum += tp[0] + tp[1] + tp[2];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) sp[t + 3 * (ix + 2 * k)] = pacc[ix][k];
  __syncwarp(mask);

  // ---------------------------------------------------------------- phase C
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double *sc = s + c * kR1CompSmem;
    double Y[3][3], Yd[3][3];
#pragma unroll
    for (int qy = 0; qy < 3; ++qy) {
      double T[3], Tx[3], Ty[3];
#pragma unroll
      for (int qx = 0; qx < 3; ++qx) {
        T[qx] = sc[0 + r1_idx(qx, qy, t)];
        Tx[qx] = sc[9 + r1_idx(qx, qy, t)];
        Ty[qx] = sc[18 + r1_idx(qx, qy, t)];
      }
      q2::val_t(T[0], T[1], T[2], Y[qy][0], Y[qy][1], Y[qy][2]);
      q2::der_t_add(Tx[0], Tx[1], Tx[2], Y[qy][0], Y[qy][1], Y[qy][2]);
      q2::val_t(Ty[0], Ty[1], Ty[2], Yd[qy][0], Yd[qy][1], Yd[qy][2]);
    }
    double o[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      q2::val_t(Y[0][i], Y[1][i], Y[2][i], o[i], o[3 + i], o[6 + i]);
      q2::der_t_add(Yd[0][i], Yd[1][i], Yd[2][i], o[i], o[3 + i], o[6 + i]);
    }
    if (HANGING) {
      if (hn_z_warp) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          double r0 = __shfl_sync(mask, o[j], 3 * cw), r1 = __shfl_sync(mask, o[j], 3 * cw + 1), r2 = __shfl_sync(mask, o[j], 3 * cw + 2);
          hn_q2_t(hn.sub(2), r0, r1, r2);
          if (hn.any() && hn.line(2, j % 3, j / 3, 3)) o[j] = t == 0 ? r0 : (t == 1 ? r1 : r2);
        }
      }
      if (hn.any()) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (hn.line(1, t, i, 3)) hn_q2_t(hn.sub(1), o[i], o[3 + i], o[6 + i]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (hn.line(0, i, t, 3)) hn_q2_t(hn.sub(0), o[3 * i], o[3 * i + 1], o[3 * i + 2]);
      }
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const uint32_t nd = node[j];
      if (!((nd >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) red_add(A.dst_u + 3 * (size_t)(nd & ADAFLO_B200_NODE_MASK) + c, o[j]);
    }
  }
  {
    double o[4] = {0, 0, 0, 0};
    if (t < 2) {
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const double a0 = sp[0 + 3 * (ix + 2 * t)], a1 = sp[1 + 3 * (ix + 2 * t)], a2 = sp[2 + 3 * (ix + 2 * t)];
        o[ix] = NP[0][0] * a0 + NP[1][0] * a1 + NP[2][0] * a2;
        o[2 + ix] = NP[0][1] * a0 + NP[1][1] * a1 + NP[2][1] * a2;
      }
    }
    if (HANGING) {
      if (hnp_z_warp) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double r0 = __shfl_sync(mask, o[j], 3 * cw), r1 = __shfl_sync(mask, o[j], 3 * cw + 1);
          hn_q1_t(hnp.sub(2), r0, r1);
          if (hnp.any() && hnp.line(2, j % 2, j / 2, 2) && t < 2) o[j] = t == 0 ? r0 : r1;
        }
      }
      if (hnp.any() && t < 2) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (hnp.line(1, t, i, 2)) hn_q1_t(hnp.sub(1), o[i], o[2 + i]);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (hnp.line(0, i, t, 2)) hn_q1_t(hnp.sub(0), o[2 * i], o[2 * i + 1]);
      }
    }
    if (t < 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!(pnode[j] >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (pnode[j] & ADAFLO_B200_NODE_MASK), o[j]);
    }
    if (A.np_cell == 9) {
      double sum = dg0_sum;
      sum += __shfl_sync(mask, dg0_sum, 3 * cw + (t + 1) % 3) + __shfl_sync(mask, dg0_sum, 3 * cw + (t + 2) % 3);
      if (t == 0 && !(dg0_node >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (dg0_node & ADAFLO_B200_NODE_MASK), sum);
    }
  }
}

// persistent blocks of kResWarps independent warps (2 per SM sub-partition: up to 255 registers, no spills); decades are
// drawn from the global ticket counter, two tickets ahead
template <int GEOM, bool VAR, bool HANGING>
__global__ void __launch_bounds__(kResWarps * 32, 1) ns_residual_r1_kernel(const NSArgs A)
{
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned n_dec = (unsigned)(A.decade_end - A.decade_begin);
  double *warp_smem = smem + warp * kResWarpSmem;
  unsigned i = 0, i_n1 = 0;
  if (lane == 0) {
    i = atomicAdd(A.work_counter, 1u);
    i_n1 = atomicAdd(A.work_counter, 1u);
  }
  i = __shfl_sync(0xffffffffu, i, 0);
  i_n1 = __shfl_sync(0xffffffffu, i_n1, 0);
  while (i < n_dec) {
    unsigned i_n2 = 0;
    if (lane == 0) i_n2 = atomicAdd(A.work_counter, 1u);
    r1_residual_decade<GEOM, VAR, HANGING>(A, A.decade_begin + i, warp_smem, lane);
    __syncwarp();  // the tiles of this decade have been read before the next decade overwrites them
    i = i_n1;
    i_n1 = __shfl_sync(0xffffffffu, i_n2, 0);
  }
}

template <int GEOM, bool VAR, bool HANGING>
static int launch_res_r1_inst(const NSArgs &a, cudaStream_t s)
{
  auto kern = ns_residual_r1_kernel<GEOM, VAR, HANGING>;
  const size_t smem = sizeof(double) * kResWarps * kResWarpSmem;
  int dev = 0, n_sm = 0;
  AB_CUDA(cudaGetDevice(&dev));
  AB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!a.work_counter) return fail(ADAFLO_B200_ERR_STATE, "ns_residual_r1: no work counter");
  NSArgs b = a;
  const size_t n_dec = n_groups(a.n_cells);
  b.decade_begin = 0;
  b.decade_end = n_dec;
  if (n_dec == 0) return 0;
  const size_t blocks_needed = (n_dec + kResWarps - 1) / kResWarps;
  const unsigned grid = (unsigned)(blocks_needed < (size_t)n_sm ? blocks_needed : (size_t)n_sm);
  AB_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(unsigned), s));
  kern<<<grid, kResWarps * 32, smem, s>>>(b);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// Newton / Picard residual (no extrapolated field); the caller falls back to the round-1 kernel otherwise
int launch_ns_residual_r1(const NSArgs &a, int geom, bool var, cudaStream_t s)
{
  const bool hn = a.hn_u != nullptr;
#define AB_CASE(G, V, H) \
  if (geom == G && var == V && hn == H) return launch_res_r1_inst<G, V, H>(a, s);
  AB_CASE(0, false, false) AB_CASE(0, false, true) AB_CASE(0, true, false) AB_CASE(0, true, true)
  AB_CASE(1, false, false) AB_CASE(1, false, true) AB_CASE(1, true, false) AB_CASE(1, true, true)
  AB_CASE(2, false, false) AB_CASE(2, false, true) AB_CASE(2, true, false) AB_CASE(2, true, true)
#undef AB_CASE
  return fail(ADAFLO_B200_ERR_ARG, "launch_ns_residual_r1: bad variant");
}

}  // namespace ab200
