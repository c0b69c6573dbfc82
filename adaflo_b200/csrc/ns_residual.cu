// ns_residual.cu -- nonlinear residual cell kernel (LocalOps::residual),
// source/navier_stokes_matrix.cc:266-293 and the q-point block :725-800.
// Besides the residual it PRODUCES the linearisation point the vmult kernels
// consume (linearized_velocities, :778-798), which is why it lives on the device.
//
// Same thread mapping as the vmult kernel (ns_q2q1.cuh).  The BDF history enters
// linearly, so w_old*u_old + w_oldold*u_oldold (and, for the extrapolating
// linearisations, the extrapolated velocity) are combined at dof level and
// evaluated once instead of evaluating every old vector separately.
#include <cstdlib>

#include "ns_q2q1.cuh"

namespace ab200 {

// values only (one tile)
__device__ __forceinline__ void evaluate_u_values(double *s, int a, int b, unsigned mask, const double (&col)[3][3],
                                                  double (&val)[3][3])
{
  AB_TABLES_Q2G3
  __syncwarp(mask);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      s[c * kAS + tidx(q, b, a)] = N[q][0] * col[c][0] + N[q][1] * col[c][1] + N[q][2] * col[c][2];
  __syncwarp(mask);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double A0 = s[c * kAS + tidx(b, a, 0)], A1 = s[c * kAS + tidx(b, a, 1)], A2 = s[c * kAS + tidx(b, a, 2)];
#pragma unroll
    for (int q = 0; q < 3; ++q) s[c * kAS + tidx(b, a, q)] = N[q][0] * A0 + N[q][1] * A1 + N[q][2] * A2;
  }
  __syncwarp(mask);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double A0 = s[c * kAS + tidx(b, 0, a)], A1 = s[c * kAS + tidx(b, 1, a)], A2 = s[c * kAS + tidx(b, 2, a)];
#pragma unroll
    for (int q = 0; q < 3; ++q) val[c][q] = N[q][0] * A0 + N[q][1] * A1 + N[q][2] * A2;
  }
}

template <int GEOM, bool VAR, bool EXTRAP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2) ns_residual_kernel(const NSArgs A)
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

  const bool stokes = A.physical_type == ADAFLO_B200_STOKES;
  const bool use_old = A.physical_type == ADAFLO_B200_INCOMPRESSIBLE;
  const int lz = A.linearization;

  // ---- gather, plain (read_dof_values_plain, :663-664,692-693)
  double col[3][3];
  uint32_t node[3];
  gather_u(A.cell_u, A.src_u, cell, t9, false, col, node);
  // hanging nodes: the plain read also needs the interpolated values (the hanging dofs themselves are not stored)
  const HnInfo hn(A.hn_u ? __ldg(A.hn_u + cell) : 0u), hnp(A.hn_p ? __ldg(A.hn_p + cell) : 0u);
  const bool hn_warp = A.hn_u && __any_sync(mask, hn.any() || hnp.any());
  hn_resolve9_u(hn, hn_warp, a, b, 9 * cw, mask, col);
  double pc[2] = {0, 0};
  uint32_t pnode[2] = {0, 0};
  if (a < 2 && b < 2) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      pnode[k] = __ldg(A.cell_p + (size_t)cell * A.np_cell + 4 * k + 2 * b + a);
      pc[k] = A.src_p[pnode[k] & ADAFLO_B200_NODE_MASK];
    }
  }
  hn_resolve9_p(hnp, hn_warp, a, b, 9 * cw, mask, pc);
  double p_dg0 = 0;
  uint32_t dg0_node = 0;
  if (A.np_cell == 9) {
    dg0_node = __ldg(A.cell_p + (size_t)cell * 9 + 8);
    p_dg0 = A.src_p[dg0_node & ADAFLO_B200_NODE_MASK];
  }
  Geometry<GEOM> G;
  G.init(A, cell);

  double val[3][3], dx[3][3], dy[3][3], dz[3][3];
  evaluate_u<true>(s, a, b, mask, col, val, dx, dy, dz);

  // BDF history: hist = w_old u_old + w_oo u_oo ; extrapolation e = ex_old u_old + ex_oo u_oo
  double hist[3][3];
  double ev[3][3], ex[3][3], ey[3][3], ez[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int q = 0; q < 3; ++q) hist[c][q] = 0;
  if (use_old && !stokes) {
    double co[3][3], coo[3][3], cc[3][3];
    uint32_t dummy[3];
    gather_u(A.cell_u, A.old_u, cell, t9, false, co, dummy);
    gather_u(A.cell_u, A.old_old_u, cell, t9, false, coo, dummy);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) cc[c][k] = A.w_old * co[c][k] + A.w_oo * coo[c][k];
    hn_resolve9_u(hn, hn_warp, a, b, 9 * cw, mask, cc);
    evaluate_u_values(s, a, b, mask, cc, hist);
    if (EXTRAP) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) cc[c][k] = A.ex_old * co[c][k] + A.ex_oo * coo[c][k];
      hn_resolve9_u(hn, hn_warp, a, b, 9 * cw, mask, cc);
      evaluate_u<true>(s, a, b, mask, cc, ev, ex, ey, ez);
    }
  }
  double pv[3];
  evaluate_p(sp, a, b, mask, pc, pv);
#pragma unroll
  for (int q = 0; q < 3; ++q) pv[q] += p_dg0;

  double tp[3];
#pragma unroll
  for (int qy = 0; qy < 3; ++qy) {
    double *lin = A.lin ? A.lin + lin_index(cell, a, qy, b) : nullptr;  // row r at lin_row(r)
    const size_t q1 = coef_index(cell, 0, 1, a, qy, b);
    const double jxw = G.at(a, qy, b, W[a] * W[qy] * W[b]);
    double g[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double r[3] = {dx[d][qy], dy[d][qy], dz[d][qy]};
      G.to_real(r, g[d]);
    }
    const double div = g[0][0] + g[1][1] + g[2][2];
    double conv[3] = {0, 0, 0};
    if (!stokes) {
      const double u[3] = {val[0][qy], val[1][qy], val[2][qy]};
#pragma unroll
      for (int d = 0; d < 3; ++d) conv[d] = A.gamma * u[d] + hist[d][qy];
      if (EXTRAP) {
        double eg[3][3], eu[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double r[3] = {ex[d][qy], ey[d][qy], ez[d][qy]};
          G.to_real(r, eg[d]);
          eu[d] = ev[d][qy];
        }
        const double ediv = eg[0][0] + eg[1][1] + eg[2][2];
        if (lz == ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            double res = A.beta * ediv * eu[d];
#pragma unroll
            for (int e = 0; e < 3; ++e) res += eu[e] * eg[d][e];
            conv[d] += A.tau1 * res;
          }
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            double res = A.beta * ediv * u[d];
#pragma unroll
            for (int e = 0; e < 3; ++e) res += eu[e] * g[d][e];
            conv[d] += A.tau1 * res;
            lin[lin_row(d)] = eu[d];
          }
          lin[lin_row(3)] = ediv;
        }
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double res = A.beta * div * u[d];
#pragma unroll
          for (int e = 0; e < 3; ++e) res += u[e] * g[d][e];
          conv[d] += A.tau1 * res;
        }
        if (lz == ADAFLO_B200_COUPLED_IMPLICIT_NEWTON) {
          // the 12 rows are interleaved in pairs (lin_index): six 16-byte stores
          const double L[12] = {u[0], u[1], u[2], g[0][0], g[0][1], g[0][2], g[1][0], g[1][1], g[1][2], g[2][0], g[2][1], g[2][2]};
#pragma unroll
          for (int r = 0; r < 6; ++r) *reinterpret_cast<double2 *>(lin + lin_row(2 * r)) = make_double2(L[2 * r], L[2 * r + 1]);
        } else {
          *reinterpret_cast<double2 *>(lin + lin_row(0)) = make_double2(u[0], u[1]);
          *reinterpret_cast<double2 *>(lin + lin_row(2)) = make_double2(u[2], div);
        }
      }
      const double rho = VAR ? __ldg(A.rho + q1) : A.rho0;
      const double damping = VAR ? __ldg(A.damp + q1) : A.damp0;
#pragma unroll
      for (int d = 0; d < 3; ++d) conv[d] = (conv[d] * rho - damping * u[d]) * jxw;
    }
    const double tmu = (VAR ? __ldg(A.mu + q1) : A.mu0) * A.tau1;
    double S[3][3];
    S[0][1] = S[1][0] = tmu * (g[0][1] + g[1][0]);
    S[0][2] = S[2][0] = tmu * (g[0][2] + g[2][0]);
    S[1][2] = S[2][1] = tmu * (g[1][2] + g[2][1]);
#pragma unroll
    for (int d = 0; d < 3; ++d) S[d][d] = 2. * tmu * g[d][d] + A.tau_gd * div - pv[qy];
    tp[qy] = -div * jxw;
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

  double out[3][3];
  __syncwarp(mask);
  integrate_u<true>(s, a, b, mask, val, dx, dy, dz, out);
  hn_distribute9_u(hn, hn_warp, a, b, 9 * cw, mask, out);
  scatter_u(A.dst_u, node, out);
  if (lz != ADAFLO_B200_PROJECTION) {  // :902-907
    double outp[2];
    integrate_p(sp, a, b, mask, tp, outp);
    hn_distribute9_p(hnp, hn_warp, a, b, 9 * cw, mask, outp);
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

template <int GEOM, bool VAR, bool EXTRAP>
static int launch_res_inst(const NSArgs &a, cudaStream_t s)
{
  auto kern = ns_residual_kernel<GEOM, VAR, EXTRAP>;
  const size_t smem = sizeof(double) * kWarpsPerBlock * kWarpSmem;
  {  // per device and cheap: set on every launch (a process may drive several GPUs)
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const size_t cells_per_block = (size_t)kWarpsPerBlock * kCellsPerWarp;
  const unsigned grid = (unsigned)((a.n_cells + cells_per_block - 1) / cells_per_block);
  if (grid == 0) return 0;
  kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(a);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_ns_residual(const NSArgs &a, int geom, bool var, cudaStream_t s)
{
  const int lz = a.linearization;
  const bool extrap = (lz == ADAFLO_B200_PROJECTION || lz == ADAFLO_B200_COUPLED_VELOCITY_SEMI_IMPLICIT ||
                       lz == ADAFLO_B200_COUPLED_VELOCITY_EXPLICIT) &&
                      a.physical_type != ADAFLO_B200_STOKES;
  if (extrap && a.physical_type != ADAFLO_B200_INCOMPRESSIBLE)
    return fail(ADAFLO_B200_ERR_UNSUPPORTED, "extrapolating linearization requires physical_type == incompressible");
  // Newton / Picard / Stokes: the kernel on the vmult mapping (ns_residual_r1.cu); the extrapolating linearisations need a
  // third evaluated field and stay on this file's 9-lane kernel.  ADAFLO_B200_RESIDUAL_KERNEL=v1 forces the latter (A/B).
  static const bool force_v1 = [] {
    const char *e = std::getenv("ADAFLO_B200_RESIDUAL_KERNEL");
    return e && std::string(e) == "v1";
  }();
  if (!extrap && !force_v1 && a.cell_u_blk && a.cell_p_blk) return launch_ns_residual_r1(a, geom, var, s);
#define AB_CASE(G, V, E) \
  if (geom == G && var == V && extrap == E) return launch_res_inst<G, V, E>(a, s);
  AB_CASE(0, false, false) AB_CASE(0, false, true) AB_CASE(0, true, false) AB_CASE(0, true, true)
  AB_CASE(1, false, false) AB_CASE(1, false, true) AB_CASE(1, true, false) AB_CASE(1, true, true)
  AB_CASE(2, false, false) AB_CASE(2, false, true) AB_CASE(2, true, false) AB_CASE(2, true, true)
#undef AB_CASE
  return fail(ADAFLO_B200_ERR_ARG, "launch_ns_residual: bad variant");
}

}  // namespace ab200
