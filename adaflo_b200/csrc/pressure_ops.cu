// pressure_ops.cu -- the sub-block operators the block preconditioner calls:
//   divergence_vmult_add      source/navier_stokes_matrix.cc:300-332, kernel :920-961
//   pressure_mass_vmult       :421-455, kernel :1036-1071
//   pressure_poisson_vmult    :386-417, kernel :965-1032
//   pressure_convdiff_vmult   :459-483, kernel :1099-1140
//   local_pressure_mass_weight :1075-1095 (null-space weights in initialize)
//
// The pressure space is tiny next to the velocity space (1 of 25 dofs per cell),
// so the scalar Q1 operators use one thread per cell with the 8 (+1 DG0) local
// dofs in registers; the divergence reuses the 9-threads-per-cell evaluation of
// the coupled kernel because it reads the velocity.
#include "ns_q2q1.cuh"

namespace ab200 {

// ---------------------------------------------------------------- divergence
template <int GEOM, int WEIGHT /*0: -1, 1: -mu0, 2: -mu_q*/>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4) divergence_kernel(const PArgs P, const NSArgs A)
{
  AB_TABLES_Q2G3
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = lane / 9;
  const int t9 = lane - 9 * cw;
  const int a = t9 % 3, b = t9 / 3;
  const size_t triple = (size_t)blockIdx.x * kWarpsPerBlock + warp;
  const size_t cell_l = triple * kCellsPerWarp + cw;
  const bool active = (cw < kCellsPerWarp) && (cell_l < P.n_cells);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t cell = (uint32_t)cell_l;
  double *s = smem + warp * kWarpSmem + cw * kCellSmem;
  double *sp = s + 9 * kAS;

  double col[3][3];
  uint32_t node[3];
  gather_u(P.cell_u, P.src, cell, t9, P.resolve != 0, col, node);
  const HnInfo hn(P.hn_u ? __ldg(P.hn_u + cell) : 0u), hnp(P.hn_p ? __ldg(P.hn_p + cell) : 0u);
  const bool hn_warp = P.hn_u && __any_sync(mask, hn.any() || hnp.any());
  hn_resolve9_u(hn, hn_warp, a, b, 9 * cw, mask, col);
  Geometry<GEOM> G;
  G.init(A, cell);
  double val[3][3], dx[3][3], dy[3][3], dz[3][3];
  evaluate_u<false>(s, a, b, mask, col, val, dx, dy, dz);
  double tp[3];
#pragma unroll
  for (int qy = 0; qy < 3; ++qy) {
    const double jxw = G.at(a, qy, b, W[a] * W[qy] * W[b]);
    double div = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double r[3] = {dx[d][qy], dy[d][qy], dz[d][qy]};
      double g[3];
      G.to_real(r, g);
      div += g[d];
    }
    const double w = WEIGHT == 0 ? -1.0 : (WEIGHT == 1 ? -P.c0 : -__ldg(P.coeff_q + coef_index(cell, 0, 1, a, qy, b)));
    tp[qy] = w * div * jxw;
  }
  double outp[2];
  integrate_p(sp, a, b, mask, tp, outp);
  hn_distribute9_p(hnp, hn_warp, a, b, 9 * cw, mask, outp);
  if (a < 2 && b < 2) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t pn = __ldg(P.cell_p + (size_t)cell * P.np_cell + 4 * k + 2 * b + a);
      if (!(pn >> ADAFLO_B200_FLAG_SHIFT)) red_add(P.dst + (pn & ADAFLO_B200_NODE_MASK), outp[k]);
    }
  }
  if (P.np_cell == 9) {
    const double sum = cell_sum9(tp[0] + tp[1] + tp[2], 9 * cw, t9, mask);
    const uint32_t pn = __ldg(P.cell_p + (size_t)cell * 9 + 8);
    if (t9 == 0 && !(pn >> ADAFLO_B200_FLAG_SHIFT)) red_add(P.dst + (pn & ADAFLO_B200_NODE_MASK), sum);
  }
}

template <int GEOM>
static int launch_div_geom(const PArgs &p, const NSArgs &a, int weight, cudaStream_t s)
{
  const size_t smem = sizeof(double) * kWarpsPerBlock * kWarpSmem;
  const size_t cpb = (size_t)kWarpsPerBlock * kCellsPerWarp;
  const unsigned grid = (unsigned)((p.n_cells + cpb - 1) / cpb);
  if (grid == 0) return 0;
#define AB_L(Wt)                                                                                          \
  {                                                                                                       \
    auto kern = divergence_kernel<GEOM, Wt>;                                                              \
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(p, a);                                                   \
  }
  if (weight == 0) AB_L(0) else if (weight == 1) AB_L(1) else AB_L(2)
#undef AB_L
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_divergence(const PArgs &p, bool weight_by_viscosity, cudaStream_t s)
{
  NSArgs a{};
  a.geom_row = p.geom_row;
  a.geom_table = p.geom_table;
  a.inv_jac = p.inv_jac;
  a.jxw = p.jxw;
  const int weight = !weight_by_viscosity ? 0 : (p.coeff_q ? 2 : 1);
  if (p.geom == ADAFLO_B200_GEOM_CARTESIAN) return launch_div_geom<ADAFLO_B200_GEOM_CARTESIAN>(p, a, weight, s);
  if (p.geom == ADAFLO_B200_GEOM_AFFINE) return launch_div_geom<ADAFLO_B200_GEOM_AFFINE>(p, a, weight, s);
  return launch_div_geom<ADAFLO_B200_GEOM_GENERAL>(p, a, weight, s);
}

// ------------------------------------------------- scalar Q1 pressure kernels
// kind 0: mass (q <- c p), 1: laplace (grad q <- c grad p), 2: mass weight (q <- 1)
// coeff_mode: 0: c = c0
//             1: c = 1 / (coeff_q[cell][centre] + c1)         pressure mass, variable viscosity (:1056-1063)
//             2: c = 1 / (c0 * coeff_q[cell][q])              pressure Poisson, variable density per point (:992-996)
//             3: c = 1 / (c0 * coeff_q[cell][centre])         pressure Poisson, cell value (:1014-1023)
//             4: c = coeff_q[cell][centre]                    pressure conv-diff (:1123-1130)
template <int NQ1>
struct Q1Tables;
template <>
struct Q1Tables<2> {
  static __device__ __forceinline__ double n(int q, int i)
  {
    constexpr double T[2][2] = {{0.7886751345948129, 0.2113248654051871}, {0.2113248654051871, 0.7886751345948129}};
    return T[q][i];
  }
  static __device__ __forceinline__ double w(int) { return 0.5; }
};
template <>
struct Q1Tables<3> {
  static __device__ __forceinline__ double n(int q, int i)
  {
    constexpr double T[3][2] = {{0.8872983346207417, 0.11270166537925831}, {0.5, 0.5}, {0.11270166537925831, 0.8872983346207417}};
    return T[q][i];
  }
  static __device__ __forceinline__ double w(int q)
  {
    constexpr double Wt[3] = {0.2777777777777778, 0.4444444444444444, 0.2777777777777778};
    return Wt[q];
  }
};

template <int NQ1, int KIND>
__global__ void __launch_bounds__(128) pressure_kernel(const PArgs P)
{
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.n_cells) return;
  constexpr int NQ = NQ1 * NQ1 * NQ1;
  const bool dg0 = P.np_cell == 9;
  uint32_t pn[9];
  double pl[9], out[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    pl[i] = 0;
    out[i] = 0;
    pn[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) pn[i] = __ldg(P.cell_p + (size_t)cell * P.np_cell + i);
  if (dg0) pn[8] = __ldg(P.cell_p + (size_t)cell * 9 + 8);
  if (KIND != 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double v = P.src[pn[i] & ADAFLO_B200_NODE_MASK];
      pl[i] = (pn[i] >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : v;
    }
    if (dg0) pl[8] = P.src[pn[8] & ADAFLO_B200_NODE_MASK];
  }
  const HnInfo hnp(P.hn_p ? __ldg(P.hn_p + cell) : 0u);
  if (KIND != 2 && hnp.any()) {
    // all lines of one direction before the next direction (single lines do not commute)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(0, i, j, 2)) hn_q1(hnp.sub(0), pl[2 * i + 4 * j], pl[1 + 2 * i + 4 * j]);   // x line at (iy = i, iz = j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(1, i, j, 2)) hn_q1(hnp.sub(1), pl[j + 4 * i], pl[j + 2 + 4 * i]);           // y line at (iz = i, ix = j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(2, i, j, 2)) hn_q1(hnp.sub(2), pl[i + 2 * j], pl[i + 2 * j + 4]);           // z line at (ix = i, iy = j)
  }
  // geometry
  double K[3][3] = {{0}}, det = 0;
  if (P.geom == ADAFLO_B200_GEOM_CARTESIAN) {
    const double *row = P.geom_table + 4 * (size_t)P.geom_row[cell];
    K[0][0] = row[0]; K[1][1] = row[1]; K[2][2] = row[2]; det = row[3];
  } else if (P.geom == ADAFLO_B200_GEOM_AFFINE) {
    const double *row = P.geom_table + 10 * (size_t)P.geom_row[cell];
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int e = 0; e < 3; ++e) K[f][e] = row[3 * f + e];
    det = row[9];
  }
  double c_cell = P.c0;
  if (P.coeff_mode == 1) c_cell = 1.0 / (P.coeff_q[coef_index(cell, 0, 1, 1, 1, 1)] + P.c1);
  if (P.coeff_mode == 3) c_cell = 1.0 / (P.c0 * P.coeff_q[coef_index(cell, 0, 1, 1, 1, 1)]);
  if (P.coeff_mode == 4) c_cell = P.coeff_q[coef_index(cell, 0, 1, 1, 1, 1)];

#pragma unroll
  for (int qz = 0; qz < NQ1; ++qz)
#pragma unroll
    for (int qy = 0; qy < NQ1; ++qy)
#pragma unroll
      for (int qx = 0; qx < NQ1; ++qx) {
        const int q = qx + NQ1 * (qy + NQ1 * qz);
        double jxw;
        if (P.geom == ADAFLO_B200_GEOM_GENERAL) {
#pragma unroll
          for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int e = 0; e < 3; ++e)
              K[f][e] = NQ1 == 3 ? P.inv_jac[coef_index(cell, 3 * f + e, 9, qx, qy, qz)] : P.inv_jac[((size_t)cell * 9 + 3 * f + e) * NQ + q];
          jxw = NQ1 == 3 ? P.jxw[coef_index(cell, 0, 1, qx, qy, qz)] : P.jxw[(size_t)cell * NQ + q];
        } else
          jxw = det * Q1Tables<NQ1>::w(qx) * Q1Tables<NQ1>::w(qy) * Q1Tables<NQ1>::w(qz);
        const double c = (P.coeff_mode == 2) ? 1.0 / (P.c0 * P.coeff_q[coef_index(cell, 0, 1, qx, qy, qz)]) : c_cell;
        if (KIND == 0 || KIND == 2) {
          double v = 1.0;
          if (KIND == 0) {
            v = 0;
#pragma unroll
            for (int iz = 0; iz < 2; ++iz)
#pragma unroll
              for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                for (int ix = 0; ix < 2; ++ix)
                  v += Q1Tables<NQ1>::n(qx, ix) * Q1Tables<NQ1>::n(qy, iy) * Q1Tables<NQ1>::n(qz, iz) * pl[ix + 2 * iy + 4 * iz];
            v = (v + pl[8]) * c;
          }
          v *= jxw;
#pragma unroll
          for (int iz = 0; iz < 2; ++iz)
#pragma unroll
            for (int iy = 0; iy < 2; ++iy)
#pragma unroll
              for (int ix = 0; ix < 2; ++ix)
                out[ix + 2 * iy + 4 * iz] += Q1Tables<NQ1>::n(qx, ix) * Q1Tables<NQ1>::n(qy, iy) * Q1Tables<NQ1>::n(qz, iz) * v;
          out[8] += v;
        } else {
          double gr[3] = {0, 0, 0};
#pragma unroll
          for (int iz = 0; iz < 2; ++iz)
#pragma unroll
            for (int iy = 0; iy < 2; ++iy)
#pragma unroll
              for (int ix = 0; ix < 2; ++ix) {
                const double v = pl[ix + 2 * iy + 4 * iz];
                const double sx = ix ? 1.0 : -1.0, sy = iy ? 1.0 : -1.0, sz = iz ? 1.0 : -1.0;
                gr[0] += sx * Q1Tables<NQ1>::n(qy, iy) * Q1Tables<NQ1>::n(qz, iz) * v;
                gr[1] += Q1Tables<NQ1>::n(qx, ix) * sy * Q1Tables<NQ1>::n(qz, iz) * v;
                gr[2] += Q1Tables<NQ1>::n(qx, ix) * Q1Tables<NQ1>::n(qy, iy) * sz * v;
              }
          double g[3], t[3];
          if (P.geom == ADAFLO_B200_GEOM_CARTESIAN) {
#pragma unroll
            for (int e = 0; e < 3; ++e) g[e] = gr[e] * K[e][e] * c;
#pragma unroll
            for (int f = 0; f < 3; ++f) t[f] = K[f][f] * g[f] * jxw;
          } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) g[e] = (gr[0] * K[0][e] + gr[1] * K[1][e] + gr[2] * K[2][e]) * c;
#pragma unroll
            for (int f = 0; f < 3; ++f) t[f] = (K[f][0] * g[0] + K[f][1] * g[1] + K[f][2] * g[2]) * jxw;
          }
#pragma unroll
          for (int iz = 0; iz < 2; ++iz)
#pragma unroll
            for (int iy = 0; iy < 2; ++iy)
#pragma unroll
              for (int ix = 0; ix < 2; ++ix) {
                const double sx = ix ? 1.0 : -1.0, sy = iy ? 1.0 : -1.0, sz = iz ? 1.0 : -1.0;
                out[ix + 2 * iy + 4 * iz] += sx * Q1Tables<NQ1>::n(qy, iy) * Q1Tables<NQ1>::n(qz, iz) * t[0] +
                                             Q1Tables<NQ1>::n(qx, ix) * sy * Q1Tables<NQ1>::n(qz, iz) * t[1] +
                                             Q1Tables<NQ1>::n(qx, ix) * Q1Tables<NQ1>::n(qy, iy) * sz * t[2];
              }
        }
      }
  if (hnp.any()) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(2, i, j, 2)) hn_q1_t(hnp.sub(2), out[i + 2 * j], out[i + 2 * j + 4]);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(1, i, j, 2)) hn_q1_t(hnp.sub(1), out[j + 4 * i], out[j + 2 + 4 * i]);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (hnp.line(0, i, j, 2)) hn_q1_t(hnp.sub(0), out[2 * i + 4 * j], out[1 + 2 * i + 4 * j]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (!(pn[i] >> ADAFLO_B200_FLAG_SHIFT)) red_add(P.dst + (pn[i] & ADAFLO_B200_NODE_MASK), out[i]);
  if (dg0 && !(pn[8] >> ADAFLO_B200_FLAG_SHIFT)) red_add(P.dst + (pn[8] & ADAFLO_B200_NODE_MASK), out[8]);
}

int launch_pressure_op(const PArgs &p, int kind, int n_q_1d, cudaStream_t s)
{
  const unsigned grid = (p.n_cells + 127) / 128;
  if (grid == 0) return 0;
#define AB_L(NQ1, KIND) pressure_kernel<NQ1, KIND><<<grid, 128, 0, s>>>(p)
  if (n_q_1d == 2 && kind == 0) AB_L(2, 0);
  else if (n_q_1d == 2 && kind == 1) AB_L(2, 1);
  else if (n_q_1d == 2 && kind == 2) AB_L(2, 2);
  else if (n_q_1d == 3 && kind == 0) AB_L(3, 0);
  else if (n_q_1d == 3 && kind == 1) AB_L(3, 1);
  else if (n_q_1d == 3 && kind == 2) AB_L(3, 2);
  else return fail(ADAFLO_B200_ERR_ARG, "launch_pressure_op: bad variant");
#undef AB_L
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
