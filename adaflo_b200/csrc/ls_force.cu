// ls_force.cu -- surface tension / gravity force and Heaviside function:
//   LevelSetOKZSolver::local_compute_force   source/level_set_okz.cc:317-409
//   LevelSetOKZSolver::compute_heaviside     source/level_set_okz.cc:479-540
//   discrete_heaviside                       include/adaflo/level_set_base.h:122-145
//
// The force kernel lives on the Navier-Stokes quadrature (QGauss(3), 27 points): it evaluates the
// level-set-space fields H (Heaviside) and kappa (curvature) there, writes rho_q / mu_q into the
// operator's decade-blocked coefficient arrays and tests sigma kappa grad(H) - g rho e_z with the Q2
// velocity basis.  It reuses the 9-lanes-per-cell integration of ns_q2q1.cuh (a 0.3 % kernel).
#include "ls_ops.cuh"
#include "ns_q2q1.cuh"

namespace ab200 {

template <int GEOM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ls_force_kernel(const ForceArgs F)
{
  AB_TABLES_Q2G3
  const NSArgs &A = F.ns;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = lane / 9;
  const int t9 = lane - 9 * cw;
  const int a = t9 % 3, b = t9 / 3;  // (qx, qz) during the point loop, (ix, iy) at the scatter
  const size_t cell_l = ((size_t)blockIdx.x * kWarpsPerBlock + warp) * kCellsPerWarp + cw;
  const bool active = (cw < kCellsPerWarp) && (cell_l < A.n_cells);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t cell = (uint32_t)cell_l;
  double *s = smem + warp * kWarpSmem + cw * kCellSmem;
  const int n1 = F.n_sub + 1, NLS = n1 * n1 * n1;
  const bool ext = F.hang.cell_ext != nullptr;  // extended slot indices: real dof or n_real + hanging id
  const uint32_t *cl = (ext ? F.hang.cell_ext : F.cell_ls) + (size_t)cell * NLS;
  const uint32_t n_real = F.hang.n_real;
  auto heav = [&](uint32_t e) { return (ext && e >= n_real) ? F.heaviside_hang[e - n_real] : F.heaviside[e]; };
  auto curv = [&](uint32_t e) { return (ext && e >= n_real) ? F.curvature_hang[e - n_real] : F.curvature[e]; };
  Geometry<GEOM> G;
  G.init(A, cell);

  // level set -> pressure interpolation of the Heaviside values (8 corner values of a Q1 function)
  double hp[8];
  if (F.interp) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hp[i] = 0;
    for (int j = 0; j < NLS; ++j) {
      const double h = heav(ext ? cl[j] : (cl[j] & ADAFLO_B200_NODE_MASK));
#pragma unroll
      for (int i = 0; i < 8; ++i) hp[i] = fma(F.interp[i * NLS + j], h, hp[i]);
    }
  }
  double tv[3][3], zero3[3][3];
#pragma unroll
  for (int qy = 0; qy < 3; ++qy) {
    // dense evaluation of H, grad H and kappa at the point (qx = a, qy, qz = b)
    double H = 0, kap = 0, gr[3] = {0, 0, 0};
    for (int k = 0; k < n1; ++k)
      for (int j = 0; j < n1; ++j)
        for (int i = 0; i < n1; ++i) {
          const double vx = F.hat_v[a][i], vy = F.hat_v[qy][j], vz = F.hat_v[b][k];
          const double w = vx * vy * vz;
          if (w == 0.0 && F.hat_d[a][i] * vy * vz == 0.0 && vx * F.hat_d[qy][j] * vz == 0.0 && vx * vy * F.hat_d[b][k] == 0.0) continue;
          const uint32_t d = ext ? cl[i + n1 * (j + n1 * k)] : (cl[i + n1 * (j + n1 * k)] & ADAFLO_B200_NODE_MASK);
          const double h = heav(d);
          H = fma(w, h, H);
          kap = fma(w, curv(d), kap);
          gr[0] = fma(F.hat_d[a][i] * vy * vz, h, gr[0]);
          gr[1] = fma(vx * F.hat_d[qy][j] * vz, h, gr[1]);
          gr[2] = fma(vx * vy * F.hat_d[b][k], h, gr[2]);
        }
    if (F.interp) {  // gradient of the Q1 interpolant instead, level_set_okz.cc:367-378
      const double x = a == 0 ? 0.11270166537925831 : (a == 1 ? 0.5 : 0.8872983346207417);
      const double y = qy == 0 ? 0.11270166537925831 : (qy == 1 ? 0.5 : 0.8872983346207417);
      const double z = b == 0 ? 0.11270166537925831 : (b == 1 ? 0.5 : 0.8872983346207417);
      gr[0] = gr[1] = gr[2] = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double vx = (i & 1) ? x : 1 - x, vy = (i & 2) ? y : 1 - y, vz = (i & 4) ? z : 1 - z;
        gr[0] = fma(((i & 1) ? 1.0 : -1.0) * vy * vz, hp[i], gr[0]);
        gr[1] = fma(vx * ((i & 2) ? 1.0 : -1.0) * vz, hp[i], gr[1]);
        gr[2] = fma(vx * vy * ((i & 4) ? 1.0 : -1.0), hp[i], gr[2]);
      }
    }
    const double jxw = G.at(a, qy, b, W[a] * W[qy] * W[b]);
    double g[3];
    G.to_real(gr, g);
    double rho = F.density;
    if (F.rho_q) {  // :351-364
      rho = F.density + F.density_diff * H;
      const size_t q1 = coef_index(cell, 0, 1, a, qy, b);
      F.rho_q[q1] = rho;
      F.mu_q[q1] = F.viscosity + F.viscosity_diff * H;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double f = F.surface_tension * kap * g[d];
      if (d == F.gravity_dir) f -= F.gravity * rho;  // force[dim - 1], :396-400
      tv[d][qy] = f * jxw;
      zero3[d][qy] = 0;
    }
  }
  double out[3][3];
  __syncwarp(mask);
  integrate_u<true>(s, a, b, mask, tv, zero3, zero3, zero3, out);
  uint32_t node[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) node[k] = __ldg(A.cell_u + (size_t)cell * 27 + 9 * k + t9);
  const HnInfo hn(A.hn_u ? __ldg(A.hn_u + cell) : 0u);
  const bool hn_warp = A.hn_u && __any_sync(mask, hn.any());
  hn_distribute9_u(hn, hn_warp, a, b, 9 * cw, mask, out);
  scatter_u(A.dst_u, node, out);
}

int launch_ls_force(const ForceArgs &f, int geom, cudaStream_t s)
{
  const size_t smem = sizeof(double) * kWarpsPerBlock * kWarpSmem;
  const size_t cpb = (size_t)kWarpsPerBlock * kCellsPerWarp;
  const unsigned grid = (unsigned)((f.ns.n_cells + cpb - 1) / cpb);
  if (grid == 0) return 0;
  if (geom == ADAFLO_B200_GEOM_CARTESIAN)
    ls_force_kernel<ADAFLO_B200_GEOM_CARTESIAN><<<grid, kWarpsPerBlock * 32, smem, s>>>(f);
  else if (geom == ADAFLO_B200_GEOM_AFFINE)
    ls_force_kernel<ADAFLO_B200_GEOM_AFFINE><<<grid, kWarpsPerBlock * 32, smem, s>>>(f);
  else
    ls_force_kernel<ADAFLO_B200_GEOM_GENERAL><<<grid, kWarpsPerBlock * 32, smem, s>>>(f);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ Heaviside
// integral of Peskin's discrete delta, include/adaflo/level_set_base.h:122-145
__device__ double discrete_heaviside(double x)
{
  const bool pos = x > 0;
  if (pos) x = -x;
  double r;
  if (x < -2.)
    r = 0;
  else if (x < -1.)
    r = 1. / 8. * (5. * x + x * x) + 1. / 32. * (-3. - 2. * x) * sqrt(-7. - 12. * x - 4. * x * x) -
        1. / 16 * asin(sqrt(2.) * (x + 3. / 2.)) + 23. / 32. - 3.14159265358979323846 / 64.;
  else
    r = 1. / 8. * (3. * x + x * x) - 1. / 32. * (-1. - 2. * x) * sqrt(1. - 4. * x - 4. * x * x) +
        1. / 16 * asin(sqrt(2.) * (x + 1. / 2.)) + 15. / 32. - 3.14159265358979323846 / 64.;
  return pos ? 1. - r : r;
}

// value of the (extended) dof e: real dofs from the vector, hanging dofs from the interpolated side array
__device__ __forceinline__ double ext_value(const double *c, const double *c_hang, uint32_t n_real, uint32_t e)
{
  return e < n_real ? c[e] : c_hang[e - n_real];
}
// hanging dof = mean of its parents' values (FE_Q_iso_Q1 hats): out[h] = sum vec[idx] / n
__global__ void k_hanging_mean(const LsHanging hang, const double *__restrict__ vec, double *__restrict__ out)
{
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= hang.n_hang) return;
  const uint32_t n = hang.hang_n[h];
  double v = 0;
  for (uint32_t i = 0; i < n; ++i) v += vec[hang.hang_idx[(size_t)h * 8 + i]];
  out[h] = v / (double)n;
}
// pass 1: per cell, "is any |c_i| below the cut-off" and the sign of the first dof (level_set_okz.cc:497-503,530-536)
__global__ void k_heaviside_cells(const uint32_t *__restrict__ cell_ls, const double *__restrict__ c, const double *__restrict__ c_hang,
                                  uint32_t n_real, bool ext, int N, uint32_t n_cells, double cutoff, signed char *__restrict__ cell_state)
{
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  bool consider = false;
  for (int i = 0; i < N; ++i) {
    const uint32_t e = cell_ls[(size_t)cell * N + i];
    consider |= fabs(ext ? ext_value(c, c_hang, n_real, e) : c[e & ADAFLO_B200_NODE_MASK]) < cutoff;
  }
  const uint32_t e0 = cell_ls[(size_t)cell * N];
  const double c0 = ext ? ext_value(c, c_hang, n_real, e0) : c[e0 & ADAFLO_B200_NODE_MASK];
  cell_state[cell] = consider ? 2 : (c0 < 0 ? 0 : 1);
}
// pass 2: per dof (real, then hanging), the value written by the LAST cell of the cell loop that contains the dof
__global__ void k_heaviside_dofs(const uint32_t *__restrict__ last_cell, const signed char *__restrict__ cell_state, const double *__restrict__ c,
                                 const double *__restrict__ c_hang, double scale, double cutoff, double *__restrict__ heaviside,
                                 double *__restrict__ heaviside_hang, uint64_t n_real, uint64_t n)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const signed char st = cell_state[last_cell[i]];
  double h;
  if (st == 2) {
    const double cv = i < n_real ? c[i] : c_hang[i - n_real];
    double distance = cv < -cutoff ? -3. : (cv > cutoff ? 3. : log((1 + cv) / (1 - cv)));
    h = discrete_heaviside(distance * scale);  // :516-521
  } else
    h = st;
  if (i < n_real)
    heaviside[i] = h;
  else
    heaviside_hang[i - n_real] = h;
}

int launch_hanging_mean(const LsHanging &hang, const double *vec, double *out, cudaStream_t s)
{
  if (!hang.n_hang) return 0;
  k_hanging_mean<<<(hang.n_hang + 127) / 128, 128, 0, s>>>(hang, vec, out);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_heaviside(const uint32_t *cell_ls, const uint32_t *last_cell, signed char *cell_state, const double *c, double *heaviside, int N,
                     uint32_t n_cells, uint64_t n_dofs, double scale, const LsHanging &hang, double *c_hang, double *heaviside_hang,
                     cudaStream_t s)
{
  const double cutoff = tanh(2.0);
  const bool ext = hang.cell_ext != nullptr;
  if (ext) AB_CHECK(launch_hanging_mean(hang, c, c_hang, s));  // constraints.distribute: hanging values of the level set
  if (n_cells) {
    k_heaviside_cells<<<(n_cells + 127) / 128, 128, 0, s>>>(ext ? hang.cell_ext : cell_ls, c, c_hang, hang.n_real, ext, N, n_cells, cutoff,
                                                           cell_state);
    AB_COUNT_LAUNCH();
  }
  const uint64_t n = n_dofs + (ext ? hang.n_hang : 0);
  if (n) {
    k_heaviside_dofs<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(last_cell, cell_state, c, c_hang, scale, cutoff, heaviside, heaviside_hang,
                                                                  n_dofs, n);
    AB_COUNT_LAUNCH();
  }
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
