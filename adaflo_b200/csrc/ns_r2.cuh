// ns_r2.cuh -- vmult cell kernel of the 3-D Q2/Q1 Navier-Stokes operator, third design (conforming meshes).
//
// Same cell mapping and arithmetic as ns_r1.cuh (3 lanes per cell, a warp per decade of 10 cells, phases A / B / C
// through a conflict-free shared-memory tile).  profiles/r1_g showed that kernel waiting: 40 % of its stall samples were
// long-scoreboard stalls on global loads (linearisation rows 18 %, node lists + gather 12 %, pressure 4 %, ticket 5 %)
// at 11 warps per SM with every byte of shared memory and every register in use.  This design trades three warps for a
// software pipeline that takes those round trips off the critical path of a warp:
//  * the linearisation rows of a decade (25.9 KB, point-major: lin_index in common.cuh) are staged by the TMA engine
//    (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) into a ring of three 2880-byte slots per warp, one
//    slot per quadrature point of all 10 cells, three points ahead of their use (the first three points of the NEXT
//    decade are requested during the last line of the current one); a lane reads its 12 rows with six 16-byte
//    shared-memory loads: no global LSU wavefronts, no L1 tag look-ups, 24 instead of 72 registers;
//  * the node lists of the next decade are loaded during phase B and its source values are gathered into registers
//    before phase C, so that phase A of the next decade starts from registers;
//  * 8 warps per SM allow 255 registers per thread: nothing is spilled by the pipeline;
//  * the tile is dense (cell stride 243 = 3 mod 16 like 259), the pressure needs 8 slots per cell (the transposed
//    accumulators travel through the cell's own tile entries after phase C has consumed them).
// Hanging nodes and compute_diagonal stay on ns_r1.cuh.
#pragma once

#include "ns_r1.cuh"

namespace ab200 {

constexpr int kR2Warps = 8;
constexpr int kR2CellSmem = 243;
constexpr int kR2POff = kGroup * kR2CellSmem;                 // 2430: 8 pressure slots per cell
constexpr int kR2RingOff = kR2POff + 8 * kGroup;              // 2510 doubles = 20080 bytes (16-byte aligned)
constexpr int kR2Slots = 2;                                   // ring slots of one quadrature point (2880 bytes) each
constexpr int kR2ListOff = kR2RingOff + kR2Slots * kLinPointStride;  // 3230: staged for the next decade, [14][32] uint32: 9 + 4 node lists, geometry row
constexpr int kR2PCurOff = kR2ListOff + 14 * 32 / 2;          // 3454: pressure node list of the current decade, [4][32] uint32
constexpr int kR2GeoOff = kR2PCurOff + 4 * 32 / 2;            // 3518: geometry table rows of the 10 cells, [10][10] doubles
constexpr int kR2BarOff = kR2GeoOff + 10 * kGroup;            // 3618
constexpr int kR2WarpSmem = kR2BarOff + 4;                    // 3622 doubles = 28976 bytes; 8 warps: 231808 bytes (2 barriers, a zero word)

__device__ __forceinline__ uint32_t r2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r2_mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(r2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void r2_mbar_wait(uint64_t *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(r2_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void r2_tma_load(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(r2_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r2_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(r2_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t ld_pinned_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_pinned_src(const double *p)  // source vector: written by other kernels only
{
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned shfl_pinned(unsigned v, int src_lane)
{
  unsigned r;
  asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(src_lane));
  return r;
}
// ticket draw as inline asm: an atomicAdd on a warp-uniform address is turned into a warp-aggregated atomic whose result
// is shuffled (= waited for) at once
__device__ __forceinline__ unsigned ticket_draw(unsigned *counter)
{
  unsigned v;
  asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
  return v;
}
__device__ __forceinline__ void r2_lds2(const double *p, double &a, double &b)
{
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(r2_smem_u32(p)));
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE>
__global__ void __launch_bounds__(kR2Warps * 32, 1) ns_vmult_r2_kernel(const NSArgs A)
{
  AB_TABLES_Q2G3
  constexpr int pairs = MODE == MODE_NEWTON ? 6 : (MODE == MODE_PICARD ? 2 : 0);
  constexpr unsigned kPointBytes = pairs * kLinPairStride * 8;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned n_dec = (unsigned)(A.decade_end - A.decade_begin);
  double *ws = smem + warp * kR2WarpSmem;
  double *ring = ws + kR2RingOff;
  uint64_t *bar = reinterpret_cast<uint64_t *>(ws + kR2BarOff);
  const int cw = lane / 3, t = lane - 3 * cw;
  double *s = ws + cw * kR2CellSmem;
  double *sp = ws + kR2POff + 8 * cw;

  if (lane == 0) bar[2] = 0;  // a zero the compiler does not know (see the ticket shuffle at the end of the decade loop)
  if (pairs && lane == 0) {
    r2_mbar_init(bar, 1);
    r2_mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // two tickets are held (current and next decade), the third is drawn while the current decade is processed
  unsigned i = 0, i_n1 = 0;
  if (lane == 0) {
    i = atomicAdd(A.work_counter, 1u);
    i_n1 = atomicAdd(A.work_counter, 1u);
  }
  i = __shfl_sync(0xffffffffu, i, 0);
  i_n1 = __shfl_sync(0xffffffffu, i_n1, 0);
  if (i >= n_dec) return;

  uint32_t node[9];
  double P[3][9];
  uint32_t *lists = reinterpret_cast<uint32_t *>(ws + kR2ListOff) + lane;  // this lane's column of the staged node lists
  // Node lists of a decade (blocked copies: 120 contiguous bytes per warp-level access) are copied global -> shared
  // memory asynchronously (cp.async: no destination registers that could be spilled while the copy is in flight);
  // every lane reads back only what it copied itself, so its own wait_group is all the synchronisation needed.
  auto stage_lists = [&](const size_t g) {
#pragma unroll
    for (int j = 0; j < 9; ++j)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(r2_smem_u32(lists + 32 * j)), "l"(A.cell_u_blk + (g * 9 + j) * 30 + lane) : "memory");
    if (WITH_P && t < 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(r2_smem_u32(lists + 32 * (9 + j))), "l"(A.cell_p_blk + (g * 4 + j) * 30 + lane) : "memory");
    }
    if (GEOM != ADAFLO_B200_GEOM_GENERAL && t == 0)  // row of the cell in the compressed geometry table
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(r2_smem_u32(lists + 32 * 13)), "l"(A.geom_row + g * kGroup + cw) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t *pcur = reinterpret_cast<uint32_t *>(ws + kR2PCurOff) + lane;
  double *geo = ws + kR2GeoOff + 10 * cw;
  auto read_lists = [&]() {  // the pressure list is used twice, far apart: it stays in shared memory (in registers it was spilled)
#pragma unroll
    for (int j = 0; j < 9; ++j) node[j] = lists[32 * j];
    if (WITH_P && t < 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) pcur[32 * j] = lists[32 * (9 + j)];
    }
  };
  // read_dof_values of the staged decade into registers, raw: the constraint flags are applied when phase A consumes
  // them.  The loads are volatile asm: they stay where they are written.
  auto gather_staged = [&]() {
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const double *p = A.src_u + 3 * (size_t)(lists[32 * j] & ADAFLO_B200_NODE_MASK);
#pragma unroll
      for (int c = 0; c < 3; ++c) P[c][j] = ld_pinned_src(p + c);
    }
    if (WITH_P && t < 2) {  // the 8 vertex values of the cell go straight into their shared-memory slots (vertex j = 2 iy + ix of plane iz = t)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(r2_smem_u32(sp + t + 2 * j)),
                     "l"(A.src_p + (lists[32 * (9 + j)] & ADAFLO_B200_NODE_MASK))
                     : "memory");
    }
    if (GEOM != ADAFLO_B200_GEOM_GENERAL && t == 0) {  // table row of the cell -> geo (two dependent global loads off the critical path)
      constexpr int n16 = GEOM == ADAFLO_B200_GEOM_CARTESIAN ? 2 : 5;
      const double *row = A.geom_table + (size_t)(2 * n16) * lists[32 * 13];
#pragma unroll
      for (int k = 0; k < n16; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(r2_smem_u32(geo + 2 * k)), "l"(row + 2 * k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  size_t g = A.decade_begin + i;
  {
    const bool act0 = (lane < 3 * kGroup) && (g * kGroup + cw < A.n_cells);
    if (act0) {
      stage_lists(g);
      if (pairs && lane == 0) {
#pragma unroll
        for (int q = 0; q < kR2Slots; ++q) r2_tma_load(ring + q * kLinPointStride, A.lin + g * kLinDecade + q * kLinPointStride, kPointBytes, bar + q);
      }
      gather_staged();
    }
  }

  for (unsigned d = 0;; ++d) {
    const bool has_next = i_n1 < n_dec;
    const size_t g1 = A.decade_begin + i_n1;
    unsigned i_n2 = 0;
    const size_t cell_l = g * kGroup + cw;
    const bool active = (lane < 3 * kGroup) && (cell_l < A.n_cells);
    const bool act1 = has_next && (lane < 3 * kGroup) && (g1 * kGroup + cw < A.n_cells);
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (active) {
      const uint32_t cell = (uint32_t)cell_l;
      read_lists();  // staged while the previous decade was processed
      if (lane == 0) {  // remaining points of this decade and the other per-point streams: into the L2 now
        if (MODE == MODE_NEWTON) l2_prefetch(A.lin + g * kLinDecade + kR2Slots * kLinPointStride, (9 - kR2Slots) * kLinPointStride * 8);
        if (MODE == MODE_PICARD) {
#pragma unroll
          for (int n = kR2Slots; n < 9; ++n) l2_prefetch(A.lin + g * kLinDecade + n * kLinPointStride, kPointBytes);
        }
        if (VAR) {
          if (MODE != MODE_STOKES) l2_prefetch(A.rho + g * 270, 270 * 8);
          if (MODE != MODE_STOKES) l2_prefetch(A.damp + g * 270, 270 * 8);
          l2_prefetch(A.mu + g * 270, 270 * 8);
        }
        if (GEOM == ADAFLO_B200_GEOM_GENERAL) {
          l2_prefetch(A.inv_jac + g * (9 * 270), 9 * 270 * 8);
          l2_prefetch(A.jxw + g * 270, 270 * 8);
        }
      }
      Geometry<GEOM> G;
      if (GEOM == ADAFLO_B200_GEOM_GENERAL) G.init_blk(A, cell);
      double p_dg0 = 0;
      uint32_t dg0_node = 0;
      if (WITH_P && A.np_cell == 9) {
        dg0_node = __ldg(A.cell_p + (size_t)cell * 9 + 8);
        p_dg0 = A.src_p[dg0_node & ADAFLO_B200_NODE_MASK];
      }

      // ---------------------------------------------------------------- phase A (from registers)
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const uint32_t flags = node[j] >> ADAFLO_B200_FLAG_SHIFT;
#pragma unroll
        for (int c = 0; c < 3; ++c) P[c][j] = ((flags >> c) & 1u) ? 0.0 : P[c][j];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double Y[3][3], Yd[3][3];  // [qy][ix]: y contraction of the node plane (P index = 3 iy + ix)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          q2::val(P[c][k], P[c][3 + k], P[c][6 + k], Y[0][k], Y[1][k], Y[2][k]);
          q2::der(P[c][k], P[c][3 + k], P[c][6 + k], Yd[0][k], Yd[1][k], Yd[2][k]);
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
      asm volatile("cp.async.wait_all;" ::: "memory");  // pressure values and geometry row staged by this lane
      if (WITH_P && t < 2) {  // the 8 vertex values of the cell were copied into sp by this lane: constrained entries read as zero
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (pcur[32 * j] >> ADAFLO_B200_FLAG_SHIFT) sp[t + 2 * j] = 0.0;
      }
      // node lists of the next decade: in flight during phase B (this decade's are in registers now)
      if (act1) stage_lists(g1);
      __syncwarp(mask);
      if (GEOM != ADAFLO_B200_GEOM_GENERAL) {  // geometry of the cell from its staged table row
        if (GEOM == ADAFLO_B200_GEOM_CARTESIAN) {
          G.K[0][0] = geo[0]; G.K[1][1] = geo[1]; G.K[2][2] = geo[2]; G.det = geo[3];
        } else {
#pragma unroll
          for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int e = 0; e < 3; ++e) G.K[f][e] = geo[3 * f + e];
          G.det = geo[9];
        }
        G.cell_ = cell;
      }
      double py[2][2] = {{0, 0}, {0, 0}}, pacc[2][2] = {{0, 0}, {0, 0}};  // [ix][iz]; pacc: transposed accumulators
      const double npt0 = t == 0 ? NP[0][0] : (t == 1 ? 0.5 : NP[2][0]), npt1 = 1.0 - npt0;  // NP[t][.]
      if (WITH_P) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          py[0][k] = npt0 * sp[k] + npt1 * sp[k + 4];
          py[1][k] = npt0 * sp[k + 2] + npt1 * sp[k + 6];
        }
      }

      // ---------------------------------------------------------------- phase B
      const double wt = (t == 1) ? W[1] : W[0];  // Gauss weight of qy = t
      double dg0_sum = 0;
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const double wj = (j == 1) ? W[1] : W[0];
        double val[3][3], dx[3][3], dy[3][3], dz[3][3];  // [c][qz]
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
          if (MODE != MODE_STOKES) q2::val(T[0], T[1], T[2], val[c][0], val[c][1], val[c][2]);
          q2::der(T[0], T[1], T[2], dz[c][0], dz[c][1], dz[c][2]);
          q2::val(Tx[0], Tx[1], Tx[2], dx[c][0], dx[c][1], dx[c][2]);
          q2::val(Ty[0], Ty[1], Ty[2], dy[c][0], dy[c][1], dy[c][2]);
        }
        double pv[3] = {0, 0, 0};
        if (WITH_P) {
          const double npj0 = j == 0 ? NP[0][0] : (j == 1 ? 0.5 : NP[2][0]), npj1 = 1.0 - npj0;  // NP[j][.]
          const double p0 = npj0 * py[0][0] + npj1 * py[1][0], p1 = npj0 * py[0][1] + npj1 * py[1][1];
#pragma unroll
          for (int q = 0; q < 3; ++q) pv[q] = NP[q][0] * p0 + NP[q][1] * p1 + p_dg0;
        }
        double tp[3];
        const size_t cbase = (size_t)(3 * j) * 30 + lane;
        const unsigned q0 = 9u * d + 3u * (unsigned)j;  // running point count of this warp: point q lives in ring slot q % 2
#pragma unroll
        for (int qz = 0; qz < 3; ++qz) {
          const size_t slot = cbase + (size_t)qz * 30;
          double L[12];
          const unsigned rs = (q0 + qz) & 1u;
          if (pairs) {  // rows of point (qx = j, qy = t, qz): staged by the TMA engine
            r2_mbar_wait(bar + rs, ((q0 + qz) >> 1) & 1u);
            const double *lp = ring + rs * kLinPointStride + 2 * lane;
#pragma unroll
            for (int r = 0; r < pairs; ++r) r2_lds2(lp + r * kLinPairStride, L[2 * r], L[2 * r + 1]);
          }
          const double jxw = G.at_blk(A, g, slot, wj * W[qz] * wt);
          double gr[3][3];
#pragma unroll
          for (int dd = 0; dd < 3; ++dd) {
            double r[3];  // reference derivatives in mesh order (kernel axes -> mesh axes, common.cuh)
            r[kAxX] = dx[dd][qz];
            r[kAxY] = dy[dd][qz];
            r[kAxZ] = dz[dd][qz];
            G.to_real(r, gr[dd]);
          }
          const double div = gr[0][0] + gr[1][1] + gr[2][2];
          double conv[3] = {0, 0, 0};
          double Sadd[3][3];
          if (MODE != MODE_STOKES) {
            const double u[3] = {val[0][qz], val[1][qz], val[2][qz]};
            const double rho = VAR ? __ldg(A.rho + g * 270 + slot) : A.rho0;
            const double damping = VAR ? __ldg(A.damp + g * 270 + slot) : A.damp0;
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) conv[dd] = A.gamma * u[dd];
            if (MODE == MODE_NEWTON) {
              double lu[3], lg[3][3];
#pragma unroll
              for (int dd = 0; dd < 3; ++dd) lu[dd] = L[dd];
#pragma unroll
              for (int dd = 0; dd < 3; ++dd)
#pragma unroll
                for (int e = 0; e < 3; ++e) lg[dd][e] = L[3 + 3 * dd + e];
              const double ltr = lg[0][0] + lg[1][1] + lg[2][2];
              if (!TRANSPOSE) {
                const double f1 = A.beta * div, f2 = A.beta * ltr;
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) {
                  double res = f1 * lu[dd] + f2 * u[dd];
#pragma unroll
                  for (int e = 0; e < 3; ++e) res += lu[e] * gr[dd][e] + u[e] * lg[dd][e];
                  conv[dd] += A.tau1 * res;
                }
              } else {
                const double udl = u[0] * lu[0] + u[1] * lu[1] + u[2] * lu[2];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                  double res = A.beta * ltr * u[e];
#pragma unroll
                  for (int dd = 0; dd < 3; ++dd) res += u[dd] * lg[dd][e];
                  conv[e] += A.tau1 * res;
                }
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) {
#pragma unroll
                  for (int e = 0; e < 3; ++e) Sadd[dd][e] = A.tau1 * rho * u[dd] * lu[e];
                  Sadd[dd][dd] += A.tau1 * rho * A.beta * udl;
                }
              }
            } else if (MODE == MODE_PICARD) {
              double lu[3];
#pragma unroll
              for (int dd = 0; dd < 3; ++dd) lu[dd] = L[dd];
              const double ldiv = L[3];
#pragma unroll
              for (int dd = 0; dd < 3; ++dd) {
                double res = A.beta * ldiv * u[dd];
                if (!TRANSPOSE) {
#pragma unroll
                  for (int e = 0; e < 3; ++e) res += lu[e] * gr[dd][e];
                } else {
#pragma unroll
                  for (int e = 0; e < 3; ++e) Sadd[dd][e] = A.tau1 * rho * u[dd] * lu[e];
                }
                conv[dd] += A.tau1 * res;
              }
            }
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) conv[dd] = (conv[dd] * rho - damping * u[dd]) * jxw;
          }
          if (pairs && lane == 0) {
            // the rows of this point are in registers (every lane's: the loads above are one warp instruction):
            // refill the slot with the point two ahead -- of this decade, or the first two of the next decade
            const int n2 = 3 * j + qz + kR2Slots;
            if (n2 < 9)
              r2_tma_load(ring + rs * kLinPointStride, A.lin + g * kLinDecade + n2 * kLinPointStride, kPointBytes, bar + rs);
            else if (has_next)
              r2_tma_load(ring + rs * kLinPointStride, A.lin + g1 * kLinDecade + (n2 - 9) * kLinPointStride, kPointBytes, bar + rs);
          }
          const double tmu = (VAR ? __ldg(A.mu + g * 270 + slot) : A.mu0) * A.tau1;
          double S[3][3];
          S[0][1] = S[1][0] = tmu * (gr[0][1] + gr[1][0]);
          S[0][2] = S[2][0] = tmu * (gr[0][2] + gr[2][0]);
          S[1][2] = S[2][1] = tmu * (gr[1][2] + gr[2][1]);
#pragma unroll
          for (int dd = 0; dd < 3; ++dd) {
            S[dd][dd] = 2. * tmu * gr[dd][dd] + A.tau_gd * div;
            if (WITH_P) S[dd][dd] -= pv[qz];
          }
          if (TRANSPOSE && (MODE == MODE_NEWTON || MODE == MODE_PICARD)) {
#pragma unroll
            for (int dd = 0; dd < 3; ++dd)
#pragma unroll
              for (int e = 0; e < 3; ++e) S[dd][e] += Sadd[dd][e];
          }
          tp[qz] = -div * jxw;
#pragma unroll
          for (int dd = 0; dd < 3; ++dd) {
            double tr[3];
            G.to_ref(S[dd], jxw, tr);
            val[dd][qz] = conv[dd];
            dx[dd][qz] = tr[kAxX];
            dy[dd][qz] = tr[kAxY];
            dz[dd][qz] = tr[kAxZ];
          }
        }
        // z contraction back, overwriting this lane's own tile entries
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double *sc = s + c * kR1CompSmem + 3 * j + kR1QyStride * t;
          double r[3] = {0, 0, 0}, rx[3], ry[3];
          if (MODE != MODE_STOKES) q2::val_t(val[c][0], val[c][1], val[c][2], r[0], r[1], r[2]);
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
        if (WITH_P) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const double pz = NP[0][k] * tp[0] + NP[1][k] * tp[1] + NP[2][k] * tp[2];  // z^T, then x^T accumulated over the lines
            const double npj0 = j == 0 ? NP[0][0] : (j == 1 ? 0.5 : NP[2][0]);
            pacc[0][k] = fma(npj0, pz, pacc[0][k]);
            pacc[1][k] = fma(1.0 - npj0, pz, pacc[1][k]);
          }
          dg0_sum += tp[0] + tp[1] + tp[2];
        }
      }
      // source values of the next decade: gathered into the registers phase A has released, in flight during phase C;
      // the ticket after next is drawn here as well (its latency is covered by phase C)
      __syncwarp(mask);  // the pressure slots have been read by all lanes of the cell
      if (act1) gather_staged();
      if (lane == 0) {
        // ptxas aggregates atomics on a provably warp-uniform address and shuffles (= waits for) the result at once:
        // offset the address by a zero it cannot see through
        unsigned z;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(z) : "r"(r2_smem_u32(bar + 2)) : "memory");
        i_n2 = ticket_draw(A.work_counter + z);
      }
      __syncwarp(mask);

      // ---------------------------------------------------------------- phase C
      // rolled over the components: unrolled, the scheduler hoists all 81 tile loads and spills the pipeline registers
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const double *sc = s + c * kR1CompSmem;
        double Y[3][3], Yd[3][3];  // [qy][ix]
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
        double o[9];  // o[3 iy + ix] = N^T Y + D^T Yd along y
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          q2::val_t(Y[0][k], Y[1][k], Y[2][k], o[k], o[3 + k], o[6 + k]);
          q2::der_t_add(Yd[0][k], Yd[1][k], Yd[2][k], o[k], o[3 + k], o[6 + k]);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const uint32_t nd = node[j];
          if (!((nd >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u)) red_add(A.dst_u + 3 * (size_t)(nd & ADAFLO_B200_NODE_MASK) + c, o[j]);
        }
      }
      if (WITH_P && !A.skip_p_test) {
        __syncwarp(mask);  // the tile has been consumed: its first 12 entries per cell carry the transposed accumulators
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int ix = 0; ix < 2; ++ix) s[t + 3 * (ix + 2 * k)] = pacc[ix][k];
        __syncwarp(mask);
        if (t < 2) {  // y^T over the three lanes' accumulators of plane iz = t
          double o[4];
#pragma unroll
          for (int ix = 0; ix < 2; ++ix) {
            const double a0 = s[0 + 3 * (ix + 2 * t)], a1 = s[1 + 3 * (ix + 2 * t)], a2 = s[2 + 3 * (ix + 2 * t)];
            o[ix] = NP[0][0] * a0 + NP[1][0] * a1 + NP[2][0] * a2;
            o[2 + ix] = NP[0][1] * a0 + NP[1][1] * a1 + NP[2][1] * a2;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t pn = pcur[32 * j];
            if (!(pn >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (pn & ADAFLO_B200_NODE_MASK), o[j]);
          }
        }
        if (A.np_cell == 9) {
          double sum = dg0_sum;
          sum += __shfl_sync(mask, dg0_sum, 3 * cw + (t + 1) % 3) + __shfl_sync(mask, dg0_sum, 3 * cw + (t + 2) % 3);
          if (t == 0 && !(dg0_node >> ADAFLO_B200_FLAG_SHIFT)) red_add(A.dst_p + (dg0_node & ADAFLO_B200_NODE_MASK), sum);
        }
      }
      __syncwarp(mask);  // the tile is free for the next decade
    }
    if (!has_next) break;
    g = g1;
    // The source lane (0) is read from shared memory with a volatile load: the shuffle cannot be scheduled before it, i.e.
    // not right behind the atomic whose latency phase C is meant to cover.
    unsigned zero;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(zero) : "r"(r2_smem_u32(bar + 2)) : "memory");
    i_n1 = shfl_pinned(i_n2, (int)zero);
  }
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE>
int launch_vmult_r2_inst(const NSArgs &a, cudaStream_t s)
{
  auto kern = ns_vmult_r2_kernel<WITH_P, MODE, GEOM, VAR, TRANSPOSE>;
  const size_t smem = sizeof(double) * kR2Warps * kR2WarpSmem;
  static int n_sm[64] = {};
  int dev = 0;
  AB_CUDA(cudaGetDevice(&dev));
  if (dev >= 64) return fail(ADAFLO_B200_ERR_ARG, "ns_vmult_r2: device ordinal out of range");
  if (!n_sm[dev]) {
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AB_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
  }
  if (!a.work_counter) return fail(ADAFLO_B200_ERR_STATE, "ns_vmult_r2: no work counter");
  NSArgs b = a;
  const size_t n_dec = n_groups(a.n_cells);
  if (b.decade_end == 0 || b.decade_end > n_dec) b.decade_end = n_dec;  // default: all decades
  if (b.decade_begin >= b.decade_end) return 0;
  const size_t blocks_needed = (b.decade_end - b.decade_begin + kR2Warps - 1) / kR2Warps;
  const unsigned grid = (unsigned)(blocks_needed < (size_t)n_sm[dev] ? blocks_needed : (size_t)n_sm[dev]);
  AB_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(unsigned), s));
  kern<<<grid, kR2Warps * 32, smem, s>>>(b);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// conforming meshes, no diagonal pass; everything else stays on ns_r1.cuh
template <int GEOM>
int launch_vmult_r2_geom(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  const bool tr = a.transpose && (mode == MODE_NEWTON || mode == MODE_PICARD);  // other modes are symmetric
  if (a.hn_u || a.diag || (tr && !with_p)) return launch_vmult_r1_geom<GEOM>(a, with_p, mode, var, s);
  if (tr) {
    if (mode == MODE_NEWTON && !var) return launch_vmult_r2_inst<true, MODE_NEWTON, GEOM, false, true>(a, s);
    if (mode == MODE_NEWTON && var) return launch_vmult_r2_inst<true, MODE_NEWTON, GEOM, true, true>(a, s);
    if (mode == MODE_PICARD && !var) return launch_vmult_r2_inst<true, MODE_PICARD, GEOM, false, true>(a, s);
    return launch_vmult_r2_inst<true, MODE_PICARD, GEOM, true, true>(a, s);
  }
#define AB_CASE(P, M, V) \
  if (with_p == P && mode == M && var == V) return launch_vmult_r2_inst<P, M, GEOM, V, false>(a, s);
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
  return fail(ADAFLO_B200_ERR_ARG, "launch_vmult_r2_geom: bad variant");
}

}  // namespace ab200
