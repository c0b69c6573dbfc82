// ns_r1.cuh -- vmult cell kernel of the 3-D Q2/Q1 Navier-Stokes operator, second design.
//
// Profiling of the first design (ns_q2q1.cuh: 9 lanes per cell, line ownership,
// profiles/r1_a_*) showed the LSU data pipe at 85 % (two thirds of it shared-memory
// transposes with bank conflicts, one third uncoalesced per-point loads) while HBM sat
// at 28 %.  This design halves the shared-memory traffic and makes every per-point
// stream fully coalesced:
//
//  * 3 lanes own one cell, a warp owns a "decade" of 10 consecutive cells (30 lanes).
//  * phase A: lane t owns the node plane iz = t (9 nodes x 3 components in registers) and
//    runs both in-plane contractions (x and y) in registers; it publishes three tiles
//    per component (interpolated, d/dx, d/dy) to shared memory.
//  * phase B: lane t owns the three z-lines (qx = j, qy = t), j = 0..2; per line it reads
//    the 3 x 9 plane values, contracts in z (values + 3 derivatives at 3 points), applies
//    the physics and contracts back in z, overwriting the same tile entries.
//  * phase C: lane t owns node plane t again, transposed in-plane contractions, scatter.
//    => one shared-memory round trip per direction pair instead of two.
//  * tile entry (array, component, qx, qy, k) at 9 arr + 27 c + k + 3 qx + 81 qy, cell stride 259 doubles: 81 = 1 and
//    259 = 3 (mod 16) make the 8-byte bank of every access 3 cell + t in both ownership patterns (conflict free,
//    243 of 259 slots used); the pressure travels as its 8 vertex values / 12 transposed accumulators.
//  * all per-quadrature-point data (linearisation point, rho/mu/damping, Jacobians) and the cell->node lists are
//    stored decade-blocked, [decade][row][3 qx + qz][30 lanes], so each warp-level load is contiguous (coef_index in
//    common.cuh); the 12 linearisation rows are interleaved in pairs and fetched with 16-byte loads (lin_index).
//  * the kernel is persistent: one block of 10-11 independent warps per SM, every warp draws decades from a global
//    ticket counter; a launch can be restricted to a decade range (host-vector pipeline, halo overlap).
// Measured history, rejected variants and the ncu evidence: DESIGN.md section 3.
#pragma once

#include "hanging.cuh"
#include "ns_q2q1.cuh"

namespace ab200 {

// Shared-memory layout of one cell: entry (arr, c, qx, qy, k) of the velocity tiles at
//   9 arr + 27 c + k + 3 qx + 81 qy            (arr: value / d/dx / d/dy part, c: component, k: z index)
// and 12 slots for the pressure behind them.  81 = 1 and 259 = 3 (mod 16): the lanes (cell, t) of a warp hit the
// 8-byte bank 3 cell + t both when t is the y index (phase B) and when it is k (phases A, C) -> no bank conflicts.
#ifndef AB_R1_MERGE
#define AB_R1_MERGE 0
#endif
#ifndef AB_R1_EVF   // linearisation rows are read once: mark their L2 lines evict-first when they are consumed (A/B: no gain)
#define AB_R1_EVF 0
#endif
#ifndef AB_R1_PFLANE  // L2 prefetch by per-lane prefetch instructions instead of one bulk request of the TMA engine
#define AB_R1_PFLANE 0
#endif
#ifndef AB_R1_IDXEV  // cell->node lists: read once, same policy
#define AB_R1_IDXEV 0
#endif
#ifndef AB_R1_PFEV   // the L2 prefetch itself carries the evict-first policy
#define AB_R1_PFEV 0
#endif
#ifndef AB_R1_RING   // register ring for the linearisation rows (see phase B)
#define AB_R1_RING 0
#endif
#ifndef AB_R1_UNROLLJ  // line loop of phase B unrolled: no loop-carried register constraints (no spills at 168 registers)
#define AB_R1_UNROLLJ 0
#endif
#ifndef AB_R1_SCHED  // 1: tickets handed out by a scheduler warp (see the kernel); 0: every worker draws its own
#define AB_R1_SCHED 0
#endif
#ifndef AB_R1_PLAIN_TICKET  // ticket atomic that ptxas cannot aggregate (see the kernel)
#define AB_R1_PLAIN_TICKET 1
#endif
#ifndef AB_R1_PASYNC  // pressure vertex values by cp.async into their shared-memory slots (see phase A)
#define AB_R1_PASYNC 1
#endif
#ifndef AB_R1_LEAN   // lean per-point physics on Cartesian cells and predicated scatter (round 2: fewer instructions per cell)
#define AB_R1_LEAN 1
#endif
#ifndef AB_R1_PFSRC  // paced L2 prefetch of src_u this many decades ahead of the ticket (0: off)
#define AB_R1_PFSRC 0
#endif
#ifndef AB_R1_RING0  // rows of the first point of a line requested one line ahead (24 registers stay live)
#define AB_R1_RING0 0
#endif
#ifndef AB_R1_STAGE  // cell->node lists of the warp's next decade staged through dead tile slots (see phase A)
#define AB_R1_STAGE 0
#endif
#ifndef AB_R1_WARPS
#define AB_R1_WARPS 11
#endif
constexpr int kR1Warps = AB_R1_WARPS;                // warps per block: one persistent block per SM, 11 x 20.7 KB of shared memory
constexpr int kR1CompSmem = 27;
constexpr int kR1QyStride = 81;
constexpr int kR1PSmem = 243;               // offset of the pressure slots
constexpr int kR1CellSmem = 259;
constexpr int kR1WarpSmem = kGroup * kR1CellSmem;
constexpr int kR1MaxReg = 168;              // 3 warps per SM sub-partition (16384 registers each)

__device__ __forceinline__ int r1_idx(int qx, int qy, int k) { return k + 3 * qx + kR1QyStride * qy; }

// bulk L2 prefetch of a contiguous block (TMA engine, no registers, no LSU wavefronts)
__device__ __forceinline__ void l2_prefetch(const void *ptr, unsigned bytes)
{
#if AB_R1_PFEV
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(ptr), "r"(bytes), "l"(pol) : "memory");
#else
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
#endif
}

// ---------------------------------------------------------------------------
// Structured 1-D kernels for the Q2 Lagrange basis on {0, 1/2, 1} at the 3 Gauss points.
// The tables have the rows N[1] = (0, 1, 0), D[1] = (-1, 0, 1) and the symmetries
// N[2][i] = N[0][2-i], D[2][i] = -D[0][2-i]; written out, a 3x3 apply costs 5-6 flops
// instead of 9 (IEEE rules forbid the compiler to drop the multiplications by 0 and 1).
namespace q2 {
constexpr double a = 0.6872983346207416, b = 0.4, c = -0.08729833462074169;            // N[0][:]
constexpr double d0 = -2.549193338482967, d1 = 3.0983866769659336, d2 = -0.5491933384829668;  // D[0][:]
// o = N x
__device__ __forceinline__ void val(const double x0, const double x1, const double x2, double &o0, double &o1, double &o2)
{
  const double t = b * x1;
  o0 = fma(a, x0, fma(c, x2, t));
  o1 = x1;
  o2 = fma(c, x0, fma(a, x2, t));
}
// o = D x
__device__ __forceinline__ void der(const double x0, const double x1, const double x2, double &o0, double &o1, double &o2)
{
  const double t = d1 * x1;
  o0 = fma(d0, x0, fma(d2, x2, t));
  o1 = x2 - x0;
  o2 = fma(-d2, x0, fma(-d0, x2, -t));
}
// o = N^T i
__device__ __forceinline__ void val_t(const double i0, const double i1, const double i2, double &o0, double &o1, double &o2)
{
  o0 = fma(a, i0, c * i2);
  o1 = fma(b, i0 + i2, i1);
  o2 = fma(c, i0, a * i2);
}
// o += D^T i
__device__ __forceinline__ void der_t_add(const double i0, const double i1, const double i2, double &o0, double &o1, double &o2)
{
  o0 += fma(d0, i0, fma(-d2, i2, -i1));
  o1 = fma(d1, i0 - i2, o1);
  o2 += fma(d2, i0, fma(-d0, i2, i1));
}
}  // namespace q2

// pinned (volatile) read-only load: keeps the software prefetch where it is written
__device__ __forceinline__ double ld_pinned(const double *p)
{
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// L2 eviction policy "evict first" for data that is read exactly once per launch (the per-point streams, the cell->node
// lists): once consumed, their lines are the first to go, so the lines the TMA prefetch has brought in for the decades
// in flight and the node planes of the vectors (reused by the next cell layer) survive.  Measured: -12 % kernel time.
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// pinned 16-byte load of a row pair of the linearisation array
__device__ __forceinline__ void ld_pinned2(const double *p, double &a, double &b, const uint64_t pol)
{
#if AB_R1_EVF
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(a), "=d"(b) : "l"(p), "l"(pol));
#else
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
#endif
}

// read-once 4-byte load (cell->node lists)
__device__ __forceinline__ uint32_t ld_once(const uint32_t *p, const uint64_t pol)
{
#if AB_R1_IDXEV
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
#else
  return __ldg(p);
#endif
}

// rows of the linearisation point of quadrature point n = 3 qx + qz of this lane (row pairs, see lin_index)
template <int MODE>
__device__ __forceinline__ void load_lin(double (&L)[12], const double *lin_lane, int n, const uint64_t pol)
{
  constexpr int pairs = MODE == MODE_NEWTON ? 6 : (MODE == MODE_PICARD ? 2 : 0);
#pragma unroll
  for (int r = 0; r < pairs; ++r) ld_pinned2(lin_lane + r * kLinPairStride + n * kLinPointStride, L[2 * r], L[2 * r + 1], pol);
}

// refinement mask (hanging.cuh: three 3-bit groups indexed by mesh direction) re-indexed by kernel axis; the rotation
// of the axes is cyclic, so HnInfo::line (which uses the cyclic successors of a direction) stays valid
__device__ __forceinline__ uint32_t hn_to_kernel_axes(uint32_t m)
{
  if (kAxX == 0) return m;
  uint32_t r = 0;
#pragma unroll
  for (int g = 0; g < 9; g += 3) {
    const uint32_t grp = (m >> g) & 7u;
    r |= (((grp >> kAxX) & 1u) | (((grp >> kAxY) & 1u) << 1) | (((grp >> kAxZ) & 1u) << 2)) << g;
  }
  return r;
}

template <int MODE, int GEOM, bool VAR>
__device__ __forceinline__ void r1_prefetch(const NSArgs &A, const size_t g)
{
  if (MODE == MODE_NEWTON) l2_prefetch(A.lin + g * kLinDecade, kLinDecade * 8);
  if (MODE == MODE_PICARD) {  // rows 0-3 = the first two row pairs of every point
#pragma unroll
    for (int n = 0; n < 9; ++n) l2_prefetch(A.lin + g * kLinDecade + n * kLinPointStride, 2 * kLinPairStride * 8);
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

// One decade (10 cells) of the cell loop, executed by one warp: 3 lanes per cell.  x, y, z, ix, qx ... below are
// KERNEL axes (common.cuh: kernel (X, Y, Z) = mesh (y, z, x)); only the geometry sees mesh axes.
// idx_staged: the cell->node lists of this decade were staged into shared memory by the previous decade of this warp;
// g_next: the warp's next decade (kNoDecade: none), whose lists are staged during phase C (AB_R1_STAGE)
constexpr size_t kNoDecade = ~(size_t)0;
template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE, bool HANGING>
__device__ __forceinline__ void r1_decade(const NSArgs &A, const size_t g, double *warp_smem, const int lane, const bool idx_staged = false,
                                          const size_t g_next = kNoDecade)
{
  AB_TABLES_Q2G3
  const int cw = lane / 3;
  const int t = lane - 3 * cw;
  const size_t cell_l = g * kGroup + cw;
  const bool active = (lane < 3 * kGroup) && (cell_l < A.n_cells);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t cell = (uint32_t)cell_l;
  double *s = warp_smem + cw * kR1CellSmem;
  double *sp = s + kR1PSmem;

  // The per-point streams of this decade are consumed in phase B; ask the L2 for them now (TMA engine) so that their
  // HBM latency overlaps the gather and the in-plane contractions of phase A.  (Asking for the warp's next decade
  // instead -- at the start of phase A, B or C -- was measured 5-30 % slower: the data is evicted before use.)
#if AB_R1_PFLANE
  if (MODE == MODE_NEWTON) {
    const char *blk = reinterpret_cast<const char *>(A.lin + g * kLinDecade);
#pragma unroll
    for (int i = 0; i < 7; ++i)
      if (lane + 30 * i < 203) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + (size_t)(lane + 30 * i) * 128));  // 30 active lanes
  } else if (lane == 0)
    r1_prefetch<MODE, GEOM, VAR>(A, g);
#else
  if (lane == 0) r1_prefetch<MODE, GEOM, VAR>(A, g);
#endif

  const double *lin_lane = A.lin + g * kLinDecade + 2 * lane;
  Geometry<GEOM> G;  // table row of the cell (two dependent loads): issue before the gather
  G.init_blk(A, cell);
  const HnInfo hn(hn_to_kernel_axes((HANGING && A.hn_u) ? __ldg(A.hn_u + cell) : 0u));
  const HnInfo hnp(hn_to_kernel_axes((HANGING && WITH_P && A.hn_p) ? __ldg(A.hn_p + cell) : 0u));
  const bool hn_z_warp = HANGING && __any_sync(mask, hn.any() && (hn.face(0) || hn.face(1) || hn.edge(2)));
  const bool hnp_z_warp = HANGING && WITH_P && __any_sync(mask, hnp.any() && (hnp.face(0) || hnp.face(1) || hnp.edge(2)));

  // ---------------------------------------------------------------- phase A
  uint32_t node[9];
  uint64_t pol = policy_evict_first();  // re-created where it is needed: it must not occupy registers across the phases
#if AB_R1_STAGE
  // Staged lists: lane (cell, t) keeps the 9 + 4 words of its NEXT decade in tile slots that are private to it once it
  // has read them in phase C (entries with k = t of component 0: value array for the velocity nodes, d/dx array for the
  // pressure vertices).  The list loads -- first touches, i.e. HBM latency, 10 % of the stall samples of profiles/r2_i --
  // then overlap two thirds of phase C instead of opening the decade.
  constexpr bool kStage = !(HANGING && !WITH_P);
  const uint32_t *stage_u = reinterpret_cast<const uint32_t *>(s + t);
  if (kStage && idx_staged) {
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 9; ++j) node[j] = stage_u[2 * (3 * (j % 3) + kR1QyStride * (j / 3))];
  } else
#endif
#pragma unroll
  for (int j = 0; j < 9; ++j) node[j] = ld_once(A.cell_u_blk + (g * 9 + j) * 30 + lane, pol);
#if AB_R1_MERGE
  // x-neighbours of a decade share a node plane: lane (cw + 1, t = 0) and lane (cw, t = 2) then scatter to the same
  // nodes.  The upper lane hands its contribution to the lower one (one shuffle), which issues a single atomic.
  unsigned merge_recv = 0, merge_send = 0;  // bit j: this lane receives from lane + 1 / gives to lane - 1
  if (!HANGING) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const uint32_t up = __shfl_down_sync(mask, node[j], 1), dn = __shfl_up_sync(mask, node[j], 1);
      if (t == 2 && lane + 1 < 3 * kGroup && ((mask >> (lane + 1)) & 1u) && up == node[j]) merge_recv |= 1u << j;
      if (t == 0 && lane > 0 && dn == node[j]) merge_send |= 1u << j;
    }
  }
#endif
#if AB_R1_LEAN
  // does any lane of the decade see a constrained velocity entry?  (interior decades: no -> no selects, no predicates)
  const bool any_con =
    __any_sync(mask, ((node[0] | node[1] | node[2] | node[3] | node[4] | node[5] | node[6] | node[7] | node[8]) >> ADAFLO_B200_FLAG_SHIFT) != 0u);
#endif
  uint32_t pnode[4] = {0, 0, 0, 0};
  double p_dg0 = 0;
  uint32_t dg0_node = 0;
  if (WITH_P) {
    if (t < 2) {
#if AB_R1_STAGE
      if (kStage && idx_staged) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pnode[j] = stage_u[2 * (9 + 3 * j)];
      } else
#endif
#pragma unroll
      for (int j = 0; j < 4; ++j) pnode[j] = ld_once(A.cell_p_blk + (g * 4 + j) * 30 + lane, pol);
#if AB_R1_PASYNC
      // the 8 vertex values of the cell go straight into their shared-memory slots (cp.async: no registers, nothing waits
      // for them before the end of phase A); 4 % of the stall samples of profiles/r1_g sat on this gather
      if (!HANGING) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sp + t + 2 * j)),
                       "l"(A.src_p + (pnode[j] & ADAFLO_B200_NODE_MASK))
                       : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
#endif
    }
    if (A.np_cell == 9) {
      dg0_node = __ldg(A.cell_p + (size_t)cell * 9 + 8);
      p_dg0 = A.src_p[dg0_node & ADAFLO_B200_NODE_MASK];
    }
  }
  // Diagonal mode (compute_diagonal, velocity block only, instantiated with the hanging-node variant):
  // one pass per cell-local slot with a unit source at that slot; only that slot's result is scattered,
  // which yields diag(C^T A_cell C) including the hanging-node weights.
  double dg0_sum = 0;
  constexpr bool kDiagCapable = HANGING && !WITH_P;
  const int n_pass = (kDiagCapable && A.diag) ? 81 : 1;
#pragma unroll 1
  for (int pass = 0; pass < n_pass; ++pass) {
  const int dg_c = pass / 27, dg_t = (pass % 27) / 9, dg_j = pass % 9;
  {
    double P[3][9];
#if AB_R1_LEAN
    if (!kDiagCapable) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const size_t n = node[j] & ADAFLO_B200_NODE_MASK;
#pragma unroll
        for (int c = 0; c < 3; ++c) P[c][j] = A.src_u[3 * n + c];
      }
      if (any_con) {  // warp-uniform: interior decades skip the 27 selects
#pragma unroll
        for (int j = 0; j < 9; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if ((node[j] >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u) P[c][j] = 0.0;
      }
    } else
#endif
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const size_t n = node[j] & ADAFLO_B200_NODE_MASK;
      const uint32_t flags = node[j] >> ADAFLO_B200_FLAG_SHIFT;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double v;
        if (kDiagCapable && A.diag)
          v = (c == dg_c && j == dg_j && t == dg_t) ? 1.0 : 0.0;
        else
          v = A.src_u[3 * n + c];
        P[c][j] = ((flags >> c) & 1u) ? 0.0 : v;
      }
    }
    if (HANGING) {  // read_dof_values: interpolate the parent's face / edge values onto this cell's nodes
      if (hn.any()) {
        // all lines of one direction before the next direction (single lines do not commute)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (hn.line(0, i, t, 3)) {  // x line at iy = i
#pragma unroll
            for (int c = 0; c < 3; ++c) hn_q2(hn.sub(0), P[c][3 * i], P[c][3 * i + 1], P[c][3 * i + 2]);
          }
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (hn.line(1, t, i, 3)) {  // y line at ix = i
#pragma unroll
            for (int c = 0; c < 3; ++c) hn_q2(hn.sub(1), P[c][i], P[c][3 + i], P[c][6 + i]);
          }
      }
      if (hn_z_warp) {  // z lines live across the 3 lanes of the cell
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
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double Y[3][3], Yd[3][3];  // [qy][ix]: y contraction of the node plane (P index = 3 iy + ix)
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
#if AB_R1_PASYNC
    if (WITH_P && !HANGING) {
      if (t < 2) {
        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (pnode[j] >> ADAFLO_B200_FLAG_SHIFT) sp[t + 2 * j] = 0.0;  // constrained entries read as zero
      }
    } else
#endif
    if (WITH_P) {
      double pl[4] = {0, 0, 0, 0};
      if (t < 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double v = A.src_p[pnode[j] & ADAFLO_B200_NODE_MASK];
          pl[j] = (pnode[j] >> ADAFLO_B200_FLAG_SHIFT) ? 0.0 : v;
        }
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
      if (t < 2) {  // the 8 vertex values of the cell (vertex j = 2 iy + ix of plane iz = t)
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[t + 2 * j] = pl[j];
      }
    }
  }
  __syncwarp(mask);
  // pressure: all three lanes read the 8 vertex values and interpolate in y to their qy = t
  double py[2][2] = {{0, 0}, {0, 0}}, pacc[2][2] = {{0, 0}, {0, 0}};  // [ix][iz]; pacc: transposed accumulators
  const double npt0 = t == 0 ? NP[0][0] : (t == 1 ? 0.5 : NP[2][0]), npt1 = 1.0 - npt0;  // NP[t][.]
  if (WITH_P) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      py[0][k] = npt0 * sp[k] + npt1 * sp[k + 4];
      py[1][k] = npt0 * sp[k + 2] + npt1 * sp[k + 6];
    }
    __syncwarp(mask);  // the slots are reused for the transposed direction
  }

  // ---------------------------------------------------------------- phase B
  const double wt = (t == 1) ? W[1] : W[0];  // Gauss weight of qy = t
  // The line loop stays rolled (a third of the code: instruction cache).  The linearisation rows of line j + 1 are
  // requested as soon as the physics of line j has consumed its rows -- before the z contraction back and the tile
  // writes of line j, the tile reads and the z contraction of line j + 1 -- at no cost in registers (the rows of line
  // j are dead by then); the rows of line 0 are requested before the pressure interpolation.
  double LQ[3][12];
#if AB_R1_LEAN
  const double cA = fma(A.gamma, A.rho0, -A.damp0), cB = A.tau1 * A.rho0, tmu0 = A.mu0 * A.tau1;  // constant coefficients
#endif
  pol = policy_evict_first();
#if AB_R1_RING
#pragma unroll
  for (int qz = 0; qz < 3; ++qz) load_lin<MODE>(LQ[qz], lin_lane, qz, pol);
#endif
#if AB_R1_RING0
  load_lin<MODE>(LQ[0], lin_lane, 0, pol);
#endif
#if AB_R1_UNROLLJ
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int j = 0; j < 3; ++j) {
    const double wj = (j == 1) ? W[1] : W[0];
#if AB_R1_LEAN
    const double wline = G.det * (wj * wt);
#endif
#if AB_R1_RING0
    // rows of point qz = 0 were requested a line ago (see the end of its physics below); qz = 1, 2 now
#pragma unroll
    for (int qz = 1; qz < 3; ++qz) load_lin<MODE>(LQ[qz], lin_lane, 3 * j + qz, pol);
#elif !AB_R1_RING
#pragma unroll
    for (int qz = 0; qz < 3; ++qz) load_lin<MODE>(LQ[qz], lin_lane, 3 * j + qz, pol);
#endif
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
    // points (qx = j, qy = t, qz): coefficient slot 3 j + qz of this lane
    const size_t cbase = (size_t)(3 * j) * 30 + lane;
#pragma unroll
    for (int qz = 0; qz < 3; ++qz) {
      const size_t slot = cbase + (size_t)qz * 30;
      const double(&L)[12] = LQ[qz];
#if AB_R1_LEAN
      if constexpr (GEOM == ADAFLO_B200_GEOM_CARTESIAN && !TRANSPOSE) {
        // Lean formulation of the per-point physics for axis-aligned cells (same terms as the general branch below,
        // navier_stokes_matrix.cc:702-893, with the constant factors folded): the kernel is bound by its instruction
        // count, and the general branch spends 110 fp64 instructions per point where 71 suffice:
        //   value part   w [(gamma rho - damping) u + tau1 rho res]      = wA u + wB res
        //   flux part    K_ee w S_de with S = tmu (grad + grad^T) + (tau_gd div - p) I
        const double w = wline * W[qz];  // JxW
        const double g0 = G.K[0][0] * w, g1 = G.K[1][1] * w, g2 = G.K[2][2] * w;
        double gr[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          gr[d][kAxX] = dx[d][qz] * G.K[kAxX][kAxX];
          gr[d][kAxY] = dy[d][qz] * G.K[kAxY][kAxY];
          gr[d][kAxZ] = dz[d][qz] * G.K[kAxZ][kAxZ];
        }
        const double div = gr[0][0] + gr[1][1] + gr[2][2];
        if (MODE != MODE_STOKES) {
          const double u[3] = {val[0][qz], val[1][qz], val[2][qz]};
          double wA, wB;
          if (VAR) {
            const double rho = __ldg(A.rho + g * 270 + slot), damping = __ldg(A.damp + g * 270 + slot);
            wA = fma(A.gamma, rho, -damping) * w;
            wB = (A.tau1 * rho) * w;
          } else {
            wA = cA * w;
            wB = cB * w;
          }
          if (MODE == MODE_NEWTON) {
            const double ltr = L[3] + L[7] + L[11];
            const double f1 = A.beta * div, f2 = A.beta * ltr;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double res = f1 * L[d];
              res = fma(f2, u[d], res);
#pragma unroll
              for (int e = 0; e < 3; ++e) res = fma(L[e], gr[d][e], fma(u[e], L[3 + 3 * d + e], res));
              val[d][qz] = fma(wB, res, wA * u[d]);
            }
          } else if (MODE == MODE_PICARD) {
            const double f2 = A.beta * L[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double res = f2 * u[d];
#pragma unroll
              for (int e = 0; e < 3; ++e) res = fma(L[e], gr[d][e], res);
              val[d][qz] = fma(wB, res, wA * u[d]);
            }
          } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) val[d][qz] = wA * u[d];
          }
        }
        const double tmu = VAR ? __ldg(A.mu + g * 270 + slot) * A.tau1 : tmu0;
        const double tmu2 = tmu + tmu;
        const double q = WITH_P ? fma(A.tau_gd, div, -pv[qz]) : A.tau_gd * div;
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
#if AB_R1_RING0
        // the rows of this point are consumed: request the same point of the next line now, a whole line ahead of its
        // use (the first consumer of a line's rows otherwise waits for them: 23 % of the stall samples of profiles/r2_i)
        if (qz == 0 && j < 2) load_lin<MODE>(LQ[0], lin_lane, 3 * (j + 1), pol);
#endif
        continue;
      }
#endif
      const double jxw = G.at_blk(A, g, slot, wj * W[qz] * wt);
      double gr[3][3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double r[3];  // reference derivatives in mesh order (kernel axes -> mesh axes, common.cuh)
        r[kAxX] = dx[d][qz];
        r[kAxY] = dy[d][qz];
        r[kAxZ] = dz[d][qz];
        G.to_real(r, gr[d]);
      }
      const double div = gr[0][0] + gr[1][1] + gr[2][2];
      double conv[3] = {0, 0, 0};
      double Sadd[3][3];
      if (MODE != MODE_STOKES) {
        const double u[3] = {val[0][qz], val[1][qz], val[2][qz]};
        const double rho = VAR ? __ldg(A.rho + g * 270 + slot) : A.rho0;
        const double damping = VAR ? __ldg(A.damp + g * 270 + slot) : A.damp0;
#pragma unroll
        for (int d = 0; d < 3; ++d) conv[d] = A.gamma * u[d];
        if (MODE == MODE_NEWTON) {
          double lu[3], lg[3][3];
#pragma unroll
          for (int d = 0; d < 3; ++d) lu[d] = L[d];
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int e = 0; e < 3; ++e) lg[d][e] = L[3 + 3 * d + e];
          const double ltr = lg[0][0] + lg[1][1] + lg[2][2];
          if (!TRANSPOSE) {
            const double f1 = A.beta * div, f2 = A.beta * ltr;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double res = f1 * lu[d] + f2 * u[d];
#pragma unroll
              for (int e = 0; e < 3; ++e) res += lu[e] * gr[d][e] + u[e] * lg[d][e];
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
          for (int d = 0; d < 3; ++d) lu[d] = L[d];
          const double ldiv = L[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            double res = A.beta * ldiv * u[d];
            if (!TRANSPOSE) {
#pragma unroll
              for (int e = 0; e < 3; ++e) res += lu[e] * gr[d][e];
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
      const double tmu = (VAR ? __ldg(A.mu + g * 270 + slot) : A.mu0) * A.tau1;
      double S[3][3];
      S[0][1] = S[1][0] = tmu * (gr[0][1] + gr[1][0]);
      S[0][2] = S[2][0] = tmu * (gr[0][2] + gr[2][0]);
      S[1][2] = S[2][1] = tmu * (gr[1][2] + gr[2][1]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        S[d][d] = 2. * tmu * gr[d][d] + A.tau_gd * div;
        if (WITH_P) S[d][d] -= pv[qz];
      }
      if (TRANSPOSE && (MODE == MODE_NEWTON || MODE == MODE_PICARD)) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int e = 0; e < 3; ++e) S[d][e] += Sadd[d][e];
      }
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
#if AB_R1_RING0
      if (qz == 0 && j < 2) load_lin<MODE>(LQ[0], lin_lane, 3 * (j + 1), pol);
#endif
    }
#if AB_R1_RING
    if (j < 2) {  // all rows of this line are consumed: request the next line's now
#pragma unroll
      for (int qz = 0; qz < 3; ++qz) load_lin<MODE>(LQ[qz], lin_lane, 3 * (j + 1) + qz, pol);
    }
#endif
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
  if (WITH_P) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) sp[t + 3 * (ix + 2 * k)] = pacc[ix][k];
  }
  __syncwarp(mask);

  // ---------------------------------------------------------------- phase C
#pragma unroll
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
      // x contraction back: Y[qy][ix] = N^T T + D^T Tx, Yd[qy][ix] = N^T Ty
      q2::val_t(T[0], T[1], T[2], Y[qy][0], Y[qy][1], Y[qy][2]);
      q2::der_t_add(Tx[0], Tx[1], Tx[2], Y[qy][0], Y[qy][1], Y[qy][2]);
      q2::val_t(Ty[0], Ty[1], Ty[2], Yd[qy][0], Yd[qy][1], Yd[qy][2]);
    }
#if AB_R1_STAGE
    if (kStage && c == 0 && g_next != kNoDecade && g_next * kGroup + cw < A.n_cells) {
      // this lane's entries of component 0 are dead now: stage the lists of its next decade into them
      const unsigned dst0 = (unsigned)__cvta_generic_to_shared(s + t);
#pragma unroll
      for (int j = 0; j < 9; ++j)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 8u * (3 * (j % 3) + kR1QyStride * (j / 3))),
                     "l"(A.cell_u_blk + (g_next * 9 + j) * 30 + lane)
                     : "memory");
      if (WITH_P && t < 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 8u * (9 + 3 * j)), "l"(A.cell_p_blk + (g_next * 4 + j) * 30 + lane)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#endif
    double o[9];  // o[3 iy + ix] = N^T Y + D^T Yd along y
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      q2::val_t(Y[0][i], Y[1][i], Y[2][i], o[i], o[3 + i], o[6 + i]);
      q2::der_t_add(Yd[0][i], Yd[1][i], Yd[2][i], o[i], o[3 + i], o[6 + i]);
    }
    if (HANGING) {  // distribute_local_to_global: transposed interpolation onto the parent's nodes
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
#if AB_R1_LEAN
    if (!kDiagCapable && !any_con && !AB_R1_MERGE) {
#pragma unroll
      for (int j = 0; j < 9; ++j) red_add(A.dst_u + 3 * (size_t)(node[j] & ADAFLO_B200_NODE_MASK) + c, o[j]);
      continue;
    }
#endif
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const uint32_t nd = node[j];
      if (kDiagCapable && A.diag && !(c == dg_c && j == dg_j && t == dg_t)) continue;
#if AB_R1_MERGE
      if (!HANGING) {
        const double from_up = __shfl_down_sync(mask, o[j], 1);
        if ((merge_recv >> j) & 1u) o[j] += from_up;
        if ((merge_send >> j) & 1u) continue;
      }
#endif
      if (!((nd >> (ADAFLO_B200_FLAG_SHIFT + c)) & 1u))
        red_add(A.dst_u + 3 * (size_t)(nd & ADAFLO_B200_NODE_MASK) + c, o[j]);
    }
  }
  if (kDiagCapable && A.diag) __syncwarp(mask);
  }  // pass
  if (WITH_P && !A.skip_p_test) {
    double o[4] = {0, 0, 0, 0};
    if (t < 2) {  // y^T over the three lanes' accumulators of plane iz = t
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

// Persistent kernel: one block per SM with kR1Warps independent worker warps and one scheduler warp.  Decades of the
// launch's range [decade_begin, decade_end) are handed out dynamically from a global ticket counter (A.work_counter,
// zeroed by the launcher): tickets go out in time order, so the decades in flight are neighbours in memory, and a warp
// that is delayed (hanging-node cells, a slow SM) simply draws fewer.  (A static assignment with a dynamic tail was
// measured 8 % slower.)  The workers do not draw tickets themselves: ptxas turns an atomic on a warp-uniform address
// into a warp-aggregated one whose result is shuffled -- or spilled -- at once, which exposed the atomic's latency once
// per decade (11 % of the stall samples in profiles/r2_a).  The scheduler warp keeps two shared-memory slots per worker
// filled with that worker's next tickets instead; the worker takes one with a shared-memory load at the end of a decade.
constexpr unsigned kNoTicket = 0xffffffffu;

__device__ __forceinline__ unsigned ld_slot(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_slot(unsigned *p, unsigned v)
{
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE, bool HANGING>
__global__ void __maxnreg__(kR1MaxReg) ns_vmult_r1_kernel(const NSArgs A)
{
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned n_dec = (unsigned)(A.decade_end - A.decade_begin);
#if AB_R1_SCHED
  __shared__ unsigned slot[2 * kR1Warps];  // two ticket slots per worker: a refill (one atomic round trip) is never waited for
  if (threadIdx.x < 2 * kR1Warps) slot[threadIdx.x] = kNoTicket;
  __syncthreads();
  if (warp == kR1Warps) {  // scheduler warp: lane l keeps slot l filled until it has handed out a ticket beyond the range
    if (lane < 2 * kR1Warps) {
      for (;;) {
        if (ld_slot(slot + lane) == kNoTicket) {
          const unsigned t = atomicAdd(A.work_counter, 1u);
          st_slot(slot + lane, t);
          if (t >= n_dec) break;
        } else
          __nanosleep(100);
      }
    }
    return;
  }
  double *warp_smem = smem + warp * kR1WarpSmem;
  unsigned k = 0, finished = 0;  // slot in turn; bit k set: slot k has delivered its final ticket
  while (finished != 3u) {
    if ((finished >> k) & 1u) k ^= 1u;
    unsigned i = 0;
    if (lane == 0) {
      unsigned *const my = slot + 2 * warp + k;
      while ((i = ld_slot(my)) == kNoTicket) {
      }
      st_slot(my, kNoTicket);
    }
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_dec) {
      finished |= 1u << k;
      continue;
    }
    r1_decade<WITH_P, MODE, GEOM, VAR, TRANSPOSE, HANGING>(A, A.decade_begin + i, warp_smem, lane);
    __syncwarp();  // the tiles of this decade have been read before the next decade overwrites them
    k ^= 1u;
  }
#else
  double *warp_smem = smem + warp * kR1WarpSmem;
  // two tickets are held (current and next decade), the third is drawn while the current decade is processed.
  // ptxas turns an atomic on a provably warp-uniform address into a warp-aggregated one and shuffles -- i.e. waits for --
  // its result at once (5 % of the stall samples of profiles/r1_g sat on that shuffle): the address is therefore offset by
  // a zero read from shared memory, which the compiler cannot see through, and the result is consumed after the decade.
  __shared__ unsigned opaque_zero[32];  // indexed by warp: a block-uniform address would make the loaded value provably uniform
  if (threadIdx.x < 32) opaque_zero[threadIdx.x] = 0;
  __syncthreads();
  unsigned i = 0, i_n1 = 0;
  if (lane == 0) {
    i = atomicAdd(A.work_counter, 1u);
    i_n1 = atomicAdd(A.work_counter, 1u);
  }
  i = __shfl_sync(0xffffffffu, i, 0);
  i_n1 = __shfl_sync(0xffffffffu, i_n1, 0);
#if AB_R1_STAGE
  bool staged = false;
#endif
  while (i < n_dec) {
    unsigned i_n2 = 0;
#if AB_R1_PLAIN_TICKET
    if (lane == 0) {
      unsigned z;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(z) : "r"((unsigned)__cvta_generic_to_shared(opaque_zero + warp)) : "memory");
      asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(i_n2) : "l"(A.work_counter + z) : "memory");
    }
#else
    if (lane == 0) i_n2 = atomicAdd(A.work_counter, 1u);
#endif
#if AB_R1_PFSRC
    // paced L2 prefetch of the source vector: where node numbers follow the cell order, the entries first touched by the
    // decades AB_R1_PFSRC tickets ahead lie about that many strides ahead (a hint only: any numbering stays correct)
    if (lane == 0 && A.pf_stride) {
      const size_t off = (A.decade_begin + i + AB_R1_PFSRC) * A.pf_stride;
      if (off + A.pf_stride <= A.pf_u_entries) l2_prefetch(A.src_u + off, (unsigned)(A.pf_stride * 8));
    }
#endif
#if AB_R1_STAGE
    r1_decade<WITH_P, MODE, GEOM, VAR, TRANSPOSE, HANGING>(A, A.decade_begin + i, warp_smem, lane, staged,
                                                           i_n1 < n_dec ? A.decade_begin + i_n1 : kNoDecade);
    staged = i_n1 < n_dec;
#else
    r1_decade<WITH_P, MODE, GEOM, VAR, TRANSPOSE, HANGING>(A, A.decade_begin + i, warp_smem, lane);
#endif
    __syncwarp();
    i = i_n1;
    i_n1 = __shfl_sync(0xffffffffu, i_n2, 0);
  }
#endif
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE, bool HANGING>
int launch_vmult_r1_hn(const NSArgs &a, cudaStream_t s)
{
  auto kern = ns_vmult_r1_kernel<WITH_P, MODE, GEOM, VAR, TRANSPOSE, HANGING>;
  const size_t smem = sizeof(double) * kR1Warps * kR1WarpSmem;
  // function attributes and the SM count are per device: a process may drive several GPUs through several contexts
  static int n_sm_of[64] = {};
  int dev = 0;
  AB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(ADAFLO_B200_ERR_ARG, "ns_vmult_r1: device ordinal out of range");
  if (!n_sm_of[dev]) {
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AB_CUDA(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_sm = n_sm_of[dev];
  if (!a.work_counter) return fail(ADAFLO_B200_ERR_STATE, "ns_vmult_r1: no work counter");
  NSArgs b = a;
  const size_t n_dec = n_groups(a.n_cells);
  if (b.decade_end == 0 || b.decade_end > n_dec) b.decade_end = n_dec;  // default: all decades
  if (b.decade_begin >= b.decade_end) return 0;
  const size_t blocks_needed = (b.decade_end - b.decade_begin + kR1Warps - 1) / kR1Warps;
  const unsigned grid = (unsigned)(blocks_needed < (size_t)n_sm ? blocks_needed : (size_t)n_sm);
  AB_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(unsigned), s));
  kern<<<grid, (kR1Warps + AB_R1_SCHED) * 32, smem, s>>>(b);
  AB_COUNT_LAUNCH();
  AB_CUDA(cudaGetLastError());
  return 0;
}

template <bool WITH_P, int MODE, int GEOM, bool VAR, bool TRANSPOSE>
int launch_vmult_r1_inst(const NSArgs &a, cudaStream_t s)
{
  if (a.hn_u || a.diag) return launch_vmult_r1_hn<WITH_P, MODE, GEOM, VAR, TRANSPOSE, true>(a, s);
  return launch_vmult_r1_hn<WITH_P, MODE, GEOM, VAR, TRANSPOSE, false>(a, s);
}

template <int GEOM>
int launch_vmult_r1_geom(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  const bool tr = a.transpose && (mode == MODE_NEWTON || mode == MODE_PICARD);  // other modes are symmetric
  if (tr) {
    if (!with_p) return fail(ADAFLO_B200_ERR_UNSUPPORTED, "transposed velocity block not instantiated");
    if (mode == MODE_NEWTON && !var) return launch_vmult_r1_inst<true, MODE_NEWTON, GEOM, false, true>(a, s);
    if (mode == MODE_NEWTON && var) return launch_vmult_r1_inst<true, MODE_NEWTON, GEOM, true, true>(a, s);
    if (mode == MODE_PICARD && !var) return launch_vmult_r1_inst<true, MODE_PICARD, GEOM, false, true>(a, s);
    return launch_vmult_r1_inst<true, MODE_PICARD, GEOM, true, true>(a, s);
  }
#define AB_CASE(P, M, V) \
  if (with_p == P && mode == M && var == V) return launch_vmult_r1_inst<P, M, GEOM, V, false>(a, s);
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
  return fail(ADAFLO_B200_ERR_ARG, "launch_vmult_r1_geom: bad variant");
}

}  // namespace ab200
