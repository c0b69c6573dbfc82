// hanging.cuh -- in-kernel resolution of 2:1 hanging-node constraints.
//
// deal.II resolves hanging nodes inside FEEvaluation::read_dof_values /
// distribute_local_to_global from a per-cell refinement configuration.
// Same idea here: on a hanging face (edge) the cell's node list holds the PARENT's face (edge)
// nodes; the gathered slot values are interpolated direction by direction with the 1-D
// parent->child matrix, every line at most once, and the transposed matrices are applied before
// the scatter.  Mask bits (adaflo_b200/mesh.py HN_*): 0-2 position of the cell in its parent per
// direction, 3-5 hanging outer face with normal x/y/z, 6-8 hanging outer edge along x/y/z without
// an adjacent hanging face.
#pragma once

#include <stdint.h>

namespace ab200 {

struct HnInfo {
  uint32_t m;
  __host__ __device__ __forceinline__ explicit HnInfo(uint32_t mask) : m(mask) {}
  __host__ __device__ __forceinline__ bool any() const { return m != 0; }
  __host__ __device__ __forceinline__ int sub(int d) const { return (m >> d) & 1; }
  __host__ __device__ __forceinline__ bool face(int d) const { return (m >> (3 + d)) & 1; }
  __host__ __device__ __forceinline__ bool edge(int d) const { return (m >> (6 + d)) & 1; }
  // local index of the outer (parent-boundary) face in direction d for n1 nodes per direction
  __host__ __device__ __forceinline__ int outer(int d, int n1) const { return sub(d) ? n1 - 1 : 0; }
  // does the line along direction d through (ie, jf) [indices along d+1, d+2 (cyclic)] need interpolation?
  __host__ __device__ __forceinline__ bool line(int d, int ie, int jf, int n1) const
  {
    const int e = (d + 1) % 3, f = (d + 2) % 3;
    const bool on_e = ie == outer(e, n1), on_f = jf == outer(f, n1);
    return (face(e) && on_e) || (face(f) && on_f) || (edge(d) && on_e && on_f);
  }
};

// Q2: parent nodes {0, 1/2, 1} -> child nodes; child 0 = [0,1/2]: rows (1,0,0), (3/8,3/4,-1/8), (0,1,0);
// child 1 = [1/2,1]: rows (0,1,0), (-1/8,3/4,3/8), (0,0,1)
__device__ __forceinline__ void hn_q2(const int sub, double &v0, double &v1, double &v2)
{
  const double mid = sub ? fma(0.375, v2, -0.125 * v0) : fma(0.375, v0, -0.125 * v2);
  const double n1 = fma(0.75, v1, mid);
  const double n0 = sub ? v1 : v0, n2 = sub ? v2 : v1;
  v0 = n0;
  v1 = n1;
  v2 = n2;
}
__device__ __forceinline__ void hn_q2_t(const int sub, double &r0, double &r1, double &r2)
{
  const double o0 = sub ? -0.125 * r1 : fma(0.375, r1, r0);
  const double o1 = sub ? fma(0.75, r1, r0) : fma(0.75, r1, r2);
  const double o2 = sub ? fma(0.375, r1, r2) : -0.125 * r1;
  r0 = o0;
  r1 = o1;
  r2 = o2;
}
// Q1: child 0: rows (1,0), (1/2,1/2); child 1: rows (1/2,1/2), (0,1)
__device__ __forceinline__ void hn_q1(const int sub, double &v0, double &v1)
{
  const double h = 0.5 * (v0 + v1);
  if (sub) v0 = h; else v1 = h;
}
__device__ __forceinline__ void hn_q1_t(const int sub, double &r0, double &r1)
{
  if (sub) {
    r1 = fma(0.5, r0, r1);
    r0 = 0.5 * r0;
  } else {
    r0 = fma(0.5, r1, r0);
    r1 = 0.5 * r1;
  }
}

}  // namespace ab200

namespace ab200 {

// ---- 9-lanes-per-cell kernels (residual, divergence): lane (a, b) = (ix, iy) owns the z column ----
// x / y lines live across lanes (lane_base + a + 3 b), z lines in registers.
__device__ __forceinline__ double hn_pick3(int i, double v0, double v1, double v2) { return i == 0 ? v0 : (i == 1 ? v1 : v2); }

__device__ __forceinline__ void hn_resolve9_u(const HnInfo hn, bool warp_any, int a, int b, int lane_base, unsigned mask, double (&col)[3][3])
{
  if (!warp_any) return;
  if (hn.any() && hn.line(2, a, b, 3)) {
#pragma unroll
    for (int c = 0; c < 3; ++c) hn_q2(hn.sub(2), col[c][0], col[c][1], col[c][2]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool nx = hn.any() && hn.line(0, b, k, 3), ny = hn.any() && hn.line(1, k, a, 3);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v0 = __shfl_sync(mask, col[c][k], lane_base + 3 * b), v1 = __shfl_sync(mask, col[c][k], lane_base + 3 * b + 1),
             v2 = __shfl_sync(mask, col[c][k], lane_base + 3 * b + 2);
      hn_q2(hn.sub(0), v0, v1, v2);
      if (nx) col[c][k] = hn_pick3(a, v0, v1, v2);
      v0 = __shfl_sync(mask, col[c][k], lane_base + a);
      v1 = __shfl_sync(mask, col[c][k], lane_base + a + 3);
      v2 = __shfl_sync(mask, col[c][k], lane_base + a + 6);
      hn_q2(hn.sub(1), v0, v1, v2);
      if (ny) col[c][k] = hn_pick3(b, v0, v1, v2);
    }
  }
}

__device__ __forceinline__ void hn_distribute9_u(const HnInfo hn, bool warp_any, int a, int b, int lane_base, unsigned mask, double (&out)[3][3])
{
  if (!warp_any) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool nx = hn.any() && hn.line(0, b, k, 3), ny = hn.any() && hn.line(1, k, a, 3);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double r0 = __shfl_sync(mask, out[c][k], lane_base + a), r1 = __shfl_sync(mask, out[c][k], lane_base + a + 3),
             r2 = __shfl_sync(mask, out[c][k], lane_base + a + 6);
      hn_q2_t(hn.sub(1), r0, r1, r2);
      if (ny) out[c][k] = hn_pick3(b, r0, r1, r2);
      r0 = __shfl_sync(mask, out[c][k], lane_base + 3 * b);
      r1 = __shfl_sync(mask, out[c][k], lane_base + 3 * b + 1);
      r2 = __shfl_sync(mask, out[c][k], lane_base + 3 * b + 2);
      hn_q2_t(hn.sub(0), r0, r1, r2);
      if (nx) out[c][k] = hn_pick3(a, r0, r1, r2);
    }
  }
  if (hn.any() && hn.line(2, a, b, 3)) {
#pragma unroll
    for (int c = 0; c < 3; ++c) hn_q2_t(hn.sub(2), out[c][0], out[c][1], out[c][2]);
  }
}

// pressure (Q1): lanes a, b < 2 hold the column pc[k], k = 0, 1; every lane of the cell takes part in the shuffles
__device__ __forceinline__ void hn_resolve9_p(const HnInfo hn, bool warp_any, int a, int b, int lane_base, unsigned mask, double (&pc)[2])
{
  if (!warp_any) return;
  const bool mine = a < 2 && b < 2;
  if (mine && hn.any() && hn.line(2, a, b, 2)) hn_q1(hn.sub(2), pc[0], pc[1]);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double v0 = __shfl_sync(mask, pc[k], lane_base + 3 * (b & 1)), v1 = __shfl_sync(mask, pc[k], lane_base + 3 * (b & 1) + 1);
    hn_q1(hn.sub(0), v0, v1);
    if (mine && hn.any() && hn.line(0, b, k, 2)) pc[k] = a == 0 ? v0 : v1;
    v0 = __shfl_sync(mask, pc[k], lane_base + (a & 1));
    v1 = __shfl_sync(mask, pc[k], lane_base + (a & 1) + 3);
    hn_q1(hn.sub(1), v0, v1);
    if (mine && hn.any() && hn.line(1, k, a, 2)) pc[k] = b == 0 ? v0 : v1;
  }
}

__device__ __forceinline__ void hn_distribute9_p(const HnInfo hn, bool warp_any, int a, int b, int lane_base, unsigned mask, double (&out)[2])
{
  if (!warp_any) return;
  const bool mine = a < 2 && b < 2;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double r0 = __shfl_sync(mask, out[k], lane_base + (a & 1)), r1 = __shfl_sync(mask, out[k], lane_base + (a & 1) + 3);
    hn_q1_t(hn.sub(1), r0, r1);
    if (mine && hn.any() && hn.line(1, k, a, 2)) out[k] = b == 0 ? r0 : r1;
    r0 = __shfl_sync(mask, out[k], lane_base + 3 * (b & 1));
    r1 = __shfl_sync(mask, out[k], lane_base + 3 * (b & 1) + 1);
    hn_q1_t(hn.sub(0), r0, r1);
    if (mine && hn.any() && hn.line(0, b, k, 2)) out[k] = a == 0 ? r0 : r1;
  }
  if (mine && hn.any() && hn.line(2, a, b, 2)) hn_q1_t(hn.sub(2), out[0], out[1]);
}

}  // namespace ab200
