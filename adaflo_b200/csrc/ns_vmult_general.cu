// ns_vmult_general.cu -- instantiations of the Q2/Q1 vmult kernels for general geometry
// (one translation unit per geometry type so the variants compile in parallel).
// Default: the 3-lanes-per-cell kernel (ns_r1.cuh).  ADAFLO_B200_VMULT_KERNEL=r2 selects the software-pipelined variant
// with TMA-staged linearisation rows (ns_r2.cuh, 8 instead of 11 warps per SM: measured 4-5 % slower, DESIGN.md section
// 3), =v1 the first design (9 lanes per cell, ns_q2q1.cuh); both are kept for A/B measurements.
#include <cstdlib>
#include <cstring>

#include "ns_r2.cuh"

namespace ab200 {
int launch_ns_vmult_general(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  static const int which = [] {
    const char *e = std::getenv("ADAFLO_B200_VMULT_KERNEL");
    return !e ? 1 : (!std::strcmp(e, "v1") ? 0 : (!std::strcmp(e, "r2") ? 2 : 1));
  }();
  if (which == 0) return launch_vmult_geom<ADAFLO_B200_GEOM_GENERAL>(a, with_p, mode, var, s);
  if (which == 1) return launch_vmult_r1_geom<ADAFLO_B200_GEOM_GENERAL>(a, with_p, mode, var, s);
  return launch_vmult_r2_geom<ADAFLO_B200_GEOM_GENERAL>(a, with_p, mode, var, s);
}
}  // namespace ab200
