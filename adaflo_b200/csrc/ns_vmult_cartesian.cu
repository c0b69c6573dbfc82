// ns_vmult_cartesian.cu -- instantiations of the Q2/Q1 vmult kernels for cartesian geometry
// (one translation unit per geometry type so the variants compile in parallel).
// Default: the 3-lanes-per-cell kernel (ns_r1.cuh); ADAFLO_B200_VMULT_KERNEL=v1 selects the
// first design (9 lanes per cell, ns_q2q1.cuh), kept for A/B measurements.
#include <cstdlib>
#include <cstring>

#include "ns_r1.cuh"

namespace ab200 {
int launch_ns_vmult_cartesian(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  static const bool use_v1 = [] { const char *e = std::getenv("ADAFLO_B200_VMULT_KERNEL"); return e && !std::strcmp(e, "v1"); }();
  if (use_v1) return launch_vmult_geom<ADAFLO_B200_GEOM_CARTESIAN>(a, with_p, mode, var, s);
  return launch_vmult_r1_geom<ADAFLO_B200_GEOM_CARTESIAN>(a, with_p, mode, var, s);
}
}  // namespace ab200
