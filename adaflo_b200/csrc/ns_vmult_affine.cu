// ns_vmult_affine.cu -- instantiations of the Q2/Q1 vmult kernels for affine geometry
// (one translation unit per geometry type so the variants compile in parallel).
#include "ns_q2q1.cuh"

namespace ab200 {
int launch_ns_vmult_affine(const NSArgs &a, bool with_p, int mode, bool var, cudaStream_t s)
{
  return launch_vmult_geom<ADAFLO_B200_GEOM_AFFINE>(a, with_p, mode, var, s);
}
}  // namespace ab200
