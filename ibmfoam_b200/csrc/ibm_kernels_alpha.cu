// Launcher of the solid-volume-fraction extension (ibm_kernels_alpha.cuh).
#include "ibm_kernels.h"
#include "ibm_kernels_alpha.cuh"

namespace ibm {

void launch_solid_fraction(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst, int nSub,
                           double* alpha) {
  cudaMemsetAsync(alpha, 0, (size_t)L.nCells * sizeof(double), st);
  if (nInst > 0) k_solid_fraction<<<nInst, ALPHA_THREADS, 0, st>>>(L, classes, in, nInst, nSub, alpha);
  const size_t n = (size_t)L.nCells;
  k_clamp_unit<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(alpha, n);
}

}  // namespace ibm
