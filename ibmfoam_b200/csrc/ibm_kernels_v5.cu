// Launchers of the fifth-generation kernels (ibm_kernels_v5.cuh).
#include "ibm_kernels.h"
#include "ibm_kernels_v5.cuh"

namespace ibm {

void launch_forcing_v5(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, const int* shellCell, const int* shellJK, const double* U, const double* Uold,
                       const double* S, double dT, double* F, int* dFallback, bool cacheStencils) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(k_forcing_v5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f5_smem_bytes());
    cudaFuncSetAttribute(k_forcing_v5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f5_smem_bytes());
  }
  if (nInst <= 0) return;
  if (cacheStencils)
    k_forcing_v5<true><<<nInst, F5_THREADS, f5_smem_bytes(), st>>>(L, classes, in, nInst, s.order, shellCell, shellJK, U, Uold, S, dT, F, dFallback);
  else
    k_forcing_v5<false><<<nInst, F5_THREADS, f5_smem_bytes(), st>>>(L, classes, in, nInst, s.order, shellCell, shellJK, U, Uold, S, dT, F, dFallback);
}

void launch_stress_v5(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr, int* dFallback) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(k_stress_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s5_smem_bytes());
  if (nInst > 0)
    k_stress_v5<<<nInst, S5_THREADS, s5_smem_bytes(), st>>>(L, classes, in, nInst, s.order, P, Tau, pRefOverRho, rhoRef, Tcorr, dFallback);
}

}  // namespace ibm
