// TEST INFRASTRUCTURE — runs the fifth-generation CUDA kernels (ibmfoam_b200/csrc/ibm_kernels_v5.cuh, the very source
// nvcc compiles for sm_100a) under the CPU emulation of tools/cuda_emu, behind a small C surface for ctypes.
// Built by tests/test_emu_v5.py:  g++ -O1 -ffp-contract=off -Itools/cuda_emu -Iibmfoam_b200/csrc -shared -fPIC
#include <cstdlib>
#include <cuda_runtime.h>  // tools/cuda_emu/cuda_runtime.h

#include <vector>

#include "ibm_kernels_v5.cuh"
#include "ibm_kernels_alpha.cuh"

using namespace ibm;

namespace {
struct HostClass {
  ShapeClassDev dev;
  std::vector<double> cosT, sinT, sinP, cosP, ck;
};

// the table part of build_class() in ibm_capi.cu (sphere.C:119-195, cylinder.C:107-152)
void build_class(HostClass& h, int kind, int nTheta, int nkIn, double d, double delta, double l) {
  const double Pi = IBM_PI;
  ShapeClassDev& s = h.dev;
  s = ShapeClassDev{};
  s.kind = kind; s.nTheta = nTheta; s.nk = (kind == 0) ? nTheta / 2 : 1; s.N = 3 * nTheta * s.nk;
  (void)nkIn;
  s.d = d; s.a = 0; s.b = 0; s.l = l; s.delta = delta; s.Rbound = d / 2.0;
  s.dTheta = 2.0 * Pi / nTheta;
  const double dTheta = 2.0 * Pi / (nTheta);
  h.cosT.resize(nTheta); h.sinT.resize(nTheta);
  for (int j = 0; j < nTheta; ++j) { double th = double(j) * dTheta; h.cosT[j] = std::cos(th); h.sinT[j] = std::sin(th); }
  h.sinP.assign(s.nk, 0.0); h.cosP.assign(s.nk, 1.0); h.ck.assign(s.nk, 1.0);
  if (kind == 0) {
    const int nk = s.nk;
    const double dPhi = Pi / (nk - 1);
    for (int k = 0; k < nk; ++k) {
      double phi = dPhi * double(k);
      h.sinP[k] = std::sin(phi); h.cosP[k] = std::cos(phi);
      if (k == 0) h.ck[k] = (std::cos(phi) - std::cos(phi + dPhi / 2.0));
      else if (k == nk - 1) h.ck[k] = (std::cos(phi - dPhi / 2.0) - std::cos(phi));
      else h.ck[k] = (std::cos(phi - dPhi / 2.0) - std::cos(phi + dPhi / 2.0));
    }
  }
  s.cosT = h.cosT.data(); s.sinT = h.sinT.data(); s.sinP = h.sinP.data(); s.cosP = h.cosP.data(); s.ck = h.ck.data();
}

LatticeDev lattice(const int* n, const double* origin, const double* delta, int kz0) {
  LatticeDev L{};
  L.nx = n[0]; L.ny = n[1]; L.nz = n[2]; L.nCells = n[0] * n[1] * n[2];
  L.ox = origin[0]; L.oy = origin[1]; L.oz = origin[2];
  L.dx = delta[0]; L.dy = delta[1]; L.dz = delta[2];
  L.latToCell = nullptr; L.cellToLat = nullptr; L.kz0 = kz0;
  return L;
}

struct Inst {
  std::vector<int> cls;
  InstanceSoA in{};
};
}  // namespace

extern "C" {

// SoA inputs: centre/v/omega [3][nInst]; rot [nInst][9]; F/T out [3][nInst] each
int emu_forcing_v5(const int* n, const double* origin, const double* delta, int kz0, int kind, int nTheta, double d, double dl,
                   int nInst, const int* order, double* centre, double* v, double* om, int* onLocal, int* nShell, int* shellOff,
                   const int* shellCell, const int* shellJK, const double* U, const double* Uold, const double* S, double dt, double* F) {
  LatticeDev L = lattice(n, origin, delta, kz0);
  HostClass hc;
  build_class(hc, kind, nTheta, 1, d, dl, std::fabs(n[2] * delta[2]));
  std::vector<int> cls(nInst, 0);
  InstanceSoA in{};
  in.cls = cls.data();
  for (int q = 0; q < 3; ++q) { in.c[q] = centre + (size_t)q * nInst; in.v[q] = v + (size_t)q * nInst; in.om[q] = om + (size_t)q * nInst; }
  in.onLocal = onLocal; in.nShell = nShell; in.shellOff = shellOff;
  int fallback = 0;
  const ShapeClassDev* classes = &hc.dev;
  emu::launch((unsigned)nInst, F5_THREADS, f5_smem_bytes(), [&]() {
    // EMU_V5_CACHE=1: the instantiation that keeps the stencils between field phases (kernel variant 6)
    if (std::getenv("EMU_V5_CACHE")) k_forcing_v5<true>(L, classes, in, nInst, order, shellCell, shellJK, U, Uold, S, dt, F, &fallback);
    else k_forcing_v5<false>(L, classes, in, nInst, order, shellCell, shellJK, U, Uold, S, dt, F, &fallback);
  });
  return fallback;
}

int emu_stress_v5(const int* n, const double* origin, const double* delta, int kz0, int kind, int nTheta, double d, double dl,
                  int nInst, const int* order, double* centre, double* rot, int* onLocal, const double* P, const double* Tau,
                  double pRefOverRho, double rhoRef, double* Fout, double* Tout, double* Tcorr) {
  LatticeDev L = lattice(n, origin, delta, kz0);
  HostClass hc;
  build_class(hc, kind, nTheta, 1, d, dl, std::fabs(n[2] * delta[2]));
  std::vector<int> cls(nInst, 0);
  InstanceSoA in{};
  in.cls = cls.data();
  for (int q = 0; q < 3; ++q) { in.c[q] = centre + (size_t)q * nInst; in.F[q] = Fout + (size_t)q * nInst; in.T[q] = Tout + (size_t)q * nInst; }
  in.rot = rot; in.onLocal = onLocal;
  int fallback = 0;
  const ShapeClassDev* classes = &hc.dev;
  emu::launch((unsigned)nInst, S5_THREADS, s5_smem_bytes(), [&]() {
    k_stress_v5(L, classes, in, nInst, order, P, Tau, pRefOverRho, rhoRef, Tcorr, &fallback);
  });
  return fallback;
}

// solid volume fraction (extension): kind 0/1 via d, kind 2 via a, b and the per-instance rotation matrices
void emu_solid_fraction(const int* n, const double* origin, const double* delta, int kz0, int kind, double d, double a, double b,
                        int nInst, double* centre, double* rot, int nSub, double* alpha) {
  LatticeDev L = lattice(n, origin, delta, kz0);
  ShapeClassDev sc{};
  sc.kind = kind; sc.d = d; sc.a = a; sc.b = b; sc.Rbound = (kind == 2) ? std::fmax(a, b) : d / 2.0;
  std::vector<int> cls(nInst, 0);
  InstanceSoA in{};
  in.cls = cls.data();
  for (int q = 0; q < 3; ++q) in.c[q] = centre + (size_t)q * nInst;
  in.rot = rot;
  const ShapeClassDev* classes = &sc;
  emu::launch((unsigned)nInst, ALPHA_THREADS, 0, [&]() { k_solid_fraction(L, classes, in, nInst, nSub, alpha); });
  const size_t C = (size_t)L.nCells;
  emu::launch(4, 64, 0, [&]() { k_clamp_unit(alpha, C); });
}

}  // extern "C"
