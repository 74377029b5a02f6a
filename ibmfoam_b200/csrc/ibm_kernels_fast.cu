// Tuned kernels for loops 2+3 (direct forcing) and 4 (surface-stress integral).
//
// Same arithmetic as the reference-order kernels up to floating-point re-association (bounded by the
// north-star tolerances: forcing 1e-12, force/torque 1e-10 relative):
//   * the linear face interpolation and the inverse-distance gather are fused into one 7-point stencil
//     per shape point:  val = cSelf*phi[cell] + sum_d cNb[d]*phi[nb_d],  cNb[d] = (1/2)(w_d/W),
//     cSelf = sum_d cNb[d]   (lambda = 1/2 on the uniform lattice, particleIBMTmp.C:47 + fvc::interpolate);
//   * 1/|x| is evaluated with rsqrt, products are contracted to FMAs;
//   * reductions use a fixed shuffle/shared-memory tree => deterministic for a given launch shape.
// Point -> cell membership uses the same exact arithmetic as the marking kernel (find_cell).
#include "ibm_kernels.h"

namespace ibm {

struct Stencil7 {
  int self;      // lattice index, -1 = off mesh
  int nb[6];     // lattice index per direction (+x,-x,+y,-y,+z,-z), -1 = no internal face
  double cSelf;
  double cNb[6];
};

__device__ __forceinline__ void make_stencil7(const LatticeDev& L, const V3& p, Stencil7& s) {
  int I, J, K;
  s.self = find_cell(L, p.x, p.y, p.z, I, J, K);
  if (s.self < 0) return;
  const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
  const double ex = Cx - p.x, ey = Cy - p.y, ez = Cz - p.z;  // cell centre - point
  const double hx = 0.5 * L.dx, hy = 0.5 * L.dy, hz = 0.5 * L.dz;
  const double ex2 = ex * ex, ey2 = ey * ey, ez2 = ez * ez;
  double d2[6];
  d2[0] = (ex + hx) * (ex + hx) + ey2 + ez2;
  d2[1] = (ex - hx) * (ex - hx) + ey2 + ez2;
  d2[2] = ex2 + (ey + hy) * (ey + hy) + ez2;
  d2[3] = ex2 + (ey - hy) * (ey - hy) + ez2;
  d2[4] = ex2 + ey2 + (ez + hz) * (ez + hz);
  d2[5] = ex2 + ey2 + (ez - hz) * (ez - hz);
  const bool ok[6] = {I + 1 < L.nx, I > 0, J + 1 < L.ny, J > 0, K + 1 < L.nz, K > 0};
  const int step[6] = {1, -1, L.nx, -L.nx, L.nx * L.ny, -L.nx * L.ny};
  double w[6], W = 0.0;
  int hot = -1;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    s.nb[d] = ok[d] ? s.self + step[d] : -1;
    w[d] = ok[d] ? rsqrt(d2[d]) : 0.0;
    if (ok[d] && d2[d] < IBM_SMALL * IBM_SMALL && hot < 0) hot = d;  // particleShape.C:69-77
    W += w[d];
  }
  if (hot >= 0) {
#pragma unroll
    for (int d = 0; d < 6; ++d) w[d] = (d == hot) ? 1.0 : 0.0;
    W = 1.0;
  }
  const double hW = (W > 0.0) ? 0.5 / W : 0.0;
  double cs = 0.0;
#pragma unroll
  for (int d = 0; d < 6; ++d) { s.cNb[d] = w[d] * hW; cs += s.cNb[d]; }
  s.cSelf = cs;
}

__device__ __forceinline__ double gather7(const Stencil7& s, const double* __restrict__ f) {
  double v = s.cSelf * __ldg(f + s.self);
#pragma unroll
  for (int d = 0; d < 6; ++d)
    if (s.nb[d] >= 0) v = fma(s.cNb[d], __ldg(f + s.nb[d]), v);
  return v;
}

__device__ __forceinline__ V3 rigid_v(const V3& pt, double cx, double cy, double cz, const V3& om, const V3& v) {
  // v + omega x (pt - c)  ==  (-(pt - c) ^ omega) + v     particleIBM.C:387-390
  double rx = pt.x - cx, ry = pt.y - cy, rz = pt.z - cz;
  return V3{v.x + (om.y * rz - om.z * ry), v.y + (om.z * rx - om.x * rz), v.z + (om.x * ry - om.y * rx)};
}

// ------------------------------------------------------------------------------------------------
// K2/K3 v1: one CTA per instance, 4 lanes per shell cell (lane q owns corner q of the polar bin:
// inner point Is[q], surface point Is[q+4], outer point Os[q]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_forcing_v1(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                                                    const int* __restrict__ shellCell, const int* __restrict__ shellJK,
                                                    const double* __restrict__ U, const double* __restrict__ Uold,
                                                    const double* __restrict__ S, double dT, double* __restrict__ F) {
  const int b = blockIdx.x;
  if (b >= nInst) return;
  if (in.onLocal[b] == 0) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const V3 vp{in.v[0][b], in.v[1][b], in.v[2][b]}, om{in.om[0][b], in.om[1][b], in.om[2][b]};
  const int off = in.shellOff[b], n = in.nShell[b];
  const size_t C = (size_t)L.nCells;
  const double rdT = 1.0 / dT;
  const int q = threadIdx.x & 3, grp = threadIdx.x >> 2, ngrp = blockDim.x >> 2;
  const unsigned gmask = 0xfu << ((threadIdx.x & 31) & ~3);
  const int nIter = (n + ngrp - 1) / ngrp;
  for (int it = 0; it < nIter; ++it) {
    const int c = it * ngrp + grp;
    const bool live = c < n;
    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
    int lat = 0;
    if (live) {
      lat = shellCell[off + c];
      const int pk = shellJK[off + c];
      const int j = pk & 0xffff, k = pk >> 16;
      const int I = lat % L.nx, J = (lat / L.nx) % L.ny, K = lat / (L.nx * L.ny);
      const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
      // corner q of the bin (particleShape.C:150-180)
      int pI, pS, pO;
      if (sc.nk != 1 && k == sc.nk - 1) {
        pI = pS = pO = 0;  // App. B #3: row left all-zero
      } else {
        const int jq = (q & 1) ? (j + 1) % sc.nTheta : j;
        const int kq = (sc.nk != 1 && (q & 2)) ? k + 1 : k;
        pI = pt_index(sc.nTheta, 0, jq, kq);
        pS = pI + 1;
        pO = pI + 2;
      }
      const V3 xI = shape_point(sc, rot, cx, cy, cz, pI);
      const V3 xS = shape_point(sc, rot, cx, cy, cz, pS);
      const V3 xO = shape_point(sc, rot, cx, cy, cz, pO);
      // wFromLocal_: distances to the inner (q) and surface (q+4) points   particleShape.C:118-130
      const double dI2 = (Cx - xI.x) * (Cx - xI.x) + (Cy - xI.y) * (Cy - xI.y) + (Cz - xI.z) * (Cz - xI.z);
      const double dS2 = (Cx - xS.x) * (Cx - xS.x) + (Cy - xS.y) * (Cy - xS.y) + (Cz - xS.z) * (Cz - xS.z);
      double wI = rsqrt(dI2), wS = rsqrt(dS2);
      // one-hot rule: first q (0..7) with diff < SMALL wins
      const unsigned hotI = __ballot_sync(gmask, dI2 < IBM_SMALL * IBM_SMALL) & gmask;
      const unsigned hotS = __ballot_sync(gmask, dS2 < IBM_SMALL * IBM_SMALL) & gmask;
      if (hotI | hotS) {
        const int sh = (threadIdx.x & 31) & ~3;
        int first = hotI ? (__ffs(hotI >> sh) - 1) : (4 + __ffs(hotS >> sh) - 1);
        wI = (first == q) ? 1.0 : 0.0;
        wS = (first == q + 4) ? 1.0 : 0.0;
      }
      double W = wI + wS;
      W += __shfl_xor_sync(gmask, W, 1);
      W += __shfl_xor_sync(gmask, W, 2);
      const double rW = 1.0 / W;
      Stencil7 sI, sS, sO;
      make_stencil7(L, xI, sI);
      make_stencil7(L, xS, sS);
      make_stencil7(L, xO, sO);
      const V3 vO = rigid_v(xO, cx, cy, cz, om, vp);
      const V3 vS = rigid_v(xS, cx, cy, cz, om, vp);
      const double aO = wI * rW, aS = wS * rW;  // note: the outer group is weighted by the inner distance (App. B #7)
#pragma unroll
      for (int cmp = 0; cmp < 3; ++cmp) {
        const double* u = U + cmp * C;
        const double* uo = Uold + cmp * C;
        const double* s = S + cmp * C;
        const double gUO = sO.self >= 0 ? gather7(sO, u) : 0.0;
        const double gSO = sO.self >= 0 ? gather7(sO, s) : 0.0;
        const double gUoI = sI.self >= 0 ? gather7(sI, uo) : 0.0;
        const double gUoS = sS.self >= 0 ? gather7(sS, uo) : 0.0;
        const double gSS = sS.self >= 0 ? gather7(sS, s) : 0.0;
        const double vo = cmp == 0 ? vO.x : (cmp == 1 ? vO.y : vO.z);
        const double vs = cmp == 0 ? vS.x : (cmp == 1 ? vS.y : vS.z);
        const double t = aO * ((vo - gUO - gUoI) * rdT + gSO) + aS * ((vs - gUoS) * rdT + gSS);
        if (cmp == 0) Fx = t; else if (cmp == 1) Fy = t; else Fz = t;
      }
    }
    // all 4 lanes of a group are live together (c depends on grp only)
    Fx += __shfl_xor_sync(gmask, Fx, 1); Fx += __shfl_xor_sync(gmask, Fx, 2);
    Fy += __shfl_xor_sync(gmask, Fy, 1); Fy += __shfl_xor_sync(gmask, Fy, 2);
    Fz += __shfl_xor_sync(gmask, Fz, 1); Fz += __shfl_xor_sync(gmask, Fz, 2);
    if (live && q == 0) {
      atomicAdd(&F[lat], Fx);
      atomicAdd(&F[C + lat], Fy);
      atomicAdd(&F[2 * C + lat], Fz);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v1: one CTA per instance, one thread per surface point, fixed-tree reduction
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stress_v1(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                                                   const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho,
                                                   double rhoRef, double* __restrict__ Tcorr) {
  __shared__ double red[8][9];
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const size_t C = (size_t)L.nCells;
  const int nS = sc.nTheta * sc.nk;
  const bool on = in.onLocal[b] != 0;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < nS; s += blockDim.x) {
    const int i = s / sc.nk, k = s % sc.nk;
    const int i2 = i + sc.nTheta * k;
    const V3 pt = shape_point(sc, rot, cx, cy, cz, pt_index(sc.nTheta, 1, i, k));
    double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
    if (on) {
      Stencil7 st;
      make_stencil7(L, pt, st);
      if (st.self >= 0) {
#pragma unroll
        for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gather7(st, Tau + cmp * C) * rhoRef;
        // sum of the 7 coefficients is 1 when the cell has internal faces (else 0: nothing gathered)
        const double csum = 2.0 * st.cSelf;
        pp = (gather7(st, P) - csum * pRefOverRho) * rhoRef;
      }
    }
    const V3 m2 = shape_point(sc, rot, cx, cy, cz, i2);  // App. B #5: lever arm as coded
    const double Rx = m2.x - cx, Ry = m2.y - cy, Rz = m2.z - cz;
    const double Rcx = pt.x - cx, Rcy = pt.y - cy, Rcz = pt.z - cz;
    const double sx = __ldg(sc.Sf + 3 * i2), sy = __ldg(sc.Sf + 3 * i2 + 1), sz = __ldg(sc.Sf + 3 * i2 + 2);
    const double Sx = rot[0] * sx + rot[1] * sy + rot[2] * sz, Sy = rot[3] * sx + rot[4] * sy + rot[5] * sz,
                 Sz = rot[6] * sx + rot[7] * sy + rot[8] * sz;
    const double rmag = rsqrt(Sx * Sx + Sy * Sy + Sz * Sz);
    const double nx = Sx * rmag, ny = Sy * rmag, nz = Sz * rmag;
    const double fx = pp * sx + (tau[0] * Sx + tau[1] * Sy + tau[2] * Sz);
    const double fy = pp * sy + (tau[1] * Sx + tau[3] * Sy + tau[4] * Sz);
    const double fz = pp * sz + (tau[2] * Sx + tau[4] * Sy + tau[5] * Sz);
    const double fn = fx * nx + fy * ny + fz * nz;
    const double ftx = fx - fn * nx, fty = fy - fn * ny, ftz = fz - fn * nz;
    acc[0] += fx; acc[1] += fy; acc[2] += fz;
    acc[3] -= fty * Rz - ftz * Ry; acc[4] -= ftz * Rx - ftx * Rz; acc[5] -= ftx * Ry - fty * Rx;
    acc[6] -= fty * Rcz - ftz * Rcy; acc[7] -= ftz * Rcx - ftx * Rcz; acc[8] -= ftx * Rcy - fty * Rcx;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    const int q = threadIdx.x;
    if (q < 3) in.F[q][b] = v;
    else if (q < 6) in.T[q - 3][b] = v;
    else if (Tcorr) Tcorr[3 * b + (q - 6)] = v;
  }
}

void launch_forcing_fast(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                         const int* shellCell, const int* shellJK, const double* U, const double* Uold, const double* S, double dT,
                         double* F) {
  if (nInst > 0) k_forcing_v1<<<nInst, 128, 0, st>>>(L, classes, in, nInst, shellCell, shellJK, U, Uold, S, dT, F);
}
void launch_stress_fast(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                        const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr) {
  if (nInst > 0) k_stress_v1<<<nInst, 256, 0, st>>>(L, classes, in, nInst, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

}  // namespace ibm
