// Tuned kernels, second generation (loops 2+3 and 4).
//
// Forcing (k_forcing_v2).  Work item = up to FORCING_CHUNK shell cells of one particle instance; items are
// laid out along a Morton-sorted instance order so CTAs that run together touch neighbouring parts of the
// fields (L2 reuse).  Inside an item:
//   A  every shell cell registers the four corner "columns" (j,k) of its polar bin in a shared-memory hash;
//      a column is the radial triple (inner, surface, outer) of shape points, shared by up to four bins;
//   B  one thread per distinct column evaluates the three 7-point gather stencils ONCE and stores the two
//      combined terms the scatter needs,
//          tO = (v(x_O) - U(x_O) - Uold(x_I))/dt + S(x_O)        tS = (v(x_S) - Uold(x_S))/dt + S(x_S)
//      (particleIBM.C:238-260; note App. B #7: the outer term mixes in Uold at the INNER point);
//   C  one thread per shell cell forms F_i = sum_q (wI_q tO_q + wS_q tS_q)/W from shared memory and adds it
//      to the forcing field.
// Arithmetic differs from the reference-order kernel only by re-association (<= 1e-12 relative, tested).
#include <cstdlib>

#include "ibm_kernels.h"

namespace ibm {


struct Stencil7 {
  int self;
  int nb[6];
  double cSelf;
  double cNb[6];
};

__device__ __forceinline__ void make_stencil7(const LatticeDev& L, double idx, double idy, double idz, const V3& p, Stencil7& s) {
  int I, J, K;
  s.self = find_cell_fast(L, idx, idy, idz, p.x, p.y, p.z, I, J, K);
  if (s.self < 0) return;
  const double ex = (L.ox + (I + 0.5) * L.dx) - p.x, ey = (L.oy + (J + 0.5) * L.dy) - p.y, ez = (L.oz + (K + L.kz0 + 0.5) * L.dz) - p.z;
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
  double rx = pt.x - cx, ry = pt.y - cy, rz = pt.z - cz;
  return V3{v.x + (om.y * rz - om.z * ry), v.y + (om.z * rx - om.x * rz), v.z + (om.x * ry - om.y * rx)};
}

template <int HBITS>
__device__ __forceinline__ unsigned hash_col(int key) { return ((unsigned)key * 2654435761u) >> (32 - HBITS); }

// schedule lookup: position with chunkOff[pos] <= item < chunkOff[pos+1]
__device__ __forceinline__ int find_pos(const int* __restrict__ off, int n, int item) {
  int lo = 0, hi = n;  // invariant: off[lo] <= item < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= item) lo = mid; else hi = mid;
  }
  return lo;
}

// one work item: shell cells [c0, c0 + CH) of instance b (block-wide; ends with a barrier so that the caller may loop)
template <int CH, int HBITS>
__device__ __forceinline__ void forcing_v2_item(const LatticeDev& L, const ShapeClassDev* __restrict__ classes, const InstanceSoA& in, int b, int c0,
                                                const int* __restrict__ shellCell, const int* __restrict__ shellJK,
                                                const double* __restrict__ U, const double* __restrict__ Uold, const double* __restrict__ S,
                                                double dT, double* __restrict__ F, double* smem2) {
  constexpr int FHASH = 1 << HBITS, FMAXCOL = 4 * CH + 1;
  double (*colVal)[6] = (double (*)[6])smem2;             // [FMAXCOL][6]
  int* hKey = (int*)(smem2 + 6 * FMAXCOL);                // [FHASH]
  int* colKey = hKey + FHASH;                             // [FMAXCOL]
  unsigned short* hSlot = (unsigned short*)(colKey + FMAXCOL + 1);  // [FHASH]
  __shared__ int nCol;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const int off = in.shellOff[b] + c0;
  const int n = min(CH, in.nShell[b] - c0);
  const size_t C = (size_t)L.nCells;
  const int nT = sc.nTheta;
  const int degenerateKey = nT * sc.nk;  // App. B #3: bins with k == nk-1 reference point 0 twelve times

  for (int t = threadIdx.x; t < FHASH; t += blockDim.x) hKey[t] = -1;
  if (threadIdx.x == 0) nCol = 0;
  // (shape points come from the class template in global memory: shared-memory trig tables were measured to cost
  //  more than they save here — every KB of shared memory is a KB less L1 for the gathers)
  __syncthreads();

  // ---- A: register columns ----
  int myLat = 0, myKeys[4] = {0, 0, 0, 0};
  const bool live = threadIdx.x < n;
  if (live) {
    myLat = shellCell[off + threadIdx.x];
    const int pk = shellJK[off + threadIdx.x];
    const int j = pk & 0xffff, k = pk >> 16;
    const int jp1 = (j + 1) % nT;
    if (sc.nk != 1 && k == sc.nk - 1) {
      myKeys[0] = myKeys[1] = myKeys[2] = myKeys[3] = degenerateKey;
    } else {
      const int k1 = (sc.nk != 1) ? k + 1 : k;
      myKeys[0] = j + nT * k; myKeys[1] = jp1 + nT * k; myKeys[2] = j + nT * k1; myKeys[3] = jp1 + nT * k1;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int key = myKeys[q];
      unsigned h = hash_col<HBITS>(key);
      while (true) {
        int old = atomicCAS(&hKey[h], -1, key);
        if (old == -1) {
          int slot = atomicAdd(&nCol, 1);
          hSlot[h] = (unsigned short)slot;
          colKey[slot] = key;
          break;
        }
        if (old == key) break;
        h = (h + 1) & (FHASH - 1);
      }
    }
  }
  __syncthreads();

  // ---- B: one thread per distinct column ----
  const int ncol = nCol;
  const V3 vp{in.v[0][b], in.v[1][b], in.v[2][b]}, om{in.om[0][b], in.om[1][b], in.om[2][b]};
  const double rdT = 1.0 / dT;
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
  for (int s = threadIdx.x; s < ncol; s += blockDim.x) {
    const int key = colKey[s];
    int pI, pS, pO;
    const bool deg = key == degenerateKey;
    const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
    if (deg) { pI = pS = pO = 0; }
    else { pI = pt_index(nT, 0, jq, kq); pS = pI + 1; pO = pI + 2; }
    double tO[3], tS[3];
    Stencil7 st;
    {  // inner point: Uold only
      const V3 x = shape_point(sc, rot, cx, cy, cz, pI);
      make_stencil7(L, idx, idy, idz, x, st);
#pragma unroll
      for (int q = 0; q < 3; ++q) tO[q] = st.self >= 0 ? -gather7(st, Uold + q * C) : 0.0;
    }
    {  // outer point: U and S, rigid-body velocity
      const V3 x = shape_point(sc, rot, cx, cy, cz, pO);
      make_stencil7(L, idx, idy, idz, x, st);
      const V3 vv = rigid_v(x, cx, cy, cz, om, vp);
      const double v3[3] = {vv.x, vv.y, vv.z};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double gu = st.self >= 0 ? gather7(st, U + q * C) : 0.0;
        const double gs = st.self >= 0 ? gather7(st, S + q * C) : 0.0;
        tO[q] = (v3[q] - gu + tO[q]) * rdT + gs;
      }
    }
    {  // surface point: Uold and S
      const V3 x = shape_point(sc, rot, cx, cy, cz, pS);
      make_stencil7(L, idx, idy, idz, x, st);
      const V3 vv = rigid_v(x, cx, cy, cz, om, vp);
      const double v3[3] = {vv.x, vv.y, vv.z};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double guo = st.self >= 0 ? gather7(st, Uold + q * C) : 0.0;
        const double gs = st.self >= 0 ? gather7(st, S + q * C) : 0.0;
        tS[q] = (v3[q] - guo) * rdT + gs;
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) { colVal[s][q] = tO[q]; colVal[s][3 + q] = tS[q]; }
  }
  __syncthreads();

  // ---- C: one thread per shell cell ----
  if (live) {
    const int I = myLat % L.nx, J = (myLat / L.nx) % L.ny, K = myLat / (L.nx * L.ny);
    const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
    double Fx = 0.0, Fy = 0.0, Fz = 0.0, W = 0.0;
    double wI[4], wS[4];
    int slot[4];
    int first = -1;  // one-hot rule of particleShape.C:121-127: first q (0..7) within SMALL wins; q < 4 are inner points
    int firstS = -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int key = myKeys[q];
      unsigned h = hash_col<HBITS>(key);
      while (hKey[h] != key) h = (h + 1) & (FHASH - 1);
      slot[q] = hSlot[h];
      const bool deg = key == degenerateKey;
      const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
      const int pI = deg ? 0 : pt_index(nT, 0, jq, kq);
      const int pS = deg ? 0 : pI + 1;
      const V3 xI = shape_point(sc, rot, cx, cy, cz, pI);
      const V3 xS = shape_point(sc, rot, cx, cy, cz, pS);
      const double dI2 = (Cx - xI.x) * (Cx - xI.x) + (Cy - xI.y) * (Cy - xI.y) + (Cz - xI.z) * (Cz - xI.z);
      const double dS2 = (Cx - xS.x) * (Cx - xS.x) + (Cy - xS.y) * (Cy - xS.y) + (Cz - xS.z) * (Cz - xS.z);
      wI[q] = rsqrt(dI2);
      wS[q] = rsqrt(dS2);
      if (dI2 < IBM_SMALL * IBM_SMALL && first < 0) first = q;
      if (dS2 < IBM_SMALL * IBM_SMALL && firstS < 0) firstS = 4 + q;
    }
    if (first < 0) first = firstS;
    if (first >= 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { wI[q] = (first == q) ? 1.0 : 0.0; wS[q] = (first == 4 + q) ? 1.0 : 0.0; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      W += wI[q] + wS[q];
      const double* cv = colVal[slot[q]];
      Fx = fma(wI[q], cv[0], fma(wS[q], cv[3], Fx));
      Fy = fma(wI[q], cv[1], fma(wS[q], cv[4], Fy));
      Fz = fma(wI[q], cv[2], fma(wS[q], cv[5], Fz));
    }
    const double rW = 1.0 / W;
    atomicAdd(&F[myLat], Fx * rW);
    atomicAdd(&F[C + myLat], Fy * rW);
    atomicAdd(&F[2 * C + myLat], Fz * rW);
  }
  __syncthreads();
}

template <int CH, int HBITS, int MINB>
__global__ void __launch_bounds__(CH, MINB)
k_forcing_v2(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
             const int* __restrict__ chunkOff, const int* __restrict__ shellCell, const int* __restrict__ shellJK,
             const double* __restrict__ U, const double* __restrict__ Uold, const double* __restrict__ S, double dT,
             double* __restrict__ F) {
  extern __shared__ double smem2[];
  const int item = blockIdx.x;
  const int pos = find_pos(chunkOff, nInst, item);
  forcing_v2_item<CH, HBITS>(L, classes, in, order[pos], (item - chunkOff[pos]) * CH, shellCell, shellJK, U, Uold, S, dT, F, smem2);
}

// the instances a staged kernel (ibm_kernels_v7.cu) could not take, from a device-side list: persistent blocks, every work
// item of every listed instance (usually the list is empty and the kernel returns at once)
template <int CH, int HBITS>
__global__ void __launch_bounds__(CH, 4)
k_forcing_v2_list(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, const int* __restrict__ list,
                  const int* __restrict__ listCount, const int* __restrict__ shellCell, const int* __restrict__ shellJK,
                  const double* __restrict__ U, const double* __restrict__ Uold, const double* __restrict__ S, double dT,
                  double* __restrict__ F) {
  extern __shared__ double smem2[];
  const int n = *listCount;
  // work items are enumerated (instance, chunk) with a fixed bound of chunks per instance so that no prefix sum is needed
  for (int w = blockIdx.x;; w += gridDim.x) {
    const int e = w / 64, ch = w % 64;
    if (e >= n) break;
    const int b = list[e];
    const int ns = in.nShell[b];
    for (int c0 = ch * CH; c0 < ns; c0 += 64 * CH) forcing_v2_item<CH, HBITS>(L, classes, in, b, c0, shellCell, shellJK, U, Uold, S, dT, F, smem2);
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v2: one CTA per instance in Morton order, one thread per surface point, fixed-tree reduction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stress_v2_instance(const LatticeDev& L, const ShapeClassDev* __restrict__ classes, const InstanceSoA& in, int b,
                                                   const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho,
                                                   double rhoRef, double* __restrict__ Tcorr) {
  __shared__ double red[8][9];
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const size_t C = (size_t)L.nCells;
  const int nS = sc.nTheta * sc.nk;
  const bool on = in.onLocal[b] != 0;
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < nS; s += blockDim.x) {
    const int i = s / sc.nk, k = s % sc.nk;
    const int i2 = i + sc.nTheta * k;
    const V3 pt = shape_point(sc, rot, cx, cy, cz, pt_index(sc.nTheta, 1, i, k));
    double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
    if (on) {
      Stencil7 st;
      make_stencil7(L, idx, idy, idz, pt, st);
      if (st.self >= 0 && (!has_halo(L) || owned_plane(L, st.self / (L.nx * L.ny)))) {  // halo: integrated by the owner of the point's plane
#pragma unroll
        for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gather7(st, Tau + cmp * C) * rhoRef;
        pp = (gather7(st, P) - 2.0 * st.cSelf * pRefOverRho) * rhoRef;
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
  __syncthreads();
}

__global__ void __launch_bounds__(256, 3)
k_stress_v2(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
            const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef, double* __restrict__ Tcorr) {
  if ((int)blockIdx.x >= nInst) return;
  stress_v2_instance(L, classes, in, order[blockIdx.x], P, Tau, pRefOverRho, rhoRef, Tcorr);
}

__global__ void __launch_bounds__(256, 3)
k_stress_v2_list(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, const int* __restrict__ list,
                 const int* __restrict__ listCount, const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho,
                 double rhoRef, double* __restrict__ Tcorr) {
  const int n = *listCount;
  for (int e = blockIdx.x; e < n; e += gridDim.x) stress_v2_instance(L, classes, in, list[e], P, Tau, pRefOverRho, rhoRef, Tcorr);
}

void launch_forcing_v2_list(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, const int* list,
                            const int* listCount, const int* shellCell, const int* shellJK, const double* U, const double* Uold,
                            const double* S, double dT, double* F) {
  constexpr int CH = 128, HB = 10;
  const size_t sm = (size_t)(4 * CH + 1) * 6 * sizeof(double) + (size_t)((1 << HB) + 4 * CH + 2) * sizeof(int) + (1 << HB) * sizeof(unsigned short);
  k_forcing_v2_list<CH, HB><<<592, CH, sm, st>>>(L, classes, in, list, listCount, shellCell, shellJK, U, Uold, S, dT, F);
}
void launch_stress_v2_list(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, const int* list,
                           const int* listCount, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr) {
  k_stress_v2_list<<<444, 256, 0, st>>>(L, classes, in, list, listCount, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

void launch_forcing_v2(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, int totalChunks, const int* shellCell, const int* shellJK, const double* U,
                       const double* Uold, const double* S, double dT, double* F) {
  if (nInst <= 0 || totalChunks <= 0) return;
  if (s.chunk == 256) {
    constexpr int CH = 256, HB = 11;
    const size_t sm = (size_t)(4 * CH + 1) * 6 * sizeof(double) + (size_t)((1 << HB) + 4 * CH + 2) * sizeof(int) + (1 << HB) * sizeof(unsigned short);
    static PerDeviceOnce attr;
    if (attr.first()) cudaFuncSetAttribute(k_forcing_v2<CH, HB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_forcing_v2<CH, HB, 2><<<totalChunks, CH, sm, st>>>(L, classes, in, nInst, s.order, s.chunkOff, shellCell, shellJK, U, Uold, S, dT, F);
  } else {
    constexpr int CH = 128, HB = 10;
    const size_t sm = (size_t)(4 * CH + 1) * 6 * sizeof(double) + (size_t)((1 << HB) + 4 * CH + 2) * sizeof(int) + (1 << HB) * sizeof(unsigned short);
    static PerDeviceOnce attr2;
    if (attr2.first()) {
      // the gathers are served by L1: leave it as much of the unified 228 KB as the resident CTAs allow
      int pct = getenv("IBM_V2_CARVEOUT") ? atoi(getenv("IBM_V2_CARVEOUT")) : -1;
      if (pct >= 0) cudaFuncSetAttribute(k_forcing_v2<CH, HB, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    k_forcing_v2<CH, HB, 5><<<totalChunks, CH, sm, st>>>(L, classes, in, nInst, s.order, s.chunkOff, shellCell, shellJK, U, Uold, S, dT, F);
  }
}
void launch_stress_v2(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr) {
  if (nInst > 0) k_stress_v2<<<nInst, 256, 0, st>>>(L, classes, in, nInst, s.order, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

}  // namespace ibm
