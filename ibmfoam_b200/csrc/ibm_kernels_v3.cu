// Tuned kernels, third generation: shared-memory staging of the work item's bounding box.
//
// ncu on the second generation showed the L1 tag stage saturated (14.6 distinct 128-byte lines per warp-level
// gather: every lane reads a different cell of a structure-of-arrays plane).  Here the cell data a work item
// needs is staged into shared memory with row-coalesced loads and every gather is served from shared memory.
//
// Forcing.  The gather is linear in the field, so the three layers of shape points only need
//     inner   points:  B = Uold                    -> contributes  -g_B(x_I)/dt   to tO
//     outer   points:  A = S - U/dt                -> tO = v(x_O)/dt + g_A(x_O) - g_B(x_I)/dt
//     surface points:  C = S - Uold/dt             -> tS = v(x_S)/dt + g_C(x_S)
// (tO, tS as in ibm_kernels_v2.cu / particleIBM.C:238-260).  A pass stages the three components of one combined
// field for the item's box (pitch padded to an odd number of doubles: conflict-free row pairs), then one thread
// per distinct column evaluates that layer's 7-point stencil from shared memory.  Three passes, 3 planes resident.
//
// Stress.  One CTA per instance; p and the six devRhoReff components of the surface-layer box are staged once
// (7 planes), one thread per surface point, fixed-tree reduction.
// Items whose box exceeds the shared-memory budget fall back to the second-generation (global gather) path.
#include "ibm_kernels.h"

namespace ibm {

#define FHASH 1024
#define FMAXCOL (4 * FORCING_CHUNK + 1)
#define FBOXCAP 2560            // cells per staged plane (forcing): 14 x 13 x 13 padded
#define SBOXCAP 1792            // cells per staged plane (stress):  12 x 12 x 12 padded
#define F3_THREADS 256
#define S3_THREADS 256

struct BoxS {
  int i0, j0, k0;     // lattice origin of the box
  int bx, by, bz;     // extents
  int px;             // odd pitch >= bx
};

struct St7 {
  int self;           // index inside the staged box (or lattice index on the fallback path), -1 = off mesh
  int nb[6];
  double cSelf;
  double cNb[6];
};

// 7-point stencil of one shape point.  STAGED: indices address the shared-memory box; else the lattice.
template <bool STAGED>
__device__ __forceinline__ void make_st7(const LatticeDev& L, double idx, double idy, double idz, const BoxS& bxs, const V3& p, St7& s) {
  int I, J, K;
  int lat = find_cell_fast(L, idx, idy, idz, p.x, p.y, p.z, I, J, K);
  s.self = lat;
  if (lat < 0) return;
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
  int step[6];
  if (STAGED) {
    s.self = (I - bxs.i0) + bxs.px * ((J - bxs.j0) + bxs.by * (K - bxs.k0));
    step[0] = 1; step[1] = -1; step[2] = bxs.px; step[3] = -bxs.px; step[4] = bxs.px * bxs.by; step[5] = -bxs.px * bxs.by;
  } else {
    step[0] = 1; step[1] = -1; step[2] = L.nx; step[3] = -L.nx; step[4] = L.nx * L.ny; step[5] = -L.nx * L.ny;
  }
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

__device__ __forceinline__ double gat(const St7& s, const double* f) {
  double v = s.cSelf * f[s.self];
#pragma unroll
  for (int d = 0; d < 6; ++d)
    if (s.nb[d] >= 0) v = fma(s.cNb[d], f[s.nb[d]], v);
  return v;
}
__device__ __forceinline__ double gatg(const St7& s, const double* __restrict__ f) {
  double v = s.cSelf * __ldg(f + s.self);
#pragma unroll
  for (int d = 0; d < 6; ++d)
    if (s.nb[d] >= 0) v = fma(s.cNb[d], __ldg(f + s.nb[d]), v);
  return v;
}

__device__ __forceinline__ V3 rigid_v3(const V3& pt, double cx, double cy, double cz, const V3& om, const V3& v) {
  double rx = pt.x - cx, ry = pt.y - cy, rz = pt.z - cz;
  return V3{v.x + (om.y * rz - om.z * ry), v.y + (om.z * rx - om.x * rz), v.z + (om.x * ry - om.y * rx)};
}
__device__ __forceinline__ unsigned hash_col3(int key) { return ((unsigned)key * 2654435761u) >> 22; }
__device__ __forceinline__ int find_pos3(const int* __restrict__ off, int n, int item) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= item) lo = mid; else hi = mid;
  }
  return lo;
}

// box bounds of a set of points: per-thread min/max folded into shared memory with atomics
__device__ __forceinline__ void fold_cell(int* sb, int I, int J, int K) {
  atomicMin(&sb[0], I); atomicMin(&sb[1], J); atomicMin(&sb[2], K);
  atomicMax(&sb[3], I); atomicMax(&sb[4], J); atomicMax(&sb[5], K);
}

// stage one combined vector field (3 components) of the box:  out = a*X + b*Y  per component (Y may be null)
__device__ __forceinline__ void stage3(const LatticeDev& L, const BoxS& B, const double* __restrict__ X, double a,
                                       const double* __restrict__ Y, double b, double* sm, int planeStride) {
  const size_t C = (size_t)L.nCells;
  const int rows = B.by * B.bz;
  // 16 lanes per row (box widths up to 16), two rows per warp; wider boxes take several sweeps of a row
  const int lane16 = threadIdx.x & 15, rowLane = threadIdx.x >> 4, rowStep = blockDim.x >> 4;
  for (int r = rowLane; r < rows; r += rowStep) {
    const int bj = r % B.by, bk = r / B.by;
    const size_t g0 = (size_t)B.i0 + (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk));
    const int s0 = B.px * (bj + B.by * bk);
    for (int bi = lane16; bi < B.bx; bi += 16) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        double v = a * __ldg(X + q * C + g0 + bi);
        if (Y) v = fma(b, __ldg(Y + q * C + g0 + bi), v);
        sm[q * planeStride + s0 + bi] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(F3_THREADS, 2)
k_forcing_v3(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
             const int* __restrict__ chunkOff, const int* __restrict__ shellCell, const int* __restrict__ shellJK,
             const double* __restrict__ U, const double* __restrict__ Uold, const double* __restrict__ S, double dT,
             double* __restrict__ F) {
  extern __shared__ double smem[];
  double* box = smem;                                     // [3][FBOXCAP]
  double* colVal = smem + 3 * FBOXCAP;                    // [FMAXCOL][6]
  int* hKey = (int*)(colVal + FMAXCOL * 6);               // [FHASH]
  int* colKey = hKey + FHASH;                             // [FMAXCOL]
  unsigned short* hSlot = (unsigned short*)(colKey + FMAXCOL + 1);  // [FHASH]
  __shared__ int nCol;
  __shared__ int sb[6];
  const int item = blockIdx.x;
  const int pos = find_pos3(chunkOff, nInst, item);
  const int b = order[pos];
  const int c0 = (item - chunkOff[pos]) * FORCING_CHUNK;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const int off = in.shellOff[b] + c0;
  const int n = min(FORCING_CHUNK, in.nShell[b] - c0);
  const size_t C = (size_t)L.nCells;
  const int nT = sc.nTheta;
  const int degenerateKey = nT * sc.nk;  // App. B #3
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;

  for (int t = threadIdx.x; t < FHASH; t += blockDim.x) hKey[t] = -1;
  if (threadIdx.x == 0) { nCol = 0; sb[0] = sb[1] = sb[2] = 0x7fffffff; sb[3] = sb[4] = sb[5] = -1; }
  __syncthreads();

  // ---- A: register the corner columns of every shell cell (two threads per cell: corners {0,1} and {2,3}) ----
  const int cellT = threadIdx.x >> 1, half = threadIdx.x & 1;
  const bool live = cellT < n;
  int myLat = 0, myKeys[2] = {0, 0};
  if (live) {
    myLat = shellCell[off + cellT];
    const int pk = shellJK[off + cellT];
    const int j = pk & 0xffff, k = pk >> 16;
    const int jp1 = (j + 1) % nT;
    if (sc.nk != 1 && k == sc.nk - 1) {
      myKeys[0] = myKeys[1] = degenerateKey;
    } else {
      const int kk = (sc.nk != 1 && half) ? k + 1 : k;
      myKeys[0] = j + nT * kk; myKeys[1] = jp1 + nT * kk;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int key = myKeys[q];
      unsigned h = hash_col3(key);
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
  const int ncol = nCol;

  // ---- A2: exact bounding box of the cells the item's shape points live in ----
  {
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    for (int s = threadIdx.x; s < ncol; s += blockDim.x) {
      const int key = colKey[s];
      const int pI = (key == degenerateKey) ? 0 : pt_index(nT, 0, key % nT, key / nT);
      const int np = (key == degenerateKey) ? 1 : 3;
      for (int l = 0; l < np; ++l) {
        const V3 x = shape_point(sc, rot, cx, cy, cz, pI + l);
        int I, J, K;
        if (find_cell_fast(L, idx, idy, idz, x.x, x.y, x.z, I, J, K) >= 0) {
          lo[0] = min(lo[0], I); lo[1] = min(lo[1], J); lo[2] = min(lo[2], K);
          hi[0] = max(hi[0], I); hi[1] = max(hi[1], J); hi[2] = max(hi[2], K);
        }
      }
    }
    if (hi[0] >= 0) { atomicMin(&sb[0], lo[0]); atomicMin(&sb[1], lo[1]); atomicMin(&sb[2], lo[2]);
                      atomicMax(&sb[3], hi[0]); atomicMax(&sb[4], hi[1]); atomicMax(&sb[5], hi[2]); }
  }
  __syncthreads();
  BoxS B;
  bool staged = false, anyOn = sb[3] >= 0;
  if (anyOn) {
    B.i0 = max(sb[0] - 1, 0); B.j0 = max(sb[1] - 1, 0); B.k0 = max(sb[2] - 1, 0);
    B.bx = min(sb[3] + 1, L.nx - 1) - B.i0 + 1; B.by = min(sb[4] + 1, L.ny - 1) - B.j0 + 1; B.bz = min(sb[5] + 1, L.nz - 1) - B.k0 + 1;
    B.px = B.bx | 1;
    staged = B.px * B.by * B.bz <= FBOXCAP;
  } else {
    B.i0 = B.j0 = B.k0 = 0; B.bx = B.by = B.bz = 1; B.px = 1;
  }

  const V3 vp{in.v[0][b], in.v[1][b], in.v[2][b]}, om{in.om[0][b], in.om[1][b], in.om[2][b]};
  const double rdT = 1.0 / dT;

  // ---- B: three passes (inner / outer / surface) ----
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    // pass 0: B = Uold (inner points); pass 1: A = S - U/dt (outer); pass 2: C = S - Uold/dt (surface)
    if (staged) {
      if (pass == 0) stage3(L, B, Uold, 1.0, nullptr, 0.0, box, FBOXCAP);
      else if (pass == 1) stage3(L, B, S, 1.0, U, -rdT, box, FBOXCAP);
      else stage3(L, B, S, 1.0, Uold, -rdT, box, FBOXCAP);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < ncol; s += blockDim.x) {
      const int key = colKey[s];
      const int pI = (key == degenerateKey) ? 0 : pt_index(nT, 0, key % nT, key / nT);
      const int layer = (pass == 0) ? 0 : (pass == 1 ? 2 : 1);
      const int pid = (key == degenerateKey) ? 0 : pI + layer;
      const V3 x = shape_point(sc, rot, cx, cy, cz, pid);
      St7 st;
      double g[3] = {0.0, 0.0, 0.0};
      if (staged) {
        make_st7<true>(L, idx, idy, idz, B, x, st);
        if (st.self >= 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) g[q] = gat(st, box + q * FBOXCAP);
        }
      } else {
        make_st7<false>(L, idx, idy, idz, B, x, st);
        if (st.self >= 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (pass == 0) g[q] = gatg(st, Uold + q * C);
            else if (pass == 1) g[q] = gatg(st, S + q * C) - rdT * gatg(st, U + q * C);
            else g[q] = gatg(st, S + q * C) - rdT * gatg(st, Uold + q * C);
          }
        }
      }
      double* cv = colVal + 6 * s;
      if (pass == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) cv[q] = -rdT * g[q];
      } else {
        const V3 vv = rigid_v3(x, cx, cy, cz, om, vp);
        const double v3[3] = {vv.x, vv.y, vv.z};
        if (pass == 1) {
#pragma unroll
          for (int q = 0; q < 3; ++q) cv[q] += fma(v3[q], rdT, g[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 3; ++q) cv[3 + q] = fma(v3[q], rdT, g[q]);
        }
      }
    }
    __syncthreads();
  }

  // ---- C: two threads per shell cell (corners {0,1} / {2,3}), combined with one shuffle ----
  {
    double Fx = 0.0, Fy = 0.0, Fz = 0.0, W = 0.0;
    int hotq = 0x7fffffff;  // lowest q in 0..7 whose distance is below SMALL (inner q, surface 4+q)
    double wI[2] = {0, 0}, wS[2] = {0, 0};
    int slot[2] = {0, 0};
    if (live) {
      const int I = myLat % L.nx, J = (myLat / L.nx) % L.ny, K = myLat / (L.nx * L.ny);
      const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int key = myKeys[q];
        unsigned h = hash_col3(key);
        while (hKey[h] != key) h = (h + 1) & (FHASH - 1);
        slot[q] = hSlot[h];
        const int pI = (key == degenerateKey) ? 0 : pt_index(nT, 0, key % nT, key / nT);
        const int pS = (key == degenerateKey) ? 0 : pI + 1;
        const V3 xI = shape_point(sc, rot, cx, cy, cz, pI);
        const V3 xS = shape_point(sc, rot, cx, cy, cz, pS);
        const double dI2 = (Cx - xI.x) * (Cx - xI.x) + (Cy - xI.y) * (Cy - xI.y) + (Cz - xI.z) * (Cz - xI.z);
        const double dS2 = (Cx - xS.x) * (Cx - xS.x) + (Cy - xS.y) * (Cy - xS.y) + (Cz - xS.z) * (Cz - xS.z);
        wI[q] = rsqrt(dI2);
        wS[q] = rsqrt(dS2);
        const int gq = 2 * half + q;  // corner number 0..3
        if (dI2 < IBM_SMALL * IBM_SMALL) hotq = min(hotq, gq);
        if (dS2 < IBM_SMALL * IBM_SMALL) hotq = min(hotq, 4 + gq);
      }
    }
    hotq = min(hotq, __shfl_xor_sync(0xffffffffu, hotq, 1));
    if (hotq != 0x7fffffff) {  // one-hot rule of particleShape.C:121-127
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int gq = 2 * half + q;
        wI[q] = (hotq == gq) ? 1.0 : 0.0;
        wS[q] = (hotq == 4 + gq) ? 1.0 : 0.0;
      }
    }
    if (live) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        W += wI[q] + wS[q];
        const double* cv = colVal + 6 * slot[q];
        Fx = fma(wI[q], cv[0], fma(wS[q], cv[3], Fx));
        Fy = fma(wI[q], cv[1], fma(wS[q], cv[4], Fy));
        Fz = fma(wI[q], cv[2], fma(wS[q], cv[5], Fz));
      }
    }
    W += __shfl_xor_sync(0xffffffffu, W, 1);
    Fx += __shfl_xor_sync(0xffffffffu, Fx, 1);
    Fy += __shfl_xor_sync(0xffffffffu, Fy, 1);
    Fz += __shfl_xor_sync(0xffffffffu, Fz, 1);
    if (live && half == 0) {
      const double rW = 1.0 / W;
      atomicAdd(&F[myLat], Fx * rW);
      atomicAdd(&F[C + myLat], Fy * rW);
      atomicAdd(&F[2 * C + myLat], Fz * rW);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v3
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S3_THREADS, 2)
k_stress_v3(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
            const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef, double* __restrict__ Tcorr) {
  extern __shared__ double smem[];
  double* box = smem;  // [7][SBOXCAP]: p, then the six devRhoReff components
  __shared__ double red[8][9];
  __shared__ int sb[6];
  if ((int)blockIdx.x >= nInst) return;
  const int b = order[blockIdx.x];
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const size_t C = (size_t)L.nCells;
  const int nS = sc.nTheta * sc.nk;
  const bool on = in.onLocal[b] != 0;
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
  if (threadIdx.x == 0) { sb[0] = sb[1] = sb[2] = 0x7fffffff; sb[3] = sb[4] = sb[5] = -1; }
  __syncthreads();
  if (on) {
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    for (int s = threadIdx.x; s < nS; s += blockDim.x) {
      const V3 x = shape_point(sc, rot, cx, cy, cz, pt_index(sc.nTheta, 1, s / sc.nk, s % sc.nk));
      int I, J, K;
      if (find_cell_fast(L, idx, idy, idz, x.x, x.y, x.z, I, J, K) >= 0) {
        lo[0] = min(lo[0], I); lo[1] = min(lo[1], J); lo[2] = min(lo[2], K);
        hi[0] = max(hi[0], I); hi[1] = max(hi[1], J); hi[2] = max(hi[2], K);
      }
    }
    if (hi[0] >= 0) { atomicMin(&sb[0], lo[0]); atomicMin(&sb[1], lo[1]); atomicMin(&sb[2], lo[2]);
                      atomicMax(&sb[3], hi[0]); atomicMax(&sb[4], hi[1]); atomicMax(&sb[5], hi[2]); }
  }
  __syncthreads();
  BoxS B;
  bool staged = false;
  if (sb[3] >= 0) {
    B.i0 = max(sb[0] - 1, 0); B.j0 = max(sb[1] - 1, 0); B.k0 = max(sb[2] - 1, 0);
    B.bx = min(sb[3] + 1, L.nx - 1) - B.i0 + 1; B.by = min(sb[4] + 1, L.ny - 1) - B.j0 + 1; B.bz = min(sb[5] + 1, L.nz - 1) - B.k0 + 1;
    B.px = B.bx | 1;
    staged = B.px * B.by * B.bz <= SBOXCAP;
  } else {
    B.i0 = B.j0 = B.k0 = 0; B.bx = B.by = B.bz = 1; B.px = 1;
  }
  if (staged) {
    const int rows = B.by * B.bz;
    const int lane16 = threadIdx.x & 15, rowLane = threadIdx.x >> 4, rowStep = blockDim.x >> 4;
    for (int r = rowLane; r < rows; r += rowStep) {
      const int bj = r % B.by, bk = r / B.by;
      const size_t g0 = (size_t)B.i0 + (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk));
      const int s0 = B.px * (bj + B.by * bk);
      for (int bi = lane16; bi < B.bx; bi += 16) {
        box[s0 + bi] = __ldg(P + g0 + bi);
#pragma unroll
        for (int q = 0; q < 6; ++q) box[(1 + q) * SBOXCAP + s0 + bi] = __ldg(Tau + q * C + g0 + bi);
      }
    }
  }
  __syncthreads();
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < nS; s += blockDim.x) {
    const int i = s / sc.nk, k = s % sc.nk;
    const int i2 = i + sc.nTheta * k;
    const V3 pt = shape_point(sc, rot, cx, cy, cz, pt_index(sc.nTheta, 1, i, k));
    double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
    if (on) {
      St7 st;
      if (staged) {
        make_st7<true>(L, idx, idy, idz, B, pt, st);
        if (st.self >= 0) {
#pragma unroll
          for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gat(st, box + (1 + cmp) * SBOXCAP) * rhoRef;
          pp = (gat(st, box) - 2.0 * st.cSelf * pRefOverRho) * rhoRef;
        }
      } else {
        make_st7<false>(L, idx, idy, idz, B, pt, st);
        if (st.self >= 0) {
#pragma unroll
          for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gatg(st, Tau + cmp * C) * rhoRef;
          pp = (gatg(st, P) - 2.0 * st.cSelf * pRefOverRho) * rhoRef;
        }
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

static size_t f3_smem() {
  return (size_t)(3 * FBOXCAP + FMAXCOL * 6) * sizeof(double) + (size_t)(FHASH + FMAXCOL + 1) * sizeof(int) + FHASH * sizeof(unsigned short) + 16;
}
static size_t s3_smem() { return (size_t)7 * SBOXCAP * sizeof(double); }

void launch_forcing_v3(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, int totalChunks, const int* shellCell, const int* shellJK, const double* U,
                       const double* Uold, const double* S, double dT, double* F) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(k_forcing_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f3_smem());
  if (nInst > 0 && totalChunks > 0)
    k_forcing_v3<<<totalChunks, F3_THREADS, f3_smem(), st>>>(L, classes, in, nInst, s.order, s.chunkOff, shellCell, shellJK, U, Uold, S,
                                                              dT, F);
}
void launch_stress_v3(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(k_stress_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3_smem());
  if (nInst > 0) k_stress_v3<<<nInst, S3_THREADS, s3_smem(), st>>>(L, classes, in, nInst, s.order, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

}  // namespace ibm
