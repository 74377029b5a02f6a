// Tuned kernels, fifth generation: ONE CTA PER PARTICLE, asynchronous box staging (cp.async / LDGSTS).
//
// What the ncu captures of the earlier generations said (profiles/README.md):
//   v2 (direct gathers)   bound by the L1 tag stage: every lane of a warp-level gather reads a different cell of an SoA
//                         plane (14.6 distinct 128-B lines per request);
//   v3/v4 (staged boxes)  the gathers leave L1, but the box is staged with synchronous loads through registers
//                         (each thread stalls on a handful of loads at a time), three times per work item, and a d = 8
//                         sphere is TWO work items (128 + 25 shell cells) that each stage the whole box: 12 of 20 ms.
// This generation keeps the staged box and fixes how it is filled:
//   * one CTA owns one particle instance (all its shell cells, in chunks of F5_CELLS), so a box is staged once per field;
//   * the box is filled with cp.async: a thread issues all its 8-byte copies back to back and only then waits, nothing
//     passes through registers, and the second CTA resident on the SM computes while this one waits;
//   * the forcing is split BY FIELD, using the linearity of the gather (particleIBM.C:238-260):
//         phase 0  planes Uold:  tO -= G(x_I)/dt          tS -= G(x_S)/dt
//         phase 1  planes S:     tO += G(x_O)             tS += G(x_S)
//         phase 2  planes U:     tO += (v(x_O) - G(x_O))/dt   tS += v(x_S)/dt
//     so only THREE planes (53 KB for a 13^3 box) are resident at a time and two CTAs fit an SM; every shell cell
//     accumulates its three partial sums in registers and issues one atomicAdd per component, as before.
// Shape points come from the class's trig tables in shared memory (bit-identical to centeredMesh_ + centre, see
// ibm_tables.cuh); ellipses and classes whose box does not fit keep using the earlier generations (host check).
//
// The header is also compiled for the CPU emulation in tools/cuda_emu (tests/test_emu_v5.py), which runs the very same
// source as fibers and checks it against the oracle on machines without a GPU.
#pragma once
#include <cuda_pipeline.h>

#include "ibm_device.cuh"
#include "ibm_tables.cuh"

#ifdef IBM_CUDA_EMU
#define IBM_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(::emu::dyn_smem())
#else
#define IBM_DYN_SMEM(T, name)                                   \
  extern __shared__ __align__(16) unsigned char ibm_dyn_smem_raw[]; \
  T* name = reinterpret_cast<T*>(ibm_dyn_smem_raw)
#endif

namespace ibm {

#define F5_THREADS 256
#define F5_CELLS 160                    // shell cells per chunk (a d = 8 sphere has ~153)
#define F5_MAXCOL (4 * F5_CELLS + 1)    // distinct corner columns of a chunk, worst case
#define F5_HASH 1024
#define F5_BOX 2208                     // cells per staged plane: 13 x 13 x 13 = 2197, padded
#define S5_THREADS 256
#define S5_BOX SBOX4                    // stress: box of the surface layer, 12^3 padded (same as v4)

struct Box5 { int i0, j0, k0, bx, by, bz, px; };

// cells that can hold a point within `reach` of the centre, plus one stencil cell each side, clipped to the lattice
__device__ __forceinline__ Box5 box5_around(const LatticeDev& L, double cx, double cy, double cz, double reach) {
  Box5 B;
  const int i0 = (int)floor((cx - reach - L.ox) / L.dx) - 1, i1 = (int)floor((cx + reach - L.ox) / L.dx) + 1;
  const int j0 = (int)floor((cy - reach - L.oy) / L.dy) - 1, j1 = (int)floor((cy + reach - L.oy) / L.dy) + 1;
  const int k0 = (int)floor((cz - reach - L.oz) / L.dz) - 1 - L.kz0, k1 = (int)floor((cz + reach - L.oz) / L.dz) + 1 - L.kz0;
  B.i0 = max(i0, 0); B.j0 = max(j0, 0); B.k0 = max(k0, 0);
  B.bx = min(i1, L.nx - 1) - B.i0 + 1; B.by = min(j1, L.ny - 1) - B.j0 + 1; B.bz = min(k1, L.nz - 1) - B.k0 + 1;
  B.px = B.bx | 1;  // odd row pitch: neighbouring rows start in different banks
  return B;
}
__device__ __forceinline__ bool box5_fits(const Box5& B, int cap) { return B.bx > 0 && B.by > 0 && B.bz > 0 && B.px * B.by * B.bz <= cap; }

// asynchronous fill of nPlanes component planes of the box: 16 lanes walk a row, rows are walked incrementally
// (one integer division per thread); every copy is 8 bytes because a row may start on an odd cell
__device__ __forceinline__ void stage5_async(const LatticeDev& L, const Box5& B, const double* __restrict__ src, size_t compStride,
                                             int nPlanes, double* sm, int planeStride) {
  const int lane16 = threadIdx.x & 15, rowLane = threadIdx.x >> 4, rowStep = blockDim.x >> 4;
  const int rows = B.by * B.bz;
  int bj = rowLane % B.by, bk = rowLane / B.by;
  const int dj = rowStep % B.by, dk = rowStep / B.by;
  for (int r = rowLane; r < rows; r += rowStep) {
    const size_t g0 = (size_t)B.i0 + (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk));
    const int s0 = B.px * (bj + B.by * bk);
    for (int bi = lane16; bi < B.bx; bi += 16)
      for (int q = 0; q < nPlanes; ++q) __pipeline_memcpy_async(sm + q * planeStride + s0 + bi, src + q * compStride + g0 + bi, 8);
    bj += dj; bk += dk;
    if (bj >= B.by) { bj -= B.by; ++bk; }
  }
}

struct St5 { int self; int ok; double cSelf; double cNb[6]; };  // neighbour = self +- {1, px, px*by}; ok = existence bits

// 7-point gather stencil of a point, relative to the staged box: half the IDW-weighted face values
//   sum_f w_f/W * (phi_P + phi_N(f))/2      (particleIBMTmp.C:26-50 with the linear face value, lambda = 1/2)
// over the INTERNAL faces of the point's cell (particleShape.C:57-90).  false = point off the (local) mesh.
__device__ __forceinline__ bool make_st5(const LatticeDev& L, double idx, double idy, double idz, const Box5& B, const V3& p, St5& s) {
  int I, J, K;
  if (find_cell_fast(L, idx, idy, idz, p.x, p.y, p.z, I, J, K) < 0) return false;
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
  const int ok = (I + 1 < L.nx) | ((I > 0) << 1) | ((J + 1 < L.ny) << 2) | ((J > 0) << 3) | ((K + 1 < L.nz) << 4) | ((K > 0) << 5);
  s.self = (I - B.i0) + B.px * ((J - B.j0) + B.by * (K - B.k0));
  s.ok = ok;
  double w[6], W = 0.0;
  int hot = -1;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    const bool e = (ok >> d) & 1;
    w[d] = e ? rsqrt(d2[d]) : 0.0;
    if (e && d2[d] < IBM_SMALL * IBM_SMALL && hot < 0) hot = d;  // one-hot rule, particleShape.C:69-77
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
  return true;
}

// gather from a staged plane; a missing neighbour has coefficient 0 and reads the point's own cell (always staged)
__device__ __forceinline__ double gat5(const St5& s, const double* f, int px, int pxy) {
  const double* c = f + s.self;
  double v = s.cSelf * c[0];
  v = fma(s.cNb[0], c[(s.ok & 1) ? 1 : 0], v);
  v = fma(s.cNb[1], c[(s.ok & 2) ? -1 : 0], v);
  v = fma(s.cNb[2], c[(s.ok & 4) ? px : 0], v);
  v = fma(s.cNb[3], c[(s.ok & 8) ? -px : 0], v);
  v = fma(s.cNb[4], c[(s.ok & 16) ? pxy : 0], v);
  v = fma(s.cNb[5], c[(s.ok & 32) ? -pxy : 0], v);
  return v;
}

__device__ __forceinline__ unsigned hash5(int key) { return ((unsigned)key * 2654435761u) >> 22; }  // 10 bits = F5_HASH

// shared-memory footprint of the forcing kernel (host and device agree through this one expression)
__host__ __device__ __forceinline__ size_t f5_smem_bytes() {
  return (size_t)(3 * F5_BOX + 6 * F5_MAXCOL) * sizeof(double) + sizeof(Tab4) + (size_t)(F5_HASH + F5_MAXCOL + 1) * sizeof(int) +
         (size_t)F5_HASH * sizeof(unsigned short) + 16;
}

// ------------------------------------------------------------------------------------------------
// K2/K3 v5: direct forcing (particleCloud.C:1158-1208, particleIBM.C:210-264, :520-539), one CTA per instance
// ------------------------------------------------------------------------------------------------
// CACHE (variant 6): of the five stencil evaluations per column only three are distinct points - the surface layer is
// gathered for U_old and again for S, the outer layer for S and again for U.  With CACHE every thread keeps the stencil of
// each of its (at most F5_COLS_PER_THREAD) columns in a thread-private array from one field phase to the next (surface:
// U_old -> S, outer: S -> U; one slot per column, overwritten after its last use) and all three phases use the outer box,
// so that a cached cell index stays valid.  Same arithmetic, same results; it trades 40 % of the stencil fp64 work for
// 64 bytes of local-memory traffic per column and phase, and a slightly larger box in the U_old phase.
#define F5_COLS_PER_THREAD ((F5_MAXCOL + F5_THREADS - 1) / F5_THREADS)
template <bool CACHE>
__global__ void __launch_bounds__(F5_THREADS, 2)
k_forcing_v5(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
             const int* __restrict__ shellCell, const int* __restrict__ shellJK, const double* __restrict__ U,
             const double* __restrict__ Uold, const double* __restrict__ S, double dT, double* __restrict__ F, int* __restrict__ fallback) {
  IBM_DYN_SMEM(double, smem5);
  double* box = smem5;                                      // [3][F5_BOX]
  double* colVal = smem5 + 3 * F5_BOX;                      // [F5_MAXCOL][6]: tO partial (3), tS partial (3)
  Tab4& tab = *(Tab4*)(colVal + 6 * F5_MAXCOL);
  int* hKey = (int*)(&tab + 1);                             // [F5_HASH]
  int* colKey = hKey + F5_HASH;                             // [F5_MAXCOL]
  unsigned short* hSlot = (unsigned short*)(colKey + F5_MAXCOL + 1);  // [F5_HASH]
  __shared__ int nCol;
  if ((int)blockIdx.x >= nInst) return;
  const int b = order[blockIdx.x];
  if (in.onLocal[b] == 0) return;                           // particleIBM.C:218-221 (block-uniform)
  const int nS = in.nShell[b];
  if (nS == 0) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const size_t C = (size_t)L.nCells;
  const int nT = sc.nTheta;
  const int degenerateKey = nT * sc.nk;                     // App. B #3: bins with k == nk-1 use point 0 twelve times
  const double R = sc.d / 2.0;
  const Box5 Bo = box5_around(L, cx, cy, cz, sc.Rbound + sc.delta);  // outer + surface points
  const Box5 Bs = box5_around(L, cx, cy, cz, sc.Rbound);             // surface + inner points
  if (!tab4_ok(sc) || !box5_fits(Bo, F5_BOX) || !box5_fits(Bs, F5_BOX)) {  // the host routes such classes elsewhere
    if (threadIdx.x == 0) atomicExch(fallback, 1);
    return;
  }
  const double vx = in.v[0][b], vy = in.v[1][b], vz = in.v[2][b], wx = in.om[0][b], wy = in.om[1][b], wz = in.om[2][b];
  const double rdT = 1.0 / dT;
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
  const int off0 = in.shellOff[b];
  tab4_load(sc, tab);

#pragma unroll 1
  for (int c0 = 0; c0 < nS; c0 += F5_CELLS) {
    const int n = min(F5_CELLS, nS - c0);
    // the first field of the chunk starts loading now: the box is free (every gather of the previous chunk finished
    // before its last barrier) and the copies fly while the column hash and the scatter weights are built
    stage5_async(L, CACHE ? Bo : Bs, Uold, C, 3, box, F5_BOX);
    __pipeline_commit();
    for (int t = threadIdx.x; t < F5_HASH; t += blockDim.x) hKey[t] = -1;
    if (threadIdx.x == 0) nCol = 0;
    __syncthreads();  // hash cleared, tables loaded; the previous chunk has finished with colVal

    // ---- A: every shell cell registers the four corner columns (j,k) of its polar bin ----
    const bool live = (int)threadIdx.x < n;
    int myLat = 0, myKeys[4] = {0, 0, 0, 0};
    if (live) {
      myLat = shellCell[off0 + c0 + threadIdx.x];
      const int pk = shellJK[off0 + c0 + threadIdx.x];
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
        unsigned h = hash5(key);
        while (true) {
          const int old = atomicCAS(&hKey[h], -1, key);
          if (old == -1) {
            const int slot = atomicAdd(&nCol, 1);
            hSlot[h] = (unsigned short)slot;
            colKey[slot] = key;
            break;
          }
          if (old == key) break;
          h = (h + 1) & (F5_HASH - 1);
        }
      }
    }
    __syncthreads();
    const int ncol = nCol;

    // ---- scatter weights of my shell cell (particleShape.C:110-133): distances to the INNER points weigh the outer
    //      terms (App. B #7), distances to the surface points the surface terms; first point within SMALL is one-hot ----
    double wI[4] = {0, 0, 0, 0}, wS[4] = {0, 0, 0, 0}, Wtot = 0.0;
    int slot[4] = {0, 0, 0, 0};
    if (live) {
      const int I = myLat % L.nx, J = (myLat / L.nx) % L.ny, K = myLat / (L.nx * L.ny);
      const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
      int first = -1, firstS = -1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int key = myKeys[q];
        unsigned h = hash5(key);
        while (hKey[h] != key) h = (h + 1) & (F5_HASH - 1);
        slot[q] = hSlot[h];
        const bool deg = key == degenerateKey;
        const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
        const V3 xI = tab4_point(sc, tab, R, cx, cy, cz, 0, jq, kq);
        const V3 xS = tab4_point(sc, tab, R, cx, cy, cz, deg ? 0 : 1, jq, kq);
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
      for (int q = 0; q < 4; ++q) Wtot += wI[q] + wS[q];
    }

    double ax = 0.0, ay = 0.0, az = 0.0;
    St5 cache[CACHE ? F5_COLS_PER_THREAD : 1];  // (thread-private: lives in local memory, backed by L1 / L2)
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
      const double* fld = (phase == 0) ? Uold : (phase == 1 ? S : U);
      const Box5& B = (CACHE || phase != 0) ? Bo : Bs;
      const int pxy = B.px * B.by;
      if (phase > 0) {  // (phase 0 was issued at the top of the chunk)
        stage5_async(L, B, fld, C, 3, box, F5_BOX);
        __pipeline_commit();
      }
      __pipeline_wait_prior(0);
      __syncthreads();  // the three planes of this field are in shared memory

      int it = 0;
      for (int s = threadIdx.x; s < ncol; s += blockDim.x, ++it) {
        const int key = colKey[s];
        const bool deg = key == degenerateKey;
        const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
        double* cv = colVal + 6 * s;
        // layer of the point that plays the "outer" role (tO) and the "surface" role (tS) in this phase
        const int layO = deg ? 0 : (phase == 0 ? 0 : 2);
        const int layS = deg ? 0 : 1;
        St5 st;
        {  // tS partial (first: in the S phase its cached stencil is read before the slot takes the outer one)
          const V3 x = tab4_point(sc, tab, R, cx, cy, cz, layS, jq, kq);
          if (phase == 2) {
            const double rx = x.x - cx, ry = x.y - cy, rz = x.z - cz;
            cv[3] = (vx + (wy * rz - wz * ry)) * rdT; cv[4] = (vy + (wz * rx - wx * rz)) * rdT; cv[5] = (vz + (wx * ry - wy * rx)) * rdT;
          } else {
            bool on;
            if (CACHE && phase == 1) { st = cache[it]; on = st.self >= 0; }
            else {
              on = make_st5(L, idx, idy, idz, B, x, st);
              if (CACHE) { if (!on) { st = St5(); st.self = -1; } cache[it] = st; }
            }
            double g0 = 0.0, g1 = 0.0, g2 = 0.0;
            if (on) {
              g0 = gat5(st, box, B.px, pxy); g1 = gat5(st, box + F5_BOX, B.px, pxy); g2 = gat5(st, box + 2 * F5_BOX, B.px, pxy);
            }
            if (phase == 0) { cv[3] = -rdT * g0; cv[4] = -rdT * g1; cv[5] = -rdT * g2; }
            else { cv[3] = g0; cv[4] = g1; cv[5] = g2; }
          }
        }
        {  // tO partial
          const V3 x = tab4_point(sc, tab, R, cx, cy, cz, layO, jq, kq);
          bool on;
          if (CACHE && phase == 2) { st = cache[it]; on = st.self >= 0; }
          else {
            on = make_st5(L, idx, idy, idz, B, x, st);
            if (CACHE && phase == 1) { if (!on) { st = St5(); st.self = -1; } cache[it] = st; }
          }
          double g0 = 0.0, g1 = 0.0, g2 = 0.0;
          if (on) {
            g0 = gat5(st, box, B.px, pxy); g1 = gat5(st, box + F5_BOX, B.px, pxy); g2 = gat5(st, box + 2 * F5_BOX, B.px, pxy);
          }
          if (phase == 0) { cv[0] = -rdT * g0; cv[1] = -rdT * g1; cv[2] = -rdT * g2; }
          else if (phase == 1) { cv[0] = g0; cv[1] = g1; cv[2] = g2; }
          else {
            // v(pt) = v + omega x (pt - c)     (particleIBM.C:387-390), always present, also for a point off the mesh
            const double rx = x.x - cx, ry = x.y - cy, rz = x.z - cz;
            const double u0 = vx + (wy * rz - wz * ry), u1 = vy + (wz * rx - wx * rz), u2 = vz + (wx * ry - wy * rx);
            cv[0] = (u0 - g0) * rdT; cv[1] = (u1 - g1) * rdT; cv[2] = (u2 - g2) * rdT;
          }
        }
      }
      __syncthreads();  // column values of this phase complete

      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double* cv = colVal + 6 * slot[q];
          ax = fma(wI[q], cv[0], fma(wS[q], cv[3], ax));
          ay = fma(wI[q], cv[1], fma(wS[q], cv[4], ay));
          az = fma(wI[q], cv[2], fma(wS[q], cv[5], az));
        }
      }
      // no barrier here: the next writes to the box come from copies whose readers (the gather loop) all passed the
      // barrier above; colVal is next written after the barrier that follows the staging (or, in the next chunk, two
      // barriers later), which a thread only reaches after this accumulation; the hash arrays are not read in the phases.
    }

    if (live) {  // F[i] += Fi   (particleIBM.C:262)
      const double rW = 1.0 / Wtot;
      atomicAdd(&F[myLat], ax * rW);
      atomicAdd(&F[C + myLat], ay * rW);
      atomicAdd(&F[2 * C + myLat], az * rW);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v5: surface-stress integral (particleIBM.C:541-581), one CTA per instance in Morton order.
// Same arithmetic as k_stress_v4; the seven planes (p and the six devRhoReff components) of the surface-layer box are
// filled asynchronously, and the trig tables load while the copies are in flight.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t s5_smem_bytes() { return (size_t)7 * S5_BOX * sizeof(double) + sizeof(Tab4); }

__global__ void __launch_bounds__(S5_THREADS, 2)
k_stress_v5(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
            const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef, double* __restrict__ Tcorr,
            int* __restrict__ fallback) {
  IBM_DYN_SMEM(double, smem5);
  double* box = smem5;                        // [7][S5_BOX]: p, then the six devRhoReff components
  Tab4& tab = *(Tab4*)(smem5 + 7 * S5_BOX);
  __shared__ double red[S5_THREADS / 32][9];
  if ((int)blockIdx.x >= nInst) return;
  const int b = order[blockIdx.x];
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const size_t C = (size_t)L.nCells;
  const int nSurf = sc.nTheta * sc.nk;
  const bool on = in.onLocal[b] != 0;
  const double R = sc.d / 2.0;
  const Box5 B = box5_around(L, cx, cy, cz, sc.Rbound);
  const bool fits = tab4_ok(sc) && box5_fits(B, S5_BOX);
  // An instance without local points (another slab's instance on a multi-GPU run) does no work but MUST write zero
  // partials: its F/T slots hold last step's combined totals and the cross-rank sum adds whatever is in them.
  if (on && !fits && threadIdx.x == 0) atomicExch(fallback, 1);
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (on && fits) {
    stage5_async(L, B, P, C, 1, box, S5_BOX);
    stage5_async(L, B, Tau, C, 6, box + S5_BOX, S5_BOX);
    __pipeline_commit();
    tab4_load(sc, tab);
    double rot[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) rot[q] = in.rot[9 * (size_t)b + q];
    const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
    const int pxy = B.px * B.by;
    const double sfScale = (sc.kind == 0) ? (R * R) * sc.dTheta : R * sc.dTheta * sc.l;
    __pipeline_wait_prior(0);
    __syncthreads();
    for (int s = threadIdx.x; s < nSurf; s += blockDim.x) {
      const int i = s / sc.nk, k = s - i * sc.nk;
      const int i2 = i + sc.nTheta * k;
      const V3 pt = tab4_point(sc, tab, R, cx, cy, cz, 1, i, k);
      double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
      St5 st;
      if (make_st5(L, idx, idy, idz, B, pt, st)) {
#pragma unroll
        for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gat5(st, box + (1 + cmp) * S5_BOX, B.px, pxy) * rhoRef;
        pp = (gat5(st, box, B.px, pxy) - 2.0 * st.cSelf * pRefOverRho) * rhoRef;  // (p_f - pRef/rho) rho, particleCloud.C:1213-1227
      }
      const V3 m2 = tab4_point_id(sc, tab, R, cx, cy, cz, i2);  // App. B #5: lever arm as coded (index2 into mesh_)
      const double Rx = m2.x - cx, Ry = m2.y - cy, Rz = m2.z - cz;
      const double Rcx = pt.x - cx, Rcy = pt.y - cy, Rcz = pt.z - cz;
      double sx, sy, sz;  // Sf_[index2(i,k)]   (sphere.C:134-157, cylinder.C:116-122)
      if (sc.kind == 0) {
        const double f = sfScale * tab.ck[k];
        sx = -(tab.cosT[i] * tab.sinP[k]) * f; sy = -(tab.sinT[i] * tab.sinP[k]) * f; sz = -tab.cosP[k] * f;
      } else {
        sx = -tab.cosT[i] * sfScale; sy = -tab.sinT[i] * sfScale; sz = 0.0;
      }
      const double Sx = rot[0] * sx + rot[1] * sy + rot[2] * sz, Sy = rot[3] * sx + rot[4] * sy + rot[5] * sz,
                   Sz = rot[6] * sx + rot[7] * sy + rot[8] * sz;
      const double rmag = rsqrt(Sx * Sx + Sy * Sy + Sz * Sz);
      const double nx = Sx * rmag, ny = Sy * rmag, nz = Sz * rmag;
      const double fx = pp * sx + (tau[0] * Sx + tau[1] * Sy + tau[2] * Sz);  // App. B #6: unrotated Sf_ in the pressure term
      const double fy = pp * sy + (tau[1] * Sx + tau[3] * Sy + tau[4] * Sz);
      const double fz = pp * sz + (tau[2] * Sx + tau[4] * Sy + tau[5] * Sz);
      const double fn = fx * nx + fy * ny + fz * nz;
      const double ftx = fx - fn * nx, fty = fy - fn * ny, ftz = fz - fn * nz;
      acc[0] += fx; acc[1] += fy; acc[2] += fz;
      acc[3] -= fty * Rz - ftz * Ry; acc[4] -= ftz * Rx - ftx * Rz; acc[5] -= ftx * Ry - fty * Rx;
      acc[6] -= fty * Rcz - ftz * Rcy; acc[7] -= ftz * Rcx - ftx * Rcz; acc[8] -= ftx * Rcy - fty * Rcx;
    }
  }
  // fixed reduction tree: shuffle inside a warp, then the warps in index order
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
    for (int w = 0; w < S5_THREADS / 32; ++w) v += red[w][threadIdx.x];
    const int q = threadIdx.x;
    if (q < 3) in.F[q][b] = v;
    else if (q < 6) in.T[q - 3][b] = v;
    else if (Tcorr) Tcorr[3 * b + (q - 6)] = v;
  }
}

}  // namespace ibm
