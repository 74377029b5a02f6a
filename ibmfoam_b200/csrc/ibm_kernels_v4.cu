// Tuned kernels, fourth generation: shared-memory staging done leanly.
//
// Changes against ibm_kernels_v3.cu, each answering a line of its ncu profile (profiles/r1_lines_*_v3_*.txt):
//   * the class's trig tables (cos/sin of theta_j, sin/cos of phi_k) live in shared memory and shape points are
//     evaluated from them with the reference's exact products ((r*cos(theta))*sin(phi), sphere.C:184-190) — no
//     template loads from global (was 20 % of the stall samples) and bit-identical point coordinates;
//   * the staged box is an analytic superset (no extra pass over all points, no barrier for it);
//   * staging walks rows incrementally (no integer division; was 11.6 % of the instructions);
//   * the rotation matrix and the per-instance constants sit in registers.
// Shapes whose tables do not fit (nTheta > 256 or nk > 128) and ellipses take the v3 path.
#include <cstdlib>

#include "ibm_kernels.h"
#include "ibm_tables.cuh"

namespace ibm {

struct Box4 { int i0, j0, k0, bx, by, bz, px; };

// cells that can hold a point within `reach` of the centre, plus one stencil cell, clipped to the lattice
__device__ __forceinline__ Box4 box4_around(const LatticeDev& L, double cx, double cy, double cz, double reach) {
  Box4 B;
  const int i0 = (int)floor((cx - reach - L.ox) / L.dx) - 1, i1 = (int)floor((cx + reach - L.ox) / L.dx) + 1;
  const int j0 = (int)floor((cy - reach - L.oy) / L.dy) - 1, j1 = (int)floor((cy + reach - L.oy) / L.dy) + 1;
  const int k0 = (int)floor((cz - reach - L.oz) / L.dz) - 1 - L.kz0, k1 = (int)floor((cz + reach - L.oz) / L.dz) + 1 - L.kz0;
  B.i0 = max(i0, 0); B.j0 = max(j0, 0); B.k0 = max(k0, 0);
  B.bx = min(i1, L.nx - 1) - B.i0 + 1; B.by = min(j1, L.ny - 1) - B.j0 + 1; B.bz = min(k1, L.nz - 1) - B.k0 + 1;
  B.px = B.bx | 1;
  return B;
}

struct St4 { int self; int ok; double cSelf; double cNb[6]; };  // neighbour = self +- {1, px, px*by}; ok = existence bits

// 7-point gather stencil of one point in the staged box.  Returns false when the point is off the (local) mesh.
__device__ __forceinline__ bool make_st4(const LatticeDev& L, double idx, double idy, double idz, const Box4& B, const V3& p, St4& s) {
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
    if (e && d2[d] < IBM_SMALL * IBM_SMALL && hot < 0) hot = d;  // particleShape.C:69-77
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

// gather from a staged plane; a missing neighbour has coefficient 0 and reads the point's own cell (always valid)
__device__ __forceinline__ double gat4(const St4& s, const double* f, int px, int pxy) {
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

// stage nPlanes component planes of the box; rows are walked incrementally (no division), 16 lanes per row
__device__ __forceinline__ void stage4(const LatticeDev& L, const Box4& B, const double* __restrict__ src, size_t compStride, int nPlanes,
                                       double* sm, int planeStride) {
  const int lane16 = threadIdx.x & 15, rowLane = threadIdx.x >> 4, rowStep = blockDim.x >> 4;
  const int rows = B.by * B.bz;
  int bj = rowLane % B.by, bk = rowLane / B.by;  // one division per thread, then incremental
  const int dj = rowStep % B.by, dk = rowStep / B.by;
  for (int r = rowLane; r < rows; r += rowStep) {
    const size_t g0 = (size_t)B.i0 + (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk));
    const int s0 = B.px * (bj + B.by * bk);
    for (int bi = lane16; bi < B.bx; bi += 16)
      for (int q = 0; q < nPlanes; ++q) sm[q * planeStride + s0 + bi] = __ldg(src + q * compStride + g0 + bi);
    bj += dj; bk += dk;
    if (bj >= B.by) { bj -= B.by; ++bk; }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v4: surface-stress integral (particleIBM.C:541-581), one CTA per instance in Morton order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S4_THREADS, 2)
k_stress_v4(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
            const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef, double* __restrict__ Tcorr,
            int* __restrict__ fallback) {
  extern __shared__ double smem4[];
  double* box = smem4;                        // [7][SBOX4]: p, then the six devRhoReff components
  Tab4& tab = *(Tab4*)(smem4 + 7 * SBOX4);
  __shared__ double red[S4_THREADS / 32][9];
  if ((int)blockIdx.x >= nInst) return;
  const int b = order[blockIdx.x];
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const size_t C = (size_t)L.nCells;
  const int nS = sc.nTheta * sc.nk;
  const bool on = in.onLocal[b] != 0;
  const double R = sc.d / 2.0;
  const Box4 B = box4_around(L, cx, cy, cz, sc.Rbound);
  const bool fits = tab4_ok(sc) && B.bx > 0 && B.by > 0 && B.bz > 0 && B.px * B.by * B.bz <= SBOX4;
  // Instances without local points (other slabs' instances on a multi-GPU run) do no work but MUST still write
  // zero partials below: their F/T slots hold last step's combined totals, and the cross-rank sum adds whatever is
  // in them.  `fits` is guaranteed by the host for every instance that is on the local mesh (ibm_set_particles).
  if (on && !fits) {
    if (threadIdx.x == 0) atomicExch(fallback, 1);
  }
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (on && fits) {
    tab4_load(sc, tab);
    stage4(L, B, P, C, 1, box, SBOX4);
    stage4(L, B, Tau, C, 6, box + SBOX4, SBOX4);
    __syncthreads();
    double rot[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) rot[q] = in.rot[9 * (size_t)b + q];
    const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
    const int pxy = B.px * B.by;
    const double sfScale = (sc.kind == 0) ? (R * R) * sc.dTheta : R * sc.dTheta * sc.l;
    for (int s = threadIdx.x; s < nS; s += blockDim.x) {
      const int i = s / sc.nk, k = s - i * sc.nk;
      const int i2 = i + sc.nTheta * k;
      const V3 pt = tab4_point(sc, tab, R, cx, cy, cz, 1, i, k);
      double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
      St4 st;
      if (make_st4(L, idx, idy, idz, B, pt, st)) {
#pragma unroll
        for (int cmp = 0; cmp < 6; ++cmp) tau[cmp] = gat4(st, box + (1 + cmp) * SBOX4, B.px, pxy) * rhoRef;
        pp = (gat4(st, box, B.px, pxy) - 2.0 * st.cSelf * pRefOverRho) * rhoRef;
      }
      const V3 m2 = tab4_point_id(sc, tab, R, cx, cy, cz, i2);  // App. B #5: lever arm as coded (index2 into mesh_)
      const double Rx = m2.x - cx, Ry = m2.y - cy, Rz = m2.z - cz;
      const double Rcx = pt.x - cx, Rcy = pt.y - cy, Rcz = pt.z - cz;
      // Sf_[index2(i,k)]: -(cos(theta) sin(phi), sin(theta) sin(phi), cos(phi)) r^2 dTheta c_k   (sphere.C:134-157)
      //                   -(cos(theta), sin(theta), 0) r dTheta l                                  (cylinder.C:116-122)
      double sx, sy, sz;
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
    for (int w = 0; w < S4_THREADS / 32; ++w) v += red[w][threadIdx.x];
    const int q = threadIdx.x;
    if (q < 3) in.F[q][b] = v;
    else if (q < 6) in.T[q - 3][b] = v;
    else if (Tcorr) Tcorr[3 * b + (q - 6)] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K2/K3 v4: direct forcing, one CTA per work item (<= FORCING_CHUNK shell cells of one instance)
// ------------------------------------------------------------------------------------------------
#define FBOX4 2560
#define F4_THREADS 256
#define F4_HASH 1024
#define F4_MAXCOL (4 * FORCING_CHUNK + 1)

// stage the three components of  a*X + b*Y  (Y may be null) for the box
__device__ __forceinline__ void stage4_combo(const LatticeDev& L, const Box4& B, const double* __restrict__ X, double a,
                                             const double* __restrict__ Y, double b, double* sm) {
  const size_t C = (size_t)L.nCells;
  const int lane16 = threadIdx.x & 15, rowLane = threadIdx.x >> 4, rowStep = blockDim.x >> 4;
  const int rows = B.by * B.bz;
  int bj = rowLane % B.by, bk = rowLane / B.by;
  const int dj = rowStep % B.by, dk = rowStep / B.by;
  for (int r = rowLane; r < rows; r += rowStep) {
    const size_t g0 = (size_t)B.i0 + (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk));
    const int s0 = B.px * (bj + B.by * bk);
    for (int bi = lane16; bi < B.bx; bi += 16) {
      const size_t g = g0 + bi;
      if (Y) {
        const double x0 = __ldg(X + g), x1 = __ldg(X + C + g), x2 = __ldg(X + 2 * C + g);
        const double y0 = __ldg(Y + g), y1 = __ldg(Y + C + g), y2 = __ldg(Y + 2 * C + g);
        sm[s0 + bi] = fma(b, y0, a * x0); sm[FBOX4 + s0 + bi] = fma(b, y1, a * x1); sm[2 * FBOX4 + s0 + bi] = fma(b, y2, a * x2);
      } else {
        const double x0 = __ldg(X + g), x1 = __ldg(X + C + g), x2 = __ldg(X + 2 * C + g);
        sm[s0 + bi] = a * x0; sm[FBOX4 + s0 + bi] = a * x1; sm[2 * FBOX4 + s0 + bi] = a * x2;
      }
    }
    bj += dj; bk += dk;
    if (bj >= B.by) { bj -= B.by; ++bk; }
  }
}

__device__ __forceinline__ unsigned hash4(int key) { return ((unsigned)key * 2654435761u) >> 22; }

__global__ void __launch_bounds__(F4_THREADS, 2)
k_forcing_v4(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order,
             const int* __restrict__ chunkOff, const int* __restrict__ itemPos, const int* __restrict__ shellCell,
             const int* __restrict__ shellJK, const double* __restrict__ U, const double* __restrict__ Uold,
             const double* __restrict__ S, double dT, double* __restrict__ F, int dbg) {
  extern __shared__ double smem4[];
  double* box = smem4;                                      // [3][FBOX4]
  double* colVal = smem4 + 3 * FBOX4;                       // [F4_MAXCOL][6]
  Tab4& tab = *(Tab4*)(colVal + F4_MAXCOL * 6);
  int* hKey = (int*)(&tab + 1);                             // [F4_HASH]
  int* colKey = hKey + F4_HASH;                             // [F4_MAXCOL]
  unsigned short* hSlot = (unsigned short*)(colKey + F4_MAXCOL + 1);  // [F4_HASH]
  __shared__ int nCol;
  const int item = blockIdx.x;
  const int pos = itemPos[item];
  const int b = order[pos];
  const int c0 = (item - chunkOff[pos]) * FORCING_CHUNK;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const int off = in.shellOff[b] + c0;
  const int n = min(FORCING_CHUNK, in.nShell[b] - c0);
  const size_t C = (size_t)L.nCells;
  const int nT = sc.nTheta;
  const int degenerateKey = nT * sc.nk;  // App. B #3
  const double R = sc.d / 2.0;

  for (int t = threadIdx.x; t < F4_HASH; t += blockDim.x) hKey[t] = -1;
  if (threadIdx.x == 0) nCol = 0;
  tab4_load(sc, tab);
  __syncthreads();

  // ---- A: register the corner columns (two threads per shell cell: corners {0,1} and {2,3}) ----
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
      unsigned h = hash4(key);
      while (true) {
        const int old = atomicCAS(&hKey[h], -1, key);
        if (old == -1) {
          const int slot = atomicAdd(&nCol, 1);
          hSlot[h] = (unsigned short)slot;
          colKey[slot] = key;
          break;
        }
        if (old == key) break;
        h = (h + 1) & (F4_HASH - 1);
      }
    }
  }
  const Box4 B = box4_around(L, cx, cy, cz, sc.Rbound + sc.delta);
  const int pxy = B.px * B.by;
  const double vx = in.v[0][b], vy = in.v[1][b], vz = in.v[2][b], ox = in.om[0][b], oy = in.om[1][b], oz = in.om[2][b];
  const double rdT = 1.0 / dT;
  const double idx = 1.0 / L.dx, idy = 1.0 / L.dy, idz = 1.0 / L.dz;
  __syncthreads();
  const int ncol = nCol;

  // ---- B: three passes; pass 0 inner points with B = Uold, pass 1 outer points with A = S - U/dt,
  //         pass 2 surface points with C = S - Uold/dt (linearity of the gather) ----
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    if (dbg != 1) {
      if (pass == 0) stage4_combo(L, B, Uold, 1.0, nullptr, 0.0, box);
      else if (pass == 1) stage4_combo(L, B, S, 1.0, U, -rdT, box);
      else stage4_combo(L, B, S, 1.0, Uold, -rdT, box);
    }
    __syncthreads();
    if (dbg == 2) continue;
    const int layer = (pass == 0) ? 0 : (pass == 1 ? 2 : 1);
    for (int s = threadIdx.x; s < ncol; s += blockDim.x) {
      const int key = colKey[s];
      const bool deg = key == degenerateKey;
      const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
      const V3 x = tab4_point(sc, tab, R, cx, cy, cz, deg ? 0 : layer, jq, kq);
      St4 st;
      double g0 = 0.0, g1 = 0.0, g2 = 0.0;
      if (make_st4(L, idx, idy, idz, B, x, st)) {
        g0 = gat4(st, box, B.px, pxy); g1 = gat4(st, box + FBOX4, B.px, pxy); g2 = gat4(st, box + 2 * FBOX4, B.px, pxy);
      }
      double* cv = colVal + 6 * s;
      if (pass == 0) {
        cv[0] = -rdT * g0; cv[1] = -rdT * g1; cv[2] = -rdT * g2;
      } else {
        // v(pt) = v + omega x (pt - c)     (particleIBM.C:387-390)
        const double rx = x.x - cx, ry = x.y - cy, rz = x.z - cz;
        const double w0 = vx + (oy * rz - oz * ry), w1 = vy + (oz * rx - ox * rz), w2 = vz + (ox * ry - oy * rx);
        if (pass == 1) { cv[0] += fma(w0, rdT, g0); cv[1] += fma(w1, rdT, g1); cv[2] += fma(w2, rdT, g2); }
        else { cv[3] = fma(w0, rdT, g0); cv[4] = fma(w1, rdT, g1); cv[5] = fma(w2, rdT, g2); }
      }
    }
    __syncthreads();
  }

  // ---- C: scatter, two threads per shell cell ----
  {
    double Fx = 0.0, Fy = 0.0, Fz = 0.0, W = 0.0;
    int hotq = 0x7fffffff;
    double wI[2] = {0, 0}, wS[2] = {0, 0};
    int slot[2] = {0, 0};
    if (live) {
      const int I = myLat % L.nx, J = (myLat / L.nx) % L.ny, K = myLat / (L.nx * L.ny);
      const double Cx = L.ox + (I + 0.5) * L.dx, Cy = L.oy + (J + 0.5) * L.dy, Cz = L.oz + (K + L.kz0 + 0.5) * L.dz;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int key = myKeys[q];
        unsigned h = hash4(key);
        while (hKey[h] != key) h = (h + 1) & (F4_HASH - 1);
        slot[q] = hSlot[h];
        const bool deg = key == degenerateKey;
        const int kq = deg ? 0 : key / nT, jq = deg ? 0 : key - kq * nT;
        const V3 xI = tab4_point(sc, tab, R, cx, cy, cz, 0, jq, kq);
        const V3 xS = tab4_point(sc, tab, R, cx, cy, cz, deg ? 0 : 1, jq, kq);
        const double dI2 = (Cx - xI.x) * (Cx - xI.x) + (Cy - xI.y) * (Cy - xI.y) + (Cz - xI.z) * (Cz - xI.z);
        const double dS2 = (Cx - xS.x) * (Cx - xS.x) + (Cy - xS.y) * (Cy - xS.y) + (Cz - xS.z) * (Cz - xS.z);
        wI[q] = rsqrt(dI2);
        wS[q] = rsqrt(dS2);
        const int gq = 2 * half + q;
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

static size_t f4_smem() {
  return (size_t)(3 * FBOX4 + F4_MAXCOL * 6) * sizeof(double) + sizeof(Tab4) + (size_t)(F4_HASH + F4_MAXCOL + 1) * sizeof(int) +
         F4_HASH * sizeof(unsigned short) + 16;
}

void launch_forcing_v4(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, int totalChunks, const int* shellCell, const int* shellJK, const double* U,
                       const double* Uold, const double* S, double dT, double* F) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(k_forcing_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f4_smem());
  if (nInst > 0 && totalChunks > 0)
  {
    static int dbg = getenv("IBM_V4_DBG") ? atoi(getenv("IBM_V4_DBG")) : 0;  // profiling switch: 1 = skip staging, 2 = skip gathers
    k_forcing_v4<<<totalChunks, F4_THREADS, f4_smem(), st>>>(L, classes, in, nInst, s.order, s.chunkOff, s.itemPos, shellCell, shellJK, U,
                                                              Uold, S, dT, F, dbg);
  }
}

static size_t s4_smem() { return (size_t)7 * SBOX4 * sizeof(double) + sizeof(Tab4); }

void launch_stress_v4(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr, int* dFallback) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(k_stress_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s4_smem());
  if (nInst > 0)
    k_stress_v4<<<nInst, S4_THREADS, s4_smem(), st>>>(L, classes, in, nInst, s.order, P, Tau, pRefOverRho, rhoRef, Tcorr, dFallback);
}

}  // namespace ibm
