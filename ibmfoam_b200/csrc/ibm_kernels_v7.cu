// Tuned kernels, seventh generation (loops 2+3 and 4): one CTA per particle instance, the cells the instance can
// touch staged ONCE into shared memory as ball rows, gathers served from shared memory.
//
// What the round-1 profile of k_forcing_v2 said (profiles/r1_lines_forcing_v2_big512_final.txt): the kernel is bound by
// the L1 tag stage — 105 scattered 8-byte loads per column, 14.6 distinct 128-byte lines per warp-level load, ~2 cycles per
// line.  The staged generations v3-v5 lost because they staged the bounding CUBE (2.7x the cells that are needed) with
// 8-byte LDGSTS (8 cycles each) and evaluated every stencil once per field.  Here:
//
//   * the region is the exact set of lattice rows (x-runs) that a layer of shape points can reach: for every (y,z) row
//     of the bounding box, the x-range of cells whose box intersects the ball of the layer's radius ("own" range), widened
//     by the six face neighbours (own range +-1, united with the own ranges of the four neighbouring rows) — a superset of
//     every cell a 7-point gather stencil of that layer reads (particleShape.C:49-66), ~45 % of the cube;
//   * the gather is linear (particleIBMTmp.C:26-50), so the three fields are combined WHILE staging into the three
//     plane-sets the scatter needs (particleIBM.C:238-260, App. B #7):
//         A = S - U/dt      on the reach of the outer layer    (gathered at outer points)
//         B = S - Uold/dt   on the reach of the surface layer  (gathered at surface points)
//         C =   - Uold/dt   on the reach of the inner layer    (gathered at inner points)
//     so that  tO = v(x_O)/dt + g_O(A) + g_I(C),  tS = v(x_S)/dt + g_S(B):  63 shared-memory loads per column instead of
//     105 global ones, and every global value is read exactly once per instance with coalesced row loads;
//   * the radial point triples ("columns") referenced by the shell cells are de-duplicated through a direct map over
//     (j,k) and a block scan (no hash, no atomics), each column's three stencils are evaluated once;
//   * shape points come from the class's trig tables in shared memory with the reference's exact products
//     (sphere.C:184-190): bit-identical coordinates, hence bit-identical point->cell membership.
//
// Instances the fast path cannot take (ellipses, classes whose tables or region exceed the shared-memory budget, a
// region or column count above the launch's capacities) are appended to a device list and processed by the direct-gather
// kernel of the second generation in the same stream (k_forcing_v2_list): same results, no host round trip.
//
// Arithmetic differs from the reference-order kernel only by re-association (<= 1e-12 relative, tested).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ibm_kernels.h"

namespace ibm {

#define V7_MAXT 320

#if defined(IBM_CUDA_EMU)
#define IBM_V7_GUARD 1  // the CPU emulation checks that every gather stencil stays inside the staged region
#endif

// shared-memory carve-up, identical on host and device
struct V7Layout {
  int planes, colVal, tab, rows, colKey, colSlot, elemRow, total;  // byte offsets
};
__host__ __device__ inline V7Layout v7_layout(const V7Caps& c) {
  V7Layout l;
  int o = 0;
  l.planes = o; o += 8 * c.planeCap;
  l.colVal = o; o += 8 * 6 * c.colCap;
  l.tab = o; o += 8 * (2 * c.nTcap + 3 * c.nkCap);
  l.rows = o; o += 4 * 10 * c.rowsCap;
  l.colKey = o; o += 4 * c.colCap;
  l.colSlot = o; o += 2 * ((c.mapCap + 3) & ~3);
  l.elemRow = o; o += 2 * ((c.elemCap + 3) & ~3);
  l.total = (o + 15) & ~15;
  return l;
}

struct Tab7 {
  const double *cosT, *sinT, *sinP, *cosP, *ck;
};
__device__ __forceinline__ Tab7 tab7_load(const ShapeClassDev& sc, double* w, const V7Caps& cap) {
  for (int j = threadIdx.x; j < sc.nTheta; j += blockDim.x) { w[j] = sc.cosT[j]; w[cap.nTcap + j] = sc.sinT[j]; }
  double* p = w + 2 * cap.nTcap;
  for (int k = threadIdx.x; k < sc.nk; k += blockDim.x) { p[k] = sc.sinP[k]; p[cap.nkCap + k] = sc.cosP[k]; p[2 * cap.nkCap + k] = sc.ck[k]; }
  return Tab7{w, w + cap.nTcap, p, p + cap.nkCap, p + 2 * cap.nkCap};
}

// mesh_[index(layer,j,k)] of a sphere / cylinder: the same roundings as the host template (sphere.C:184-190, cylinder.C:143-148)
__device__ __forceinline__ V3 tab7_point(int kind, const Tab7& t, double r, double cx, double cy, double cz, int j, int k) {
  if (kind == 0)
    return V3{xadd(xmul(xmul(r, t.cosT[j]), t.sinP[k]), cx), xadd(xmul(xmul(r, t.sinT[j]), t.sinP[k]), cy), xadd(xmul(r, t.cosP[k]), cz)};
  return V3{xadd(xmul(r, t.cosT[j]), cx), xadd(xmul(r, t.sinT[j]), cy), xadd(0.0, cz)};
}

// 1/sqrt(x) and 1/x for well-scaled positive x: hardware seed (2^-22) + one third-order / two second-order steps; no
// special-case handling (the arguments are squared distances and weight sums of order delta^2 and 1/delta)
__device__ __forceinline__ double rsqrt7(double x) {
#if defined(IBM_CUDA_EMU)
  return 1.0 / sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);     // 1 - x y^2
  const double p = fma(e, 0.375, 0.5);
  return fma(p * e, y, y);                  // y (1 + e/2 + 3 e^2/8)
#endif
}
__device__ __forceinline__ double rcp7(double x) {
#if defined(IBM_CUDA_EMU)
  return 1.0 / x;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
#if !defined(IBM_CUDA_EMU)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

struct Box7 {
  int j0, k0;  // first row / plane of the box, clipped to the lattice
  int by, bz;  // rows per plane, planes
  int i0, bx;  // first column and width of the bounding box (prefetch only)
  int edge;    // the unclipped box reaches the lattice boundary: stencil neighbours may be missing, points may be off the mesh
};

struct Geo7 {  // lattice constants of a launch, in registers
  double ox, oy, oz, dx, dy, dz, idx, idy, idz, hx, hy, hz;
  int nx, ny, nz, kz0, ownK0, ownK1;
};
__device__ __forceinline__ Geo7 geo7(const LatticeDev& L) {
  return Geo7{L.ox, L.oy, L.oz, L.dx, L.dy, L.dz, 1.0 / L.dx, 1.0 / L.dy, 1.0 / L.dz, 0.5 * L.dx, 0.5 * L.dy, 0.5 * L.dz, L.nx, L.ny, L.nz, L.kz0, L.ownK0, L.ownK1};
}

// exclusive scan of two ints per thread over the block; every thread gets the block totals
template <int NT>
__device__ __forceinline__ void block_scan7(int& v0, int& v1, int* wsum, int& tot0, int& tot1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int i0 = v0, i1 = v1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n0 = __shfl_up_sync(0xffffffffu, i0, o), n1 = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { i0 += n0; i1 += n1; }
  }
  if (lane == 31) { wsum[2 * warp] = i0; wsum[2 * warp + 1] = i1; }
  __syncthreads();
  int b0 = 0, b1 = 0, t0 = 0, t1 = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const int s0 = wsum[2 * w], s1 = wsum[2 * w + 1];
    if (w < warp) { b0 += s0; b1 += s1; }
    t0 += s0; t1 += s1;
  }
  tot0 = t0; tot1 = t1;
  __syncthreads();  // wsum may be reused
  v0 = b0 + i0 - v0; v1 = b1 + i1 - v1;
}

// own x-range of row (J,K) for a layer of radius rho around (cx,cy,cz): cells whose box intersects the closed ball
__device__ __forceinline__ void own_range7(const LatticeDev& L, double idx, double cx, double cy, double cz, double rho, int J, int K, int& lo, int& hi) {
  const double ylo = L.oy + J * L.dy, yhi = L.oy + (J + 1) * L.dy;
  const double zlo = L.oz + (K + L.kz0) * L.dz, zhi = L.oz + (K + L.kz0 + 1) * L.dz;
  const double ey = fmax(fmax(ylo - cy, cy - yhi), 0.0), ez = fmax(fmax(zlo - cz, cz - zhi), 0.0);
  const double re = rho * (1.0 + 1e-9) + 1e-9 * L.dx;  // rounding of the point coordinates, of the cell search and of the products below
  const double h2 = re * re - ey * ey - ez * ez;
  if (h2 < 0.0) { lo = 1; hi = 0; return; }
  const double h = sqrt(h2);
  lo = (int)floor((cx - h - L.ox) * idx);
  hi = (int)floor((cx + h - L.ox) * idx);
}

// cell index along one axis and the position inside the cell (fraction of the cell size).  floor(q * inv) equals
// floor(fl(q / d)) (the DEFINED findCell) unless q/d lies within rounding of an integer; those cases take the exact division.
__device__ __forceinline__ int cell_axis7(double q, double d, double inv, double& fr) {
  const double t = q * inv;
  const double s = (t - 0.5) + 6755399441055744.0;  // 1.5 * 2^52: the low word is rn(t - 0.5) as a two's-complement integer
  int I = __double2loint(s);
  const double f = s - 6755399441055744.0;
  fr = t - f;
  if (!(fr > 1e-9 && fr < 1.0 - 1e-9 && fabs(t) < 1.0e9)) {
    const double f2 = floor(xdiv(q, d));
    I = (f2 > -2.0e9 && f2 < 2.0e9) ? (int)f2 : -0x40000000;  // (far outside or not a number: off the mesh)
    fr = t - f2;
  }
  return I;
}

struct St7 { int idx[7]; double c[7]; };  // self, +x, -x, +y, -y, +z, -z: record index in a plane-set, coefficient

// 7-point gather stencil of point p in a plane-set with row table rowBase (particleShape.C:49-91 with the linear face
// weights lambda = 1/2 folded in).  Returns false when the point is off the local mesh (it contributes nothing).
template <bool EDGE>
__device__ __forceinline__ bool make_st7(const Geo7& G, const Box7& B, const int* __restrict__ rowBase, const int* __restrict__ rowLo,
                                         const int* __restrict__ rowHi, const V3& p, St7& s, int& guard, const bool ownedOnly = false) {
  double fx, fy, fz;
  const int I = cell_axis7(p.x - G.ox, G.dx, G.idx, fx);
  const int J = cell_axis7(p.y - G.oy, G.dy, G.idy, fy);
  const int K = cell_axis7(p.z - G.oz, G.dz, G.idz, fz) - G.kz0;
  if (EDGE) {
    if ((unsigned)I >= (unsigned)G.nx || (unsigned)J >= (unsigned)G.ny || (unsigned)K >= (unsigned)G.nz) return false;
  }
  // the stress integral with a halo (ibm_set_slab_halo): a surface point belongs to the rank that owns the plane of its cell
  if (ownedOnly && (K < G.ownK0 || K >= G.ownK1)) return false;
  const int bj = J - B.j0, bk = K - B.k0;
  const int row = bj + B.by * bk;
#if defined(IBM_V7_GUARD)
  if (bj < 0 || bj >= B.by || bk < 0 || bk >= B.bz) { guard = 1; return false; }
  if (rowLo[row] > I - (I > 0) || rowHi[row] < I + (I + 1 < G.nx)) { guard = 1; return false; }
  if (J + 1 < G.ny && (bj + 1 >= B.by || I < rowLo[row + 1] || I > rowHi[row + 1])) { guard = 1; return false; }
  if (J > 0 && (bj < 1 || I < rowLo[row - 1] || I > rowHi[row - 1])) { guard = 1; return false; }
  if (K + 1 < G.nz && (bk + 1 >= B.bz || I < rowLo[row + B.by] || I > rowHi[row + B.by])) { guard = 1; return false; }
  if (K > 0 && (bk < 1 || I < rowLo[row - B.by] || I > rowHi[row - B.by])) { guard = 1; return false; }
  if (!EDGE && !(I > 0 && I + 1 < G.nx && J > 0 && J + 1 < G.ny && K > 0 && K + 1 < G.nz)) { guard = 1; return false; }
#endif
  const double ex = (0.5 - fx) * G.dx, ey = (0.5 - fy) * G.dy, ez = (0.5 - fz) * G.dz;  // cell centre - point
  const double ex2 = ex * ex, ey2 = ey * ey, ez2 = ez * ez;
  const double exy = ex2 + ey2, exz = ex2 + ez2, eyz = ey2 + ez2;
  double d2[6];
  d2[0] = fma(ex + G.hx, ex + G.hx, eyz);
  d2[1] = fma(ex - G.hx, ex - G.hx, eyz);
  d2[2] = fma(ey + G.hy, ey + G.hy, exz);
  d2[3] = fma(ey - G.hy, ey - G.hy, exz);
  d2[4] = fma(ez + G.hz, ez + G.hz, exy);
  d2[5] = fma(ez - G.hz, ez - G.hz, exy);
  const int self = rowBase[row] + I;
  double w[6];
  if (EDGE) {
    const bool ok[6] = {I + 1 < G.nx, I > 0, J + 1 < G.ny, J > 0, K + 1 < G.nz, K > 0};
    s.idx[0] = self;
    s.idx[1] = ok[0] ? self + 1 : self;
    s.idx[2] = ok[1] ? self - 1 : self;
    s.idx[3] = ok[2] ? rowBase[row + 1] + I : self;
    s.idx[4] = ok[3] ? rowBase[row - 1] + I : self;
    s.idx[5] = ok[4] ? rowBase[row + B.by] + I : self;
    s.idx[6] = ok[5] ? rowBase[row - B.by] + I : self;
#pragma unroll
    for (int d = 0; d < 6; ++d) w[d] = ok[d] ? rsqrt7(d2[d]) : 0.0;
  } else {
    s.idx[0] = self; s.idx[1] = self + 1; s.idx[2] = self - 1;
    s.idx[3] = rowBase[row + 1] + I; s.idx[4] = rowBase[row - 1] + I;
    s.idx[5] = rowBase[row + B.by] + I; s.idx[6] = rowBase[row - B.by] + I;
#pragma unroll
    for (int d = 0; d < 6; ++d) w[d] = rsqrt7(d2[d]);
  }
  double W = ((w[0] + w[1]) + (w[2] + w[3])) + (w[4] + w[5]);
  if (!(W <= 0.5e15)) {  // a face centre within SMALL of the point (the seeded rsqrt7 turns d2 = 0 into NaN: caught here too)
    int hot = -1;        // one-hot on the first such face (particleShape.C:69-77)
    const bool ok6[6] = {!EDGE || I + 1 < G.nx, !EDGE || I > 0, !EDGE || J + 1 < G.ny, !EDGE || J > 0, !EDGE || K + 1 < G.nz, !EDGE || K > 0};
    W = 0.0;
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      w[d] = ok6[d] ? rsqrt(d2[d]) : 0.0;  // library version: exact handling of 0
      if (ok6[d] && d2[d] < IBM_SMALL * IBM_SMALL && hot < 0) hot = d;
      W += w[d];
    }
    if (hot >= 0) {
#pragma unroll
      for (int d = 0; d < 6; ++d) w[d] = (d == hot) ? 1.0 : 0.0;
      W = 1.0;
    }
  }
  const double hW = (!EDGE || W > 0.0) ? 0.5 * rcp7(W) : 0.0;
#pragma unroll
  for (int d = 0; d < 6; ++d) s.c[1 + d] = w[d] * hW;
  s.c[0] = ((s.c[1] + s.c[2]) + (s.c[3] + s.c[4])) + (s.c[5] + s.c[6]);
  return true;
}

// gather NC component planes (component q of cell idx at f[q * n + idx])
template <int NC>
__device__ __forceinline__ void gat7_planes(const St7& s, const double* __restrict__ f, int n, double* out) {
#pragma unroll
  for (int q = 0; q < NC; ++q) out[q] = s.c[0] * f[q * n + s.idx[0]];
#pragma unroll
  for (int d = 1; d < 7; ++d) {
#pragma unroll
    for (int q = 0; q < NC; ++q) out[q] = fma(s.c[d], f[q * n + s.idx[d]], out[q]);
  }
}

// gather NC interleaved components of a record plane (record = NC doubles per cell)
template <int NC>
__device__ __forceinline__ void gat7(const St7& s, const double* __restrict__ f, double* out) {
#pragma unroll
  for (int q = 0; q < NC; ++q) out[q] = s.c[0] * f[NC * s.idx[0] + q];
#pragma unroll
  for (int d = 1; d < 7; ++d) {
#pragma unroll
    for (int q = 0; q < NC; ++q) out[q] = fma(s.c[d], f[NC * s.idx[d] + q], out[q]);
  }
}

// Region tables of one instance: for every set (0 = outer reach, 1 = surface reach, 2 = inner reach) and row of the box
//   rowLo/rowHi = first / last staged column, rowBase = record index of column 0 of the row (record = rowBase + I).
// Cell counts per set in n[3].  first/last limit the sets (the stress kernel needs set 1 only).  rows <= NT (caller).
template <int NT>
__device__ __forceinline__ void build_region7(const LatticeDev& L, const Box7& B, double cx, double cy, double cz, const double* rho,
                                              int first, int last, int rowsCap, int* rowBase, int* rowLo, int* rowHi, int* rowG, int* wsum, int* n) {
  const int rows = B.by * B.bz;
  const int r = threadIdx.x;
  const int bj = r % B.by, bk = r / B.by;
  const double idx = 1.0 / L.dx;
  if (r < rows) {
    for (int s = first; s <= last; ++s) {
      int lo, hi;
      own_range7(L, idx, cx, cy, cz, rho[s], B.j0 + bj, B.k0 + bk, lo, hi);
      rowLo[s * rowsCap + r] = lo;
      rowHi[s * rowsCap + r] = hi;
    }
    rowG[r] = L.nx * ((B.j0 + bj) + L.ny * (B.k0 + bk));  // lattice index of column 0 of the row
  }
  __syncthreads();
  int tl[3] = {1, 1, 1}, th[3] = {0, 0, 0};
  if (r < rows) {
    for (int s = first; s <= last; ++s) {
      const int* lo = rowLo + s * rowsCap;
      const int* hi = rowHi + s * rowsCap;
      int a = 0x7fffffff, b = -0x7fffffff;
      if (lo[r] <= hi[r]) { a = lo[r] - 1; b = hi[r] + 1; }
      if (bj + 1 < B.by && lo[r + 1] <= hi[r + 1]) { a = min(a, lo[r + 1]); b = max(b, hi[r + 1]); }
      if (bj > 0 && lo[r - 1] <= hi[r - 1]) { a = min(a, lo[r - 1]); b = max(b, hi[r - 1]); }
      if (bk + 1 < B.bz && lo[r + B.by] <= hi[r + B.by]) { a = min(a, lo[r + B.by]); b = max(b, hi[r + B.by]); }
      if (bk > 0 && lo[r - B.by] <= hi[r - B.by]) { a = min(a, lo[r - B.by]); b = max(b, hi[r - B.by]); }
      a = max(a, 0); b = min(b, L.nx - 1);
      if (a > b) { a = 1; b = 0; }
      tl[s] = a; th[s] = b;
    }
  }
  __syncthreads();
  // one scan for all three sets: set 0 alone, sets 1 and 2 packed (16 bits each; the caller guarantees < 65536 cells per set)
  int v0 = max(th[0] - tl[0] + 1, 0), v1 = max(th[1] - tl[1] + 1, 0) | (max(th[2] - tl[2] + 1, 0) << 16);
  int t0, t1;
  block_scan7<NT>(v0, v1, wsum, t0, t1);
  if (r < rows) {
    const int off[3] = {v0, v1 & 0xffff, (v1 >> 16) & 0xffff};
    for (int s = first; s <= last; ++s) {
      rowLo[s * rowsCap + r] = tl[s];
      rowHi[s * rowsCap + r] = th[s];
      rowBase[s * rowsCap + r] = off[s] - tl[s];
    }
  }
  n[0] = t0; n[1] = t1 & 0xffff; n[2] = (t1 >> 16) & 0xffff;
  __syncthreads();
}

__device__ __forceinline__ bool box7_around(const LatticeDev& L, double cx, double cy, double cz, double reach, Box7& B) {
  const double re = reach * (1.0 + 1e-9) + 1e-9 * L.dx;  // (covers the rounding of the products with the inverse spacings)
  const double ix = 1.0 / L.dx, iy = 1.0 / L.dy, iz = 1.0 / L.dz;
  const int i0 = (int)floor((cx - re - L.ox) * ix) - 1, i1 = (int)floor((cx + re - L.ox) * ix) + 1;
  const int j0 = (int)floor((cy - re - L.oy) * iy) - 1, j1 = (int)floor((cy + re - L.oy) * iy) + 1;
  const int k0 = (int)floor((cz - re - L.oz) * iz) - 1 - L.kz0, k1 = (int)floor((cz + re - L.oz) * iz) + 1 - L.kz0;
  B.j0 = max(j0, 0); B.k0 = max(k0, 0);
  B.by = min(j1, L.ny - 1) - B.j0 + 1;
  B.bz = min(k1, L.nz - 1) - B.k0 + 1;
  B.i0 = max(i0, 0); B.bx = min(i1, L.nx - 1) - B.i0 + 1;
  // interior: every point's cell AND its six neighbours exist (the box carries one cell of margin around the points' cells)
  B.edge = !(i0 >= 0 && i1 <= L.nx - 1 && j0 >= 0 && j1 <= L.ny - 1 && k0 >= 0 && k1 <= L.nz - 1);
  return B.by > 0 && B.bz > 0;
}

// ------------------------------------------------------------------------------------------------
// Launch records (V7Hdr + packed row tables), once per update
// ------------------------------------------------------------------------------------------------
#define V7_F_LOCAL 1
#define V7_F_FITS 2
#define V7_F_EDGE 4

// 64 threads per instance: the kernel is a chain of dependent loads and barriers (latency, not work), so the point is to
// have many instances in flight per SM (32 CTAs of two warps)
#define K7P_NT 64
#define K7P_ROWS 256
__global__ void __launch_bounds__(K7P_NT)
k7_prepare(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const int* __restrict__ order, V7Caps cap,
           V7Hdr* __restrict__ hdr, unsigned* __restrict__ rowsOut, int rowsCapG) {
  __shared__ short lo0[3][K7P_ROWS], hi0[3][K7P_ROWS];  // own ranges, relative to the box
  __shared__ short lo1[3][K7P_ROWS], hi1[3][K7P_ROWS];  // widened by the face neighbours
  __shared__ int offs[3][K7P_ROWS];
  __shared__ int laneTot[3][32];
  if ((int)blockIdx.x >= nInst) return;
  const int b = order[blockIdx.x];
  const int cls = in.cls[b];
  const ShapeClassDev sc = classes[cls];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  {
    // The outer layer cannot reach the owned planes (other ranks' instances on a multi-GPU run): the kernels only need to know
    // that.  The test of the marking kernel (mark_finish), so that this kernel needs nothing the marking kernel writes and
    // can run beside it; k7_finish adds the shell segment and clears the flag of instances without local points.
    double rOut = sc.Rbound + sc.delta;
    rOut += 1e-9 * (fabs(cx) + fabs(cy) + fabs(cz) + rOut) + 1e-300;
    const bool outside = cx + rOut < L.ox || cx - rOut > L.ox + L.nx * L.dx || cy + rOut < L.oy || cy - rOut > L.oy + L.ny * L.dy ||
                         cz + rOut < own_zlo(L) || cz - rOut > own_zhi(L);
    if (outside) {
      if (threadIdx.x == 0) {
        V7Hdr h{};
        h.b = b; h.cls = cls;
        h.cx = cx; h.cy = cy; h.cz = cz;
        hdr[blockIdx.x] = h;
      }
      return;
    }
  }
  const double R = sc.d / 2.0;
  const double rho[3] = {xadd(R, sc.delta), xadd(R, 0.0), xadd(R, -1 * sc.delta)};  // set 0 = outer reach ... set 2 = inner reach
  Box7 B;
  B.j0 = B.k0 = B.by = B.bz = B.i0 = B.bx = B.edge = 0;
  bool fits = sc.kind != 2 && sc.nTheta <= cap.nTcap && sc.nk <= cap.nkCap && box7_around(L, cx, cy, cz, rho[0], B);
  const int rows = B.by * B.bz;
  if (fits) fits = rows <= rowsCapG && rows <= K7P_ROWS && B.bx <= 250;
  int nS[3] = {0, 0, 0};
  if (fits) {  // (block-uniform)
    const double idx = 1.0 / L.dx;
    for (int r = threadIdx.x; r < rows; r += K7P_NT) {
      const int bk = r / B.by, bj = r - bk * B.by;
      for (int s = 0; s < 3; ++s) {
        int lo, hi;
        own_range7(L, idx, cx, cy, cz, rho[s], B.j0 + bj, B.k0 + bk, lo, hi);
        // (a range is only compared and widened by one: clamping it to the box's neighbourhood keeps it in 16 bits)
        lo = max(lo, B.i0 - 2); hi = min(hi, B.i0 + B.bx + 1);
        if (lo > hi) { lo = B.i0 + 1; hi = B.i0; }
        lo0[s][r] = (short)(lo - B.i0); hi0[s][r] = (short)(hi - B.i0);
      }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += K7P_NT) {
      const int bk = r / B.by, bj = r - bk * B.by;
      for (int s = 0; s < 3; ++s) {
        const short* lo = lo0[s];
        const short* hi = hi0[s];
        int a = 0x7fff, e = -0x7fff;
        if (lo[r] <= hi[r]) { a = lo[r] - 1; e = hi[r] + 1; }
        if (bj + 1 < B.by && lo[r + 1] <= hi[r + 1]) { a = min(a, (int)lo[r + 1]); e = max(e, (int)hi[r + 1]); }
        if (bj > 0 && lo[r - 1] <= hi[r - 1]) { a = min(a, (int)lo[r - 1]); e = max(e, (int)hi[r - 1]); }
        if (bk + 1 < B.bz && lo[r + B.by] <= hi[r + B.by]) { a = min(a, (int)lo[r + B.by]); e = max(e, (int)hi[r + B.by]); }
        if (bk > 0 && lo[r - B.by] <= hi[r - B.by]) { a = min(a, (int)lo[r - B.by]); e = max(e, (int)hi[r - B.by]); }
        a = max(a, -B.i0); e = min(e, L.nx - 1 - B.i0);  // the lattice
        if (a > e) { a = 1; e = 0; }
        lo1[s][r] = (short)a; hi1[s][r] = (short)e;
      }
    }
    __syncthreads();
    // exclusive offsets of the rows, per set: warp 0, eight consecutive rows per lane
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x, r0 = lane * (K7P_ROWS / 32);
      int tot[3] = {0, 0, 0};
      for (int q = 0; q < K7P_ROWS / 32; ++q) {
        const int r = r0 + q;
        for (int s = 0; s < 3; ++s) {
          if (r < rows) { offs[s][r] = tot[s]; tot[s] += max(hi1[s][r] - lo1[s][r] + 1, 0); }
        }
      }
      for (int s = 0; s < 3; ++s) {
        int incl = tot[s];
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        laneTot[s][lane] = incl - tot[s];
        nS[s] = __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32 && nS[0] > 65535) nS[0] = -1;
    for (int r = threadIdx.x; r < rows; r += K7P_NT) {
      const int lane = r / (K7P_ROWS / 32);
      for (int s = 0; s < 3; ++s) {
        const int lo = lo1[s][r], hi = hi1[s][r];
        unsigned w = 0xffu;  // empty: first > last
        if (lo <= hi) w = (unsigned)lo | ((unsigned)hi << 8) | ((unsigned)((offs[s][r] + laneTot[s][lane]) & 0xffff) << 16);
        rowsOut[((size_t)blockIdx.x * rowsCapG + r) * 3 + s] = w;
      }
    }
  }
  if (threadIdx.x == 0) {
    if (fits && nS[0] < 0) fits = false;
    V7Hdr h;
    h.b = b; h.cls = cls; h.nShell = 0; h.shellOff = 0;  // (k7_finish)
    h.flags = V7_F_LOCAL | (fits ? V7_F_FITS : 0) | (B.edge ? V7_F_EDGE : 0);
    h.j0 = B.j0; h.k0 = B.k0; h.by = B.by; h.bz = B.bz; h.i0 = B.i0; h.bx = B.bx;
    h.nA = max(nS[0], 0); h.nB = nS[1]; h.nC = nS[2]; h.pad0 = 0; h.pad1 = 0;
    h.cx = cx; h.cy = cy; h.cz = cz; h.pad2 = 0.0;
    hdr[blockIdx.x] = h;
  }
}

__global__ void k7_finish(InstanceSoA in, int nInst, const int* __restrict__ order, V7Hdr* __restrict__ hdr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nInst) return;
  const int b = order[t];
  hdr[t].nShell = in.nShell[b];
  hdr[t].shellOff = in.shellOff[b];
  if (in.onLocal[b] == 0) hdr[t].flags &= ~V7_F_LOCAL;
}

__global__ void k7_local_flags(int nInst, const V7Hdr* __restrict__ hdr, int* __restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nInst) return;
  flag[t] = (t < nInst && (hdr[t].flags & V7_F_LOCAL)) ? 1 : 0;  // (one entry past the end for the scan's total)
}
__global__ void k7_local_compact(int nInst, const V7Hdr* __restrict__ hdr, const int* __restrict__ flag, const int* __restrict__ off,
                                 V7Hdr* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nInst || !flag[t]) return;
  V7Hdr h = hdr[t];
  h.pad0 = t;
  out[off[t]] = h;
}
// force / torque partials of every instance <- 0 (the staged stress kernel then only visits the instances with local points)
__global__ void k7_zero_ft(InstanceSoA in, int nInst, double* __restrict__ Tcorr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nInst) return;
  for (int q = 0; q < 3; ++q) { in.F[q][t] = 0.0; in.T[q][t] = 0.0; if (Tcorr) Tcorr[3 * (size_t)t + q] = 0.0; }
}

// row tables of a CTA from its packed record (sets first..last); rowG as build_region7 leaves it.  No barrier inside.
__device__ __forceinline__ void load_region7(const LatticeDev& L, const Box7& B, const unsigned* __restrict__ packed, int first, int last, int rowsCap,
                                             int* rowBase, int* rowLo, int* rowHi, int* rowG) {
  const int r = threadIdx.x;
  if (r < B.by * B.bz) {
    for (int s = first; s <= last; ++s) {
      const unsigned w = __ldg(packed + 3 * r + s);
      const int lo = B.i0 + (int)(w & 0xffu), hi = B.i0 + (int)((w >> 8) & 0xffu);
      rowLo[s * rowsCap + r] = lo;
      rowHi[s * rowsCap + r] = hi;
      rowBase[s * rowsCap + r] = (int)(w >> 16) - lo;
    }
    const int bk = r / B.by, bj = r - bk * B.by;
    rowG[r] = L.nx * ((B.j0 + bj) + L.ny * (B.k0 + bk));
  }
}

// ------------------------------------------------------------------------------------------------
// K2/K3 v7: direct forcing, one CTA per instance
// ------------------------------------------------------------------------------------------------
struct Task7Ctx {
  Geo7 G; Box7 B; Tab7 tab;
  const int *baseA, *loA, *hiA, *baseB, *loB, *hiB, *baseC, *loC, *hiC;
  const double *pA, *pB, *pC;
  double cx, cy, cz, rdT;
  double rI, rS, rO;
  V3 vp, om;
  int kind, nT, nk;
};

// The shell cells of the instance, four lanes per cell (one per corner column of its polar bin): each lane evaluates the
// three gather stencils of its column and weighs the two column terms with the cell's inverse distances to the inner and
// surface points (particleShape.C:113-131); the four lanes are summed by two shuffles and lane 0 adds F_i to the field
// (particleIBM.C:230-263).  A d = 8 sphere has 584 distinct columns for 4 x 155 corner references: sharing columns between
// cells would save 6 % of the stencils and cost a 32 KB table, a compaction pass and a barrier.
template <bool EDGE, int NT>
__device__ __forceinline__ int cells7(const Task7Ctx& X, const int* __restrict__ shellCell, const int* __restrict__ shellJK, int nShell,
                                      size_t C, double* __restrict__ F) {
  int guard = 0;
  const int nTask = 4 * nShell;
  const int nxy = X.G.nx * X.G.ny;
  const float rnxy = 1.0f / (float)nxy, rnx = 1.0f / (float)X.G.nx;
  for (int base = 0; base < nTask; base += NT) {
    const int task = base + threadIdx.x;
    const bool valid = task < nTask;
    double fx = 0.0, fy = 0.0, fz = 0.0, wsum = 0.0, tOx = 0.0, tOy = 0.0, tOz = 0.0, tSx = 0.0, tSy = 0.0, tSz = 0.0;
    int lat = 0, tinyI = 0, tinyS = 0;
    if (valid) {
      const int cell = task >> 2, q = task & 3;
      lat = shellCell[cell];
      const int pk = shellJK[cell];
      const int j = pk & 0xffff, k = pk >> 16;
      const bool deg = X.nk != 1 && k == X.nk - 1;  // App. B #3: bins with k == nk-1 reference point 0 twelve times
      int jq = (q & 1) ? (j + 1 == X.nT ? 0 : j + 1) : j;
      int kq = (q & 2) ? (X.nk != 1 ? k + 1 : k) : k;
      if (deg) { jq = 0; kq = 0; }
      // (a degenerate bin uses point 0 = the inner point of (0,0) for all three roles, particleShape.C:148)
      const V3 xI = tab7_point(X.kind, X.tab, X.rI, X.cx, X.cy, X.cz, jq, kq);
      const V3 xS = deg ? xI : tab7_point(X.kind, X.tab, X.rS, X.cx, X.cy, X.cz, jq, kq);
      const V3 xO = deg ? xI : tab7_point(X.kind, X.tab, X.rO, X.cx, X.cy, X.cz, jq, kq);
      double g[3];
      St7 st;
      if (make_st7<EDGE>(X.G, X.B, X.baseC, X.loC, X.hiC, xI, st, guard)) { gat7<3>(st, X.pC, g); tOx = g[0]; tOy = g[1]; tOz = g[2]; }
      if (make_st7<EDGE>(X.G, X.B, X.baseA, X.loA, X.hiA, xO, st, guard)) { gat7<3>(st, X.pA, g); tOx += g[0]; tOy += g[1]; tOz += g[2]; }
      {
        const double rx = xO.x - X.cx, ry = xO.y - X.cy, rz = xO.z - X.cz;
        tOx = fma(X.vp.x + (X.om.y * rz - X.om.z * ry), X.rdT, tOx);
        tOy = fma(X.vp.y + (X.om.z * rx - X.om.x * rz), X.rdT, tOy);
        tOz = fma(X.vp.z + (X.om.x * ry - X.om.y * rx), X.rdT, tOz);
      }
      if (make_st7<EDGE>(X.G, X.B, X.baseB, X.loB, X.hiB, xS, st, guard)) { gat7<3>(st, X.pB, g); tSx = g[0]; tSy = g[1]; tSz = g[2]; }
      {
        const double rx = xS.x - X.cx, ry = xS.y - X.cy, rz = xS.z - X.cz;
        tSx = fma(X.vp.x + (X.om.y * rz - X.om.z * ry), X.rdT, tSx);
        tSy = fma(X.vp.y + (X.om.z * rx - X.om.x * rz), X.rdT, tSy);
        tSz = fma(X.vp.z + (X.om.x * ry - X.om.y * rx), X.rdT, tSz);
      }
      // lattice index -> (I,J,K) by reciprocal multiplication and one correction (the float quotient is within one of the
      // truth: relative error 2^-23 on a quotient below 2^21) instead of two integer divisions
      int K = (int)((float)lat * rnxy), rem = lat - K * nxy;
      if (rem < 0) { --K; rem += nxy; } else if (rem >= nxy) { ++K; rem -= nxy; }
      int J = (int)((float)rem * rnx), I = rem - J * X.G.nx;
      if (I < 0) { --J; I += X.G.nx; } else if (I >= X.G.nx) { ++J; I -= X.G.nx; }
      const double Cx = X.G.ox + (I + 0.5) * X.G.dx, Cy = X.G.oy + (J + 0.5) * X.G.dy, Cz = X.G.oz + (K + X.G.kz0 + 0.5) * X.G.dz;
      const double dI2 = (Cx - xI.x) * (Cx - xI.x) + (Cy - xI.y) * (Cy - xI.y) + (Cz - xI.z) * (Cz - xI.z);
      const double dS2 = (Cx - xS.x) * (Cx - xS.x) + (Cy - xS.y) * (Cy - xS.y) + (Cz - xS.z) * (Cz - xS.z);
      tinyI = dI2 < IBM_SMALL * IBM_SMALL;
      tinyS = dS2 < IBM_SMALL * IBM_SMALL;
      const double wI = rsqrt7(fmax(dI2, 1e-300)), wS = rsqrt7(fmax(dS2, 1e-300));
      wsum = wI + wS;
      fx = fma(wI, tOx, wS * tSx); fy = fma(wI, tOy, wS * tSy); fz = fma(wI, tOz, wS * tSz);
    }
    // the four corners of a cell sit in four adjacent lanes (NT and the warp size are multiples of 4)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned quad = 0xfu << (lane & ~3u);
    const unsigned anyI = __ballot_sync(0xffffffffu, tinyI) & quad, anyS = __ballot_sync(0xffffffffu, tinyS) & quad;
    if (anyI | anyS) {
      // one-hot rule of particleShape.C:121-127: the first q (0..7) within SMALL wins; q < 4 are the inner points
      const int first = anyI ? (__ffs(anyI) - 1) - (int)(lane & ~3u) : 4 + (__ffs(anyS) - 1) - (int)(lane & ~3u);
      const int q = lane & 3;
      const double wI = (first == q) ? 1.0 : 0.0, wS = (first == 4 + q) ? 1.0 : 0.0;
      wsum = wI + wS;
      fx = fma(wI, tOx, wS * tSx); fy = fma(wI, tOy, wS * tSy); fz = fma(wI, tOz, wS * tSz);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, o);
      fy += __shfl_xor_sync(0xffffffffu, fy, o);
      fz += __shfl_xor_sync(0xffffffffu, fz, o);
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (valid && (lane & 3u) == 0) {
      const double rW = rcp7(wsum);
      atomicAdd(&F[lat], fx * rW);
      atomicAdd(&F[C + lat], fy * rW);
      atomicAdd(&F[2 * C + lat], fz * rW);
    }
  }
  return guard;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_forcing_v7(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const V7Hdr* __restrict__ hdrs,
             const unsigned* __restrict__ rowsPacked, int rowsCapG, int identity,
             const int* __restrict__ shellCell, const int* __restrict__ shellJK, const double* __restrict__ U,
             const double* __restrict__ Uold, const double* __restrict__ S, double dT, double* __restrict__ F, V7Caps cap,
             int* __restrict__ slowList, int* __restrict__ slowCount, int* __restrict__ guardFlag) {
  extern __shared__ double smem7[];
  if ((int)blockIdx.x >= nInst) return;
  const V7Hdr h = hdrs[blockIdx.x];
  const int b = h.b;
  const int nShell = h.nShell;
  if (!(h.flags & V7_F_LOCAL) || nShell <= 0) return;
  const V7Layout lay = v7_layout(cap);
  char* base = (char*)smem7;
  double* planes = (double*)(base + lay.planes);
  double* tabD = (double*)(base + lay.tab);
  int* rowBase = (int*)(base + lay.rows);
  int* rowLo = rowBase + 3 * cap.rowsCap;
  int* rowHi = rowLo + 3 * cap.rowsCap;
  int* rowG = rowHi + 3 * cap.rowsCap;
  unsigned short* elemRow = (unsigned short*)(base + lay.elemRow);

  Box7 B;
  B.j0 = h.j0; B.k0 = h.k0; B.by = h.by; B.bz = h.bz; B.i0 = h.i0; B.bx = h.bx; B.edge = (h.flags & V7_F_EDGE) ? 1 : 0;
  const int nA = h.nA, nB = h.nB, nC = h.nC;
  if (!(h.flags & V7_F_FITS) || B.by * B.bz > cap.rowsCap || B.by * B.bz > NT || 3 * (nA + nB + nC) > cap.planeCap || nA > cap.elemCap) {
    if (threadIdx.x == 0) slowList[atomicAdd(slowCount, 1)] = b;
    return;
  }
  // the lines of the bounding rows on their way into L2 while the tables are unpacked (the staging loads then wait for an L2
  // hit instead of a DRAM access)
  if (cap.prefetch) {
    const size_t C = (size_t)L.nCells;
    const int rows = B.by * B.bz;
    for (int r = threadIdx.x; r < 2 * rows; r += NT) {  // two 128-byte lines cover a row of up to 16 cells
      const int row = r >> 1, bk = row / B.by, bj = row - bk * B.by;
      const size_t g = (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk)) + (size_t)B.i0 + ((r & 1) ? (size_t)(B.bx - 1) : 0);
#pragma unroll
      for (int q = 0; q < 3; ++q) { prefetch_l2(S + q * C + g); prefetch_l2(U + q * C + g); prefetch_l2(Uold + q * C + g); }
    }
  }
  const ShapeClassDev sc = classes[h.cls];
  const double cx = h.cx, cy = h.cy, cz = h.cz;
  const int nT = sc.nTheta, nk = sc.nk;
  const double R = sc.d / 2.0;
  const double rL[3] = {xadd(R, -1 * sc.delta), xadd(R, 0.0), xadd(R, sc.delta)};  // R + delta[i]  (sphere.C:167,173)
  const V3 vp{in.v[0][b], in.v[1][b], in.v[2][b]}, omv{in.om[0][b], in.om[1][b], in.om[2][b]};
  const int off = h.shellOff;
  const Tab7 tab = tab7_load(sc, tabD, cap);
  // (compacted launch on several ranks: the record carries its schedule position; on one GPU it is blockIdx, known at once)
  const int pos = identity ? (int)blockIdx.x : h.pad0;
  load_region7(L, B, rowsPacked + (size_t)pos * rowsCapG * 3, 0, 2, cap.rowsCap, rowBase, rowLo, rowHi, rowG);
  // elemRow: row of every element of set 0 (the staging loop runs over the elements of the widest set)
  if ((int)threadIdx.x < B.by * B.bz) {
    const int r = threadIdx.x;
    const int lo = rowLo[r], hi = rowHi[r], o = rowBase[r] + lo;
    for (int x = lo; x <= hi; ++x) elemRow[o + x - lo] = (unsigned short)r;
  }
  __syncthreads();

  // ---- stage: every needed cell value read once, combined on the way (records of three components) ----
  const size_t C = (size_t)L.nCells;
  const double rdT = 1.0 / dT;
  double* pA = planes;
  double* pB = planes + 3 * nA;
  double* pC = pB + 3 * nB;
  const int* loB = rowLo + cap.rowsCap; const int* hiB = rowHi + cap.rowsCap; const int* baseB = rowBase + cap.rowsCap;
  const int* loC = rowLo + 2 * cap.rowsCap; const int* hiC = rowHi + 2 * cap.rowsCap; const int* baseC = rowBase + 2 * cap.rowsCap;
#pragma unroll 2
  for (int a = threadIdx.x; a < nA; a += NT) {
    const int row = elemRow[a];
    const int I = a - rowBase[row];
    const size_t g = (size_t)(rowG[row] + I);
    const double s0 = __ldg(S + g), s1 = __ldg(S + C + g), s2 = __ldg(S + 2 * C + g);
    const double u0 = __ldg(U + g), u1 = __ldg(U + C + g), u2 = __ldg(U + 2 * C + g);
    const bool inB = I >= loB[row] && I <= hiB[row];
    double o0 = 0.0, o1 = 0.0, o2 = 0.0;
    if (inB) { o0 = __ldg(Uold + g); o1 = __ldg(Uold + C + g); o2 = __ldg(Uold + 2 * C + g); }
    pA[3 * a] = fma(-rdT, u0, s0); pA[3 * a + 1] = fma(-rdT, u1, s1); pA[3 * a + 2] = fma(-rdT, u2, s2);
    if (inB) {
      const int ib = 3 * (baseB[row] + I);
      pB[ib] = fma(-rdT, o0, s0); pB[ib + 1] = fma(-rdT, o1, s1); pB[ib + 2] = fma(-rdT, o2, s2);
      if (I >= loC[row] && I <= hiC[row]) {
        const int ic = 3 * (baseC[row] + I);
        pC[ic] = -rdT * o0; pC[ic + 1] = -rdT * o1; pC[ic + 2] = -rdT * o2;
      }
    }
  }
  __syncthreads();

  Task7Ctx X;
  X.G = geo7(L); X.B = B; X.tab = tab;
  X.baseA = rowBase; X.loA = rowLo; X.hiA = rowHi; X.baseB = baseB; X.loB = loB; X.hiB = hiB; X.baseC = baseC; X.loC = loC; X.hiC = hiC;
  X.pA = pA; X.pB = pB; X.pC = pC;
  X.cx = cx; X.cy = cy; X.cz = cz; X.rdT = rdT;
  X.rI = rL[0]; X.rS = rL[1]; X.rO = rL[2];
  X.vp = vp; X.om = omv;
  X.kind = sc.kind; X.nT = nT; X.nk = nk;
  const int bad = B.edge ? cells7<true, NT>(X, shellCell + off, shellJK + off, nShell, C, F) : cells7<false, NT>(X, shellCell + off, shellJK + off, nShell, C, F);
  if (bad) atomicExch(guardFlag, 1);
}

// ------------------------------------------------------------------------------------------------
// K4 v7: surface-stress integral (particleIBM.C:541-581), one CTA per instance: p and the six devRhoReff components staged
// on the reach of the surface layer only (the round-1 kernel staged the solid 11^3 box: 3.6x the bytes that are needed)
// ------------------------------------------------------------------------------------------------
struct Str7Ctx {
  Geo7 G; Box7 B; Tab7 tab;
  const int *baseB, *loB, *hiB;
  const double* planes;
  double cx, cy, cz, R, delta, rS, pRefOverRho, rhoRef, sfScale;
  double rot[9];
  int kind, nT, nk, nB;
};

template <bool EDGE, int NT, bool SOA>
__device__ __forceinline__ int stress_points7(const Str7Ctx& X, double* acc) {
  int guard = 0;
  const int nSurf = X.nT * X.nk;
  // the index splits below by reciprocal multiplication (the integer divisions were a tenth of the kernel's instructions):
  // exact for the 16-bit ranges they are used on - (n + 1/2)/d is at least 1/(2d) away from an integer, 6e-8 n/d of rounding is less
  const bool small = nSurf <= 32767;
  const float invNT = 1.0f / (float)X.nT, inv2NT = 0.5f / (float)X.nT;
  // Order of the points: two circles of latitude at a time, alternating between them, so that the 16 lanes of a half-warp
  // hold an 8 x 2 patch of the polar grid.  Neighbouring points are about half a cell apart: such a patch lies in 2-4 cells,
  // lanes in the same cell read the same shared-memory words (one broadcast), and the bank conflicts between the few distinct
  // cells that remain are rare - 1.1 wavefronts per half-warp load against 2.0 for a walk along a meridian (measured: the
  // gather of the first version was 761 M wavefronts for 333 M loads and the kernel was bound by them).
  // (an odd last circle goes alone at the end: no empty slots - 1250 points of a d = 8 sphere are five rounds of 256 threads,
  // 1300 slots would be six)
  const int nPaired = 2 * X.nT * (X.nk >> 1);
  for (int s = threadIdx.x; s < nSurf; s += NT) {
    int i, k;
    if (s < nPaired) {
      const int kp = small ? (int)(((float)s + 0.5f) * inv2NT) : s / (2 * X.nT);
      const int rem = s - kp * 2 * X.nT;
      i = rem >> 1; k = 2 * kp + (rem & 1);
    } else {
      i = s - nPaired; k = X.nk - 1;
    }
    const int i2 = i + X.nT * k;
    const V3 pt = tab7_point(X.kind, X.tab, X.rS, X.cx, X.cy, X.cz, i, k);
    double g[7] = {0, 0, 0, 0, 0, 0, 0};
    double pp = 0.0;
    St7 st;
    if (make_st7<EDGE>(X.G, X.B, X.baseB, X.loB, X.hiB, pt, st, guard, true)) {
      if (SOA) gat7_planes<7>(st, X.planes, X.nB, g); else gat7<7>(st, X.planes, g);
      pp = (g[0] - 2.0 * st.c[0] * X.pRefOverRho) * X.rhoRef;
#pragma unroll
      for (int q = 1; q < 7; ++q) g[q] *= X.rhoRef;
    }
    const double* tau = g + 1;
    // App. B #5: lever arm as coded, mesh_[index2(i,k)] read through index(i,j,k)
    int lay2, j2, k2;
    if (small) {
      const int q3 = (int)(((unsigned)i2 * 43691u) >> 17);  // i2 / 3 for i2 < 2^16
      lay2 = i2 - 3 * q3;
      k2 = (int)(((float)q3 + 0.5f) * invNT);
      j2 = q3 - k2 * X.nT;
    } else {
      lay2 = i2 % 3; j2 = (i2 / 3) % X.nT; k2 = i2 / (3 * X.nT);
    }
    const double r2 = xadd(X.R, lay2 == 0 ? -1 * X.delta : (lay2 == 1 ? 0.0 : X.delta));
    const V3 m2 = tab7_point(X.kind, X.tab, r2, X.cx, X.cy, X.cz, j2, k2);
    const double Rx = m2.x - X.cx, Ry = m2.y - X.cy, Rz = m2.z - X.cz;
    const double Rcx = pt.x - X.cx, Rcy = pt.y - X.cy, Rcz = pt.z - X.cz;
    // Sf_[index2(i,k)]   (sphere.C:134-157, cylinder.C:116-122)
    double sx, sy, sz;
    if (X.kind == 0) {
      const double f = X.sfScale * X.tab.ck[k];
      sx = -(X.tab.cosT[i] * X.tab.sinP[k]) * f; sy = -(X.tab.sinT[i] * X.tab.sinP[k]) * f; sz = -X.tab.cosP[k] * f;
    } else {
      sx = -X.tab.cosT[i] * X.sfScale; sy = -X.tab.sinT[i] * X.sfScale; sz = 0.0;
    }
    const double* rot = X.rot;
    const double Sx = rot[0] * sx + rot[1] * sy + rot[2] * sz, Sy = rot[3] * sx + rot[4] * sy + rot[5] * sz,
                 Sz = rot[6] * sx + rot[7] * sy + rot[8] * sz;
    const double rmag = rsqrt7(Sx * Sx + Sy * Sy + Sz * Sz);
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
  return guard;
}

template <int NT, int MINB, bool SOA>
__global__ void __launch_bounds__(NT, MINB)
k_stress_v7(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, const V7Hdr* __restrict__ hdrs,
            const unsigned* __restrict__ rowsPacked, int rowsCapG, int identity,
            const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef, double* __restrict__ Tcorr,
            V7Caps cap, int* __restrict__ slowList, int* __restrict__ slowCount, int* __restrict__ guardFlag,
            double* __restrict__ Fclear, const int* __restrict__ shellCell) {
  extern __shared__ double smem7[];
  __shared__ double red[NT / 32][9];
  if ((int)blockIdx.x >= nInst) return;
  const V7Hdr h = hdrs[blockIdx.x];
  const int b = h.b;
  const bool on = (h.flags & V7_F_LOCAL) != 0;
  if (Fclear) {
    // solve(): the consumer of the forcing field is done with it, and the field has to be zero again at the cells the last
    // forcing() wrote before the shell lists are rebuilt.  Every instance re-zeroes its own shell cells here - stores that
    // nobody waits for - instead of a separate kernel over all of them beside this one (0.87 ms of contended side-stream work).
    const size_t C = (size_t)L.nCells;
    for (int q = threadIdx.x; q < h.nShell; q += NT) {
      const int lat = __ldg(shellCell + h.shellOff + q);
      Fclear[lat] = 0.0; Fclear[C + lat] = 0.0; Fclear[2 * C + lat] = 0.0;
    }
  }
  const V7Layout lay = v7_layout(cap);
  char* base = (char*)smem7;
  double* planes = (double*)(base + lay.planes);
  double* tabD = (double*)(base + lay.tab);
  int* rowBase = (int*)(base + lay.rows);
  int* rowLo = rowBase + 3 * cap.rowsCap;
  int* rowHi = rowLo + 3 * cap.rowsCap;
  int* rowG = rowHi + 3 * cap.rowsCap;
  unsigned short* elemRow = (unsigned short*)(base + lay.elemRow);
  // the box of the OUTER reach (the launch record is shared with the forcing kernel): the rows beyond the surface reach are
  // empty in set 1
  Box7 B;
  B.j0 = h.j0; B.k0 = h.k0; B.by = h.by; B.bz = h.bz; B.i0 = h.i0; B.bx = h.bx; B.edge = (h.flags & V7_F_EDGE) ? 1 : 0;
  const int nB = h.nB;
  // Instances without local points (other slabs' instances on a multi-GPU run) do no work but MUST still write zero
  // partials below: their F/T slots hold last step's combined totals, and the cross-rank sum adds whatever is in them.
  if (on && (!(h.flags & V7_F_FITS) || B.by * B.bz > cap.rowsCap || B.by * B.bz > NT || 7 * nB > cap.planeCap || nB > cap.elemCap)) {
    if (threadIdx.x == 0) slowList[atomicAdd(slowCount, 1)] = b;
    return;  // the list kernel writes this instance's F/T
  }
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (on) {
    if (cap.prefetch) {
      const size_t C = (size_t)L.nCells;
      const int rows = B.by * B.bz;
      for (int r = threadIdx.x; r < 2 * rows; r += NT) {
        const int row = r >> 1, bk = row / B.by, bj = row - bk * B.by;
        const size_t g = (size_t)L.nx * ((size_t)(B.j0 + bj) + (size_t)L.ny * (size_t)(B.k0 + bk)) + (size_t)B.i0 + ((r & 1) ? (size_t)(B.bx - 1) : 0);
        prefetch_l2(P + g);
#pragma unroll
        for (int q = 0; q < 6; ++q) prefetch_l2(Tau + q * C + g);
      }
    }
    const ShapeClassDev sc = classes[h.cls];
    const double cx = h.cx, cy = h.cy, cz = h.cz;
    const int nT = sc.nTheta, nk = sc.nk;
    const double R = sc.d / 2.0;
    const double rhoS = xadd(R, 0.0);
    double rot9[9];  // (asked for early: the answer is back long before the points need it)
#pragma unroll
    for (int q = 0; q < 9; ++q) rot9[q] = __ldg(in.rot + 9 * (size_t)b + q);
    const Tab7 tab = tab7_load(sc, tabD, cap);
    const int pos = identity ? (int)blockIdx.x : h.pad0;
    load_region7(L, B, rowsPacked + (size_t)pos * rowsCapG * 3, 1, 1, cap.rowsCap, rowBase, rowLo, rowHi, rowG);
    const int* loB = rowLo + cap.rowsCap; const int* hiB = rowHi + cap.rowsCap; const int* baseB = rowBase + cap.rowsCap;
    if ((int)threadIdx.x < B.by * B.bz) {
      const int r = threadIdx.x;
      const int lo = loB[r], hi = hiB[r], o = baseB[r] + lo;
      for (int x = lo; x <= hi; ++x) elemRow[o + x - lo] = (unsigned short)r;
    }
    __syncthreads();
    const size_t C = (size_t)L.nCells;
#pragma unroll 2
    for (int a = threadIdx.x; a < nB; a += NT) {
      const int row = elemRow[a];
      const int I = a - baseB[row];
      const size_t g = (size_t)(rowG[row] + I);
      double v[7];
      v[0] = __ldg(P + g);
#pragma unroll
      for (int q = 0; q < 6; ++q) v[1 + q] = __ldg(Tau + q * C + g);
#pragma unroll
      for (int q = 0; q < 7; ++q) planes[SOA ? q * nB + a : 7 * a + q] = v[q];
    }
    __syncthreads();
    Str7Ctx X;
    X.G = geo7(L); X.B = B; X.tab = tab;
    X.baseB = baseB; X.loB = loB; X.hiB = hiB; X.planes = planes;
    X.cx = cx; X.cy = cy; X.cz = cz; X.R = R; X.delta = sc.delta; X.rS = rhoS; X.pRefOverRho = pRefOverRho; X.rhoRef = rhoRef;
    X.sfScale = (sc.kind == 0) ? (R * R) * sc.dTheta : R * sc.dTheta * sc.l;
#pragma unroll
    for (int q = 0; q < 9; ++q) X.rot[q] = rot9[q];
    X.kind = sc.kind; X.nT = nT; X.nk = nk; X.nB = nB;
    const int bad = B.edge ? stress_points7<true, NT, SOA>(X, acc) : stress_points7<false, NT, SOA>(X, acc);
    if (bad) atomicExch(guardFlag, 1);
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
    for (int w = 0; w < NT / 32; ++w) v += red[w][threadIdx.x];
    const int q = threadIdx.x;
    if (q < 3) in.F[q][b] = v;
    else if (q < 6) in.T[q - 3][b] = v;
    else if (Tcorr) Tcorr[3 * b + (q - 6)] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Upper estimate of the staged cells of a ball of radius rho (lattice units per axis given by d[]) over centre offsets
// within a cell: the same row construction as build_region7, evaluated on the host for a lattice of sub-cell offsets.
static long long region_cells_host(const double* d, double rho, int nGeoD) {
  long long worst = 0;
  const int NS = 5;
  for (int sz = 0; sz < (nGeoD == 3 ? NS : 1); ++sz)
    for (int sy = 0; sy < NS; ++sy)
      for (int sx = 0; sx < NS; ++sx) {
        const double c[3] = {(sx + 0.013) / NS * d[0], (sy + 0.037) / NS * d[1], nGeoD == 3 ? (sz + 0.021) / NS * d[2] : 0.5 * d[2]};
        const int j0 = (int)std::floor((c[1] - rho) / d[1]) - 2, j1 = (int)std::floor((c[1] + rho) / d[1]) + 2;
        const int k0 = nGeoD == 3 ? (int)std::floor((c[2] - rho) / d[2]) - 2 : 0, k1 = nGeoD == 3 ? (int)std::floor((c[2] + rho) / d[2]) + 2 : 0;
        auto own = [&](int J, int K, int& lo, int& hi) {
          const double ey = std::max(std::max(J * d[1] - c[1], c[1] - (J + 1) * d[1]), 0.0);
          const double ez = nGeoD == 3 ? std::max(std::max(K * d[2] - c[2], c[2] - (K + 1) * d[2]), 0.0) : 0.0;
          const double h2 = rho * rho * (1.0 + 1e-8) - ey * ey - ez * ez;
          if (h2 < 0) { lo = 1; hi = 0; return; }
          const double h = std::sqrt(h2);
          lo = (int)std::floor((c[0] - h) / d[0]); hi = (int)std::floor((c[0] + h) / d[0]);
        };
        long long tot = 0;
        for (int K = k0; K <= k1; ++K)
          for (int J = j0; J <= j1; ++J) {
            int a = 0x7fffffff, b = -0x7fffffff, lo, hi;
            own(J, K, lo, hi); if (lo <= hi) { a = lo - 1; b = hi + 1; }
            own(J + 1, K, lo, hi); if (lo <= hi) { a = std::min(a, lo); b = std::max(b, hi); }
            own(J - 1, K, lo, hi); if (lo <= hi) { a = std::min(a, lo); b = std::max(b, hi); }
            if (nGeoD == 3) {
              own(J, K + 1, lo, hi); if (lo <= hi) { a = std::min(a, lo); b = std::max(b, hi); }
              own(J, K - 1, lo, hi); if (lo <= hi) { a = std::min(a, lo); b = std::max(b, hi); }
            }
            if (a <= b) tot += b - a + 1;
          }
        worst = std::max(worst, tot);
      }
  return worst;
}

// capacities of a launch from the shape classes (host copies); returns false when no class can take the fast path
bool v7_plan(const LatticeDev& L, const ShapeClassDev* hostClasses, int nClasses, int nGeoD, V7Caps& forcing, V7Caps& stress) {
  const int SMEM_BUDGET_F = 113 * 1024, SMEM_BUDGET_S = 113 * 1024;  // at least two resident CTAs per SM (228 KB - 1 KB per CTA)
  V7Caps f{}, s{};
  bool any = false;
  const double d[3] = {L.dx, L.dy, L.dz};
  for (int q = 0; q < nClasses; ++q) {
    const ShapeClassDev& sc = hostClasses[q];
    if (sc.kind == 2) return false;  // mixed populations with ellipses take the second-generation path as a whole
    const double R = sc.d / 2.0;
    const long long nA = region_cells_host(d, R + sc.delta, nGeoD), nB = region_cells_host(d, R, nGeoD),
                    nC = region_cells_host(d, std::max(R - sc.delta, 0.0), nGeoD);
    auto rowsOf = [&](double reach) {
      long long by = (long long)std::ceil(2.0 * reach / d[1]) + 4, bz = nGeoD == 3 ? (long long)std::ceil(2.0 * reach / d[2]) + 4 : 1;
      return std::min<long long>(by, L.ny) * std::min<long long>(bz, L.nz);
    };
    const double cellV = d[0] * d[1] * (nGeoD == 2 ? 1.0 : d[2]);
    const double r0 = std::max(0.0, R - sc.delta);
    const double vol = nGeoD == 2 ? IBM_PI * (R * R - r0 * r0) : 4.0 / 3.0 * IBM_PI * (R * R * R - r0 * r0 * r0);
    // distinct columns of a d = 8 cells sphere: 584 on average, 631 at most, of 4 x 155 corner references (measured over
    // 3000 random centres); beyond the capacity an instance takes the list path, so this is a performance bound only
    const long long colEst = std::min<long long>((long long)sc.nTheta * sc.nk + 1, (long long)(4.0 * vol / cellV) + 56);
    auto up = [](long long v) { return (int)std::min<long long>((v + v / 16 + 16 + 1) & ~1LL, 0x3fffffff); };
    f.planeCap = std::max(f.planeCap, up(3 * (nA + nB + nC)));
    (void)colEst;
    f.rowsCap = std::max(f.rowsCap, (int)std::min<long long>(rowsOf(R + sc.delta), 192));

    f.nTcap = std::max(f.nTcap, sc.nTheta);
    f.nkCap = std::max(f.nkCap, sc.nk);
    f.elemCap = std::max(f.elemCap, up(nA));
    s.planeCap = std::max(s.planeCap, up(7 * nB));
    s.rowsCap = std::max(s.rowsCap, (int)std::min<long long>(rowsOf(R), 192));
    s.nTcap = f.nTcap; s.nkCap = f.nkCap;
    s.elemCap = std::max(s.elemCap, up(nB));
    any = true;
  }
  s.colCap = 0; s.mapCap = 0;
  s.rowsCap = std::max(s.rowsCap, f.rowsCap);  // the launch records carry the box of the OUTER reach for both kernels
  if (!any) return false;
  if (v7_layout(f).total > SMEM_BUDGET_F || v7_layout(s).total > SMEM_BUDGET_S) return false;
  if (f.elemCap > 65535 || s.elemCap > 65535) return false;
  {
    const char* e = getenv("IBM_V7_PREFETCH");  // developer switch: bit 0 forcing, bit 1 stress
    const int pf = e ? atoi(e) : 1;  // measured (gpurun_out/r2y): forcing 5.66 ms with, 5.81 without; stress 5.74 with, 5.34 without
    f.prefetch = pf & 1; s.prefetch = (pf >> 1) & 1;
  }
  forcing = f; stress = s;
  return true;
}

static int v7_threads(bool stress) {  // developer switches: IBM_V7_THREADS (both kernels), IBM_V7_THREADS_F / _S (one of them)
  static int nt[2] = {-1, -1};
  if (nt[0] < 0) {
    const char* e = getenv("IBM_V7_THREADS");
    const char* ef = getenv("IBM_V7_THREADS_F");
    const char* es = getenv("IBM_V7_THREADS_S");
    nt[0] = ef ? atoi(ef) : (e ? atoi(e) : 256);
    nt[1] = es ? atoi(es) : (e ? atoi(e) : 256);
    for (int q = 0; q < 2; ++q)
      if (nt[q] != 256 && nt[q] != 320 && nt[q] != 192 && nt[q] != 224) nt[q] = 256;
  }
  return nt[stress ? 1 : 0];
}

int v7_prep_alloc(V7Prep& p, int cap, int rowsCap) {
  v7_prep_free(p);
  if (cap <= 0 || rowsCap <= 0) return 0;
  if (cudaMalloc(&p.hdr, (size_t)cap * sizeof(V7Hdr)) != cudaSuccess) return 1;
  if (cudaMalloc(&p.rows, (size_t)cap * rowsCap * 3 * sizeof(unsigned)) != cudaSuccess) { cudaFree(p.hdr); p.hdr = nullptr; return 1; }
  if (cudaMalloc(&p.hdrLocal, (size_t)cap * sizeof(V7Hdr)) != cudaSuccess || cudaMalloc(&p.flag, ((size_t)cap + 1) * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&p.flagOff, ((size_t)cap + 1) * sizeof(int)) != cudaSuccess) { v7_prep_free(p); return 1; }
  p.cap = cap; p.rowsCap = rowsCap; p.nLocal = -1;
  return 0;
}
void v7_prep_free(V7Prep& p) {
  if (p.hdr) cudaFree(p.hdr);
  if (p.rows) cudaFree(p.rows);
  if (p.hdrLocal) cudaFree(p.hdrLocal);
  if (p.flag) cudaFree(p.flag);
  if (p.flagOff) cudaFree(p.flagOff);
  p = V7Prep{};
}
void launch_v7_prepare(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, const V7Caps& capF, const V7Prep& prep) {
  if (nInst <= 0) return;
  k7_prepare<<<nInst, K7P_NT, 0, st>>>(L, classes, in, nInst, s.order, capF, prep.hdr, prep.rows, prep.rowsCap);
}

void launch_v7_finish(cudaStream_t st, const InstanceSoA& in, int nInst, const Schedule& s, const V7Prep& prep) {
  if (nInst <= 0) return;
  k7_finish<<<(nInst + 255) / 256, 256, 0, st>>>(in, nInst, s.order, prep.hdr);
}

void launch_v7_compact(cudaStream_t st, int nInst, Schedule& s, const V7Prep& prep, int* dCount) {
  if (nInst <= 0) { cudaMemsetAsync(dCount, 0, sizeof(int), st); return; }
  k7_local_flags<<<(nInst + 256) / 256, 256, 0, st>>>(nInst, prep.hdr, prep.flag);
  launch_exclusive_scan_total(st, s, prep.flag, prep.flagOff, nInst, dCount);
  k7_local_compact<<<(nInst + 255) / 256, 256, 0, st>>>(nInst, prep.hdr, prep.flag, prep.flagOff, prep.hdrLocal);
}

template <int NT, int MINB>
static void forcing_v7_nt(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                          const V7Prep& prep, const int* shellCell, const int* shellJK, const double* U, const double* Uold, const double* S,
                          double dT, double* F, const V7Caps& cap, int* slowList, int* slowCount, int* guard, int sm) {
  static PerDeviceOnce attr;
  static int attrBytes = 0;
  if (attr.first() || sm > attrBytes) {
    cudaFuncSetAttribute(k_forcing_v7<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    attrBytes = std::max(attrBytes, sm);
  }
  const bool ident = prep.nLocal < 0;
  const int n = ident ? nInst : prep.nLocal;
  if (n > 0)
    k_forcing_v7<NT, MINB><<<n, NT, sm, st>>>(L, classes, in, n, ident ? prep.hdr : prep.hdrLocal, prep.rows, prep.rowsCap, ident ? 1 : 0, shellCell, shellJK,
                                             U, Uold, S, dT, F, cap, slowList, slowCount, guard);
}
template <int NT, int MINB, bool SOA>
static void stress_v7_nt(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst, const V7Prep& prep,
                         const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr, const V7Caps& cap, int* slowList,
                         int* slowCount, int* guard, int sm, double* Fclear, const int* shellCell) {
  static PerDeviceOnce attr;
  static int attrBytes = 0;
  if (attr.first() || sm > attrBytes) {
    cudaFuncSetAttribute(k_stress_v7<NT, MINB, SOA>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    attrBytes = std::max(attrBytes, sm);
  }
  const bool ident = prep.nLocal < 0;
  const int n = ident ? nInst : prep.nLocal;
  if (!ident) k7_zero_ft<<<(nInst + 255) / 256, 256, 0, st>>>(in, nInst, Tcorr);  // the instances without local points contribute zero partials
  if (n > 0)
    k_stress_v7<NT, MINB, SOA><<<n, NT, sm, st>>>(L, classes, in, n, ident ? prep.hdr : prep.hdrLocal, prep.rows, prep.rowsCap, ident ? 1 : 0, P, Tau, pRefOverRho,
                                                 rhoRef, Tcorr, cap, slowList, slowCount, guard, Fclear, shellCell);
}

void launch_forcing_v7(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, const V7Prep& prep, const int* shellCell, const int* shellJK, const double* U, const double* Uold,
                       const double* S, double dT, double* F, const V7Caps& cap, int* slowList, int* slowCount, int* guard) {
  if (nInst <= 0) return;
  (void)s;
  const int sm = v7_layout(cap).total;
  cudaMemsetAsync(slowCount, 0, sizeof(int), st);
  const int nt = v7_threads(false);
  // resident CTAs per SM follow from the staged planes (a d = 8 cells sphere: 71 KB -> three); the register bound goes with it
  const bool three = 3 * (sm + 1024) <= 228 * 1024;
#define IBM_F7(NT_, MB_) forcing_v7_nt<NT_, MB_>(st, L, classes, in, nInst, prep, shellCell, shellJK, U, Uold, S, dT, F, cap, slowList, slowCount, guard, sm)
  if (nt == 320) IBM_F7(320, 2);
  else if (nt == 192) { if (three) IBM_F7(192, 3); else IBM_F7(192, 2); }
  else if (nt == 224) { if (three) IBM_F7(224, 3); else IBM_F7(224, 2); }
  else { if (three) IBM_F7(256, 3); else IBM_F7(256, 2); }
#undef IBM_F7
  launch_forcing_v2_list(st, L, classes, in, slowList, slowCount, shellCell, shellJK, U, Uold, S, dT, F);
}

void launch_stress_v7(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const V7Prep& prep, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr,
                      const V7Caps& cap, int* slowList, int* slowCount, int* guard, double* Fclear, const int* shellCell) {
  if (nInst <= 0) return;
  (void)s;
  const int sm = v7_layout(cap).total;
  cudaMemsetAsync(slowCount, 0, sizeof(int), st);
  const int nt = v7_threads(true);
  const bool three = 3 * (sm + 1024) <= 228 * 1024;
  static int soa = -1;
  if (soa < 0) { const char* e = getenv("IBM_V7_STRESS_SOA"); soa = e ? atoi(e) : 0; }
#define IBM_S7(NT_, MB_) do { if (soa) stress_v7_nt<NT_, MB_, true>(st, L, classes, in, nInst, prep, P, Tau, pRefOverRho, rhoRef, Tcorr, cap, slowList, slowCount, guard, sm, Fclear, shellCell); \
                              else stress_v7_nt<NT_, MB_, false>(st, L, classes, in, nInst, prep, P, Tau, pRefOverRho, rhoRef, Tcorr, cap, slowList, slowCount, guard, sm, Fclear, shellCell); } while (0)
  if (nt == 320) IBM_S7(320, 2);
  else if (nt == 192) { if (three) IBM_S7(192, 3); else IBM_S7(192, 2); }
  else if (nt == 224) { if (three) IBM_S7(224, 3); else IBM_S7(224, 2); }
  else { if (three) IBM_S7(256, 3); else IBM_S7(256, 2); }
#undef IBM_S7
  launch_stress_v2_list(st, L, classes, in, slowList, slowCount, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

}  // namespace ibm
