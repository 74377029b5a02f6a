// Reference-order kernels (translation unit compiled with -fmad=false).
//
//   K1  marking:   k1_count / k1_fill / k1_indicator      ({sphere,cylinder,ellipse}::updateCellLists,
//                  particleShape::setNeighbours / setWeights status, particleCloud::updateI)
//   taps:          k_tap_shell / k_tap_points              (parity inspection of the implicit tables)
//   K2/K3 v0:      k_forcing_ref                           (particleIBM::forcing, op-for-op)
//   K4 v0:         k_stress_ref                            (particleIBM::integrateSurfaceStress, op-for-op,
//                                                           flat i-outer/k-inner accumulation order)
// The v0 kernels follow the reference's evaluation order exactly (every sum left-associated, every
// division kept) and are the on-device yardstick for the tuned kernels in ibm_kernels_fast.cu.
#include "ibm_kernels.h"

namespace ibm {

// ------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------
// particleShape::setRotationMatrix   particleShape.C:186-223
__device__ void set_rotation_matrix(double tx, double ty, double tz, double* R) {
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  auto mul = [&](const double* B) {
    double A[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) A[q] = R[q];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        R[3 * r + c] = xadd(xadd(xmul(A[3 * r], B[c]), xmul(A[3 * r + 1], B[3 + c])), xmul(A[3 * r + 2], B[6 + c]));
  };
  if (fabs(tx) > IBM_SMALL) { double c = cos(tx), s = sin(tx); double B[9] = {1.0, 0.0, 0.0, 0.0, c, -s, 0.0, s, c}; mul(B); }
  if (fabs(ty) > IBM_SMALL) { double c = cos(ty), s = sin(ty); double B[9] = {c, 0.0, s, 0.0, 1.0, 0.0, -s, 0.0, c}; mul(B); }
  if (fabs(tz) > IBM_SMALL) { double c = cos(tz), s = sin(tz); double B[9] = {c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0}; mul(B); }
}

// shell test + polar bin of one cell centre.  sphere.C:210-234, cylinder.C:167-187, ellipse.C:165-194
__device__ __forceinline__ bool mark_cell(const ShapeClassDev& sc, double cx, double cy, double cz, double thz,
                                          double Cx, double Cy, double Cz, int& j, int& k) {
  double dx = xsub(Cx, cx), dy = xsub(Cy, cy), dz = xsub(Cz, cz);
  double r = xmag(dx, dy, dz);
  if (sc.kind == 0) {
    double R = sc.d / 2.0;
    double innerR = xsub(R, sc.delta);
    if (!(r >= innerR && r <= R)) return false;
    double theta = atan2(dy, dx);
    double phi = acos(xdiv(dz, r));
    if (theta < 0) theta = xadd(theta, 2.0 * IBM_PI);
    if (phi < 0) phi = xadd(phi, 2.0 * IBM_PI);
    j = (int)xdiv(xmul(theta, (double)sc.nTheta), 2.0 * IBM_PI);
    k = (int)xdiv(xmul(phi, (double)sc.nk), IBM_PI);
    if (j >= sc.nTheta) j = 0;      // App. B #3 DEFINED clamp
    if (k >= sc.nk) k = sc.nk - 1;
    return true;
  } else if (sc.kind == 1) {
    double R = sc.d / 2.0;
    double innerR = xsub(R, sc.delta);
    if (!(r >= innerR && r <= R)) return false;
    double theta = atan2(dy, dx);
    if (theta < 0) theta = xadd(theta, IBM_TWOPI);
    j = (int)xdiv(xmul(theta, (double)(sc.nTheta - 1)), IBM_TWOPI);
    k = 0;
    return true;
  } else {
    double R = shape_r(sc, cx, cy, thz, Cx, Cy);
    double innerR = xsub(R, sc.delta);
    if (!(r >= innerR && r <= R)) return false;
    double theta = xsub(atan2(dy, dx), thz);
    while (theta < 0) theta = xadd(theta, IBM_TWOPI);
    while (theta > IBM_TWOPI) theta = xsub(theta, IBM_TWOPI);
    j = (int)xdiv(xmul(theta, (double)(sc.nTheta - 1)), IBM_TWOPI);
    k = 0;
    return true;
  }
}

struct Box { int i0, i1, j0, j1, k0, k1; };
// same candidate box as the oracle's fast search (any superset of the shell gives the same list)
__device__ __forceinline__ Box bounding_box(const LatticeDev& L, double cx, double cy, double cz, double R) {
  Box b;
  b.i0 = max(0, (int)floor((cx - R - L.ox) / L.dx) - 1); b.i1 = min(L.nx - 1, (int)floor((cx + R - L.ox) / L.dx) + 1);
  b.j0 = max(0, (int)floor((cy - R - L.oy) / L.dy) - 1); b.j1 = min(L.ny - 1, (int)floor((cy + R - L.oy) / L.dy) + 1);
  b.k0 = max(L.ownK0, (int)floor((cz - R - L.oz) / L.dz) - 1 - L.kz0); b.k1 = min(L.ownK1 - 1, (int)floor((cz + R - L.oz) / L.dz) + 1 - L.kz0);  // owned planes only
  return b;
}

__device__ __forceinline__ int block_sum(int v, int* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  int tot = 0;
  for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += sm[q];
  return tot;
}

// ------------------------------------------------------------------------------------------------
// K1 pass 1: rotation matrix, shell count, points-on-mesh count, status, active flag
// ------------------------------------------------------------------------------------------------
__global__ void k1_count(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst) {
  __shared__ int sm[32];
  __shared__ double srot[9];
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  if (threadIdx.x == 0) {
    // particleShape::moveMesh: rotation matrix from the current theta (particleShape.C:291-295)
    set_rotation_matrix(in.th[0][b], in.th[1][b], in.th[2][b], srot);
    for (int q = 0; q < 9; ++q) in.rot[9 * (size_t)b + q] = srot[q];
    if (sc.kind == 2) {  // ellipse::updateCellLists wraps theta.z first (ellipse.C:155-157)
      double tz = in.th[2][b];
      if (tz < 0) tz = xadd(tz, IBM_TWOPI);
      if (tz > IBM_TWOPI) tz = xsub(tz, IBM_TWOPI);
      in.th[2][b] = tz;
    }
  }
  __syncthreads();
  const double thz = in.th[2][b];
  Box bx = bounding_box(L, cx, cy, cz, sc.Rbound);
  int bi = bx.i1 - bx.i0 + 1, bj = bx.j1 - bx.j0 + 1, bk = bx.k1 - bx.k0 + 1;
  int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  int cnt = 0;
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    int I = bx.i0 + t % bi, J = bx.j0 + (t / bi) % bj, K = bx.k0 + t / (bi * bj);
    int j, k;
    if (mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k)) ++cnt;
  }
  int nShell = block_sum(cnt, sm);
  int on = 0;
  for (int n = threadIdx.x; n < sc.N; n += blockDim.x) {
    V3 p = shape_point(sc, srot, cx, cy, cz, n);
    int I, J, K;
    if (find_cell(L, p.x, p.y, p.z, I, J, K) >= 0 && owned_plane(L, K)) ++on;
  }
  int nOn = block_sum(on, sm);
  if (threadIdx.x == 0) {
    int I, J, K;
    int cOn = find_cell(L, cx, cy, cz, I, J, K) >= 0 && owned_plane(L, K);
    in.nShell[b] = nShell;
    in.nPtsOnMesh[b] = nOn;  // local count; summed over the slabs before k1_status on several GPUs
    in.onLocal[b] = (nOn == 0 && !cOn) ? 0 : (nOn == sc.N ? 2 : 1);
  }
}

// second half of the marking kernels: on-mesh point count, status, per-instance outputs (block-wide; `off` < 0 = shell
// buffer overflow)
__device__ __forceinline__ void mark_finish(const LatticeDev& L, const ShapeClassDev& sc, InstanceSoA& in, int b, double cx, double cy,
                                            double cz, const double* srot, int* sm, int off, int nShell) {
  // on-mesh point count (particleShape.C:41-91): all N when the outer layer is strictly inside the lattice
  double rOut = sc.Rbound + sc.delta;
  const double mx = 1e-9 * (fabs(cx) + fabs(cy) + fabs(cz) + rOut) + 1e-300;
  rOut += mx;
  const bool inside = cx - rOut > L.ox && cx + rOut < L.ox + L.nx * L.dx && cy - rOut > L.oy && cy + rOut < L.oy + L.ny * L.dy &&
                      (L.nz == 1 && sc.kind != 0 ? (cz > own_zlo(L) && cz < own_zhi(L))
                                                 : (cz - rOut > own_zlo(L) && cz + rOut < own_zhi(L)));
  // ... and none when the outer layer cannot reach the owned planes at all (other slabs' instances on a multi-GPU run)
  const bool outside = cx + rOut < L.ox || cx - rOut > L.ox + L.nx * L.dx || cy + rOut < L.oy || cy - rOut > L.oy + L.ny * L.dy ||
                       cz + rOut < own_zlo(L) || cz - rOut > own_zhi(L);
  int nOn;
  if (inside) {
    nOn = sc.N;
  } else if (outside) {
    nOn = 0;
  } else {
    int on = 0;
    for (int n = threadIdx.x; n < sc.N; n += blockDim.x) {
      V3 p = shape_point(sc, srot, cx, cy, cz, n);
      int I, J, K;
      if (find_cell(L, p.x, p.y, p.z, I, J, K) >= 0 && owned_plane(L, K)) ++on;
    }
    nOn = block_sum(on, sm);
  }
  if (threadIdx.x == 0) {
    int I, J, K;
    const int cOn = find_cell(L, cx, cy, cz, I, J, K) >= 0 && owned_plane(L, K);
    in.nShell[b] = off >= 0 ? nShell : 0;
    in.shellOff[b] = off >= 0 ? off : 0;
    in.nPtsOnMesh[b] = nOn;
    in.onLocal[b] = (nOn == 0 && !cOn) ? 0 : (nOn == sc.N ? 2 : 1);
  }
}

// ------------------------------------------------------------------------------------------------
// K1 single pass: rotation matrix, status counts and the shell list in ONE kernel.  Shell cells found in the
// bounding box go to a shared-memory list; one atomicAdd on a global cursor reserves the instance's segment of the
// shell buffer (segments are contiguous per instance; their order in the buffer is irrelevant).  Shells larger
// than the shared list are re-evaluated and written directly.  The on-mesh point count is skipped for instances
// whose outermost point layer lies strictly inside the lattice (then all N points are on the mesh).
// ------------------------------------------------------------------------------------------------
#define K1_LIST 1024
__global__ void k1_mark(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                        int* __restrict__ shellCell, int* __restrict__ shellJK, int shellCap, int* __restrict__ cursor,
                        int* __restrict__ overflow) {
  __shared__ int sm[32];
  __shared__ double srot[9];
  __shared__ int listCell[K1_LIST], listJK[K1_LIST];
  __shared__ int nList, segOff, nDirect;
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  if (threadIdx.x == 0) {
    set_rotation_matrix(in.th[0][b], in.th[1][b], in.th[2][b], srot);  // particleShape::moveMesh (particleShape.C:291-295)
    for (int q = 0; q < 9; ++q) in.rot[9 * (size_t)b + q] = srot[q];
    if (sc.kind == 2) {  // ellipse::updateCellLists wraps theta.z first (ellipse.C:155-157)
      double tz = in.th[2][b];
      if (tz < 0) tz = xadd(tz, IBM_TWOPI);
      if (tz > IBM_TWOPI) tz = xsub(tz, IBM_TWOPI);
      in.th[2][b] = tz;
    }
    nList = 0; nDirect = 0;
  }
  __syncthreads();
  const double thz = in.th[2][b];
  Box bx = bounding_box(L, cx, cy, cz, sc.Rbound);
  const int bi = bx.i1 - bx.i0 + 1, bj = bx.j1 - bx.j0 + 1, bk = bx.k1 - bx.k0 + 1;
  const int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  int cnt = 0;
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    const int I = bx.i0 + t % bi, J = bx.j0 + (t / bi) % bj, K = bx.k0 + t / (bi * bj);
    int j, k;
    if (mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k)) {
      ++cnt;
      const int s = atomicAdd(&nList, 1);
      if (s < K1_LIST) { listCell[s] = lat_index(L, I, J, K); listJK[s] = j | (k << 16); }
    }
  }
  const int nShell = block_sum(cnt, sm);
  if (threadIdx.x == 0) {
    const int o = atomicAdd(cursor, nShell);
    segOff = o;
    if (o + nShell > shellCap) { atomicExch(overflow, 1); segOff = -1; }
  }
  __syncthreads();
  const int off = segOff;
  if (off >= 0) {
    if (nShell <= K1_LIST) {
      for (int s = threadIdx.x; s < nShell; s += blockDim.x) { shellCell[off + s] = listCell[s]; shellJK[off + s] = listJK[s]; }
    } else {
      for (int t = threadIdx.x; t < nb; t += blockDim.x) {
        const int I = bx.i0 + t % bi, J = bx.j0 + (t / bi) % bj, K = bx.k0 + t / (bi * bj);
        int j, k;
        if (mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k)) {
          const int s = atomicAdd(&nDirect, 1);
          shellCell[off + s] = lat_index(L, I, J, K);
          shellJK[off + s] = j | (k << 16);
        }
      }
    }
  }
  mark_finish(L, sc, in, b, cx, cy, cz, srot, sm, off, nShell);
}

// ------------------------------------------------------------------------------------------------
// K1 single pass, second version (option mark_variant = 1; same outputs).  What the line-level profile of k1_mark said
// (profiles/r1_lines_k1_mark_big512_final.txt): the kernel is instruction-bound, and the instructions are the atan2 /
// acos of the polar bin executed by warps in which ~1 lane in 5 holds a shell cell (the shell is 21 % of its bounding
// box), plus the runtime integer divisions that turn the box index into (I, J, K) and a block-wide reduction.  Here
//   phase A  every thread runs only the radial test over its box cells and appends the box index of each shell cell to
//            the shared list (spheres / cylinders; an ellipse's radius needs the angle, so it takes mark_cell at once),
//   phase B  the threads walk the COMPACT list and compute the polar bin with all lanes busy,
// the box index is split with a float reciprocal and one correction step, and the shell count is the list counter.
// The per-cell arithmetic is mark_cell's, unchanged, so the lists are the same sets of (cell, j, k).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_box_index(int t, int bi, int bij, float rbi, float rbij, int& i, int& j, int& k) {
  k = (int)((float)t * rbij);
  int rem = t - k * bij;
  if (rem < 0) { --k; rem += bij; } else if (rem >= bij) { ++k; rem -= bij; }
  j = (int)((float)rem * rbi);
  i = rem - j * bi;
  if (i < 0) { --j; i += bi; } else if (i >= bi) { ++j; i -= bi; }
}

// Several ranks: most instances of the directory are far from a rank's slab.  One THREAD per instance settles those (what
// the marking kernel writes for an instance whose outer layer cannot reach the owned planes: rotation matrix, the ellipse's
// wrapped angle, zero counts) and flags the others; the marking kernel then runs over the compacted list of the flagged
// ones (its grid is still the directory - the list length lives on the device - but a block beyond the list leaves after
// one load instead of a class record, a centre and three sincos).
__global__ void k1_far(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, int* __restrict__ nearFlag) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nInst) return;
  if (b == nInst) { nearFlag[b] = 0; return; }  // (one entry past the end for the scan's total)
  const ShapeClassDev& sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  double rOut = sc.Rbound + sc.delta;
  rOut += 1e-9 * (fabs(cx) + fabs(cy) + fabs(cz) + rOut) + 1e-300;
  const bool outside = cx + rOut < L.ox || cx - rOut > L.ox + L.nx * L.dx || cy + rOut < L.oy || cy - rOut > L.oy + L.ny * L.dy ||
                       cz + rOut < own_zlo(L) || cz - rOut > own_zhi(L);
  nearFlag[b] = outside ? 0 : 1;
  if (!outside) return;
  double r9[9];
  set_rotation_matrix(in.th[0][b], in.th[1][b], in.th[2][b], r9);
  for (int q = 0; q < 9; ++q) in.rot[9 * (size_t)b + q] = r9[q];
  if (sc.kind == 2) {
    double tz = in.th[2][b];
    if (tz < 0) tz = xadd(tz, IBM_TWOPI);
    if (tz > IBM_TWOPI) tz = xsub(tz, IBM_TWOPI);
    in.th[2][b] = tz;
  }
  in.nShell[b] = 0; in.shellOff[b] = 0; in.nPtsOnMesh[b] = 0; in.onLocal[b] = 0;
}
__global__ void k1_near_list(int nInst, const int* __restrict__ nearFlag, const int* __restrict__ nearOff, int* __restrict__ nearList) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nInst && nearFlag[b]) nearList[nearOff[b]] = b;
}

__global__ void k1_mark2(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                         int* __restrict__ shellCell, int* __restrict__ shellJK, int shellCap, int* __restrict__ cursor,
                         int* __restrict__ overflow, const int* __restrict__ nearList, const int* __restrict__ nearCount) {
  __shared__ int sm[32];
  __shared__ double srot[9];
  __shared__ int listCell[K1_LIST], listJK[K1_LIST];
  __shared__ int nList, segOff, nDirect;
  if ((int)blockIdx.x >= nInst) return;
  if (nearList && (int)blockIdx.x >= *nearCount) return;  // (block-uniform)
  const int b = nearList ? nearList[blockIdx.x] : (int)blockIdx.x;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  // An instance whose outer layer cannot reach the owned planes (other slabs' instances on a multi-GPU run: 7 of 8 on eight
  // GPUs) has no shell cell and no point here: thread 0 writes what mark_finish would and the block is gone, without the
  // four barriers below.  Same test as mark_finish.
  double rOutQ = sc.Rbound + sc.delta;
  rOutQ += 1e-9 * (fabs(cx) + fabs(cy) + fabs(cz) + rOutQ) + 1e-300;
  const bool outsideQ = cx + rOutQ < L.ox || cx - rOutQ > L.ox + L.nx * L.dx || cy + rOutQ < L.oy || cy - rOutQ > L.oy + L.ny * L.dy ||
                        cz + rOutQ < own_zlo(L) || cz - rOutQ > own_zhi(L);
  if (threadIdx.x == 0) {
    set_rotation_matrix(in.th[0][b], in.th[1][b], in.th[2][b], srot);  // particleShape::moveMesh (particleShape.C:291-295)
    for (int q = 0; q < 9; ++q) in.rot[9 * (size_t)b + q] = srot[q];
    if (sc.kind == 2) {  // ellipse::updateCellLists wraps theta.z first (ellipse.C:155-157)
      double tz = in.th[2][b];
      if (tz < 0) tz = xadd(tz, IBM_TWOPI);
      if (tz > IBM_TWOPI) tz = xsub(tz, IBM_TWOPI);
      in.th[2][b] = tz;
    }
    nList = 0; nDirect = 0;
    if (outsideQ) { in.nShell[b] = 0; in.shellOff[b] = 0; in.nPtsOnMesh[b] = 0; in.onLocal[b] = 0; }
  }
  if (outsideQ) return;  // (block-uniform)
  __syncthreads();
  const double thz = in.th[2][b];
  Box bx = bounding_box(L, cx, cy, cz, sc.Rbound);
  const int bi = bx.i1 - bx.i0 + 1, bj = bx.j1 - bx.j0 + 1, bk = bx.k1 - bx.k0 + 1;
  const int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  const int bij = bi * bj;
  const bool fastSplit = nb < (1 << 20);  // float quotients are within one of the truth well beyond this
  const float rbi = 1.0f / (float)max(bi, 1), rbij = 1.0f / (float)max(bij, 1);
  const bool radial = sc.kind != 2;
  const double Rr = sc.d / 2.0, innerR = xsub(Rr, sc.delta);
#define cell_of(t, I, J, K)                                                \
  do {                                                                     \
    int i_, j_, k_;                                                        \
    if (fastSplit) split_box_index((t), bi, bij, rbi, rbij, i_, j_, k_);   \
    else { i_ = (t) % bi; j_ = ((t) / bi) % bj; k_ = (t) / bij; }          \
    I = bx.i0 + i_; J = bx.j0 + j_; K = bx.k0 + k_;                        \
  } while (0)
  // phase A
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    int I, J, K;
    cell_of(t, I, J, K);
    const double Cx = cc(L.ox, I, L.dx), Cy = cc(L.oy, J, L.dy), Cz = cc(L.oz, K + L.kz0, L.dz);
    if (radial) {
      const double r = xmag(xsub(Cx, cx), xsub(Cy, cy), xsub(Cz, cz));  // the test of mark_cell, sphere.C:212-216, cylinder.C:169-173
      if (r >= innerR && r <= Rr) {
        const int s = atomicAdd(&nList, 1);
        if (s < K1_LIST) listCell[s] = t;
      }
    } else {
      int j, k;
      if (mark_cell(sc, cx, cy, cz, thz, Cx, Cy, Cz, j, k)) {
        const int s = atomicAdd(&nList, 1);
        if (s < K1_LIST) { listCell[s] = lat_index(L, I, J, K); listJK[s] = j | (k << 16); }
      }
    }
  }
  __syncthreads();
  const int nShell = nList;
  // phase B: polar bins of the compact list
  if (radial && nShell <= K1_LIST) {
    for (int s = threadIdx.x; s < nShell; s += blockDim.x) {
      int I, J, K, j = 0, k = 0;
      cell_of(listCell[s], I, J, K);
      mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k);
      listCell[s] = lat_index(L, I, J, K);
      listJK[s] = j | (k << 16);
    }
  }
  if (threadIdx.x == 0) {
    const int o = atomicAdd(cursor, nShell);
    segOff = o;
    if (o + nShell > shellCap) { atomicExch(overflow, 1); segOff = -1; }
  }
  __syncthreads();
  const int off = segOff;
  if (off >= 0) {
    if (nShell <= K1_LIST) {
      for (int s = threadIdx.x; s < nShell; s += blockDim.x) { shellCell[off + s] = listCell[s]; shellJK[off + s] = listJK[s]; }
    } else {
      for (int t = threadIdx.x; t < nb; t += blockDim.x) {
        int I, J, K, j, k;
        cell_of(t, I, J, K);
        if (mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k)) {
          const int s = atomicAdd(&nDirect, 1);
          shellCell[off + s] = lat_index(L, I, J, K);
          shellJK[off + s] = j | (k << 16);
        }
      }
    }
  }
  mark_finish(L, sc, in, b, cx, cy, cz, srot, sm, off, nShell);
#undef cell_of
}

// particleShape.C:97-108 and particleIBM.C:607-614 on the WHOLE mesh: nPtsOnMesh holds the point count over all
// slabs, Lg is the global lattice.  On one GPU this reproduces the local status exactly.
__global__ void k1_status(LatticeDev Lg, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nInst) return;
  int I, J, K;
  int cOn = find_cell(Lg, in.c[0][b], in.c[1][b], in.c[2][b], I, J, K) >= 0;
  int nOn = in.nPtsOnMesh[b];
  in.onMesh[b] = (nOn == 0 && !cOn) ? 0 : (nOn == classes[in.cls[b]].N ? 2 : 1);
  in.active[b] = cOn;
}

// ------------------------------------------------------------------------------------------------
// K1 pass 2: ordered fill of the shell list (ascending lattice index == the reference's scan order)
// ------------------------------------------------------------------------------------------------
__global__ void k1_fill(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                        int* __restrict__ shellCell, int* __restrict__ shellJK) {
  __shared__ int warpTot[32];
  __shared__ int base;
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b], thz = in.th[2][b];
  const int off = in.shellOff[b];
  Box bx = bounding_box(L, cx, cy, cz, sc.Rbound);
  int bi = bx.i1 - bx.i0 + 1, bj = bx.j1 - bx.j0 + 1, bk = bx.k1 - bx.k0 + 1;
  int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t0 = 0; t0 < nb; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int j = 0, k = 0, lat = -1;
    bool hit = false;
    if (t < nb) {
      int I = bx.i0 + t % bi, J = bx.j0 + (t / bi) % bj, K = bx.k0 + t / (bi * bj);
      hit = mark_cell(sc, cx, cy, cz, thz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), j, k);
      lat = lat_index(L, I, J, K);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warpTot[warp] = __popc(m);
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int q = 0; q < nw; ++q) { if (q < warp) wbase += warpTot[q]; tot += warpTot[q]; }
    int b0 = base;
    if (hit) {
      shellCell[off + b0 + wbase + pre] = lat;
      shellJK[off + b0 + wbase + pre] = j | (k << 16);
    }
    __syncthreads();
    if (threadIdx.x == 0) base = b0 + tot;
    __syncthreads();
  }
}

// particleCloud::updateI   particleCloud.C:952-969   (field pre-zeroed; lattice order)
__global__ void k1_indicator(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                             double* __restrict__ Ifield) {
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b], thz = in.th[2][b];
  Box bx = bounding_box(L, cx, cy, cz, sc.Rbound);
  int bi = bx.i1 - bx.i0 + 1, bj = bx.j1 - bx.j0 + 1, bk = bx.k1 - bx.k0 + 1;
  int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    int I = bx.i0 + t % bi, J = bx.j0 + (t / bi) % bj, K = bx.k0 + t / (bi * bj);
    double Cx = cc(L.ox, I, L.dx), Cy = cc(L.oy, J, L.dy), Cz = cc(L.oz, K + L.kz0, L.dz);
    double r = shape_r(sc, cx, cy, thz, Cx, Cy);
    if (r > xmag(xsub(cx, Cx), xsub(cy, Cy), xsub(cz, Cz))) Ifield[lat_index(L, I, J, K)] = 1.0;
  }
}

// ------------------------------------------------------------------------------------------------
// gather stencil of one shape point, reference order (particleShape.C:42-91)
// ------------------------------------------------------------------------------------------------
struct Stencil {
  int n;        // faces listed
  int self;     // lattice index of findCell(pt), -1 = off mesh
  int nb[6];    // lattice index of the cell across face f
  int selfOwner;  // bit f set: the point's cell owns face f
  double w[6];
  double W;
};

__device__ __forceinline__ void make_stencil(const LatticeDev& L, const V3& p, Stencil& s, int* nbIdOut = nullptr, int* selfIdOut = nullptr) {
  int I, J, K;
  s.self = find_cell(L, p.x, p.y, p.z, I, J, K);
  s.n = 0; s.W = 0.0; s.selfOwner = 0;
  if (s.self < 0) return;
  int dirs[6], nbId[6], selfId;
  int n = cell_faces(L, I, J, K, dirs, s.nb, nbId, selfId);
  if (selfIdOut) *selfIdOut = selfId;
  double W = 0.0;
  int wi = 0;
  for (int f = 0; f < n; ++f) {
    V3 c = face_centre(L, I, J, K, dirs[f]);
    double diff = xmag(xsub(c.x, p.x), xsub(c.y, p.y), xsub(c.z, p.z));
    if (selfId < nbId[f]) s.selfOwner |= (1 << f);
    if (nbIdOut) nbIdOut[f] = nbId[f];
    if (diff < IBM_SMALL) {
      for (int q = 0; q < f; ++q) s.w[q] = 0.0;
      s.w[f] = 1.0; W = 1.0; wi = f + 1;
      break;
    }
    s.w[f] = xdiv(1.0, diff);
    W = xadd(W, s.w[f]);
    wi = f + 1;
  }
  s.n = wi; s.W = W;
}

// interpolateFromMesh for one point and one scalar component stored as a lattice SoA plane
// (particleIBMTmp.C:42-49 with the linear face value w*(P-N)+N, lambda = 1/2 on the uniform lattice)
__device__ __forceinline__ double gather_ref(const Stencil& s, const double* __restrict__ f) {
  double val = 0.0;
  if (s.self < 0) return val;
  double self = f[s.self];
  for (int q = 0; q < s.n; ++q) {
    double other = f[s.nb[q]];
    double P = ((s.selfOwner >> q) & 1) ? self : other;
    double N = ((s.selfOwner >> q) & 1) ? other : self;
    double face = xadd(xmul(0.5, xsub(P, N)), N);
    val = xadd(val, xmul(xdiv(s.w[q], s.W), face));
  }
  return val;
}

// wFromLocal_ / WFromLocal_ of one shell cell   particleShape.C:110-133
__device__ __forceinline__ void scatter_weights(const ShapeClassDev& sc, const double* rot, double cx, double cy, double cz,
                                                double Cx, double Cy, double Cz, const int* Is, double* w, double& W) {
  W = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) w[q] = 0.0;
  for (int q = 0; q < 8; ++q) {
    V3 p = shape_point(sc, rot, cx, cy, cz, Is[q]);
    double diff = xmag(xsub(Cx, p.x), xsub(Cy, p.y), xsub(Cz, p.z));
    if (diff < IBM_SMALL) {
      for (int r = 0; r < 8; ++r) w[r] = 0.0;
      w[q] = 1.0; W = 1.0;
      break;
    }
    w[q] = xdiv(1.0, diff);
    W = xadd(W, w[q]);
  }
}

// ------------------------------------------------------------------------------------------------
// taps
// ------------------------------------------------------------------------------------------------
__global__ void k_tap_shell(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int inst,
                            const int* __restrict__ shellCell, const int* __restrict__ shellJK, int* cells, int* jk,
                            int* IsOut, int* OsOut, double* wOut, double* WOut) {
  const ShapeClassDev sc = classes[in.cls[inst]];
  const double cx = in.c[0][inst], cy = in.c[1][inst], cz = in.c[2][inst];
  const double* rot = in.rot + 9 * (size_t)inst;
  const int off = in.shellOff[inst], n = in.nShell[inst];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    int lat = shellCell[off + c], pk = shellJK[off + c];
    int j = pk & 0xffff, k = pk >> 16;
    int I = lat % L.nx, J = (lat / L.nx) % L.ny, K = lat / (L.nx * L.ny);
    int Is[8], Os[4];
    neighbour_ids(sc, j, k, Is, Os);
    double w[8], W;
    scatter_weights(sc, rot, cx, cy, cz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), Is, w, W);
    cells[c] = caller_id(L, lat);
    jk[2 * c] = j; jk[2 * c + 1] = k;
    for (int q = 0; q < 8; ++q) { IsOut[8 * c + q] = Is[q]; wOut[8 * c + q] = w[q]; }
    for (int q = 0; q < 4; ++q) OsOut[4 * c + q] = Os[q];
    WOut[c] = W;
  }
}

__global__ void k_tap_points(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int inst, double* pts,
                             int* cell, int* nFaces, int* otherCell, double* wOut, double* WOut) {
  const ShapeClassDev sc = classes[in.cls[inst]];
  const double cx = in.c[0][inst], cy = in.c[1][inst], cz = in.c[2][inst];
  const double* rot = in.rot + 9 * (size_t)inst;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < sc.N; n += gridDim.x * blockDim.x) {
    V3 p = shape_point(sc, rot, cx, cy, cz, n);
    Stencil s;
    int nbId[6], selfId = -1;
    make_stencil(L, p, s, nbId, &selfId);
    pts[3 * n] = p.x; pts[3 * n + 1] = p.y; pts[3 * n + 2] = p.z;
    cell[n] = s.self < 0 ? -1 : selfId;
    nFaces[n] = s.n;
    for (int q = 0; q < 6; ++q) {
      otherCell[6 * n + q] = q < s.n ? nbId[q] : -1;
      wOut[6 * n + q] = q < s.n ? s.w[q] : 0.0;
    }
    WOut[n] = s.W;
  }
}

// ------------------------------------------------------------------------------------------------
// K2/K3 v0: forcing, reference order (particleIBM.C:210-264 + particleIBMTmp.C:26-50)
// one thread per shell cell; the 12 neighbour-point gathers are evaluated in place
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 rigid_velocity(const V3& pt, double cx, double cy, double cz, const V3& om, const V3& v) {
  // (-(pt - center) ^ omega) + v      particleIBM.C:387-390
  double ax = -xsub(pt.x, cx), ay = -xsub(pt.y, cy), az = -xsub(pt.z, cz);
  return V3{xadd(xsub(xmul(ay, om.z), xmul(az, om.y)), v.x), xadd(xsub(xmul(az, om.x), xmul(ax, om.z)), v.y),
            xadd(xsub(xmul(ax, om.y), xmul(ay, om.x)), v.z)};
}

__global__ void k_forcing_ref(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                              const int* __restrict__ shellCell, const int* __restrict__ shellJK,
                              const double* __restrict__ U, const double* __restrict__ Uold, const double* __restrict__ S,
                              double dT, double* __restrict__ F) {
  const int b = blockIdx.x;
  if (b >= nInst) return;
  if (in.onLocal[b] == 0) return;  // particleIBM.C:218-221 (rank-local status, like a decomposed run)
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const V3 vp{in.v[0][b], in.v[1][b], in.v[2][b]}, om{in.om[0][b], in.om[1][b], in.om[2][b]};
  const int off = in.shellOff[b], n = in.nShell[b];
  const size_t C = (size_t)L.nCells;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    int lat = shellCell[off + c], pk = shellJK[off + c];
    int j = pk & 0xffff, k = pk >> 16;
    int I = lat % L.nx, J = (lat / L.nx) % L.ny, K = lat / (L.nx * L.ny);
    int Is[8], Os[4];
    neighbour_ids(sc, j, k, Is, Os);
    double w[8], W;
    scatter_weights(sc, rot, cx, cy, cz, cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz), Is, w, W);
    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
    for (int q = 0; q < 4; ++q) {
      V3 pO = shape_point(sc, rot, cx, cy, cz, Os[q]);
      V3 pI = shape_point(sc, rot, cx, cy, cz, Is[q]);
      Stencil sO, sI;
      make_stencil(L, pO, sO);
      make_stencil(L, pI, sI);
      V3 vv = rigid_velocity(pO, cx, cy, cz, om, vp);
      double ww = xdiv(w[q], W);
      double t;
      t = xadd(xdiv(xsub(xsub(vv.x, gather_ref(sO, U)), gather_ref(sI, Uold)), dT), gather_ref(sO, S));
      Fx = xadd(Fx, xmul(ww, t));
      t = xadd(xdiv(xsub(xsub(vv.y, gather_ref(sO, U + C)), gather_ref(sI, Uold + C)), dT), gather_ref(sO, S + C));
      Fy = xadd(Fy, xmul(ww, t));
      t = xadd(xdiv(xsub(xsub(vv.z, gather_ref(sO, U + 2 * C)), gather_ref(sI, Uold + 2 * C)), dT), gather_ref(sO, S + 2 * C));
      Fz = xadd(Fz, xmul(ww, t));
    }
    for (int q = 4; q < 8; ++q) {
      V3 pS = shape_point(sc, rot, cx, cy, cz, Is[q]);
      Stencil sS;
      make_stencil(L, pS, sS);
      V3 vv = rigid_velocity(pS, cx, cy, cz, om, vp);
      double ww = xdiv(w[q], W);
      double t;
      t = xadd(xdiv(xsub(vv.x, gather_ref(sS, Uold)), dT), gather_ref(sS, S));
      Fx = xadd(Fx, xmul(ww, t));
      t = xadd(xdiv(xsub(vv.y, gather_ref(sS, Uold + C)), dT), gather_ref(sS, S + C));
      Fy = xadd(Fy, xmul(ww, t));
      t = xadd(xdiv(xsub(vv.z, gather_ref(sS, Uold + 2 * C)), dT), gather_ref(sS, S + 2 * C));
      Fz = xadd(Fz, xmul(ww, t));
    }
    // F[i] += Fi (particleIBM.C:262).  Shells of distinct instances share a cell only when bodies
    // overlap; two contributions commute exactly, so the result is order-independent.
    atomicAdd(&F[lat], Fx);
    atomicAdd(&F[C + lat], Fy);
    atomicAdd(&F[2 * C + lat], Fz);
  }
}

// ------------------------------------------------------------------------------------------------
// K4 v0: surface-stress integral, reference order (particleIBM.C:541-581)
// pScaled/tauScaled semantics: face value f -> (f - pRef/rho)*rhoRef and f*rhoRef (particleCloud.C:1213-1227)
// ------------------------------------------------------------------------------------------------
#define STRESS_CHUNK 512
__global__ void k_stress_ref(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst,
                             const double* __restrict__ P, const double* __restrict__ Tau, double pRefOverRho, double rhoRef,
                             double* __restrict__ Tcorr) {
  __shared__ double contrib[STRESS_CHUNK][9];
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const size_t C = (size_t)L.nCells;
  const int nS = sc.nTheta * sc.nk;
  double Fx = 0, Fy = 0, Fz = 0, Tx = 0, Ty = 0, Tz = 0, Ux = 0, Uy = 0, Uz = 0;
  const bool on = in.onLocal[b] != 0;
  for (int s0 = 0; s0 < nS; s0 += STRESS_CHUNK) {
    int s = s0 + threadIdx.x;
    if (threadIdx.x < STRESS_CHUNK && s < nS) {
      // flat order: i outer (theta), k inner
      int i = s / sc.nk, k = s % sc.nk;
      int i2 = i + sc.nTheta * k;  // index2(i,k)
      V3 pt = shape_point(sc, rot, cx, cy, cz, pt_index(sc.nTheta, 1, i, k));
      double tau[6] = {0, 0, 0, 0, 0, 0}, pp = 0.0;
      if (on) {  // interpolateFromMesh returns early for OFF_MESH particles: fields stay zero
        Stencil st;
        make_stencil(L, pt, st);
        // (with a halo a surface point is integrated by the rank that owns the plane of its cell, and by no other)
        if (st.self >= 0 && (!has_halo(L) || owned_plane(L, st.self / (L.nx * L.ny)))) {
          for (int q = 0; q < st.n; ++q) {
            double ww = xdiv(st.w[q], st.W);
            bool own = (st.selfOwner >> q) & 1;
            for (int cmp = 0; cmp < 6; ++cmp) {
              double a = Tau[cmp * C + st.self], o = Tau[cmp * C + st.nb[q]];
              double Pv = own ? a : o, Nv = own ? o : a;
              double face = xmul(xadd(xmul(0.5, xsub(Pv, Nv)), Nv), rhoRef);
              tau[cmp] = xadd(tau[cmp], xmul(ww, face));
            }
            double a = P[st.self], o = P[st.nb[q]];
            double Pv = own ? a : o, Nv = own ? o : a;
            double face = xmul(xsub(xadd(xmul(0.5, xsub(Pv, Nv)), Nv), pRefOverRho), rhoRef);
            pp = xadd(pp, xmul(ww, face));
          }
        }
      }
      // R = mesh_[index2(i,k)] - center (App. B #5 as coded); corrected lever arm uses index(1,i,k)
      V3 m2 = shape_point(sc, rot, cx, cy, cz, i2);
      double Rx = xsub(m2.x, cx), Ry = xsub(m2.y, cy), Rz = xsub(m2.z, cz);
      double Rcx = xsub(pt.x, cx), Rcy = xsub(pt.y, cy), Rcz = xsub(pt.z, cz);
      double sx = sc.Sf[3 * i2], sy = sc.Sf[3 * i2 + 1], sz = sc.Sf[3 * i2 + 2];
      // Sfi = rotationMatrix_ & Sf_   (particleShapeI.H:35-38)
      double Sx = xdot(rot[0], rot[1], rot[2], sx, sy, sz), Sy = xdot(rot[3], rot[4], rot[5], sx, sy, sz),
             Sz = xdot(rot[6], rot[7], rot[8], sx, sy, sz);
      double magSf = xmag(Sx, Sy, Sz);
      double nx = xdiv(Sx, magSf), ny = xdiv(Sy, magSf), nz = xdiv(Sz, magSf);
      // F = p*Sf_ + (tau & Sfi)    (App. B #6 as coded: unrotated Sf_ in the pressure term)
      double fx = xadd(xmul(pp, sx), xdot(tau[0], tau[1], tau[2], Sx, Sy, Sz));
      double fy = xadd(xmul(pp, sy), xdot(tau[1], tau[3], tau[4], Sx, Sy, Sz));
      double fz = xadd(xmul(pp, sz), xdot(tau[2], tau[4], tau[5], Sx, Sy, Sz));
      double fn = xdot(fx, fy, fz, nx, ny, nz);
      double ftx = xsub(fx, xmul(fn, nx)), fty = xsub(fy, xmul(fn, ny)), ftz = xsub(fz, xmul(fn, nz));
      double* o = contrib[threadIdx.x];
      o[0] = fx; o[1] = fy; o[2] = fz;
      // Ft ^ R
      o[3] = xsub(xmul(fty, Rz), xmul(ftz, Ry)); o[4] = xsub(xmul(ftz, Rx), xmul(ftx, Rz)); o[5] = xsub(xmul(ftx, Ry), xmul(fty, Rx));
      o[6] = xsub(xmul(fty, Rcz), xmul(ftz, Rcy)); o[7] = xsub(xmul(ftz, Rcx), xmul(ftx, Rcz)); o[8] = xsub(xmul(ftx, Rcy), xmul(fty, Rcx));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int m = min(STRESS_CHUNK, nS - s0);
      for (int q = 0; q < m; ++q) {
        const double* o = contrib[q];
        Fx = xadd(Fx, o[0]); Fy = xadd(Fy, o[1]); Fz = xadd(Fz, o[2]);
        Tx = xsub(Tx, o[3]); Ty = xsub(Ty, o[4]); Tz = xsub(Tz, o[5]);
        Ux = xsub(Ux, o[6]); Uy = xsub(Uy, o[7]); Uz = xsub(Uz, o[8]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    in.F[0][b] = Fx; in.F[1][b] = Fy; in.F[2][b] = Fz;
    in.T[0][b] = Tx; in.T[1][b] = Ty; in.T[2][b] = Tz;
    if (Tcorr) { Tcorr[3 * b] = Ux; Tcorr[3 * b + 1] = Uy; Tcorr[3 * b + 2] = Uz; }
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
void launch_k1_count(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst) {
  if (nInst > 0) k1_count<<<nInst, 128, 0, st>>>(L, classes, in, nInst);
}
void launch_k1_far(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst, Schedule& s,
                   int* nearFlag, int* nearOff, int* nearList, int* dNearCount) {
  if (nInst <= 0) return;
  k1_far<<<(nInst + 256) / 256, 256, 0, st>>>(L, classes, in, nInst, nearFlag);
  launch_exclusive_scan_total(st, s, nearFlag, nearOff, nInst, dNearCount);
  k1_near_list<<<(nInst + 255) / 256, 256, 0, st>>>(nInst, nearFlag, nearOff, nearList);
}

void launch_k1_mark(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                    int* shellCell, int* shellJK, int shellCap, int* cursor, int* overflow, int variant, const int* nearList,
                    const int* nearCount) {
  if (nInst <= 0) return;
  if (variant == 1) k1_mark2<<<nInst, 128, 0, st>>>(L, classes, in, nInst, shellCell, shellJK, shellCap, cursor, overflow, nearList, nearCount);
  else k1_mark<<<nInst, 128, 0, st>>>(L, classes, in, nInst, shellCell, shellJK, shellCap, cursor, overflow);
}
void launch_k1_status(cudaStream_t st, const LatticeDev& Lg, const ShapeClassDev* classes, const InstanceSoA& in, int nInst) {
  if (nInst > 0) k1_status<<<(nInst + 127) / 128, 128, 0, st>>>(Lg, classes, in, nInst);
}
void launch_k1_fill(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                    int* shellCell, int* shellJK) {
  if (nInst > 0) k1_fill<<<nInst, 128, 0, st>>>(L, classes, in, nInst, shellCell, shellJK);
}
void launch_k1_indicator(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                         double* Ifield) {
  if (nInst > 0) k1_indicator<<<nInst, 128, 0, st>>>(L, classes, in, nInst, Ifield);
}
void launch_tap_shell(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int inst,
                      const int* shellCell, const int* shellJK, int* cells, int* jk, int* Is, int* Os, double* w, double* W) {
  k_tap_shell<<<64, 128, 0, st>>>(L, classes, in, inst, shellCell, shellJK, cells, jk, Is, Os, w, W);
}
void launch_tap_points(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int inst,
                       double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W) {
  k_tap_points<<<64, 128, 0, st>>>(L, classes, in, inst, pts, cell, nFaces, otherCell, w, W);
}
void launch_forcing_ref(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                        const int* shellCell, const int* shellJK, const double* U, const double* Uold, const double* S,
                        double dT, double* F) {
  if (nInst > 0) k_forcing_ref<<<nInst, 64, 0, st>>>(L, classes, in, nInst, shellCell, shellJK, U, Uold, S, dT, F);
}
void launch_stress_ref(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr) {
  if (nInst > 0) k_stress_ref<<<nInst, STRESS_CHUNK, 0, st>>>(L, classes, in, nInst, P, Tau, pRefOverRho, rhoRef, Tcorr);
}

}  // namespace ibm
