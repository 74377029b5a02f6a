// Device-side data model of the B200 immersed-boundary path.
//
//  * LatticeDev   — the fvMesh as a uniform hex lattice; geometry is evaluated, never stored
//                   (replaces polyMesh::cellCentres/faceCentres/cells/findCell, particleShape.C:39-66).
//  * ShapeClassDev — one entry per distinct (kind, d|a,b, nTheta, nk, delta): the body-fitted three-layer
//                   point template centeredMesh_ and the area vectors Sf_ (particleShape.H:119-129),
//                   built once on the host with the C library's libm exactly as the reference's
//                   discretize()/calcSf() do, shared by every particle of the class.
//  * InstanceSoA  — Cloud<particleIBM> as structure-of-arrays (one slot per particle instance, cloud order).
//
// "x*" helpers are IEEE round-to-nearest operations that the compiler may not contract into FMAs: the
// membership arithmetic (cell marking, point->cell) must be bit-identical to an x86-64 build of the
// reference, which has no FMA contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IBM_SMALL 1.0e-15
#define IBM_PI 3.14159265358979323846
#define IBM_TWOPI (2.0 * IBM_PI)

namespace ibm {

struct LatticeDev {
  int nx, ny, nz;
  int nCells;
  double ox, oy, oz;
  double dx, dy, dz;
  const int* latToCell;  // nullptr = identity
  const int* cellToLat;  // nullptr = identity
  int kz0;               // slab decomposition: global z-plane of local plane 0 (nz is the LOCAL plane count);
                         // geometry is always evaluated with the global plane index K + kz0, exactly as a
                         // decomposed OpenFOAM mesh keeps the global cell centres
  int ownK0, ownK1;      // local planes [ownK0, ownK1) are OWNED by this rank, the others are a read-only halo of the
                         // neighbouring slabs (ibm_set_slab_halo): shell cells are marked, shape points counted and surface
                         // points integrated in owned planes only; gathers read the halo.  Without a halo: 0, nz.
};
// z-extent of the owned planes; does lattice cell latIdx lie in one
__device__ __forceinline__ double own_zlo(const LatticeDev& L) { return L.oz + (L.kz0 + L.ownK0) * L.dz; }
__device__ __forceinline__ double own_zhi(const LatticeDev& L) { return L.oz + (L.kz0 + L.ownK1) * L.dz; }
__device__ __forceinline__ bool has_halo(const LatticeDev& L) { return L.ownK0 > 0 || L.ownK1 < L.nz; }
__device__ __forceinline__ bool owned_plane(const LatticeDev& L, int K) { return K >= L.ownK0 && K < L.ownK1; }

struct ShapeClassDev {
  int kind, nTheta, nk, N;
  double d, a, b, l, delta;
  double Rbound;   // bounding radius: d/2, or max(a,b) for the ellipse
  double V, I, A;  // volume, inertia per unit mass (as coded), frontal area
  const double* centered;  // [N][3]   centeredMesh_
  const double* Sf;        // [nTheta*nk][3]
  const double* cosT;      // [nTheta]  cos(j*dTheta)
  const double* sinT;      // [nTheta]
  const double* sinP;      // [nk]      sin(k*dPhi)   (sphere)
  const double* cosP;      // [nk]
  const double* Rtab;      // [3][nTheta] layer radii (ellipse: rho(theta_j)+delta_i; else R+delta_i)
  const double* ck;        // [nk]      polar band factor of Sf_ (sphere.C:142-157); 1 for 2-D shapes
  double dTheta;           // 2*pi/nTheta
};

struct InstanceSoA {
  int* index; int* copy; int* active; int* onMesh; int* cls;
  double* c[3]; double* th[3]; double* v[3]; double* om[3];
  double* rho; double* age;
  double* F[3]; double* T[3];
  double* rot;        // [cap][9] rotationMatrix_
  int* nShell; int* shellOff;
  int* nPtsOnMesh;
  int* onLocal;       // rank-local location status (== onMesh on one GPU): which instances loops 1-4 skip on this slab
};

__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }
// OpenFOAM magSqr / mag / inner product, left-associated
__device__ __forceinline__ double xmagSqr(double x, double y, double z) {
  return xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));
}
__device__ __forceinline__ double xmag(double x, double y, double z) { return xsqrt(xmagSqr(x, y, z)); }
__device__ __forceinline__ double xdot(double ax, double ay, double az, double bx, double by, double bz) {
  return xadd(xadd(xmul(ax, bx), xmul(ay, by)), xmul(az, bz));
}

struct V3 { double x, y, z; };

__device__ __forceinline__ int lat_index(const LatticeDev& L, int I, int J, int K) { return I + L.nx * (J + L.ny * K); }
__device__ __forceinline__ int caller_id(const LatticeDev& L, int lat) { return L.latToCell ? L.latToCell[lat] : lat; }

// cell centre of the mini-mesh definition: o + (I + 0.5)*d  (mul then add)
__device__ __forceinline__ double cc(double o, int I, double d) { return xadd(o, xmul((double)I + 0.5, d)); }
__device__ __forceinline__ double fc(double o, int I, double d) { return xadd(o, xmul((double)I, d)); }  // face plane I

// DEFINED findCell: floor((pt - origin)/delta) with bounds check.  Returns lattice index or -1.
__device__ __forceinline__ int find_cell(const LatticeDev& L, double x, double y, double z, int& I, int& J, int& K) {
  double tx = floor(xdiv(xsub(x, L.ox), L.dx));
  double ty = floor(xdiv(xsub(y, L.oy), L.dy));
  double tz = floor(xdiv(xsub(z, L.oz), L.dz)) - (double)L.kz0;
  if (!(tx >= 0.0 && tx < (double)L.nx && ty >= 0.0 && ty < (double)L.ny && tz >= 0.0 && tz < (double)L.nz)) return -1;
  I = (int)tx; J = (int)ty; K = (int)tz;
  return lat_index(L, I, J, K);
}

// Same result as find_cell without the three fp64 divisions on the common path: floor(q*inv) equals
// floor(fl(q/d)) unless q/d lies within a few ulp of an integer; those cases fall back to the exact division.
__device__ __forceinline__ bool floor_div_exact(double q, double d, double inv, int n, int& I, int off = 0) {
  double t = q * inv;
  double f = floor(t);
  double fr = t - f;
  if (fr < 1e-9 || fr > 1.0 - 1e-9) f = floor(xdiv(q, d));
  f -= (double)off;
  if (!(f >= 0.0 && f < (double)n)) return false;
  I = (int)f;
  return true;
}
__device__ __forceinline__ int find_cell_fast(const LatticeDev& L, double idx, double idy, double idz, double x, double y, double z,
                                              int& I, int& J, int& K) {
  if (!floor_div_exact(xsub(x, L.ox), L.dx, idx, L.nx, I)) return -1;
  if (!floor_div_exact(xsub(y, L.oy), L.dy, idy, L.ny, J)) return -1;
  if (!floor_div_exact(xsub(z, L.oz), L.dz, idz, L.nz, K, L.kz0)) return -1;
  return lat_index(L, I, J, K);
}

// mesh_[n] = rotate(centeredMesh_[n]) + center   (particleShape.C:291-302; ellipse.C:61-64)
__device__ __forceinline__ V3 shape_point(const ShapeClassDev& sc, const double* rot, double cx, double cy, double cz, int n) {
  double px = sc.centered[3 * n], py = sc.centered[3 * n + 1], pz = sc.centered[3 * n + 2];
  if (sc.kind == 2) {
    double rx = xdot(rot[0], rot[1], rot[2], px, py, pz);
    double ry = xdot(rot[3], rot[4], rot[5], px, py, pz);
    double rz = xdot(rot[6], rot[7], rot[8], px, py, pz);
    px = rx; py = ry; pz = rz;
  }
  return V3{xadd(px, cx), xadd(py, cy), xadd(pz, cz)};
}

// radius of the shape toward pt: sphere.H:115-118, cylinder.H:117-120, ellipse.C:222-227
__device__ __forceinline__ double shape_r(const ShapeClassDev& sc, double cx, double cy, double thz, double px, double py) {
  if (sc.kind != 2) return sc.d / 2.0;
  double th = xsub(atan2(xsub(py, cy), xsub(px, cx)), thz);
  double c = xdiv(cos(th), sc.a), s = xdiv(sin(th), sc.b);
  return xsqrt(xdiv(1.0, xadd(xmul(c, c), xmul(s, s))));
}

// Neighbour directions: 0 +x, 1 -x, 2 +y, 3 -y, 4 +z, 5 -z
__device__ __forceinline__ bool nb_in_range(const LatticeDev& L, int I, int J, int K, int d, int& I2, int& J2, int& K2) {
  I2 = I + (d == 0) - (d == 1);
  J2 = J + (d == 2) - (d == 3);
  K2 = K + (d == 4) - (d == 5);
  return I2 >= 0 && I2 < L.nx && J2 >= 0 && J2 < L.ny && K2 >= 0 && K2 < L.nz;
}

// cells()[c] restricted to internal faces, in OpenFOAM order: faces to higher-numbered neighbours
// (ascending neighbour id) then faces to lower-numbered neighbours (ascending id).  Writes direction
// codes and neighbour lattice indices; returns the face count.
__device__ __forceinline__ int cell_faces(const LatticeDev& L, int I, int J, int K, int* dirs, int* nbLat, int* nbId, int& selfId) {
  selfId = caller_id(L, lat_index(L, I, J, K));
  int n = 0;
  int idv[6], dv[6], lv[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    int I2, J2, K2;
    if (nb_in_range(L, I, J, K, d, I2, J2, K2)) {
      int l = lat_index(L, I2, J2, K2);
      idv[n] = caller_id(L, l); dv[n] = d; lv[n] = l; ++n;
    }
  }
  // insertion sort by key: (id < self) then id
  for (int a = 1; a < n; ++a) {
    int id = idv[a], d = dv[a], l = lv[a];
    long long key = ((long long)(id < selfId) << 32) | (unsigned)id;
    int b = a - 1;
    while (b >= 0) {
      long long kb = ((long long)(idv[b] < selfId) << 32) | (unsigned)idv[b];
      if (kb <= key) break;
      idv[b + 1] = idv[b]; dv[b + 1] = dv[b]; lv[b + 1] = lv[b]; --b;
    }
    idv[b + 1] = id; dv[b + 1] = d; lv[b + 1] = l;
  }
  for (int a = 0; a < n; ++a) { dirs[a] = dv[a]; nbLat[a] = lv[a]; nbId[a] = idv[a]; }
  return n;
}

// face centre of cell (I,J,K) in direction d
__device__ __forceinline__ V3 face_centre(const LatticeDev& L, int I, int J, int K, int d) {
  V3 f{cc(L.ox, I, L.dx), cc(L.oy, J, L.dy), cc(L.oz, K + L.kz0, L.dz)};
  if (d == 0) f.x = fc(L.ox, I + 1, L.dx);
  if (d == 1) f.x = fc(L.ox, I, L.dx);
  if (d == 2) f.y = fc(L.oy, J + 1, L.dy);
  if (d == 3) f.y = fc(L.oy, J, L.dy);
  if (d == 4) f.z = fc(L.oz, K + L.kz0 + 1, L.dz);
  if (d == 5) f.z = fc(L.oz, K + L.kz0, L.dz);
  return f;
}

// particleShape::index / index2   particleShape.H:164-180
__device__ __forceinline__ int pt_index(int nTheta, int i, int j, int k) { return i + j * 3 + k * 3 * nTheta; }

// Is_[0..7] / Os_[0..3] of a shell cell with bin (j,k)   particleShape.C:136-183
__device__ __forceinline__ void neighbour_ids(const ShapeClassDev& sc, int j, int k, int* Is, int* Os) {
  int nT = sc.nTheta;
  int jp1 = (j + 1) % nT;
  if (sc.nk != 1) {
    if (k == sc.nk - 1) {
#pragma unroll
      for (int q = 0; q < 8; ++q) Is[q] = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) Os[q] = 0;
      return;
    }
    Is[0] = pt_index(nT, 0, j, k); Is[1] = pt_index(nT, 0, jp1, k); Is[2] = pt_index(nT, 0, j, k + 1); Is[3] = pt_index(nT, 0, jp1, k + 1);
    Is[4] = pt_index(nT, 1, j, k); Is[5] = pt_index(nT, 1, jp1, k); Is[6] = pt_index(nT, 1, j, k + 1); Is[7] = pt_index(nT, 1, jp1, k + 1);
    Os[0] = pt_index(nT, 2, j, k); Os[1] = pt_index(nT, 2, jp1, k); Os[2] = pt_index(nT, 2, j, k + 1); Os[3] = pt_index(nT, 2, jp1, k + 1);
  } else {
    Is[0] = pt_index(nT, 0, j, k); Is[1] = pt_index(nT, 0, jp1, k); Is[2] = Is[0]; Is[3] = Is[1];
    Is[4] = pt_index(nT, 1, j, k); Is[5] = pt_index(nT, 1, jp1, k); Is[6] = Is[4]; Is[7] = Is[5];
    Os[0] = pt_index(nT, 2, j, k); Os[1] = pt_index(nT, 2, jp1, k); Os[2] = Os[0]; Os[3] = Os[1];
  }
}

}  // namespace ibm
