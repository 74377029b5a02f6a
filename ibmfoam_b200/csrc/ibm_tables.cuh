// Shape points of spheres / cylinders from the class's trig tables held in shared memory: the same two roundings
// as the host-built template ((r*cos(theta))*sin(phi), sphere.C:184-190), so coordinates are bit-identical to
// centeredMesh_ + centre without touching global memory.
#pragma once
#include "ibm_device.cuh"

namespace ibm {

#define T4_NT 256
#define T4_NK 128
#define SBOX4 1792       // cells per staged plane (stress): 12^3 padded
#define S4_THREADS 256

struct Tab4 {
  double cosT[T4_NT], sinT[T4_NT];
  double sinP[T4_NK], cosP[T4_NK], ck[T4_NK];
};

__device__ __forceinline__ bool tab4_ok(const ShapeClassDev& sc) { return sc.kind != 2 && sc.nTheta <= T4_NT && sc.nk <= T4_NK; }

__device__ __forceinline__ void tab4_load(const ShapeClassDev& sc, Tab4& t) {
  for (int j = threadIdx.x; j < sc.nTheta; j += blockDim.x) { t.cosT[j] = sc.cosT[j]; t.sinT[j] = sc.sinT[j]; }
  for (int k = threadIdx.x; k < sc.nk; k += blockDim.x) { t.sinP[k] = sc.sinP[k]; t.cosP[k] = sc.cosP[k]; t.ck[k] = sc.ck[k]; }
}

// mesh_[index(i,j,k)] of a sphere / cylinder from the tables: the same two roundings as the host template
__device__ __forceinline__ V3 tab4_point(const ShapeClassDev& sc, const Tab4& t, double R, double cx, double cy, double cz, int layer,
                                         int j, int k) {
  const double r = xadd(R, layer == 0 ? -1 * sc.delta : (layer == 1 ? 0.0 : sc.delta));  // R + delta[i]   (sphere.C:167,173)
  if (sc.kind == 0)
    return V3{xadd(xmul(xmul(r, t.cosT[j]), t.sinP[k]), cx), xadd(xmul(xmul(r, t.sinT[j]), t.sinP[k]), cy), xadd(xmul(r, t.cosP[k]), cz)};
  return V3{xadd(xmul(r, t.cosT[j]), cx), xadd(xmul(r, t.sinT[j]), cy), xadd(0.0, cz)};
}
__device__ __forceinline__ V3 tab4_point_id(const ShapeClassDev& sc, const Tab4& t, double R, double cx, double cy, double cz, int id) {
  const int layer = id % 3, j = (id / 3) % sc.nTheta, k = id / (3 * sc.nTheta);
  return tab4_point(sc, t, R, cx, cy, cz, layer, j, k);
}


}  // namespace ibm
