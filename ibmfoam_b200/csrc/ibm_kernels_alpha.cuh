// Solid volume fraction alpha_s (EXTENSION - no counterpart in the reference).
//
// BASELINE.json's north star lists "solid volume-fraction computation over the mesh cells each particle overlaps" next
// to the marking loop; the reference only has the binary indicator I.ibm (1 if the cell CENTRE is inside a particle,
// particleCloud.C:952-969).  This kernel computes the fractional field the wording asks for:
//   alpha_s[cell] = min(1, sum over instances of  (sub-cell sample points inside the body) / nSub^D )
// with nSub^D sample points per cell on a regular sub-lattice (D = 3, or 2 on a one-cell-thick lattice), the same
// definition as oracle/alpha_oracle.py, so the result is bit-exact (the counts are integers).  Cells that are entirely
// inside / outside the body are classified by the distance of their nearest / farthest point, with a relative margin of
// 1e-9 so that rounding can never disagree with the sampling.  One CTA per instance over the cells of its bounding box.
// Also compiled for the CPU emulation (tests/test_emu_v5.py).
#pragma once
#include "ibm_device.cuh"

namespace ibm {

#define ALPHA_THREADS 128

// is the point inside the body?  sphere / cylinder: |x - c| < R (in the plane for the cylinder);
// ellipse: the point rotated back into the body frame satisfies (x'/a)^2 + (y'/b)^2 < 1  (a, b semi-axes, ellipse.C:135-137)
__device__ __forceinline__ bool alpha_inside(const ShapeClassDev& sc, const double* rot, double cx, double cy, double cz, double x,
                                             double y, double z) {
  const double dx = xsub(x, cx), dy = xsub(y, cy), dz = xsub(z, cz);
  if (sc.kind == 0) { const double R = sc.d / 2.0; return xmagSqr(dx, dy, dz) < xmul(R, R); }
  if (sc.kind == 1) { const double R = sc.d / 2.0; return xadd(xmul(dx, dx), xmul(dy, dy)) < xmul(R, R); }
  const double bx = xadd(xmul(rot[0], dx), xmul(rot[3], dy)), by = xadd(xmul(rot[1], dx), xmul(rot[4], dy));  // R^T (dx, dy)
  const double u = xdiv(bx, sc.a), v = xdiv(by, sc.b);
  return xadd(xmul(u, u), xmul(v, v)) < 1.0;
}

__global__ void __launch_bounds__(ALPHA_THREADS)
k_solid_fraction(LatticeDev L, const ShapeClassDev* __restrict__ classes, InstanceSoA in, int nInst, int nSub, double* __restrict__ alpha) {
  const int b = blockIdx.x;
  if (b >= nInst) return;
  const ShapeClassDev sc = classes[in.cls[b]];
  const double cx = in.c[0][b], cy = in.c[1][b], cz = in.c[2][b];
  const double* rot = in.rot + 9 * (size_t)b;
  const double R = sc.Rbound;
  const bool flat = (sc.kind != 0);  // cylinders / ellipses span the lattice thickness: sample in the plane
  const int i0 = max((int)floor((cx - R - L.ox) / L.dx), 0), i1 = min((int)floor((cx + R - L.ox) / L.dx), L.nx - 1);
  const int j0 = max((int)floor((cy - R - L.oy) / L.dy), 0), j1 = min((int)floor((cy + R - L.oy) / L.dy), L.ny - 1);
  const int k0 = flat ? 0 : max((int)floor((cz - R - L.oz) / L.dz) - L.kz0, 0);
  const int k1 = flat ? L.nz - 1 : min((int)floor((cz + R - L.oz) / L.dz) - L.kz0, L.nz - 1);
  const int bi = i1 - i0 + 1, bj = j1 - j0 + 1, bk = k1 - k0 + 1;
  const int nb = (bi > 0 && bj > 0 && bk > 0) ? bi * bj * bk : 0;
  const double inv = 1.0 / (double)nSub;
  const int nz = flat ? 1 : nSub;
  const double total = (double)(nSub * nSub * nz);
  const bool iso = (sc.kind != 2);
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    const int I = i0 + t % bi, J = j0 + (t / bi) % bj, K = k0 + t / (bi * bj);
    const double Cx = cc(L.ox, I, L.dx), Cy = cc(L.oy, J, L.dy), Cz = cc(L.oz, K + L.kz0, L.dz);
    int count = -1;
    if (iso) {  // whole-cell shortcuts for round bodies
      const double ax = fabs(Cx - cx), ay = fabs(Cy - cy), az = flat ? 0.0 : fabs(Cz - cz);
      const double hx = 0.5 * L.dx, hy = 0.5 * L.dy, hz = flat ? 0.0 : 0.5 * L.dz;
      const double fx = ax + hx, fy = ay + hy, fz = az + hz;                                  // farthest corner
      const double nx = fmax(ax - hx, 0.0), ny = fmax(ay - hy, 0.0), nzr = fmax(az - hz, 0.0);  // nearest point
      const double R2 = R * R;
      if (fx * fx + fy * fy + fz * fz < R2 * (1.0 - 1e-9)) count = nSub * nSub * nz;
      else if (nx * nx + ny * ny + nzr * nzr > R2 * (1.0 + 1e-9)) count = 0;
    }
    if (count < 0) {
      count = 0;
      for (int c = 0; c < nz; ++c) {
        const double z = flat ? cz : xadd(Cz, xmul(xsub(xmul((double)c + 0.5, inv), 0.5), L.dz));
        for (int q = 0; q < nSub; ++q) {
          const double y = xadd(Cy, xmul(xsub(xmul((double)q + 0.5, inv), 0.5), L.dy));
          for (int p = 0; p < nSub; ++p) {
            const double x = xadd(Cx, xmul(xsub(xmul((double)p + 0.5, inv), 0.5), L.dx));
            count += alpha_inside(sc, rot, cx, cy, cz, x, y, z) ? 1 : 0;
          }
        }
      }
    }
    if (count > 0) atomicAdd(&alpha[lat_index(L, I, J, K)], (double)count / total);
  }
}

// overlapping bodies: the fraction of a cell cannot exceed 1
__global__ void k_clamp_unit(double* __restrict__ a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (a[i] > 1.0) a[i] = 1.0;
}

}  // namespace ibm
