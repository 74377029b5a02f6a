// Loop 5 — collisions, wall hits, periodic copies, sub-stepped integration, force combination.
// (translation unit compiled with -fmad=false: every operation below is a plain IEEE +,-,*,/,sqrt in the
//  reference's evaluation order, so sphere / cylinder trajectories are bit-identical to an x86-64 build.)
//
//   particleCloud::computeCollisions  particleCloud.C:482-581   -> k5_bin + k5_contacts + k5_merge
//   particleIBM::collision            particleIBM.C:41-205      -> pair_collision()
//   particleCloud::computeWallHits    particleCloud.C:584-656   -> k5_walls (+ k5_merge)
//   particleCloud::computeCyclicHits  particleCloud.C:659-937   -> k5_cyclic_detect + scan + k5_append
//   particleIBM::solve                particleIBM.C:393-450     -> k5_solve
//   particleCloud::setPositions       particleCloud.C:316-444   -> k5_set_positions
//   particleCloud::combineForces      particleCloud.C:171-313   -> k_combine_forces
//   particleCloud::solve (loop+log)   particleCloud.C:1247-1324 -> motion_advance
//
// The reference keeps dense nParticles x nParticles force tables and visits all ordered pairs; here a uniform
// cell list (bin edge >= largest diameter, fixed-capacity bins refilled every sub-step) yields the same
// contact set, stored sparsely per instance and summed in partner-index order (the reference's sum() order).
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstring>

#include "ibm_ctx.h"

namespace cg = cooperative_groups;

namespace ibm {

#define CONTACT_CAP 24

struct PatchDev {
  int type, side, neighb, nFaces;
  const double *Cf, *Sf, *magSf;
  int latticeOrdered;  // faces in transverse lattice order (fastest index first): candidate faces by index box
  int nA, nB;          // transverse extents
  double lo[3], hi[3]; // bounding box of the face centres (cheap rejection: |centre - Cf| >= distance to the box)
  int planar;          // every face has bit-identical Sf and magSf: Sf/magSf is one vector for the whole patch
  double nrm[3];       // that vector
  int uniformShift;    // cyclic pair that is a pure translation: Cf_neighbour[f] - Cf[f] is one vector for every face
  double shift[3];     // that vector (the image of a centre c is c + shift up to rounding, whichever face is hit)
};

struct MotionDev {
  int cap, nParticles, nWalls, nCyc;
  int* nInst;  // device-side instance count (appends happen on the device)
  // cell list
  double bx0, by0, bz0, binEdge;
  int nbx, nby, nbz;
  int binCap; int* binCnt; int* binItems; double* binPos;   // fixed-capacity cell list (entry = instance + its centre)
  int* pairCount; int* pairI; int* pairJ; int pairCap;       // near-contact candidate pairs of the current sub-step
  int* cPos;                                                 // cloud position of the partner behind each row entry
  unsigned* keys; unsigned* keysSorted; int* vals; int* valsSorted;  // scratch for the same-index list build
  // sparse contact tables (sorted by partner index)
  int* cCnt; int* cIdx; double* cF; double* cT;
  int* cCnt2; int* cIdx2; double* cF2; double* cT2;
  int* overflow;
  // wall tables [cap][nWalls][3]
  double* Fw; double* Tw; double* Fw2; double* Tw2;
  // same-index groups in cloud order
  int* idxHead; int* idxTail; int* instNext;
  // originals (origParticles_) and copy counters
  const int* origCls; const double* origRho; int* maxCopies;
  // cyclic hits
  int* candCount; int* candKey; int candCap;      // (patch, instance) pairs within reach of a cyclic patch
  int* hitCount; int* hitKey; double* hitCenter; int* blocksDone;  // compact list of new periodic copies (key = patchSlot*cap + instance)
  const PatchDev* patches; const int* wallPatch; const int* cycPatch;
  double gx, gy, gz, e;
  int moving, rotating;
  // EXTENSION (no counterpart in the reference): soft-sphere contact law instead of the hard-sphere impulse
  int model;                       // 0 = particleCloud.C:505-521 as coded, 1 = linear spring - dashpot - Coulomb slider
  double kn, gn, gt, mu;           // normal stiffness, normal / tangential damping, friction coefficient
  // quiet / active split of a fluid step (option active_loop5): see the block comment above k5q_bounds
  int activeMode;                  // 1 while the split is in force (append_link then keeps the active list up to date)
  int* act; int* nAct;             // instances that can take part in a contact, a wall / patch hit or a copy group this step
  int* quiet;                      // 1 = integrates alone, all sub-steps in one kernel
  double* c0;                      // [3][cap] centre at the start of the step (at creation for copies appended later)
  unsigned long long* vaBits;      // bit pattern of the largest own displacement bound s_i over the instances
  double* budget;                  // S = safety * max s_i: no instance is assumed to move farther than this during the step
  double* sOwn; double* uMax;      // [cap] own ballistic bound s_i; its maximum over the instance and its possible partners
  double* bud;                     // [cap] B_i: displacement budget of instance i (two-hop maximum of s, times safety)
  int* violated;                   // an active instance moved farther than its B_i: the step is redone without the split
  int* noSplit;                    // the start configuration defeats the classification (see k5q_ghosts): everything is active
  double safety, check;            // budget factor; fraction of B_i the check allows (1 except in tests of the fallback)
  // cluster mode of the split (option cluster_loop5, k5c_*): the active instances are partitioned into groups that cannot
  // interact with one another during the step; every group runs ALL its sub-steps inside one thread block
  int* edgeA; int* edgeB; int* nEdge; int edgeCap;   // pairs of active instances that may interact this step
  int* label;                      // [cap] component root of an active instance
  int* compCnt; int* compOff; int* compCur;  // [cap] members per root, exclusive offsets, placement cursors
  int* sorted;                     // [cap] active instances grouped by component
  int* ctaBegin; int* ctaEnd;      // [ctaBins] range of `sorted` a block of k5c_run owns
  int ctaBins;
  int* nextSlot;                   // next free instance slot (copies created by the blocks take provisional slots)
  int* evSub; int* evKey;          // [cap] per provisional slot: sub-step of creation, patchSlot * cap + position of the source
  unsigned long long* clusterStat; // [2] block sub-steps in total / on the general path (diagnostics of the last steps)
  int* clusterFail;                // a block ran out of room (members, candidates, hits): the step is redone unsplit
  // multi-GPU, cluster mode (option owned_loop5): loop 5 is partitioned by subdomain.  An instance that integrates alone
  // belongs to the rank whose z-slab holds its centre at the start of the step, a block of k5c_run to the rank that owns its
  // first member; every rank advances what it owns and the new states (and new periodic copies) are exchanged once per
  // fluid step (k5x_pack / all-reduce / k5x_unpack) - the counterpart of particleCloud.C:316-444, 838-899
  int owned, rank, nRanks;
  int slabK0[17];                  // first global z-plane of every rank's slab; [nRanks] = planes of the global mesh
  double ownOz, ownDz; int ownNz;
  int* mine;                       // [cap] 1 = this rank advanced the instance in this step
};

struct Motion {
  MotionDev d{};
  std::vector<void*> allocs;
  void* cubTemp = nullptr;
  size_t cubTempBytes = 0;
  PatchDev* dPatches = nullptr;
  bool groupsValid = false;
  double* dDts = nullptr;
  int fusedBlocks = 0;
  int fusedActBlocks = 0;       // upper bound of the grid of k5a_fused (co-residency)
  // CUDA graph of the unfused sub-step sequence, cached for one schedule of local time steps
  cudaGraphExec_t graphExec = nullptr;
  std::vector<double> graphDts;
  // A captured graph freezes its kernel arguments (the InstanceSoA pointers, which ctx_update swaps with the second set
  // when it compacts the instance list) and its grid sizes (k5_pair_force is sized from the instance count): it is only
  // replayed while both still hold.
  bool graphValid = false;
  int graphNInst = 0;
  long long graphLaunches = 0;
  int* dIdxKeys = nullptr; int* dIdxKeys2 = nullptr;
  double* dStats = nullptr;
  // quiet / active split: what is needed to redo a step whose displacement budget was exceeded
  double* snap = nullptr;       // [12][cap] c, theta, v, omega at the start of the step
  int* snapActive = nullptr;    // [cap]
  int* snapMaxCopies = nullptr; // [nParticles]
  bool graphActive = false;     // the captured graph is of the split sequence
  double graphSafety = 0.0;     // ... with this safety factor (MotionDev is captured by value)
  double graphDt = 0.0;         // ... and this fluid step (an argument of k5q_bounds)
  bool splitAllowed = true;     // every cyclic patch is a pure translation (k5q_ghosts needs the image of a centre)
  long long nFallbacks = 0;
  int lastActive = 0;           // active instances of the previous split step (sizes the grid of k5a_fused)
  void* scanTemp = nullptr; size_t scanTempBytes = 0;  // cub::DeviceScan scratch of the cluster mode
  int* hEv = nullptr;           // pinned: creation records of the copies of a step (ordering them is a host-side sort)
  double* dFtPack = nullptr;    // [6][cap] F / T partials packed for the cross-rank sum
  int* dPerm = nullptr;         // [cap] provisional slot (relative) that goes to final slot r
  long long nClusterSteps = 0;
  cudaStream_t side = nullptr;  // classification of the coming advance beside the stress integral (motion_prepare_async)
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  bool prepared = false; double preparedDt = 0.0; int preparedNInst = 0;
  int lastFlags = 0;            // clusterFail + 1000 violated + 1e6 overflow of the last cluster step (diagnostics)
  long long* dX = nullptr;      // exchange buffer of the owned mode (k5x_pack)
  size_t xWords = 0;
  bool slabsKnown = false;
  long long nOwnedSteps = 0;
};

// ------------------------------------------------------------------------------------------------
// small vector helpers in the reference's evaluation order
// ------------------------------------------------------------------------------------------------
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator-(D3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double magSqr3(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double mag3(D3 a) { return sqrt(magSqr3(a)); }
__device__ __forceinline__ D3 ld3(double* const* p, int i) { return {p[0][i], p[1][i], p[2][i]}; }
__device__ __forceinline__ void st3(double* const* p, int i, D3 v) { p[0][i] = v.x; p[1][i] = v.y; p[2][i] = v.z; }
__device__ __forceinline__ D3 rotv(const double* R, D3 v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
__device__ __forceinline__ D3 classD(const ShapeClassDev& s) { return s.kind == 2 ? D3{s.a, s.b, s.l} : D3{s.d, s.d, s.d}; }
__device__ __forceinline__ double cmptMax3(D3 a) { return fmax(a.x, fmax(a.y, a.z)); }

__device__ __forceinline__ double shape_r_at(const ShapeClassDev& sc, D3 c, double thz, D3 pt) {
  if (sc.kind != 2) return sc.d / 2.0;
  D3 diff = pt - c;
  double th = atan2(diff.y, diff.x) - thz;
  double ca = cos(th) / sc.a, sb = sin(th) / sc.b;
  return sqrt(1.0 / (ca * ca + sb * sb));
}

// particleShape::setRotationMatrix (particleShape.C:186-223) — used for freshly appended copies
__device__ void rot_from_theta(double tx, double ty, double tz, double* R) {
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  auto mul = [&](const double* B) {
    double A[9];
    for (int q = 0; q < 9; ++q) A[q] = R[q];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  };
  if (fabs(tx) > IBM_SMALL) { double c = cos(tx), s = sin(tx); double B[9] = {1.0, 0.0, 0.0, 0.0, c, -s, 0.0, s, c}; mul(B); }
  if (fabs(ty) > IBM_SMALL) { double c = cos(ty), s = sin(ty); double B[9] = {c, 0.0, s, 0.0, 1.0, 0.0, -s, 0.0, c}; mul(B); }
  if (fabs(tz) > IBM_SMALL) { double c = cos(tz), s = sin(tz); double B[9] = {c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0}; mul(B); }
}

// ------------------------------------------------------------------------------------------------
// particleIBM::collision(p1, p2, collisionPt)   particleIBM.C:41-205
// ------------------------------------------------------------------------------------------------
struct PView {
  D3 c; D3 D; double thz; const double* R; const ShapeClassDev* sc;
};

__device__ bool pair_collision(const PView& p1, const PView& p2, D3& collisionPt) {
  const double pi = IBM_PI;
  D3 center1 = p1.c, center2 = p2.c;
  D3 diff = center2 - center1;
  D3 d1 = p1.D, d2 = p2.D;
  if (mag3(diff) > 0.5 * (cmptMax3(d1) + cmptMax3(d2))) return false;
  if ((d1.x + d1.y + d1.z) / 3.0 == cmptMax3(d1) && (d2.x + d2.y + d2.z) / 3.0 == cmptMax3(d2))  // App. B #21
    return 0.5 * (cmptMax3(d1) + cmptMax3(d2)) > mag3(diff);
  D3 pt1 = center1, pt2 = center2;
  const double* R1 = p1.R;
  const double* R2 = p2.R;
  double gamma1 = atan2(diff.y, diff.x) - p1.thz;
  double gamma2 = atan2(diff.y, diff.x) - p2.thz;
  if (center2.x < center1.x) gamma2 -= pi; else gamma1 -= pi;
  double r1 = shape_r_at(*p1.sc, p1.c, p1.thz, pt2), r2 = shape_r_at(*p2.sc, p2.c, p2.thz, pt1);
  double x1 = r1 * cos(gamma1), x2 = r2 * cos(gamma2);
  const bool flip = tan(p1.thz) * tan(p2.thz) < 0;
  if (flip) { x1 *= -1.0; x2 *= -1.0; }
  bool converged = false, hit = false;
  int itt = 0;
  while (!converged) {
    itt++;
    pt1 = rotv(R1, D3{x1, 0, 0}) + center1;
    pt2 = rotv(R2, D3{x2, 0, 0}) + center2;
    D3 oldDiff = diff;
    diff = pt2 - pt1;
    double alpha1 = atan2(diff.y, diff.x) - p1.thz;
    double alpha2 = atan2(diff.y, diff.x) - p2.thz;
    if (center2.x < center1.x) alpha2 -= pi; else alpha1 -= pi;
    D3 rp1 = pt1, rp1Local = rp1;
    {
      double m = tan(alpha1), I = -m * x1;
      double a = m * m + (d1.y / d1.x) * (d1.y / d1.x), b = 2.0 * m * I, c = I * I - d1.y * d1.y;
      double posRoot = (-b + sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
      double negRoot = (-b - sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
      D3 rpPos = rotv(R1, D3{posRoot, I + posRoot * m, 0}) + center1;
      D3 rpNeg = rotv(R1, D3{negRoot, I + negRoot * m, 0}) + center1;
      if (mag3(rpPos - center2) < mag3(rpNeg - center2)) { rp1 = rpPos; rp1Local = D3{posRoot, I + posRoot * m, 0}; }
      else { rp1 = rpNeg; rp1Local = D3{negRoot, I + negRoot * m, 0}; }
      gamma1 = atan2(rp1Local.y, rp1Local.x);
    }
    D3 rp2 = pt2, rp2Local = rp2;
    {
      double m = tan(alpha2), I = -m * x2;
      double a = m * m + (d2.y / d2.x) * (d2.y / d2.x), b = 2.0 * m * I, c = I * I - d2.y * d2.y;
      double posRoot = (-b + sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
      double negRoot = (-b - sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
      D3 rpPos = rotv(R2, D3{posRoot, I + posRoot * m, 0}) + center2;
      D3 rpNeg = rotv(R2, D3{negRoot, I + negRoot * m, 0}) + center2;
      if (mag3(rpPos - center1) < mag3(rpNeg - center1)) { rp2 = rpPos; rp2Local = D3{posRoot, I + posRoot * m, 0}; }
      else { rp2 = rpNeg; rp2Local = D3{negRoot, I + negRoot * m, 0}; }
      gamma2 = atan2(rp2Local.y, rp2Local.x);
    }
    r1 = mag3(rp1Local);
    r2 = mag3(rp2Local);
    x1 = r1 * cos(gamma1);
    x2 = r2 * cos(gamma2);
    if (flip) { x1 *= -1.0; x2 *= -1.0; }
    if (dot3(diff, rp2 - rp1) < 0) {
      converged = true;
      hit = true;
      collisionPt = 0.5 * (rp1 + rp2);
    }
    if (mag3(oldDiff) < mag3(diff) || itt > 10) converged = true;
  }
  return hit;
}

// ------------------------------------------------------------------------------------------------
// cell list: key per instance, sorted by a stable radix sort (ties keep cloud order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bin_of(const MotionDev& m, D3 c, int& bx, int& by, int& bz) {
  bx = min(max((int)floor((c.x - m.bx0) / m.binEdge), 0), m.nbx - 1);
  by = min(max((int)floor((c.y - m.by0) / m.binEdge), 0), m.nby - 1);
  bz = min(max((int)floor((c.z - m.bz0) / m.binEdge), 0), m.nbz - 1);
}

__device__ __forceinline__ void d5_bin(MotionDev m, InstanceSoA in, int i) {
  if (i >= *m.nInst || in.onMesh[i] == 0) return;
  int bx, by, bz;
  const D3 c = ld3(in.c, i);
  bin_of(m, c, bx, by, bz);
  int b = bx + m.nbx * (by + m.nby * bz);
  int slot = atomicAdd(&m.binCnt[b], 1);
  if (slot < m.binCap) {
    const size_t o = (size_t)b * m.binCap + slot;
    m.binItems[o] = i;
    m.binPos[3 * o] = c.x; m.binPos[3 * o + 1] = c.y; m.binPos[3 * o + 2] = c.z;  // centre travels with the bin entry
  } else atomicExch(m.overflow, 3);
}

// particleCloud::computeCollisions, pair part (particleCloud.C:484-523), in three light phases:
//   d5_find_pairs   thread per instance: 27-bin neighbourhood, centre-distance rejection -> (i, j) candidate list
//                   (the reference's early-out, particleIBM.C:56-59, fails for every pair farther apart than the
//                   largest diameter == bin edge, so the list holds little more than the actual contacts);
//   d5_pair_force   thread per candidate: the pair exactly as the reference evaluates it (instance i builds ITS
//                   row: as p1 when its index is the larger one, as p2 otherwise) -> appended to row i;
//   d5_rows_finish  thread per instance with more than one entry: duplicate partner indices (periodic copies) keep
//                   the entry written last in the reference's loop order (largest cloud position of the partner),
//                   then sort by partner index (sum(collisionForces_) order, particleIBM.C:408-411).
__device__ __forceinline__ void d5_find_pairs(MotionDev m, InstanceSoA in, int t) {
  // three threads per instance: one per z-layer of the 27-bin neighbourhood
  const int i = t / 3, dz0 = t % 3 - 1;
  if (i >= *m.nInst || in.onMesh[i] == 0) return;
  const D3 c = ld3(in.c, i);
  int bx, by, bz;
  bin_of(m, c, bx, by, bz);
  const double cut = m.binEdge * m.binEdge * (1.0 + 1e-9);
  for (int dz = dz0; dz <= dz0; ++dz) {
    const int z = bz + dz;
    if (z < 0 || z >= m.nbz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = by + dy;
      if (y < 0 || y >= m.nby) continue;
      for (int x = max(bx - 1, 0); x <= min(bx + 1, m.nbx - 1); ++x) {
        const int bin = x + m.nbx * (y + m.nby * z);
        const int nb = min(m.binCnt[bin], m.binCap);
        for (int q = 0; q < nb; ++q) {
          const size_t o = (size_t)bin * m.binCap + q;
          const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
          if (ddx * ddx + ddy * ddy + ddz * ddz > cut) continue;
          const int j = m.binItems[o];
          if (j == i) continue;
          const int h = atomicAdd(m.pairCount, 1);
          if (h < m.pairCap) { m.pairI[h] = i; m.pairJ[h] = j; }
          else atomicExch(m.overflow, 5);
        }
      }
    }
  }
}

__device__ __forceinline__ void d5_pair_eval(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double deltaT, int i, int j);

__device__ __forceinline__ void d5_pair_force(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double deltaT, int h) {
  if (h >= min(*m.pairCount, m.pairCap)) return;
  d5_pair_eval(m, in, classes, deltaT, m.pairI[h], m.pairJ[h]);
}

// the pair (i, j) as the reference evaluates it, appended to the contact row of instance i
__device__ __forceinline__ void d5_pair_eval(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double deltaT, int i, int j) {
  const int idxI = in.index[i], idxJ = in.index[j];
  if (idxJ == idxI || in.onMesh[j] == 0) return;  // particleCloud.C:502
  const ShapeClassDev* sci = classes + in.cls[i];
  const ShapeClassDev* scj = classes + in.cls[j];
  PView me{ld3(in.c, i), classD(*sci), in.th[2][i], in.rot + 9 * (size_t)i, sci};
  PView other{ld3(in.c, j), classD(*scj), in.th[2][j], in.rot + 9 * (size_t)j, scj};
  const D3 vi = ld3(in.v, i), vj = ld3(in.v, j);
  const double mi = sci->V * in.rho[i], mj = scj->V * in.rho[j];
  const bool iAmP1 = idxI > idxJ;
  const PView& p1 = iAmP1 ? me : other;
  const PView& p2 = iAmP1 ? other : me;
  D3 x12 = p1.c - p2.c;
  D3 v12 = iAmP1 ? (vi - vj) : (vj - vi);
  D3 hitPt{0, 0, 0};
  if (!pair_collision(p1, p2, hitPt)) return;
  D3 myF, myT;
  if (m.model == 1 && sci->kind != 2 && scj->kind != 2) {
    // EXTENSION: linear spring - dashpot in the normal direction, viscous tangential force capped by Coulomb friction
    // (no tangential spring, hence no contact history), acting for as long as the bodies overlap.  Evaluated once in
    // the canonical order (p1 = larger index) so that the two rows of a pair hold exactly opposite forces.
    const D3 w1 = iAmP1 ? ld3(in.om, i) : ld3(in.om, j), w2 = iAmP1 ? ld3(in.om, j) : ld3(in.om, i);
    const double r1 = 0.5 * cmptMax3(p1.D), r2 = 0.5 * cmptMax3(p2.D);
    const double dist = mag3(x12);
    if (!(dist > 0.0)) return;
    const D3 n = x12 / dist;                       // from p2 to p1
    const double overlap = (r1 + r2) - dist;       // > 0 (pair_collision)
    const D3 a1 = -(r1 - 0.5 * overlap) * n;       // contact point relative to the centre of p1
    const D3 a2 = (r2 - 0.5 * overlap) * n;        // ... of p2
    const D3 vrel = (v12 + cross3(w1, a1)) - cross3(w2, a2);
    const double vn = dot3(vrel, n);
    double Fn = m.kn * overlap - m.gn * vn;
    if (Fn < 0.0) Fn = 0.0;                        // no attraction while the bodies separate
    const D3 vt = vrel - vn * n;
    const double vtMag = mag3(vt);
    D3 F1 = Fn * n;
    if (vtMag > 0.0) {
      const double Ft = fmin(m.gt * vtMag, m.mu * Fn);
      F1 = F1 - (Ft / vtMag) * vt;
    }
    myF = iAmP1 ? F1 : -F1;
    myT = iAmP1 ? cross3(a1, F1) : cross3(a2, -F1);
  } else {
    if (!(dot3(x12, v12) < 0)) return;
    double mass1 = iAmP1 ? mi : mj, mass2 = iAmP1 ? mj : mi;
    double M = mass1 * mass2 / (mass1 + mass2);
    double x12Dotv12 = dot3(x12, v12);
    double x12Sqr = magSqr3(x12);
    D3 F = 2.0 * M * x12Dotv12 / x12Sqr * x12 * m.e / deltaT;  // particleCloud.C:513
    myF = iAmP1 ? -F : F;
    myT = iAmP1 ? cross3(-F, hitPt) : cross3(F, hitPt);
  }
  const int slot = atomicAdd(&m.cCnt[i], 1);
  if (slot >= CONTACT_CAP) { atomicExch(m.overflow, 1); return; }
  const size_t o = (size_t)i * CONTACT_CAP + slot;
  m.cIdx[o] = idxJ;
  m.cPos[o] = j;
  m.cF[3 * o] = myF.x; m.cF[3 * o + 1] = myF.y; m.cF[3 * o + 2] = myF.z;
  m.cT[3 * o] = myT.x; m.cT[3 * o + 1] = myT.y; m.cT[3 * o + 2] = myT.z;
}

__device__ __forceinline__ void d5_rows_finish(MotionDev m, InstanceSoA in, int i) {
  if (i >= *m.nInst) return;
  int cnt = min(m.cCnt[i], CONTACT_CAP);
  if (cnt < 2) { if (m.cCnt[i] > CONTACT_CAP) m.cCnt[i] = CONTACT_CAP; return; }
  const size_t r0 = (size_t)i * CONTACT_CAP;
  // duplicates of one partner index: keep the entry of the partner latest in cloud order
  for (int a = 0; a < cnt; ++a)
    for (int b2 = a + 1; b2 < cnt;) {
      if (m.cIdx[r0 + b2] == m.cIdx[r0 + a]) {
        const int keep = m.cPos[r0 + b2] > m.cPos[r0 + a] ? b2 : a;
        if (keep == b2) {
          m.cPos[r0 + a] = m.cPos[r0 + b2];
          for (int q = 0; q < 3; ++q) { m.cF[3 * (r0 + a) + q] = m.cF[3 * (r0 + b2) + q]; m.cT[3 * (r0 + a) + q] = m.cT[3 * (r0 + b2) + q]; }
        }
        --cnt;  // move the last entry into b2
        m.cIdx[r0 + b2] = m.cIdx[r0 + cnt]; m.cPos[r0 + b2] = m.cPos[r0 + cnt];
        for (int q = 0; q < 3; ++q) { m.cF[3 * (r0 + b2) + q] = m.cF[3 * (r0 + cnt) + q]; m.cT[3 * (r0 + b2) + q] = m.cT[3 * (r0 + cnt) + q]; }
      } else ++b2;
    }
  // insertion sort by partner index
  for (int a = 1; a < cnt; ++a) {
    const int ki = m.cIdx[r0 + a], kp = m.cPos[r0 + a];
    double f[3], t[3];
    for (int q = 0; q < 3; ++q) { f[q] = m.cF[3 * (r0 + a) + q]; t[q] = m.cT[3 * (r0 + a) + q]; }
    int b2 = a - 1;
    while (b2 >= 0 && m.cIdx[r0 + b2] > ki) {
      m.cIdx[r0 + b2 + 1] = m.cIdx[r0 + b2]; m.cPos[r0 + b2 + 1] = m.cPos[r0 + b2];
      for (int q = 0; q < 3; ++q) { m.cF[3 * (r0 + b2 + 1) + q] = m.cF[3 * (r0 + b2) + q]; m.cT[3 * (r0 + b2 + 1) + q] = m.cT[3 * (r0 + b2) + q]; }
      --b2;
    }
    m.cIdx[r0 + b2 + 1] = ki; m.cPos[r0 + b2 + 1] = kp;
    for (int q = 0; q < 3; ++q) { m.cF[3 * (r0 + b2 + 1) + q] = f[q]; m.cT[3 * (r0 + b2 + 1) + q] = t[q]; }
  }
  m.cCnt[i] = cnt;
}

// find_pairs + pair_force + rows_finish of ONE instance by one thread (used where the instances are few and a phase
// boundary is a grid-wide barrier, k5a_fused): same candidates (27 bins, centre distance <= bin edge), same pair
// evaluation, same finished row - the order in which entries are appended differs, the finished row does not depend on it.
__device__ __forceinline__ void d5_row_direct(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double deltaT, int i) {
  if (i >= *m.nInst || in.onMesh[i] == 0) return;
  const D3 c = ld3(in.c, i);
  int bx, by, bz;
  bin_of(m, c, bx, by, bz);
  const double cut = m.binEdge * m.binEdge * (1.0 + 1e-9);
  for (int z = max(bz - 1, 0); z <= min(bz + 1, m.nbz - 1); ++z)
    for (int y = max(by - 1, 0); y <= min(by + 1, m.nby - 1); ++y)
      for (int x = max(bx - 1, 0); x <= min(bx + 1, m.nbx - 1); ++x) {
        const int bin = x + m.nbx * (y + m.nby * z);
        const int nb = min(m.binCnt[bin], m.binCap);
        for (int q = 0; q < nb; ++q) {
          const size_t o = (size_t)bin * m.binCap + q;
          const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
          if (ddx * ddx + ddy * ddy + ddz * ddz > cut) continue;
          const int j = m.binItems[o];
          if (j != i) d5_pair_eval(m, in, classes, deltaT, i, j);
        }
      }
  d5_rows_finish(m, in, i);
}

// particleIBM::wallHit for every wall slot (particleIBM.C:453-488, particleCloud.C:591-597)
__device__ __forceinline__ void face_range(const MotionDev& m, const LatticeDev& L, const PatchDev& p, D3 c, double r, int& a0, int& a1,
                                           int& b0, int& b1) {
  a0 = 0; a1 = p.nA - 1; b0 = 0; b1 = p.nB - 1;
  if (!p.latticeOrdered) { a0 = 0; a1 = p.nFaces - 1; b0 = b1 = 0; return; }
  const int axis = p.side / 2;
  const double o3[3] = {L.ox, L.oy, L.oz}, d3[3] = {L.dx, L.dy, L.dz}, c3[3] = {c.x, c.y, c.z};
  const int ta = (axis == 0) ? 1 : 0, tb = (axis == 2) ? 1 : 2;
  a0 = max(0, (int)floor((c3[ta] - r - o3[ta]) / d3[ta]) - 1); a1 = min(p.nA - 1, (int)floor((c3[ta] + r - o3[ta]) / d3[ta]) + 1);
  b0 = max(0, (int)floor((c3[tb] - r - o3[tb]) / d3[tb]) - 1); b1 = min(p.nB - 1, (int)floor((c3[tb] + r - o3[tb]) / d3[tb]) + 1);
}

__device__ __forceinline__ bool patch_in_reach(const PatchDev& p, D3 c, double r) {
  double dx = fmax(fmax(p.lo[0] - c.x, c.x - p.hi[0]), 0.0), dy = fmax(fmax(p.lo[1] - c.y, c.y - p.hi[1]), 0.0),
         dz = fmax(fmax(p.lo[2] - c.z, c.z - p.hi[2]), 0.0);
  return dx * dx + dy * dy + dz * dz <= r * r * (1.0 + 1e-9) + 1e-300;
}

__device__ __forceinline__ void d5_walls(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt, int i) {
  if (i >= *m.nInst) return;
  if (in.onMesh[i] == 0) return;
  const ShapeClassDev* sc = classes + in.cls[i];
  const D3 c = ld3(in.c, i), v = ld3(in.v, i), om = ld3(in.om, i);
  const double thz = in.th[2][i];
  const double mass = sc->V * in.rho[i];
  for (int w = 0; w < m.nWalls; ++w) {
    const PatchDev& p = m.patches[m.wallPatch[w]];
    if (!patch_in_reach(p, c, sc->Rbound)) continue;
    int a0, a1, b0, b1;
    face_range(m, L, p, c, sc->Rbound, a0, a1, b0, b1);
    bool done = false;
    for (int b = b0; b <= b1 && !done; ++b)
      for (int a = a0; a <= a1 && !done; ++a) {
        int facei = p.latticeOrdered ? a + p.nA * b : a;
        D3 fc{p.Cf[3 * facei], p.Cf[3 * facei + 1], p.Cf[3 * facei + 2]};
        D3 R = c - fc;
        if (magSqr3(R) > sc->Rbound * sc->Rbound * (1.0 + 1e-9)) continue;  // cannot pass mag(R) <= r(faceCentre)
        D3 sf{p.Sf[3 * facei], p.Sf[3 * facei + 1], p.Sf[3 * facei + 2]};
        D3 norm = sf / p.magSf[facei];
        D3 rv = cross3(-(fc - c), om) + v;  // v(faceCentre)
        if (mag3(R) <= shape_r_at(*sc, c, thz, fc) && dot3(rv, norm) > 0) {
          D3 vNorm = norm * dot3(norm, rv);
          D3 rHat = R / mag3(R);
          D3 F = -2.0 * mass * vNorm / dt;
          D3 Fw = dot3(F, rHat) * rHat;
          D3 Tw = cross3(F - Fw, R);
          size_t o = 3 * ((size_t)i * m.nWalls + w);
          m.Fw[o] = Fw.x; m.Fw[o + 1] = Fw.y; m.Fw[o + 2] = Fw.z;
          m.Tw[o] = Tw.x; m.Tw[o + 1] = Tw.y; m.Tw[o + 2] = Tw.z;
          done = true;  // only one hit per patch
        }
      }
  }
}

// Copies of one index share the larger-magnitude entry (particleCloud.C:525-554, 599-620).  The reference's
// in-place pair sweep leaves every member of a same-index group with the entry of largest magnitude, ties going
// to the member latest in cloud order; evaluated here per instance from the pre-merge tables (double buffered).
__device__ __forceinline__ void d5_merge(MotionDev m, InstanceSoA in, int i) {
  if (i >= *m.nInst) return;
  const int idx = in.index[i];
  const int head = m.idxHead[idx];
  if (head < 0 || m.instNext[head] < 0) return;  // single instance of this index: k5_solve reads the raw tables
  int cnt = 0;
  int eIdx[CONTACT_CAP];
  D3 eF[CONTACT_CAP], eT[CONTACT_CAP];
  double eMag[CONTACT_CAP];
  for (int mem = head; mem >= 0; mem = m.instNext[mem]) {
    int c = m.cCnt[mem];
    for (int a = 0; a < c; ++a) {
      size_t o = (size_t)mem * CONTACT_CAP + a;
      int k = m.cIdx[o];
      D3 f{m.cF[3 * o], m.cF[3 * o + 1], m.cF[3 * o + 2]};
      double mg = mag3(f);
      int slot = -1;
      for (int t = 0; t < cnt; ++t) if (eIdx[t] == k) slot = t;
      if (slot < 0) {
        if (cnt >= CONTACT_CAP) { atomicExch(m.overflow, 1); continue; }
        slot = cnt++;
        eIdx[slot] = k; eMag[slot] = -1.0;
      }
      if (mg >= eMag[slot]) {  // members are visited in cloud order: ">=" lets the later member win ties
        eMag[slot] = mg; eF[slot] = f;
        eT[slot] = D3{m.cT[3 * o], m.cT[3 * o + 1], m.cT[3 * o + 2]};
      }
    }
  }
  for (int a = 1; a < cnt; ++a) {
    int ki = eIdx[a]; D3 f = eF[a], t = eT[a];
    int b = a - 1;
    while (b >= 0 && eIdx[b] > ki) { eIdx[b + 1] = eIdx[b]; eF[b + 1] = eF[b]; eT[b + 1] = eT[b]; --b; }
    eIdx[b + 1] = ki; eF[b + 1] = f; eT[b + 1] = t;
  }
  m.cCnt2[i] = cnt;
  for (int a = 0; a < cnt; ++a) {
    size_t o = (size_t)i * CONTACT_CAP + a;
    m.cIdx2[o] = eIdx[a];
    m.cF2[3 * o] = eF[a].x; m.cF2[3 * o + 1] = eF[a].y; m.cF2[3 * o + 2] = eF[a].z;
    m.cT2[3 * o] = eT[a].x; m.cT2[3 * o + 1] = eT[a].y; m.cT2[3 * o + 2] = eT[a].z;
  }
  for (int w = 0; w < m.nWalls; ++w) {
    double best = -1.0;
    D3 bf{0, 0, 0}, bt{0, 0, 0};
    for (int mem = head; mem >= 0; mem = m.instNext[mem]) {
      size_t o = 3 * ((size_t)mem * m.nWalls + w);
      D3 f{m.Fw[o], m.Fw[o + 1], m.Fw[o + 2]};
      double mg = mag3(f);
      if (mg >= best) { best = mg; bf = f; bt = D3{m.Tw[o], m.Tw[o + 1], m.Tw[o + 2]}; }
    }
    size_t o = 3 * ((size_t)i * m.nWalls + w);
    m.Fw2[o] = bf.x; m.Fw2[o + 1] = bf.y; m.Fw2[o + 2] = bf.z;
    m.Tw2[o] = bt.x; m.Tw2[o + 1] = bt.y; m.Tw2[o + 2] = bt.z;
  }
}

// particleIBM::patchHit + found() for every (cyclic patch, instance)   particleIBM.C:490-517, particleCloud.C:447-479, 685-730
// stage 1: (patch, instance) pairs whose centre is within reach of the patch
__device__ __forceinline__ void d5_cyclic_candidates(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, int i) {
  if (i >= *m.nInst || in.onMesh[i] == 0) return;
  const D3 c = ld3(in.c, i);
  const double r = classes[in.cls[i]].Rbound;
  for (int slot = 0; slot < m.nCyc; ++slot) {
    const PatchDev& p = m.patches[m.cycPatch[slot]];
    if (p.type != IBM_CYCLIC) continue;  // processor patches are handled by the slab exchange, not here
    if (!patch_in_reach(p, c, r)) continue;
    if (p.planar) {  // patchHit needs (v & norm) > 0 for the hit face (particleIBM.C:511): same test, once per patch
      const D3 v = ld3(in.v, i);
      if (!(dot3(v, D3{p.nrm[0], p.nrm[1], p.nrm[2]}) > 0)) continue;
    }
    if (p.uniformShift) {
      // found() (particleCloud.C:447-479, App. B #15) rejects the new copy when an instance of the same index lies
      // within 1e-4 of the image centre R + Cf_neighbour[face].  For a pure translation that centre is c + shift up to
      // ~1e-12 whichever face is hit, so an instance within HALF that distance settles it here, without the face scan
      // (in steady state nearly every candidate already has its copy); anything closer to the threshold goes on to the
      // exact test of stage 2.
      const D3 nc = c + D3{p.shift[0], p.shift[1], p.shift[2]};
      bool have = false;
      for (int mem = m.idxHead[in.index[i]]; mem >= 0 && !have; mem = m.instNext[mem])
        have = magSqr3(nc - ld3(in.c, mem)) < 0.25e-8;
      if (have) continue;
    }
    int h = atomicAdd(m.candCount, 1);
    if (h < m.candCap) m.candKey[h] = slot * m.cap + i;
    else atomicExch(m.overflow, 4);
  }
}

// stage 2: one warp per candidate; lanes scan the candidate faces in patch order, the first qualifying face
// (lowest position in that order) is the reference's patchHit() result.
__device__ __forceinline__ void d5_cyclic_detect(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes,
                                                 int tid, int nth) {
  const int lane = threadIdx.x & 31;
  const int nCand = min(*m.candCount, m.candCap);
  for (int w = tid >> 5; w < nCand; w += nth >> 5) {
    const int key = m.candKey[w];
    const int slot = key / m.cap, i = key % m.cap;
    const PatchDev& p = m.patches[m.cycPatch[slot]];
    const ShapeClassDev* sc = classes + in.cls[i];
    const D3 c = ld3(in.c, i), v = ld3(in.v, i);
    const double thz = in.th[2][i];
    int a0, a1, b0, b1;
    face_range(m, L, p, c, sc->Rbound, a0, a1, b0, b1);
    const int na = a1 - a0 + 1, nf = na * (b1 - b0 + 1);
    int firstPos = 0x7fffffff;
    for (int base = 0; base < nf && firstPos == 0x7fffffff; base += 32) {
      int pos = base + lane;
      bool ok = false;
      if (pos < nf) {
        int a = a0 + pos % na, b = b0 + pos / na;
        int facei = p.latticeOrdered ? a + p.nA * b : a;
        D3 fc{p.Cf[3 * facei], p.Cf[3 * facei + 1], p.Cf[3 * facei + 2]};
        D3 R = c - fc;
        if (magSqr3(R) <= sc->Rbound * sc->Rbound * (1.0 + 1e-9)) {
          D3 sf{p.Sf[3 * facei], p.Sf[3 * facei + 1], p.Sf[3 * facei + 2]};
          D3 norm = sf / p.magSf[facei];
          ok = mag3(R) <= shape_r_at(*sc, c, thz, fc) && dot3(v, norm) > 0;
        }
      }
      unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (bal) firstPos = base + __ffs(bal) - 1;
    }
    if (firstPos != 0x7fffffff && lane == 0) {
      int a = a0 + firstPos % na, b = b0 + firstPos / na;
      int facei = p.latticeOrdered ? a + p.nA * b : a;
      D3 fc{p.Cf[3 * facei], p.Cf[3 * facei + 1], p.Cf[3 * facei + 2]};
      D3 R = c - fc;
      const PatchDev& q = m.patches[p.neighb];
      D3 nc = R + D3{q.Cf[3 * facei], q.Cf[3 * facei + 1], q.Cf[3 * facei + 2]};
      bool f = false;
      for (int mem = m.idxHead[in.index[i]]; mem >= 0; mem = m.instNext[mem])
        if (mag3(nc - ld3(in.c, mem)) < 1e-4) f = true;  // App. B #15
      if (!f) {
        int h = atomicAdd(m.hitCount, 1);
        if (h < m.cap) {
          m.hitKey[h] = key;
          m.hitCenter[3 * (size_t)h] = nc.x; m.hitCenter[3 * (size_t)h + 1] = nc.y; m.hitCenter[3 * (size_t)h + 2] = nc.z;
        }
      }
    }
  }
}

// append the new periodic copies in the reference's order (patch-major, cloud order within a patch):
// particleIBM copy constructor (particleIBM.C:318-350) from origParticles_[index]   particleCloud.C:815-832.
// The hit list is unordered (atomics); each hit finds its rank by counting smaller keys.
__device__ __forceinline__ void d5_append(MotionDev m, InstanceSoA in, int tid, int nth) {
  const int total = *m.hitCount;
  if (total == 0) return;
  const int n0 = *m.nInst;
  // the rank of a hit is found by counting smaller keys (O(total^2)): bounded so that a diverged run fails fast
  if (total > m.cap - n0 || total > 131072) { if (tid == 0) atomicExch(m.overflow, 2); return; }
  for (int h = tid; h < total; h += nth) {
    const int key = m.hitKey[h];
    int rank = 0;
    for (int q = 0; q < total; ++q) rank += m.hitKey[q] < key;
    int src = key % m.cap;
    int dst = n0 + rank;
    int idx = in.index[src];
    in.index[dst] = idx;
    in.copy[dst] = -1;  // assigned sequentially below
    in.active[dst] = 0;
    in.onMesh[dst] = IBM_PARTIAL;
    in.cls[dst] = m.origCls[idx];
    for (int q = 0; q < 3; ++q) {
      in.c[q][dst] = m.hitCenter[3 * (size_t)h + q];
      in.th[q][dst] = in.th[q][src];
      in.v[q][dst] = in.v[q][src];
      in.om[q][dst] = in.om[q][src];
      in.F[q][dst] = 0.0;  // integratedForce_ of origParticles_[index] (never updated after construction)
      in.T[q][dst] = 0.0;
    }
    in.rho[dst] = m.origRho[idx];
    in.age[dst] = in.age[src];
    in.nShell[dst] = 0; in.shellOff[dst] = 0; in.nPtsOnMesh[dst] = 0;
    rot_from_theta(in.th[0][src], in.th[1][src], in.th[2][src], in.rot + 9 * (size_t)dst);
    if (m.activeMode) m.bud[dst] = m.bud[src];  // quiet / active split: the copy inherits its source's displacement budget
    m.cCnt[dst] = 0;
    m.cCnt2[dst] = 0;
    for (int w = 0; w < m.nWalls; ++w)
      for (int q = 0; q < 3; ++q) {
        size_t o = 3 * ((size_t)dst * m.nWalls + w) + q;
        m.Fw[o] = 0.0; m.Tw[o] = 0.0; m.Fw2[o] = 0.0; m.Tw2[o] = 0.0;
      }
  }
}

// sequential part of the append: copy ids (maxCopies_[index]+1, ++) and the same-index lists, in append order
__device__ __forceinline__ void d5_append_link(MotionDev m, InstanceSoA in) {
  *m.candCount = 0;
  const int total = *m.hitCount;
  if (total == 0) return;
  *m.hitCount = 0;
  if (total > m.cap - *m.nInst || total > 131072) return;
  const int n0 = *m.nInst;
  for (int t = 0; t < total; ++t) {
    int dst = n0 + t;
    int idx = in.index[dst];
    in.copy[dst] = m.maxCopies[idx] + 1;
    m.maxCopies[idx] += 1;
    m.instNext[dst] = -1;
    int tail = m.idxTail[idx];
    if (tail >= 0) m.instNext[tail] = dst; else m.idxHead[idx] = dst;
    m.idxTail[idx] = dst;
    if (m.activeMode) {  // a copy belongs to a group of more than one instance: active from its creation on
      m.act[*m.nAct] = dst;
      *m.nAct += 1;
      m.quiet[dst] = 0;
      for (int q = 0; q < 3; ++q) m.c0[(size_t)q * m.cap + dst] = in.c[q][dst];
    }
  }
  *m.nInst = n0 + total;
}

// particleIBM::solve(dt, g, moving, rotation)   particleIBM.C:393-450
__device__ __forceinline__ void d5_solve(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt,
                         const int* cCnt, const int* cIdx, const double* cF, const double* cT, double* Fw, double* Tw, int* cCntA, int* cCntB, int i) {
  if (i >= *m.nInst) return;
  if (in.onMesh[i] == 0) return;
  const ShapeClassDev* sc = classes + in.cls[i];
  const double mass = sc->V * in.rho[i];
  {
    // instances whose index has copies read the merged tables (k5_merge), all others the raw ones
    const int head = m.idxHead[in.index[i]];
    const bool multi = head >= 0 && m.instNext[head] >= 0;
    if (!multi) { cCnt = m.cCnt; cF = m.cF; cT = m.cT; Fw = m.Fw; Tw = m.Tw; }
  }
  const int cnt = cCnt[i];
  if (m.moving) {
    D3 sFc{0, 0, 0}, sFw{0, 0, 0};
    for (int a = 0; a < cnt; ++a) { size_t o = 3 * ((size_t)i * CONTACT_CAP + a); sFc = sFc + D3{cF[o], cF[o + 1], cF[o + 2]}; }
    for (int w = 0; w < m.nWalls; ++w) { size_t o = 3 * ((size_t)i * m.nWalls + w); sFw = sFw + D3{Fw[o], Fw[o + 1], Fw[o + 2]}; }
    D3 F = ld3(in.F, i) + sFc + sFw;
    D3 v = ld3(in.v, i);
    v = v + dt * (F / mass + D3{m.gx, m.gy, m.gz});
    st3(in.v, i, v);
    D3 dx = dt * v;
    D3 c = ld3(in.c, i) + dx;
    st3(in.c, i, c);
    int I, J, K;
    in.active[i] = find_cell(L, c.x, c.y, c.z, I, J, K) >= 0;  // particleIBM.C:417-424
  }
  if (m.rotating) {
    D3 sTc{0, 0, 0}, sTw{0, 0, 0};
    for (int a = 0; a < cnt; ++a) { size_t o = 3 * ((size_t)i * CONTACT_CAP + a); sTc = sTc + D3{cT[o], cT[o + 1], cT[o + 2]}; }
    for (int w = 0; w < m.nWalls; ++w) { size_t o = 3 * ((size_t)i * m.nWalls + w); sTw = sTw + D3{Tw[o], Tw[o + 1], Tw[o + 2]}; }
    D3 T = ld3(in.T, i) + sTc + sTw;
    D3 om = ld3(in.om, i);
    om = om + dt * T / (mass * sc->I);
    st3(in.om, i, om);
    D3 th = ld3(in.th, i) + om * dt;
    if (th.z < 0) th.z += IBM_TWOPI;
    if (th.z >= IBM_TWOPI) th.z -= IBM_TWOPI;
    st3(in.th, i, th);
  }
  // collisionForces_ = Zero; ... (particleIBM.C:446-449)
  cCntA[i] = 0;
  cCntB[i] = 0;
  for (int w = 0; w < m.nWalls; ++w)
    for (int q = 0; q < 3; ++q) {
      size_t o = 3 * ((size_t)i * m.nWalls + w) + q;
      m.Fw[o] = 0.0; m.Tw[o] = 0.0; m.Fw2[o] = 0.0; m.Tw2[o] = 0.0;
    }
}

// particleCloud::setPositions, serial branch (particleCloud.C:411-443): processed in cloud order, every inactive
// instance pushes its theta/v/omega to the other copies of its index.
__device__ __forceinline__ void d5_set_positions(MotionDev m, InstanceSoA in, int idx) {
  if (idx >= m.nParticles) return;
  int head = m.idxHead[idx];
  if (head < 0 || m.instNext[head] < 0) return;  // single instance: nothing to copy
  for (int p = head; p >= 0; p = m.instNext[p]) {
    if (in.active[p]) continue;
    D3 th = ld3(in.th, p), v = ld3(in.v, p), om = ld3(in.om, p);
    int cp = in.copy[p];
    for (int p2 = head; p2 >= 0; p2 = m.instNext[p2])
      if (in.copy[p2] != cp) { st3(in.th, p2, th); st3(in.v, p2, v); st3(in.om, p2, om); }
  }
}

__global__ void k5_bin(MotionDev m, InstanceSoA in) {
  d5_bin(m, in, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_bin_cand(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, int doCyclic) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  d5_bin(m, in, i);
  if (doCyclic) d5_cyclic_candidates(m, in, classes, i);
}
__global__ void k5_find_pairs(MotionDev m, InstanceSoA in) { d5_find_pairs(m, in, blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void k5_pair_force(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double deltaT) {
  // grid-stride: the grid is sized from a typical candidate count (and frozen in a captured graph), the list may be longer
  const int np = min(*m.pairCount, m.pairCap);
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < np; h += gridDim.x * blockDim.x) d5_pair_force(m, in, classes, deltaT, h);
}
__global__ void k5_rows_finish(MotionDev m, InstanceSoA in) {
  d5_rows_finish(m, in, blockIdx.x * blockDim.x + threadIdx.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) *m.pairCount = 0;  // pair_force is done: ready for the next sub-step
}
__global__ void k5_solve_clear(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt,
                               const int* cCnt, const int* cIdx, const double* cF, const double* cT, double* Fw, double* Tw, int* cCntA,
                               int* cCntB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  d5_solve(m, L, in, classes, dt, cCnt, cIdx, cF, cT, Fw, Tw, cCntA, cCntB, i);
  const int nBins = m.nbx * m.nby * m.nbz;
  for (int b = i; b < nBins; b += gridDim.x * blockDim.x) m.binCnt[b] = 0;  // contacts are done: bins ready for the next sub-step
}
// append + its sequential tail in one launch: the block that finishes last links the new instances
__global__ void k5_append_all(MotionDev m, InstanceSoA in) {
  __shared__ int last;
  d5_append(m, in, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(m.blocksDone, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *m.blocksDone = 0;
    __threadfence();
    d5_append_link(m, in);
  }
}

__global__ void k5_walls(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt) {
  d5_walls(m, L, in, classes, dt, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_merge(MotionDev m, InstanceSoA in) {
  d5_merge(m, in, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_cyclic_candidates(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes) {
  d5_cyclic_candidates(m, in, classes, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_solve(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt, const int* cCnt, const int* cIdx, const double* cF, const double* cT, double* Fw, double* Tw, int* cCntA, int* cCntB) {
  d5_solve(m, L, in, classes, dt, cCnt, cIdx, cF, cT, Fw, Tw, cCntA, cCntB, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_set_positions(MotionDev m, InstanceSoA in) {
  d5_set_positions(m, in, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k5_cyclic_detect(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes) {
  d5_cyclic_detect(m, L, in, classes, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void k5_append(MotionDev m, InstanceSoA in) { d5_append(m, in, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
__global__ void k5_append_link(MotionDev m, InstanceSoA in) {
  if (threadIdx.x == 0 && blockIdx.x == 0) d5_append_link(m, in);
}

// ------------------------------------------------------------------------------------------------
// Quiet / active split of a fluid step (option active_loop5; same results, less work per sub-step).
//
// The hydrodynamic force and torque are frozen over the ~100 sub-steps of particleCloud::solve (particleCloud.C:1247-1275),
// so an instance that cannot touch anything during the step - no other body, no wall, no periodic patch - and that is the
// only instance of its index just integrates  v += dt (F/m + g); x += dt v; ...  (particleIBM.C:393-450) on its own: all
// its sub-steps are done by ONE thread in ONE kernel (k5q_quiet), with the very arithmetic of the per-sub-step path.  Only
// the remaining ACTIVE instances (a few per cent of a dilute-to-moderate suspension) go through the per-sub-step chain of
// bins, pair search, wall / patch tests, copy handling and merging, on a compact list.
//
// "Cannot touch anything" is decided at the start of the step from displacement budgets.  With dt the fluid step,
//     s_i = 1.02 (|v_i| dt + |F_i/m_i + g| dt^2 / 2)       what instance i covers if it feels no impulse (the 2 % pays for the
//                                                         symplectic-Euler sum  sum_k dt_k |v_0 + a t_{k+1}|  of ~100 sub-steps)
//     S   = safety * max_i s_i
//     B_i = safety * max{ s_j : j within two hops of i in the graph of pairs closer than (Dmax_i + Dmax_j)/2 + 2 S }
// (an elastic impulse hands a body at most its partners' momentum, so the neighbourhood maximum is what an instance can
// plausibly reach; `safety` > sqrt(2) covers one exchange among equal bodies outright; the neighbourhood of i includes the
// places where a copy of i may appear, and the instances of one index share the largest B of their group).  Instance i is ACTIVE when
//   * its index has several instances (periodic copies share and exchange state, particleCloud.C:316-444, 525-554),
//   * a wall or periodic patch is within its reach + B_i,
//   * another body j is within (Dmax_i + Dmax_j)/2 + B_i + B_j (particleIBM.C:56-59 is the reference's own early-out at
//     (Dmax_i + Dmax_j)/2), or
//   * the IMAGE of a body g that may spawn a periodic copy this step (g within its reach + S of a cyclic patch, whichever
//     way it moves now: a contact can turn it round) is within (Dmax_i + Dmax_g)/2 + B_i + (1 + n) B_g, n = number of
//     translations composed in the image (a copy is created within B_g of the image of g's start and then has B_g of its own).
// A quiet instance covers at most s_i <= B_i by construction; an active one is CHECKED after every sub-step: farther than
// B_i from where it started (or was created) raises `violated` and the host redoes the whole step from a snapshot without
// the split.  Under that check a quiet instance can never come within contact distance of anything, so the split changes
// nothing but the work done.  Cyclic patches that are not pure translations switch the split off (host side).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmax_of(const ShapeClassDev& s) { return s.kind == 2 ? fmax(s.a, fmax(s.b, s.l)) : s.d; }

__global__ void k5q_bounds(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dtStep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double sb = 0.0;
  if (i < *m.nInst) {
    if (in.onMesh[i] != 0) {
      const ShapeClassDev* sc = classes + in.cls[i];
      const double mass = sc->V * in.rho[i];
      sb = 1.02 * (mag3(ld3(in.v, i)) * dtStep + 0.5 * mag3(ld3(in.F, i) / mass + D3{m.gx, m.gy, m.gz}) * dtStep * dtStep);
    }
    m.sOwn[i] = sb;
    for (int q = 0; q < 3; ++q) m.c0[(size_t)q * m.cap + i] = in.c[q][i];
  }
  for (int o = 16; o > 0; o >>= 1) sb = fmax(sb, __shfl_xor_sync(0xffffffffu, sb, o));
  if ((threadIdx.x & 31) == 0) atomicMax(m.vaBits, (unsigned long long)__double_as_longlong(sb));  // non-negative doubles order like their bit patterns
}

__device__ __forceinline__ double split_S(const MotionDev& m) {
  return m.safety * __longlong_as_double((long long)m.vaBits[0]) + 1e-12 * m.binEdge;
}

// bin entries < 0 are images: -(1 + 4 g + (n - 1)) = the image of instance g under n composed translations
__device__ __forceinline__ bool bin_entry(int item, int& inst, int& mult) {
  if (item >= 0) { inst = item; mult = 1; return false; }
  const int e = -item - 1;
  inst = e >> 2;
  mult = (e & 3) + 2;
  return true;
}

// translations of the cyclic patches within `reach` of c (at most three: a corner); false = both sides of a direction
__device__ __forceinline__ bool image_shifts(const MotionDev& m, D3 c, double reach, D3* sh, int& n) {
  n = 0;
  for (int s = 0; s < m.nCyc; ++s) {
    const PatchDev& p = m.patches[m.cycPatch[s]];
    if (!patch_in_reach(p, c, reach)) continue;
    if (n == 3) return false;
    sh[n++] = D3{p.shift[0], p.shift[1], p.shift[2]};
  }
  return true;
}

// images of the instances that may spawn a periodic copy this step, entered into the bins next to the real centres
__global__ void k5q_ghosts(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *m.nInst || in.onMesh[i] == 0) return;
  const D3 c = ld3(in.c, i);
  const double reach = classes[in.cls[i]].Rbound + split_S(m);
  D3 sh[3];
  int n;
  if (!image_shifts(m, c, reach, sh, n)) { atomicExch(m.noSplit, 1); return; }  // both sides of one direction in reach: a domain a few bodies wide
  for (int sub = 1; sub < (1 << n); ++sub) {
    D3 g = c;
    int cnt = 0;
    for (int q = 0; q < n; ++q)
      if (sub & (1 << q)) { g = g + sh[q]; ++cnt; }
    int bx, by, bz;
    bin_of(m, g, bx, by, bz);
    const int b = bx + m.nbx * (by + m.nby * bz);
    const int slot = atomicAdd(&m.binCnt[b], 1);
    if (slot < m.binCap) {  // (a full bin keeps its count above the capacity: the classification then decides nothing from it)
      const size_t o = (size_t)b * m.binCap + slot;
      m.binItems[o] = -(1 + 4 * i + (cnt - 1));
      m.binPos[3 * o] = g.x; m.binPos[3 * o + 1] = g.y; m.binPos[3 * o + 2] = g.z;
    }
  }
}

// bins that can hold an entry within `cut` of c, never more than the +-2 neighbourhood the callers are entitled to
// (2 S <= binEdge is checked by them): most instances need +-1, 27 bins instead of 125
struct BinRange { int x0, x1, y0, y1, z0, z1; };
__device__ __forceinline__ BinRange bin_range(const MotionDev& m, D3 c, double cut) {
  int bx, by, bz;
  bin_of(m, c, bx, by, bz);
  const double inv = 1.0 / m.binEdge, pad = cut * (1.0 + 1e-9) + 1e-9 * m.binEdge;
  BinRange r;
  r.x0 = max(max(bx - 2, 0), (int)floor((c.x - pad - m.bx0) * inv)); r.x1 = min(min(bx + 2, m.nbx - 1), (int)floor((c.x + pad - m.bx0) * inv));
  r.y0 = max(max(by - 2, 0), (int)floor((c.y - pad - m.by0) * inv)); r.y1 = min(min(by + 2, m.nby - 1), (int)floor((c.y + pad - m.by0) * inv));
  r.z0 = max(max(bz - 2, 0), (int)floor((c.z - pad - m.bz0) * inv)); r.z1 = min(min(bz + 2, m.nbz - 1), (int)floor((c.z + pad - m.bz0) * inv));
  if (m.nbz == 1) { r.z0 = 0; r.z1 = 0; }
  // an instance clamped into an edge bin (centre outside the binned region) keeps the full neighbourhood
  if (r.x0 > bx || r.x1 < bx) { r.x0 = max(bx - 2, 0); r.x1 = min(bx + 2, m.nbx - 1); }
  if (r.y0 > by || r.y1 < by) { r.y0 = max(by - 2, 0); r.y1 = min(by + 2, m.nby - 1); }
  if (r.z0 > bz || r.z1 < bz) { r.z0 = max(bz - 2, 0); r.z1 = min(bz + 2, m.nbz - 1); }
  return r;
}

// maximum of src over the bodies and images within Rij + 2 S of position c (instance i itself and its own images excluded)
__device__ __forceinline__ double nbr_scan_max(const MotionDev& m, const InstanceSoA& in, const ShapeClassDev* __restrict__ classes,
                                               const double* __restrict__ src, D3 c, double di, int i, double S, double u) {
  const BinRange br = bin_range(m, c, 0.5 * (di + m.binEdge) + 2.0 * S);  // (binEdge = the largest diameter)
  for (int z = br.z0; z <= br.z1; ++z)
    for (int y = br.y0; y <= br.y1; ++y)
      for (int x = br.x0; x <= br.x1; ++x) {
        const int bin = x + m.nbx * (y + m.nby * z);
        const int nb = min(m.binCnt[bin], m.binCap);
        for (int q = 0; q < nb; ++q) {
          const size_t o = (size_t)bin * m.binCap + q;
          int j, mult;
          bin_entry(m.binItems[o], j, mult);
          if (j == i) continue;
          const double cut = 0.5 * (di + dmax_of(classes[in.cls[j]])) + 2.0 * S;
          const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
          if (ddx * ddx + ddy * ddy + ddz * ddz <= cut * cut * (1.0 + 1e-9)) u = fmax(u, src[j]);
        }
      }
  return u;
}

// one hop of the neighbourhood maximum: dst_i = scale * max(src_i, src_j over the bodies and images near instance i AND near
// the places where a copy of i may appear this step) + floor
// (threads follow the Morton order of the schedule, `order`: neighbouring threads scan neighbouring bins)
__global__ void k5q_nbrmax(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ src,
                           double* __restrict__ dst, double scale, double floorB, int doCyclic, const int* __restrict__ order, int nOrder) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m.nInst) return;
  const int i = (t < nOrder) ? order[t] : t;
  double u = src[i];
  const double S = split_S(m);
  if (in.onMesh[i] != 0 && m.moving && 2.0 * S <= m.binEdge) {
    const D3 c = ld3(in.c, i);
    const ShapeClassDev* sc = classes + in.cls[i];
    const double di = dmax_of(*sc);
    u = nbr_scan_max(m, in, classes, src, c, di, i, S, u);
    if (doCyclic) {
      D3 sh[3];
      int n;
      if (image_shifts(m, c, sc->Rbound + S, sh, n))
        for (int sub = 1; sub < (1 << n); ++sub) {
          D3 g = c;
          for (int q = 0; q < n; ++q)
            if (sub & (1 << q)) g = g + sh[q];
          u = nbr_scan_max(m, in, classes, src, g, di, i, S, u);
        }
    }
  }
  dst[i] = scale * u + floorB;
}

// the instances of one index are one body (they exchange state every sub-step): they share the largest budget of the group
__global__ void k5q_group_budget(MotionDev m) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m.nParticles) return;
  const int head = m.idxHead[idx];
  if (head < 0 || m.instNext[head] < 0) return;
  double b = 0.0;
  for (int p = head; p >= 0; p = m.instNext[p]) b = fmax(b, m.bud[p]);
  for (int p = head; p >= 0; p = m.instNext[p]) m.bud[p] = b;
}

__global__ void k5q_classify(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const int* __restrict__ order, int nOrder) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double S = split_S(m);
  if (t == 0) *m.budget = S;
  if (t >= *m.nInst) return;
  const int i = (t < nOrder) ? order[t] : t;
  const D3 c = ld3(in.c, i);
  int quiet = 1;
  if (in.onMesh[i] != 0) {  // (an instance that is off the mesh does not move at all: particleIBM.C:400-403)
    const ShapeClassDev* sc = classes + in.cls[i];
    const double Bi = m.bud[i];
    const int head = m.idxHead[in.index[i]];
    if (head >= 0 && m.instNext[head] >= 0) quiet = 0;
    if (*m.noSplit) quiet = 0;
    if (quiet && m.moving) {
      const double reach = sc->Rbound + Bi;
      for (int w = 0; w < m.nWalls && quiet; ++w)
        if (patch_in_reach(m.patches[m.wallPatch[w]], c, reach)) quiet = 0;
      for (int s = 0; s < m.nCyc && quiet; ++s)
        if (patch_in_reach(m.patches[m.cycPatch[s]], c, reach)) quiet = 0;
    }
    if (quiet && m.moving) {
      if (2.0 * S > m.binEdge) quiet = 0;  // the +-2-bin search below no longer covers the cut-off: no split this step
      else {
        const double di = dmax_of(*sc);
        // (every budget is at most S: bud = safety * (a maximum of s) + floor, split_S; an image counts up to 4 budgets)
        const BinRange br = bin_range(m, c, 0.5 * (di + m.binEdge) + Bi + 4.0 * S);
        for (int z = br.z0; z <= br.z1 && quiet; ++z)
          for (int y = br.y0; y <= br.y1 && quiet; ++y)
            for (int x = br.x0; x <= br.x1 && quiet; ++x) {
              const int bin = x + m.nbx * (y + m.nby * z);
              const int nb = m.binCnt[bin];
              if (nb > m.binCap) { quiet = 0; break; }  // overflowing bin: entries are missing, decide nothing from it
              for (int q = 0; q < nb; ++q) {
                const size_t o = (size_t)bin * m.binCap + q;
                int j, mult;
                bin_entry(m.binItems[o], j, mult);
                if (j == i) continue;
                const double cut = 0.5 * (di + dmax_of(classes[in.cls[j]])) + Bi + mult * m.bud[j];
                const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
                if (ddx * ddx + ddy * ddy + ddz * ddz <= cut * cut * (1.0 + 1e-9)) { quiet = 0; break; }
              }
            }
      }
    }
  }
  m.quiet[i] = quiet;
  if (!quiet) m.act[atomicAdd(m.nAct, 1)] = i;
}

// every sub-step of a quiet instance, in order, by one thread
__global__ void k5q_quiet(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ dts,
                          int nSub) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (runs before the first sub-step: *m.nInst is the count at the start of the step)
  if (i >= *m.nInst || !m.quiet[i] || in.onMesh[i] == 0) return;
  for (int s = 0; s < nSub; ++s) d5_solve(m, L, in, classes, dts[s], m.cCnt, m.cIdx, m.cF, m.cT, m.Fw, m.Tw, m.cCnt, m.cCnt2, i);
}

// the per-sub-step chain on the compact list of active instances
#define IBM_ACT_I                                                  \
  const int t_ = blockIdx.x * blockDim.x + threadIdx.x;            \
  if (t_ >= *m.nAct) return;                                       \
  const int i = m.act[t_]
__global__ void k5a_bin_cand(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, int doCyclic) {
  IBM_ACT_I;
  d5_bin(m, in, i);
  if (doCyclic) d5_cyclic_candidates(m, in, classes, i);
}
__global__ void k5a_rows_direct(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt) {
  IBM_ACT_I;
  d5_row_direct(m, in, classes, dt, i);  // pair search, pair evaluation and row finalisation of this instance
  if (m.nWalls > 0) d5_walls(m, L, in, classes, dt, i);
}
__global__ void k5a_merge(MotionDev m, InstanceSoA in) {
  IBM_ACT_I;
  d5_merge(m, in, i);
}
__global__ void k5a_solve_clear(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, double dt,
                                const int* cCnt, const int* cIdx, const double* cF, const double* cT, double* Fw, double* Tw, int* cCntA,
                                int* cCntB) {
  IBM_ACT_I;
  if (i < *m.nInst && in.onMesh[i] != 0) {  // the bin this instance was put in (its centre has not changed since): ready for the next sub-step
    int bx, by, bz;
    bin_of(m, ld3(in.c, i), bx, by, bz);
    m.binCnt[bx + m.nbx * (by + m.nby * bz)] = 0;
  }
  d5_solve(m, L, in, classes, dt, cCnt, cIdx, cF, cT, Fw, Tw, cCntA, cCntB, i);
  if (i < *m.nInst) {
    const double B = m.check * m.bud[i];
    const D3 d0 = ld3(in.c, i) - D3{m.c0[i], m.c0[(size_t)m.cap + i], m.c0[2 * (size_t)m.cap + i]};
    if (magSqr3(d0) > B * B) atomicExch(m.violated, 1);
  }
}
__global__ void k5a_set_positions(MotionDev m, InstanceSoA in) {
  IBM_ACT_I;
  if (i >= *m.nInst) return;
  const int idx = in.index[i];
  if (m.idxHead[idx] != i) return;  // the first instance of a group does the group
  d5_set_positions(m, in, idx);
}

__global__ void k5q_snapshot(MotionDev m, InstanceSoA in, double* __restrict__ snap, int* __restrict__ snapActive, int* __restrict__ snapMaxCopies,
                             int restore) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t cap = (size_t)m.cap;
  if (i < m.cap) {
    for (int q = 0; q < 3; ++q) {
      if (restore) { in.c[q][i] = snap[q * cap + i]; in.th[q][i] = snap[(3 + q) * cap + i]; in.v[q][i] = snap[(6 + q) * cap + i]; in.om[q][i] = snap[(9 + q) * cap + i]; }
      else { snap[q * cap + i] = in.c[q][i]; snap[(3 + q) * cap + i] = in.th[q][i]; snap[(6 + q) * cap + i] = in.v[q][i]; snap[(9 + q) * cap + i] = in.om[q][i]; }
    }
    if (restore) in.active[i] = snapActive[i]; else snapActive[i] = in.active[i];
  }
  if (i < m.nParticles) {
    if (restore) m.maxCopies[i] = snapMaxCopies[i]; else snapMaxCopies[i] = m.maxCopies[i];
  }
}


// ------------------------------------------------------------------------------------------------
// Cluster mode of the quiet / active split (option cluster_loop5).
//
// The split above still pays ~7 dependent launches (or 3-4 grid-wide barriers) per sub-step, 101 times per fluid step,
// whatever the size of the active list: loop 5 is LATENCY bound.  But the classification knows more than "active": two
// active instances can only influence each other during the step if they are linked by a chain of possible interactions
// (a pair within (Dmax_i + Dmax_j)/2 + B_i + B_j, an image of one within reach of the other, the instances of one index).
// The connected components of that graph are independent, so every component runs ALL its sub-steps inside ONE thread
// block with block-level barriers — no kernel boundary, no grid barrier, as many blocks as there are groups of components:
//
//   k5c_edges        active instance -> its possible partners (the very cut-offs of k5q_classify) and its group head
//   k5c_components   connected components (hook roots + pointer jumping, one block)
//   k5c_count/place  members grouped by component, components packed into blocks by their offset
//   k5c_run          per block: the sub-step loop of particleCloud::solve over its members with the same device functions
//                    as every other mode (d5_pair_eval, d5_rows_finish, d5_walls, d5_merge, d5_cyclic_*, d5_solve,
//                    d5_set_positions); candidates of a pair search are the members of the same component
//   host             periodic copies created by different blocks took provisional slots in arbitrary order: their creation
//                    records (sub-step, patch, source) are sorted into the reference's append order and the new instances
//                    permuted accordingly (the order within one index group, which the physics depends on, is already
//                    the reference's inside a block).
// The budgets are CHECKED as in the plain split; a violation, or a block that runs out of room, redoes the step unsplit.
// ------------------------------------------------------------------------------------------------
#define K5C_THREADS 128
#define K5C_CAP 160        // members a block can hold (including the copies it creates)
#define K5C_FILL 56        // components are assigned to blocks by offset / K5C_FILL
#define K5C_CAND 96
#define K5C_HIT 48
#define K5C_EVN 4096      // creation records read back with the flags of a step (more: a second copy)

__device__ __forceinline__ void k5c_emit(const MotionDev& m, int a, int b) {
  const int h = atomicAdd(m.nEdge, 1);
  if (h < m.edgeCap) { m.edgeA[h] = a; m.edgeB[h] = b; }
  else atomicExch(m.clusterFail, 1);
}

// k5q_classify and k5c_edges in one pass (cluster mode): quiet / active decision with the cut-offs of k5q_classify, every
// possible partner emitted as an edge, the group head linked, label initialised
__global__ void k5c_classify(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const int* __restrict__ order, int nOrder) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double S = split_S(m);
  if (t == 0) *m.budget = S;
  if (t >= *m.nInst) return;
  const int i = (t < nOrder) ? order[t] : t;
  int quiet = 1;
  if (in.onMesh[i] != 0) {  // (an instance that is off the mesh does not move at all: particleIBM.C:400-403)
    const D3 c = ld3(in.c, i);
    const ShapeClassDev* sc = classes + in.cls[i];
    const double Bi = m.bud[i];
    const int head = m.idxHead[in.index[i]];
    const bool multi = head >= 0 && m.instNext[head] >= 0;
    if (multi) quiet = 0;
    if (*m.noSplit) { quiet = 0; atomicExch(m.clusterFail, 1); }
    if (m.moving) {
      if (quiet) {
        const double reach = sc->Rbound + Bi;
        for (int w = 0; w < m.nWalls && quiet; ++w)
          if (patch_in_reach(m.patches[m.wallPatch[w]], c, reach)) quiet = 0;
        for (int s = 0; s < m.nCyc && quiet; ++s)
          if (patch_in_reach(m.patches[m.cycPatch[s]], c, reach)) quiet = 0;
      }
      if (2.0 * S > m.binEdge) { quiet = 0; atomicExch(m.clusterFail, 1); }  // no bound on the interactions: no clusters this step
      else {
        const double di = dmax_of(*sc);
        const BinRange br = bin_range(m, c, 0.5 * (di + m.binEdge) + Bi + 4.0 * S);
        for (int z = br.z0; z <= br.z1; ++z)
          for (int y = br.y0; y <= br.y1; ++y)
            for (int x = br.x0; x <= br.x1; ++x) {
              const int bin = x + m.nbx * (y + m.nby * z);
              const int nb = m.binCnt[bin];
              if (nb > m.binCap) { quiet = 0; atomicExch(m.clusterFail, 1); continue; }  // entries are missing
              for (int q = 0; q < nb; ++q) {
                const size_t o = (size_t)bin * m.binCap + q;
                int j, mult;
                bin_entry(m.binItems[o], j, mult);
                if (j == i) continue;
                const double cut = 0.5 * (di + dmax_of(classes[in.cls[j]])) + Bi + mult * m.bud[j];
                const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
                if (ddx * ddx + ddy * ddy + ddz * ddz <= cut * cut * (1.0 + 1e-9)) { quiet = 0; k5c_emit(m, i, j); }
              }
            }
      }
    }
    if (multi && head != i) k5c_emit(m, i, head);
  }
  m.quiet[i] = quiet;
  if (!quiet) { m.act[atomicAdd(m.nAct, 1)] = i; m.label[i] = i; }
}

// possible partners of every active instance: same scan and cut-offs as k5q_classify, all hits instead of the first
__global__ void k5c_edges(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m.nAct) return;
  const int i = m.act[t];
  m.label[i] = i;
  if (in.onMesh[i] == 0) return;
  const int head = m.idxHead[in.index[i]];
  if (head >= 0 && head != i) k5c_emit(m, i, head);
  if (!m.moving) return;
  const double S = split_S(m);
  if (*m.noSplit || 2.0 * S > m.binEdge) { atomicExch(m.clusterFail, 1); return; }  // everything is active for lack of a bound: no clusters
  const ShapeClassDev* sc = classes + in.cls[i];
  const D3 c = ld3(in.c, i);
  const double Bi = m.bud[i], di = dmax_of(*sc);
  const BinRange br = bin_range(m, c, 0.5 * (di + m.binEdge) + Bi + 4.0 * S);
  for (int z = br.z0; z <= br.z1; ++z)
    for (int y = br.y0; y <= br.y1; ++y)
      for (int x = br.x0; x <= br.x1; ++x) {
        const int bin = x + m.nbx * (y + m.nby * z);
        const int nb = m.binCnt[bin];
        if (nb > m.binCap) { atomicExch(m.clusterFail, 1); continue; }
        for (int q = 0; q < nb; ++q) {
          const size_t o = (size_t)bin * m.binCap + q;
          int j, mult;
          bin_entry(m.binItems[o], j, mult);
          if (j == i) continue;
          const double cut = 0.5 * (di + dmax_of(classes[in.cls[j]])) + Bi + mult * m.bud[j];
          const double ddx = m.binPos[3 * o] - c.x, ddy = m.binPos[3 * o + 1] - c.y, ddz = m.binPos[3 * o + 2] - c.z;
          if (ddx * ddx + ddy * ddy + ddz * ddz <= cut * cut * (1.0 + 1e-9)) k5c_emit(m, i, j);
        }
      }
}

// connected components of the active instances: label[i] = smallest instance of the component (one block)
__global__ void __launch_bounds__(1024) k5c_components(MotionDev m) {
  __shared__ int changed;
  const int nE = min(*m.nEdge, m.edgeCap), nA = *m.nAct;
  // an edge may name an instance that is NOT in the active list only if the classification is inconsistent: guard
  for (int e = threadIdx.x; e < nE; e += blockDim.x)
    if (m.quiet[m.edgeB[e]] || m.quiet[m.edgeA[e]]) atomicExch(m.clusterFail, 1);
  __syncthreads();
  if (*m.clusterFail) return;
  for (int it = 0; it < 4096; ++it) {
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < nE; e += blockDim.x) {
      const int ra = m.label[m.edgeA[e]], rb = m.label[m.edgeB[e]];  // roots after the compression below
      if (ra != rb) { atomicMin(&m.label[max(ra, rb)], min(ra, rb)); changed = 1; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nA; t += blockDim.x) {
      const int i = m.act[t];
      int l = m.label[i];
      while (m.label[l] != l) l = m.label[l];
      m.label[i] = l;
    }
    __syncthreads();
    if (!changed) break;
    __syncthreads();
  }
}

// the same in rounds of two grid-wide kernels (hook the roots of every edge, then compress every path): the single block above
// spends 0.19 ms on 2.5 x 10^4 active instances; a last hook pass with `check` raises clusterFail if an edge still joins two roots
__global__ void k5c_cc_guard(MotionDev m) {
  const int nE = min(*m.nEdge, m.edgeCap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x)
    if (m.quiet[m.edgeB[e]] || m.quiet[m.edgeA[e]]) atomicExch(m.clusterFail, 1);
}
__global__ void k5c_cc_hook(MotionDev m, int check) {
  const int nE = min(*m.nEdge, m.edgeCap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
    const int ra = m.label[m.edgeA[e]], rb = m.label[m.edgeB[e]];
    if (ra != rb) {
      if (check) atomicExch(m.clusterFail, 1);
      else atomicMin(&m.label[max(ra, rb)], min(ra, rb));
    }
  }
}
__global__ void k5c_cc_compress(MotionDev m) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m.nAct) return;
  const int i = m.act[t];
  int l = m.label[i];
  while (m.label[l] != l) l = m.label[l];
  m.label[i] = l;
}

__global__ void k5c_count(MotionDev m) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m.nAct) return;
  atomicAdd(&m.compCnt[m.label[m.act[t]]], 1);
}
__global__ void k5c_place(MotionDev m) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m.nAct) return;
  const int i = m.act[t], r = m.label[i];
  m.sorted[m.compOff[r] + atomicAdd(&m.compCur[r], 1)] = i;
  if (i == r) {  // the root places its component into a block
    const int off = m.compOff[r], cnt = m.compCnt[r];
    const int b = off / K5C_FILL;
    if (b >= m.ctaBins) { atomicExch(m.clusterFail, 1); return; }
    atomicMin(&m.ctaBegin[b], off);
    atomicMax(&m.ctaEnd[b], off + cnt);
  }
}

// rank that owns a centre at height z (owned mode): the slab that holds the z-plane of the cell the centre lies in, centres
// below / above the mesh going to the first / last slab
__device__ __forceinline__ int owner_of(const MotionDev& m, double z) {
  const double q = floor((z - m.ownOz) / m.ownDz);
  int K = q > 0.0 ? (q < (double)(m.ownNz - 1) ? (int)q : m.ownNz - 1) : 0;  // (a NaN ends up in plane 0 on every rank)
  int r = 0;
  while (r + 1 < m.nRanks && K >= m.slabK0[r + 1]) ++r;
  return r;
}

// every sub-step of a quiet instance by one thread, state in registers: the arithmetic of d5_solve with empty contact and
// wall tables (sums of nothing are exact zeros, particleIBM.C:408-414), operation for operation
__global__ void k5c_quiet(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ dts,
                          int nSub) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *m.nInst || !m.quiet[i] || in.onMesh[i] == 0) return;
  if (m.owned) {
    if (owner_of(m, in.c[2][i]) != m.rank) return;  // somebody else's: its new state arrives with the exchange
    m.mine[i] = 1;
  }
  const ShapeClassDev* sc = classes + in.cls[i];
  const double mass = sc->V * in.rho[i];
  const D3 Fh = ld3(in.F, i), Th = ld3(in.T, i);
  D3 v = ld3(in.v, i), c = ld3(in.c, i), om = ld3(in.om, i), th = ld3(in.th, i);
  const double mI = mass * sc->I;
  // F/m + g and dt T/(m I) have the same operands in every sub-step (one dt for all but possibly the last): evaluated once
  D3 acc{0, 0, 0}, dW{0, 0, 0};
  double dtW = -1.0;
  if (m.moving) {
    D3 sFc{0, 0, 0}, sFw{0, 0, 0};
    for (int w = 0; w < m.nWalls; ++w) sFw = sFw + D3{0.0, 0.0, 0.0};
    const D3 F = Fh + sFc + sFw;
    acc = F / mass + D3{m.gx, m.gy, m.gz};
  }
  for (int s = 0; s < nSub; ++s) {
    const double dt = dts[s];
    if (m.moving) {
      v = v + dt * acc;
      const D3 dx = dt * v;
      c = c + dx;
    }
    if (m.rotating) {
      if (dtW != dt) {
        D3 sTc{0, 0, 0}, sTw{0, 0, 0};
        for (int w = 0; w < m.nWalls; ++w) sTw = sTw + D3{0.0, 0.0, 0.0};
        const D3 T = Th + sTc + sTw;
        dW = dt * T / mI;
        dtW = dt;
      }
      om = om + dW;
      th = th + om * dt;
      if (th.z < 0) th.z += IBM_TWOPI;
      if (th.z >= IBM_TWOPI) th.z -= IBM_TWOPI;
    }
  }
  if (nSub > 0) {
    if (m.moving) {
      st3(in.v, i, v); st3(in.c, i, c);
      int I, J, K;
      in.active[i] = find_cell(L, c.x, c.y, c.z, I, J, K) >= 0;
    }
    if (m.rotating) { st3(in.om, i, om); st3(in.th, i, th); }
  }
}

#define K5C_NP 6            // possible contact partners a member keeps in its list (more: it scans its whole component)
#define K5C_PATCH 8         // wall / cyclic patches a block keeps in shared memory (more: every sub-step takes the general path)

struct K5cPatch { double lo[3], hi[3], nrm[3], shift[3]; int planar, uniformShift, type, lattice, axis; };

struct K5cShared {
  // state of the members, resident for the whole fluid step
  double px[K5C_CAP], py[K5C_CAP], pz[K5C_CAP], vx[K5C_CAP], vy[K5C_CAP], vz[K5C_CAP];
  double wx[K5C_CAP], wy[K5C_CAP], wz[K5C_CAP], tx[K5C_CAP], ty[K5C_CAP], tz[K5C_CAP];
  double Fx[K5C_CAP], Fy[K5C_CAP], Fz[K5C_CAP], Tx[K5C_CAP], Ty[K5C_CAP], Tz[K5C_CAP];
  double mass[K5C_CAP], mI[K5C_CAP], rad[K5C_CAP], dmx[K5C_CAP], bud[K5C_CAP];
  double c0x[K5C_CAP], c0y[K5C_CAP], c0z[K5C_CAP];
  int mem[K5C_CAP];       // instance (cloud position or provisional slot) of every member
  int comp[K5C_CAP];      // its component root
  int index[K5C_CAP];
  int flags[K5C_CAP];     // K5F_*
  unsigned char nbr[K5C_CAP][K5C_NP], nNbr[K5C_CAP];  // members that can come within contact distance this step (from c0 and the budgets)
  unsigned short pmask[K5C_CAP];            // walls (bits 0-7) and cyclic patches (bits 8-15) within reach + budget of where it started
  unsigned short ring[K5C_CAP];             // next member slot with the same index (circular; itself when alone)
  unsigned short cb[K5C_CAP], ce[K5C_CAP];  // slots [cb, ce) hold the component of the member as it was at the start of the step;
                                            // copies created since sit in [nMem0, nMem)
  K5cPatch wall[K5C_PATCH], cyc[K5C_PATCH];
  int candKey[K5C_CAND];
  int hitKey[K5C_HIT];
  double hitCenter[3 * K5C_HIT];
  int nMem, nMem0, candCount, hitCount, bigN, base, fail, patchesFit, evMask[2];
};
#define K5F_ONMESH 1
#define K5F_ACTIVE 2
#define K5F_MULTI 4     // its index has several instances
#define K5F_ISO 8       // isotropic in the sense of particleIBM.C:60-74 (App. B #21)
#define K5F_SCANALL 32  // more possible partners than the list holds
#define K5F_NOPATCH 16  // no wall / cyclic patch within reach + budget of where it started: patch tests skipped for the whole step

// member slot <- global instance arrays
__device__ __forceinline__ void k5c_load(K5cShared& sh, int t, const MotionDev& m, const InstanceSoA& in, const ShapeClassDev* __restrict__ classes,
                                         int doCyclic) {
  const int i = sh.mem[t];
  const ShapeClassDev* sc = classes + in.cls[i];
  sh.px[t] = in.c[0][i]; sh.py[t] = in.c[1][i]; sh.pz[t] = in.c[2][i];
  sh.vx[t] = in.v[0][i]; sh.vy[t] = in.v[1][i]; sh.vz[t] = in.v[2][i];
  sh.wx[t] = in.om[0][i]; sh.wy[t] = in.om[1][i]; sh.wz[t] = in.om[2][i];
  sh.tx[t] = in.th[0][i]; sh.ty[t] = in.th[1][i]; sh.tz[t] = in.th[2][i];
  sh.Fx[t] = in.F[0][i]; sh.Fy[t] = in.F[1][i]; sh.Fz[t] = in.F[2][i];
  sh.Tx[t] = in.T[0][i]; sh.Ty[t] = in.T[1][i]; sh.Tz[t] = in.T[2][i];
  const double mass = sc->V * in.rho[i];
  sh.mass[t] = mass;
  sh.mI[t] = mass * sc->I;
  sh.rad[t] = sc->Rbound;
  const D3 D = classD(*sc);
  sh.dmx[t] = cmptMax3(D);
  sh.bud[t] = m.bud[i];
  sh.c0x[t] = m.c0[i]; sh.c0y[t] = m.c0[(size_t)m.cap + i]; sh.c0z[t] = m.c0[2 * (size_t)m.cap + i];
  const int idx = in.index[i];
  sh.index[t] = idx;
  const int head = m.idxHead[idx];
  int f = 0;
  if (in.onMesh[i] != 0) f |= K5F_ONMESH;
  if (in.active[i]) f |= K5F_ACTIVE;
  if (head >= 0 && m.instNext[head] >= 0) f |= K5F_MULTI;
  if ((D.x + D.y + D.z) / 3.0 == cmptMax3(D)) f |= K5F_ISO;
  if (sh.patchesFit) {
    const D3 c0{sh.c0x[t], sh.c0y[t], sh.c0z[t]};
    const double reach = sc->Rbound + 1.001 * sh.bud[t] + 1e-9 * m.binEdge;
    int mask = 0;
    for (int w = 0; w < m.nWalls; ++w) {
      const K5cPatch& p = sh.wall[w];
      const double dx = fmax(fmax(p.lo[0] - c0.x, c0.x - p.hi[0]), 0.0), dy = fmax(fmax(p.lo[1] - c0.y, c0.y - p.hi[1]), 0.0),
                   dz = fmax(fmax(p.lo[2] - c0.z, c0.z - p.hi[2]), 0.0);
      if (dx * dx + dy * dy + dz * dz <= reach * reach) mask |= 1 << w;
    }
    for (int q = 0; q < m.nCyc && doCyclic; ++q) {
      const K5cPatch& p = sh.cyc[q];
      const double dx = fmax(fmax(p.lo[0] - c0.x, c0.x - p.hi[0]), 0.0), dy = fmax(fmax(p.lo[1] - c0.y, c0.y - p.hi[1]), 0.0),
                   dz = fmax(fmax(p.lo[2] - c0.z, c0.z - p.hi[2]), 0.0);
      if (dx * dx + dy * dy + dz * dz <= reach * reach) mask |= 256 << q;
    }
    sh.pmask[t] = (unsigned short)mask;
    if (!mask) f |= K5F_NOPATCH;
  }
  sh.flags[t] = f;
}

// kinematic state of a member slot <-> global arrays (around a sub-step that takes the general path)
__device__ __forceinline__ void k5c_store_state(const K5cShared& sh, int t, const InstanceSoA& in) {
  const int i = sh.mem[t];
  in.c[0][i] = sh.px[t]; in.c[1][i] = sh.py[t]; in.c[2][i] = sh.pz[t];
  in.v[0][i] = sh.vx[t]; in.v[1][i] = sh.vy[t]; in.v[2][i] = sh.vz[t];
  in.om[0][i] = sh.wx[t]; in.om[1][i] = sh.wy[t]; in.om[2][i] = sh.wz[t];
  in.th[0][i] = sh.tx[t]; in.th[1][i] = sh.ty[t]; in.th[2][i] = sh.tz[t];
  in.active[i] = (sh.flags[t] & K5F_ACTIVE) ? 1 : 0;
}
__device__ __forceinline__ void k5c_reload_state(K5cShared& sh, int t, const MotionDev& m, const InstanceSoA& in) {
  const int i = sh.mem[t];
  sh.px[t] = in.c[0][i]; sh.py[t] = in.c[1][i]; sh.pz[t] = in.c[2][i];
  sh.vx[t] = in.v[0][i]; sh.vy[t] = in.v[1][i]; sh.vz[t] = in.v[2][i];
  sh.wx[t] = in.om[0][i]; sh.wy[t] = in.om[1][i]; sh.wz[t] = in.om[2][i];
  sh.tx[t] = in.th[0][i]; sh.ty[t] = in.th[1][i]; sh.tz[t] = in.th[2][i];
  int f = sh.flags[t] & ~(K5F_ACTIVE | K5F_MULTI);
  if (in.active[i]) f |= K5F_ACTIVE;
  const int head = m.idxHead[sh.index[t]];
  if (head >= 0 && m.instNext[head] >= 0) f |= K5F_MULTI;  // (its index may have got a copy in this sub-step)
  sh.flags[t] = f;
}

// slots that can hold a member of the same component as slot t: its component's range at the start of the step, then the
// copies the block has created since
#define K5C_FOR_COMPONENT(sh, t, u) \
  for (int u = (sh).cb[t], u##_e = (sh).ce[t]; u < (sh).nMem; u = (u + 1 == u##_e && u##_e <= (sh).nMem0) ? max((sh).nMem0, u##_e) : u + 1)

// Possible contact partners of slot t during this step: both bodies stay within their budget of where they started (checked
// every sub-step), so only pairs whose START centres are within (Dmax_t + Dmax_u)/2 + B_t + B_u can pass particleIBM.C:56-59.
__device__ __forceinline__ void k5c_build_nbrs(K5cShared& sh, int t) {
  int n = 0, f = sh.flags[t] & ~K5F_SCANALL;
  const D3 c0{sh.c0x[t], sh.c0y[t], sh.c0z[t]};
  K5C_FOR_COMPONENT(sh, t, u) {
    if (u == t || sh.comp[u] != sh.comp[t] || sh.index[u] == sh.index[t]) continue;
    const double lim = 0.5 * (sh.dmx[t] + sh.dmx[u]) + 1.0001 * (sh.bud[t] + sh.bud[u]) + 1e-9 * (sh.dmx[t] + sh.dmx[u]);
    if (magSqr3(D3{sh.c0x[u], sh.c0y[u], sh.c0z[u]} - c0) > lim * lim) continue;
    if (n < K5C_NP) sh.nbr[t][n++] = (unsigned char)u; else f |= K5F_SCANALL;
  }
  sh.nNbr[t] = (unsigned char)n;
  sh.flags[t] = f;
}
// the slots to test against slot t for a contact: its list, or its whole component when the list overflowed
#define K5C_FOR_PARTNERS(sh, t, u, ...)                                                      \
  if ((sh).flags[t] & K5F_SCANALL) { K5C_FOR_COMPONENT(sh, t, u) { __VA_ARGS__ } }           \
  else { for (int u##_q = 0; u##_q < (int)(sh).nNbr[t]; ++u##_q) { const int u = (sh).nbr[t][u##_q]; __VA_ARGS__ } }

// next member with the same index after slot t, cyclically (needs index[] of every member)
__device__ __forceinline__ void k5c_link_ring(K5cShared& sh, int t) {
  int nxt = -1, first = -1;
  K5C_FOR_COMPONENT(sh, t, u) {
    if (sh.index[u] != sh.index[t]) continue;
    if (first < 0) first = u;
    if (u > t && nxt < 0) nxt = u;
  }
  sh.ring[t] = (unsigned short)(nxt >= 0 ? nxt : first);
}

// Is a face centre of the patch within r of c?  patchHit / wallHit need one (particleIBM.C:466-472, 505-511).  Exact-enough for
// lattice-ordered planar patches (face centres = transverse cell centres), "maybe" otherwise: a superset either way.
__device__ __forceinline__ bool k5c_face_in_reach(const K5cPatch& p, const LatticeDev& L, D3 inv, D3 c, double r) {
  const double dx = fmax(fmax(p.lo[0] - c.x, c.x - p.hi[0]), 0.0), dy = fmax(fmax(p.lo[1] - c.y, c.y - p.hi[1]), 0.0),
               dz = fmax(fmax(p.lo[2] - c.z, c.z - p.hi[2]), 0.0);
  const double r2 = r * r * (1.0 + 1e-9) + 1e-300;
  if (!(dx * dx + dy * dy + dz * dz <= r2)) return false;
  if (!p.lattice) return true;
  // distance to the nearest face centre: the patch plane along its axis, the centre of the containing transverse cell otherwise
  // (products with the inverse spacing: a rounding across a cell boundary picks a centre that is equally far away)
  auto axisDist = [&](int a, double ca, double o, double d, double id, int n, double plane) {
    if (a == p.axis) return ca - plane;
    double t = floor((ca - o) * id);
    t = fmin(fmax(t, 0.0), (double)(n - 1));
    return ca - (o + (t + 0.5) * d);
  };
  const double ex = axisDist(0, c.x, L.ox, L.dx, inv.x, L.nx, p.lo[0]), ey = axisDist(1, c.y, L.oy, L.dy, inv.y, L.ny, p.lo[1]),
               ez = axisDist(2, c.z, L.oz, L.dz, inv.z, L.nz, p.lo[2]);
  const double d2 = ex * ex + ey * ey + ez * ez;
  return d2 <= r2 * (1.0 + 1e-9) + 1e-18 * (L.dx * L.dx + L.dy * L.dy + L.dz * L.dz);
}

// 0 = nothing, 1 = only contacts between isotropic single-instance bodies under the reference's impulse law (the resident path
// evaluates them itself, k5c_contact_sum), 2 = anything else (the sub-step takes the general path)
__device__ __forceinline__ int k5c_event(const K5cShared& sh, int t, const MotionDev& m, const LatticeDev& L, D3 inv, int doCyclic, long long& tPair) {
  const long long tp0 = clock64();
  const int f = sh.flags[t];
  if (!(f & K5F_ONMESH) || !m.moving) return 0;
  int code = 0;
  const int comp = sh.comp[t], idx = sh.index[t];
  const D3 c{sh.px[t], sh.py[t], sh.pz[t]}, v{sh.vx[t], sh.vy[t], sh.vz[t]};
  bool other = false;
  K5C_FOR_PARTNERS(sh, t, u, {
    if (u == t || sh.comp[u] != comp || sh.index[u] == idx || !(sh.flags[u] & K5F_ONMESH)) continue;
    const D3 cu{sh.px[u], sh.py[u], sh.pz[u]};
    const double R = 0.5 * (sh.dmx[t] + sh.dmx[u]);
    if (magSqr3(cu - c) > R * R * (1.0 + 1e-9)) continue;             // (surely mag > R: spares the square root)
    if (mag3(cu - c) > R) continue;                                   // particleIBM.C:56-59
    if (m.model != 0 || !(f & K5F_ISO) || !(sh.flags[u] & K5F_ISO)) { other = true; continue; }  // overlap force without history / ellipse iteration
    if (!(R > mag3(cu - c))) continue;                                // particleIBM.C:60-74
    const D3 vu{sh.vx[u], sh.vy[u], sh.vz[u]};
    const bool iAmP1 = idx > sh.index[u];
    const D3 x12 = iAmP1 ? (c - cu) : (cu - c);
    const D3 v12 = iAmP1 ? (v - vu) : (vu - v);
    if (dot3(x12, v12) < 0) {
      if ((f | sh.flags[u]) & K5F_MULTI) other = true;  // copies share the larger entry (particleCloud.C:525-554): general path
      else code = 1;
    }
  })
  if (other) return 2;
  tPair += clock64() - tp0;
  if (f & K5F_NOPATCH) return code;
  if (!sh.patchesFit) return 2;
  const double r = sh.rad[t];
  const int pmask = sh.pmask[t];
  for (int w = 0; w < m.nWalls; ++w)
    if (((pmask >> w) & 1) && k5c_face_in_reach(sh.wall[w], L, inv, c, r)) return 2;
  if (doCyclic)
    for (int q = 0; q < m.nCyc; ++q) {  // d5_cyclic_candidates + the reach of a face centre (d5_cyclic_detect)
      if (!((pmask >> (8 + q)) & 1)) continue;
      const K5cPatch& p = sh.cyc[q];
      if (p.type != IBM_CYCLIC) continue;
      // every condition is necessary for a new copy; the cheap ones first (most bodies near a patch move away from it or have
      // their copy already)
      if (p.planar && !(dot3(v, D3{p.nrm[0], p.nrm[1], p.nrm[2]}) > 0)) continue;
      if (p.uniformShift) {
        const D3 nc = c + D3{p.shift[0], p.shift[1], p.shift[2]};
        bool have = false;
        for (int u = sh.ring[t]; u != t && !have; u = sh.ring[u]) have = magSqr3(nc - D3{sh.px[u], sh.py[u], sh.pz[u]}) < 0.25e-8;
        if (have) continue;
      }
      if (!k5c_face_in_reach(p, L, inv, c, r)) continue;
      return 2;
    }
  return code;
}

// sum(collisionForces_) of member t for a sub-step whose only events are contacts of isotropic single-instance bodies:
// every entry as d5_pair_eval writes it (particleCloud.C:505-521), summed in partner-index order (particleIBM.C:408-411).
// The torque entries are -+F ^ hitPt with hitPt = 0 (particleIBM.C:56-74): exact zeros.
__device__ __forceinline__ D3 k5c_contact_sum(const K5cShared& sh, int t, const MotionDev& m, double deltaT, bool& overflow) {
  D3 sum{0, 0, 0};
  if (!(sh.flags[t] & K5F_ONMESH)) return sum;
  const int comp = sh.comp[t], idx = sh.index[t];
  const D3 c{sh.px[t], sh.py[t], sh.pz[t]}, v{sh.vx[t], sh.vy[t], sh.vz[t]};
  int eIdx[4];
  D3 eF[4];
  int n = 0;
  K5C_FOR_PARTNERS(sh, t, u, {
    if (u == t || sh.comp[u] != comp || sh.index[u] == idx || !(sh.flags[u] & K5F_ONMESH)) continue;
    const D3 cu{sh.px[u], sh.py[u], sh.pz[u]};
    const double R = 0.5 * (sh.dmx[t] + sh.dmx[u]);
    if (magSqr3(cu - c) > R * R * (1.0 + 1e-9)) continue;
    if (mag3(cu - c) > R) continue;
    if (!(R > mag3(cu - c))) continue;
    const D3 vu{sh.vx[u], sh.vy[u], sh.vz[u]};
    const bool iAmP1 = idx > sh.index[u];
    const D3 x12 = iAmP1 ? (c - cu) : (cu - c);
    const D3 v12 = iAmP1 ? (v - vu) : (vu - v);
    if (!(dot3(x12, v12) < 0)) continue;
    const double mi = sh.mass[t], mj = sh.mass[u];
    double mass1 = iAmP1 ? mi : mj, mass2 = iAmP1 ? mj : mi;
    double M = mass1 * mass2 / (mass1 + mass2);
    double x12Dotv12 = dot3(x12, v12);
    double x12Sqr = magSqr3(x12);
    D3 F = 2.0 * M * x12Dotv12 / x12Sqr * x12 * m.e / deltaT;  // particleCloud.C:513
    const D3 myF = iAmP1 ? -F : F;
    if (n >= 4) { overflow = true; continue; }
    int q = n++;
    while (q > 0 && eIdx[q - 1] > sh.index[u]) { eIdx[q] = eIdx[q - 1]; eF[q] = eF[q - 1]; --q; }
    eIdx[q] = sh.index[u]; eF[q] = myF;
  })
  for (int q = 0; q < n; ++q) sum = sum + eF[q];
  return sum;
}

// member t over one sub-step: d5_solve with the given sum of contact forces (zero in free flight), zero contact torques and
// empty wall tables, operation for operation
// Free flight repeats two quotients with the same operands in every sub-step: F/m + g (contact sum zero) and dt T/(m I)
// (one dt for all sub-steps but possibly the last); a thread keeps them for the members it steps.
struct K5cCache { D3 acc; D3 dW; double dtW; int validAcc; };

__device__ __forceinline__ void k5c_free_flight(K5cShared& sh, int t, const MotionDev& m, const LatticeDev& L, double dt, D3 sFc, bool contact,
                                                bool needActive, K5cCache& ca) {
  if (!(sh.flags[t] & K5F_ONMESH)) return;
  if (m.moving) {
    D3 acc;
    if (contact || !ca.validAcc) {
      D3 sFw{0, 0, 0};
      for (int w = 0; w < m.nWalls; ++w) sFw = sFw + D3{0.0, 0.0, 0.0};
      const D3 F = D3{sh.Fx[t], sh.Fy[t], sh.Fz[t]} + sFc + sFw;
      acc = F / sh.mass[t] + D3{m.gx, m.gy, m.gz};
      if (!contact) { ca.acc = acc; ca.validAcc = 1; }
    } else acc = ca.acc;
    D3 v{sh.vx[t], sh.vy[t], sh.vz[t]};
    v = v + dt * acc;
    sh.vx[t] = v.x; sh.vy[t] = v.y; sh.vz[t] = v.z;
    const D3 dx = dt * v;
    const D3 c = D3{sh.px[t], sh.py[t], sh.pz[t]} + dx;
    sh.px[t] = c.x; sh.py[t] = c.y; sh.pz[t] = c.z;
    if (needActive) {
      // particleIBM.C:417-424.  active_ is read by setPositions of the copies of an index and by the log block after the last
      // sub-step: instances without copies only need the last one (a general sub-step recomputes it for everybody)
      int I, J, K;
      const bool act = find_cell(L, c.x, c.y, c.z, I, J, K) >= 0;
      sh.flags[t] = act ? (sh.flags[t] | K5F_ACTIVE) : (sh.flags[t] & ~K5F_ACTIVE);
    }
  }
  if (m.rotating) {
    D3 sTc{0, 0, 0}, sTw{0, 0, 0};
    for (int w = 0; w < m.nWalls; ++w) sTw = sTw + D3{0.0, 0.0, 0.0};
    if (ca.dtW != dt) {
      const D3 T = D3{sh.Tx[t], sh.Ty[t], sh.Tz[t]} + sTc + sTw;
      ca.dW = dt * T / sh.mI[t];
      ca.dtW = dt;
    }
    D3 om{sh.wx[t], sh.wy[t], sh.wz[t]};
    om = om + ca.dW;
    sh.wx[t] = om.x; sh.wy[t] = om.y; sh.wz[t] = om.z;
    D3 th = D3{sh.tx[t], sh.ty[t], sh.tz[t]} + om * dt;
    if (th.z < 0) th.z += IBM_TWOPI;
    if (th.z >= IBM_TWOPI) th.z -= IBM_TWOPI;
    sh.tx[t] = th.x; sh.ty[t] = th.y; sh.tz[t] = th.z;
  }
}

__global__ void __launch_bounds__(K5C_THREADS)
k5c_run(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ dts, int nSub,
        int doCyclic, int generalOnly) {
  __shared__ K5cShared sh;
  const long long tLaunch = clock64();
  const int b = blockIdx.x;
  if (b >= m.ctaBins) return;
  const int beg = m.ctaBegin[b], end = m.ctaEnd[b];
  if (end <= beg) return;
  if (*m.clusterFail) return;
  // a block belongs to the owner of the root (smallest instance) of its first component: the components sit at offsets that
  // follow from the labels alone, whereas the order of the members inside one is that of an atomic counter and differs
  // from rank to rank
  if (m.owned && owner_of(m, m.c0[2 * (size_t)m.cap + m.label[m.sorted[beg]]]) != m.rank) return;
  const int tid = threadIdx.x;
  if (end - beg > K5C_CAP) { if (tid == 0) atomicExch(m.clusterFail, 1); return; }
  if (tid == 0) {
    sh.nMem = end - beg; sh.nMem0 = end - beg; sh.evMask[0] = 0; sh.evMask[1] = 0; sh.candCount = 0; sh.hitCount = 0; sh.bigN = 0x7fffffff; sh.fail = 0;
    sh.patchesFit = m.nWalls <= K5C_PATCH && m.nCyc <= K5C_PATCH;
  }
  if (m.nWalls <= K5C_PATCH && m.nCyc <= K5C_PATCH && tid < m.nWalls + m.nCyc) {
    const bool isWall = tid < m.nWalls;
    const PatchDev& p = m.patches[isWall ? m.wallPatch[tid] : m.cycPatch[tid - m.nWalls]];
    K5cPatch& q = isWall ? sh.wall[tid] : sh.cyc[tid - m.nWalls];
    for (int a = 0; a < 3; ++a) { q.lo[a] = p.lo[a]; q.hi[a] = p.hi[a]; q.nrm[a] = p.nrm[a]; q.shift[a] = p.shift[a]; }
    q.planar = p.planar; q.uniformShift = p.uniformShift; q.type = p.type; q.lattice = p.latticeOrdered; q.axis = p.side / 2;
  }
  for (int t = tid; t < end - beg; t += K5C_THREADS) {
    const int i = m.sorted[beg + t], root = m.label[i];
    sh.mem[t] = i; sh.comp[t] = root;
    sh.cb[t] = (unsigned short)(m.compOff[root] - beg); sh.ce[t] = (unsigned short)(m.compOff[root] - beg + m.compCnt[root]);
  }
  __syncthreads();
  for (int t = tid; t < sh.nMem; t += K5C_THREADS) k5c_load(sh, t, m, in, classes, doCyclic);
  __syncthreads();
  for (int t = tid; t < sh.nMem; t += K5C_THREADS) { k5c_link_ring(sh, t); k5c_build_nbrs(sh, t); }
  __syncthreads();
  // this block's view of the motion data: private candidate / hit lists, no instance-count bound (members are valid by construction)
  MotionDev ml = m;
  ml.nInst = &sh.bigN;
  ml.candCount = &sh.candCount; ml.candKey = sh.candKey; ml.candCap = K5C_CAND;
  ml.hitCount = &sh.hitCount; ml.hitKey = sh.hitKey; ml.hitCenter = sh.hitCenter;
  ml.overflow = &sh.fail;
  const bool copies = m.nCyc > 0;
  const double cut = m.binEdge * m.binEdge * (1.0 + 1e-9);
  long long tEv = 0, tFl = 0, tSp = 0, tGen = 0, tPair = 0, tEvOwn = 0;
  const D3 invSpacing{1.0 / L.dx, 1.0 / L.dy, 1.0 / L.dz};
  K5cCache cache[(K5C_CAP + K5C_THREADS - 1) / K5C_THREADS];
  for (auto& ca : cache) { ca.validAcc = 0; ca.dtW = -1.0; }
  const long long tStart = clock64();
  for (int s = 0; s < nSub; ++s) {
    const double dt = dts[s];
    const long long t0 = clock64();
    // ---- can anything but free flight happen to a member in this sub-step? ----
    int evl = generalOnly ? 2 : 0;
    if (!evl)
      for (int t = tid; t < sh.nMem; t += K5C_THREADS) evl = max(evl, k5c_event(sh, t, m, L, invSpacing, doCyclic, tPair));
    tEvOwn += clock64() - t0;
    if (evl) atomicOr(&sh.evMask[s & 1], evl);
    __syncthreads();
    const int ev = sh.evMask[s & 1];
    if (tid == 0) sh.evMask[(s + 1) & 1] = 0;  // (nobody reads or writes the other word before the next barrier)
    const long long t1 = clock64();
    tEv += t1 - t0;
    if (tid == 0) { atomicAdd(&m.clusterStat[0], 1ULL); if (ev & 2) atomicAdd(&m.clusterStat[1], 1ULL); }
    if (!(ev & 2)) {
      // ---- no, or contacts the resident path evaluates itself: every member integrates (particleIBM.C:393-450), then copies
      //      exchange state (particleCloud.C:411-443) ----
      D3 sums[(K5C_CAP + K5C_THREADS - 1) / K5C_THREADS];
      bool over = false;
      if (ev & 1) {  // contact forces from the positions and velocities at the start of the sub-step, for everybody, THEN move
        int q = 0;
        for (int t = tid; t < sh.nMem; t += K5C_THREADS) sums[q++] = k5c_contact_sum(sh, t, m, dt, over);
        if (over) sh.fail = 8;
        __syncthreads();
      }
      int q = 0;
      for (int t = tid; t < sh.nMem; t += K5C_THREADS) {
        k5c_free_flight(sh, t, m, L, dt, (ev & 1) ? sums[q] : D3{0, 0, 0}, (ev & 1) != 0, s == nSub - 1 || (sh.flags[t] & K5F_MULTI) != 0, cache[q]);
        ++q;
        const double B = m.check * sh.bud[t];
        const D3 d0 = D3{sh.px[t], sh.py[t], sh.pz[t]} - D3{sh.c0x[t], sh.c0y[t], sh.c0z[t]};
        if (magSqr3(d0) > B * B) atomicExch(m.violated, 1);
      }
      __syncthreads();
      const long long t2 = clock64();
      tFl += t2 - t1;
      if (copies) {
        // setPositions in cloud order: the first inactive instance of an index hands theta, v, omega to all the others
        // (every later inactive one then holds that very state already).  Its own state does not change: no race.
        for (int t = tid; t < sh.nMem; t += K5C_THREADS) {
          if (!(sh.flags[t] & K5F_MULTI)) continue;
          int src = -1, srcInst = 0x7fffffff;
          if (!(sh.flags[t] & K5F_ACTIVE)) { src = t; srcInst = sh.mem[t]; }
          for (int u = sh.ring[t]; u != t; u = sh.ring[u])
            if (!(sh.flags[u] & K5F_ACTIVE) && sh.mem[u] < srcInst) { src = u; srcInst = sh.mem[u]; }
          if (src >= 0 && src != t) {
            sh.tx[t] = sh.tx[src]; sh.ty[t] = sh.ty[src]; sh.tz[t] = sh.tz[src];
            sh.vx[t] = sh.vx[src]; sh.vy[t] = sh.vy[src]; sh.vz[t] = sh.vz[src];
            sh.wx[t] = sh.wx[src]; sh.wy[t] = sh.wy[src]; sh.wz[t] = sh.wz[src];
          }
        }
        __syncthreads();
      }
      tSp += clock64() - t2;
      if (sh.fail) break;
      continue;
    }
    // ---- yes: this sub-step takes the general path on the global arrays ----
    const int nMemBefore = sh.nMem;
    for (int t = tid; t < sh.nMem; t += K5C_THREADS) k5c_store_state(sh, t, in);
    __syncthreads();
    if (m.moving) {
      const int nMem = sh.nMem;
      for (int t = tid; t < nMem; t += K5C_THREADS) {
        const int i = sh.mem[t];
        if (in.onMesh[i] != 0) {
          const D3 c = ld3(in.c, i);
          const int comp = sh.comp[t];
          K5C_FOR_COMPONENT(sh, t, u) {
            if (u == t || sh.comp[u] != comp) continue;
            const int j = sh.mem[u];
            const D3 dd = ld3(in.c, j) - c;
            if (dd.x * dd.x + dd.y * dd.y + dd.z * dd.z > cut) continue;  // the bin neighbourhood's cut-off (d5_find_pairs)
            if (in.onMesh[j] == 0) continue;                               // (off-mesh instances are never entered into the bins)
            d5_pair_eval(ml, in, classes, dt, i, j);
          }
          d5_rows_finish(ml, in, i);
          if (m.nWalls > 0) d5_walls(ml, L, in, classes, dt, i);
          if (doCyclic) d5_cyclic_candidates(ml, in, classes, i);
        }
      }
      __syncthreads();  // rows complete before anybody merges them or moves a centre
      if (copies) {
        for (int t = tid; t < nMem; t += K5C_THREADS) d5_merge(ml, in, sh.mem[t]);
        if (doCyclic) d5_cyclic_detect(ml, L, in, classes, tid, K5C_THREADS);
        __syncthreads();
        if (doCyclic) {
          const int total = sh.hitCount;
          if (total > 0) {
            // provisional slots for the new copies, in this block's (patch, source) order: the reference's order among the
            // instances this block owns; the order across blocks is settled on the host after the step
            if (tid == 0) {
              int base = -1;
              if (total <= K5C_HIT && sh.nMem + total <= K5C_CAP) {
                base = atomicAdd(m.nextSlot, total);
                if (base + total > m.cap) base = -1;
              }
              sh.base = base;
              if (base < 0) { sh.fail = 9; sh.hitCount = 0; sh.candCount = 0; }
            }
            __syncthreads();
            const int base = sh.base;
            if (base >= 0) {
              for (int h = tid; h < total; h += K5C_THREADS) {
                const int key = sh.hitKey[h];
                int rank = 0;
                for (int q = 0; q < total; ++q) rank += sh.hitKey[q] < key;
                const int src = key % m.cap, dst = base + rank;
                const int idx = in.index[src];
                in.index[dst] = idx;
                in.copy[dst] = -1;
                in.active[dst] = 0;
                in.onMesh[dst] = IBM_PARTIAL;
                in.cls[dst] = m.origCls[idx];
                for (int q = 0; q < 3; ++q) {
                  in.c[q][dst] = sh.hitCenter[3 * h + q];
                  in.th[q][dst] = in.th[q][src];
                  in.v[q][dst] = in.v[q][src];
                  in.om[q][dst] = in.om[q][src];
                  in.F[q][dst] = 0.0;  // integratedForce_ of origParticles_[index] (never updated after construction)
                  in.T[q][dst] = 0.0;
                  m.c0[(size_t)q * m.cap + dst] = sh.hitCenter[3 * h + q];
                }
                in.rho[dst] = m.origRho[idx];
                in.age[dst] = in.age[src];
                in.nShell[dst] = 0; in.shellOff[dst] = 0; in.nPtsOnMesh[dst] = 0;
                rot_from_theta(in.th[0][src], in.th[1][src], in.th[2][src], in.rot + 9 * (size_t)dst);
                m.bud[dst] = m.bud[src];
                m.quiet[dst] = 0;
                m.cCnt[dst] = 0;
                m.cCnt2[dst] = 0;
                for (int w = 0; w < m.nWalls; ++w)
                  for (int q = 0; q < 3; ++q) {
                    const size_t o = 3 * ((size_t)dst * m.nWalls + w) + q;
                    m.Fw[o] = 0.0; m.Tw[o] = 0.0; m.Fw2[o] = 0.0; m.Tw2[o] = 0.0;
                  }
                m.evSub[dst] = s;
                m.evKey[dst] = key;
                // the copy joins the component of its source
                int compSrc = 0, cbS = 0, ceS = 0;
                for (int u = 0; u < sh.nMem; ++u) if (sh.mem[u] == src) { compSrc = sh.comp[u]; cbS = sh.cb[u]; ceS = sh.ce[u]; }
                sh.mem[sh.nMem + rank] = dst;
                sh.comp[sh.nMem + rank] = compSrc;
                sh.cb[sh.nMem + rank] = (unsigned short)cbS; sh.ce[sh.nMem + rank] = (unsigned short)ceS;
              }
              __syncthreads();
              if (tid == 0) {  // sequential part: copy ids and the same-index lists, in append order (d5_append_link)
                for (int t2 = 0; t2 < total; ++t2) {
                  const int dst = base + t2;
                  const int idx = in.index[dst];
                  in.copy[dst] = m.maxCopies[idx] + 1;
                  m.maxCopies[idx] += 1;
                  m.instNext[dst] = -1;
                  const int tail = m.idxTail[idx];
                  if (tail >= 0) m.instNext[tail] = dst; else m.idxHead[idx] = dst;
                  m.idxTail[idx] = dst;
                }
                sh.nMem += total;
                sh.hitCount = 0;
                sh.candCount = 0;
              }
              __syncthreads();
              for (int t = sh.nMem - total + tid; t < sh.nMem; t += K5C_THREADS) k5c_load(sh, t, m, in, classes, doCyclic);  // the new members
            }
            __syncthreads();
          } else if (tid == 0) {
            sh.candCount = 0;
          }
        }
      }
    }
    {
      const int nMem = sh.nMem;
      for (int t = tid; t < nMem; t += K5C_THREADS) {
        const int i = sh.mem[t];
        d5_solve(ml, L, in, classes, dt, copies ? m.cCnt2 : m.cCnt, copies ? m.cIdx2 : m.cIdx, copies ? m.cF2 : m.cF, copies ? m.cT2 : m.cT,
                 copies ? m.Fw2 : m.Fw, copies ? m.Tw2 : m.Tw, m.cCnt, m.cCnt2, i);
        const double B = m.check * m.bud[i];
        const D3 d0 = ld3(in.c, i) - D3{m.c0[i], m.c0[(size_t)m.cap + i], m.c0[2 * (size_t)m.cap + i]};
        if (magSqr3(d0) > B * B) atomicExch(m.violated, 1);
      }
    }
    __syncthreads();  // every centre final
    if (copies) {
      const int nMem = sh.nMem;
      for (int t = tid; t < nMem; t += K5C_THREADS) {
        const int i = sh.mem[t];
        if (m.idxHead[in.index[i]] == i) d5_set_positions(ml, in, in.index[i]);  // the first instance of a group does the group
      }
      __syncthreads();
    }
    const bool grew = sh.nMem != nMemBefore;  // (block-uniform)
    if (grew) for (auto& ca : cache) { ca.validAcc = 0; ca.dtW = -1.0; }
    for (int t = tid; t < sh.nMem; t += K5C_THREADS) { k5c_reload_state(sh, t, m, in); k5c_link_ring(sh, t); if (grew) k5c_build_nbrs(sh, t); }
    __syncthreads();
    tGen += clock64() - t1;
    if (sh.fail) break;  // (block-uniform: written before the last barrier)
  }
  atomicMax(&m.clusterStat[9], (unsigned long long)tPair); atomicMax(&m.clusterStat[10], (unsigned long long)tEvOwn);
  if (tid == 0) {  // diagnostics: the slowest block's cycles per phase
    atomicMax(&m.clusterStat[2], (unsigned long long)tEv); atomicMax(&m.clusterStat[3], (unsigned long long)tFl);
    atomicMax(&m.clusterStat[4], (unsigned long long)tSp); atomicMax(&m.clusterStat[5], (unsigned long long)tGen);
    atomicMax(&m.clusterStat[6], (unsigned long long)(clock64() - tStart)); atomicMax(&m.clusterStat[7], (unsigned long long)sh.nMem);
    atomicMax(&m.clusterStat[8], (unsigned long long)(tStart - tLaunch));
  }
  for (int t = tid; t < sh.nMem; t += K5C_THREADS) {
    k5c_store_state(sh, t, in);
    if (m.owned) m.mine[sh.mem[t]] = 1;
  }
  if (tid == 0 && sh.fail) atomicExch(m.clusterFail, 1);
}

// ------------------------------------------------------------------------------------------------
// Owned mode: the exchange of a fluid step.  One buffer of 64-bit words, summed over the ranks as integers (every word is
// written by at most one rank with a value, by the others with 0, so bit patterns travel unchanged):
//   header [32]            [0] new copies in total, [1] clusterFail, [2] violated, [3] overflow, [8 + r] new copies of rank r
//   touched [n0]           1 where the rank advanced the instance
//   state [13][n0]         centre, v, omega, theta, active of the instances it advanced
//   records [nRanks][Q]    the periodic copies it created (provisional slots n0, n0 + 1, ... of that rank), K5X_REC words each
// ------------------------------------------------------------------------------------------------
#define K5X_HEAD 32
#define K5X_REC 32
#define K5X_Q 256
__device__ __forceinline__ long long dbits(double x) { return __double_as_longlong(x); }
__device__ __forceinline__ double bitsd(long long x) { return __longlong_as_double(x); }

__global__ void k5x_pack(MotionDev m, InstanceSoA in, int n0, long long* __restrict__ X) {
  const size_t N0 = (size_t)n0;
  long long* T = X + K5X_HEAD;
  long long* S = T + N0;
  long long* R = S + 13 * N0;
  const int cnt = *m.nextSlot - n0;
  const int nRec = m.nRanks * K5X_Q;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n0 + nRec + K5X_HEAD; t += gridDim.x * blockDim.x) {
    if (t < n0) {
      const int i = t;
      const bool me = m.mine[i] != 0;
      T[i] = me ? 1 : 0;
      for (int q = 0; q < 3; ++q) {
        S[(size_t)q * N0 + i] = me ? dbits(in.c[q][i]) : 0;
        S[(size_t)(3 + q) * N0 + i] = me ? dbits(in.v[q][i]) : 0;
        S[(size_t)(6 + q) * N0 + i] = me ? dbits(in.om[q][i]) : 0;
        S[(size_t)(9 + q) * N0 + i] = me ? dbits(in.th[q][i]) : 0;
      }
      S[(size_t)12 * N0 + i] = me ? (long long)in.active[i] : 0;
    } else if (t < n0 + nRec) {
      const int rq = t - n0, r = rq / K5X_Q, q = rq - r * K5X_Q;
      long long* rec = R + (size_t)rq * K5X_REC;
      const bool me = r == m.rank && q < cnt;
      if (!me) {
        for (int w = 0; w < K5X_REC; ++w) rec[w] = 0;
      } else {
        const int s = n0 + q;
        rec[0] = in.index[s]; rec[1] = in.copy[s]; rec[2] = in.active[s]; rec[3] = in.onMesh[s]; rec[4] = in.cls[s];
        rec[5] = m.evSub[s]; rec[6] = m.evKey[s];
        for (int a = 0; a < 3; ++a) {
          rec[7 + a] = dbits(in.c[a][s]); rec[10 + a] = dbits(in.th[a][s]); rec[13 + a] = dbits(in.v[a][s]); rec[16 + a] = dbits(in.om[a][s]);
        }
        rec[19] = dbits(in.rho[s]); rec[20] = dbits(in.age[s]); rec[21] = dbits(m.bud[s]);
        for (int a = 0; a < 9; ++a) rec[22 + a] = dbits(in.rot[9 * (size_t)s + a]);
        rec[31] = 1;
      }
    } else {
      const int h = t - n0 - nRec;
      long long v = 0;
      if (h == 0) v = cnt > 0 ? cnt : 0;
      else if (h == 1) v = (*m.clusterFail != 0 || cnt > K5X_Q || cnt < 0) ? 1 : 0;
      else if (h == 2) v = *m.violated != 0;
      else if (h == 3) v = *m.overflow != 0;
      else if (h == 8 + m.rank) v = cnt > 0 ? cnt : 0;
      X[h] = v;
    }
  }
}

__global__ void k5x_unpack(MotionDev m, InstanceSoA in, int n0, const long long* __restrict__ X) {
  const size_t N0 = (size_t)n0;
  const long long* T = X + K5X_HEAD;
  const long long* S = T + N0;
  const long long* R = S + 13 * N0;
  const int nRec = m.nRanks * K5X_Q;
  const long long total = X[0];
  const bool bad = X[1] != 0 || (long long)n0 + total > (long long)m.cap;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n0 + nRec + 1; t += gridDim.x * blockDim.x) {
    if (t < n0) {
      const int i = t;
      m.mine[i] = 0;  // (ready for the next step)
      if (T[i] == 1) {
        for (int q = 0; q < 3; ++q) {
          in.c[q][i] = bitsd(S[(size_t)q * N0 + i]);
          in.v[q][i] = bitsd(S[(size_t)(3 + q) * N0 + i]);
          in.om[q][i] = bitsd(S[(size_t)(6 + q) * N0 + i]);
          in.th[q][i] = bitsd(S[(size_t)(9 + q) * N0 + i]);
        }
        in.active[i] = (int)S[(size_t)12 * N0 + i];
      } else if (T[i] != 0) {
        atomicExch(m.clusterFail, 1);  // two ranks advanced the same instance: the ownership rule was not the same everywhere
      }
    } else if (t < n0 + nRec) {
      if (bad) continue;
      const int rq = t - n0, r = rq / K5X_Q, q = rq - r * K5X_Q;
      if (q >= X[8 + r]) continue;
      int off = 0;
      for (int r2 = 0; r2 < r; ++r2) off += (int)X[8 + r2];
      const long long* rec = R + (size_t)rq * K5X_REC;
      const int dst = n0 + off + q;
      const int idx = (int)rec[0];
      in.index[dst] = idx; in.copy[dst] = (int)rec[1]; in.active[dst] = (int)rec[2]; in.onMesh[dst] = (int)rec[3]; in.cls[dst] = (int)rec[4];
      m.evSub[dst] = (int)rec[5];
      {
        // the source of a copy may itself be a copy created earlier in the step: its provisional slot moves with its rank's offset
        const int key = (int)rec[6], patch = key / m.cap;
        int src = key - patch * m.cap;
        if (src >= n0) src += off;
        m.evKey[dst] = patch * m.cap + src;
      }
      for (int a = 0; a < 3; ++a) {
        in.c[a][dst] = bitsd(rec[7 + a]); in.th[a][dst] = bitsd(rec[10 + a]); in.v[a][dst] = bitsd(rec[13 + a]); in.om[a][dst] = bitsd(rec[16 + a]);
        in.F[a][dst] = 0.0; in.T[a][dst] = 0.0;
        m.c0[(size_t)a * m.cap + dst] = bitsd(rec[7 + a]);
      }
      in.rho[dst] = bitsd(rec[19]); in.age[dst] = bitsd(rec[20]); m.bud[dst] = bitsd(rec[21]);
      for (int a = 0; a < 9; ++a) in.rot[9 * (size_t)dst + a] = bitsd(rec[22 + a]);
      in.nShell[dst] = 0; in.shellOff[dst] = 0; in.nPtsOnMesh[dst] = 0;
      m.quiet[dst] = 0; m.cCnt[dst] = 0; m.cCnt2[dst] = 0; m.mine[dst] = 0;
      for (int w = 0; w < m.nWalls; ++w)
        for (int a = 0; a < 3; ++a) {
          const size_t o = 3 * ((size_t)dst * m.nWalls + w) + a;
          m.Fw[o] = 0.0; m.Tw[o] = 0.0; m.Fw2[o] = 0.0; m.Tw2[o] = 0.0;
        }
      atomicMax(&m.maxCopies[idx], (int)rec[1]);  // (the copies of one index all come from one block, hence from one rank)
    } else {
      *m.nextSlot = bad ? n0 : n0 + (int)total;
      if (bad) *m.clusterFail = 1;
      if (X[2] != 0) *m.violated = 1;
      if (X[3] != 0) *m.overflow = 1;
    }
  }
}

// new instances [n0, n0 + total) into the reference's append order: dst slot n0 + r takes provisional slot n0 + perm[r]
__global__ void k5c_permute(InstanceSoA s, InstanceSoA d, const int* __restrict__ perm, int n0, int total, int back) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total) return;
  const int i = back ? n0 + r : n0 + perm[r], j = n0 + r;
  d.index[j] = s.index[i]; d.copy[j] = s.copy[i]; d.active[j] = s.active[i]; d.onMesh[j] = s.onMesh[i]; d.cls[j] = s.cls[i];
  d.nShell[j] = s.nShell[i]; d.shellOff[j] = s.shellOff[i]; d.nPtsOnMesh[j] = s.nPtsOnMesh[i]; d.onLocal[j] = s.onLocal[i];
  d.rho[j] = s.rho[i]; d.age[j] = s.age[i];
  for (int q = 0; q < 3; ++q) {
    d.c[q][j] = s.c[q][i]; d.th[q][j] = s.th[q][i]; d.v[q][j] = s.v[q][i]; d.om[q][j] = s.om[q][i];
    d.F[q][j] = s.F[q][i]; d.T[q][j] = s.T[q][i];
  }
  for (int q = 0; q < 9; ++q) d.rot[9 * (size_t)j + q] = s.rot[9 * (size_t)i + q];
}

// ------------------------------------------------------------------------------------------------
// The whole sub-step loop of particleCloud::solve (particleCloud.C:1251-1275) as ONE persistent cooperative
// kernel: phases are separated by grid-wide barriers instead of kernel launches (the unfused sequence above
// costs ~8 launches x 100 sub-steps per fluid step and was launch-bound).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k5_fused(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ dts, int nSub,
         int doCyclic) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int nBins = m.nbx * m.nby * m.nbz;
  const bool copies = m.nCyc > 0;
  for (int s = 0; s < nSub; ++s) {
    const double dt = dts[s];
    if (m.moving) {
      int n = *m.nInst;  // only the link phase changes it, and a grid barrier follows
      for (int i = tid; i < n; i += nth) d5_bin(m, in, i);
      grid.sync();
      for (int t = tid; t < 3 * n; t += nth) d5_find_pairs(m, in, t);
      for (int i = tid; i < n; i += nth) {
        if (m.nWalls > 0) d5_walls(m, L, in, classes, dt, i);
        if (copies && doCyclic) d5_cyclic_candidates(m, in, classes, i);
      }
      grid.sync();
      {
        const int np = min(*m.pairCount, m.pairCap);
        for (int h = tid; h < np; h += nth) d5_pair_force(m, in, classes, dt, h);
      }
      grid.sync();
      for (int i = tid; i < n; i += nth) d5_rows_finish(m, in, i);
      if (tid == 0) *m.pairCount = 0;
      grid.sync();
      if (copies) {
        for (int i = tid; i < n; i += nth) d5_merge(m, in, i);
        if (doCyclic) d5_cyclic_detect(m, L, in, classes, tid, nth);
        grid.sync();
        if (doCyclic) {
          const int total = *m.hitCount;  // complete after the barrier, identical for every thread
          if (total > 0) {
            d5_append(m, in, tid, nth);
            grid.sync();
            if (tid == 0) d5_append_link(m, in);
            grid.sync();
          } else if (tid == 0) {
            *m.candCount = 0;
          }
        }
      }
    }
    {
      const int n = *m.nInst;
      const int* cCnt = copies ? m.cCnt2 : m.cCnt;
      for (int i = tid; i < n; i += nth)
        d5_solve(m, L, in, classes, dt, cCnt, copies ? m.cIdx2 : m.cIdx, copies ? m.cF2 : m.cF, copies ? m.cT2 : m.cT,
                 copies ? m.Fw2 : m.Fw, copies ? m.Tw2 : m.Tw, m.cCnt, m.cCnt2, i);
      for (int b = tid; b < nBins; b += nth) m.binCnt[b] = 0;  // contacts are done: ready for the next sub-step
    }
    grid.sync();
    if (copies) {
      for (int idx = tid; idx < m.nParticles; idx += nth) d5_set_positions(m, in, idx);
      grid.sync();
    }
  }
}

// The sub-step loop of the ACTIVE instances of a split step (k5q_* above) as one persistent cooperative kernel: the active
// list is a fraction of the cloud, so a small grid holds it and its grid-wide barriers are cheap next to the ~9 launches
// per sub-step of the unfused chain.  Phases are those of substep_launch_active, except that every instance builds its own
// contact row in one go (d5_row_direct) instead of the pair list + pair kernel + row finalisation, and the bins of a sub-step
// are filled in the last phase of the previous one: 3 barriers per sub-step without periodic patches, 4 with them.
__global__ void __launch_bounds__(256)
k5a_fused(MotionDev m, LatticeDev L, InstanceSoA in, const ShapeClassDev* __restrict__ classes, const double* __restrict__ dts, int nSub,
          int doCyclic) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const bool copies = m.nCyc > 0;
  // Barriers per sub-step: rows | [merge + patch scan |] integrate | copy states + bins of the NEXT sub-step.  The bins only
  // need the centres, which are final after the integration, so they share a phase with setPositions (which writes angles
  // and velocities of copies only); the periodic-patch candidates need those velocities and go with the rows.
  if (m.moving && nSub > 0) {
    const int na = *m.nAct;
    for (int t = tid; t < na; t += nth) d5_bin(m, in, m.act[t]);
  }
  grid.sync();
  for (int s = 0; s < nSub; ++s) {
    const double dt = dts[s];
    if (m.moving) {
      const int na = *m.nAct;  // only the link phase changes it, and a grid barrier follows
      for (int t = tid; t < na; t += nth) {  // every instance builds and finishes its own contact row (no pair list, no atomics)
        const int i = m.act[t];
        d5_row_direct(m, in, classes, dt, i);
        if (m.nWalls > 0) d5_walls(m, L, in, classes, dt, i);
        if (doCyclic) d5_cyclic_candidates(m, in, classes, i);
      }
      grid.sync();  // rows complete before anybody merges them or moves a centre
      if (copies) {
        for (int t = tid; t < na; t += nth) d5_merge(m, in, m.act[t]);
        if (doCyclic) d5_cyclic_detect(m, L, in, classes, tid, nth);
        grid.sync();  // merged rows read the raw rows of other instances, which the integration clears
        if (doCyclic) {
          const int total = *m.hitCount;  // complete after the barrier, identical for every thread
          if (total > 0) {
            d5_append(m, in, tid, nth);
            grid.sync();
            if (tid == 0) d5_append_link(m, in);
            grid.sync();
          } else if (tid == 0) {
            *m.candCount = 0;  // (the next candidates phase is two barriers away)
          }
        }
      }
    }
    {
      const int na = *m.nAct;
      for (int t = tid; t < na; t += nth) {
        const int i = m.act[t];
        if (i >= *m.nInst) continue;
        if (in.onMesh[i] != 0) {  // the bin this instance was put in (its centre has not changed since)
          int bx, by, bz;
          bin_of(m, ld3(in.c, i), bx, by, bz);
          m.binCnt[bx + m.nbx * (by + m.nby * bz)] = 0;
        }
        d5_solve(m, L, in, classes, dt, copies ? m.cCnt2 : m.cCnt, copies ? m.cIdx2 : m.cIdx, copies ? m.cF2 : m.cF, copies ? m.cT2 : m.cT,
                 copies ? m.Fw2 : m.Fw, copies ? m.Tw2 : m.Tw, m.cCnt, m.cCnt2, i);
        const double B = m.check * m.bud[i];
        const D3 d0 = ld3(in.c, i) - D3{m.c0[i], m.c0[(size_t)m.cap + i], m.c0[2 * (size_t)m.cap + i]};
        if (magSqr3(d0) > B * B) atomicExch(m.violated, 1);
      }
    }
    grid.sync();  // every bin is empty and every centre final
    {
      const int na = *m.nAct;
      for (int t = tid; t < na; t += nth) {
        const int i = m.act[t];
        if (i >= *m.nInst) continue;
        if (copies && m.idxHead[in.index[i]] == i) d5_set_positions(m, in, in.index[i]);  // the first instance of a group does the group
        if (m.moving && s + 1 < nSub) d5_bin(m, in, i);                                    // bins of the next sub-step
      }
    }
    grid.sync();
  }
}

// the six F / T components of every instance in one buffer [6][n] and back: ONE all-reduce instead of six
__global__ void k_pack_ft(InstanceSoA in, int n, double* __restrict__ buf, int unpack) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (unpack) { in.F[q][i] = buf[(size_t)q * n + i]; in.T[q][i] = buf[(size_t)(3 + q) * n + i]; }
    else { buf[(size_t)q * n + i] = in.F[q][i]; buf[(size_t)(3 + q) * n + i] = in.T[q][i]; }
  }
}

// particleCloud::combineForces, serial branch (particleCloud.C:268-312)
__global__ void k_combine_forces(MotionDev m, InstanceSoA in) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m.nParticles) return;
  int head = m.idxHead[idx];
  if (head < 0) return;
  D3 totF{0, 0, 0}, totT{0, 0, 0};
  for (int p = head; p >= 0; p = m.instNext[p]) { totF = totF + ld3(in.F, p); totT = totT + ld3(in.T, p); }
  for (int p = head; p >= 0; p = m.instNext[p]) { st3(in.F, p, totF); st3(in.T, p, totT); }
}

// same-index lists from a stable sort of (index, position)
__global__ void k_group_keys(InstanceSoA in, int n, int cap, int* keys, int* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  keys[i] = i < n ? in.index[i] : 0x7fffffff;
  vals[i] = i;
}
__global__ void k_group_link(const int* __restrict__ keys, const int* __restrict__ vals, int n, MotionDev m) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int k = keys[q], pos = vals[q];
  bool first = (q == 0) || keys[q - 1] != k;
  bool last = (q == n - 1) || keys[q + 1] != k;
  if (first) m.idxHead[k] = pos;
  if (last) m.idxTail[k] = pos;
  m.instNext[pos] = last ? -1 : vals[q + 1];
}

// end-of-solve log block (particleCloud.C:1296-1308): fixed-shape tree => deterministic.  Two levels: every block reduces a
// fixed contiguous segment of the instances, one block then adds the segment sums in order.
#define STATS_BLOCKS 64
__global__ void __launch_bounds__(256) k_stats_partial(MotionDev m, InstanceSoA in, const ShapeClassDev* __restrict__ classes, int n, double* part) {
  __shared__ double red[8][10];
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int seg = (n + STATS_BLOCKS - 1) / STATS_BLOCKS;
  const int i0 = blockIdx.x * seg, i1 = min(i0 + seg, n);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    double act = in.active[i] ? 1.0 : 0.0;
    D3 v = ld3(in.v, i), om = ld3(in.om, i);
    double mass = classes[in.cls[i]].V * in.rho[i];
    acc[0] += om.x * act; acc[1] += om.y * act; acc[2] += om.z * act;
    acc[3] += v.x * act; acc[4] += v.y * act; acc[5] += v.z * act;
    acc[6] += v.x * mass * act; acc[7] += v.y * mass * act; acc[8] += v.z * mass * act;
    acc[9] += 0.5 * mass * dot3(v, v) * act;
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < 10; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    part[blockIdx.x * 10 + threadIdx.x] = v;
  }
}
__global__ void k_stats_final(const double* __restrict__ part, double* out) {
  if (threadIdx.x < 10) {
    double v = 0;
    for (int b = 0; b < STATS_BLOCKS; ++b) v += part[b * 10 + threadIdx.x];
    out[threadIdx.x] = v;
  }
}

__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_copy_int(int* dst, const int* src) { *dst = *src; }

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <class T>
static bool dalloc(ibm_ctx* c, Motion* M, T** p, size_t n) {
  if (!ctx_cuda_ok(c, cudaMalloc(p, std::max<size_t>(n, 1) * sizeof(T)), "cudaMalloc(motion)")) return false;
  cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
  M->allocs.push_back(*p);
  return true;
}

static int motion_init(ibm_ctx* c) {
  if (c->motion) return IBM_OK;
  if (c->cap == 0) return ctx_fail(c, IBM_E_STATE, "loop 5: particles not set");
  Motion* M = new Motion;
  c->motion = M;
  MotionDev& d = M->d;
  d.cap = c->cap; d.nParticles = c->nParticles;
  d.nWalls = (int)c->walls.size(); d.nCyc = (int)c->cyclics.size();
  d.gx = c->cfg.g[0]; d.gy = c->cfg.g[1]; d.gz = c->cfg.g[2]; d.e = c->cfg.e;
  d.moving = c->cfg.moving; d.rotating = c->cfg.rotating;
  // cell list geometry: bin edge = largest cmptMax(D) over the shape classes
  double dmax = 0;
  for (auto& h : c->classes) dmax = std::max(dmax, h.dev.kind == 2 ? std::max(h.dev.a, std::max(h.dev.b, h.dev.l)) : h.dev.d);
  if (!(dmax > 0)) dmax = 1.0;
  const LatticeDev& L = c->Lg;
  d.binEdge = dmax;
  d.bx0 = L.ox - dmax; d.by0 = L.oy - dmax; d.bz0 = L.oz - dmax;
  auto nb = [&](double len) { return std::max(1, (int)std::ceil((len + 2 * dmax) / dmax)); };
  d.nbx = nb(L.nx * L.dx); d.nby = nb(L.ny * L.dy); d.nbz = c->nGeometricD == 2 ? 1 : nb(L.nz * L.dz);
  if (c->nGeometricD == 2) d.bz0 = L.oz - 1e300;
  while ((double)d.nbx * d.nby * d.nbz > 4.0e9) { d.binEdge *= 2; d.nbx = (d.nbx + 1) / 2; d.nby = (d.nby + 1) / 2; d.nbz = (d.nbz + 1) / 2; }
  {
    // capacity per bin: bodies of the smallest class that fit (with slack for overlap) in a bin of the largest
    double dmin = dmax;
    for (auto& h : c->classes) dmin = std::min(dmin, h.dev.kind == 2 ? std::min(h.dev.a, h.dev.b) : h.dev.d);
    double ratio = d.binEdge / std::max(dmin, 1e-300);
    double est = (c->nGeometricD == 2 ? 4.0 * ratio * ratio : 8.0 * ratio * ratio * ratio) + 4.0;
    d.binCap = (int)std::min(64.0, std::max(12.0, std::ceil(est)));
  }
  const size_t cap = (size_t)c->cap;
  d.candCap = (int)(3 * cap);
  d.pairCap = (int)(16 * cap);
  bool ok = dalloc(c, M, &d.nInst, 1) && dalloc(c, M, &d.keys, cap) && dalloc(c, M, &d.keysSorted, cap) && dalloc(c, M, &d.vals, cap) &&
            dalloc(c, M, &d.valsSorted, cap) && dalloc(c, M, &d.cCnt, cap) && dalloc(c, M, &d.cIdx, cap * CONTACT_CAP) &&
            dalloc(c, M, &d.cF, 3 * cap * CONTACT_CAP) && dalloc(c, M, &d.cT, 3 * cap * CONTACT_CAP) && dalloc(c, M, &d.cCnt2, cap) &&
            dalloc(c, M, &d.cIdx2, cap * CONTACT_CAP) && dalloc(c, M, &d.cF2, 3 * cap * CONTACT_CAP) &&
            dalloc(c, M, &d.cT2, 3 * cap * CONTACT_CAP) && dalloc(c, M, &d.overflow, 1) && dalloc(c, M, &d.Fw, 3 * cap * std::max(1, d.nWalls)) &&
            dalloc(c, M, &d.Tw, 3 * cap * std::max(1, d.nWalls)) && dalloc(c, M, &d.Fw2, 3 * cap * std::max(1, d.nWalls)) &&
            dalloc(c, M, &d.Tw2, 3 * cap * std::max(1, d.nWalls)) && dalloc(c, M, &d.idxHead, (size_t)std::max(1, c->nParticles)) &&
            dalloc(c, M, &d.idxTail, (size_t)std::max(1, c->nParticles)) && dalloc(c, M, &d.instNext, cap) &&
            dalloc(c, M, &d.maxCopies, (size_t)std::max(1, c->nParticles)) && dalloc(c, M, &d.hitCount, 1) && dalloc(c, M, &d.blocksDone, 1) && dalloc(c, M, &d.candCount, 1) &&
            dalloc(c, M, &d.candKey, 3 * cap) &&
            dalloc(c, M, &d.hitKey, cap) && dalloc(c, M, &d.hitCenter, 3 * cap) &&
            dalloc(c, M, &d.binCnt, (size_t)d.nbx * d.nby * d.nbz) && dalloc(c, M, &d.binItems, (size_t)d.nbx * d.nby * d.nbz * d.binCap) &&
            dalloc(c, M, &d.binPos, (size_t)3 * d.nbx * d.nby * d.nbz * d.binCap) && dalloc(c, M, &d.pairCount, 1) &&
            dalloc(c, M, &d.pairI, (size_t)d.pairCap) && dalloc(c, M, &d.pairJ, (size_t)d.pairCap) && dalloc(c, M, &d.cPos, cap * CONTACT_CAP) &&
            dalloc(c, M, &M->dIdxKeys, cap) && dalloc(c, M, &M->dIdxKeys2, cap) && dalloc(c, M, &M->dStats, 16 + 10 * STATS_BLOCKS) &&
            dalloc(c, M, &d.act, cap) && dalloc(c, M, &d.nAct, 1) && dalloc(c, M, &d.quiet, cap) && dalloc(c, M, &d.c0, 3 * cap) &&
            dalloc(c, M, &d.vaBits, 2) && dalloc(c, M, &d.budget, 1) && dalloc(c, M, &d.violated, 1) &&
            dalloc(c, M, &d.sOwn, cap) && dalloc(c, M, &d.uMax, cap) && dalloc(c, M, &d.bud, cap) && dalloc(c, M, &d.noSplit, 1) && dalloc(c, M, &M->snap, 12 * cap) &&
            dalloc(c, M, &M->snapActive, cap) && dalloc(c, M, &M->snapMaxCopies, (size_t)std::max(1, c->nParticles));
  if (!ok) return IBM_E_CUDA;
  // cluster mode of the split
  d.edgeCap = (int)std::min<size_t>(8 * cap + 1024, 0x3fffffff);
  d.ctaBins = (int)(cap / K5C_FILL) + 2;
  ok = dalloc(c, M, &d.edgeA, (size_t)d.edgeCap) && dalloc(c, M, &d.edgeB, (size_t)d.edgeCap) && dalloc(c, M, &d.nEdge, 1) &&
       dalloc(c, M, &d.label, cap) && dalloc(c, M, &d.compCnt, cap) && dalloc(c, M, &d.compOff, cap + 1) && dalloc(c, M, &d.compCur, cap) &&
       dalloc(c, M, &d.sorted, cap) && dalloc(c, M, &d.ctaBegin, (size_t)d.ctaBins) && dalloc(c, M, &d.ctaEnd, (size_t)d.ctaBins) &&
       dalloc(c, M, &d.nextSlot, 1) && dalloc(c, M, &d.evSub, cap) && dalloc(c, M, &d.evKey, cap) && dalloc(c, M, &d.clusterFail, 1) && dalloc(c, M, &d.clusterStat, 16) &&
       dalloc(c, M, &M->dPerm, std::max<size_t>(cap, 32)) && dalloc(c, M, &M->dFtPack, 6 * cap) && dalloc(c, M, &d.mine, cap);
  if (!ok) return IBM_E_CUDA;
  d.owned = 0; d.rank = 0; d.nRanks = 1;
  {
    size_t sb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sb, d.compCnt, d.compOff, c->cap);
    M->scanTempBytes = sb + 256;
    if (!ctx_cuda_ok(c, cudaMalloc(&M->scanTemp, M->scanTempBytes), "cudaMalloc(scan)")) return IBM_E_CUDA;
    M->allocs.push_back(M->scanTemp);
    if (!ctx_cuda_ok(c, cudaMallocHost(&M->hEv, (2 * K5C_EVN + 16) * sizeof(int)), "cudaMallocHost(events)")) return IBM_E_CUDA;
  }
  d.activeMode = 0;
  d.safety = 2.0;
  d.check = 1.0;
  // originals
  int* oc = nullptr; double* orho = nullptr; int* wp = nullptr; int* cp = nullptr;
  ok = dalloc(c, M, &oc, (size_t)std::max(1, c->nParticles)) && dalloc(c, M, &orho, (size_t)std::max(1, c->nParticles)) &&
       dalloc(c, M, &wp, (size_t)std::max(1, d.nWalls)) && dalloc(c, M, &cp, (size_t)std::max(1, d.nCyc));
  if (!ok) return IBM_E_CUDA;
  std::vector<double> rho(c->nParticles);
  for (int i = 0; i < c->nParticles; ++i) rho[i] = c->orig[i].rho;
  if (c->nParticles) {
    cudaMemcpy(oc, c->origCls.data(), c->nParticles * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(orho, rho.data(), c->nParticles * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d.maxCopies, c->maxCopies.data(), c->nParticles * sizeof(int), cudaMemcpyHostToDevice);
  }
  if (d.nWalls) cudaMemcpy(wp, c->walls.data(), d.nWalls * sizeof(int), cudaMemcpyHostToDevice);
  if (d.nCyc) cudaMemcpy(cp, c->cyclics.data(), d.nCyc * sizeof(int), cudaMemcpyHostToDevice);
  d.origCls = oc; d.origRho = orho; d.wallPatch = wp; d.cycPatch = cp;
  // patches
  std::vector<PatchDev> pd;
  for (auto& ph : c->patches) {
    PatchDev q{};
    q.type = ph.type; q.side = ph.side; q.neighb = ph.neighb; q.nFaces = ph.nFaces();
    q.Cf = ph.dCf; q.Sf = ph.dSf; q.magSf = ph.dMagSf;
    const int axis = ph.side / 2;
    q.nA = (axis == 0) ? L.ny : L.nx;
    q.nB = (axis == 2) ? L.ny : L.nz;
    // lattice-ordered iff the face centres match the mini-mesh order (then candidate faces come from an index box)
    for (int a = 0; a < 3; ++a) { q.lo[a] = 1e300; q.hi[a] = -1e300; }
    for (int f = 0; f < q.nFaces; ++f)
      for (int a = 0; a < 3; ++a) { q.lo[a] = std::min(q.lo[a], ph.Cf[3 * (size_t)f + a]); q.hi[a] = std::max(q.hi[a], ph.Cf[3 * (size_t)f + a]); }
    q.planar = q.nFaces > 0 ? 1 : 0;
    for (int f = 1; f < q.nFaces && q.planar; ++f)
      if (ph.Sf[3 * (size_t)f] != ph.Sf[0] || ph.Sf[3 * (size_t)f + 1] != ph.Sf[1] || ph.Sf[3 * (size_t)f + 2] != ph.Sf[2] ||
          ph.magSf[f] != ph.magSf[0]) q.planar = 0;
    for (int a = 0; a < 3; ++a) q.nrm[a] = q.nFaces > 0 ? ph.Sf[a] / ph.magSf[0] : 0.0;
    q.latticeOrdered = 0;
    if (q.nFaces == q.nA * q.nB) {
      const double o3[3] = {L.ox, L.oy, L.oz}, d3[3] = {L.dx, L.dy, L.dz};
      const int ta = (axis == 0) ? 1 : 0, tb = (axis == 2) ? 1 : 2;
      bool match = true;
      for (int b = 0; b < q.nB && match; ++b)
        for (int a = 0; a < q.nA && match; ++a) {
          const double* f = &ph.Cf[3 * ((size_t)a + (size_t)q.nA * b)];
          double ea = o3[ta] + (a + 0.5) * d3[ta], eb = o3[tb] + (b + 0.5) * d3[tb];
          if (std::fabs(f[ta] - ea) > 1e-9 * d3[ta] || std::fabs(f[tb] - eb) > 1e-9 * d3[tb]) match = false;
        }
      q.latticeOrdered = match ? 1 : 0;
    }
    q.uniformShift = 0;
    if (ph.type == IBM_CYCLIC && ph.neighb >= 0 && ph.neighb < (int)c->patches.size() && q.nFaces > 0 &&
        c->patches[ph.neighb].nFaces() == q.nFaces) {
      const auto& nb = c->patches[ph.neighb];
      double scale = 0.0;
      for (int a = 0; a < 3; ++a) { q.shift[a] = nb.Cf[a] - ph.Cf[a]; scale = std::max(scale, std::fabs(q.shift[a])); }
      bool uniform = true;
      for (int f = 1; f < q.nFaces && uniform; ++f)
        for (int a = 0; a < 3; ++a)
          if (std::fabs((nb.Cf[3 * (size_t)f + a] - ph.Cf[3 * (size_t)f + a]) - q.shift[a]) > 1e-9 * scale) uniform = false;
      q.uniformShift = uniform ? 1 : 0;
    }
    if (ph.type == IBM_CYCLIC && !q.uniformShift) M->splitAllowed = false;  // images are not translations: no quiet / active split
    pd.push_back(q);
  }
  if (!pd.empty()) {
    if (!dalloc(c, M, &M->dPatches, pd.size())) return IBM_E_CUDA;
    cudaMemcpy(M->dPatches, pd.data(), pd.size() * sizeof(PatchDev), cudaMemcpyHostToDevice);
  }
  d.patches = M->dPatches;
  // cub temp storage
  size_t b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b2, M->dIdxKeys, M->dIdxKeys2, d.vals, d.valsSorted, c->cap);
  M->cubTempBytes = b2 + 256;
  if (!dalloc(c, M, &M->dDts, 256)) return IBM_E_CUDA;
  {
    int perSm = 0, sms = 0, coop = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k5_fused, 256, 0) == cudaSuccess && perSm > 0)
      M->fusedBlocks = sms * std::min(perSm, 4);
    perSm = 0;
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k5a_fused, 256, 0) == cudaSuccess && perSm > 0)
      M->fusedActBlocks = sms * std::min(perSm, 2);
  }
  if (!ctx_cuda_ok(c, cudaMalloc(&M->cubTemp, M->cubTempBytes), "cudaMalloc(cub)")) return IBM_E_CUDA;
  return IBM_OK;
}

static int blocks_for(int n, int b = 128) { return std::max(1, (n + b - 1) / b); }

// (re)build the same-index lists after the instance array changed on the host side (set_particles, compaction)
static int rebuild_groups(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  cudaStream_t st = c->stream;
  IBM_CUDA(c, cudaMemsetAsync(d.idxHead, 0xff, (size_t)std::max(1, c->nParticles) * sizeof(int), st));
  IBM_CUDA(c, cudaMemsetAsync(d.idxTail, 0xff, (size_t)std::max(1, c->nParticles) * sizeof(int), st));
  k_set_int<<<1, 1, 0, st>>>(d.nInst, c->nInst);
  if (c->nInst > 0) {
    k_group_keys<<<blocks_for(c->cap), 128, 0, st>>>(c->in, c->nInst, c->cap, M->dIdxKeys, d.vals);
    size_t tb = M->cubTempBytes;
    cub::DeviceRadixSort::SortPairs(M->cubTemp, tb, M->dIdxKeys, M->dIdxKeys2, d.vals, d.valsSorted, c->cap, 0, 32, st);
    k_group_link<<<blocks_for(c->nInst), 128, 0, st>>>(M->dIdxKeys2, d.valsSorted, c->nInst, d);
    c->nLaunch += 4;
  }
  M->groupsValid = true;
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}

static int ensure_motion(ibm_ctx* c) {
  int rc = motion_init(c);
  if (rc) return rc;
  Motion* M = (Motion*)c->motion;
  if (!M->groupsValid) return rebuild_groups(c);
  return IBM_OK;
}

void motion_invalidate_groups(ibm_ctx* c) {
  if (c->motion) {
    ((Motion*)c->motion)->groupsValid = false;
    ((Motion*)c->motion)->graphValid = false;  // the instance arrays were exchanged / rewritten on the host side
  }
}

bool motion_counter(ibm_ctx* c, const char* name, long long* value) {
  Motion* M = (Motion*)c->motion;
  if (std::strcmp(name, "loop5_fallbacks") == 0) { *value = M ? M->nFallbacks : 0; return true; }
  if (M && std::strncmp(name, "loop5_stat", 10) == 0) {
    unsigned long long h[16];
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(h, M->d.clusterStat, sizeof(h), cudaMemcpyDeviceToHost);
    const int q = std::atoi(name + 10);
    *value = (q >= 0 && q < 16) ? (long long)h[q] : -1;
    return true;
  }
  if (M && (std::strcmp(name, "loop5_block_substeps") == 0 || std::strcmp(name, "loop5_general_substeps") == 0)) {
    unsigned long long h[2] = {0, 0};
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(h, M->d.clusterStat, sizeof(h), cudaMemcpyDeviceToHost);
    *value = (long long)h[std::strcmp(name, "loop5_general_substeps") == 0 ? 1 : 0];
    return true;
  }
  if (std::strcmp(name, "loop5_cluster_flags") == 0) { *value = M ? M->lastFlags : 0; return true; }
  if (std::strcmp(name, "loop5_cluster_steps") == 0) { *value = M ? M->nClusterSteps : 0; return true; }
  if (std::strcmp(name, "loop5_owned_steps") == 0) { *value = M ? M->nOwnedSteps : 0; return true; }
  if (std::strcmp(name, "loop5_active_instances") == 0) { *value = (M && M->d.activeMode) ? c->hScal[11] : -1; return true; }  // of the last advance
  return false;
}

int combine_forces(ibm_ctx* c) {
  int rc = ensure_motion(c);
  if (rc) return rc;
  Motion* M = (Motion*)c->motion;
  if (comm_ranks(c) > 1 && c->nInst > 0) {
    // cross-rank part of combineForces (particleCloud.C:236-237): every rank holds every instance, so the partials
    // of one instance over the slabs are summed element-wise
    k_pack_ft<<<blocks_for(c->nInst), 128, 0, c->stream>>>(c->in, c->nInst, M->dFtPack, 0);
    if ((rc = comm_allreduce_sum(c, M->dFtPack, (size_t)6 * c->nInst, 0))) return rc;
    k_pack_ft<<<blocks_for(c->nInst), 128, 0, c->stream>>>(c->in, c->nInst, M->dFtPack, 1);
    c->nLaunch += 2;
  }
  if (c->nParticles > 0) {
    k_combine_forces<<<blocks_for(c->nParticles), 128, 0, c->stream>>>(M->d, c->in);
    c->nLaunch++;
  }
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}

static int substep_launch(ibm_ctx* c, double localDt) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  cudaStream_t st = c->stream;
  const int cap = c->cap;
  const bool copiesPossible = d.nCyc > 0;
  const bool doCyclic = copiesPossible && (d.nWalls > 0 || c->cfg.cyclicWithoutWalls);  // App. B #10
  const int* cCnt = d.cCnt; const int* cIdx = d.cIdx; const double* cF = d.cF; const double* cT = d.cT;
  double* Fw = d.Fw; double* Tw = d.Tw;
  if (c->cfg.moving) {
    // computeCollisions (bins are cleared by the previous k5_solve_clear / by ensure_bins_clear)
    k5_bin_cand<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, doCyclic ? 1 : 0);
    k5_find_pairs<<<blocks_for(3 * cap), 128, 0, st>>>(d, c->in);
    k5_pair_force<<<blocks_for(std::min(d.pairCap, 4 * c->nInst + 1024), 64), 64, 0, st>>>(d, c->in, c->dClasses, localDt);
    k5_rows_finish<<<blocks_for(cap), 128, 0, st>>>(d, c->in);
    c->nLaunch += 4;
    // computeWallHits
    if (d.nWalls > 0) { k5_walls<<<blocks_for(cap), 128, 0, st>>>(d, c->Lg, c->in, c->dClasses, localDt); c->nLaunch++; }
    if (copiesPossible) {
      k5_merge<<<blocks_for(cap, 64), 64, 0, st>>>(d, c->in);
      c->nLaunch++;
      cCnt = d.cCnt2; cIdx = d.cIdx2; cF = d.cF2; cT = d.cT2; Fw = d.Fw2; Tw = d.Tw2;
      if (doCyclic) {  // computeCyclicHits
        k5_cyclic_detect<<<148, 256, 0, st>>>(d, c->Lg, c->in, c->dClasses);
        k5_append_all<<<32, 128, 0, st>>>(d, c->in);
        c->nLaunch += 2;
      }
    }
  }
  k5_solve_clear<<<blocks_for(cap), 128, 0, st>>>(d, c->Lg, c->in, c->dClasses, localDt, cCnt, cIdx, cF, cT, Fw, Tw, d.cCnt, d.cCnt2);
  c->nLaunch++;
  if (copiesPossible && c->nParticles > 0) { k5_set_positions<<<blocks_for(c->nParticles), 128, 0, st>>>(d, c->in); c->nLaunch++; }
  return IBM_OK;
}

// the same chain over the compact list of active instances (quiet / active split)
static int substep_launch_active(ibm_ctx* c, double localDt) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  cudaStream_t st = c->stream;
  const int cap = c->cap;
  const bool copiesPossible = d.nCyc > 0;
  const bool doCyclic = copiesPossible && (d.nWalls > 0 || c->cfg.cyclicWithoutWalls);  // App. B #10
  const int* cCnt = d.cCnt; const int* cIdx = d.cIdx; const double* cF = d.cF; const double* cT = d.cT;
  double* Fw = d.Fw; double* Tw = d.Tw;
  if (c->cfg.moving) {
    k5a_bin_cand<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, doCyclic ? 1 : 0);
    k5a_rows_direct<<<blocks_for(cap, 64), 64, 0, st>>>(d, c->Lg, c->in, c->dClasses, localDt);  // (+ wall hits)
    c->nLaunch += 2;
    if (copiesPossible) {
      k5a_merge<<<blocks_for(cap, 64), 64, 0, st>>>(d, c->in);
      c->nLaunch++;
      cCnt = d.cCnt2; cIdx = d.cIdx2; cF = d.cF2; cT = d.cT2; Fw = d.Fw2; Tw = d.Tw2;
      if (doCyclic) {
        k5_cyclic_detect<<<148, 256, 0, st>>>(d, c->Lg, c->in, c->dClasses);
        k5_append_all<<<32, 128, 0, st>>>(d, c->in);
        c->nLaunch += 2;
      }
    }
  }
  k5a_solve_clear<<<blocks_for(cap), 128, 0, st>>>(d, c->Lg, c->in, c->dClasses, localDt, cCnt, cIdx, cF, cT, Fw, Tw, d.cCnt, d.cCnt2);
  c->nLaunch++;
  if (copiesPossible && c->nParticles > 0) { k5a_set_positions<<<blocks_for(cap), 128, 0, st>>>(d, c->in); c->nLaunch++; }
  return IBM_OK;
}

// start of a split step: bounds, bins at the start positions, classification
static void active_classify(ibm_ctx* c, double dtStep, bool clusterMode = false, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  cudaStream_t st = c->stream;
  const int cap = c->cap;
  const bool doCyclic = d.nCyc > 0 && (d.nWalls > 0 || c->cfg.cyclicWithoutWalls);  // App. B #10: copies are created at all
  cudaMemsetAsync(d.vaBits, 0, 2 * sizeof(unsigned long long), st);
  cudaMemsetAsync(d.nAct, 0, sizeof(int), st);
  cudaMemsetAsync(d.violated, 0, sizeof(int), st);
  cudaMemsetAsync(d.noSplit, 0, sizeof(int), st);
  k5q_bounds<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, dtStep);  // (reads the hydrodynamic force: before anything rewrites it)
  if (side) {  // everything below only needs positions, velocities and the bounds: it may run beside the caller's next kernels
    cudaEventRecord(fork, st);
    cudaStreamWaitEvent(side, fork, 0);
    st = side;
  }
  k5_bin<<<blocks_for(cap), 128, 0, st>>>(d, c->in);
  if (doCyclic) { k5q_ghosts<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses); c->nLaunch++; }
  // the schedule's Morton order is a permutation of the instances the last update() saw (c->schedN of them)
  k5q_nbrmax<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, d.sOwn, d.uMax, 1.0, 0.0, doCyclic ? 1 : 0, c->sched.order, c->schedN);
  k5q_nbrmax<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, d.uMax, d.bud, d.safety, 1e-12 * d.binEdge, doCyclic ? 1 : 0, c->sched.order, c->schedN);
  if (c->nParticles > 0) { k5q_group_budget<<<blocks_for(c->nParticles), 128, 0, st>>>(d); c->nLaunch++; }
  if (!clusterMode) k5q_classify<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, c->sched.order, c->schedN);
  c->nLaunch += 5;
}

// ... then every sub-step of the quiet instances (the per-sub-step chain of the active ones follows)
static void active_prologue(ibm_ctx* c, double dtStep, int nSub) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  cudaStream_t st = c->stream;
  active_classify(c, dtStep);
  cudaMemsetAsync(d.binCnt, 0, (size_t)d.nbx * d.nby * d.nbz * sizeof(int), st);
  k5q_quiet<<<blocks_for(c->cap), 128, 0, st>>>(d, c->Lg, c->in, c->dClasses, M->dDts, nSub);
  c->nLaunch += 1;
}

// cluster mode: classification, components, packing, quiet instances, then ONE launch for every sub-step of the active ones
// cluster mode, part 1: classification, components, packing.  With a side stream (motion_prepare_async) everything after
// the displacement bounds runs beside the caller's next kernels - the surface-stress integral of the same solve() - and
// `join` is recorded behind it.
static void cluster_prepare(ibm_ctx* c, double dtStep, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  const int cap = c->cap;
  cudaMemsetAsync(d.nEdge, 0, sizeof(int), c->stream);
  cudaMemsetAsync(d.clusterFail, 0, sizeof(int), c->stream);
  k_copy_int<<<1, 1, 0, c->stream>>>(d.nextSlot, d.nInst);
  active_classify(c, dtStep, true, side, fork);  // budgets; the decision itself is k5c_classify below (it also emits the edges)
  cudaStream_t st = side ? side : c->stream;
  cudaMemsetAsync(d.compCnt, 0, (size_t)cap * sizeof(int), st);
  cudaMemsetAsync(d.compCur, 0, (size_t)cap * sizeof(int), st);
  cudaMemsetAsync(d.ctaBegin, 0x7f, (size_t)d.ctaBins * sizeof(int), st);
  cudaMemsetAsync(d.ctaEnd, 0, (size_t)d.ctaBins * sizeof(int), st);
  k5c_classify<<<blocks_for(cap), 128, 0, st>>>(d, c->in, c->dClasses, c->sched.order, c->schedN);
  if (c->clusterOneBlockCC) {
    k5c_components<<<1, 1024, 0, st>>>(d);
  } else {
    k5c_cc_guard<<<148, 256, 0, st>>>(d);
    for (int round = 0; round < 12; ++round) {
      k5c_cc_hook<<<148, 256, 0, st>>>(d, 0);
      k5c_cc_compress<<<blocks_for(cap), 128, 0, st>>>(d);
    }
    k5c_cc_hook<<<148, 256, 0, st>>>(d, 1);  // not converged after 12 rounds (a chain of thousands of bodies): the step is redone unsplit
    c->nLaunch += 25;
  }
  k5c_count<<<blocks_for(cap), 128, 0, st>>>(d);
  size_t tb = M->scanTempBytes;
  cub::DeviceScan::ExclusiveSum(M->scanTemp, tb, d.compCnt, d.compOff, cap, st);
  k5c_place<<<blocks_for(cap), 128, 0, st>>>(d);
  cudaMemsetAsync(d.binCnt, 0, (size_t)d.nbx * d.nby * d.nbz * sizeof(int), st);
  c->nLaunch += 7;
  if (side) cudaEventRecord(join, side);
}

// cluster mode, part 2: every sub-step of the quiet instances in one kernel, every sub-step of the active ones in another
static void cluster_execute(ibm_ctx* c, int nSub) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  const int doCyclic = (d.nCyc > 0 && (d.nWalls > 0 || c->cfg.cyclicWithoutWalls)) ? 1 : 0;
  k5c_quiet<<<blocks_for(c->cap), 128, 0, c->stream>>>(d, c->Lg, c->in, c->dClasses, M->dDts, nSub);
  k5c_run<<<d.ctaBins, K5C_THREADS, 0, c->stream>>>(d, c->Lg, c->in, c->dClasses, M->dDts, nSub, doCyclic, c->clusterGeneralOnly ? 1 : 0);
  c->nLaunch += 2;
}

static int finish_substeps(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  // instance count and overflow flags back to the host
  IBM_CUDA(c, cudaMemcpyAsync(c->hScal + 8, M->d.nInst, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  IBM_CUDA(c, cudaMemcpyAsync(c->hScal + 9, M->d.overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  IBM_CUDA(c, cudaMemcpyAsync(c->hScal + 10, M->d.violated, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  IBM_CUDA(c, cudaMemcpyAsync(c->hScal + 11, M->d.nAct, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  IBM_CUDA(c, cudaGetLastError());
  if (c->hScal[9] == 1) return ctx_fail(c, IBM_E_CAPACITY, "loop 5: more than CONTACT_CAP simultaneous contacts on one particle");
  if (c->hScal[9] == 5) return ctx_fail(c, IBM_E_CAPACITY, "loop 5: candidate pair list overflow");
  if (c->hScal[9] == 4) return ctx_fail(c, IBM_E_CAPACITY, "loop 5: periodic-patch candidate list overflow");
  if (c->hScal[9] == 3) return ctx_fail(c, IBM_E_CAPACITY, "loop 5: cell-list bin capacity exceeded (strongly polydisperse or overlapping bodies)");
  if (c->hScal[9] == 2) return ctx_fail(c, IBM_E_CAPACITY, "loop 5: instance capacity exhausted (raise ibm_cloud_cfg.maxInstances)");
  c->nInst = c->hScal[8];
  return IBM_OK;
}

static void ensure_bins_clear(ibm_ctx* c) {
  MotionDev& d = ((Motion*)c->motion)->d;
  cudaMemsetAsync(d.binCnt, 0, (size_t)d.nbx * d.nby * d.nbz * sizeof(int), c->stream);
  cudaMemsetAsync(d.pairCount, 0, sizeof(int), c->stream);
  cudaMemsetAsync(d.candCount, 0, sizeof(int), c->stream);
  cudaMemsetAsync(d.hitCount, 0, sizeof(int), c->stream);
}

static void refresh_contact_model(ibm_ctx* c) {
  MotionDev& d = ((Motion*)c->motion)->d;
  d.model = c->collisionModel; d.kn = c->softKn; d.gn = c->softGammaN; d.gt = c->softGammaT; d.mu = c->softMu;
}

void motion_reset_copies(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  if (M && c->nParticles > 0) cudaMemcpy(M->d.maxCopies, c->maxCopies.data(), c->nParticles * sizeof(int), cudaMemcpyHostToDevice);
}

void motion_invalidate_graph(ibm_ctx* c) {
  if (c->motion) ((Motion*)c->motion)->graphValid = false;  // MotionDev is captured by value
}

int motion_substep(ibm_ctx* c, double localDt) {
  int rc = ensure_motion(c);
  if (rc) return rc;
  refresh_contact_model(c);
  ensure_bins_clear(c);
  ((Motion*)c->motion)->d.activeMode = 0;  // a single literal sub-step is always the unsplit chain
  ctx_tic(c, T_ADVANCE);
  substep_launch(c, localDt);
  ctx_toc(c);
  return finish_substeps(c);
}

// literal loop of particleCloud.C:1247-1275 (App. B #22): the schedule of local time steps is plain fp64 scalar arithmetic
static int substep_schedule(double dt, double* dts) {
  int nSub = 0;
  double localDt = dt / 100.0, localTime = localDt;
  while (localTime < dt) {
    if (nSub < 256) dts[nSub] = localDt;
    ++nSub;
    if (localTime + localDt > dt) localDt = dt - localTime;
    else localTime += localDt;
  }
  return nSub;
}

// Owned mode (multi-GPU cluster steps): who owns what.  The slabs of the ranks are gathered once (every rank knows only its
// own, ibm_set_slab); ranks whose slabs do not tile the z-planes in rank order (every rank holding the whole mesh, say) fall
// back to equal shares of the planes.
static int owned_setup(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  const int nR = comm_ranks(c);
  d.owned = 0;
  if (!c->ownedLoop5 || nR < 2 || nR > 16) return IBM_OK;
  if (!M->slabsKnown) {
    int h[17] = {0};
    h[comm_rank(c)] = c->L.kz0 + c->L.ownK0;  // first OWNED plane
    int* dTmp = (int*)M->dPerm;  // (scratch of the step's end, free here)
    IBM_CUDA(c, cudaMemcpyAsync(dTmp, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    int rc = comm_allreduce_sum(c, dTmp, 17, 1);
    if (rc) return rc;
    IBM_CUDA(c, cudaMemcpyAsync(h, dTmp, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    IBM_CUDA(c, cudaStreamSynchronize(c->stream));
    const int nz = c->Lg.nz;
    bool tiles = h[0] == 0;
    for (int r = 1; r < nR; ++r) tiles = tiles && h[r] > h[r - 1] && h[r] < nz;
    for (int r = 0; r < nR; ++r) d.slabK0[r] = tiles ? h[r] : (int)((long long)nz * r / nR);
    d.slabK0[nR] = nz;
    d.ownOz = c->Lg.oz; d.ownDz = c->Lg.dz; d.ownNz = nz;
    d.rank = comm_rank(c); d.nRanks = nR;
    M->slabsKnown = true;
  }
  d.owned = 1;
  return IBM_OK;
}

// the exchange of an owned step: after it every rank holds the new state of every instance, the new copies of all ranks at
// provisional slots in rank order, and the OR of the failure flags
static int owned_exchange(ibm_ctx* c, int n0) {
  Motion* M = (Motion*)c->motion;
  MotionDev& d = M->d;
  const size_t words = (size_t)K5X_HEAD + 14 * (size_t)n0 + (size_t)d.nRanks * K5X_Q * K5X_REC;
  if (M->xWords < words) {
    const size_t want = (size_t)K5X_HEAD + 14 * (size_t)c->cap + (size_t)d.nRanks * K5X_Q * K5X_REC;
    if (M->dX) cudaFree(M->dX);
    M->dX = nullptr; M->xWords = 0;
    IBM_CUDA(c, cudaMalloc(&M->dX, want * sizeof(long long)));
    M->xWords = want;
  }
  const int threads = n0 + d.nRanks * K5X_Q + K5X_HEAD;
  k5x_pack<<<blocks_for(threads), 128, 0, c->stream>>>(d, c->in, n0, M->dX);
  ctx_mark(c, "advance:packed");
  int rc = comm_allreduce_sum(c, M->dX, words, 2);
  if (rc) return rc;
  ctx_mark(c, "advance:all-reduced");
  k5x_unpack<<<blocks_for(threads), 128, 0, c->stream>>>(d, c->in, n0, M->dX);
  if (getenv("IBM_DEBUG_OWNED")) {  // developer switch: the summed header of the exchange
    long long hx[K5X_HEAD];
    cudaMemcpyAsync(hx, M->dX, sizeof(hx), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    fprintf(stderr, "[owned rank %d] n0 %d words %zu total %lld fail %lld violated %lld overflow %lld counts %lld %lld %lld %lld slabs %d %d %d\n", d.rank, n0, words,
            hx[0], hx[1], hx[2], hx[3], hx[8], hx[9], hx[10], hx[11], d.slabK0[0], d.slabK0[1], d.slabK0[2]);
  }
  c->nLaunch += 2;
  M->nOwnedSteps++;
  return IBM_OK;
}

static bool cluster_applies(ibm_ctx* c, int nSub) {
  Motion* M = (Motion*)c->motion;
  return !c->fusedLoop5 && c->activeLoop5 && c->clusterLoop5 && M->splitAllowed && nSub <= 256 && c->nInst > 0;
}

// Called by ibm_solve BEFORE the surface-stress integral: the classification of the coming advance (bounds from the force
// and torque of the previous solve - any bound is safe, the budgets are checked), on a side stream beside the stress
// kernel; the sparse re-zeroing of the forcing field rides along.  motion_advance picks the result up.
int motion_prepare_async(ibm_ctx* c, double dt) {
  int rc = ensure_motion(c);
  if (rc) return rc;
  Motion* M = (Motion*)c->motion;
  M->prepared = false;
  if (!c->asyncPrepare) return IBM_OK;
  double dts[256];
  const int nSub = substep_schedule(dt, dts);
  if (!cluster_applies(c, nSub)) return IBM_OK;
  if (!M->side) {
    // highest priority: its small kernels must be scheduled as blocks of the (SM-filling) stress kernel retire, not after it
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    IBM_CUDA(c, cudaStreamCreateWithPriority(&M->side, cudaStreamNonBlocking, prHi));
    IBM_CUDA(c, cudaEventCreateWithFlags(&M->evFork, cudaEventDisableTiming));
    IBM_CUDA(c, cudaEventCreateWithFlags(&M->evJoin, cudaEventDisableTiming));
  }
  refresh_contact_model(c);
  M->d.activeMode = 1;
  M->d.safety = std::max(1.0, 1e-6 * c->activeSafetyPpm);
  M->d.check = std::min(1.0, 1e-6 * c->activeCheckPpm);
  ensure_bins_clear(c);
  IBM_CUDA(c, cudaMemcpyAsync(M->dDts, dts, nSub * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k5q_snapshot<<<blocks_for(std::max(c->cap, c->nParticles)), 128, 0, c->stream>>>(M->d, c->in, M->snap, M->snapActive, M->snapMaxCopies, 0);
  c->nLaunch++;
  cluster_prepare(c, dt, M->side, M->evFork, M->evJoin);
  M->prepared = true;
  M->preparedDt = dt;
  M->preparedNInst = c->nInst;
  IBM_CUDA(c, cudaGetLastError());
  return IBM_OK;
}
cudaStream_t motion_side_stream(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  return (M && M->prepared) ? M->side : nullptr;
}

int motion_advance(ibm_ctx* c, double dt, ibm_solve_stats* stats) {
  int rc = ensure_motion(c);
  if (rc) return rc;
  Motion* M = (Motion*)c->motion;
  refresh_contact_model(c);
  ctx_tic(c, T_ADVANCE);
  double dts[256];
  int nSub = substep_schedule(dt, dts);
  const bool prepared = M->prepared && M->preparedDt == dt && M->preparedNInst == c->nInst && cluster_applies(c, nSub);
  M->prepared = false;
  if (!prepared) ensure_bins_clear(c);
  if (c->fusedLoop5 && nSub <= 256 && M->fusedBlocks > 0) {
    MotionDev& d = M->d;
    IBM_CUDA(c, cudaMemcpyAsync(M->dDts, dts, nSub * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IBM_CUDA(c, cudaMemsetAsync(d.binCnt, 0, (size_t)d.nbx * d.nby * d.nbz * sizeof(int), c->stream));
    int doCyclic = (d.nWalls > 0 || c->cfg.cyclicWithoutWalls) ? 1 : 0;
    const ShapeClassDev* cls = c->dClasses;
    const double* dd = M->dDts;
    void* args[] = {&d, &c->Lg, &c->in, &cls, &dd, &nSub, &doCyclic};
    IBM_CUDA(c, cudaLaunchCooperativeKernel((void*)k5_fused, dim3(M->fusedBlocks), dim3(256), args, 0, c->stream));
    c->nLaunch++;
  } else {
    // quiet / active split (see k5q_bounds): same results, the per-sub-step chain only over the instances that can interact
    const bool split = c->activeLoop5 && M->splitAllowed && nSub <= 256 && c->nInst > 0;
    const int nInst0 = c->nInst;
    M->d.activeMode = split ? 1 : 0;
    M->d.safety = std::max(1.0, 1e-6 * c->activeSafetyPpm);  // (a quiet instance needs B_i >= s_i)
    M->d.check = std::min(1.0, 1e-6 * c->activeCheckPpm);
    if (split && !prepared) {
      IBM_CUDA(c, cudaMemcpyAsync(M->dDts, dts, nSub * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      k5q_snapshot<<<blocks_for(std::max(c->cap, c->nParticles)), 128, 0, c->stream>>>(M->d, c->in, M->snap, M->snapActive, M->snapMaxCopies, 0);
      c->nLaunch++;
    }
    auto sequence = [&]() {
      if (split) active_prologue(c, dt, nSub);
      for (int s = 0; s < nSub; ++s) {
        const double ls = s < 256 ? dts[s] : dts[255];
        if (split) substep_launch_active(c, ls); else substep_launch(c, ls);
      }
    };
    const bool cluster = split && c->clusterLoop5;
    bool clusterRedo = false;
    if (cluster) {
      // one launch sequence per fluid step, no per-sub-step kernel boundary; flags, the instance count and the creation
      // records of new copies come back with ONE synchronisation
      if (prepared) IBM_CUDA(c, cudaStreamWaitEvent(c->stream, M->evJoin, 0));  // classified beside the stress integral
      else cluster_prepare(c, dt);
      if ((rc = owned_setup(c))) { ctx_toc(c); return rc; }
      if (M->d.owned) IBM_CUDA(c, cudaMemsetAsync(M->d.mine, 0, (size_t)c->cap * sizeof(int), c->stream));
      ctx_mark(c, "advance:classified");
      cluster_execute(c, nSub);
      ctx_mark(c, "advance:sub-steps done");
      if (M->d.owned && (rc = owned_exchange(c, nInst0))) { ctx_toc(c); return rc; }
      ctx_mark(c, "advance:exchanged");
      M->nClusterSteps++;
      MotionDev& d = M->d;
      int* h = M->hEv;
      IBM_CUDA(c, cudaMemcpyAsync(h + 0, d.nextSlot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaMemcpyAsync(h + 1, d.clusterFail, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaMemcpyAsync(h + 2, d.violated, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaMemcpyAsync(h + 3, d.nAct, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      IBM_CUDA(c, cudaMemcpyAsync(h + 4, d.overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      const int evn = std::min(K5C_EVN, c->cap - nInst0);
      if (evn > 0) {
        IBM_CUDA(c, cudaMemcpyAsync(h + 16, d.evSub + nInst0, (size_t)evn * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IBM_CUDA(c, cudaMemcpyAsync(h + 16 + K5C_EVN, d.evKey + nInst0, (size_t)evn * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      }
      IBM_CUDA(c, cudaStreamSynchronize(c->stream));
      ctx_mark(c, "advance:flags read");
      IBM_CUDA(c, cudaGetLastError());
      M->lastActive = h[3];
      c->hScal[11] = h[3];
      M->lastFlags = h[1] * 1 + h[2] * 1000 + h[4] * 1000000;
      if (h[1] || h[2] || h[4]) {
        clusterRedo = true;  // no room in a block, a budget outrun, or a table overflow: redo the step unsplit (which reports real overflows)
      } else {
        const int total = h[0] - nInst0;
        if (total > 0) {
          std::vector<int> sub(total), key(total);
          if (total <= evn) {
            std::copy(h + 16, h + 16 + total, sub.begin());
            std::copy(h + 16 + K5C_EVN, h + 16 + K5C_EVN + total, key.begin());
          } else {
            IBM_CUDA(c, cudaMemcpy(sub.data(), d.evSub + nInst0, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
            IBM_CUDA(c, cudaMemcpy(key.data(), d.evKey + nInst0, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
          }
          // the reference appends, sub-step by sub-step, patch-major and in cloud order of the source within a patch
          // (particleCloud.C:685-832); a source may itself be a copy created in an earlier sub-step of this step
          std::vector<int> ord(total), fin(total, -1), perm(total);
          for (int r = 0; r < total; ++r) ord[r] = r;
          std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return sub[a] < sub[b]; });
          int placed = 0;
          for (int a = 0; a < total;) {
            int b = a;
            while (b < total && sub[ord[b]] == sub[ord[a]]) ++b;
            auto srcPos = [&](int r) {
              const int src = key[r] % c->cap;
              return src < nInst0 ? src : nInst0 + fin[src - nInst0];
            };
            std::sort(ord.begin() + a, ord.begin() + b, [&](int x, int y) {
              const int px = key[x] / c->cap, py = key[y] / c->cap;
              if (px != py) return px < py;
              return srcPos(x) < srcPos(y);
            });
            for (int q = a; q < b; ++q) { fin[ord[q]] = placed; perm[placed] = ord[q]; ++placed; }
            a = b;
          }
          bool identity = true;
          for (int r = 0; r < total; ++r) identity = identity && perm[r] == r;
          if (!identity) {
            IBM_CUDA(c, cudaMemcpyAsync(M->dPerm, perm.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            k5c_permute<<<blocks_for(total), 128, 0, c->stream>>>(c->in, c->in2, M->dPerm, nInst0, total, 0);
            k5c_permute<<<blocks_for(total), 128, 0, c->stream>>>(c->in2, c->in, M->dPerm, nInst0, total, 1);
            c->nLaunch += 2;
            IBM_CUDA(c, cudaStreamSynchronize(c->stream));  // perm is a local
          }
          k_set_int<<<1, 1, 0, c->stream>>>(d.nInst, nInst0 + total);
          c->nLaunch++;
          c->nInst = nInst0 + total;
          M->groupsValid = false;  // the same-index lists hold provisional slots: relink from the final order
          M->graphValid = false;
        }
      }
    } else if (split && c->fusedActive && M->fusedActBlocks > 0) {
      // prologue as plain launches, then ONE cooperative kernel for every sub-step of the active list.  Its grid is sized
      // from the active count of the previous step (the kernel strides, so any size is correct) and bounded by co-residency.
      active_prologue(c, dt, nSub);
      const int guess = M->lastActive > 0 ? M->lastActive : std::max(1, c->nInst / 3);
      const int blocks = std::max(std::min(M->fusedActBlocks, 4), std::min(M->fusedActBlocks, (int)((guess + guess / 4 + 255) / 256)));
      MotionDev& d = M->d;
      int doCyclic = (d.nCyc > 0 && (d.nWalls > 0 || c->cfg.cyclicWithoutWalls)) ? 1 : 0;
      const ShapeClassDev* cls = c->dClasses;
      const double* dd = M->dDts;
      void* args[] = {&d, &c->Lg, &c->in, &cls, &dd, &nSub, &doCyclic};
      if (cudaLaunchCooperativeKernel((void*)k5a_fused, dim3(blocks), dim3(256), args, 0, c->stream) == cudaSuccess) {
        c->nLaunch++;
      } else {
        // refused on this device / in this context (co-residency cannot be guaranteed): the unfused chain, now and from here on
        cudaGetLastError();
        M->fusedActBlocks = 0;
        for (int s = 0; s < nSub; ++s) substep_launch_active(c, s < 256 ? dts[s] : dts[255]);
      }
    } else if (c->graphLoop5 && nSub <= 256) {
      // the launch sequence of one fluid step is fixed (same kernels, same grids, same pointers): capture it once
      // per schedule of local time steps and replay it — the sub-step kernels are a few microseconds each
      bool same = M->graphExec && M->graphValid && M->graphActive == split && M->graphSafety == M->d.safety * M->d.check && M->graphDt == dt && c->nInst <= M->graphNInst && (int)M->graphDts.size() == nSub;
      for (int s = 0; same && s < nSub; ++s) same = M->graphDts[s] == dts[s];
      bool graphOk = true;
      if (!same) {
        if (M->graphExec) { cudaGraphExecDestroy(M->graphExec); M->graphExec = nullptr; }
        M->graphValid = false;
        cudaGraph_t g = nullptr;
        // no early return between BeginCapture and EndCapture: an error inside must end the capture, drop the partial graph
        // and leave the timer pair balanced (the step then runs as plain launches)
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
          const long long before = c->nLaunch;
          sequence();
          M->graphLaunches = c->nLaunch - before;
          c->nLaunch = before;
          e = cudaStreamEndCapture(c->stream, &g);
          if (e == cudaSuccess) e = cudaGraphInstantiate(&M->graphExec, g, 0);
          if (g) cudaGraphDestroy(g);
        }
        if (e != cudaSuccess) {
          cudaGetLastError();
          M->graphExec = nullptr;
          graphOk = false;
        } else {
          M->graphDts.assign(dts, dts + nSub);
          M->graphValid = true;
          M->graphActive = split;
          M->graphSafety = M->d.safety * M->d.check;
          M->graphDt = dt;
          M->graphNInst = c->nInst;
        }
      }
      if (graphOk) {
        if (!ctx_cuda_ok(c, cudaGraphLaunch(M->graphExec, c->stream), "cudaGraphLaunch")) { ctx_toc(c); return IBM_E_CUDA; }
        c->nLaunch += M->graphLaunches;
      } else {
        sequence();
      }
    } else {
      sequence();
    }
    if (split) {
      if (!cluster) {
        if ((rc = finish_substeps(c))) { ctx_toc(c); return rc; }
        M->lastActive = c->hScal[11];
      }
      if (cluster ? clusterRedo : c->hScal[10] != 0) {
        // an active instance outran the displacement budget: put the start of the step back and redo it without the split
        M->nFallbacks++;
        k5q_snapshot<<<blocks_for(std::max(c->cap, c->nParticles)), 128, 0, c->stream>>>(M->d, c->in, M->snap, M->snapActive, M->snapMaxCopies, 1);
        k_set_int<<<1, 1, 0, c->stream>>>(M->d.nInst, nInst0);
        c->nLaunch += 2;
        c->nInst = nInst0;
        M->d.activeMode = 0;
        M->groupsValid = false;  // copies appended by the abandoned attempt are gone: relink the same-index lists
        if ((rc = ensure_motion(c))) { ctx_toc(c); return rc; }
        ensure_bins_clear(c);
        for (int s = 0; s < nSub; ++s) substep_launch(c, s < 256 ? dts[s] : dts[255]);
      }
    }
  }
  ctx_toc(c);
  ctx_mark(c, "advance:copies ordered");
  if ((rc = finish_substeps(c))) return rc;
  ctx_mark(c, "advance:finished");
  // update() for every instance, deleteInActiveParticles, (indicator), then the log block
  if ((rc = ctx_update(c))) return rc;  // invalidates the same-index lists when instances were deleted
  if ((rc = ensure_motion(c))) return rc;
  if (stats) {
    k_stats_partial<<<STATS_BLOCKS, 256, 0, c->stream>>>(M->d, c->in, c->dClasses, c->nInst, M->dStats + 16);
    k_stats_final<<<1, 32, 0, c->stream>>>(M->dStats + 16, M->dStats);
    c->nLaunch += 2;
    double h[10];
    IBM_CUDA(c, cudaMemcpyAsync(h, M->dStats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    IBM_CUDA(c, cudaStreamSynchronize(c->stream));
    const double np = c->nParticles > 0 ? (double)c->nParticles : 1.0;
    stats->nSubSteps = nSub;
    stats->nInstances = c->nInst;
    for (int q = 0; q < 3; ++q) { stats->avgOmega[q] = h[q] / np; stats->avgV[q] = h[3 + q] / np; stats->avgP[q] = h[6 + q] / np; }
    stats->totK = h[9];
  }
  return IBM_OK;
}

void motion_destroy(ibm_ctx* c) {
  Motion* M = (Motion*)c->motion;
  if (!M) return;
  if (M->graphExec) cudaGraphExecDestroy(M->graphExec);
  if (M->hEv) cudaFreeHost(M->hEv);
  if (M->side) { cudaStreamSynchronize(M->side); cudaStreamDestroy(M->side); cudaEventDestroy(M->evFork); cudaEventDestroy(M->evJoin); }
  for (void* p : M->allocs) cudaFree(p);
  if (M->cubTemp) cudaFree(M->cubTemp);
  if (M->dX) cudaFree(M->dX);
  delete M;
  c->motion = nullptr;
}

}  // namespace ibm
