// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// OpenFOAM-free mini-mesh: a uniform hex lattice with blockMesh cell numbering, expanded into
// the explicit arrays the reference reads from polyMesh/fvMesh:
//   cellCentres()  faceCentres()  cells()  nInternalFaces()  faceOwner()/faceNeighbour()
//   surfaceInterpolation weights   boundary patches (Cf, Sf, magSf)   findCell()
// Conventions restated from OpenFOAM (un-vendored dependency, Foundation "dev" ~ v6):
//   * blockMesh numbers cells i + nx*(j + ny*k) inside a block, blocks concatenated;
//   * internal faces are in upper-triangular order (owner ascending, then neighbour ascending);
//   * primitiveMesh::calcCells: cells()[c] = faces owned by c (ascending id) then faces of which
//     c is the neighbour (ascending id);  boundary faces (id >= nInternalFaces) come after all
//     internal ones and are skipped by the reference (particleShape.C:59), so only internal
//     faces are listed here, in that relative order;
//   * linear interpolation weights  w = |Sf.(C_N-Cf)| / (|Sf.(Cf-C_P)| + |Sf.(C_N-Cf)|)
//     (surfaceInterpolation::makeWeights), face value = w*(phi_P - phi_N) + phi_N.
// Mini-mesh definitions that are OURS (no OpenFOAM counterpart to copy):
//   * cell centre (I,J,K) = (x0 + (I+0.5)*dx, y0 + (J+0.5)*dy, z0 + (K+0.5)*dz)  [mul then add];
//   * face centre on the plane between I and I+1: x = x0 + (I+1)*dx;
//   * patch faces are ordered by transverse lattice index, fastest first
//     (x-sides: J then K; y-sides: I then K; z-sides: I then J), so cyclic pairs match by index;
//   * findCell(pt) == floor((pt - origin)/delta) per direction with a bounds check (DEFINED for
//     points exactly on a face; agrees with OpenFOAM's nearest-centre + pointInCell search for
//     points in generic position).  The "faithful" variant below does the linear scans.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "orc_math.hpp"

namespace orc {

enum PatchType { PATCH = 0, WALL = 1, CYCLIC = 2, EMPTY = 3, PROCESSOR = 4 };

struct Patch {
  int type = PATCH;
  int side = 0;          // 0 xmin, 1 xmax, 2 ymin, 3 ymax, 4 zmin, 5 zmax
  int neighbPatch = -1;  // cyclic partner (patch id)
  std::vector<vec3> Cf, Sf;
  std::vector<double> magSf;
  std::vector<int> faceCells;
  int size() const { return (int)Cf.size(); }
};

struct Mesh {
  int nx = 0, ny = 0, nz = 0;
  int kz0 = 0;  // slab of a decomposed mesh: global z-plane of local plane 0; geometry uses the global index
                // (decomposePar keeps the global points, so processor-mesh cell centres equal the undecomposed ones)
  vec3 origin{0, 0, 0}, delta{1, 1, 1};
  int nCells = 0, nInternalFaces = 0;
  std::vector<vec3> C;        // cellCentres()
  std::vector<double> V;      // cell volumes
  std::vector<int> owner, neighbour;
  std::vector<vec3> Cf, Sf;   // internal faces only
  std::vector<double> lambda; // linear weights
  std::vector<int> cellFaceStart, cellFaces;  // cells(), internal faces only
  std::vector<int> latticeToCell, cellToLattice;
  std::vector<Patch> patches;
  vec3 bbMin{0, 0, 0}, bbMax{0, 0, 0};  // min/max(mesh.points())
  int nGeometricD = 3;

  int lat(int I, int J, int K) const { return I + nx * (J + ny * K); }

  // DEFINED findCell (floor rule)
  int findCell(const vec3& p) const {
    double tx = std::floor((p.x - origin.x) / delta.x);
    double ty = std::floor((p.y - origin.y) / delta.y);
    double tz = std::floor((p.z - origin.z) / delta.z) - kz0;
    if (!(tx >= 0 && tx < nx && ty >= 0 && ty < ny && tz >= 0 && tz < nz)) return -1;
    return latticeToCell[lat((int)tx, (int)ty, (int)tz)];
  }

  bool pointInCell(const vec3& p, int celli) const {
    // primitiveMesh::pointInCell is a face-plane test; for the hex lattice the six planes are the
    // cell box.  Ties (point exactly on a plane) are implementation-defined in OpenFOAM (tet
    // decomposition + tolerances); DEFINED here by the same floor arithmetic as findCell so that
    // the linear-scan search and the ijk search agree on every input.
    int l = cellToLattice[celli];
    int I = l % nx, J = (l / nx) % ny, K = l / (nx * ny);
    return std::floor((p.x - origin.x) / delta.x) == I && std::floor((p.y - origin.y) / delta.y) == J &&
           std::floor((p.z - origin.z) / delta.z) - kz0 == K;
  }

  // polyMesh::findCell as the reference calls it: nearest cell centre by linear scan (first
  // minimum wins), test that cell, otherwise scan every cell (O(nCells) per call).
  int findCellFaithful(const vec3& p) const {
    if (nCells == 0) return -1;
    int best = 0;
    double bestD = magSqr(C[0] - p);
    for (int c = 1; c < nCells; ++c) {
      double d2 = magSqr(C[c] - p);
      if (d2 < bestD) { bestD = d2; best = c; }
    }
    if (pointInCell(p, best)) return best;
    for (int c = 0; c < nCells; ++c)
      if (pointInCell(p, c)) return c;
    return -1;
  }
};

struct BlockDesc { int i0, i1, j0, j1, k0, k1; };

inline Mesh* buildMesh(int nx, int ny, int nz, vec3 origin, vec3 delta, const std::vector<BlockDesc>& blocksIn,
                       const std::vector<int>& patchSide, const std::vector<int>& patchType,
                       const std::vector<int>& patchNeighb, int kz0 = 0) {
  Mesh* m = new Mesh;
  m->nx = nx; m->ny = ny; m->nz = nz; m->origin = origin; m->delta = delta; m->kz0 = kz0;
  std::vector<BlockDesc> blocks = blocksIn;
  if (blocks.empty()) blocks.push_back({0, nx, 0, ny, 0, nz});
  const int64_t n = (int64_t)nx * ny * nz;
  m->nCells = (int)n;
  m->latticeToCell.assign(n, -1);
  m->cellToLattice.assign(n, -1);
  int next = 0;
  for (const auto& b : blocks)
    for (int K = b.k0; K < b.k1; ++K)
      for (int J = b.j0; J < b.j1; ++J)
        for (int I = b.i0; I < b.i1; ++I) {
          int l = m->lat(I, J, K);
          m->latticeToCell[l] = next;
          m->cellToLattice[next] = l;
          ++next;
        }
  m->C.resize(n);
  m->V.assign(n, delta.x * delta.y * delta.z);
  for (int c = 0; c < (int)n; ++c) {
    int l = m->cellToLattice[c];
    int I = l % nx, J = (l / nx) % ny, K = l / (nx * ny);
    m->C[c] = {origin.x + (I + 0.5) * delta.x, origin.y + (J + 0.5) * delta.y, origin.z + (K + kz0 + 0.5) * delta.z};
  }
  m->bbMin = {origin.x, origin.y, origin.z + kz0 * delta.z};
  m->bbMax = {origin.x + nx * delta.x, origin.y + ny * delta.y, origin.z + (kz0 + nz) * delta.z};

  // internal faces, upper-triangular order
  std::vector<std::vector<int>> nbrFaces(n);  // faces for which the cell is the neighbour
  m->cellFaceStart.assign(n + 1, 0);
  std::vector<std::vector<int>> ownFaces(n);
  const int dI[6] = {1, -1, 0, 0, 0, 0}, dJ[6] = {0, 0, 1, -1, 0, 0}, dK[6] = {0, 0, 0, 0, 1, -1};
  for (int c = 0; c < (int)n; ++c) {
    int l = m->cellToLattice[c];
    int I = l % nx, J = (l / nx) % ny, K = l / (nx * ny);
    struct Cand { int nb, dir; };
    Cand cand[6];
    int nc = 0;
    for (int d = 0; d < 6; ++d) {
      int I2 = I + dI[d], J2 = J + dJ[d], K2 = K + dK[d];
      if (I2 < 0 || I2 >= nx || J2 < 0 || J2 >= ny || K2 < 0 || K2 >= nz) continue;
      int nb = m->latticeToCell[m->lat(I2, J2, K2)];
      if (nb > c) cand[nc++] = {nb, d};
    }
    std::sort(cand, cand + nc, [](const Cand& a, const Cand& b) { return a.nb < b.nb; });
    for (int q = 0; q < nc; ++q) {
      int f = (int)m->owner.size();
      m->owner.push_back(c);
      m->neighbour.push_back(cand[q].nb);
      int d = cand[q].dir;
      vec3 cf = m->C[c];
      vec3 sf{0, 0, 0};
      if (d == 0) { cf.x = origin.x + (I + 1) * delta.x; sf.x = delta.y * delta.z; }
      if (d == 1) { cf.x = origin.x + I * delta.x;       sf.x = -(delta.y * delta.z); }
      if (d == 2) { cf.y = origin.y + (J + 1) * delta.y; sf.y = delta.x * delta.z; }
      if (d == 3) { cf.y = origin.y + J * delta.y;       sf.y = -(delta.x * delta.z); }
      if (d == 4) { cf.z = origin.z + (K + kz0 + 1) * delta.z; sf.z = delta.x * delta.y; }
      if (d == 5) { cf.z = origin.z + (K + kz0) * delta.z;     sf.z = -(delta.x * delta.y); }
      m->Cf.push_back(cf);
      m->Sf.push_back(sf);
      ownFaces[c].push_back(f);
      nbrFaces[cand[q].nb].push_back(f);
    }
  }
  m->nInternalFaces = (int)m->owner.size();
  m->lambda.resize(m->nInternalFaces);
  for (int f = 0; f < m->nInternalFaces; ++f) {
    double SfdOwn = std::fabs(dot(m->Sf[f], m->Cf[f] - m->C[m->owner[f]]));
    double SfdNei = std::fabs(dot(m->Sf[f], m->C[m->neighbour[f]] - m->Cf[f]));
    m->lambda[f] = SfdNei / (SfdOwn + SfdNei);
  }
  for (int c = 0; c < (int)n; ++c) {
    m->cellFaceStart[c] = (int)m->cellFaces.size();
    for (int f : ownFaces[c]) m->cellFaces.push_back(f);
    for (int f : nbrFaces[c]) m->cellFaces.push_back(f);  // already ascending (faces created in id order)
  }
  m->cellFaceStart[n] = (int)m->cellFaces.size();

  // boundary patches
  int nEmptyDirs = 0;
  bool emptyDir[3] = {false, false, false};
  for (size_t p = 0; p < patchSide.size(); ++p) {
    Patch pt;
    pt.side = patchSide[p];
    pt.type = patchType[p];
    pt.neighbPatch = patchNeighb[p];
    int s = pt.side;
    int axis = s / 2;
    bool hiSide = (s & 1);
    if (pt.type == EMPTY) emptyDir[axis] = true;
    int nA = (axis == 0) ? ny : nx;              // fastest transverse extent
    int nB = (axis == 2) ? ny : nz;              // slowest transverse extent
    for (int b = 0; b < nB; ++b)
      for (int a = 0; a < nA; ++a) {
        int I, J, K;
        if (axis == 0) { I = hiSide ? nx - 1 : 0; J = a; K = b; }
        else if (axis == 1) { I = a; J = hiSide ? ny - 1 : 0; K = b; }
        else { I = a; J = b; K = hiSide ? nz - 1 : 0; }
        int c = m->latticeToCell[m->lat(I, J, K)];
        vec3 cf = m->C[c];
        vec3 sf{0, 0, 0};
        double sgn = hiSide ? 1.0 : -1.0;
        if (axis == 0) { cf.x = hiSide ? origin.x + nx * delta.x : origin.x; sf.x = sgn * (delta.y * delta.z); }
        if (axis == 1) { cf.y = hiSide ? origin.y + ny * delta.y : origin.y; sf.y = sgn * (delta.x * delta.z); }
        if (axis == 2) { cf.z = hiSide ? origin.z + (kz0 + nz) * delta.z : origin.z + kz0 * delta.z; sf.z = sgn * (delta.x * delta.y); }
        pt.Cf.push_back(cf);
        pt.Sf.push_back(sf);
        pt.magSf.push_back(mag(sf));
        pt.faceCells.push_back(c);
      }
    m->patches.push_back(std::move(pt));
  }
  for (int a = 0; a < 3; ++a) nEmptyDirs += emptyDir[a] ? 1 : 0;
  m->nGeometricD = 3 - nEmptyDirs;
  return m;
}

}  // namespace orc
