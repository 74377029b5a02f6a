// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// PARITY PINNED against the reference's own code: oracle/_ref (the reference's sources compiled against
// oracle/mockfoam) agrees with this restatement bit for bit (tests/test_oracle_vs_reference.py), and the vectors that
// build produced are committed (tests/golden/, tests/test_golden_oracle.py).  OpenFOAM's own arithmetic (mesh geometry,
// findCell, linear face interpolation) is restated, not pinned: see oracle/README.md.
//
// CPU restatement of IBMFoam's immersed-boundary hot path: one function per reference function,
// same loop order and floating-point evaluation order.  Citations are relative to /root/reference.
// Deviations from "as coded" are the DEFINED rows of SURVEY.md Appendix B and are marked [B#n].
#pragma once
#include <cstdio>
#include <list>
#include <map>
#include <memory>
#include <vector>

#include "orc_mesh.hpp"

namespace orc {

enum ShapeKind { SPHERE = 0, CYLINDER = 1, ELLIPSE = 2 };
enum LocationType { OFF_MESH = 0, PARTIAL = 1, ON_MESH = 2 };  // particleShape.H:62-67

struct ShapeDict {
  int kind = SPHERE;
  double d = 0, a = 0, b = 0;
  int nTheta = 0, nk = 1;
  double delta = 0;
  vec3 theta{0, 0, 0};
};

struct ParticleDict {
  ShapeDict shape;
  vec3 position{0, 0, 0}, v{0, 0, 0}, omega{0, 0, 0};
  double rho = 0;
};

struct SearchMode {
  bool faithful = false;  // true: linear scans exactly like the reference; false: ijk + bounding box
  bool staleTables = false;  // true: quirk B#8 AS CODED (gather tables of points that left the mesh keep their old
                             // entries) - only used to compare against the compiled reference (oracle/_ref)
};

// ---------------------------------------------------------------------------------------------
// particleShape (+ sphere / cylinder / ellipse)   src/particles/particleShapes/
// ---------------------------------------------------------------------------------------------
struct Shape {
  const Mesh* pMesh = nullptr;
  SearchMode mode;
  int kind = SPHERE;
  int onMesh = ON_MESH;
  vec3 center{0, 0, 0};
  double momentOfInertia = HUGE_VAL;
  static constexpr int nRadial = 3;
  int nTheta = 0, nk = 1, N = 0;
  vec3 theta{0, 0, 0};
  double delta = 0;
  std::vector<vec3> centeredMesh, mesh;
  tensor rotationMatrix = TensorI;
  std::vector<vec3> Sf;
  std::vector<int> shellCells;
  std::vector<ivec3> neighbourPoints;
  std::vector<std::array<int, 8>> Is;
  std::vector<std::array<int, 4>> Os;
  std::vector<std::vector<double>> wToLocal;
  std::vector<double> WToLocal;
  std::vector<std::vector<int>> facesToLocal;
  std::vector<int> pointCell;  // tap only: findCell(mesh[pti]) as seen by setWeights
  std::vector<std::array<double, 8>> wFromLocal;
  std::vector<double> WFromLocal;
  double d_ = 0, a_ = 0, b_ = 0, l_ = 0;

  // particleShape.H:164-180
  int index(int i, int j, int k) const { return i + j * nRadial + k * nRadial * nTheta; }
  int index2(int i, int j) const { return i + j * nTheta; }

  int findCell(const vec3& p) const { return mode.faithful ? pMesh->findCellFaithful(p) : pMesh->findCell(p); }

  void resizeForN() {
    centeredMesh.assign(N, Zero3);
    mesh.assign(N, Zero3);
    Sf.assign(nTheta * nk, Zero3);
    neighbourPoints.assign(N, ivec3{0, 0, 0});
    wToLocal.assign(N, {});
    WToLocal.assign(N, 0.0);
    facesToLocal.assign(N, {});
    pointCell.assign(N, -1);
  }

  // dictionary constructors: particleShape.C:228-253, sphere.C:53-92, cylinder.C:53-82, ellipse.C:68-98
  void constructFromDict(const Mesh* m, SearchMode sm, const ShapeDict& dict, const vec3& c) {
    pMesh = m;
    mode = sm;
    kind = dict.kind;
    onMesh = ON_MESH;
    center = c;
    momentOfInertia = HUGE_VAL;
    nTheta = dict.nTheta;
    nk = dict.nk;
    N = nRadial * nTheta * nk;
    theta = dict.theta;
    delta = dict.delta;
    if (kind == SPHERE) {
      d_ = dict.d;
      momentOfInertia = 2.0 / 5.0 * sqr(d_ / 2.0);
      nk = nTheta / 2;  // [B#2] the guard looks for "nk_", so this always runs (sphere.C:66-69)
      N = nRadial * nTheta * nk;
    } else if (kind == CYLINDER) {
      d_ = dict.d;
      momentOfInertia = sqr(d_ / 8.0);  // [B#14] as coded
      nk = 1;
      N = nRadial * nTheta * nk;
      l_ = std::fabs(pMesh->bbMax.z - pMesh->bbMin.z);
    } else {
      a_ = dict.a;
      b_ = dict.b;
      momentOfInertia = sqr(a_) + sqr(b_);  // [B#14] as coded
      nk = 1;
      N = nRadial * nTheta * nk;
      l_ = std::fabs(pMesh->bbMax.z - pMesh->bbMin.z);
    }
    resizeForN();
    discretize();
    updateCellLists();
    calcSf();
  }

  // copy constructors: particleShape.C:256-281 + sphere.C:95-107 / cylinder.C:84-97 / ellipse.C:100-114
  void constructCopy(const Shape& s, const vec3& c, const vec3& th) {
    pMesh = s.pMesh;
    mode = s.mode;
    kind = s.kind;
    onMesh = PARTIAL;
    center = c;
    momentOfInertia = s.momentOfInertia;
    nTheta = s.nTheta;
    nk = s.nk;
    N = s.N;
    theta = th;
    delta = s.delta;
    centeredMesh = s.centeredMesh;
    mesh.assign(N, Zero3);
    Sf = s.Sf;
    neighbourPoints.assign(N, ivec3{0, 0, 0});
    wToLocal.assign(N, {});
    WToLocal.assign(N, 0.0);
    facesToLocal.assign(N, {});
    pointCell.assign(N, -1);
    shellCells.clear();
    Is.clear();
    Os.clear();
    wFromLocal.clear();
    WFromLocal.clear();
    d_ = s.d_; a_ = s.a_; b_ = s.b_; l_ = s.l_;
    moveMesh(c);
  }

  // sphere.C:163-195, cylinder.C:126-152, ellipse.C:124-150
  void discretize() {
    vec3 dl{-1 * delta, 0, delta};
    double dTheta = 2.0 * PI / (nTheta);
    if (kind == SPHERE) {
      double R = d_ / 2.0;
      double dPhi = PI / (nk - 1);
      for (int i = 0; i < nRadial; i++) {
        double r = R + dl[i];
        for (int j = 0; j < nTheta; j++) {
          double th = double(j) * dTheta;
          for (int k = 0; k < nk; k++) {
            double phi = dPhi * double(k);
            centeredMesh[index(i, j, k)] = {r * std::cos(th) * std::sin(phi), r * std::sin(th) * std::sin(phi),
                                            r * std::cos(phi)};
          }
        }
      }
    } else if (kind == CYLINDER) {
      double R = d_ / 2.0;
      for (int i = 0; i < nRadial; i++) {
        double r = R + dl[i];
        for (int j = 0; j < nTheta; j++) {
          double th = double(j) * dTheta;
          centeredMesh[index(i, j, 0)] = {r * std::cos(th), r * std::sin(th), 0.0};
        }
      }
    } else {
      for (int i = 0; i < nRadial; i++)
        for (int j = 0; j < nTheta; j++) {
          double th = double(j) * dTheta;
          double R = std::sqrt(1.0 / (sqr(std::cos(th) / (a_)) + sqr(std::sin(th) / (b_)))) + dl[i];
          centeredMesh[index(i, j, 0)] = {R * std::cos(th), R * std::sin(th), 0.0};
        }
    }
    moveMesh(center);  // [B#1] sphere.C:194 uses an uninitialised shadow member; DEFINED: base centre
  }

  // sphere.C:119-160, cylinder.C:107-124, ellipse.C:50-59  (area vectors point inward)
  void calcSf() {
    double dTheta = 2.0 * PI / double(nTheta);
    if (kind == SPHERE) {
      double r = d_ / 2.0;
      double dPhi = PI / double(nk - 1);
      for (int i = 0; i < nTheta; i++) {
        double th = dTheta * double(i);
        for (int k = 0; k < nk; k++) {
          double phi = dPhi * double(k);
          vec3 s = -vec3{std::cos(th) * std::sin(phi), std::sin(th) * std::sin(phi), std::cos(phi)} * sqr(r) * dTheta;
          if (k == 0) s *= (std::cos(phi) - std::cos(phi + dPhi / 2.0));
          else if (k == nk - 1) s *= (std::cos(phi - dPhi / 2.0) - std::cos(phi));
          else s *= (std::cos(phi - dPhi / 2.0) - std::cos(phi + dPhi / 2.0));
          Sf[index2(i, k)] = s;
        }
      }
    } else if (kind == CYLINDER) {
      double r = d_ / 2.0;
      for (int i = 0; i < nTheta; i++) {
        double th = dTheta * double(i);
        Sf[index2(i, 0)] = -vec3{std::cos(th), std::sin(th), 0.0} * r * dTheta * l_;
      }
    } else {
      for (int i = 0; i < nTheta; i++) Sf[index2(i, 0)] = -centeredMesh[index(1, i, 0)] * dTheta * l_;
    }
  }

  // particleShape.C:186-223
  void setRotationMatrix() {
    rotationMatrix = TensorI;
    if (std::fabs(theta.x) > SMALL) {
      double t1 = theta.x;
      rotationMatrix = dot(rotationMatrix, tensor{1.0, 0.0, 0.0, 0.0, std::cos(t1), -std::sin(t1), 0.0, std::sin(t1), std::cos(t1)});
    }
    if (std::fabs(theta.y) > SMALL) {
      double t2 = theta.y;
      rotationMatrix = dot(rotationMatrix, tensor{std::cos(t2), 0.0, std::sin(t2), 0.0, 1.0, 0.0, -std::sin(t2), 0.0, std::cos(t2)});
    }
    if (std::fabs(theta.z) > SMALL) {
      double t3 = theta.z;
      rotationMatrix = dot(rotationMatrix, tensor{std::cos(t3), -std::sin(t3), 0.0, std::sin(t3), std::cos(t3), 0.0, 0.0, 0.0, 1.0});
    }
  }

  // particleShape.C:291-302; rotate(): particleShape.H:192-195 (identity), ellipse.C:61-64
  void moveMesh(const vec3& c) {
    center = c;
    setRotationMatrix();
    if (kind == ELLIPSE)
      for (int n = 0; n < N; ++n) mesh[n] = dot(rotationMatrix, centeredMesh[n]);
    else
      mesh = centeredMesh;
    for (int n = 0; n < N; ++n) mesh[n] += center;
  }

  // sphere.H:115-118, cylinder.H:117-120, ellipse.C:222-227
  double r(const vec3& pt) const {
    if (kind != ELLIPSE) return d_ / 2.0;
    vec3 diff = pt - center;
    double th = std::atan2(diff.y, diff.x) - theta.z;
    return std::sqrt(1.0 / (sqr(std::cos(th) / (a_)) + sqr(std::sin(th) / (b_))));
  }
  double d() const {
    if (kind != ELLIPSE) return d_;
    return 1.55 * std::pow(A(), 0.625) / (TWOPI * std::sqrt(0.5 * sqr(a_) + (sqr(b_))));  // ellipse.C:206-214
  }
  vec3 D() const { return kind == ELLIPSE ? vec3{a_, b_, l_} : vec3{d_, d_, d_}; }
  double A() const {
    if (kind == SPHERE) return PI * sqr(d_ / 2.0);
    if (kind == CYLINDER) return d_ * l_;
    return a_ * b_ * PI;
  }
  double V() const {
    if (kind == SPHERE) return 1.0 / 6.0 * PI * pow3(d_);
    if (kind == CYLINDER) return PI * sqr(d_ / 2.0) * l_;
    return PI * a_ * b_ * l_;
  }
  double boundingRadius() const { return kind == ELLIPSE ? std::fmax(a_, b_) : d_ / 2.0; }

  // candidate cells in ascending cell-id order
  void candidateCells(std::vector<int>& out) const {
    out.clear();
    const Mesh& m = *pMesh;
    if (mode.faithful) {
      out.resize(m.nCells);
      for (int c = 0; c < m.nCells; ++c) out[c] = c;
      return;
    }
    double R = boundingRadius();
    auto lo = [&](double c, double o, double dlt) { return (int)std::floor((c - R - o) / dlt) - 1; };
    auto hi = [&](double c, double o, double dlt) { return (int)std::floor((c + R - o) / dlt) + 1; };
    int i0 = std::max(0, lo(center.x, m.origin.x, m.delta.x)), i1 = std::min(m.nx - 1, hi(center.x, m.origin.x, m.delta.x));
    int j0 = std::max(0, lo(center.y, m.origin.y, m.delta.y)), j1 = std::min(m.ny - 1, hi(center.y, m.origin.y, m.delta.y));
    int k0 = std::max(0, lo(center.z, m.origin.z, m.delta.z) - m.kz0), k1 = std::min(m.nz - 1, hi(center.z, m.origin.z, m.delta.z) - m.kz0);
    for (int K = k0; K <= k1; ++K)
      for (int J = j0; J <= j1; ++J)
        for (int I = i0; I <= i1; ++I) out.push_back(m.latticeToCell[m.lat(I, J, K)]);
    std::sort(out.begin(), out.end());
  }

  // sphere.C:198-243, cylinder.C:155-196, ellipse.C:153-204
  void updateCellLists() {
    const Mesh& m = *pMesh;
    shellCells.clear();
    neighbourPoints.clear();
    std::vector<int> cand;
    candidateCells(cand);
    if (kind == SPHERE) {
      double R = d_ / 2.0;
      double innerR = R - delta;
      for (int celli : cand) {
        vec3 diff = m.C[celli] - center;  // [B#1] base-class centre
        double r = mag(diff);
        if (r >= innerR && r <= R) {
          double th = std::atan2(diff.y, diff.x);
          double phi = std::acos(diff.z / r);
          if (th < 0) th += 2.0 * PI;
          if (phi < 0) phi += 2.0 * PI;
          int j = int(th * nTheta / (2.0 * PI));
          int k = int(phi * nk / PI);
          if (j >= nTheta) j = 0;        // [B#3] DEFINED clamp of the out-of-range bin
          if (k >= nk) k = nk - 1;       // [B#3]
          shellCells.push_back(celli);
          neighbourPoints.push_back({0, j, k});
        }
      }
    } else if (kind == CYLINDER) {
      double R = d_ / 2.0;
      double innerR = R - delta;
      for (int celli : cand) {
        vec3 diff = m.C[celli] - center;
        double r = mag(diff);
        if (r >= innerR && r <= R) {
          double th = std::atan2(diff.y, diff.x);
          if (th < 0) th += TWOPI;
          int j = int(th * (nTheta - 1) / TWOPI);  // [B#4] as coded
          shellCells.push_back(celli);
          neighbourPoints.push_back({0, j, 0});
        }
      }
    } else {
      if (theta.z < 0) theta.z += TWOPI;
      if (theta.z > TWOPI) theta.z -= TWOPI;
      for (int celli : cand) {
        const vec3& CC = m.C[celli];
        vec3 diff = CC - center;
        double radius = mag(diff);
        double R = r(CC);
        double innerR = R - delta;
        if (radius >= innerR && radius <= R) {
          double th = std::atan2(diff.y, diff.x) - theta.z;
          while (th < 0) th += TWOPI;
          while (th > TWOPI) th -= TWOPI;
          int j = int(th * (nTheta - 1) / TWOPI);
          shellCells.push_back(celli);
          neighbourPoints.push_back({0, j, 0});
        }
      }
    }
    setNeighbours();
    setWeights();
  }

  // particleShape.C:136-183
  void setNeighbours() {
    Is.assign(shellCells.size(), std::array<int, 8>{0, 0, 0, 0, 0, 0, 0, 0});
    Os.assign(shellCells.size(), std::array<int, 4>{0, 0, 0, 0});
    for (size_t celli = 0; celli < shellCells.size(); ++celli) {
      int j = neighbourPoints[celli].y;
      int k = neighbourPoints[celli].z;
      int jp1 = ((j + 1) % nTheta);
      if (nk != 1) {
        if (k == nk - 1) continue;  // [B#3] as coded: row left all-zero
        Is[celli] = {index(0, j, k), index(0, jp1, k), index(0, j, k + 1), index(0, jp1, k + 1),
                     index(1, j, k), index(1, jp1, k), index(1, j, k + 1), index(1, jp1, k + 1)};
        Os[celli] = {index(2, j, k), index(2, jp1, k), index(2, j, k + 1), index(2, jp1, k + 1)};
      } else {
        Is[celli] = {index(0, j, k), index(0, jp1, k), index(0, j, k), index(0, jp1, k),
                     index(1, j, k), index(1, jp1, k), index(1, j, k), index(1, jp1, k)};
        Os[celli] = {index(2, j, k), index(2, jp1, k), index(2, j, k), index(2, jp1, k)};
      }
    }
  }

  // particleShape.C:36-134
  void setWeights() {
    const Mesh& m = *pMesh;
    int nPtsOnMesh = 0;
    for (int pti = 0; pti < N; ++pti) {
      // [B#8] DEFINED: tables cleared on every update (the reference keeps stale entries)
      if (!mode.staleTables) {
        wToLocal[pti].clear();
        facesToLocal[pti].clear();
        WToLocal[pti] = 0.0;
      }
      vec3 pt = mesh[pti];
      int celli = findCell(pt);
      pointCell[pti] = celli;
      if (celli != -1) {
        double W = 0.0;
        std::vector<double> w;
        std::vector<int> faces;
        for (int q = m.cellFaceStart[celli]; q < m.cellFaceStart[celli + 1]; ++q) {
          int f = m.cellFaces[q];  // internal faces only (particleShape.C:59)
          double diff = mag(m.Cf[f] - pt);
          if (diff < SMALL) {
            for (auto& x : w) x = 0.0;
            w.push_back(1.0);
            faces.push_back(f);
            W = 1.0;
            break;
          }
          faces.push_back(f);
          w.push_back(1.0 / diff);
          W += w.back();
        }
        wToLocal[pti] = w;
        WToLocal[pti] = W;
        facesToLocal[pti] = faces;
        nPtsOnMesh++;
      }
    }
    wFromLocal.assign(shellCells.size(), std::array<double, 8>{0, 0, 0, 0, 0, 0, 0, 0});
    WFromLocal.assign(shellCells.size(), 0.0);
    if (nPtsOnMesh == 0 && findCell(center) == -1) onMesh = OFF_MESH;
    else if (nPtsOnMesh == N) onMesh = ON_MESH;
    else onMesh = PARTIAL;
    for (size_t celli = 0; celli < shellCells.size(); ++celli) {
      vec3 CC = m.C[shellCells[celli]];
      double W = 0.0;
      std::array<double, 8> w{0, 0, 0, 0, 0, 0, 0, 0};
      for (int pti = 0; pti < 8; ++pti) {
        double diff = mag(CC - mesh[Is[celli][pti]]);
        if (diff < SMALL) {
          w = {0, 0, 0, 0, 0, 0, 0, 0};
          w[pti] = 1.0;
          W = 1.0;
          break;
        }
        w[pti] = 1.0 / diff;
        W += w[pti];
      }
      wFromLocal[celli] = w;
      WFromLocal[celli] = W;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// particleIBM   src/particles/particles/particleIBM/
// ---------------------------------------------------------------------------------------------
struct Particle {
  int index = 0, copy = 0;
  bool active = true;
  Shape shape;
  vec3 v{0, 0, 0}, omega{0, 0, 0};
  double rho = 0, age = 0;
  vec3 integratedForce{0, 0, 0}, integratedTorque{0, 0, 0};
  // dense N-by-N / N-by-nWalls tables in the reference (particleIBMI.H:191-197); stored sparse
  // here (ordered by partner index) — sum() over the ordered keys equals the dense sum.  [B#12]
  std::map<int, vec3> Fc, Tc;
  std::vector<vec3> Fw, Tw;
  bool updatedForce = false;

  bool onMesh() const { return shape.onMesh != OFF_MESH; }  // particleIBMI.H:35-38
  const vec3& center() const { return shape.center; }
  double mass() const { return shape.V() * rho; }           // particleIBMI.H:96-99
  double r(const vec3& x2) const { return shape.r(x2); }

  // particleIBM.C:269-315
  void constructFromDict(const Mesh* m, SearchMode sm, const ParticleDict& dict, int idx) {
    index = idx;
    copy = 0;
    active = true;
    shape.constructFromDict(m, sm, dict.shape, dict.position);
    v = dict.v;
    omega = dict.omega;
    rho = dict.rho;
    age = 0.0;
    integratedForce = Zero3;
    integratedTorque = Zero3;
    active = (shape.findCell(shape.center) != -1);
  }
  // particleIBM.C:318-350
  void constructCopy(const Particle& p, int cpy, const vec3& c, const vec3& th, const vec3& vv, const vec3& om,
                     double ag) {
    index = p.index;
    copy = cpy;
    active = false;
    shape.constructCopy(p.shape, c, th);
    v = vv;
    omega = om;
    rho = p.rho;
    age = ag;
    integratedForce = p.integratedForce;
    integratedTorque = p.integratedTorque;
    Fc.clear();
    Tc.clear();
    Fw.assign(p.Fw.size(), Zero3);
    Tw.assign(p.Tw.size(), Zero3);
  }
  void setCollisions(int /*nParticles*/, int nWalls) {  // particleIBMI.H:191-197
    Fc.clear();
    Tc.clear();
    Fw.assign(nWalls, Zero3);
    Tw.assign(nWalls, Zero3);
  }

  // particleIBM.C:387-390   (-(pt - center) ^ omega) + v
  vec3 vAt(const vec3& pt) const { return cross(-(pt - shape.center), omega) + v; }

  // particleIBMTmp.C:26-50
  template <class T>
  void interpolateFromMesh(const std::vector<T>& fieldF, std::vector<T>& field) const {
    if (!onMesh()) return;
    for (int pti = 0; pti < shape.N; ++pti) {
      const auto& faces = shape.facesToLocal[pti];
      for (size_t facei = 0; facei < faces.size(); ++facei)
        field[pti] += shape.wToLocal[pti][facei] / shape.WToLocal[pti] * fieldF[faces[facei]];
    }
  }

  // particleIBM.C:210-264
  void interpolateFToMesh(const std::vector<vec3>& U, const std::vector<vec3>& Uold, const std::vector<vec3>& S,
                          double dT, std::vector<vec3>& F) const {
    if (!onMesh()) return;
    const auto& Is = shape.Is;
    const auto& Os = shape.Os;
    const auto& ws = shape.wFromLocal;
    const auto& W = shape.WFromLocal;
    for (size_t celli = 0; celli < shape.shellCells.size(); ++celli) {
      int i = shape.shellCells[celli];
      vec3 Fi = Zero3;
      for (int pti = 0; pti < 4; pti++)  // [B#7] as coded
        Fi += ws[celli][pti] / W[celli] *
              ((vAt(shape.mesh[Os[celli][pti]]) - U[Os[celli][pti]] - Uold[Is[celli][pti]]) / dT + S[Os[celli][pti]]);
      for (int pti = 4; pti < 8; pti++)
        Fi += ws[celli][pti] / W[celli] *
              ((vAt(shape.mesh[Is[celli][pti]]) - Uold[Is[celli][pti]]) / dT + S[Is[celli][pti]]);
      F[i] += Fi;
    }
  }

  // particleIBM.C:520-539
  void forcing(const std::vector<vec3>& Uf, const std::vector<vec3>& Ufold, const std::vector<vec3>& Sf, double dT,
               std::vector<vec3>& F) const {
    int N = shape.N;
    std::vector<vec3> iU(N, Zero3), iUold(N, Zero3), iS(N, Zero3);
    interpolateFromMesh(Uf, iU);
    interpolateFromMesh(Ufold, iUold);
    interpolateFromMesh(Sf, iS);
    interpolateFToMesh(iU, iUold, iS, dT, F);
  }

  // particleIBM.C:541-581
  void integrateSurfaceStress(const std::vector<symm6>& tauf, const std::vector<double>& pf,
                              vec3* correctedTorque = nullptr) {
    std::vector<symm6> iTau(shape.N, ZeroS);
    std::vector<double> iP(shape.N, 0.0);
    interpolateFromMesh(tauf, iTau);
    interpolateFromMesh(pf, iP);
    std::vector<vec3> Sf(shape.Sf.size());
    for (size_t n = 0; n < Sf.size(); ++n) Sf[n] = dot(shape.rotationMatrix, shape.Sf[n]);  // particleShapeI.H:35-38
    integratedForce = Zero3;
    integratedTorque = Zero3;
    vec3 Tcorr = Zero3;
    for (int i = 0; i < shape.nTheta; i++)
      for (int k = 0; k < shape.nk; k++) {
        vec3 R = shape.mesh[shape.index2(i, k)] - shape.center;  // [B#5] as coded (index2 into mesh_)
        const vec3& Sfi = Sf[shape.index2(i, k)];
        double magSf = mag(Sfi);
        vec3 norm = Sfi / magSf;
        const symm6& tau = iTau[shape.index(1, i, k)];
        vec3 F = iP[shape.index(1, i, k)] * shape.Sf[shape.index2(i, k)] + dot(tau, Sfi);  // [B#6] as coded
        vec3 Fn = dot(F, norm) * norm;
        vec3 Ft = F - Fn;
        integratedForce += F;
        integratedTorque -= cross(Ft, R);
        Tcorr -= cross(Ft, shape.mesh[shape.index(1, i, k)] - shape.center);
      }
    if (correctedTorque) *correctedTorque = Tcorr;
  }

  // particleIBM.C:583-600
  double Cd(double rhoRef, const vec3& UInf) const {
    vec3 normUInf = UInf / mag(UInf);
    return dot(normUInf, integratedForce) * 2.0 / (rhoRef * magSqr(UInf) * shape.A());
  }

  // particleIBM.C:393-450
  void solve(double dt, const vec3& g, bool moving, bool rotation) {
    if (!onMesh()) return;
    if (moving) {
      vec3 sFc = Zero3, sFw = Zero3;
      for (const auto& kv : Fc) sFc += kv.second;
      for (const auto& f : Fw) sFw += f;
      vec3 F = integratedForce + sFc + sFw;
      v += dt * (F / mass() + g);
      vec3 dx = dt * v;
      shape.center += dx;
      active = (shape.findCell(shape.center) != -1);
    }
    if (rotation) {
      vec3 sTc = Zero3, sTw = Zero3;
      for (const auto& kv : Tc) sTc += kv.second;
      for (const auto& t : Tw) sTw += t;
      vec3 T = integratedTorque + sTc + sTw;
      omega += dt * T / (mass() * shape.momentOfInertia);
      shape.theta += omega * dt;
      if (shape.theta.z < 0) shape.theta.z += TWOPI;
      if (shape.theta.z >= TWOPI) shape.theta.z -= TWOPI;
    }
    Fc.clear();
    Tc.clear();
    for (auto& f : Fw) f = Zero3;
    for (auto& t : Tw) t = Zero3;
  }

  // particleIBM.C:453-488   slot = index into the wall list, geometry = that wall's patch [B#11]
  void wallHit(const Patch& patch, double dt, int slot) {
    if (!onMesh()) return;
    for (int facei = 0; facei < patch.size(); ++facei) {
      const vec3& faceCentre = patch.Cf[facei];
      vec3 norm = patch.Sf[facei] / patch.magSf[facei];
      vec3 rv = vAt(faceCentre);
      vec3 R = center() - faceCentre;
      if (mag(R) <= r(faceCentre) && dot(rv, norm) > 0) {
        vec3 vNorm = norm * dot(norm, rv);
        vec3 rHat = R / mag(R);
        vec3 F = -2.0 * mass() * vNorm / dt;
        Fw[slot] = dot(F, rHat) * rHat;
        Tw[slot] = cross(F - Fw[slot], R);
        return;  // only one hit per patch
      }
    }
  }

  // particleIBM.C:490-517
  int patchHit(const Patch& patch, vec3& R) const {
    if (!onMesh()) return -1;
    for (int facei = 0; facei < patch.size(); ++facei) {
      const vec3& faceCentre = patch.Cf[facei];
      vec3 norm = patch.Sf[facei] / patch.magSf[facei];
      R = center() - faceCentre;
      if (mag(R) <= r(faceCentre) && dot(v, norm) > 0) return facei;
    }
    return -1;
  }

  // particleIBM.C:602-620
  void update() {
    shape.moveMesh(shape.center);
    shape.updateCellLists();
    active = (shape.findCell(shape.center) != -1);
  }

  // particleIBM.C:41-205 (static)
  static bool collision(const Particle& p1, const Particle& p2, vec3& collisionPt) {
    double pi = PI;
    vec3 center1 = p1.center();
    vec3 center2 = p2.center();
    vec3 diff = center2 - center1;
    vec3 d1 = p1.shape.D();
    vec3 d2 = p2.shape.D();
    if (mag(diff) > 0.5 * (cmptMax(d1) + cmptMax(d2))) return false;
    if ((d1.x + d1.y + d1.z) / 3.0 == cmptMax(d1) && (d2.x + d2.y + d2.z) / 3.0 == cmptMax(d2)) {  // [B#21]
      return 0.5 * (cmptMax(d1) + cmptMax(d2)) > mag(diff);
    }
    vec3 pt1 = center1, pt2 = center2;
    tensor R1 = p1.shape.rotationMatrix, R2 = p2.shape.rotationMatrix;
    double gamma1 = std::atan2(diff.y, diff.x) - p1.shape.theta.z;
    double gamma2 = std::atan2(diff.y, diff.x) - p2.shape.theta.z;
    if (center2.x < center1.x) gamma2 -= pi; else gamma1 -= pi;
    double r1 = p1.r(pt2), r2 = p2.r(pt1);
    double x1 = r1 * std::cos(gamma1), x2 = r2 * std::cos(gamma2);
    if (std::tan(p1.shape.theta.z) * std::tan(p2.shape.theta.z) < 0) { x1 *= -1.0; x2 *= -1.0; }
    bool converged = false, hit = false;
    int itt = 0;
    while (!converged) {
      itt++;
      pt1 = dot(R1, vec3{x1, 0, 0}) + center1;
      pt2 = dot(R2, vec3{x2, 0, 0}) + center2;
      vec3 oldDiff = diff;
      diff = pt2 - pt1;
      double alpha1 = std::atan2(diff.y, diff.x) - p1.shape.theta.z;
      double alpha2 = std::atan2(diff.y, diff.x) - p2.shape.theta.z;
      if (center2.x < center1.x) alpha2 -= pi; else alpha1 -= pi;
      vec3 rp1 = pt1, rp1Local = rp1;
      {
        double m = std::tan(alpha1), I = -m * x1;
        double a = sqr(m) + sqr(d1.y / d1.x), b = 2.0 * m * I, c = sqr(I) - sqr(d1.y);
        double posRoot = (-b + std::sqrt(sqr(b) - 4.0 * a * c)) / (2.0 * a);
        double negRoot = (-b - std::sqrt(sqr(b) - 4.0 * a * c)) / (2.0 * a);
        vec3 rpPos = dot(R1, vec3{posRoot, I + posRoot * m, 0}) + center1;
        vec3 rpNeg = dot(R1, vec3{negRoot, I + negRoot * m, 0}) + center1;
        if (mag(rpPos - center2) < mag(rpNeg - center2)) { rp1 = rpPos; rp1Local = {posRoot, I + posRoot * m, 0}; }
        else { rp1 = rpNeg; rp1Local = {negRoot, I + negRoot * m, 0}; }
        gamma1 = std::atan2(rp1Local.y, rp1Local.x);
      }
      vec3 rp2 = pt2, rp2Local = rp2;
      {
        double m = std::tan(alpha2), I = -m * x2;
        double a = sqr(m) + sqr(d2.y / d2.x), b = 2.0 * m * I, c = sqr(I) - sqr(d2.y);
        double posRoot = (-b + std::sqrt(sqr(b) - 4.0 * a * c)) / (2.0 * a);
        double negRoot = (-b - std::sqrt(sqr(b) - 4.0 * a * c)) / (2.0 * a);
        vec3 rpPos = dot(R2, vec3{posRoot, I + posRoot * m, 0}) + center2;
        vec3 rpNeg = dot(R2, vec3{negRoot, I + negRoot * m, 0}) + center2;
        if (mag(rpPos - center1) < mag(rpNeg - center1)) { rp2 = rpPos; rp2Local = {posRoot, I + posRoot * m, 0}; }
        else { rp2 = rpNeg; rp2Local = {negRoot, I + negRoot * m, 0}; }
        gamma2 = std::atan2(rp2Local.y, rp2Local.x);
      }
      r1 = mag(rp1Local);
      r2 = mag(rp2Local);
      x1 = r1 * std::cos(gamma1);
      x2 = r2 * std::cos(gamma2);
      if (std::tan(p1.shape.theta.z) * std::tan(p2.shape.theta.z) < 0) { x1 *= -1.0; x2 *= -1.0; }
      if (dot(diff, rp2 - rp1) < 0) {
        converged = true;
        hit = true;
        collisionPt = 0.5 * (rp1 + rp2);
      }
      if (mag(oldDiff) < mag(diff) || itt > 10) converged = true;
    }
    return hit;
  }
};

// ---------------------------------------------------------------------------------------------
// particleCloud   src/particleCloud/particleCloud.C
// ---------------------------------------------------------------------------------------------
struct CloudCfg {
  int moving = 0, rotating = 0, coupled = 1;
  double e = 1.0;
  vec3 UInf{0, 0, 0};
  double rhoRef = 1.0, pRef = 0.0;
  vec3 g{0, 0, 0};
  int faithfulSearch = 0;
  int cyclicWithoutWalls = 0;  // [B#10] 0 = as coded (computeCyclicHits returns when no wall patch exists)
};

struct SolveStats {
  int nSubSteps = 0;
  int nInstances = 0;
  vec3 avgV{0, 0, 0}, avgOmega{0, 0, 0}, avgP{0, 0, 0};
  double totK = 0;
};

struct Cloud {
  const Mesh* mesh = nullptr;
  CloudCfg cfg;
  int nParticles = 0;
  std::vector<Particle> origParticles;
  std::list<Particle> particles;  // Cloud<particleIBM> iteration order = insertion order
  std::vector<int> patches, walls, maxCopies;
  std::vector<double> I;  // indicator field I.ibm

  // cfg.faithfulSearch: bit 0 = linear-scan searches, bit 1 = quirk B#8 as coded
  SearchMode sm() const { return SearchMode{(cfg.faithfulSearch & 1) != 0, (cfg.faithfulSearch & 2) != 0}; }

  // particleCloud.C:974-1134 (time == 0 branch)
  Cloud(const Mesh* m, const CloudCfg& c, const std::vector<ParticleDict>& dicts) : mesh(m), cfg(c) {
    nParticles = (int)dicts.size();
    maxCopies.assign(nParticles, -1);
    for (int patchi = 0; patchi < (int)mesh->patches.size(); ++patchi) {
      const Patch& p = mesh->patches[patchi];
      if ((p.type == CYCLIC || p.type == PROCESSOR) && p.size() != 0) patches.push_back(patchi);
      else if (p.type == WALL) walls.push_back(patchi);
    }
    origParticles.resize(nParticles);
    for (int i = 0; i < nParticles; ++i) {
      origParticles[i].constructFromDict(mesh, sm(), dicts[i], i);
      origParticles[i].setCollisions(nParticles, (int)walls.size());
    }
    for (int i = 0; i < nParticles; ++i) particles.push_back(origParticles[i]);  // identical construction
    deleteInActiveParticles();
    for (auto& p : particles) {
      p.setCollisions(nParticles, (int)walls.size());
      maxCopies[p.index] = 0;
    }
    I.assign(mesh->nCells, 0.0);
    updateI();
  }

  // particleCloud.C:940-950
  void deleteInActiveParticles() {
    for (auto it = particles.begin(); it != particles.end();) {
      if (!it->onMesh()) it = particles.erase(it);
      else ++it;
    }
  }

  // particleCloud.C:952-969
  void updateI() {
    std::fill(I.begin(), I.end(), 0.0);
    if (cfg.faithfulSearch & 1) {
      for (int celli = 0; celli < mesh->nCells; ++celli) {
        const vec3& pt = mesh->C[celli];
        for (auto& p : particles)
          if (p.r(pt) > mag(p.center() - pt)) I[celli] = 1.0;
      }
    } else {
      std::vector<int> cand;
      for (auto& p : particles) {
        p.shape.candidateCells(cand);
        for (int celli : cand) {
          const vec3& pt = mesh->C[celli];
          if (p.r(pt) > mag(p.center() - pt)) I[celli] = 1.0;
        }
      }
    }
  }

  template <class T>
  void interpolateToFaces(const T* cellField, std::vector<T>& faceField) const {  // fvc::interpolate, linear
    faceField.resize(mesh->nInternalFaces);
    for (int f = 0; f < mesh->nInternalFaces; ++f) {
      const T& P = cellField[mesh->owner[f]];
      const T& Nn = cellField[mesh->neighbour[f]];
      faceField[f] = mesh->lambda[f] * (P - Nn) + Nn;
    }
  }

  // particleCloud.C:1158-1208; returns the number of cell-particle updates (SURVEY §8d unit)
  long forcing(const vec3* U, const vec3* Uold, const vec3* S, double dT, vec3* Fout) {
    std::vector<vec3> F(mesh->nCells, Zero3);
    long updates = 0;
    if (cfg.coupled) {
      std::vector<vec3> Uf, Ufold, Sf;
      interpolateToFaces(U, Uf);
      interpolateToFaces(Uold, Ufold);
      interpolateToFaces(S, Sf);
      for (const auto& p : particles) {
        p.forcing(Uf, Ufold, Sf, dT, F);
        if (p.onMesh()) updates += (long)p.shape.shellCells.size();
      }
    }
    for (int c = 0; c < mesh->nCells; ++c) Fout[c] = F[c];
    return updates;
  }

  // particleCloud.C:1210-1229.  p = kinematic or static pressure cell field, devRhoReff = cell field
  // returned by particleCloud::devRhoReff() (:41-101), handed in by the caller.
  void integrateSurfaceStresses(const double* p, const symm6* devRhoReff) {
    double pRefOverRho = cfg.pRef / cfg.rhoRef;
    std::vector<symm6> tauf;
    interpolateToFaces(devRhoReff, tauf);
    std::vector<double> pf(mesh->nInternalFaces);
    for (int f = 0; f < mesh->nInternalFaces; ++f) {
      double P = p[mesh->owner[f]], Nn = p[mesh->neighbour[f]];
      pf[f] = (mesh->lambda[f] * (P - Nn) + Nn) - pRefOverRho;
    }
    for (auto& t : tauf) t = t * cfg.rhoRef;
    for (auto& x : pf) x = x * cfg.rhoRef;
    for (auto& pr : particles) pr.integrateSurfaceStress(tauf, pf);
  }

  // particleCloud.C:171-313 (serial branch)
  void combineForces() {
    for (auto& p : particles) p.updatedForce = false;
    for (auto& p : particles) {
      if (!p.updatedForce) {
        vec3 totF = Zero3, totT = Zero3;
        for (auto& p2 : particles)
          if (p2.index == p.index) { totF += p2.integratedForce; totT += p2.integratedTorque; }
        for (auto& p2 : particles)
          if (p2.index == p.index) { p2.integratedForce = totF; p2.integratedTorque = totT; p2.updatedForce = true; }
      }
    }
  }

  // particleCloud.C:316-444 (serial branch :411-443)
  void setPositions() {
    for (const auto& p : particles) {
      if (!p.active) {
        for (auto& p2 : particles)
          if (p2.index == p.index && p2.copy != p.copy) {
            p2.shape.theta = p.shape.theta;
            p2.v = p.v;
            p2.omega = p.omega;
          }
      }
    }
  }

  // particleCloud.C:447-479
  bool found(int index, int /*copy*/, const vec3& center) const {
    bool f = false;
    for (const auto& p : particles)
      if (index == p.index && mag(center - p.center()) < 1e-4) f = true;  // [B#15]
    return f;
  }

  // particleCloud.C:482-581 (serial part)
  void computeCollisions(double deltaT) {
    auto pairStep = [&](Particle& p1, Particle& p2) {
      const vec3& x1 = p1.center();
      const vec3& x2 = p2.center();
      double mass1 = p1.mass(), mass2 = p2.mass();
      vec3 x12 = x1 - x2;
      vec3 v12 = p1.v - p2.v;
      if (p1.index > p2.index && p1.onMesh() && p2.onMesh()) {
        vec3 hitPt = Zero3;
        if (Particle::collision(p1, p2, hitPt) && dot(x12, v12) < 0) {
          double M = mass1 * mass2 / (mass1 + mass2);
          double x12Dotv12 = dot(x12, v12);
          double x12Sqr = magSqr(x12);
          vec3 F = 2.0 * M * x12Dotv12 / x12Sqr * x12 * cfg.e / deltaT;
          p1.Fc[p2.index] = -F;
          p2.Fc[p1.index] = F;
          p1.Tc[p2.index] = cross(-F, hitPt);
          p2.Tc[p1.index] = cross(F, hitPt);
        }
      }
    };
    if ((cfg.faithfulSearch & 1) || particles.size() < 64) {
      for (auto& p1 : particles)
        for (auto& p2 : particles) pairStep(p1, p2);
    } else {
      // same pair visit order restricted to pairs that can pass the overlap early-out
      // (particleIBM.C:56-59): uniform bins of edge >= max diameter, 27-neighbourhood.
      double dmax = 0;
      for (auto& p : particles) dmax = std::fmax(dmax, cmptMax(p.shape.D()));
      std::vector<Particle*> ps;
      for (auto& p : particles) ps.push_back(&p);
      std::map<std::array<long, 3>, std::vector<int>> bins;
      auto key = [&](const vec3& c) {
        return std::array<long, 3>{(long)std::floor(c.x / dmax), (long)std::floor(c.y / dmax), (long)std::floor(c.z / dmax)};
      };
      for (int i = 0; i < (int)ps.size(); ++i) bins[key(ps[i]->center())].push_back(i);
      std::vector<int> cand;
      for (int i = 0; i < (int)ps.size(); ++i) {
        auto k0 = key(ps[i]->center());
        cand.clear();
        for (long a = -1; a <= 1; ++a)
          for (long b = -1; b <= 1; ++b)
            for (long c = -1; c <= 1; ++c) {
              auto it = bins.find({k0[0] + a, k0[1] + b, k0[2] + c});
              if (it != bins.end()) cand.insert(cand.end(), it->second.begin(), it->second.end());
            }
        std::sort(cand.begin(), cand.end());
        for (int j : cand) pairStep(*ps[i], *ps[j]);
      }
    }
    // copies of one index share the larger-magnitude entry (particleCloud.C:525-554)
    std::map<int, std::vector<Particle*>> groups;
    bool anyCopies = false;
    for (auto& p : particles) {
      auto& g = groups[p.index];
      g.push_back(&p);
      if (g.size() > 1) anyCopies = true;
    }
    if (anyCopies) {
      auto get = [](std::map<int, vec3>& m, int k) { auto it = m.find(k); return it == m.end() ? Zero3 : it->second; };
      for (auto& p : particles)
        for (auto& p2 : particles)
          if (p.index == p2.index && &p != &p2) {
            std::vector<int> keys;
            for (auto& kv : p.Fc) keys.push_back(kv.first);
            for (auto& kv : p2.Fc) keys.push_back(kv.first);
            for (int k : keys) {
              if (mag(get(p.Fc, k)) > mag(get(p2.Fc, k))) { p2.Fc[k] = get(p.Fc, k); p2.Tc[k] = get(p.Tc, k); }
              else { p.Fc[k] = get(p2.Fc, k); p.Tc[k] = get(p2.Tc, k); }
            }
          }
    }
  }

  // particleCloud.C:584-656 (serial part)
  void computeWallHits(double deltaT) {
    if (!walls.size()) return;
    for (auto& p : particles)
      for (int i = 0; i < (int)walls.size(); ++i) p.wallHit(mesh->patches[walls[i]], deltaT, i);  // [B#11]
    for (int i = 0; i < (int)walls.size(); ++i)
      for (auto& p : particles)
        for (auto& p2 : particles)
          if (p.index == p2.index) {
            if (mag(p.Fw[i]) > mag(p2.Fw[i])) { p2.Fw[i] = p.Fw[i]; p2.Tw[i] = p.Tw[i]; }
            else { p.Fw[i] = p2.Fw[i]; p.Tw[i] = p2.Tw[i]; }
          }
  }

  // particleCloud.C:659-937 (cyclic branch; processor branches need MPI ranks and are exercised by the
  // decomposed multi-GPU oracle instead)
  void computeCyclicHits() {
    if (!walls.size() && !cfg.cyclicWithoutWalls) return;  // [B#10]
    struct NewP { vec3 c, th, v, om; double age; int index, copy; };
    std::vector<NewP> newCyclic;
    for (int patchi : patches) {
      const Patch& patch = mesh->patches[patchi];
      if (patch.type != CYCLIC) continue;
      for (const auto& p : particles) {
        vec3 R = Zero3;
        int facei = p.patchHit(patch, R);
        if (facei != -1) {
          vec3 newCenter = R + mesh->patches[patch.neighbPatch].Cf[facei];
          if (!found(p.index, p.copy, newCenter))
            newCyclic.push_back({newCenter, p.shape.theta, p.v, p.omega, p.age, p.index, p.copy});
        }
      }
    }
    for (const auto& np : newCyclic) {
      particles.emplace_back();
      particles.back().constructCopy(origParticles[np.index], maxCopies[np.index] + 1, np.c, np.th, np.v, np.om, np.age);
      maxCopies[np.index]++;
    }
  }

  // sub-step + update part of particleCloud::solve (particleCloud.C:1247-1324)
  SolveStats advance(double dt) {
    SolveStats st;
    double localDt = dt / 100.0;
    double localTime = localDt;
    while (localTime < dt) {  // [B#22] literal loop
      if (cfg.moving) {
        computeCollisions(localDt);
        computeWallHits(localDt);
        computeCyclicHits();
      }
      for (auto& p : particles) p.solve(localDt, cfg.g, cfg.moving != 0, cfg.rotating != 0);
      setPositions();
      st.nSubSteps++;
      if (localTime + localDt > dt) localDt = dt - localTime;
      else localTime += localDt;
    }
    for (auto& p : particles) p.update();
    deleteInActiveParticles();
    updateI();  // [B#13] DEFINED: indicator refreshed with the cell lists
    st.nInstances = (int)particles.size();
    for (const auto& p : particles) {
      double act = p.active ? 1.0 : 0.0;
      st.avgOmega += p.omega * act;
      st.avgV += p.v * act;
      st.avgP += p.v * p.mass() * act;
      st.totK += 0.5 * p.mass() * dot(p.v, p.v) * act;
    }
    if (nParticles) {
      st.avgV = st.avgV / double(nParticles);
      st.avgOmega = st.avgOmega / double(nParticles);
      st.avgP = st.avgP / double(nParticles);
    }
    return st;
  }

  // particleCloud.C:1242-1324
  SolveStats solve(const double* p, const symm6* devRhoReff, double dt) {
    integrateSurfaceStresses(p, devRhoReff);
    combineForces();
    return advance(dt);
  }
};

}  // namespace orc
