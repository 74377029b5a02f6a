// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// PARITY PINNED against the reference's own code: oracle/_ref (the reference's sources compiled against
// oracle/mockfoam) agrees with this restatement bit for bit (tests/test_oracle_vs_reference.py), and the vectors that
// build produced are committed (tests/golden/, tests/test_golden_oracle.py).  OpenFOAM's own arithmetic (mesh geometry,
// findCell, linear face interpolation) is restated, not pinned: see oracle/README.md.
// extern "C" surface of the CPU restatement, loaded with ctypes by tests/ and bench.py.
#include <chrono>
#include <cstring>

#include "orc_cloud.hpp"

using namespace orc;

extern "C" {

struct orc_particle_init {
  int kind;
  double position[3];
  double d, a, b;
  int nTheta, nk;
  double delta;
  double theta[3], v[3], omega[3];
  double rho;
};

struct orc_cloud_cfg {
  int moving, rotating, coupled;
  double e;
  double UInf[3];
  double rhoRef, pRef;
  double g[3];
  int faithfulSearch;
  int cyclicWithoutWalls;
};

struct orc_instance_state {
  int index, copy, active, onMesh;
  double center[3], theta[3], v[3], omega[3];
  double rho, age;
  double F[3], T[3];
  int nShell, N;
};

void* orc_mesh_create(int nx, int ny, int nz, const double* origin, const double* delta, int nBlocks,
                      const int* blocks, int nPatches, const int* patchSide, const int* patchType,
                      const int* patchNeighb, int kz0) {
  std::vector<BlockDesc> bl;
  for (int b = 0; b < nBlocks; ++b)
    bl.push_back({blocks[6 * b], blocks[6 * b + 1], blocks[6 * b + 2], blocks[6 * b + 3], blocks[6 * b + 4], blocks[6 * b + 5]});
  std::vector<int> ps(patchSide, patchSide + nPatches), pt(patchType, patchType + nPatches),
      pn(patchNeighb, patchNeighb + nPatches);
  return buildMesh(nx, ny, nz, {origin[0], origin[1], origin[2]}, {delta[0], delta[1], delta[2]}, bl, ps, pt, pn, kz0);
}
void orc_mesh_destroy(void* h) { delete (Mesh*)h; }
int orc_mesh_ncells(void* h) { return ((Mesh*)h)->nCells; }
int orc_mesh_ninternal_faces(void* h) { return ((Mesh*)h)->nInternalFaces; }
void orc_mesh_cell_centres(void* h, double* out) { std::memcpy(out, ((Mesh*)h)->C.data(), sizeof(vec3) * ((Mesh*)h)->nCells); }
void orc_mesh_lattice_to_cell(void* h, int* out) { std::memcpy(out, ((Mesh*)h)->latticeToCell.data(), sizeof(int) * ((Mesh*)h)->nCells); }
void orc_mesh_faces(void* h, int* owner, int* neighbour, double* Cf, double* lambda) {
  Mesh* m = (Mesh*)h;
  std::memcpy(owner, m->owner.data(), sizeof(int) * m->nInternalFaces);
  std::memcpy(neighbour, m->neighbour.data(), sizeof(int) * m->nInternalFaces);
  if (Cf) std::memcpy(Cf, m->Cf.data(), sizeof(vec3) * m->nInternalFaces);
  if (lambda) std::memcpy(lambda, m->lambda.data(), sizeof(double) * m->nInternalFaces);
}
int orc_mesh_patch_size(void* h, int patchi) { return ((Mesh*)h)->patches[patchi].size(); }
void orc_mesh_patch_faces(void* h, int patchi, double* Cf, double* Sf, double* magSf) {
  const Patch& p = ((Mesh*)h)->patches[patchi];
  std::memcpy(Cf, p.Cf.data(), sizeof(vec3) * p.size());
  std::memcpy(Sf, p.Sf.data(), sizeof(vec3) * p.size());
  std::memcpy(magSf, p.magSf.data(), sizeof(double) * p.size());
}
int orc_mesh_find_cell(void* h, const double* pt, int faithful) {
  Mesh* m = (Mesh*)h;
  vec3 p{pt[0], pt[1], pt[2]};
  return faithful ? m->findCellFaithful(p) : m->findCell(p);
}

void* orc_cloud_create(void* mesh, const orc_cloud_cfg* c, int n, const orc_particle_init* init) {
  Mesh* m = (Mesh*)mesh;
  CloudCfg cfg;
  cfg.moving = c->moving; cfg.rotating = c->rotating; cfg.coupled = c->coupled; cfg.e = c->e;
  cfg.UInf = {c->UInf[0], c->UInf[1], c->UInf[2]};
  cfg.rhoRef = c->rhoRef; cfg.pRef = c->pRef;
  cfg.g = {c->g[0], c->g[1], c->g[2]};
  cfg.faithfulSearch = c->faithfulSearch;
  cfg.cyclicWithoutWalls = c->cyclicWithoutWalls;
  std::vector<ParticleDict> dicts(n);
  for (int i = 0; i < n; ++i) {
    const auto& s = init[i];
    ParticleDict& d = dicts[i];
    d.shape.kind = s.kind; d.shape.d = s.d; d.shape.a = s.a; d.shape.b = s.b;
    d.shape.nTheta = s.nTheta; d.shape.nk = s.nk; d.shape.delta = s.delta;
    d.shape.theta = {s.theta[0], s.theta[1], s.theta[2]};
    d.position = {s.position[0], s.position[1], s.position[2]};
    d.v = {s.v[0], s.v[1], s.v[2]};
    d.omega = {s.omega[0], s.omega[1], s.omega[2]};
    d.rho = s.rho;
    // sphere.C:79-84, cylinder.C:64-70, ellipse.C:80-86 (FatalError in the reference)
    if (s.kind == SPHERE && m->nGeometricD != 3) return nullptr;
    if (s.kind != SPHERE && m->nGeometricD != 2) return nullptr;
  }
  return new Cloud(m, cfg, dicts);
}
void orc_cloud_destroy(void* h) { delete (Cloud*)h; }
int orc_cloud_n_instances(void* h) { return (int)((Cloud*)h)->particles.size(); }

static Particle& nth(Cloud* c, int i) {
  auto it = c->particles.begin();
  std::advance(it, i);
  return *it;
}

void orc_cloud_get_state(void* h, orc_instance_state* out) {
  Cloud* c = (Cloud*)h;
  int i = 0;
  for (const auto& p : c->particles) {
    auto& o = out[i++];
    o.index = p.index; o.copy = p.copy; o.active = p.active; o.onMesh = p.shape.onMesh;
    for (int a = 0; a < 3; ++a) {
      o.center[a] = p.shape.center[a]; o.theta[a] = p.shape.theta[a]; o.v[a] = p.v[a]; o.omega[a] = p.omega[a];
      o.F[a] = p.integratedForce[a]; o.T[a] = p.integratedTorque[a];
    }
    o.rho = p.rho; o.age = p.age;
    o.nShell = (int)p.shape.shellCells.size(); o.N = p.shape.N;
  }
}

void orc_cloud_set_force_torque(void* h, const double* FT) {
  Cloud* c = (Cloud*)h;
  int i = 0;
  for (auto& p : c->particles) {
    p.integratedForce = {FT[6 * i], FT[6 * i + 1], FT[6 * i + 2]};
    p.integratedTorque = {FT[6 * i + 3], FT[6 * i + 4], FT[6 * i + 5]};
    ++i;
  }
}

// shell tap: cells[nShell], jk[nShell][2], Is[nShell][8], Os[nShell][4], w[nShell][8], W[nShell]
void orc_cloud_get_shell(void* h, int inst, int* cells, int* jk, int* Is, int* Os, double* w, double* W) {
  const Shape& s = nth((Cloud*)h, inst).shape;
  for (size_t c = 0; c < s.shellCells.size(); ++c) {
    if (cells) cells[c] = s.shellCells[c];
    if (jk) { jk[2 * c] = s.neighbourPoints[c].y; jk[2 * c + 1] = s.neighbourPoints[c].z; }
    for (int q = 0; q < 8; ++q) {
      if (Is) Is[8 * c + q] = s.Is[c][q];
      if (w) w[8 * c + q] = s.wFromLocal[c][q];
    }
    for (int q = 0; q < 4; ++q)
      if (Os) Os[4 * c + q] = s.Os[c][q];
    if (W) W[c] = s.WFromLocal[c];
  }
}

// point tap: pts[N][3], cell[N], nFaces[N], otherCell[N][6] (neighbour cell across each listed
// face, in facesToLocal_ order), w[N][6], W[N]
void orc_cloud_get_points(void* h, int inst, double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W) {
  Cloud* c = (Cloud*)h;
  const Shape& s = nth(c, inst).shape;
  const Mesh& m = *c->mesh;
  for (int n = 0; n < s.N; ++n) {
    if (pts) { pts[3 * n] = s.mesh[n].x; pts[3 * n + 1] = s.mesh[n].y; pts[3 * n + 2] = s.mesh[n].z; }
    if (cell) cell[n] = s.pointCell[n];
    int nf = (int)s.facesToLocal[n].size();
    if (nFaces) nFaces[n] = nf;
    for (int q = 0; q < 6; ++q) {
      if (otherCell) {
        int oc = -1;
        if (q < nf) {
          int f = s.facesToLocal[n][q];
          oc = (m.owner[f] == s.pointCell[n]) ? m.neighbour[f] : m.owner[f];
        }
        otherCell[6 * n + q] = oc;
      }
      if (w) w[6 * n + q] = q < nf ? s.wToLocal[n][q] : 0.0;
    }
    if (W) W[n] = s.WToLocal[n];
  }
}

void orc_cloud_get_sf(void* h, int inst, double* Sf) {
  const Shape& s = nth((Cloud*)h, inst).shape;
  std::memcpy(Sf, s.Sf.data(), sizeof(vec3) * s.Sf.size());
}

void orc_cloud_indicator(void* h, double* I) {
  Cloud* c = (Cloud*)h;
  std::memcpy(I, c->I.data(), sizeof(double) * c->I.size());
}

long orc_cloud_forcing(void* h, const double* U, const double* Uold, const double* S, double dt, double* F) {
  return ((Cloud*)h)->forcing((const vec3*)U, (const vec3*)Uold, (const vec3*)S, dt, (vec3*)F);
}

void orc_cloud_integrate_stress(void* h, const double* p, const double* devRhoReff6, double* FT, double* Tcorrected) {
  Cloud* c = (Cloud*)h;
  c->integrateSurfaceStresses(p, (const symm6*)devRhoReff6);
  int i = 0;
  for (auto& pr : c->particles) {
    if (FT)
      for (int a = 0; a < 3; ++a) { FT[6 * i + a] = pr.integratedForce[a]; FT[6 * i + 3 + a] = pr.integratedTorque[a]; }
    ++i;
  }
  if (Tcorrected) {
    // [B#5] report: torque with the lever arm taken from the surface point index(1,i,k)
    double pRefOverRho = c->cfg.pRef / c->cfg.rhoRef;
    std::vector<symm6> tauf;
    c->interpolateToFaces((const symm6*)devRhoReff6, tauf);
    std::vector<double> pf(c->mesh->nInternalFaces);
    for (int f = 0; f < c->mesh->nInternalFaces; ++f) {
      double P = p[c->mesh->owner[f]], Nn = p[c->mesh->neighbour[f]];
      pf[f] = ((c->mesh->lambda[f] * (P - Nn) + Nn) - pRefOverRho) * c->cfg.rhoRef;
    }
    for (auto& t : tauf) t = t * c->cfg.rhoRef;
    i = 0;
    for (auto& pr : c->particles) {
      Particle tmp = pr;
      vec3 tc;
      tmp.integrateSurfaceStress(tauf, pf, &tc);
      for (int a = 0; a < 3; ++a) Tcorrected[3 * i + a] = tc[a];
      ++i;
    }
  }
}

void orc_cloud_combine_forces(void* h) { ((Cloud*)h)->combineForces(); }

// stats[0]=nSubSteps, [1]=nInstances, [2..4]=avgV, [5..7]=avgOmega, [8..10]=avgP, [11]=totK
void orc_cloud_advance(void* h, double dt, double* stats) {
  SolveStats st = ((Cloud*)h)->advance(dt);
  if (stats) {
    stats[0] = st.nSubSteps; stats[1] = st.nInstances;
    for (int a = 0; a < 3; ++a) { stats[2 + a] = st.avgV[a]; stats[5 + a] = st.avgOmega[a]; stats[8 + a] = st.avgP[a]; }
    stats[11] = st.totK;
  }
}

void orc_cloud_solve(void* h, const double* p, const double* devRhoReff6, double dt, double* stats) {
  Cloud* c = (Cloud*)h;
  c->integrateSurfaceStresses(p, (const symm6*)devRhoReff6);
  c->combineForces();
  orc_cloud_advance(h, dt, stats);
}

// one literal pass of the moving-particle sub-step body, for known-answer tests
void orc_cloud_substep(void* h, double localDt) {
  Cloud* c = (Cloud*)h;
  if (c->cfg.moving) {
    c->computeCollisions(localDt);
    c->computeWallHits(localDt);
    c->computeCyclicHits();
  }
  for (auto& p : c->particles) p.solve(localDt, c->cfg.g, c->cfg.moving != 0, c->cfg.rotating != 0);
  c->setPositions();
}

void orc_cloud_update(void* h) {
  Cloud* c = (Cloud*)h;
  for (auto& p : c->particles) p.update();
  c->deleteInActiveParticles();
  c->updateI();
}

void orc_cloud_cd(void* h, double* out) {
  Cloud* c = (Cloud*)h;
  int i = 0;
  for (const auto& p : c->particles) out[i++] = p.Cd(c->cfg.rhoRef, c->cfg.UInf);
}

double orc_cloud_max_co_num(void* /*h*/) { return 0.0; }  // [B#9] particleCloud.C:1145-1156 always returns 0

double orc_particle_mass(void* h, int inst) { return nth((Cloud*)h, inst).mass(); }
double orc_particle_inertia(void* h, int inst) { return nth((Cloud*)h, inst).shape.momentOfInertia; }

const char* orc_impl() { return "port"; }

// collision tables of one instance after computeCollisions/computeWallHits, expanded to the reference's dense layout
// (Fc/Tc [nParticles][3] pre-zeroed by the caller, Fw/Tw [nWalls][3])
void orc_particle_collision_tables(void* h, int inst, double* Fc, double* Tc, double* Fw, double* Tw) {
  const Particle& p = nth((Cloud*)h, inst);
  for (const auto& kv : p.Fc)
    for (int a = 0; a < 3; ++a) Fc[3 * kv.first + a] = kv.second[a];
  for (const auto& kv : p.Tc)
    for (int a = 0; a < 3; ++a) Tc[3 * kv.first + a] = kv.second[a];
  for (size_t i = 0; i < p.Fw.size(); ++i)
    for (int a = 0; a < 3; ++a) { Fw[3 * i + a] = p.Fw[i][a]; Tw[3 * i + a] = p.Tw[i][a]; }
}

// particleIBM::collision (static, particleIBM.C:41-205) on two instances
int orc_pair_collision(void* h, int inst1, int inst2, double* hitPt) {
  Cloud* c = (Cloud*)h;
  vec3 pt = Zero3;
  bool hit = Particle::collision(nth(c, inst1), nth(c, inst2), pt);
  for (int a = 0; a < 3; ++a) hitPt[a] = pt[a];
  return hit ? 1 : 0;
}

// sub-step schedule of particleCloud.C:1247-1275 for a given dt: writes up to maxN local dts, returns count
int orc_substep_schedule(double dt, double* dts, int maxN) {
  int n = 0;
  double localDt = dt / 100.0, localTime = localDt;
  while (localTime < dt) {
    if (n < maxN) dts[n] = localDt;
    ++n;
    if (localTime + localDt > dt) localDt = dt - localTime;
    else localTime += localDt;
  }
  return n;
}

}  // extern "C"
