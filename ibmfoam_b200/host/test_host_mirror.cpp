// Exercises the C++ host mirror the way pimpleIBMFoam uses the reference class (createFields.H:60,
// UEqn.H:20-26, pimpleIBMFoam.C:62,87): construct, maxCoNum, forcing, solve, iterate particles.
// Prints one line per check; exit code 0 iff all pass.  Needs a CUDA device.
#include <cmath>
#include <cstdio>

#include "particleCloud.H"

using namespace ibm;

int main() {
  latticeMesh mesh;
  mesh.nx = mesh.ny = mesh.nz = 32;
  mesh.boundary = {{IBM_CYCLIC, IBM_XMIN, 1}, {IBM_CYCLIC, IBM_XMAX, 0}, {IBM_CYCLIC, IBM_YMIN, 3},
                   {IBM_CYCLIC, IBM_YMAX, 2}, {IBM_WALL, IBM_ZMIN, -1}, {IBM_WALL, IBM_ZMAX, -1}};
  IBMProperties dict;
  dict.moving = dict.rotating = true;
  for (int i = 0; i < 2; ++i) {
    ibm_particle_init p{};
    p.kind = IBM_SPHERE;
    p.position[0] = 10.3 + 11.1 * i; p.position[1] = 15.7; p.position[2] = 16.2;
    p.d = 8.0; p.nTheta = 50; p.nk = 1; p.delta = 1.0; p.rho = 1.5;
    p.v[0] = i == 0 ? 0.5 : -0.5;
    dict.particles.push_back(p);
  }
  int fails = 0;
  auto expect = [&](bool ok, const char* what) { std::printf("%s %s\n", ok ? "ok  " : "FAIL", what); if (!ok) ++fails; };
  try {
    particleCloud particles(mesh, dict, vector{{0, 0, -9.81}});
    expect(particles.nParticles() == 2, "nParticles() == 2");
    expect(particles.maxCoNum() == 0.0, "maxCoNum() == 0 (as coded)");
    const size_t C = (size_t)mesh.nCells();
    vectorField U(C, vector{{0.3, -0.2, 0.1}}), S(C, vector{{0, 0, 0}});
    vectorField F = particles.forcing(S, U, U, 0.25);
    size_t nz = 0;
    for (auto& f : F) if (f[0] != 0 || f[1] != 0 || f[2] != 0) ++nz;
    size_t shell = 0;
    for (auto& p : particles.particles()) shell += (size_t)p.nShellCells();
    expect(nz > 0 && nz <= shell, "forcing() is non-zero exactly on shell cells");
    scalarField p(C, 1.0);
    symmTensorField tau(C, std::array<double, 6>{{0, 0, 0, 0, 0, 0}});
    particles.solve(p, tau, 0.25);
    expect(particles.lastSolveStats().nSubSteps == 100 || particles.lastSolveStats().nSubSteps == 101, "solve(): literal sub-step loop");
    auto ps = particles.particles();
    expect(ps[0].v()[2] < 0 && std::fabs(ps[0].v()[2] + 9.81 * 0.25) < 1e-6 * 9.81, "gravity integrated over deltaT");
    scalarField cd;
    dict.UInf = vector{{1, 0, 0}};
    particles.Cd(cd);
    expect(cd.size() == 2, "Cd() fills one value per particle");
    bool threw = false;
    try {
      IBMProperties bad = dict;
      bad.particles[0].kind = IBM_CYLINDER;  // cylinder on a 3-D mesh: FatalError in the reference (cylinder.C:64-70)
      particleCloud badCloud(mesh, bad, vector{{0, 0, 0}});
    } catch (const FatalError& e) { threw = true; }
    expect(threw, "cylinder on a 3-D mesh is a FatalError");
  } catch (const FatalError& e) {
    std::printf("FAIL FatalError: %s\n", e.what());
    return 2;
  }
  return fails ? 1 : 0;
}
