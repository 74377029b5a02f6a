// Drop-in demonstration (TEST INFRASTRUCTURE): the solver side of the boundary, written ONCE against the public
// interface of Foam::particleCloud, and compiled twice -
//   dropin_reference : with the reference's own src/particleCloud + src/particles sources (CPU),
//   dropin_shim      : with shim/particleCloud.{H,C} + libibmfoam_b200.so (B200),
// both against the OpenFOAM look-alike in oracle/mockfoam (OpenFOAM itself is not available here).  It makes the calls
// pimpleIBMFoam makes, in its order (applications/solvers/pimpleIBMFoam):
//   createFields.H:47-60   IOdictionary IBMProperties;  particleCloud particles(mesh, IBMProperties, g);
//   pimpleIBMFoam.C:62     CoNum = max(CoNum, particles.maxCoNum());
//   UEqn.H:20-26           UEqn -= particles.forcing(S)            (nOuterCorrectors times)
//   pimpleIBMFoam.C:87     particles.solve();
//   pimpleIBMFoam.C:89     runTime.write()  ->  particles.writeObject(...)
// with synthetic U, U.oldTime(), S, p and devRhoReff fields in place of the PIMPLE solution, and dumps what comes back.
//
//   dropin_* <case file> <output file>
// case file (text): see tests/test_zz_gpu_shim_dropin.py::write_case.
#include "particleCloud.H"

#include <cstdlib>
#include <cstdio>
#include <fstream>

using namespace Foam;

#ifdef DROPIN_REFERENCE
namespace Foam { defineTypeNameAndDebug(particle, 0); }
#endif

static void dump(std::FILE* f, const char* tag, const std::vector<double>& v)
{
    std::fprintf(f, "%s %zu\n", tag, v.size());
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fprintf(f, "\n");
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s <case> <out>\n", argv[0]); return 2; }
    std::ifstream in(argv[1]);
    int nx, ny, nz, nPatches, nPart, nSteps, nOuter, moving, rotating;
    double origin[3], delta[3], dt, e, g[3], UInf[3], rhoInf, pInf;
    in >> nx >> ny >> nz >> origin[0] >> origin[1] >> origin[2] >> delta[0] >> delta[1] >> delta[2] >> nPatches;
    std::vector<int> side(nPatches), type(nPatches), nb(nPatches);
    for (int p = 0; p < nPatches; ++p) in >> side[p] >> type[p] >> nb[p];
    in >> nSteps >> nOuter >> dt >> moving >> rotating >> e >> g[0] >> g[1] >> g[2] >> UInf[0] >> UInf[1] >> UInf[2] >> rhoInf >> pInf >> nPart;
    static const char* sideName[6] = {"xMin", "xMax", "yMin", "yMax", "zMin", "zMax"};
    static const char* kindName[3] = {"sphere", "cylinder", "ellipse"};

    try
    {
        // ---- the case: mesh, time, constant/IBMProperties ----
        std::unique_ptr<orc::Mesh> mini(orc::buildMesh(nx, ny, nz, {origin[0], origin[1], origin[2]}, {delta[0], delta[1], delta[2]}, {},
                                                       side, type, nb, 0));
        Time runTime;
        runTime.setTime(0);
        runTime.setDeltaT(dt);
        fvMesh mesh(*mini, runTime, false);
        IOdictionary IBMProperties("IBMProperties", mesh);
        IBMProperties.add("nParticles", label(nPart));
        IBMProperties.add("moving", bool(moving));
        IBMProperties.add("rotating", bool(rotating));
        IBMProperties.add("coupled", true);
        IBMProperties.add("e", scalar(e));
        dictionary& flow = IBMProperties.addDict("flow");
        flow.add("UInf", vector(UInf[0], UInf[1], UInf[2]));
        flow.add("rhoInf", scalar(rhoInf));
        flow.add("pInf", scalar(pInf));
        if (!std::getenv("DROPIN_NO_LATTICE"))
        {
            // optional entry, read by the shim only (the numbers of system/blockMeshDict); without it the shim hands polyMesh's
            // cell and face centres to ibm_lattice_from_arrays, which is what the comparison test does as well
            dictionary& lat = IBMProperties.addDict("lattice");
            lat.add("n", vector(nx, ny, nz));
            lat.add("origin", vector(origin[0], origin[1], origin[2]));
            lat.add("delta", vector(delta[0], delta[1], delta[2]));
            std::vector<std::string> sides;
            for (int p = 0; p < nPatches; ++p) sides.push_back(sideName[side[p]]);
            lat.add("patchSides", sides);
        }
        for (int i = 0; i < nPart; ++i)
        {
            int kind, nTheta;
            double pos[3], d, a, b, dl, rho, th[3], v[3], om[3];
            in >> kind >> pos[0] >> pos[1] >> pos[2] >> d >> a >> b >> nTheta >> dl >> rho >> th[0] >> th[1] >> th[2] >> v[0] >> v[1] >> v[2] >>
                om[0] >> om[1] >> om[2];
            dictionary& pd = IBMProperties.addDict(IOobject::groupName("particle", Foam::name(i)));
            pd.add("particleShape", kindName[kind]);
            pd.add("position", vector(pos[0], pos[1], pos[2]));
            if (kind == 2) { pd.add("a", scalar(a)); pd.add("b", scalar(b)); }
            else pd.add("d", scalar(d));
            pd.add("nTheta", label(nTheta));
            pd.add("delta", scalar(dl));
            pd.add("rho", scalar(rho));
            pd.add("theta", vector(th[0], th[1], th[2]));
            pd.add("v", vector(v[0], v[1], v[2]));
            pd.add("omega", vector(om[0], om[1], om[2]));
        }
        if (!in) { std::fprintf(stderr, "bad case file\n"); return 2; }
        int nRestart = 0;   // optional trailing number: steps to run with a SECOND cloud constructed from the written time directory
        if (!(in >> nRestart)) nRestart = 0;

        // ---- the fields the solver owns ----
        const word t0 = runTime.timeName();
        volVectorField U(IOobject("U", t0, mesh), mesh, dimensionedVector("U", dimVelocity, Zero));
        volVectorField Uold(IOobject("U_0", t0, mesh), mesh, dimensionedVector("U", dimVelocity, Zero));
        U.setOldTime(&Uold);
        volVectorField S(IOobject("S", t0, mesh, IOobject::NO_READ, IOobject::NO_WRITE, false), mesh, dimensionedVector("S", dimAcceleration, Zero));
        volScalarField p(IOobject("p", t0, mesh), mesh, dimensionedScalar("p", dimPressure/dimDensity, 0.0));
        volSymmTensorField devRhoReff(IOobject("devRhoReff", t0, mesh, IOobject::NO_READ, IOobject::NO_WRITE, false), mesh,
                                      dimensioned<symmTensor>("tau", dimPressure, symmTensor(Zero)));
        compressible::turbulenceModel turbulence(mesh, devRhoReff);   // what particleCloud::devRhoReff() finds first
        const vectorField& C = mesh.cellCentres();
        const double Lx = nx*delta[0], Ly = ny*delta[1], Lz = nz*delta[2];
        const double kx = 6.283185307179586/Lx, ky = 6.283185307179586/Ly, kz = 6.283185307179586/Lz;
        auto fill = [&](const double t)
        {
            forAll(C, i)
            {
                const double x = kx*(C[i].x() - origin[0]), y = ky*(C[i].y() - origin[1]), z = (nz > 1) ? kz*(C[i].z() - origin[2]) : 0.0;
                U[i] = vector(std::sin(x + t)*std::cos(y)*std::cos(z), -std::cos(x + t)*std::sin(y)*std::cos(z), nz > 1 ? 0.1*std::sin(z + t) : 0.0);
                S[i] = vector(0.3*std::sin(2*x)*std::cos(z), 0.3*std::sin(2*y + t), nz > 1 ? -0.2*std::sin(2*z) : 0.0);
                p[i] = 0.0625*(std::cos(2*x + t) + std::cos(2*y))*(std::cos(2*z) + 2.0);
                devRhoReff[i] = symmTensor(-0.02*std::cos(x)*std::cos(y), 0.01*std::sin(x + t), nz > 1 ? 0.01*std::sin(z) : 0.0,
                                           0.02*std::cos(x)*std::cos(y), nz > 1 ? -0.01*std::cos(y)*std::sin(z) : 0.0, 0.0);
            }
        };

        // ---- createFields.H:60 ----
        std::unique_ptr<particleCloud> cloudPtr(new particleCloud(mesh, IBMProperties, dimensionedVector("g", dimAcceleration, vector(g[0], g[1], g[2]))));

        std::FILE* out = std::fopen(argv[2], "wb");
        std::fprintf(out, "dropin %d %d %d\n", mesh.nCells(), nSteps, nOuter);
        double CoNum = 0;
        fill(0.0);
        for (int step = 1; step <= nSteps + nRestart; ++step)
        {
            particleCloud& particles = *cloudPtr;
            static_cast<Field<vector>&>(Uold) = U;                   // U.oldTime() of the new step
            runTime.setTime(step*dt);
            fill(0.37*step);
            CoNum = max(CoNum, particles.maxCoNum());                // pimpleIBMFoam.C:62
            for (int outer = 0; outer < nOuter; ++outer)             // pimple.loop()
            {
                tmp<volVectorField> tF = particles.forcing(S);       // UEqn.H:20-26
                const volVectorField& F = tF();
                std::vector<double> f;
                forAll(F, i) { f.push_back(F[i].x()); f.push_back(F[i].y()); f.push_back(F[i].z()); }
                dump(out, "F", f);
            }
            particles.solve();                                       // pimpleIBMFoam.C:87
            IOField<scalar> Cds(particles.nParticles());
            particles.Cd(Cds);
            std::vector<double> cd;
            forAll(Cds, i) cd.push_back(Cds[i]);
            dump(out, "Cd", cd);
            if (step == nSteps && nRestart > 0)
            {
                // ---- restart (reference particleCloud.C:1119-1123): runTime.write(), the cloud goes away, a NEW cloud is constructed
                //      at a time != 0 and reads <time>/lagrangian/IBMparticles back (particleCloudIO.C:47-69) ----
                mockWrittenFields().clear();
                particles.writeObject(IOstream::ASCII, IOstream::versionNumber(), IOstream::UNCOMPRESSED, true);
                cloudPtr.reset();
                cloudPtr.reset(new particleCloud(mesh, IBMProperties, dimensionedVector("g", dimAcceleration, vector(g[0], g[1], g[2]))));
                dump(out, "restarted", {double(cloudPtr->nParticles())});
            }
        }
        particleCloud& particles = *cloudPtr;
        dump(out, "CoNum", {CoNum});
        // ---- runTime.write(): lagrangian/IBMparticles/* ----
        mockWrittenFields().clear();
        particles.writeObject(IOstream::ASCII, IOstream::versionNumber(), IOstream::UNCOMPRESSED, true);
        for (const char* name : {"index", "copy", "active", "center", "theta", "v", "omega", "rho", "age"})
        {
            auto it = mockWrittenFields().find(std::string(runTime.timeName()) + "/lagrangian/IBMparticles/" + name);
            if (it == mockWrittenFields().end()) { std::fprintf(stderr, "field %s was not written\n", name); return 3; }
            dump(out, name, it->second);
        }
        // ---- the per-particle class API (reference particleIBM.H:194-391): accessors read through the cloud's particles ----
        {
            std::vector<double> mass, vAt, rr, idx;
            const vector probe(origin[0] + 0.31*Lx, origin[1] + 0.47*Ly, origin[2] + (nz > 1 ? 0.53*Lz : 0.5*Lz));
#ifdef DROPIN_REFERENCE
            forAllConstIter(Cloud<particleIBM>, particles, iter)
            {
                const particleIBM& q = iter();
#else
            const List<particleIBM> views(particles.particles());
            forAll(views, vi)
            {
                const particleIBM& q = views[vi];
#endif
                idx.push_back(q.index());
                mass.push_back(q.mass());
                rr.push_back(q.r(probe));
                rr.push_back(q.d());
                const vector w(q.v(probe));
                vAt.push_back(w.x()); vAt.push_back(w.y()); vAt.push_back(w.z());
            }
            dump(out, "view_index", idx);
            dump(out, "view_mass", mass);
            dump(out, "view_r", rr);
            dump(out, "view_vAt", vAt);
        }
        std::fclose(out);
    }
    catch (const std::exception& ex)
    {
        std::fprintf(stderr, "%s\n", ex.what());
        return 1;
    }
    return 0;
}
