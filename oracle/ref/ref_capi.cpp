// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C surface over the REFERENCE'S OWN CLASSES (Foam::particleCloud, Foam::particleIBM, Foam::particleShape and the
// sphere/cylinder/ellipse shapes), compiled from /root/reference/src by oracle/ref/Makefile against oracle/mockfoam.
// It exports the same orc_* entry points as oracle/ibm_oracle.cpp (the restatement), so that the same Python
// front-end (oracle/pyoracle.py) drives either library and tests/test_oracle_vs_reference.py can compare them bit
// for bit.  No algorithm lives here: every number this file returns was computed by reference code.
//
// Reading the reference's private tables (shell cells, neighbour ids, weights) and calling its private member
// functions one at a time needs access the classes do not give to outsiders: this ONE translation unit is compiled
// with g++ -fno-access-control (oracle/ref/Makefile).  The flag only switches off the access check; layout and
// mangled names are untouched, so the objects are the ones the reference's own translation units build.
#include "mockfoam.H"
#include "particleCloud.H"

#include <cstring>

using namespace Foam;

namespace Foam
{
defineTypeNameAndDebug(particle, 0);

}

namespace
{
thread_local std::string g_lastError;

struct RefMesh
{
    orc::Mesh* m;
};

struct RefCloud
{
    const orc::Mesh* mini;
    Time time;
    std::unique_ptr<fvMesh> mesh;
    std::unique_ptr<IOdictionary> dict;
    std::unique_ptr<volVectorField> U, Uold, S;
    std::unique_ptr<volScalarField> p;
    std::unique_ptr<volSymmTensorField> tau;
    std::unique_ptr<compressible::turbulenceModel> turb;
    std::unique_ptr<particleCloud> cloud;
    double rhoRef, pRef;
    vector UInf, g;
    bool moving, rotating;

    ~RefCloud()
    {
        cloud.reset();  // before the fields and the mesh it refers to
        turb.reset();
    }
};

particleIBM& nth(RefCloud* c, int i)
{
    Cloud<particleIBM>::iterator it = c->cloud->begin();
    for (int k = 0; k < i; ++k) ++it;
    return it();
}

template<class F>
int guarded(F&& f)
{
    try
    {
        f();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_lastError = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

struct orc_particle_init
{
    int kind;
    double position[3];
    double d, a, b;
    int nTheta, nk;
    double delta;
    double theta[3], v[3], omega[3];
    double rho;
};

struct orc_cloud_cfg
{
    int moving, rotating, coupled;
    double e;
    double UInf[3];
    double rhoRef, pRef;
    double g[3];
    int faithfulSearch;
    int cyclicWithoutWalls;
};

struct orc_instance_state
{
    int index, copy, active, onMesh;
    double center[3], theta[3], v[3], omega[3];
    double rho, age;
    double F[3], T[3];
    int nShell, N;
};

const char* orc_last_error() { return g_lastError.c_str(); }
const char* orc_impl() { return "reference"; }

void* orc_mesh_create(int nx, int ny, int nz, const double* origin, const double* delta, int nBlocks, const int* blocks,
                      int nPatches, const int* patchSide, const int* patchType, const int* patchNeighb, int kz0)
{
    std::vector<orc::BlockDesc> bl;
    for (int b = 0; b < nBlocks; ++b)
        bl.push_back({blocks[6 * b], blocks[6 * b + 1], blocks[6 * b + 2], blocks[6 * b + 3], blocks[6 * b + 4], blocks[6 * b + 5]});
    std::vector<int> ps(patchSide, patchSide + nPatches), pt(patchType, patchType + nPatches), pn(patchNeighb, patchNeighb + nPatches);
    RefMesh* r = new RefMesh;
    r->m = orc::buildMesh(nx, ny, nz, {origin[0], origin[1], origin[2]}, {delta[0], delta[1], delta[2]}, bl, ps, pt, pn, kz0);
    return r;
}
void orc_mesh_destroy(void* h)
{
    RefMesh* r = (RefMesh*)h;
    delete r->m;
    delete r;
}
int orc_mesh_ncells(void* h) { return ((RefMesh*)h)->m->nCells; }
int orc_mesh_ninternal_faces(void* h) { return ((RefMesh*)h)->m->nInternalFaces; }

void* orc_cloud_create(void* mesh, const orc_cloud_cfg* c, int n, const orc_particle_init* init)
{
    static const char* kindName[3] = {"sphere", "cylinder", "ellipse"};
    RefCloud* r = new RefCloud;
    const int rc = guarded([&]() {
        if (c->cyclicWithoutWalls)
            throw FatalErrorException("the reference returns from computeCyclicHits when there is no wall patch (particleCloud.C:661-664)");
        r->mini = ((RefMesh*)mesh)->m;
        r->time.setTime(0);
        r->mesh.reset(new fvMesh(*r->mini, r->time, (c->faithfulSearch & 1) != 0));  // bit 1 (B#8 as coded) is what the reference does anyway
        r->dict.reset(new IOdictionary("IBMProperties", *r->mesh));
        IOdictionary& d = *r->dict;
        d.add("nParticles", label(n));
        d.add("moving", bool(c->moving));
        d.add("rotating", bool(c->rotating));
        d.add("coupled", bool(c->coupled));
        d.add("e", scalar(c->e));
        dictionary& flow = d.addDict("flow");
        flow.add("UInf", vector(c->UInf[0], c->UInf[1], c->UInf[2]));
        flow.add("rhoInf", scalar(c->rhoRef));
        flow.add("pInf", scalar(c->pRef));
        for (int i = 0; i < n; ++i)
        {
            const orc_particle_init& s = init[i];
            dictionary& pd = d.addDict(IOobject::groupName("particle", Foam::name(i)));
            pd.add("particleShape", kindName[s.kind]);
            pd.add("position", vector(s.position[0], s.position[1], s.position[2]));
            if (s.kind == 2) { pd.add("a", scalar(s.a)); pd.add("b", scalar(s.b)); }
            else pd.add("d", scalar(s.d));
            pd.add("rho", scalar(s.rho));
            pd.add("nTheta", label(s.nTheta));
            pd.add("nk", label(s.nk));
            pd.add("delta", scalar(s.delta));
            pd.add("theta", vector(s.theta[0], s.theta[1], s.theta[2]));
            pd.add("v", vector(s.v[0], s.v[1], s.v[2]));
            pd.add("omega", vector(s.omega[0], s.omega[1], s.omega[2]));
        }
        const fvMesh& fm = *r->mesh;
        const word t0 = r->time.timeName();
        r->U.reset(new volVectorField(IOobject("U", t0, fm), fm, dimensionedVector("U", dimVelocity, Zero)));
        r->Uold.reset(new volVectorField(IOobject("U_0", t0, fm), fm, dimensionedVector("U", dimVelocity, Zero)));
        r->U->setOldTime(r->Uold.get());
        r->S.reset(new volVectorField(IOobject("S", t0, fm, IOobject::NO_READ, IOobject::NO_WRITE, false), fm,
                                      dimensionedVector("S", dimAcceleration, Zero)));
        r->p.reset(new volScalarField(IOobject("p", t0, fm), fm, dimensionedScalar("p", dimPressure / dimDensity, 0.0)));
        r->tau.reset(new volSymmTensorField(IOobject("devRhoReff", t0, fm, IOobject::NO_READ, IOobject::NO_WRITE, false), fm,
                                            dimensioned<symmTensor>("tau", dimPressure, symmTensor(Zero))));
        r->turb.reset(new compressible::turbulenceModel(fm, *r->tau));
        r->rhoRef = c->rhoRef;
        r->pRef = c->pRef;
        r->UInf = vector(c->UInf[0], c->UInf[1], c->UInf[2]);
        r->g = vector(c->g[0], c->g[1], c->g[2]);
        r->moving = c->moving;
        r->rotating = c->rotating;
        r->cloud.reset(new particleCloud(fm, d, dimensionedVector("g", dimAcceleration, r->g)));
    });
    if (rc)
    {
        delete r;
        return nullptr;
    }
    return r;
}
void orc_cloud_destroy(void* h) { delete (RefCloud*)h; }
int orc_cloud_n_instances(void* h) { return ((RefCloud*)h)->cloud->size(); }

void orc_cloud_get_state(void* h, orc_instance_state* out)
{
    RefCloud* c = (RefCloud*)h;
    int i = 0;
    forAllConstIter(Cloud<particleIBM>, *c->cloud, it)
    {
        const particleIBM& p = it();
        orc_instance_state& o = out[i++];
        o.index = p.index();
        o.copy = p.copy();
        o.active = p.active();
        o.onMesh = int(p.shape_->onMesh_);
        for (int a = 0; a < 3; ++a)
        {
            o.center[a] = p.center()[a];
            o.theta[a] = p.theta()[a];
            o.v[a] = p.v()[a];
            o.omega[a] = p.omega()[a];
            o.F[a] = p.F()[a];
            o.T[a] = p.T()[a];
        }
        o.rho = p.rho();
        o.age = p.age();
        o.nShell = p.shape_->shellCells_.size();
        o.N = p.shape_->N();
    }
}

void orc_cloud_set_force_torque(void* h, const double* FT)
{
    RefCloud* c = (RefCloud*)h;
    int i = 0;
    forAllIter(Cloud<particleIBM>, *c->cloud, it)
    {
        it().F() = vector(FT[6 * i], FT[6 * i + 1], FT[6 * i + 2]);
        it().T() = vector(FT[6 * i + 3], FT[6 * i + 4], FT[6 * i + 5]);
        ++i;
    }
}

void orc_cloud_get_shell(void* h, int inst, int* cells, int* jk, int* Is, int* Os, double* w, double* W)
{
    const particleShape& s = nth((RefCloud*)h, inst).shape_();
    forAll(s.shellCells_, c)
    {
        if (cells) cells[c] = s.shellCells_[c];
        if (jk) { jk[2 * c] = s.neighbourPoints_[c].y(); jk[2 * c + 1] = s.neighbourPoints_[c].z(); }
        for (int q = 0; q < 8; ++q)
        {
            if (Is) Is[8 * c + q] = s.Is_[c][q];
            if (w) w[8 * c + q] = s.wFromLocal_[c][q];
        }
        for (int q = 0; q < 4; ++q)
            if (Os) Os[4 * c + q] = s.Os_[c][q];
        if (W) W[c] = s.WFromLocal_[c];
    }
}

// cell[n] is what setWeights saw (findCell of the point); the lists are reported as the reference holds them, i.e. for a
// point that has left the mesh whatever an earlier update left there (quirk B#8)
void orc_cloud_get_points(void* h, int inst, double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W)
{
    RefCloud* c = (RefCloud*)h;
    const particleShape& s = nth(c, inst).shape_();
    const orc::Mesh& m = *c->mini;
    for (int n = 0; n < s.N(); ++n)
    {
        const vector& pt = s.mesh_[n];
        if (pts) { pts[3 * n] = pt.x(); pts[3 * n + 1] = pt.y(); pts[3 * n + 2] = pt.z(); }
        const label celli = c->mesh->findCell(pt);
        if (cell) cell[n] = celli;
        const int nf = s.facesToLocal_[n].size();
        if (nFaces) nFaces[n] = nf;
        for (int q = 0; q < 6; ++q)
        {
            if (otherCell)
            {
                int oc = -1;
                if (q < nf)
                {
                    const int f = s.facesToLocal_[n][q];
                    oc = (m.owner[size_t(f)] == celli) ? m.neighbour[size_t(f)] : m.owner[size_t(f)];
                }
                otherCell[6 * n + q] = oc;
            }
            if (w) w[6 * n + q] = q < nf ? s.wToLocal_[n][q] : 0.0;
        }
        if (W) W[n] = s.WToLocal_[n];
    }
}

void orc_cloud_get_sf(void* h, int inst, double* Sf)
{
    const particleShape& s = nth((RefCloud*)h, inst).shape_();
    forAll(s.Sf_, i)
        for (int a = 0; a < 3; ++a) Sf[3 * i + a] = s.Sf_[i][a];
}

void orc_cloud_indicator(void* h, double* I)
{
    RefCloud* c = (RefCloud*)h;
    forAll(c->cloud->I_, i) I[i] = c->cloud->I_[i];
}

static void load(volVectorField& f, const double* src)
{
    forAll(f, i) f[i] = vector(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

long orc_cloud_forcing(void* h, const double* U, const double* Uold, const double* S, double dt, double* F)
{
    RefCloud* c = (RefCloud*)h;
    long updates = -1;
    guarded([&]() {
        load(*c->U, U);
        load(*c->Uold, Uold);
        load(*c->S, S);
        c->time.setDeltaT(dt);
        tmp<volVectorField> tF = c->cloud->forcing(*c->S);
        const volVectorField& Fv = tF();
        forAll(Fv, i)
            for (int a = 0; a < 3; ++a) F[3 * i + a] = Fv[i][a];
        updates = 0;
        if (c->cloud->coupled_)
            forAllConstIter(Cloud<particleIBM>, *c->cloud, it)
                if (it().onMesh()) updates += it().shape_->shellCells_.size();
    });
    return updates;
}

static void loadStress(RefCloud* c, const double* p, const double* tau6)
{
    forAll(*c->p, i) (*c->p)[i] = p[i];
    forAll(*c->tau, i)
        (*c->tau)[i] = symmTensor(tau6[6 * i], tau6[6 * i + 1], tau6[6 * i + 2], tau6[6 * i + 3], tau6[6 * i + 4], tau6[6 * i + 5]);
}

static void readFT(RefCloud* c, double* FT)
{
    int i = 0;
    forAllConstIter(Cloud<particleIBM>, *c->cloud, it)
    {
        for (int a = 0; a < 3; ++a) { FT[6 * i + a] = it().F()[a]; FT[6 * i + 3 + a] = it().T()[a]; }
        ++i;
    }
}

// Tcorrected (quirk B#5 report) is the restatement's own diagnostic and has no reference counterpart: must be null
int orc_cloud_integrate_stress(void* h, const double* p, const double* devRhoReff6, double* FT, double* Tcorrected)
{
    RefCloud* c = (RefCloud*)h;
    if (Tcorrected) { g_lastError = "Tcorrected has no reference counterpart"; return 1; }
    return guarded([&]() {
        loadStress(c, p, devRhoReff6);
        c->cloud->integrateSurfaceStresses();
        if (FT) readFT(c, FT);
    });
}

void orc_cloud_combine_forces(void* h) { ((RefCloud*)h)->cloud->combineForces(); }

// particleCloud::solve() end to end.  The statistics are only ever printed by the reference (particleCloud.C:1317-1322):
// they are read back from that print-out at 17 significant digits.
// stats[0]=nSubSteps (not observable: -1), [1]=nInstances, [2..4]=avgV, [5..7]=avgOmega, [8..10]=avgP, [11]=totK
int orc_cloud_solve(void* h, const double* p, const double* devRhoReff6, double dt, double* stats)
{
    RefCloud* c = (RefCloud*)h;
    return guarded([&]() {
        loadStress(c, p, devRhoReff6);
        c->time.setDeltaT(dt);
        std::ostringstream log;
        log.precision(17);
        std::ostream* saved = mockInfoStream().stdStream();
        mockInfoStream().stdStream(&log);
        try { c->cloud->solve(); }
        catch (...) { mockInfoStream().stdStream(saved); throw; }
        mockInfoStream().stdStream(saved);
        if (stats)
        {
            for (int i = 0; i < 12; ++i) stats[i] = 0;
            stats[0] = -1;
            stats[1] = c->cloud->size();
            const std::string s = log.str();
            auto vec = [&](const char* key, double* out, int n) {
                size_t at = s.rfind(key);
                if (at == std::string::npos) return;
                std::string rest = s.substr(at + std::strlen(key));
                for (char& ch : rest) if (ch == '(' || ch == ')') ch = ' ';
                std::istringstream is(rest);
                for (int i = 0; i < n; ++i) is >> out[i];
            };
            vec("Average velocity:", stats + 2, 3);
            vec("Average omega:", stats + 5, 3);
            vec("Average momentum:", stats + 8, 3);
            vec("total K:", stats + 11, 1);
        }
    });
}

// the body of the sub-step loop of particleCloud::solve (particleCloud.C:1252-1265), one pass, by calling the
// reference's own member functions in the reference's order
int orc_cloud_substep(void* h, double localDt)
{
    RefCloud* c = (RefCloud*)h;
    return guarded([&]() {
        particleCloud& pc = *c->cloud;
        if (pc.moving_)
        {
            pc.computeCollisions(localDt);
            pc.computeWallHits(localDt);
            pc.computeCyclicHits();
        }
        forAllIter(Cloud<particleIBM>, pc, it) it().solve(localDt, pc.g_, pc.moving_, pc.rotating_);
        pc.setPositions();
    });
}

// particleCloud.C:1277-1286 (the indicator is NOT refreshed by the reference: quirk B#13)
int orc_cloud_update(void* h)
{
    RefCloud* c = (RefCloud*)h;
    return guarded([&]() {
        forAllIter(Cloud<particleIBM>, *c->cloud, it) it().update();
        c->cloud->deleteInActiveParticles();
    });
}

void orc_cloud_cd(void* h, double* out)
{
    RefCloud* c = (RefCloud*)h;
    IOField<scalar> cds(c->cloud->size());
    c->cloud->Cd(cds);
    forAll(cds, i) out[i] = cds[i];
}

double orc_cloud_max_co_num(void* h) { return ((RefCloud*)h)->cloud->maxCoNum(); }
double orc_particle_mass(void* h, int inst) { return nth((RefCloud*)h, inst).mass(); }
double orc_particle_inertia(void* h, int inst) { return nth((RefCloud*)h, inst).shape_->I(); }

// the reference's own restart writer (particleCloudIO.C:30-45 -> particleIBMIO.C:58-108): the mock keeps what the IOFields
// write in memory.  orc_written_field copies the field lagrangian/IBMparticles/<name> (flattened) and returns its length.
int orc_cloud_write_fields(void* h)
{
    RefCloud* c = (RefCloud*)h;
    return guarded([&]() {
        mockWrittenFields().clear();
        c->cloud->writeObject(IOstream::ASCII, IOstream::versionNumber(), IOstream::UNCOMPRESSED, true);
    });
}
int orc_written_field(void* h, const char* name, double* out, int maxN)
{
    RefCloud* c = (RefCloud*)h;
    const std::string path = std::string(c->time.timeName()) + "/lagrangian/IBMparticles/" + name;
    auto it = mockWrittenFields().find(path);
    if (it == mockWrittenFields().end()) return -1;
    const int n = (int)it->second.size();
    for (int i = 0; i < n && i < maxN; ++i) out[i] = it->second[size_t(i)];
    return n;
}

// collision tables of one instance after computeCollisions/computeWallHits (dense, as the reference stores them)
void orc_particle_collision_tables(void* h, int inst, double* Fc, double* Tc, double* Fw, double* Tw)
{
    const particleIBM& p = nth((RefCloud*)h, inst);
    forAll(p.Fc(), i)
        for (int a = 0; a < 3; ++a) { Fc[3 * i + a] = p.Fc()[i][a]; Tc[3 * i + a] = p.Tc()[i][a]; }
    forAll(p.Fw(), i)
        for (int a = 0; a < 3; ++a) { Fw[3 * i + a] = p.Fw()[i][a]; Tw[3 * i + a] = p.Tw()[i][a]; }
}

// static bool particleIBM::collision(p1, p2, pt) on two instances (covers the ellipse iteration, particleIBM.C:76-204)
int orc_pair_collision(void* h, int inst1, int inst2, double* hitPt)
{
    RefCloud* c = (RefCloud*)h;
    vector pt = Zero;
    const bool hit = particleIBM::collision(nth(c, inst1), nth(c, inst2), pt);
    for (int a = 0; a < 3; ++a) hitPt[a] = pt[a];
    return hit ? 1 : 0;
}

}  // extern "C"
