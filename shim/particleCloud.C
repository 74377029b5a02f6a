/*---------------------------------------------------------------------------*\
  OpenFOAM-side binding of libibmfoam_b200 - see particleCloud.H.  Every member forwards to one C-ABI call
  (include/ibmfoam_b200.h); OpenFOAM keeps the mesh, the fields, the PIMPLE loop and the linear solvers.
\*---------------------------------------------------------------------------*/
#include "particleCloud.H"

namespace Foam
{
    defineTypeNameAndDebug(particleCloud, 0);
}

void Foam::particleCloud::check(const int rc, const char* where) const
{
    if (rc != IBM_OK)
    {
        FatalErrorInFunction
            << where << ": " << ibm_last_error(ctx_) << exit(FatalError);
    }
}

void Foam::particleCloud::describeMesh(ibm_mesh_desc& md, List<ibm_patch_desc>& pats, labelList& blocks) const
{
    const fvBoundaryMesh& bm = mesh_.boundary();
    pats.setSize(bm.size());
    labelList sideOf(bm.size(), label(0));
    md.nBlocks = 0; md.blocks = nullptr; md.latticeToCell = nullptr;
    if (!dict_.found("lattice"))
    {
        // What the reference has: polyMesh's own arrays (mesh.C(), mesh.Cf(): particleShape.C:39-66).  The library recognises
        // the uniform lattice in them, in whatever order blockMesh numbered the cells, and refuses anything else.
        const vectorField& cc = mesh_.C().primitiveField();
        List<const double*> cf(bm.size());
        labelList nf(bm.size());
        forAll(bm, patchi)
        {
            cf[patchi] = reinterpret_cast<const double*>(mesh_.Cf().boundaryField()[patchi].begin());
            nf[patchi] = bm[patchi].size();
        }
        int n[3], identity = 1;
        latticeToCell_.setSize(cc.size());
        check(ibm_lattice_from_arrays(reinterpret_cast<const double*>(cc.begin()), cc.size(), mesh_.nGeometricD(), bm.size(),
                                      cf.begin(), nf.begin(), n, md.origin, md.delta, sideOf.begin(), latticeToCell_.begin(), &identity),
              "mesh (uniform lattice expected)");
        md.nx = n[0]; md.ny = n[1]; md.nz = n[2];
        if (!identity) md.latticeToCell = latticeToCell_.begin();
    }
    else
    {
        // optional: IBMProperties { lattice { n (nx ny nz); origin (x y z); delta (dx dy dz); blocks ((i0 i1 j0 j1 k0 k1) ...);
        // patchSides (xMin ...); } } - the numbers of system/blockMeshDict, for callers that have them at hand
        const dictionary& lat = dict_.subDict("lattice");
        const labelVector n(lat.lookup("n"));
        const vector origin(lat.lookup("origin")), delta(lat.lookup("delta"));
        md.nx = n.x(); md.ny = n.y(); md.nz = n.z();
        forAll(origin, d) { md.origin[d] = origin[d]; md.delta[d] = delta[d]; }
        if (lat.found("blocks"))
        {
            List<labelList> b(lat.lookup("blocks"));
            blocks.setSize(6*b.size());
            forAll(b, i) forAll(b[i], j) blocks[6*i + j] = b[i][j];
        }
        md.nBlocks = blocks.size()/6;
        md.blocks = blocks.size() ? blocks.begin() : nullptr;
        const wordList sides(lat.lookup("patchSides"));   // one of xMin xMax yMin yMax zMin zMax per patch
        forAll(bm, patchi)
        {
            const word& s = sides[patchi];
            sideOf[patchi] = s == "xMin" ? IBM_XMIN : s == "xMax" ? IBM_XMAX : s == "yMin" ? IBM_YMIN
                           : s == "yMax" ? IBM_YMAX : s == "zMin" ? IBM_ZMIN : IBM_ZMAX;
        }
    }

    // boundary patches with their OpenFOAM face arrays (particleIBM.C:466-470, 504-508)
    forAll(bm, patchi)
    {
        ibm_patch_desc& p = pats[patchi];
        p.type =
            isA<wallFvPatch>(bm[patchi]) ? IBM_WALL
          : isA<cyclicFvPatch>(bm[patchi]) ? IBM_CYCLIC
          : isA<processorFvPatch>(bm[patchi]) ? IBM_PROCESSOR
          : isA<emptyFvPatch>(bm[patchi]) ? IBM_EMPTY : IBM_PATCH;
        p.side = sideOf[patchi];
        p.neighbPatch =
            isA<cyclicFvPatch>(bm[patchi])
          ? refCast<const cyclicPolyPatch>(mesh_.boundaryMesh()[patchi]).neighbPatchID() : -1;
        p.nFaces = bm[patchi].size();
        // Foam::vector is three contiguous doubles: the boundary fields can be passed as they are
        p.Cf = reinterpret_cast<const double*>(mesh_.Cf().boundaryField()[patchi].begin());
        p.Sf = reinterpret_cast<const double*>(mesh_.Sf().boundaryField()[patchi].begin());
        p.magSf = mesh_.magSf().boundaryField()[patchi].begin();
    }
    md.nPatches = pats.size();
    md.patches = pats.begin();
}

Foam::particleCloud::particleCloud
(
    const fvMesh& mesh,
    const dictionary& dict,
    const dimensionedVector& g
)
:
    regIOobject(IOobject("IBMparticles", mesh.time().timeName(), "lagrangian", mesh, IOobject::NO_READ, IOobject::AUTO_WRITE)),
    mesh_(mesh),
    dict_(dict),
    nParticles_(readLabel(dict_.lookup("nParticles"))),
    UInf_(dict_.subDict("flow").lookupOrDefault<vector>("UInf", Zero)),
    rhoRef_(dict_.subDict("flow").lookupOrDefault<scalar>("rhoInf", 1.0)),
    pRef_(dict_.subDict("flow").lookupOrDefault<scalar>("pInf", 0.0)),
    g_(g.value()),
    e_(readScalar(dict_.lookup("e"))),
    coupled_(dict_.lookup("coupled")),
    moving_(dict_.lookup("moving")),
    rotating_(dict_.lookup("rotating")),
    I_(IOobject(IOobject::groupName("I", "ibm"), mesh.time().timeName(), mesh), mesh, dimensionedScalar("0", dimless, 0.0)),
    ctx_(nullptr)
{
    ibm_mesh_desc md;
    List<ibm_patch_desc> pats;
    labelList blocks;
    describeMesh(md, pats, blocks);

    ibm_cloud_cfg cfg;
    cfg.moving = moving_; cfg.rotating = rotating_; cfg.coupled = coupled_; cfg.e = e_;
    forAll(UInf_, d) { cfg.UInf[d] = UInf_[d]; cfg.g[d] = g_[d]; }
    cfg.rhoRef = rhoRef_; cfg.pRef = pRef_;
    cfg.cyclicWithoutWalls = dict_.lookupOrDefault<Switch>("cyclicWithoutWalls", false);
    cfg.maxInstances = dict_.lookupOrDefault<label>("maxInstances", 0);

    const label device = dict_.lookupOrDefault<label>("device", Pstream::parRun() ? Pstream::myProcNo() % 8 : 0);
    if (ibm_create(&md, &cfg, device, &ctx_) != IBM_OK)
    {
        FatalErrorInFunction << ibm_last_error(nullptr) << exit(FatalError);
    }

    // EXTENSION (absent from the reference, which only names a soft-sphere model in src/particleCloud/Make/files):
    //     collisionModel softSphere;  softSphereCoeffs { kn ...; gamman ...; gammat ...; mu ...; }
    // in IBMProperties selects the spring - dashpot - slider contact law; without the keyword the reference's impulse is used
    if (dict_.lookupOrDefault<word>("collisionModel", "hardSphere") == "softSphere")
    {
        const dictionary& sDict = dict_.subDict("softSphereCoeffs");
        check(ibm_set_real_option(ctx_, "soft_sphere_kn", readScalar(sDict.lookup("kn"))), "softSphereCoeffs");
        check(ibm_set_real_option(ctx_, "soft_sphere_gamma_n", sDict.lookupOrDefault<scalar>("gamman", 0.0)), "softSphereCoeffs");
        check(ibm_set_real_option(ctx_, "soft_sphere_gamma_t", sDict.lookupOrDefault<scalar>("gammat", 0.0)), "softSphereCoeffs");
        check(ibm_set_real_option(ctx_, "soft_sphere_mu", sDict.lookupOrDefault<scalar>("mu", 0.0)), "softSphereCoeffs");
        check(ibm_set_option(ctx_, "collision_model", 1), "collisionModel");
    }

    // particle.<i> sub-dictionaries (reference particleCloud.C:1070-1117, particleIBM.C:269-297)
    List<ibm_particle_init> init(nParticles_);
    forAll(init, i)
    {
        const dictionary& pDict = dict_.subDict(IOobject::groupName("particle", Foam::name(i)));
        ibm_particle_init& p = init[i];
        const word shape(pDict.lookup("particleShape"));
        p.kind = shape == "sphere" ? IBM_SPHERE : shape == "cylinder" ? IBM_CYLINDER : shape == "ellipse" ? IBM_ELLIPSE : -1;
        const vector pos(pDict.lookup("position"));
        const vector theta(pDict.lookupOrDefault<vector>("theta", Zero));
        const vector v(pDict.lookupOrDefault<vector>("v", Zero));
        const vector omega(pDict.lookupOrDefault<vector>("omega", Zero));
        forAll(pos, d) { p.position[d] = pos[d]; p.theta[d] = theta[d]; p.v[d] = v[d]; p.omega[d] = omega[d]; }
        p.d = pDict.lookupOrDefault<scalar>("d", 0.0);
        p.a = pDict.lookupOrDefault<scalar>("a", 0.0);
        p.b = pDict.lookupOrDefault<scalar>("b", 0.0);
        p.nTheta = readLabel(pDict.lookup("nTheta"));
        p.nk = pDict.lookupOrDefault<label>("nk", 1);
        p.delta = readScalar(pDict.lookup("delta"));
        p.rho = readScalar(pDict.lookup("rho"));
    }
    check(ibm_set_particles(ctx_, init.size(), init.begin()), "particleCloud");
    // restart: the instances come from <time>/lagrangian/IBMparticles instead of the dictionary (reference particleCloud.C:1095-1123)
    if (mesh_.time().value() != 0)
    {
        readParticles();
    }
    check(ibm_get_indicator(ctx_, I_.primitiveFieldRef().begin()), "updateI");
}

Foam::particleCloud::~particleCloud()
{
    forAll(registered_, i)
    {
        ibm_host_unregister(ctx_, registered_[i]);
    }
    ibm_destroy(ctx_);
}

Foam::label Foam::particleCloud::nParticles() const
{
    int n = 0;
    ibm_n_instances(ctx_, &n);
    return n;
}

Foam::scalar Foam::particleCloud::maxCoNum() const
{
    double v = 0;
    ibm_max_co_num(ctx_, &v);
    return v;
}

Foam::tmp<Foam::volVectorField> Foam::particleCloud::forcing(const volVectorField& S)
{
    tmp<volVectorField> tmpF
    (
        new volVectorField
        (
            IOobject("F", mesh_.time().timeName(), mesh_, IOobject::NO_READ, IOobject::NO_WRITE, false),
            mesh_,
            dimensionedVector("F", dimForce/dimVolume/dimDensity, Zero)
        )
    );
    if (!coupled_) return tmpF;

    const volVectorField& U = mesh_.lookupObject<volVectorField>("U");
    // U and U.oldTime() live in the mesh registry for the whole run: their storage is registered once so that the copies run
    // at the full DMA rate (S and the returned F are temporaries: the library stages them through its pinned chunks)
    registerField(U.primitiveField().begin(), U.size()*sizeof(vector));
    registerField(U.oldTime().primitiveField().begin(), U.size()*sizeof(vector));
    // Foam::vector is three contiguous doubles: internal fields are handed over without a copy
    check
    (
        ibm_forcing
        (
            ctx_,
            reinterpret_cast<const double*>(U.primitiveField().begin()),
            reinterpret_cast<const double*>(U.oldTime().primitiveField().begin()),
            reinterpret_cast<const double*>(S.primitiveField().begin()),
            mesh_.time().deltaT().value(),
            reinterpret_cast<double*>(tmpF.ref().primitiveFieldRef().begin())
        ),
        "forcing"
    );
    return tmpF;
}

void Foam::particleCloud::integrateSurfaceStresses()
{
    const volScalarField& p = mesh_.lookupObject<volScalarField>("p");
    tmp<volSymmTensorField> tau(devRhoReff());
    check
    (
        ibm_integrate_stress
        (
            ctx_,
            p.primitiveField().begin(),
            reinterpret_cast<const double*>(tau().primitiveField().begin()),
            nullptr
        ),
        "integrateSurfaceStresses"
    );
}

void Foam::particleCloud::Cd(IOField<scalar>& Cds) const
{
    Cds.setSize(nParticles());
    ibm_cd(ctx_, Cds.begin());
}

void Foam::particleCloud::registerField(const void* ptr, const std::size_t bytes) const
{
    forAll(registered_, i)
    {
        if (registered_[i] == ptr) return;
    }
    if (ibm_host_register(ctx_, ptr, bytes) == IBM_OK)   // (a refusal only costs speed: the staged path takes over)
    {
        registered_.setSize(registered_.size() + 1);
        registered_[registered_.size() - 1] = ptr;
    }
}

void Foam::particleCloud::solve()
{
    const volScalarField& p = mesh_.lookupObject<volScalarField>("p");
    registerField(p.primitiveField().begin(), p.size()*sizeof(scalar));
    tmp<volSymmTensorField> tau(devRhoReff());
    ibm_solve_stats st;
    check
    (
        ibm_solve
        (
            ctx_,
            p.primitiveField().begin(),
            reinterpret_cast<const double*>(tau().primitiveField().begin()),
            mesh_.time().deltaT().value(),
            &st
        ),
        "solve"
    );
    check(ibm_get_indicator(ctx_, I_.primitiveFieldRef().begin()), "updateI");
    // the Info<< block of the reference (particleCloud.C:1317-1322)
    Info<< "    Actual number of active particles: " << nParticles_ << nl
        << "    Total number of solved particles: " << st.nInstances << nl
        << "    Average velocity: " << vector(st.avgV[0], st.avgV[1], st.avgV[2]) << nl
        << "    Average omega: " << vector(st.avgOmega[0], st.avgOmega[1], st.avgOmega[2]) << nl
        << "    Average momentum: " << vector(st.avgP[0], st.avgP[1], st.avgP[2]) << nl
        << "    total K: " << st.totK << endl;
}

void Foam::particleCloud::writeFields() const
{
    // lagrangian/IBMparticles/{index,copy,active,center,theta,v,omega,rho,age}  (reference particleIBMIO.C:58-108)
    const label n = nParticles();
    List<ibm_particle_state> st(n);
    if (n) ibm_get_particles(ctx_, st.begin());
    // lagrangian/IBMparticles/positions (particle::writeFields, called first by particleIBM::writeFields): the lagrangian
    // position the reference tracks with the centre (particleIBM.C:313, 426-433); its presence is what readParticles keys on
    IOField<vector> positions(IOobject("positions", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<label> index(IOobject("index", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<label> copy(IOobject("copy", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<label> active(IOobject("active", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<vector> center(IOobject("center", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<vector> theta(IOobject("theta", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<vector> v(IOobject("v", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<vector> omega(IOobject("omega", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<scalar> rho(IOobject("rho", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    IOField<scalar> age(IOobject("age", time().timeName(), "lagrangian/IBMparticles", mesh_, IOobject::NO_READ), n);
    forAll(st, i)
    {
        index[i] = st[i].index; copy[i] = st[i].copy; active[i] = st[i].active;
        center[i] = vector(st[i].center[0], st[i].center[1], st[i].center[2]);
        positions[i] = center[i];
        theta[i] = vector(st[i].theta[0], st[i].theta[1], st[i].theta[2]);
        v[i] = vector(st[i].v[0], st[i].v[1], st[i].v[2]);
        omega[i] = vector(st[i].omega[0], st[i].omega[1], st[i].omega[2]);
        rho[i] = st[i].rho; age[i] = st[i].age;
    }
    positions.write(n > 0);
    index.write(); copy.write(); active.write(); center.write(); theta.write(); v.write(); omega.write();
    rho.write(); age.write();
}

bool Foam::particleCloud::writeObject
(
    IOstream::streamFormat,
    IOstream::versionNumber,
    IOstream::compressionType,
    const bool
) const
{
    writeFields();
    return true;
}

void Foam::particleCloud::readParticles()
{
    // reference particleCloudIO.C:47-69 + particleIBMIO.C:111-213: positions decide whether there is anything to read; the
    // per-particle fields are MUST_READ; every particle gets the shape of particle.<index> at the stored centre
    const fileName dir("lagrangian/IBMparticles");
    const word t(mesh_.time().timeName());
    IOobject posIO("positions", t, dir, mesh_, IOobject::READ_IF_PRESENT, IOobject::NO_WRITE, false);
    if (!posIO.typeHeaderOk<IOField<vector>>(false))
    {
        // "Assuming the initial cloud contains 0 particles" (particleCloudIO.C:62-68)
        check(ibm_restore_particles(ctx_, 0, nullptr), "readParticles");
        return;
    }
    auto io = [&](const char* name) { return IOobject(name, t, dir, mesh_, IOobject::MUST_READ, IOobject::NO_WRITE, false); };
    IOField<label> index(io("index"), true), copy(io("copy"), true), active(io("active"), true);
    IOField<vector> center(io("center"), true), theta(io("theta"), true), v(io("v"), true), omega(io("omega"), true);
    IOField<scalar> rho(io("rho"), true), age(io("age"), true);
    const label n = index.size();
    if (copy.size() != n || active.size() != n || center.size() != n || theta.size() != n || v.size() != n || omega.size() != n
     || rho.size() != n || age.size() != n)
    {
        FatalErrorInFunction << "lagrangian/IBMparticles fields of different sizes at time " << t << exit(FatalError);
    }
    List<ibm_particle_state> st(n);
    forAll(st, i)
    {
        ibm_particle_state& s = st[i];
        s.index = index[i]; s.copy = copy[i]; s.active = active[i]; s.onMesh = IBM_ON_MESH;
        forAll(center[i], d)
        {
            s.center[d] = center[i][d]; s.theta[d] = theta[i][d]; s.v[d] = v[i][d]; s.omega[d] = omega[i][d];
            s.F[d] = 0; s.T[d] = 0;
        }
        s.rho = rho[i]; s.age = age[i]; s.nShell = 0; s.N = 0;
    }
    check(ibm_restore_particles(ctx_, n, n ? st.begin() : nullptr), "readParticles");
}

// ---- per-particle view (reference particleIBM.H:194-391): the state lives on the GPU, a view is a snapshot of one instance ----
Foam::List<Foam::particleIBM> Foam::particleCloud::particles() const
{
    const label n = nParticles();
    List<ibm_particle_state> st(n);
    if (n) ibm_get_particles(ctx_, st.begin());
    List<particleIBM> out(n);
    forAll(out, i)
    {
        double m = 0, inertia = 0;
        ibm_particle_mass(ctx_, i, &m, &inertia);
        out[i] = particleIBM(st[i], m, inertia, dict_.subDict(IOobject::groupName("particle", Foam::name(label(st[i].index)))), UInf_, rhoRef_);
    }
    return out;
}

// ---- the viscous stress stays OpenFOAM's (it is not part of the GPU path).  Same sources, in the same order of
//      preference, as the reference's private members (particleCloud.C:41-169), which follow OpenFOAM's `forces`
//      function object: a compressible turbulence model's devRhoReff, an incompressible one's rho*devReff, a fluidThermo
//      mu, a transportModel nu, or the `nu` entry of a transportProperties dictionary. ----
Foam::tmp<Foam::volScalarField> Foam::particleCloud::rho() const
{
    if (mesh_.foundObject<volScalarField>("thermo:rho"))
    {
        return mesh_.lookupObject<volScalarField>("thermo:rho");
    }
    return tmp<volScalarField>
    (
        new volScalarField
        (
            IOobject("rho", mesh_.time().timeName(), mesh_),
            mesh_,
            dimensionedScalar("rho", dimDensity, rhoRef_)
        )
    );
}

Foam::tmp<Foam::volSymmTensorField> Foam::particleCloud::devRhoReff() const
{
    typedef compressible::turbulenceModel cmpModel;
    typedef incompressible::turbulenceModel icoModel;
    if (mesh_.foundObject<cmpModel>(cmpModel::propertiesName))
    {
        return mesh_.lookupObject<cmpModel>(cmpModel::propertiesName).devRhoReff();
    }
    if (mesh_.foundObject<icoModel>(icoModel::propertiesName))
    {
        return rho()*mesh_.lookupObject<icoModel>(icoModel::propertiesName).devReff();
    }
    const volVectorField& U = mesh_.lookupObject<volVectorField>("U");
    if (mesh_.foundObject<fluidThermo>(fluidThermo::dictName))
    {
        return -mesh_.lookupObject<fluidThermo>(fluidThermo::dictName).mu()*dev(twoSymm(fvc::grad(U)));
    }
    if (mesh_.foundObject<transportModel>("transportProperties"))
    {
        return -rho()*mesh_.lookupObject<transportModel>("transportProperties").nu()*dev(twoSymm(fvc::grad(U)));
    }
    if (mesh_.foundObject<dictionary>("transportProperties"))
    {
        const dimensionedScalar nu(mesh_.lookupObject<dictionary>("transportProperties").lookup("nu"));
        return -rho()*nu*dev(twoSymm(fvc::grad(U)));
    }
    FatalErrorInFunction << "No valid model for viscous stress calculation" << exit(FatalError);
    return volSymmTensorField::null();
}
