/* ibmfoam_b200 — C ABI of the B200-native immersed-boundary hot path.
 *
 * This is the drop-in boundary for IBMFoam's per-step IB loops.  The reference has no C plugin
 * interface: its boundary is the C++ class API of libIBMparticleCloud / libparticleIBM
 * (src/particleCloud/particleCloud.H:267-358, src/particles/particles/particleIBM/particleIBM.H:194-391),
 * called by pimpleIBMFoam at three sites (createFields.H:60, UEqn.H:20-26, pimpleIBMFoam.C:62,87).
 * Every entry point below cites the reference member it replaces; shim/particleCloud.{H,C} shows
 * the OpenFOAM-side binding.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns 0 on success or a
 *     negative IBM_E_* code; ibm_last_error(ctx) gives the message (the reference aborts through
 *     FatalErrorInFunction; the shim turns a non-zero status into exactly that).
 *   - one ibm_ctx per particleCloud, bound to one CUDA device and one stream; not re-entrant.
 *   - all arithmetic is fp64 (OpenFOAM scalar), labels are 32-bit.
 *   - "host" entry points take caller-ordered AoS arrays exactly as OpenFOAM stores them
 *     (vector = 3 doubles, symmTensor = xx xy xz yy yz zz) and include the H2D/D2H copies: an array that is pinned or was
 *     registered with ibm_host_register is copied directly by the DMA engine; ordinary pageable memory (OpenFOAM's default)
 *     goes through two library-owned pinned chunks filled by host threads while the previous chunk is on the bus;
 *     "_device" entry points take device pointers to structure-of-arrays fields in LATTICE order
 *     ([component][k][j][i], i fastest) and never touch the host.
 *   - there is no CPU fallback: without a CUDA device ibm_create fails with IBM_E_CUDA.
 */
#ifndef IBMFOAM_B200_H
#define IBMFOAM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBM_ABI_VERSION 1

enum { IBM_OK = 0, IBM_E_ARG = -1, IBM_E_CUDA = -2, IBM_E_STATE = -3, IBM_E_UNSUPPORTED = -4, IBM_E_COMM = -5,
       IBM_E_CAPACITY = -6 };

enum { IBM_SPHERE = 0, IBM_CYLINDER = 1, IBM_ELLIPSE = 2 };            /* particleShape run-time table: particleShapeNew.C:27-82 */
enum { IBM_PATCH = 0, IBM_WALL = 1, IBM_CYCLIC = 2, IBM_EMPTY = 3, IBM_PROCESSOR = 4 }; /* particleCloud.C:1050-1068 */
enum { IBM_OFF_MESH = 0, IBM_PARTIAL = 1, IBM_ON_MESH = 2 };           /* particleShape.H:62-67 */
enum { IBM_XMIN = 0, IBM_XMAX = 1, IBM_YMIN = 2, IBM_YMAX = 3, IBM_ZMIN = 4, IBM_ZMAX = 5 };

typedef struct ibm_ctx ibm_ctx;

/* One boundary patch (fvPatch).  Cf/Sf/magSf are mesh.Cf()/Sf()/magSf().boundaryField()[patchi]
 * (particleIBM.C:466-470, 504-508) in patch face order; NULL = generate them for a lattice side in
 * the mini-mesh order (transverse index, fastest first). */
typedef struct {
  int type;        /* IBM_PATCH .. IBM_PROCESSOR */
  int side;        /* IBM_XMIN .. IBM_ZMAX: lattice side the patch lies on */
  int neighbPatch; /* cyclicPolyPatch::neighbPatchID(), else -1 */
  int nFaces;      /* ignored when Cf == NULL */
  const double* Cf;    /* [nFaces][3] */
  const double* Sf;    /* [nFaces][3] */
  const double* magSf; /* [nFaces]    */
} ibm_patch_desc;

/* The fvMesh as the path sees it: a uniform hex lattice (blockMesh with simpleGrading 1) whose cells
 * the caller numbers block by block.  Geometry the reference reads from polyMesh (cellCentres,
 * faceCentres, cells(), findCell — particleShape.C:39-66, sphere.C:208-216) is evaluated from the
 * lattice on the device and never stored. */
typedef struct {
  int nx, ny, nz;
  double origin[3];
  double delta[3];
  int nBlocks;            /* 0 = one block */
  const int* blocks;      /* [nBlocks][6] = i0,i1,j0,j1,k0,k1 (half-open), blockMesh order */
  const int* latticeToCell; /* optional explicit caller cell id per lattice cell (overrides blocks) */
  int nPatches;
  const ibm_patch_desc* patches;
} ibm_mesh_desc;

/* The reference reads its geometry from polyMesh (mesh.C(), mesh.Cf(): particleShape.C:39-66, sphere.C:208-216).  This helper
 * recognises the uniform lattice in those raw arrays, so that a caller need not describe it: n[3], origin[3], delta[3]
 * from the cell centres of any mesh whose cells are the cells of a uniform hex lattice, in ANY numbering (one blockMesh
 * block, several, renumbered); latticeToCell[nCells] (may be NULL) receives the caller's cell id of every lattice cell and
 * *identity tells whether it is the identity (then ibm_mesh_desc.latticeToCell may stay NULL); patchSide[p] the lattice side
 * (IBM_XMIN..IBM_ZMAX) the face centres patchCf[p][patchNFaces[p]][3] lie on.  Host-only, no CUDA call.  IBM_E_UNSUPPORTED:
 * the centres are not a uniform lattice (graded or unstructured mesh) or a patch does not lie on one side. */
int ibm_lattice_from_arrays(const double* cellCentres, int nCells, int nGeometricD, int nPatches, const double* const* patchCf,
                            const int* patchNFaces, int* n, double* origin, double* delta, int* patchSide, int* latticeToCell,
                            int* identity);

/* constant/IBMProperties top level + g (particleCloud.C:989-1005, 1034-1036; createFields.H:1) */
typedef struct {
  int moving, rotating, coupled;
  double e;
  double UInf[3];
  double rhoRef, pRef;
  double g[3];
  int cyclicWithoutWalls; /* 0 = as coded (no periodic copies unless a wall patch exists) */
  int maxInstances;       /* capacity for particle instances (originals + periodic / processor copies); 0 = 2n+64 */
} ibm_cloud_cfg;

/* constant/IBMProperties particle.<i> (particleIBM.C:269-297, particleShape.C:241-245, sphere.C:62, ellipse.C:77-78) */
typedef struct {
  int kind;
  double position[3];
  double d, a, b;
  int nTheta, nk;
  double delta;
  double theta[3], v[3], omega[3];
  double rho;
} ibm_particle_init;

/* particleIBM accessors (particleIBMI.H:25-213) for one instance, in cloud iteration order */
typedef struct {
  int index, copy, active, onMesh;
  double center[3], theta[3], v[3], omega[3];
  double rho, age;
  double F[3], T[3];
  int nShell, N;
} ibm_particle_state;

/* end-of-solve log block (particleCloud.C:1288-1323) */
typedef struct {
  int nSubSteps, nInstances;
  double avgV[3], avgOmega[3], avgP[3];
  double totK;
} ibm_solve_stats;

/* accumulated since ibm_timings_reset: kernel time and transfer time are reported SEPARATELY */
typedef struct {
  double update_ms, forcing_ms, stress_ms, advance_ms; /* kernels, CUDA events on the ctx stream */
  double h2d_ms, d2h_ms, layout_ms;                    /* pinned copies; AoS<->SoA + cell-order kernels */
  long long h2d_bytes, d2h_bytes;
  long long n_kernel_launches;
  long long n_forcing_calls, n_updates;                /* cell-particle updates (SURVEY §8d unit) */
} ibm_timing;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* particleCloud::particleCloud(mesh, dict, g)  particleCloud.C:974-1134 (mesh + top-level dict part) */
int ibm_create(const ibm_mesh_desc* mesh, const ibm_cloud_cfg* cfg, int device, ibm_ctx** out);
/* ... the per-particle part of the constructor: particleIBM(mesh, dict, index) for every entry,
 * deleteInActiveParticles, setCollisions, updateI  (particleCloud.C:1070-1133) */
int ibm_set_particles(ibm_ctx* ctx, int n, const ibm_particle_init* init);
int ibm_destroy(ibm_ctx* ctx);
const char* ibm_last_error(const ibm_ctx* ctx); /* ctx may be NULL: error of the last failed ibm_create */
int ibm_abi_version(void);

/* ---- loop 1: marking ----------------------------------------------------------------------------- */
/* particleIBM::update() for every instance + deleteInActiveParticles + updateI
 * (particleIBM.C:602-620; {sphere,cylinder,ellipse}::updateCellLists; particleShape::setNeighbours/setWeights;
 *  particleCloud.C:940-969) */
int ibm_update(ibm_ctx* ctx);

/* ---- loops 2+3: direct forcing ------------------------------------------------------------------- */
/* tmp<volVectorField> particleCloud::forcing(const volVectorField& S)   particleCloud.C:1158-1208
 * U, Uold are the "U" and U.oldTime() internal fields the reference looks up itself (:1193-1194);
 * dt = mesh.time().deltaT() (particleIBM.C:223).  F receives the full field (zero outside shells). */
int ibm_forcing(ibm_ctx* ctx, const double* U, const double* Uold, const double* S, double dt, double* F);
int ibm_forcing_device(ibm_ctx* ctx, const double* dU, const double* dUold, const double* dS, double dt, double* dF);

/* ---- loop 4: surface-stress integral ------------------------------------------------------------- */
/* particleCloud::integrateSurfaceStresses()  particleCloud.C:1210-1229 — p is the "p" field, devRhoReff the
 * cell field of particleCloud::devRhoReff() (:41-101, stays OpenFOAM's).  FT (optional) gets [nInstances][6]. */
int ibm_integrate_stress(ibm_ctx* ctx, const double* p, const double* devRhoReff6, double* FT);
int ibm_integrate_stress_device(ibm_ctx* ctx, const double* dP, const double* dTau6, double* FT_host);
/* particleCloud::combineForces()  particleCloud.C:171-313 (all-reduce of the partials when ibm_comm_init was called) */
int ibm_combine_forces(ibm_ctx* ctx);

/* ---- loop 5: collisions, walls, periodic copies, integration -------------------------------------- */
/* the sub-step loop + update part of particleCloud::solve()  particleCloud.C:1247-1324 */
int ibm_advance(ibm_ctx* ctx, double dt, ibm_solve_stats* stats);
/* void particleCloud::solve()  particleCloud.C:1242-1324  = integrate_stress + combine_forces + advance */
int ibm_solve(ibm_ctx* ctx, const double* p, const double* devRhoReff6, double dt, ibm_solve_stats* stats);
int ibm_solve_device(ibm_ctx* ctx, const double* dP, const double* dTau6, double dt, ibm_solve_stats* stats);
/* one literal pass of the sub-step body (computeCollisions, computeWallHits, computeCyclicHits,
 * particleIBM::solve, setPositions: particleCloud.C:1253-1265) — known-answer tests */
int ibm_substep(ibm_ctx* ctx, double localDt);

/* ---- accessors ----------------------------------------------------------------------------------- */
int ibm_n_instances(ibm_ctx* ctx, int* n);                    /* Cloud<particleIBM>::size(), particleCloud.H:282-285 */
int ibm_get_particles(ibm_ctx* ctx, ibm_particle_state* out); /* [nInstances] */
/* particleCloud::readParticles + the tail of the constructor on a restart (particleCloudIO.C:47-69, particleIBMIO.C:111-213,
 * particleCloud.C:1119-1133): the instance list is REPLACED by the n records read from <time>/lagrangian/IBMparticles (index, copy,
 * active, center, theta, v, omega, rho, age; shape from particle.<index> of IBMProperties, i.e. from ibm_set_particles), then
 * update(), deleteInActiveParticles and maxCopies_[index] = 0 as coded.  F, T start from zero like a freshly read particleIBM. */
int ibm_restore_particles(ibm_ctx* ctx, int n, const ibm_particle_state* st);
int ibm_set_force_torque(ibm_ctx* ctx, const double* FT);     /* particleIBM::F()/T() non-const refs, particleIBMI.H:136-149 */
int ibm_cd(ibm_ctx* ctx, double* out);                        /* particleCloud::Cd, particleCloud.C:1231-1240 */
int ibm_max_co_num(ibm_ctx* ctx, double* out);                /* particleCloud::maxCoNum — always 0, particleCloud.C:1145-1156 */
int ibm_particle_mass(ibm_ctx* ctx, int inst, double* mass, double* inertia); /* particleIBMI.H:96-99, particleShapeI.H */
int ibm_substep_schedule(double dt, double* dts, int maxN);   /* literal loop of particleCloud.C:1247-1275; returns count */

/* ---- parity taps (membership is part of the contract: bit-exact) --------------------------------- */
/* shellCells_, neighbourPoints_ (j,k), Is_, Os_, wFromLocal_, WFromLocal_ of one instance, shell cells in
 * ascending caller cell id (particleShape.H:131-158).  Any pointer may be NULL. */
int ibm_get_shell(ibm_ctx* ctx, int inst, int* cells, int* jk, int* Is, int* Os, double* w, double* W);
/* mesh_[N], findCell(mesh_[n]), number of internal faces, neighbour cell across each face (cells() order),
 * wToLocal_, WToLocal_ (particleShape.C:42-91).  Any pointer may be NULL. */
int ibm_get_points(ibm_ctx* ctx, int inst, double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W);
int ibm_get_sf(ibm_ctx* ctx, int inst, double* Sf);           /* Sf_ [nTheta*nk][3] */
int ibm_get_indicator(ibm_ctx* ctx, double* I);               /* I.ibm, caller cell order (particleCloud.C:952-969) */
/* EXTENSION, no reference counterpart (the reference only has the binary I.ibm): fractional solid volume fraction
 * alpha_s[cell] = min(1, sum over instances of the share of nSub^D regularly spaced sample points of the cell that lie
 * inside the body), caller cell order.  BASELINE.json north star, loop (1); SURVEY.md §8f row 4. */
int ibm_get_solid_fraction(ibm_ctx* ctx, int nSub, double* alpha);

/* ---- device-resident helpers (bench / multi-GPU harness) ----------------------------------------- */
int ibm_device_field_bytes(ibm_ctx* ctx, int nComponents, size_t* bytes);
/* caller AoS (host) -> lattice SoA (device) and back, using the ctx stream and pinned staging */
int ibm_upload_field(ibm_ctx* ctx, const double* hostAoS, int nComponents, double* dSoA);
int ibm_download_field(ibm_ctx* ctx, const double* dSoA, int nComponents, double* hostAoS);
/* cudaHostRegister / cudaHostUnregister of a caller-owned host array that persists across calls (the shim registers U,
 * U.oldTime() and p: fields that live in the mesh registry, UEqn.H:20-26, particleCloud.C:1193-1194, 1212); temporaries such as S,
 * devRhoReff() and the returned forcing field take the staged path */
int ibm_host_register(ibm_ctx* ctx, const void* ptr, size_t bytes);
int ibm_host_unregister(ibm_ctx* ctx, const void* ptr);
int ibm_sync(ibm_ctx* ctx);
int ibm_set_kernel_variant(ibm_ctx* ctx, int forcingVariant, int stressVariant); /* 0 = reference-order kernels, 1.. = tuned generations */
/* diagnostic counters: "loop5_active_instances" (of the last ibm_advance with the quiet/active split, else -1), "loop5_fallbacks",
 * "forcing_variant", "stress_variant", "mark_variant" (what is in force, also when set through the IBM_* environment switches) */
int ibm_get_counter(ibm_ctx* ctx, const char* name, long long* value);
/* Switches (defaults are what bench.py measures; every schedule of loop 5 gives the same results, bit for bit):
 *   "active_loop5" (1)    quiet / active split of a fluid step: instances that provably cannot touch anything integrate all
 *                         their sub-steps in one kernel; budgets are checked, a violation redoes the step unsplit
 *   "cluster_loop5" (1)   with the split: every group of possibly interacting instances runs all its sub-steps in one block
 *   "async_prepare" (1)   ibm_solve: the classification of the coming advance on a side stream beside the stress integral
 *   "owned_loop5" (1)     several ranks: a rank advances the instances and groups whose centre lies in its slab; states and
 *                         new periodic copies are exchanged once per fluid step (0: every rank advances everything)
 *   "graph_loop5" (1)     the general per-sub-step launch sequence (the fallback) replayed as a CUDA graph
 *   "fused_loop5" (0), "fused_active_loop5" (0)   that sequence / the active list as one persistent cooperative kernel
 *   "active_loop5_safety_ppm" (2000000)   displacement-budget factor in parts per million
 *   "mark_variant" (1)    1 = k1_mark2, 0 = k1_mark (same lists)
 *   "collision_model" (0) see ibm_set_real_option
 * Counters (ibm_get_counter) beyond those above: "loop5_cluster_steps", "loop5_owned_steps", "loop5_cluster_flags",
 * "loop5_block_substeps", "loop5_general_substeps", "v7_ok", "v7_slow_count", "h2d_pinned_bytes", "h2d_staged_bytes". */
int ibm_set_option(ibm_ctx* ctx, const char* name, int value);

/* EXTENSION without a counterpart in the reference (the reference names a soft-sphere model in
 * src/particleCloud/Make/files but ships none): ibm_set_option(ctx, "collision_model", 1) replaces the hard-sphere impulse of
 * particleCloud::computeCollisions (particleCloud.C:505-521) for sphere / cylinder pairs by a linear spring - dashpot -
 * Coulomb-slider contact law evaluated every sub-step while the bodies overlap:
 *   F_n = max(0, kn * overlap - gamma_n * (v_rel . n)),   F_t = -min(gamma_t |v_t|, mu F_n) v_t / |v_t|
 * with v_rel the relative velocity of the two surfaces at the contact point (rotation included).  Walls and periodic
 * copies keep the reference's treatment.  Parameters (all >= 0): "soft_sphere_kn", "soft_sphere_gamma_n",
 * "soft_sphere_gamma_t", "soft_sphere_mu".  Default collision_model = 0 = the reference, bit for bit. */
int ibm_set_real_option(ibm_ctx* ctx, const char* name, double value);

/* ---- timing -------------------------------------------------------------------------------------- */
int ibm_timings(ibm_ctx* ctx, ibm_timing* out);
int ibm_timings_reset(ibm_ctx* ctx);

/* ---- multi-GPU: slab-decomposed fields, replicated particle state (replaces the Pstream::combineReduce call
 * sites of particleCloud.C:201-264 and the decomposePar sub-meshes).  Order of calls on every
 * rank: ibm_create(WHOLE mesh) -> ibm_set_slab -> ibm_comm_init -> ibm_set_particles(ALL particles).
 * Cell fields handed to ibm_forcing / ibm_integrate_stress are then the rank's own z-planes [k0,k1). --- */
int ibm_set_slab(ibm_ctx* ctx, int k0, int k1);
/* EXTENSION (the reference's parallel results depend on the decomposition, particleShape.C:45-47,59): the same
 * slab with `halo` read-only planes of the neighbouring slabs on either side (clipped at the mesh).  The caller's cell
 * fields then cover the planes [*localK0, *localK1) returned here, halo planes filled with the neighbours' values (what a
 * processor-patch exchange provides).  Shell cells are marked, shape points counted and surface points integrated in the
 * owned planes [k0,k1) only, gathers read the halo: when the halo is deep enough that every stencil is complete, forcing
 * field, force and torque equal the SERIAL run's, whatever the decomposition.  Deep enough: (2 delta + polar-bin arc)/dz + 1
 * planes (3-4) for the regular polar bins; but the degenerate bins at the pole of a sphere reference point 0 - the OPPOSITE
 * pole - as coded (particleShape.C:148, App. B #3), so a sphere needs (d + delta)/dz + 1 planes to be exact.
 * halo = 0 is ibm_set_slab. */
int ibm_set_slab_halo(ibm_ctx* ctx, int k0, int k1, int halo, int* localK0, int* localK1);
int ibm_comm_unique_id_bytes(void);
int ibm_comm_get_unique_id(void* idBytes);               /* rank 0; broadcast the bytes with the launcher's own transport */
int ibm_comm_init(ibm_ctx* ctx, const void* idBytes, int rank, int nRanks);   /* NCCL, one process per GPU */
int ibm_comm_info(ibm_ctx* ctx, int* rank, int* nRanks);
/* in-process transport for tests: ranks are host threads, all contexts on one device */
int ibm_comm_local_group_create(int nRanks, void** group);
int ibm_comm_local_group_destroy(void* group);
int ibm_comm_init_local(ibm_ctx* ctx, void* group, int rank);

#ifdef __cplusplus
}
#endif
#endif /* IBMFOAM_B200_H */
