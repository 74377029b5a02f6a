// Host-side context of one particleCloud bound to one CUDA device + stream.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/ibmfoam_b200.h"
#include "ibm_kernels.h"

namespace ibm {

struct PatchHost {
  int type = 0, side = 0, neighb = -1;
  std::vector<double> Cf, Sf, magSf;
  double *dCf = nullptr, *dSf = nullptr, *dMagSf = nullptr;
  int nFaces() const { return (int)magSf.size(); }
};

struct ShapeClassHost {
  ShapeClassDev dev{};  // device pointers filled after upload
  std::vector<double> centered, Sf, cosT, sinT, sinP, cosP, Rtab, ck;
};

struct EventPair {
  cudaEvent_t a, b;
  int cat;
};

enum TimeCat { T_UPDATE = 0, T_FORCING, T_STRESS, T_ADVANCE, T_H2D, T_D2H, T_LAYOUT, T_NCAT };

}  // namespace ibm

struct ibm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  ibm_cloud_cfg cfg{};
  ibm::LatticeDev L{};   // lattice loops 1-4 run on: the whole mesh, or this rank's z-slab
  ibm::LatticeDev Lg{};  // the whole mesh (loop 5, patches, global location status)
  int* dLatToCell = nullptr;
  int* dCellToLat = nullptr;
  std::vector<int> hLatToCell;
  std::vector<ibm::PatchHost> patches;
  std::vector<int> walls, cyclics;  // patch ids (particleCloud.C:1050-1068)
  int nGeometricD = 3;
  double zMin = 0, zMax = 0;  // min/max(mesh.points()).z  (cylinder.C:73)

  std::vector<ibm::ShapeClassHost> classes;
  ibm::ShapeClassDev* dClasses = nullptr;
  std::vector<double*> classPools;

  int cap = 0, nInst = 0, nParticles = 0;
  ibm::InstanceSoA in{};
  ibm::InstanceSoA in2{};  // compaction / append scratch
  std::vector<void*> instAllocs;
  std::vector<ibm_particle_init> orig;  // origParticles_ (particleCloud.C:1082-1092)
  std::vector<int> origCls;
  std::vector<int> maxCopies;

  int* dShellCell = nullptr;
  int* dShellJK = nullptr;
  int shellCap = 0, totalShell = 0;
  int* dScal = nullptr;   // small device scalars (totals, counters)
  int* hScal = nullptr;   // pinned mirror
  int* dKeep = nullptr;   // keep flags / prefix
  int* dKeepOff = nullptr;

  double *dU = nullptr, *dUold = nullptr, *dS = nullptr, *dF = nullptr, *dP = nullptr, *dTau = nullptr, *dI = nullptr;
  double* dStage = nullptr;
  size_t stageDoubles = 0;
  double* dTcorr = nullptr;
  bool indicatorStale = true;  // I.ibm is computed on demand (ibm_get_indicator), not on every update
  bool fLive = false;       // dF holds a previous forcing result on the current shell cells
  bool fEverZeroed = false;
  double* hPinned = nullptr;  // small pinned scratch for results
  void* hChunk[2] = {nullptr, nullptr};          // pinned staging chunks for pageable caller arrays
  cudaEvent_t chunkEv[2] = {nullptr, nullptr};
  long long h2dPinnedBytes = 0, h2dStagedBytes = 0;  // host->device bytes copied directly from pinned / registered arrays, through the chunks
  size_t hPinnedDoubles = 0;

  // kernel generation of loops 2-4: 0 reference-order (every sum and division as coded: the on-device yardstick), 2 direct gather
  // with Morton-scheduled work items (round-1 default; now the path for clouds the staged kernels cannot take),
  // 7 one CTA per instance, ball-row staging of combined fields (round-2 default where every class fits: v7ok)
  int forcingVariant = 7, stressVariant = 7;
  bool v7ok = false;              // every shape class can take the seventh-generation kernels (no ellipse, tables + region fit)
  ibm::V7Caps v7capF{}, v7capS{};  // shared-memory capacities of their launches
  int *dNearFlag = nullptr, *dNearOff = nullptr, *dNearList = nullptr;  // [cap + 1]: marking over the instances near the slab (several ranks)
  bool clearFInStress = false;     // set by ibm_solve around its stress integral: the staged stress kernel re-zeroes the forcing field
  cudaStream_t updSide = nullptr;  // ctx_update: Morton order + launch records beside the marking kernel
  cudaEvent_t updFork = nullptr, updJoin = nullptr;
  ibm::V7Prep v7prep;             // launch records of the staged kernels (k7_prepare), rebuilt by every update
  int* dSlowList = nullptr;       // [cap] instances the staged kernel handed to the direct-gather kernel (device-side list)
  ibm::Schedule sched;
  int schedN = 0;          // instances the schedule's Morton order covers (set by ctx_update; 0 = no valid order)
  int totalChunks = 0;
  int updatesPerCall = 0;

  bool chunk256 = false;    // direct-gather forcing kernel: 256 shell cells per work item (one item per d=8 sphere)
  bool fusedLoop5 = false;  // one persistent cooperative kernel for the sub-step loop
  bool graphLoop5 = true;   // else: the per-phase launch sequence of a fluid step replayed as a CUDA graph
  bool activeLoop5 = true;  // quiet / active split of a fluid step (ibm_motion.cu, k5q_bounds): same results, less work per
                            // sub-step; written after round 1's GPU budget was spent, verified under emulation only
  int activeSafetyPpm = 2000000;  // displacement-budget safety factor in parts per million (>= 1: a quiet instance needs B_i >= s_i)
  int markVariant = 1;            // 1 = k1_mark2 (compacted polar-bin phase; 1.46 vs 2.07 ms per update at 512^3 / 10^5), 0 = k1_mark; same outputs
  int collisionModel = 0;         // EXTENSION: 1 = soft-sphere contact law (ibm_motion.cu d5_pair_force); 0 = the reference's impulse
  double softKn = 0.0, softGammaN = 0.0, softGammaT = 0.0, softMu = 0.0;
  bool asyncPrepare = true;        // ibm_solve: loop-5 classification and the re-zeroing of F on a side stream beside the stress integral
  bool clusterOneBlockCC = false;  // test hook: connected components by the single-block kernel
  bool clusterGeneralOnly = false; // test hook: every sub-step of a block takes the general path (no resident fast path)
  bool clusterLoop5 = true;       // split steps: independent groups of active instances, all sub-steps of a group in one block (k5c_run)
  bool ownedLoop5 = true;         // multi-GPU cluster steps: every rank advances the instances / groups its slab owns, one exchange per fluid step
  bool fusedActive = false;       // split steps: the active list's sub-steps in one persistent cooperative kernel (k5a_fused)
  int activeCheckPpm = 1000000;   // fraction of B_i the per-sub-step check allows (tests shrink it to force the fallback)
  // motion (loop 5) scratch lives in ibm_motion.cu
  void* motion = nullptr;
  // comm
  void* comm = nullptr;

  // developer timeline (IBM_DEBUG_TIMING=1): named marks on the stream, mean interval between consecutive marks printed at destroy
  bool dbgTiming = false;
  std::vector<std::pair<const char*, cudaEvent_t>> dbgMarks;
  std::vector<double> dbgHost;
  std::vector<std::pair<std::string, std::vector<float>>> dbgAcc, dbgAccHost;  // samples per "A -> B"
  std::vector<ibm::EventPair> events;
  std::vector<cudaEvent_t> eventPool;
  double accMs[ibm::T_NCAT] = {0, 0, 0, 0, 0, 0, 0};
  long long h2dBytes = 0, d2hBytes = 0, nLaunch = 0, nForcingCalls = 0, nUpdates = 0;
};

namespace ibm {
int ctx_fail(ibm_ctx* c, int code, const std::string& msg);
bool ctx_cuda_ok(ibm_ctx* c, cudaError_t e, const char* what);
void ctx_mark(ibm_ctx* c, const char* label);   // no-op unless IBM_DEBUG_TIMING is set
void ctx_marks_report(ibm_ctx* c);
void ctx_tic(ibm_ctx* c, int cat);
void ctx_toc(ibm_ctx* c);
int ctx_ensure_stage(ibm_ctx* c, size_t doubles);
int ctx_ensure_pinned(ibm_ctx* c, size_t doubles);
int ctx_update(ibm_ctx* c);
// ibm_motion.cu
int motion_advance(ibm_ctx* c, double dt, ibm_solve_stats* stats);
int motion_prepare_async(ibm_ctx* c, double dt);   // classification of the coming advance on a side stream (call before the stress integral)
cudaStream_t motion_side_stream(ibm_ctx* c);      // that stream while a prepared classification is pending, else null
int motion_substep(ibm_ctx* c, double localDt);
void motion_destroy(ibm_ctx* c);
void motion_invalidate_groups(ibm_ctx* c);
void motion_invalidate_graph(ibm_ctx* c);
void motion_reset_copies(ibm_ctx* c);  // copy counters of loop 5 <- c->maxCopies (after a restart)
bool motion_counter(ibm_ctx* c, const char* name, long long* value);
int combine_forces(ibm_ctx* c);
// ibm_comm.cu
void comm_destroy(ibm_ctx* c);
int comm_allreduce_sum(ibm_ctx* c, void* buf, size_t count, int isInt);
int comm_ranks(ibm_ctx* c);
int comm_rank(ibm_ctx* c);
}  // namespace ibm

#define IBM_CUDA(c, call)                                            \
  do {                                                               \
    if (!ibm::ctx_cuda_ok((c), (call), #call)) return IBM_E_CUDA;    \
  } while (0)
