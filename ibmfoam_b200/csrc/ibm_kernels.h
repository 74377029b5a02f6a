// Kernel launchers shared between translation units (reference-order TU, tuned TU, motion TU, C ABI).
#pragma once
#include "ibm_device.cuh"

namespace ibm {

// kernel attributes (dynamic shared-memory size, carve-out) belong to the device a kernel will run on: call sites set them
// once per (call site, current device), so a process that drives several devices gets them on each
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};


// ---- ibm_kernels_ref.cu (compiled with -fmad=false) ----
void launch_k1_count(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst);
void launch_k1_mark(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                    int* shellCell, int* shellJK, int shellCap, int* cursor, int* overflow, int variant = 0, const int* nearList = nullptr,
                    const int* nearCount = nullptr);
struct Schedule;
// several ranks: the instances far from the slab settled by one thread each, the others listed for the marking kernel (variant 1)
void launch_k1_far(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst, Schedule& s,
                   int* nearFlag, int* nearOff, int* nearList, int* dNearCount);
void launch_k1_status(cudaStream_t st, const LatticeDev& Lg, const ShapeClassDev* classes, const InstanceSoA& in, int nInst);
void launch_k1_fill(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                    int* shellCell, int* shellJK);
void launch_k1_indicator(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                         double* Ifield);
void launch_tap_shell(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int inst,
                      const int* shellCell, const int* shellJK, int* cells, int* jk, int* Is, int* Os, double* w, double* W);
void launch_tap_points(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int inst,
                       double* pts, int* cell, int* nFaces, int* otherCell, double* w, double* W);
void launch_forcing_ref(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                        const int* shellCell, const int* shellJK, const double* U, const double* Uold, const double* S,
                        double dT, double* F);
void launch_stress_ref(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr);

// ---- schedule (ibm_kernels_util.cu): Morton-sorted instance order + table of <=FORCING_CHUNK-cell work items ----
#define FORCING_CHUNK 128
struct Schedule {
  unsigned* keys = nullptr; unsigned* keysSorted = nullptr;
  int* iota = nullptr; int* order = nullptr;      // order[pos] = instance scheduled at position pos
  int* chunkCnt = nullptr; int* chunkOff = nullptr;  // per sorted position; chunkOff has n+1 entries
  int chunk = FORCING_CHUNK;                         // shell cells per work item of the direct-gather kernel (128, or 256 with option chunk256)
  void* temp = nullptr; size_t tempBytes = 0;
  void* temp2 = nullptr;                             // scratch of the sort (it may run on a side stream beside the scans that use temp)
  int cap = 0;
};
int schedule_alloc(Schedule& s, int cap);
void schedule_free(Schedule& s);
// builds order[] and chunkOff[]; writes the total number of work items to *dTotalChunks
void launch_build_schedule(cudaStream_t st, const LatticeDev& L, const InstanceSoA& in, int nInst, Schedule& s, int* dTotalChunks,
                           int* dTotalUpdates);
// its two halves: the Morton order needs the centres only (it can run beside the marking kernel), the work-item table the shell counts
void launch_schedule_order(cudaStream_t st, const LatticeDev& L, const InstanceSoA& in, int nInst, Schedule& s);
void launch_schedule_chunks(cudaStream_t st, const InstanceSoA& in, int nInst, Schedule& s, int* dTotalChunks, int* dTotalUpdates);
void launch_forcing_v2(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, int totalChunks, const int* shellCell, const int* shellJK, const double* U,
                       const double* Uold, const double* S, double dT, double* F);
void launch_stress_v2(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr);

// ---- ibm_kernels_v7.cu: one CTA per instance, ball-row staging of combined fields, gathers from shared memory ----
struct V7Caps { int planeCap, colCap, rowsCap, mapCap, nTcap, nkCap, elemCap, prefetch; };  // shared-memory capacities of a launch (doubles / entries)
// Per-instance launch records of the staged kernels, built once per update (k7_prepare) in SCHEDULE order, so that a CTA
// finds everything its prologue needs - which instance, its box, the staged rows of the three reaches - with loads that
// depend on nothing but blockIdx (one memory round trip instead of the chain order -> instance -> class -> box -> rows, and
// no per-call recomputation of the region: the centres only change in solve(), which ends with an update)
struct V7Hdr {
  int b, cls, nShell, shellOff, flags;  // flags: 1 has local points, 2 the staged path applies (geometry), 4 box touches the lattice boundary
  int j0, k0, by, bz, i0, bx;           // bounding box of the outer reach (Box7)
  int nA, nB, nC, pad0, pad1;           // staged cells of the outer / surface / inner reach
  double cx, cy, cz, pad2;
};
struct V7Prep {
  V7Hdr* hdr = nullptr;       // [cap]
  unsigned* rows = nullptr;   // [cap][rowsCap][3]: first column - i0 | (last column - i0) << 8 | offset of the row << 16; empty row: 0xff
  int rowsCap = 0, cap = 0;
  // multi-GPU: the records of the instances with local points, compacted in schedule order (pad0 = schedule position, which
  // indexes `rows`), so that the staged kernels launch one CTA per LOCAL instance instead of one per instance of the directory
  V7Hdr* hdrLocal = nullptr;  // [cap]
  int* flag = nullptr;        // [cap + 1]
  int* flagOff = nullptr;     // [cap + 1]
  int nLocal = -1;            // records in hdrLocal; -1 = not compacted (one GPU): the kernels take hdr / blockIdx as they are
};
// capacities from the shape classes (HOST copies of the class records); false = the fast path does not apply to this cloud
bool v7_plan(const LatticeDev& L, const ShapeClassDev* hostClasses, int nClasses, int nGeoD, V7Caps& forcing, V7Caps& stress);
int v7_prep_alloc(V7Prep& p, int cap, int rowsCap);
void v7_prep_free(V7Prep& p);
// geometry of the records (needs the centres and the order only: may run beside the marking kernel) ...
void launch_v7_prepare(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, const V7Caps& capF, const V7Prep& prep);
// ... and what the marking kernel decides: shell segment, has-local-points flag
void launch_v7_finish(cudaStream_t st, const InstanceSoA& in, int nInst, const Schedule& s, const V7Prep& prep);
// ... and, on several ranks, the compacted list of the local records (count to *dCount; the caller reads it back into prep.nLocal)
void launch_v7_compact(cudaStream_t st, int nInst, Schedule& s, const V7Prep& prep, int* dCount);
void launch_forcing_v7(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                       const Schedule& s, const V7Prep& prep, const int* shellCell, const int* shellJK, const double* U, const double* Uold,
                       const double* S, double dT, double* F, const V7Caps& cap, int* slowList, int* slowCount, int* guard);
void launch_stress_v7(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst,
                      const Schedule& s, const V7Prep& prep, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr,
                      const V7Caps& cap, int* slowList, int* slowCount, int* guard, double* Fclear = nullptr, const int* shellCell = nullptr);
// (Fclear: the forcing field, to be re-zeroed at every instance's shell cells on the way - solve() only)
// second-generation kernels over a device-side list of instances (the ones a staged kernel could not take)
void launch_forcing_v2_list(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, const int* list,
                            const int* listCount, const int* shellCell, const int* shellJK, const double* U, const double* Uold,
                            const double* S, double dT, double* F);
void launch_stress_v2_list(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, const int* list,
                           const int* listCount, const double* P, const double* Tau, double pRefOverRho, double rhoRef, double* Tcorr);

// ---- ibm_kernels_alpha.cu: fractional solid volume fraction (extension) ----
void launch_solid_fraction(cudaStream_t st, const LatticeDev& L, const ShapeClassDev* classes, const InstanceSoA& in, int nInst, int nSub,
                           double* alpha);

// ---- ibm_kernels_util.cu ----
void launch_keep_flags(cudaStream_t st, const InstanceSoA& in, int nInst, int* keep);
void launch_compact_instances(cudaStream_t st, const InstanceSoA& src, const InstanceSoA& dst, const int* keep, const int* keepOff,
                              int nInst);
void launch_exclusive_scan(cudaStream_t st, const int* in, int* out, int n, int* total);
void launch_exclusive_scan_total(cudaStream_t st, Schedule& s, const int* in, int* out, int n, int* total);
void launch_aos_to_soa(cudaStream_t st, const LatticeDev& L, const double* aos, int nComp, double* soa);
void launch_soa_to_aos(cudaStream_t st, const LatticeDev& L, const double* soa, int nComp, double* aos);
void launch_copy_int(cudaStream_t st, int* dst, const int* src);
void launch_clear_shell_cells(cudaStream_t st, const int* shellCell, int total, size_t C, double* F);

}  // namespace ibm
