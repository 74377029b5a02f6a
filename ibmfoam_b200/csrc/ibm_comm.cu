// Multi-GPU plumbing: slab-decomposed fields, replicated particle state.
//
// Replaces the Pstream::combineReduce call sites of the reference:
//   combineForces (particleCloud.C:201-264)  -> one all-reduce(sum) of the per-instance F/T partials
//   setWeights status across ranks            -> one all-reduce(sum) of the per-instance on-mesh point counts
// setPositions / the N x N collision-table reductions / the append-gathers of new copies need NO message:
// every rank holds every instance and runs loop 5 redundantly with deterministic kernels.
//
// Two transports behind one call:
//   * NCCL (one process per GPU; libnccl resolved with dlopen so the core library has no link dependency),
//   * an in-process "local group" (rank = host thread, all contexts on one device) used by the tests to
//     exercise the multi-rank path on a single GPU: host threads rendez-vous on a barrier, then one kernel
//     per rank sums all ranks' buffers in rank order (no kernel ever waits for another kernel).
#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <mutex>

#include "ibm_ctx.h"

namespace ibm {

// ---- minimal NCCL ABI (nccl.h: ncclUniqueId is 128 opaque bytes; enums are stable across 2.x) ----
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclSum = 0, kNcclInt32 = 2, kNcclInt64 = 4, kNcclFloat64 = 8 };
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef const char* (*fn_GetErrorString)(int);
typedef int (*fn_Group)(void);

struct NcclApi {
  void* h = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  fn_Group GroupStart = nullptr, GroupEnd = nullptr;
};
static NcclApi g_nccl;
static std::string g_ncclErr;

static bool load_nccl() {
  if (g_nccl.h) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) { g_ncclErr = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_GetErrorString)dlsym(g_nccl.h, "ncclGetErrorString");
  g_nccl.GroupStart = (fn_Group)dlsym(g_nccl.h, "ncclGroupStart");
  g_nccl.GroupEnd = (fn_Group)dlsym(g_nccl.h, "ncclGroupEnd");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    g_ncclErr = "libnccl is missing a required symbol";
    return false;
  }
  return true;
}

// ---- local group ----
struct LocalGroup {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<const void*> src;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    long g = generation;
    if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g; });
  }
};

struct Comm {
  int rank = 0, nRanks = 1;
  LocalGroup* lg = nullptr;
  ncclComm_t nccl = nullptr;
  void* tmp = nullptr;
  size_t tmpBytes = 0;
  const void** dSrcList = nullptr;
};

template <class T>
__global__ void k_sum_ranks(const void* const* src, int nRanks, size_t count, T* dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    T acc = ((const T*)src[0])[i];
    for (int r = 1; r < nRanks; ++r) acc += ((const T*)src[r])[i];
    dst[i] = acc;
  }
}

int comm_ranks(ibm_ctx* c) { return c->comm ? ((Comm*)c->comm)->nRanks : 1; }
int comm_rank(ibm_ctx* c) { return c->comm ? ((Comm*)c->comm)->rank : 0; }

// in-place all-reduce(sum) of `count` doubles (isInt = 0), int32 (isInt = 1) or int64 (isInt = 2: bit patterns of values
// that exactly one rank contributes travel unchanged, x + 0 + ... + 0 being exact in integers) on the ctx stream
int comm_allreduce_sum(ibm_ctx* c, void* buf, size_t count, int isInt) {
  Comm* cm = (Comm*)c->comm;
  if (!cm || cm->nRanks == 1 || count == 0) return IBM_OK;
  const size_t esz = isInt == 1 ? sizeof(int) : sizeof(double);
  if (cm->nccl) {
    int rc = g_nccl.AllReduce(buf, buf, count, isInt == 1 ? kNcclInt32 : (isInt == 2 ? kNcclInt64 : kNcclFloat64), kNcclSum, cm->nccl, c->stream);
    if (rc != 0) return ctx_fail(c, IBM_E_COMM, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return IBM_OK;
  }
  LocalGroup* lg = cm->lg;
  if (cm->tmpBytes < count * esz) {
    if (cm->tmp) cudaFree(cm->tmp);
    cm->tmp = nullptr; cm->tmpBytes = 0;
    IBM_CUDA(c, cudaMalloc(&cm->tmp, count * esz));
    cm->tmpBytes = count * esz;
  }
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));  // my contribution is complete
  { std::lock_guard<std::mutex> lk(lg->mu); lg->src[cm->rank] = buf; }
  lg->barrier();                                   // all contributions complete and published
  IBM_CUDA(c, cudaMemcpyAsync(cm->dSrcList, lg->src.data(), lg->n * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
  int blocks = (int)std::min<size_t>((count + 255) / 256, 1184);
  if (isInt == 1) k_sum_ranks<int><<<blocks, 256, 0, c->stream>>>(cm->dSrcList, lg->n, count, (int*)cm->tmp);
  else if (isInt == 2) k_sum_ranks<long long><<<blocks, 256, 0, c->stream>>>(cm->dSrcList, lg->n, count, (long long*)cm->tmp);
  else k_sum_ranks<double><<<blocks, 256, 0, c->stream>>>(cm->dSrcList, lg->n, count, (double*)cm->tmp);
  IBM_CUDA(c, cudaStreamSynchronize(c->stream));
  lg->barrier();                                   // everyone has finished reading the sources
  IBM_CUDA(c, cudaMemcpyAsync(buf, cm->tmp, count * esz, cudaMemcpyDeviceToDevice, c->stream));
  c->nLaunch++;
  return IBM_OK;
}

void comm_destroy(ibm_ctx* c) {
  Comm* cm = (Comm*)c->comm;
  if (!cm) return;
  if (cm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(cm->nccl);
  if (cm->tmp) cudaFree(cm->tmp);
  if (cm->dSrcList) cudaFree(cm->dSrcList);
  delete cm;
  c->comm = nullptr;
}

}  // namespace ibm

using namespace ibm;

extern "C" {

int ibm_comm_unique_id_bytes(void) { return (int)sizeof(NcclUniqueId); }

int ibm_comm_get_unique_id(void* idBytes) {
  if (!idBytes) return IBM_E_ARG;
  if (!load_nccl()) return ctx_fail(nullptr, IBM_E_COMM, g_ncclErr);
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return ctx_fail(nullptr, IBM_E_COMM, "ncclGetUniqueId failed");
  std::memcpy(idBytes, &id, sizeof(id));
  return IBM_OK;
}

int ibm_comm_init(ibm_ctx* c, const void* idBytes, int rank, int nRanks) {
  if (!c || !idBytes || rank < 0 || rank >= nRanks) return ctx_fail(c, IBM_E_ARG, "ibm_comm_init: bad arguments");
  if (c->comm) return ctx_fail(c, IBM_E_STATE, "ibm_comm_init: communicator already set");
  if (c->cap != 0) return ctx_fail(c, IBM_E_STATE, "ibm_comm_init: call before ibm_set_particles");
  cudaSetDevice(c->device);
  if (!load_nccl()) return ctx_fail(c, IBM_E_COMM, g_ncclErr);
  Comm* cm = new Comm;
  cm->rank = rank; cm->nRanks = nRanks;
  NcclUniqueId id;
  std::memcpy(&id, idBytes, sizeof(id));
  int rc = g_nccl.CommInitRank(&cm->nccl, nRanks, id, rank);
  if (rc != 0) { delete cm; return ctx_fail(c, IBM_E_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); }
  c->comm = cm;
  return IBM_OK;
}

int ibm_comm_local_group_create(int nRanks, void** group) {
  if (!group || nRanks < 1) return IBM_E_ARG;
  LocalGroup* lg = new LocalGroup;
  lg->n = nRanks;
  lg->src.assign(nRanks, nullptr);
  *group = lg;
  return IBM_OK;
}
int ibm_comm_local_group_destroy(void* group) {
  delete (LocalGroup*)group;
  return IBM_OK;
}
int ibm_comm_init_local(ibm_ctx* c, void* group, int rank) {
  if (!c || !group) return IBM_E_ARG;
  LocalGroup* lg = (LocalGroup*)group;
  if (rank < 0 || rank >= lg->n) return ctx_fail(c, IBM_E_ARG, "ibm_comm_init_local: bad rank");
  if (c->comm) return ctx_fail(c, IBM_E_STATE, "ibm_comm_init_local: communicator already set");
  if (c->cap != 0) return ctx_fail(c, IBM_E_STATE, "ibm_comm_init_local: call before ibm_set_particles");
  cudaSetDevice(c->device);
  Comm* cm = new Comm;
  cm->rank = rank; cm->nRanks = lg->n; cm->lg = lg;
  if (!ctx_cuda_ok(c, cudaMalloc(&cm->dSrcList, lg->n * sizeof(void*)), "cudaMalloc")) { delete cm; return IBM_E_CUDA; }
  c->comm = cm;
  return IBM_OK;
}

int ibm_comm_info(ibm_ctx* c, int* rank, int* nRanks) {
  if (!c) return IBM_E_ARG;
  if (rank) *rank = comm_rank(c);
  if (nRanks) *nRanks = comm_ranks(c);
  return IBM_OK;
}

}  // extern "C"
