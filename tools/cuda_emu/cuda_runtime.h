// TEST INFRASTRUCTURE — stands in for <cuda_runtime.h> when a kernel header is compiled for the CPU emulation
// (tools/cuda_emu/emu.hpp).  Only what ibm_device.cuh / ibm_tables.cuh / ibm_kernels_v5.cuh use.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "emu.hpp"

#define IBM_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

typedef void* cudaStream_t;

#define threadIdx (::emu::cur().tid)
#define blockIdx (::emu::blk()->blockIdx)
#define blockDim (::emu::blk()->blockDim)
#define gridDim (::emu::blk()->gridDim)

inline void __syncthreads() { ::emu::syncthreads(); }
inline int __syncthreads_or(int pred) { return ::emu::syncthreads_or(pred); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { ::emu::group_barrier(mask); }

template <class T> inline T __ldg(const T* p) { return *p; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
using std::floor;
using std::fma;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double max(double a, double b) { return a > b ? a : b; }

inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline long long clock64() { return 0; }
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
inline int atomicCAS(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }

inline double __shfl_down_sync(unsigned k, double v, int d) { return ::emu::shfl(k, v, ::emu::lane_id() + d); }
inline double __shfl_xor_sync(unsigned k, double v, int m) { return ::emu::shfl(k, v, ::emu::lane_id() ^ m); }
inline int __shfl_down_sync(unsigned k, int v, int d) { return (int)::emu::shfl(k, (double)v, ::emu::lane_id() + d); }
inline int __shfl_xor_sync(unsigned k, int v, int m) { return (int)::emu::shfl(k, (double)v, ::emu::lane_id() ^ m); }
inline double __shfl_sync(unsigned k, double v, int src) { return ::emu::shfl(k, v, src); }
inline int __shfl_sync(unsigned k, int v, int src) { return (int)::emu::shfl(k, (double)v, src); }

// ================================================================================================
// Runtime API stand-in (whole-library emulation, tools/cuda_emu/build_emu.py): device memory is host memory, streams
// are synchronous, events read the host clock, a captured graph is the list of recorded launches.
// ================================================================================================
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned X = 1, unsigned Y = 1, unsigned Z = 1) : x(X), y(Y), z(Z) {}
};
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCooperativeLaunch = 95 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };

namespace emu {
struct Event { std::chrono::steady_clock::time_point t; };
typedef std::vector<std::function<void()>> Graph;
inline std::mutex& launch_mutex() { static std::mutex m; return m; }
inline Graph*& capturing() { static thread_local Graph* g = nullptr; return g; }
inline void submit(std::function<void()> fn) {
  if (capturing()) { capturing()->push_back(std::move(fn)); return; }
  fn();
}
template <class K, class... A>
inline void rt_launch(dim3 grid, dim3 block, size_t smem, K kern, A... args) {
  auto tup = std::make_tuple(args...);  // arguments are evaluated NOW, like a real launch (also under graph capture)
  submit([=]() {
    std::lock_guard<std::mutex> lk(launch_mutex());  // one emulated kernel at a time (static __shared__ storage)
    ::emu::launch(grid.x * grid.y * grid.z, block.x * block.y * block.z, smem, [&]() { std::apply(kern, tup); });
  });
}
// cooperative launch: every block resident, grid.sync() available (reached through ibm::launch_cooperative)
template <class K, class... A>
inline cudaError_t coop_launch(dim3 grid, dim3 block, size_t smem, K kern, A... args) {
  auto tup = std::make_tuple(args...);
  submit([=]() {
    std::lock_guard<std::mutex> lk(launch_mutex());
    ::emu::launch(grid.x * grid.y * grid.z, block.x * block.y * block.z, smem, [&]() { std::apply(kern, tup); }, true);
  });
  return cudaSuccess;
}
// cudaLaunchCooperativeKernel((void*)kernel, ...) is rewritten by build_emu.py to this: the kernel keeps its type, so the
// void** argument block can be unpacked (and copied NOW, like a real launch)
template <class... P, size_t... I>
inline cudaError_t coop_unpack(void (*kern)(P...), dim3 grid, dim3 block, void** args, size_t smem, std::index_sequence<I...>) {
  return coop_launch(grid, block, smem, kern, *(typename std::remove_reference<P>::type*)args[I]...);
}
template <class... P>
inline cudaError_t coop_launch_erased(void (*kern)(P...), dim3 grid, dim3 block, void** args, size_t smem, void* /*stream*/) {
  if (std::getenv("EMU_REFUSE_COOP")) return cudaErrorNotSupported;  // tests of the product's handling of a refused launch
  return coop_unpack(kern, grid, block, args, smem, std::index_sequence_for<P...>{});
}
}  // namespace emu
#define EMU_LAUNCH(kern, grid, block, smem, stream, ...) ::emu::rt_launch(dim3(grid), dim3(block), (size_t)(smem), kern, ##__VA_ARGS__)

typedef emu::Event* cudaEvent_t;
typedef emu::Graph* cudaGraph_t;
typedef emu::Graph* cudaGraphExec_t;

inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = std::malloc(n ? n : 1);
  if (!*p) return 2;
  std::memset(*p, 0xA5, n);  // device memory is not zeroed: make reads of uninitialised memory visible
  return cudaSuccess;
}
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  emu::submit([=]() { std::memmove(d, s, n); });
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  emu::submit([=]() { std::memset(d, v, n); });
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) { *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : (a == cudaDevAttrCooperativeLaunch) ? 1 : 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu::Event; return cudaSuccess; }
#define cudaEventDisableTiming 2u
#define cudaHostRegisterDefault 0u
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu::Event; return cudaSuccess; }
// registered host arrays (by base pointer): what cudaPointerGetAttributes reports as pinned
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
inline std::vector<std::pair<const char*, size_t>>& emu_registered() { static std::vector<std::pair<const char*, size_t>> v; return v; }
inline cudaError_t cudaHostRegister(void* p, size_t n, unsigned) { emu_registered().push_back({(const char*)p, n}); return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void* p) {
  auto& v = emu_registered();
  for (size_t i = 0; i < v.size(); ++i) if (v[i].first == (const char*)p) { v.erase(v.begin() + i); return cudaSuccess; }
  return cudaErrorInvalidValue;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = const_cast<void*>(p);
  for (auto& r : emu_registered()) if ((const char*)p >= r.first && (const char*)p < r.first + r.second) a->type = cudaMemoryTypeHost;
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }  // streams are synchronous here
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaLaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
  if (emu::capturing()) return cudaErrorInvalidValue;
  emu::capturing() = new emu::Graph;
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = emu::capturing(); emu::capturing() = nullptr; return *g ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long = 0) { *e = new emu::Graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t) {
  for (auto& fn : *g) fn();
  return cudaSuccess;
}

// ---- more device intrinsics (whole-library emulation) ----
inline unsigned __ballot_sync(unsigned mask, int pred) { return ::emu::ballot(mask, pred != 0); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline void __threadfence() {}
inline int __shfl_up_sync(unsigned k, int v, int d) { const int l = ::emu::lane_id(); return (int)::emu::shfl(k, (double)v, l - d >= 0 ? l - d : l); }
inline double __shfl_up_sync(unsigned k, double v, int d) { const int l = ::emu::lane_id(); return ::emu::shfl(k, v, l - d >= 0 ? l - d : l); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { unsigned o = *p; if (o == c) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) { unsigned long long o = *p; if (o == c) *p = v; return o; }
inline int __double2loint(double d) { long long l; std::memcpy(&l, &d, 8); return (int)(unsigned)(l & 0xffffffffLL); }
inline long long __double_as_longlong(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
inline double __longlong_as_double(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
using std::fabs;
using std::sqrt;
using std::sin;
using std::cos;
using std::atan2;
using std::acos;
using std::tan;
using std::fmax;
using std::fmin;
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
