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
inline void __syncwarp(unsigned = 0xffffffffu) { ::emu::warp_barrier(); }

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
inline int atomicCAS(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }

inline double __shfl_down_sync(unsigned, double v, int d) { return ::emu::shfl(v, ::emu::lane_id() + d); }
inline double __shfl_xor_sync(unsigned, double v, int m) { return ::emu::shfl(v, ::emu::lane_id() ^ m); }
inline int __shfl_down_sync(unsigned, int v, int d) { return (int)::emu::shfl((double)v, ::emu::lane_id() + d); }
inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)::emu::shfl((double)v, ::emu::lane_id() ^ m); }
