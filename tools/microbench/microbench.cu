// Secondary ceilings of the immersed-boundary kernels, measured on the box ("measure with an fp64 FMA
// micro-benchmark since MEASURED_PEAKS.json has only bf16"):
//   fp64_fma      dependent-chain-free DFMA throughput (the forcing / stress / marking kernels are fp64-instruction bound)
//   fp64_rsqrt    rsqrt(double) throughput (six per gather stencil)
//   launch        back-to-back empty kernels in a stream and replayed from a CUDA graph (loop 5: ~9 launches x 101 sub-steps)
//   grid_sync     cooperative grid-wide barrier against the number of blocks (k5a_fused: 4-6 barriers per sub-step)
//   hbm_copy      device-to-device copy bandwidth (the roofline denominator, for comparison with MEASURED_PEAKS.json)
// One JSON line on stdout.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } \
  } while (0)

__global__ void k_fma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void k_rsqrt(double* out, int iters, double a) {
  double x0 = 1.0 + threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  for (int i = 0; i < iters; ++i) {
    x0 = rsqrt(x0) + a; x1 = rsqrt(x1) + a; x2 = rsqrt(x2) + a; x3 = rsqrt(x3) + a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (x0 + x1) + (x2 + x3);
}

__global__ void k_empty() {}

__global__ void k_gridsync(int n, int* sink) {
  cg::grid_group g = cg::this_grid();
  int v = 0;
  for (int i = 0; i < n; ++i) {
    g.sync();
    v += i;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = v;
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int dev = 0, sms = 0, coop = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  double* buf = nullptr;
  CK(cudaMalloc(&buf, (size_t)sms * 8 * 1024 * sizeof(double)));

  // fp64 FMA and rsqrt
  const int blocks = sms * 8, threads = 256, iters = 20000;
  k_fma<<<blocks, threads, 0, st>>>(buf, 100, 1.0000001, 1e-9);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  k_fma<<<blocks, threads, 0, st>>>(buf, iters, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  const double fmaPerS = (double)blocks * threads * iters * 8 / (elapsed(e0, e1) * 1e-3);
  k_rsqrt<<<blocks, threads, 0, st>>>(buf, 100, 1.5);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  k_rsqrt<<<blocks, threads, 0, st>>>(buf, iters / 4, 1.5);
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  const double rsqrtPerS = (double)blocks * threads * (iters / 4) * 4 / (elapsed(e0, e1) * 1e-3);

  // launch latency: stream and graph
  const int nL = 2000;
  for (int i = 0; i < 100; ++i) k_empty<<<1, 32, 0, st>>>();
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < nL; ++i) k_empty<<<1, 32, 0, st>>>();
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  const double usStream = elapsed(e0, e1) * 1e3 / nL;
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < nL; ++i) k_empty<<<1, 32, 0, st>>>();
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  const double usGraph = elapsed(e0, e1) * 1e3 / nL;

  // grid-wide barrier against the number of blocks (256 threads each)
  std::vector<int> gridSizes = {4, 16, 64, 148, 296};
  std::vector<double> usSync;
  int* sink = nullptr;
  CK(cudaMalloc(&sink, sizeof(int)));
  for (int gs : gridSizes) {
    double us = -1.0;
    int perSm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_gridsync, 256, 0));
    if (coop && gs <= perSm * sms) {
      int n = 2000;
      void* args[] = {&n, &sink};
      CK(cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(gs), dim3(256), args, 0, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaEventRecord(e0, st));
      CK(cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(gs), dim3(256), args, 0, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      us = elapsed(e0, e1) * 1e3 / n;
    }
    usSync.push_back(us);
  }

  // device-to-device copy
  const size_t bytes = (size_t)2 << 30;
  char *a = nullptr, *b = nullptr;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  const double copyGBs = 5.0 * 2.0 * bytes / (elapsed(e0, e1) * 1e-3) / 1e9;  // read + write

  std::printf("{\"sms\": %d, \"fp64_fma_per_s\": %.4g, \"fp64_tflops\": %.3f, \"fp64_rsqrt_per_s\": %.4g, \"launch_us_stream\": %.3f, "
              "\"launch_us_graph\": %.3f, \"grid_sync_us\": {",
              sms, fmaPerS, 2.0 * fmaPerS / 1e12, rsqrtPerS, usStream, usGraph);
  for (size_t i = 0; i < gridSizes.size(); ++i) std::printf("%s\"%d\": %.3f", i ? ", " : "", gridSizes[i], usSync[i]);
  std::printf("}, \"hbm_copy_GBs\": %.1f}\n", copyGBs);
  return 0;
}
