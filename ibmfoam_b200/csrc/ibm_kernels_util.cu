// Plumbing kernels: exclusive scan of the per-instance shell counts, and the layout kernels that turn
// OpenFOAM's caller-ordered array-of-structures fields (vector = 3 doubles, symmTensor = 6) into the
// lattice-ordered structure-of-arrays planes the hot kernels read (and back for F).
#include <cub/cub.cuh>

#include "ibm_kernels.h"

namespace ibm {

// single-block exclusive scan (n <= a few 1e5 instances; runs once per update): every thread owns one contiguous
// segment - serial sum, one block-wide scan of the 1024 segment totals, serial write - i.e. three passes and two
// barriers whatever n is (the tile-by-tile version it replaces took 3 barriers per 1024 elements: 168 us at n = 1e5)
__global__ void k_exclusive_scan(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ total) {
  __shared__ int warpSum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int seg = (n + (int)blockDim.x - 1) / (int)blockDim.x;
  const int i0 = min((int)threadIdx.x * seg, n), i1 = min(i0 + seg, n);
  int mine = 0;
  for (int i = i0; i < i1; ++i) mine += in[i];
  int x = mine;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warpSum[warp] = x;
  __syncthreads();
  int wbase = 0, tot = 0;
  for (int q = 0; q < nw; ++q) { if (q < warp) wbase += warpSum[q]; tot += warpSum[q]; }
  int run = wbase + x - mine;  // sum of everything before this thread's segment
  for (int i = i0; i < i1; ++i) { const int v = in[i]; out[i] = run; run += v; }
  if (threadIdx.x == 0) *total = tot;
}

__global__ void k_aos_to_soa(LatticeDev L, const double* __restrict__ aos, int nComp, double* __restrict__ soa) {
  const size_t C = (size_t)L.nCells;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < C; l += (size_t)gridDim.x * blockDim.x) {
    size_t c = L.latToCell ? (size_t)L.latToCell[l] : l;
    for (int q = 0; q < nComp; ++q) soa[(size_t)q * C + l] = aos[c * nComp + q];
  }
}

__global__ void k_soa_to_aos(LatticeDev L, const double* __restrict__ soa, int nComp, double* __restrict__ aos) {
  const size_t C = (size_t)L.nCells;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < C; l += (size_t)gridDim.x * blockDim.x) {
    size_t c = L.latToCell ? (size_t)L.latToCell[l] : l;
    for (int q = 0; q < nComp; ++q) aos[c * nComp + q] = soa[(size_t)q * C + l];
  }
}

// re-zero only the cells a previous forcing call touched (the persistent F stays zero elsewhere)
__global__ void k_clear_shell_cells(const int* __restrict__ shellCell, int total, size_t C, double* __restrict__ F) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lat = shellCell[i];
    F[lat] = 0.0; F[C + lat] = 0.0; F[2 * C + lat] = 0.0;
  }
}

// deleteInActiveParticles (particleCloud.C:940-950): keep = onMesh != OFF_MESH, order preserved
__global__ void k_keep_flags(InstanceSoA in, int n, int* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = in.onMesh[i] != 0;
  else if (i == n) keep[i] = 0;  // sentinel: the exclusive scan over n + 1 entries leaves the total behind the last offset
}

__global__ void k_compact_instances(InstanceSoA s, InstanceSoA d, const int* __restrict__ keep, const int* __restrict__ keepOff, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  int j = keepOff[i];
  d.index[j] = s.index[i]; d.copy[j] = s.copy[i]; d.active[j] = s.active[i]; d.onMesh[j] = s.onMesh[i]; d.cls[j] = s.cls[i];
  d.nShell[j] = s.nShell[i]; d.shellOff[j] = s.shellOff[i]; d.nPtsOnMesh[j] = s.nPtsOnMesh[i]; d.onLocal[j] = s.onLocal[i];
  d.rho[j] = s.rho[i]; d.age[j] = s.age[i];
  for (int q = 0; q < 3; ++q) {
    d.c[q][j] = s.c[q][i]; d.th[q][j] = s.th[q][i]; d.v[q][j] = s.v[q][i]; d.om[q][j] = s.om[q][i];
    d.F[q][j] = s.F[q][i]; d.T[q][j] = s.T[q][i];
  }
  for (int q = 0; q < 9; ++q) d.rot[9 * (size_t)j + q] = s.rot[9 * (size_t)i + q];
}

// ---- schedule -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void k_morton_keys(LatticeDev L, InstanceSoA in, int n, int cap, unsigned* keys, int* iota) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  iota[i] = i;
  if (i >= n) { keys[i] = 0xffffffffu; return; }
  int I = min(max((int)floor((in.c[0][i] - L.ox) / L.dx), 0), L.nx - 1);
  int J = min(max((int)floor((in.c[1][i] - L.oy) / L.dy), 0), L.ny - 1);
  int K = min(max((int)floor((in.c[2][i] - L.oz) / L.dz) - L.kz0, 0), L.nz - 1);
  // 4-cell granularity keeps 10 bits per axis up to 4096 cells
  int sh = 0;
  while (((max(L.nx, max(L.ny, L.nz)) - 1) >> sh) > 1023) ++sh;
  keys[i] = spread10((unsigned)(I >> sh)) | (spread10((unsigned)(J >> sh)) << 1) | (spread10((unsigned)(K >> sh)) << 2);
}
__global__ void k_chunk_counts(InstanceSoA in, const int* __restrict__ order, int n, int* cnt, int* totalUpdates, int chunk) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0, ns = 0;
  if (pos < n) {
    int i = order[pos];
    if (in.onLocal[i] != 0) { ns = in.nShell[i]; c = (ns + chunk - 1) / chunk; }
  }
  if (pos <= n) cnt[pos] = c;  // entry n is a zero sentinel so that the exclusive scan yields n+1 offsets
  // cell-particle updates of one forcing call = shell cells of the instances the forcing kernels process
  for (int o = 16; o > 0; o >>= 1) ns += __shfl_down_sync(0xffffffffu, ns, o);
  if ((threadIdx.x & 31) == 0 && ns) atomicAdd(totalUpdates, ns);
}
__global__ void k_last_offset(const int* off, int n, int* out) { *out = off[n]; }

int schedule_alloc(Schedule& s, int cap) {
  s.cap = cap;
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, s.keys, s.keysSorted, s.iota, s.order, cap);
  cub::DeviceScan::ExclusiveSum(nullptr, b2, s.chunkCnt, s.chunkOff, cap + 1);
  s.tempBytes = (b1 > b2 ? b1 : b2) + 256;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  A((void**)&s.keys, cap * sizeof(unsigned)); A((void**)&s.keysSorted, cap * sizeof(unsigned));
  A((void**)&s.iota, cap * sizeof(int)); A((void**)&s.order, cap * sizeof(int));
  A((void**)&s.chunkCnt, (cap + 1) * sizeof(int)); A((void**)&s.chunkOff, (cap + 1) * sizeof(int));
  A(&s.temp, s.tempBytes);
  A(&s.temp2, s.tempBytes);
  return e == cudaSuccess ? 0 : -1;
}
void schedule_free(Schedule& s) {
  void* ps[] = {s.keys, s.keysSorted, s.iota, s.order, s.chunkCnt, s.chunkOff, s.temp, s.temp2};
  for (void* p : ps) if (p) cudaFree(p);
  s = Schedule{};
}
void launch_schedule_order(cudaStream_t st, const LatticeDev& L, const InstanceSoA& in, int nInst, Schedule& s) {
  const int cap = s.cap;
  k_morton_keys<<<(cap + 255) / 256, 256, 0, st>>>(L, in, nInst, cap, s.keys, s.iota);
  size_t tb = s.tempBytes;
  cub::DeviceRadixSort::SortPairs(s.temp2, tb, s.keys, s.keysSorted, s.iota, s.order, cap, 0, 32, st);
}
void launch_schedule_chunks(cudaStream_t st, const InstanceSoA& in, int nInst, Schedule& s, int* dTotalChunks, int* dTotalUpdates) {
  cudaMemsetAsync(dTotalUpdates, 0, sizeof(int), st);
  k_chunk_counts<<<(nInst + 256) / 256, 256, 0, st>>>(in, s.order, nInst, s.chunkCnt, dTotalUpdates, s.chunk);
  size_t tb = s.tempBytes;
  cub::DeviceScan::ExclusiveSum(s.temp, tb, s.chunkCnt, s.chunkOff, nInst + 1, st);
  k_last_offset<<<1, 1, 0, st>>>(s.chunkOff, nInst, dTotalChunks);
}
void launch_build_schedule(cudaStream_t st, const LatticeDev& L, const InstanceSoA& in, int nInst, Schedule& s, int* dTotalChunks,
                           int* dTotalUpdates) {
  launch_schedule_order(st, L, in, nInst, s);
  launch_schedule_chunks(st, in, nInst, s, dTotalChunks, dTotalUpdates);
}

void launch_keep_flags(cudaStream_t st, const InstanceSoA& in, int nInst, int* keep) {
  if (nInst > 0) k_keep_flags<<<(nInst + 256) / 256, 256, 0, st>>>(in, nInst, keep);
}
void launch_compact_instances(cudaStream_t st, const InstanceSoA& src, const InstanceSoA& dst, const int* keep, const int* keepOff,
                              int nInst) {
  if (nInst > 0) k_compact_instances<<<(nInst + 255) / 256, 256, 0, st>>>(src, dst, keep, keepOff, nInst);
}

void launch_exclusive_scan(cudaStream_t st, const int* in, int* out, int n, int* total) {
  k_exclusive_scan<<<1, 1024, 0, st>>>(in, out, n, total);
}
// in[n] must be a zero sentinel and out hold n + 1 entries: out[n] is the total (device-wide scan, ~6 us at n = 1e5 where the
// single-block kernel above takes 100 us)
void launch_exclusive_scan_total(cudaStream_t st, Schedule& s, const int* in, int* out, int n, int* total) {
  size_t tb = s.tempBytes;
  cub::DeviceScan::ExclusiveSum(s.temp, tb, in, out, n + 1, st);
  k_last_offset<<<1, 1, 0, st>>>(out, n, total);
}
static int grid_for(size_t n, int block, int maxBlocks = 148 * 16) {
  size_t g = (n + block - 1) / block;
  if (g > (size_t)maxBlocks) g = maxBlocks;
  if (g < 1) g = 1;
  return (int)g;
}
void launch_aos_to_soa(cudaStream_t st, const LatticeDev& L, const double* aos, int nComp, double* soa) {
  k_aos_to_soa<<<grid_for(L.nCells, 256), 256, 0, st>>>(L, aos, nComp, soa);
}
void launch_soa_to_aos(cudaStream_t st, const LatticeDev& L, const double* soa, int nComp, double* aos) {
  k_soa_to_aos<<<grid_for(L.nCells, 256), 256, 0, st>>>(L, soa, nComp, aos);
}
__global__ void k_copy_one_int(int* dst, const int* src) { *dst = *src; }
void launch_copy_int(cudaStream_t st, int* dst, const int* src) { k_copy_one_int<<<1, 1, 0, st>>>(dst, src); }

void launch_clear_shell_cells(cudaStream_t st, const int* shellCell, int total, size_t C, double* F) {
  if (total > 0) k_clear_shell_cells<<<grid_for(total, 256), 256, 0, st>>>(shellCell, total, C, F);
}

}  // namespace ibm
