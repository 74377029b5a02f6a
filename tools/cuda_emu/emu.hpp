// TEST INFRASTRUCTURE — a minimal CPU emulation of the CUDA execution model, enough to run the sm_100a kernels of
// ibmfoam_b200/csrc/ibm_kernels_v5.cuh on a machine without a GPU and check their LOGIC (indexing, staging boxes,
// hashing, barrier placement, reductions) against the oracle.  It says nothing about performance and nothing about
// hardware memory-model effects; the GPU parity tests remain the gate.
//
// Model: blocks run one after another; the threads of a block are fibers (ucontext) scheduled round-robin by one OS
// thread.  __syncthreads() parks a fiber until every live fiber of the block has arrived; warp shuffles park a fiber
// until all 32 lanes of its warp have deposited their value.  A fiber runs undisturbed between two such points, so a
// missing barrier shows up as a deterministic wrong answer (thread 0 runs far ahead of thread 255).
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct PendingCopy { void* dst; const void* src; size_t n; };

struct Fiber {
  std::vector<PendingCopy> pending;  // cp.async copies issued and not yet waited for (performed at the wait)
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;  // 0 runnable, 1 at block barrier, 2 at warp barrier, 3 done
  long warpGen = 0;
  Dim3 tid;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  Dim3 blockIdx, blockDim, gridDim;
  std::vector<unsigned char> dynSmem;
  std::vector<long> warpGen;       // generation counter per warp
  std::vector<int> warpArrived;
  std::vector<double> shflBuf;     // [warp][32] exchange slots (doubles hold ints exactly)
  const std::function<void()>* body = nullptr;
};

inline Block*& blk() { static Block* b = nullptr; return b; }
inline Fiber& cur() { return blk()->fibers[blk()->cur]; }
inline void* dyn_smem() { return blk()->dynSmem.data(); }

inline void yield() { Block* b = blk(); swapcontext(&b->fibers[b->cur].ctx, &b->sched); }

inline void syncthreads() {
  cur().state = 1;
  yield();
}

inline int lane_id() { const Fiber& f = cur(); return (int)(f.tid.x % 32); }
inline int warp_id() { const Fiber& f = cur(); return (int)(f.tid.x / 32); }

// all 32 lanes of the warp (or all lanes that exist, for a partial last warp) rendez-vous
inline void warp_barrier() {
  Block* b = blk();
  Fiber& f = cur();
  const int w = warp_id();
  const int lanes = std::min<int>(32, (int)b->blockDim.x - 32 * w);
  const long g = b->warpGen[w];
  if (++b->warpArrived[w] == lanes) {
    b->warpArrived[w] = 0;
    ++b->warpGen[w];
    return;
  }
  f.state = 2;
  f.warpGen = g;
  yield();
}

inline double shfl(double v, int srcLane) {
  Block* b = blk();
  const int w = warp_id(), l = lane_id();
  b->shflBuf[32 * w + l] = v;
  warp_barrier();
  const int lanes = std::min<int>(32, (int)b->blockDim.x - 32 * w);
  const double r = (srcLane >= 0 && srcLane < lanes) ? b->shflBuf[32 * w + srcLane] : v;
  warp_barrier();
  return r;
}

inline void fiber_entry() {
  Block* b = blk();
  (*b->body)();
  if (!b->fibers[b->cur].pending.empty()) {
    std::fprintf(stderr, "cuda_emu: thread %u exits with asynchronous copies it never waited for\n", b->fibers[b->cur].tid.x);
    std::abort();
  }
  b->fibers[b->cur].state = 3;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

// run `body` as a kernel: grid.x blocks of block.x threads (1-D is all the kernels here use)
inline void launch(unsigned gridX, unsigned blockX, size_t smemBytes, const std::function<void()>& body) {
  static const size_t STACK = 192 * 1024;
  Block b;
  blk() = &b;
  b.gridDim.x = gridX;
  b.blockDim.x = blockX;
  b.fibers.resize(blockX);
  for (auto& f : b.fibers) f.stack = (char*)std::malloc(STACK);
  b.dynSmem.assign(smemBytes + 128, 0);
  const int nWarps = (int)(blockX + 31) / 32;
  b.body = &body;
  for (unsigned bx = 0; bx < gridX; ++bx) {
    b.blockIdx.x = bx;
    b.warpGen.assign(nWarps, 0);
    b.warpArrived.assign(nWarps, 0);
    b.shflBuf.assign(32 * (size_t)nWarps, 0.0);
    // shared memory is NOT cleared between blocks: like the hardware, a kernel must not rely on its initial content
    for (unsigned t = 0; t < blockX; ++t) {
      Fiber& f = b.fibers[t];
      f.state = 0;
      f.tid.x = t;
      f.pending.clear();
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = STACK;
      f.ctx.uc_link = &b.sched;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned done = 0;
    while (done < blockX) {
      bool progressed = false;
      unsigned atBarrier = 0, live = 0;
      for (unsigned t = 0; t < blockX; ++t) {
        Fiber& f = b.fibers[t];
        if (f.state == 3) continue;
        if (f.state == 2 && b.warpGen[t / 32] != f.warpGen) f.state = 0;
        if (f.state == 0) {
          b.cur = (int)t;
          swapcontext(&b.sched, &f.ctx);
          progressed = true;
          if (f.state == 3) ++done;
        }
      }
      for (unsigned t = 0; t < blockX; ++t) {
        const Fiber& f = b.fibers[t];
        if (f.state != 3) ++live;
        if (f.state == 1) ++atBarrier;
      }
      if (live > 0 && atBarrier == live) {  // every live thread is at the barrier: release
        for (auto& f : b.fibers)
          if (f.state == 1) f.state = 0;
        progressed = true;
      }
      if (!progressed) {
        std::fprintf(stderr, "cuda_emu: deadlock in block %u (%u live, %u at __syncthreads, the rest at a warp barrier)\n", bx, live,
                     atBarrier);
        std::abort();
      }
    }
  }
  for (auto& f : b.fibers) std::free(f.stack);
  blk() = nullptr;
}

}  // namespace emu
