// TEST INFRASTRUCTURE — a minimal CPU emulation of the CUDA execution model, enough to run the sm_100a kernels of
// ibmfoam_b200/csrc/ibm_kernels_v5.cuh on a machine without a GPU and check their LOGIC (indexing, staging boxes,
// hashing, barrier placement, reductions) against the oracle.  It says nothing about performance and nothing about
// hardware memory-model effects; the GPU parity tests remain the gate.
//
// Model: blocks run one after another (all at once for a cooperative launch); the threads of a block are fibers (ucontext) scheduled round-robin by one OS
// thread.  __syncthreads() parks a fiber until every live fiber of the block has arrived; warp shuffles park a fiber
// until all 32 lanes of its warp have deposited their value.  A fiber runs undisturbed between two such points, so a
// missing barrier shows up as a deterministic wrong answer (thread 0 runs far ahead of thread 255).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

namespace emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct PendingCopy { void* dst; const void* src; size_t n; };

struct GroupBarrier { unsigned arrived = 0; long gen = 0; };  // rendez-vous of the lanes named by one mask

struct Fiber {
  std::vector<PendingCopy> pending;  // cp.async copies issued and not yet waited for (performed at the wait)
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;  // 0 runnable, 1 at block barrier, 2 at warp-level rendez-vous, 3 done, 4 at the grid barrier
  const GroupBarrier* waitObj = nullptr;
  long warpGen = 0;
  Dim3 tid;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  Dim3 blockIdx, blockDim, gridDim;
  std::vector<unsigned char> dynSmem;
  std::vector<std::map<unsigned, GroupBarrier>> groups;  // per warp: one rendez-vous object per participation mask
  std::vector<double> shflBuf;     // [warp][32] exchange slots (doubles hold ints exactly)
  const std::function<void()>* body = nullptr;
  int orAcc = 0;  // __syncthreads_or accumulator
};

inline Block*& blk() { static Block* b = nullptr; return b; }
inline Fiber& cur() { return blk()->fibers[blk()->cur]; }
inline void* dyn_smem() { return blk()->dynSmem.data(); }

inline void yield() { Block* b = blk(); swapcontext(&b->fibers[b->cur].ctx, &b->sched); }

inline void syncthreads() {
  cur().state = 1;
  yield();
}

// block-wide OR of a predicate (three rendez-vous: deposit, read, reset)
inline int syncthreads_or(int pred) {
  Block* b = blk();
  if (pred) b->orAcc = 1;
  syncthreads();
  const int r = b->orAcc;
  syncthreads();
  b->orAcc = 0;
  syncthreads();
  return r;
}

inline int lane_id() { const Fiber& f = cur(); return (int)(f.tid.x % 32); }
inline int warp_id() { const Fiber& f = cur(); return (int)(f.tid.x / 32); }

inline unsigned existing_lanes() {
  Block* b = blk();
  const int lanes = std::min<int>(32, (int)b->blockDim.x - 32 * warp_id());
  return lanes >= 32 ? 0xffffffffu : ((1u << lanes) - 1u);
}
inline int warp_lanes() { return __builtin_popcount(existing_lanes()); }

// the lanes named in `mask` rendez-vous (sub-warp groups with disjoint masks proceed independently, as on the hardware;
// a lane that is named but has exited, or never arrives, is reported as a deadlock by the scheduler)
inline void group_barrier(unsigned mask) {
  Block* b = blk();
  Fiber& f = cur();
  mask &= existing_lanes();
  GroupBarrier& g = b->groups[warp_id()][mask];
  g.arrived |= 1u << lane_id();
  if ((g.arrived & mask) == mask) {
    g.arrived = 0;
    ++g.gen;
    return;
  }
  f.state = 2;
  f.waitObj = &g;
  f.warpGen = g.gen;
  yield();
}
inline void warp_barrier() { group_barrier(0xffffffffu); }

// every named lane deposits a predicate, all read the word (ballot)
inline unsigned ballot(unsigned mask, bool pred) {
  Block* b = blk();
  const int w = warp_id(), l = lane_id();
  mask &= existing_lanes();
  b->shflBuf[32 * w + l] = pred ? 1.0 : 0.0;
  group_barrier(mask);
  unsigned word = 0;
  for (int s = 0; s < 32; ++s)
    if (((mask >> s) & 1u) && b->shflBuf[32 * w + s] != 0.0) word |= (1u << s);
  group_barrier(mask);
  return word;
}

inline double shfl(unsigned mask, double v, int srcLane) {
  Block* b = blk();
  const int w = warp_id(), l = lane_id();
  mask &= existing_lanes();
  b->shflBuf[32 * w + l] = v;
  group_barrier(mask);
  const double r = (srcLane >= 0 && srcLane < 32 && ((mask >> srcLane) & 1u)) ? b->shflBuf[32 * w + srcLane] : v;
  group_barrier(mask);
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

// one scheduling round over the fibers of block b (blk() must point at it): returns whether anything moved
inline bool block_round(Block& b, const std::vector<unsigned>& threadOrder, unsigned& done) {
  const unsigned blockX = b.blockDim.x;
  bool progressed = false;
  unsigned atBarrier = 0, live = 0;
  for (unsigned to = 0; to < blockX; ++to) {
    const unsigned t = threadOrder[to];
    Fiber& f = b.fibers[t];
    if (f.state == 3) continue;
    if (f.state == 2 && f.waitObj->gen != f.warpGen) f.state = 0;
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
  return progressed;
}

inline void block_arm(Block& b, unsigned bx, size_t stack) {
  const int nWarps = (int)(b.blockDim.x + 31) / 32;
  b.blockIdx.x = bx;
  b.groups.assign((size_t)nWarps, {});
  b.shflBuf.assign(32 * (size_t)nWarps, 0.0);
  for (unsigned t = 0; t < b.blockDim.x; ++t) {
    Fiber& f = b.fibers[t];
    f.state = 0;
    f.tid.x = t;
    f.pending.clear();
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = stack;
    f.ctx.uc_link = &b.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
}

// grid-wide barrier of a cooperative launch (state 4): released by launch(..., cooperative) once every live thread of
// every block has arrived
inline void gridsync() {
  cur().state = 4;
  yield();
}

// run `body` as a kernel: grid.x blocks of block.x threads (1-D is all the kernels here use).  cooperative = all blocks
// are resident at once (their fibers are scheduled block after block, round after round) so that grid.sync() works;
// static __shared__ storage is one object for the whole process here, so a cooperative kernel may only use dynamic
// shared memory.
inline void launch(unsigned gridX, unsigned blockX, size_t smemBytes, const std::function<void()>& body, bool cooperative = false) {
  static const size_t STACK = 192 * 1024;
  // EMU_ORDER=reverse | shuffle:<seed> perturbs the order in which blocks run and in which the fibers of a block are
  // resumed: a kernel whose result depends on it (beyond the rounding of floating-point atomics) has a race
  static const char* orderEnv = std::getenv("EMU_ORDER");
  const bool reverse = orderEnv && std::strcmp(orderEnv, "reverse") == 0;
  const bool shuffle = orderEnv && std::strncmp(orderEnv, "shuffle", 7) == 0;
  static unsigned long long rng = (shuffle && std::strlen(orderEnv) > 8) ? std::strtoull(orderEnv + 8, nullptr, 10) * 2654435761ull + 1 : 88172645463325252ull;
  auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  std::vector<unsigned> blockOrder(gridX), threadOrder(blockX);
  for (unsigned i = 0; i < gridX; ++i) blockOrder[i] = reverse ? gridX - 1 - i : i;
  for (unsigned i = 0; i < blockX; ++i) threadOrder[i] = reverse ? blockX - 1 - i : i;
  if (shuffle) {
    for (unsigned i = gridX; i > 1; --i) std::swap(blockOrder[i - 1], blockOrder[next() % i]);
  }
  const unsigned resident = cooperative ? gridX : 1;
  std::vector<Block> blocks(resident);
  for (auto& b : blocks) {
    b.gridDim.x = gridX;
    b.blockDim.x = blockX;
    b.fibers.resize(blockX);
    for (auto& f : b.fibers) f.stack = (char*)std::malloc(STACK);
    b.dynSmem.assign(smemBytes + 128, 0);
    b.body = &body;
  }
  if (cooperative) {
    std::vector<unsigned> done(gridX, 0);
    for (unsigned bo = 0; bo < gridX; ++bo) { blk() = &blocks[bo]; block_arm(blocks[bo], blockOrder[bo], STACK); }
    unsigned finished = 0;
    while (finished < gridX) {
      bool progressed = false;
      for (unsigned bo = 0; bo < gridX; ++bo) {
        Block& b = blocks[bo];
        if (done[bo] >= blockX) continue;
        if (shuffle)
          for (unsigned i = blockX; i > 1; --i) std::swap(threadOrder[i - 1], threadOrder[next() % i]);
        blk() = &b;
        if (block_round(b, threadOrder, done[bo])) progressed = true;
        if (done[bo] >= blockX) ++finished;
      }
      size_t live = 0, atGrid = 0;
      for (auto& b : blocks)
        for (auto& f : b.fibers) {
          if (f.state != 3) ++live;
          if (f.state == 4) ++atGrid;
        }
      if (live > 0 && atGrid == live) {
        for (auto& b : blocks)
          for (auto& f : b.fibers)
            if (f.state == 4) f.state = 0;
        progressed = true;
      }
      if (!progressed) {
        std::fprintf(stderr, "cuda_emu: deadlock in a cooperative launch (%zu live threads, %zu at grid.sync)\n", live, atGrid);
        std::abort();
      }
    }
  } else {
    Block& b = blocks[0];
    blk() = &b;
    for (unsigned bo = 0; bo < gridX; ++bo) {
      const unsigned bx = blockOrder[bo];
      if (shuffle)
        for (unsigned i = blockX; i > 1; --i) std::swap(threadOrder[i - 1], threadOrder[next() % i]);
      // shared memory is NOT cleared between blocks: like the hardware, a kernel must not rely on its initial content
      block_arm(b, bx, STACK);
      unsigned done = 0;
      while (done < blockX) {
        if (!block_round(b, threadOrder, done)) {
          unsigned live = 0, atBarrier = 0;
          for (auto& f : b.fibers) { if (f.state != 3) ++live; if (f.state == 1) ++atBarrier; }
          std::fprintf(stderr, "cuda_emu: deadlock in block %u (%u live, %u at __syncthreads, the rest at a warp barrier%s)\n", bx, live,
                       atBarrier, "");
          std::abort();
        }
      }
    }
  }
  for (auto& b : blocks)
    for (auto& f : b.fibers) std::free(f.stack);
  blk() = nullptr;
}

}  // namespace emu
