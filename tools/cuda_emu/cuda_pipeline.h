// TEST INFRASTRUCTURE — <cuda_pipeline.h> for the CPU emulation.  A copy is only PERFORMED when the issuing thread
// waits for it (__pipeline_wait_prior), so data read before the wait + barrier is stale and the test fails.
#pragma once
#include "emu.hpp"
inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n) { ::emu::cur().pending.push_back({dst, src, n}); }
inline void __pipeline_commit() {}
inline void __pipeline_wait_prior(int) {
  auto& q = ::emu::cur().pending;
  for (const auto& c : q) std::memcpy(c.dst, c.src, c.n);
  q.clear();
}
