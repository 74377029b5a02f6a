// TEST INFRASTRUCTURE — the two cub device algorithms the library uses, for the CPU emulation.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* temp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int beginBit = 0,
                               int endBit = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!temp) { bytes = 16; return cudaSuccess; }
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    const unsigned long long mask = (endBit - beginBit >= 64) ? ~0ull : (((1ull << (endBit - beginBit)) - 1ull) << beginBit);
    ::emu::submit([=]() mutable {
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ((unsigned long long)kin[a] & mask) < ((unsigned long long)kin[b] & mask); });
      std::vector<K> k((size_t)n);
      std::vector<V> v((size_t)n);
      for (int i = 0; i < n; ++i) { k[(size_t)i] = kin[idx[(size_t)i]]; v[(size_t)i] = vin[idx[(size_t)i]]; }
      for (int i = 0; i < n; ++i) { kout[i] = k[(size_t)i]; vout[i] = v[(size_t)i]; }
    });
    return cudaSuccess;
  }
};
struct DeviceScan {
  template <class I, class O>
  static cudaError_t ExclusiveSum(void* temp, size_t& bytes, const I* in, O* out, int n, cudaStream_t = nullptr) {
    if (!temp) { bytes = 16; return cudaSuccess; }
    ::emu::submit([=]() {
      O acc = 0;
      for (int i = 0; i < n; ++i) { const O x = (O)in[i]; out[i] = acc; acc += x; }
    });
    return cudaSuccess;
  }
};
}  // namespace cub
