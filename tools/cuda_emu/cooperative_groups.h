// TEST INFRASTRUCTURE — grid-wide synchronisation needs all blocks resident at once; the emulation runs blocks one after
// another, so a cooperative launch is refused (cudaLaunchCooperativeKernel returns an error) and this only has to compile.
#pragma once
#include <cstdio>
#include <cstdlib>
namespace cooperative_groups {
struct grid_group {
  void sync() const { std::fprintf(stderr, "cuda_emu: grid.sync() is not emulated\n"); std::abort(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
