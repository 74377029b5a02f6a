// TEST INFRASTRUCTURE — grid-wide synchronisation for the CPU emulation: a cooperative launch (emu::coop_launch, reached
// through ibm::launch_cooperative of ibm_ctx.h) keeps every block of the grid resident, and grid.sync() parks a fiber until
// all live threads of all blocks have arrived.
#pragma once
#include "emu.hpp"
namespace cooperative_groups {
struct grid_group {
  void sync() const { ::emu::gridsync(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
