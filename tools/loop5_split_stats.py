#!/usr/bin/env python
"""Active fraction and fallback count of the loop-5 quiet/active split on the benchmark's assembly (solid fraction 0.2,
d = 8 cells, Gaussian velocities), measured by RUNNING the CUDA sources under the CPU emulation (tools/cuda_emu) - the
numbers DESIGN.md §4 quotes.  Statistics of the classification only: nothing here says anything about time.

    python tools/loop5_split_stats.py [box=256] [steps=12] [safety_ppm=2000000]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))


def main(n=256, steps=12, safety_ppm=2000000):
    import build_emu
    from ibmfoam_b200 import capi
    capi.LIB_PATH = build_emu.build()          # test infrastructure: the emulated library, never the product's loader default
    from ibmfoam_b200 import cases
    from ibmfoam_b200.cloud import ParticleCloud
    t0 = time.time()
    case = cases.assembly(n=n, phi=0.2, moving=True, cyclic_without_walls=True)
    gc = ParticleCloud(case)
    for k, v in (("graph_loop5", 0), ("fused_loop5", 0), ("active_loop5", 1), ("fused_active_loop5", 1), ("mark_variant", 1),
                 ("active_loop5_safety_ppm", safety_ppm)):
        gc.set_option(k, v)
    print(f"box {n}^3, {case.particles.n} spheres, safety {safety_ppm / 1e6}")
    for step in range(steps):
        gc.combineForces()
        gc.advance(case.dt)
        inst, act = gc.nParticles(), gc.counter("loop5_active_instances")
        print(f"step {step:3d}  instances {inst:7d}  active {act:7d} ({100.0 * act / inst:5.1f} %)  fallbacks so far "
              f"{gc.counter('loop5_fallbacks')}  [{time.time() - t0:6.1f} s]", flush=True)
    gc.close()


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
