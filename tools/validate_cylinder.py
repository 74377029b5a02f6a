#!/usr/bin/env python
"""validation/cylinderRe30 to steady state: mini flow solver (ibmfoam_b200/validation/mini_flow.py) + the IB path.

    python tools/validate_cylinder.py [--steps 8000] [--checker port|reference]

Default: the CUDA path (needs a B200).  --checker runs the CPU checker instead (test infrastructure; the restatement is
bit-identical to the reference's own code on this case, tests/test_validation_cases.py) - that is how the number in
DESIGN.md §8 was produced while no GPU was available.  Literature anchor: Cd ~ 1.74 for an unbounded cylinder at Re = 30;
the case is a transversally periodic row with 10 % blockage and only 5 d of upstream length, which raises it to ~2.1."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8000)
    ap.add_argument("--checker", choices=["port", "reference"], default=None)
    args = ap.parse_args()
    import ibm_testutil as util
    from ibmfoam_b200.validation import mini_flow
    case, flow, to_caller, from_caller = util.cylinder_re30_setup()
    if args.checker:
        cloud = util.OracleAsCloud(case, impl=args.checker)
    else:
        from ibmfoam_b200.cloud import ParticleCloud
        cloud = ParticleCloud(case)
    t0 = time.time()
    cds = mini_flow.run_cylinder(cloud, flow, args.steps, to_caller, from_caller, every=max(1, args.steps // 20),
                                 log=lambda s, c: print(f"step {s:6d}  t = {s * case.dt:6.3f} s  Cd = {c:.6f}", flush=True))
    print(f"Cd = {cds[-1][1]:.4f} after {args.steps} steps ({time.time() - t0:.0f} s wall); max |u| = {np.abs(flow.u).max():.4f}")


if __name__ == "__main__":
    main()
