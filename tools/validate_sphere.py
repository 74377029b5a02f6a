#!/usr/bin/env python
"""A fixed sphere in a periodic box (validation/sphere's mesh: every patch cyclic; d = 0.004 in a 0.02 box, U_inf = 0.1)
run to steady state with the mini flow solver + the IB path, against Schiller-Naumann  Cd = 24/Re (1 + 0.15 Re^0.687).

    python tools/validate_sphere.py [--n 48] [--Re 20] [--steps 2400] [--checker port|reference]

Default: the CUDA path (needs a B200); --checker runs the CPU checker (test infrastructure).  Two drag estimates are
printed: the reference's surface-stress integral (particleIBM.C:541-581, what `Cd()` returns) and the momentum balance of
the periodic box (the uniform body force that holds the mean flow equals the drag in steady state)."""
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
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--Re", type=float, default=20.0)
    ap.add_argument("--steps", type=int, default=2400)
    ap.add_argument("--checker", choices=["port", "reference"], default=None)
    args = ap.parse_args()
    import ibm_testutil as util
    from ibmfoam_b200.validation import mini_flow
    case, flow = util.sphere_box_setup(n=args.n, Re=args.Re)
    if args.checker:
        cloud = util.OracleAsCloud(case, impl=args.checker)
    else:
        from ibmfoam_b200.cloud import ParticleCloud
        cloud = ParticleCloud(case)
    L, d, U = 0.02, 0.004, 0.1
    A = np.pi * (d / 2) ** 2
    sn = 24.0 / args.Re * (1.0 + 0.15 * args.Re ** 0.687)
    t0 = time.time()

    def log(step, cd, f):
        print(f"step {step:6d}  t = {step * case.dt:7.4f} s  Cd(stress integral) = {cd:.4f}  Cd(momentum balance) = {f[1] * L ** 3 / (0.5 * U * U * A):.4f}",
              flush=True)

    hist = mini_flow.run_sphere(cloud, flow, args.steps, every=max(1, args.steps // 20), log=log)
    cd, f = hist[-1][1], hist[-1][2]
    print(f"Re = {args.Re:g}, d/h = {d / (L / args.n):.1f}: Cd(stress integral) = {cd:.3f}, Cd(momentum balance) = {f[1] * L ** 3 / (0.5 * U * U * A):.3f}, "
          f"Schiller-Naumann = {sn:.3f}  ({time.time() - t0:.0f} s wall)")


if __name__ == "__main__":
    main()
