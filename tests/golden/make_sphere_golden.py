"""Golden drag history of a coarse validation/sphere: mini flow solver (PeriodicBox3D) + the REFERENCE'S OWN particleCloud
(oracle/_ref).  Run in the build container:  python tests/golden/make_sphere_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ibm_testutil as util  # noqa: E402
from ibmfoam_b200.validation import mini_flow  # noqa: E402

N, N_STEPS = 24, 40

if __name__ == "__main__":
    case, flow = util.sphere_box_setup(n=N)
    cloud = util.OracleAsCloud(case, impl="reference")
    hist = mini_flow.run_sphere(cloud, flow, N_STEPS, every=1)
    out = os.path.join(HERE, "sphere_box_cd_history.npz")
    np.savez_compressed(out, n=N, cd=np.array([c for _, c, _ in hist]), body_force=np.array([f for _, _, f in hist]))
    print(out, os.path.getsize(out), "bytes; Cd(1) =", hist[0][1], "Cd(%d) =" % N_STEPS, hist[-1][1])
