"""Golden trajectory of two colliding spheres coupled to the mini flow solver, from the REFERENCE'S OWN particleCloud.
Run in the build container:  python tests/golden/make_colliding_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ibm_testutil as util  # noqa: E402
from ibmfoam_b200.validation import mini_flow  # noqa: E402

N_STEPS = 16

if __name__ == "__main__":
    case, flow = util.colliding_spheres_setup()
    cloud = util.OracleAsCloud(case, impl="reference")
    hist = mini_flow.run_moving_bodies(cloud, flow, N_STEPS, state=util.kinematics)
    out = os.path.join(HERE, "colliding_spheres_coupled.npz")
    np.savez_compressed(out, history=np.array(hist))
    h = np.array(hist)
    print(out, os.path.getsize(out), "bytes; v_x of sphere 0:", h[:, 0, 3])
