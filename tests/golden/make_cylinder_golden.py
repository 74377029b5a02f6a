"""Golden drag history of validation/cylinderRe30: the mini flow solver (ibmfoam_b200/validation/mini_flow.py) coupled to
the REFERENCE'S OWN particleCloud (oracle/_ref).  Run in the build container:  python tests/golden/make_cylinder_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ibm_testutil as util  # noqa: E402
from ibmfoam_b200.validation import mini_flow  # noqa: E402

N_STEPS = 150

if __name__ == "__main__":
    case, flow, to_caller, from_caller = util.cylinder_re30_setup()
    cloud = util.OracleAsCloud(case, impl="reference")
    cds = mini_flow.run_cylinder(cloud, flow, N_STEPS, to_caller, from_caller, every=1)
    out = os.path.join(HERE, "cylinder_re30_cd_history.npz")
    probes = np.linspace(0, flow.u.size - 1, 64).astype(np.int64)   # a few cells of the final flow field
    np.savez_compressed(out, steps=np.array([s for s, _ in cds]), cd=np.array([c for _, c in cds]), probes=probes,
                        u_probe=flow.u.ravel()[probes], v_probe=flow.v.ravel()[probes], p_probe=flow.p.ravel()[probes])
    print(out, os.path.getsize(out), "bytes; Cd(1) =", cds[0][1], "Cd(%d) =" % N_STEPS, cds[-1][1])
