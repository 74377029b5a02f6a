"""validation/cylinderRe30 end to end on the GPU: mini flow solver + the CUDA path through the host C ABI, against the
drag history the same solver produces with the REFERENCE'S OWN particleCloud (tests/golden/cylinder_re30_cd_history.npz).
Tolerance: 1e-8 relative on Cd over 150 coupled steps (force/torque are 1e-10 relative per call and the forcing field
1e-12; the flow feeds them back).  Written without access to a GPU; sorts last for that reason."""
import os

import numpy as np
import pytest

from ibmfoam_b200.cloud import ParticleCloud
from ibmfoam_b200.validation import mini_flow
import golden_util as gu
import ibm_testutil as util

pytestmark = pytest.mark.gpu


def test_cylinder_re30_drag_history_matches_reference():
    g = np.load(os.path.join(gu.GOLDEN, "cylinder_re30_cd_history.npz"))
    case, flow, to_caller, from_caller = util.cylinder_re30_setup()
    cloud = ParticleCloud(case)
    n = len(g["cd"])
    cds = mini_flow.run_cylinder(cloud, flow, n, to_caller, from_caller, every=1)
    cd = np.array([c for _, c in cds])
    rel = np.abs(cd - g["cd"]) / np.abs(g["cd"])
    assert np.max(rel) < 1e-8, float(np.max(rel))
    pr = g["probes"]
    for name, a in (("u", flow.u), ("v", flow.v), ("p", flow.p)):
        ref = g[name + "_probe"]
        assert np.max(np.abs(a.ravel()[pr] - ref)) < 1e-8 * max(np.max(np.abs(ref)), 1e-3), name
    cloud.close()


def test_sphere_box_drag_history_matches_reference():
    g = np.load(os.path.join(gu.GOLDEN, "sphere_box_cd_history.npz"))
    case, flow = util.sphere_box_setup(n=int(g["n"]))
    cloud = ParticleCloud(case)
    hist = mini_flow.run_sphere(cloud, flow, len(g["cd"]), every=1)
    cd = np.array([c for _, c, _ in hist])
    bf = np.array([f for _, _, f in hist])
    assert np.max(np.abs(cd - g["cd"]) / np.abs(g["cd"])) < 1e-8
    assert np.max(np.abs(bf - g["body_force"])) < 1e-8 * np.max(np.abs(g["body_force"]))
    cloud.close()


def test_colliding_spheres_coupled_trajectory_matches_reference():
    """validation/collidingParticles-type loop on the GPU.  Stated tolerance over the 16 fluid steps (1600 sub-steps):
    1e-8 relative on centres / velocities / angular velocities, 1e-7 on force / torque (relative to the largest entry)."""
    g = np.load(os.path.join(gu.GOLDEN, "colliding_spheres_coupled.npz"))["history"]
    case, flow = util.colliding_spheres_setup()
    cloud = ParticleCloud(case)
    hist = np.array(mini_flow.run_moving_bodies(cloud, flow, len(g), state=util.kinematics))
    assert hist.shape == g.shape
    for sl, tol in ((slice(0, 3), 1e-8), (slice(3, 6), 1e-8), (slice(6, 9), 1e-8), (slice(9, 15), 1e-7)):
        ref = g[:, :, sl]
        assert np.max(np.abs(hist[:, :, sl] - ref)) < tol * np.max(np.abs(ref)), sl
    cloud.close()
