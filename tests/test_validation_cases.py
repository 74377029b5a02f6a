"""validation/cylinderRe30 end to end on the CPU: the mini flow solver (ibmfoam_b200/validation/mini_flow.py, the caller
either side of the IB path) coupled to the CPU restatement must reproduce, bit for bit, the drag history it produces
when coupled to the REFERENCE'S OWN particleCloud - live (oracle/_ref) and from the committed golden history."""
import os

import numpy as np
import pytest

from ibmfoam_b200.validation import mini_flow
from oracle import pyoracle
import golden_util as gu
import ibm_testutil as util

GOLD = os.path.join(gu.GOLDEN, "cylinder_re30_cd_history.npz")


def _run(impl, n_steps):
    case, flow, to_caller, from_caller = util.cylinder_re30_setup()
    cloud = util.OracleAsCloud(case, impl=impl)
    cds = mini_flow.run_cylinder(cloud, flow, n_steps, to_caller, from_caller, every=1)
    return np.array([c for _, c in cds]), flow


def test_restatement_reproduces_reference_drag_history():
    g = np.load(GOLD)
    n = 60
    cd, flow = _run("port", n)
    assert np.array_equal(cd, g["cd"][:n])
    assert 2.0 < cd[-1] < 5.0 and cd[0] > 10 * cd[-1], "impulsive start: a large added-mass spike that relaxes within tens of steps"


@pytest.mark.skipif(not pyoracle.have_reference(), reason="no oracle/_ref/libibm_ref.so and no /root/reference")
def test_compiled_reference_matches_its_golden_history():
    g = np.load(GOLD)
    n = 25
    cd, _ = _run("reference", n)
    assert np.array_equal(cd, g["cd"][:n])


def test_uniform_flow_without_a_body_stays_uniform():
    """the flow solver alone: no forcing, uniform inflow -> nothing happens (guards the Poisson / ghost-cell plumbing)"""
    flow = mini_flow.Channel2D(40, 20, 0.05, 0.3, 1e-3, 1e-3)
    for _ in range(5):
        Sx, Sy = flow.explicit_terms()
        flow.project(flow.u - flow.dt * Sx, flow.v - flow.dt * Sy)
    assert np.allclose(flow.u, 0.3, atol=1e-13) and np.allclose(flow.v, 0.0, atol=1e-13) and np.allclose(flow.p, 0.0, atol=1e-10)


def test_sphere_in_periodic_box_restatement_reproduces_reference_history():
    """3-D: coarse validation/sphere (periodic box, mean flow held by a body force), 40 coupled steps"""
    g = np.load(os.path.join(gu.GOLDEN, "sphere_box_cd_history.npz"))
    case, flow = util.sphere_box_setup(n=int(g["n"]))
    cloud = util.OracleAsCloud(case)
    hist = mini_flow.run_sphere(cloud, flow, len(g["cd"]), every=1)
    assert np.array_equal([c for _, c, _ in hist], g["cd"])
    assert np.array_equal([f for _, _, f in hist], g["body_force"])
    assert hist[-1][2][1] > 0, "a positive body force along the mean flow balances the drag"


def test_colliding_spheres_coupled_loop_reproduces_reference_trajectory():
    """forcing (moving target velocity) + stress integral + collision + sub-stepped integration + re-marking, coupled to
    the flow: 16 fluid steps (1600 sub-steps), the pair collides on the way"""
    g = np.load(os.path.join(gu.GOLDEN, "colliding_spheres_coupled.npz"))["history"]
    case, flow = util.colliding_spheres_setup()
    cloud = util.OracleAsCloud(case, stale_tables=True)
    hist = np.array(mini_flow.run_moving_bodies(cloud, flow, len(g), state=util.kinematics))
    assert np.array_equal(hist, g)
    vx = hist[:, 0, 3]
    assert vx[0] > 0.5 and vx[-1] < -0.1, "the pair must have collided and separated"
