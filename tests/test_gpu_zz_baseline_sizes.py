"""Parity at the sizes BASELINE.json states ("what's weak" 1): validation/sphere at 150^3 with nTheta = 200
(N = 60 000 shape points, ~2 600 shell cells) and the ~10^3-sphere assembly at 128^3 (1 565 spheres d = 8): marking
bit-exact, forcing 1e-12 (max|dF|/max|F|), force / torque 1e-10, against the oracle's fast (ijk) search."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import ibm_testutil as util

pytestmark = pytest.mark.gpu


def _parity(case, n_shell_checks):
    f = util.caller_fields(case)
    oc, gc = OracleCloud(case), ParticleCloud(case)
    so, sg = oc.state_arrays(), gc.state_arrays()
    assert gc.nParticles() == oc.n_instances
    for k in ("index", "copy", "active", "onMesh", "nShell", "N"):
        assert np.array_equal(so[k], sg[k]), k
    for inst in np.linspace(0, oc.n_instances - 1, n_shell_checks).astype(int):
        a, b = oc.shell(int(inst)), gc.shell(int(inst))
        for k in ("cells", "jk", "Is", "Os"):
            assert np.array_equal(a[k], b[k]), (inst, k)
    assert np.array_equal(oc.indicator(), gc.indicator())
    Fo, n_upd = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    Fg = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    assert n_upd == int(so["nShell"].sum())
    assert np.array_equal(Fo.any(axis=1), Fg.any(axis=1))  # same support
    assert util.rel_err(Fg, Fo) < 1e-12, util.rel_err(Fg, Fo)
    FTo = oc.integrate_stress(f["p"], f["tau"])
    FTg = gc.integrateSurfaceStresses(f["p"], f["tau"])
    assert np.max(np.abs(FTg - FTo)) < 1e-10 * np.max(np.abs(FTo))
    gc.close()
    return n_upd


def test_validation_sphere_150():
    case = cases.sphere_case(n=150)
    assert case.mesh.n_cells == 150 ** 3 and int(case.particles.nTheta[0]) == 200
    n_upd = _parity(case, 1)
    assert 2000 < n_upd < 3500  # ~2 580 shell cells


def test_particle_assembly_128():
    case = cases.assembly(n=128)  # 1 565 fixed spheres, solid fraction 0.2
    assert case.particles.n == 1565
    n_upd = _parity(case, 40)
    assert 150 * 1565 * 0.9 < n_upd < 160 * 1565 * 1.1
