"""constant/IBMProperties reader / writer / generator (ibmfoam_b200/ibm_properties.py)."""
import os

import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200 import ibm_properties as ip
import ibm_testutil as util
import ref_cases

REF_VALIDATION = "/root/reference/validation"


def _same_table(a, b):
    for k in ("kind", "position", "d", "a", "b", "nTheta", "nk", "delta", "theta", "v", "omega", "rho"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("builder", [ref_cases.wall_cyclic_3d, lambda: ref_cases.wall_cyclic_2d(cases.ELLIPSE), util.ellipse_case,
                                     lambda: cases.assembly(n=32, n_particles=20, moving=True)])
def test_write_then_read_gives_the_same_doubles(builder):
    case = builder()
    text = ip.write_ibm_properties(case.cloud, case.particles)
    cloud, table = ip.read_ibm_properties(text, g=case.cloud.g)
    _same_table(table, case.particles)
    for k in ("moving", "rotating", "coupled", "e", "rhoRef", "pRef"):
        assert getattr(cloud, k) == getattr(case.cloud, k), k
    assert tuple(cloud.UInf) == tuple(float(x) for x in case.cloud.UInf)


@pytest.mark.skipif(not os.path.isdir(REF_VALIDATION), reason="reference validation cases not present")
def test_reads_the_reference_colliding_particles_file_verbatim():
    with open(os.path.join(REF_VALIDATION, "collidingParticles", "constant", "IBMProperties")) as f:
        cloud, table = ip.read_ibm_properties(f.read(), g=(0.0, -9.81, 0.0))
    ref = cases.colliding_particles()
    assert table.n == 45 and cloud.moving and cloud.rotating and cloud.coupled and cloud.e == 1.0
    assert tuple(cloud.UInf) == (0.0, 0.22, 0.0) and cloud.rhoRef == 1.0
    _same_table(table, ref.particles)


@pytest.mark.skipif(not os.path.isdir(REF_VALIDATION), reason="reference validation cases not present")
def test_cylinder_re30_file_lacks_the_mandatory_coupled_key():
    """App. B #16: the shipped case aborts in the reference (dict_.lookup("coupled"), particleCloud.C:1034); same here."""
    with open(os.path.join(REF_VALIDATION, "cylinderRe30", "constant", "IBMProperties")) as f:
        text = f.read()
    with pytest.raises(ip.FoamDictError, match="coupled"):
        ip.read_ibm_properties(text)
    cloud, table = ip.read_ibm_properties(text.replace("nParticles", "coupled true;\nnParticles", 1))
    ref = cases.cylinder_re30()
    _same_table(table, ref.particles)
    assert cloud.rhoRef == 1.0   # the file's `rhoRef` key is ignored; the code reads rhoInf (default 1)


def test_missing_keys_fail_like_the_reference():
    case = ref_cases.wall_cyclic_3d()
    text = ip.write_ibm_properties(case.cloud, case.particles)
    with pytest.raises(ip.FoamDictError, match="rho"):
        ip.read_ibm_properties(text.replace("rho\t1.5;", "", 1))
    with pytest.raises(ip.FoamDictError, match="Unknown particle shape"):
        ip.read_ibm_properties(text.replace("sphere", "cube", 1))
    with pytest.raises(ip.FoamDictError, match="particle.3"):
        ip.read_ibm_properties(text.replace("nParticles\t4;", "nParticles\t5;").replace("particle.3", "particle.7"))


def test_generator_obeys_the_reference_placement_rule_and_is_reproducible():
    t1 = ip.generate_monodisperse((0, 0, 0), (48, 48, 48), 8.0, volume_fraction=0.2, nTheta=50, delta=1.0, rho=1.5, seed=7,
                                  velocity_sigma=0.1, omega_sigma=0.0125)
    t2 = ip.generate_monodisperse((0, 0, 0), (48, 48, 48), 8.0, volume_fraction=0.2, nTheta=50, delta=1.0, rho=1.5, seed=7,
                                  velocity_sigma=0.1, omega_sigma=0.0125)
    _same_table(t1, t2)
    assert t1.n == round(0.2 * 48 ** 3 / (np.pi * 8.0 ** 3 / 6.0))          # monodisperse.C:64-110
    x = t1.position
    dmin = np.inf
    for i in range(t1.n):                                                    # |xi - xj| > ri + rj under minimum image (:162-178)
        dv = x[i + 1:] - x[i]
        dv -= 48.0 * np.round(dv / 48.0)
        if len(dv):
            dmin = min(dmin, np.sqrt((dv ** 2).sum(axis=1)).min())
    assert dmin > 8.0
    cyl = ip.generate_monodisperse((-0.02, -0.02, -0.001), (0.02, 0.02, 0.001), 0.004, n_particles=30, shape="cylinder", nTheta=300,
                                   delta=0.00013, seed=3, periodic=False)
    assert cyl.n == 30 and np.all(cyl.position[:, 2] == 0.0)
    for i in range(30):
        dv = cyl.position[i + 1:, :2] - cyl.position[i, :2]
        if len(dv):
            assert np.sqrt((dv ** 2).sum(axis=1)).min() > 0.004
    assert np.all(np.abs(cyl.position[:, :2]) <= 0.02 - 0.002)


def test_generated_file_runs_through_the_checker():
    """generator -> IBMProperties text -> reader -> the CPU checker: the round trip feeds the path"""
    from oracle.pyoracle import OracleCloud
    table = ip.generate_monodisperse((0, 0, 0), (32, 32, 32), 8.0, n_particles=10, nTheta=50, delta=1.0, rho=1.5, seed=2)
    base = cases.assembly(n=32, n_particles=10, seed=2)
    cloud, table2 = ip.read_ibm_properties(ip.write_ibm_properties(base.cloud, table))
    case = cases.Case("gen", base.mesh, cloud, table2, dt=0.25)
    oc = OracleCloud(case)
    assert oc.n_instances == 10 and np.all(oc.state_arrays()["nShell"] > 0)
