"""Loop 1 parity (GPU, through the C ABI): cell marking, polar bins, neighbour tables, point->cell membership,
gather face lists and the indicator field must be BIT-EXACT against the oracle; IDW weights to 1e-14."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import ibm_testutil as util

pytestmark = pytest.mark.gpu


def _cases():
    return {
        "cyl30": cases.cylinder_re30(),
        "sph48": cases.sphere_case(n=48),
        "asm32": util.small_assembly(32, 12),
        "col": cases.colliding_particles(),
        "ellipse2": util.ellipse_case(),
        "two_spheres": cases.two_spheres(n=40),
    }


@pytest.mark.parametrize("name", list(_cases()))
def test_marking_bit_exact(name):
    case = _cases()[name]
    oc = OracleCloud(case)
    gc = ParticleCloud(case)
    so, sg = oc.state_arrays(), gc.state_arrays()
    assert gc.nParticles() == oc.n_instances
    for k in ("index", "copy", "active", "onMesh", "nShell", "N"):
        assert np.array_equal(so[k], sg[k]), k
    n_check = min(oc.n_instances, 6)
    for inst in range(n_check):
        a, b = oc.shell(inst), gc.shell(inst)
        for k in ("cells", "jk", "Is", "Os"):
            assert np.array_equal(a[k], b[k]), (inst, k)
        assert np.allclose(a["w"], b["w"], rtol=1e-14, atol=0)
        assert np.allclose(a["W"], b["W"], rtol=1e-14, atol=0)
        pa, pb = oc.points(inst), gc.points(inst)
        assert np.array_equal(pa["pts"], pb["pts"]), "shape points must be bit-identical"
        for k in ("cell", "nFaces", "otherCell"):
            assert np.array_equal(pa[k], pb[k]), (inst, k)
        assert np.allclose(pa["w"], pb["w"], rtol=1e-14, atol=0)
        assert np.allclose(pa["W"], pb["W"], rtol=1e-14, atol=0)
    assert np.array_equal(oc.indicator(), gc.indicator())
    gc.close()


def test_marking_all_instances_assembly():
    case = cases.assembly(n=64, n_particles=150, seed=3)
    oc, gc = OracleCloud(case), ParticleCloud(case)
    so, sg = oc.state_arrays(), gc.state_arrays()
    assert np.array_equal(so["nShell"], sg["nShell"])
    assert np.array_equal(so["onMesh"], sg["onMesh"])
    for inst in range(oc.n_instances):
        a, b = oc.shell(inst), gc.shell(inst, so["nShell"][inst])
        assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["jk"], b["jk"])
    assert np.array_equal(oc.indicator(), gc.indicator())
    gc.close()


def test_shape_errors_like_reference():
    # sphere.C:79-84 / cylinder.C:64-70: FatalError in the reference, IbmError here
    from ibmfoam_b200.capi import IbmError
    case = cases.cylinder_re30()
    case.particles.kind[:] = cases.SPHERE
    with pytest.raises(IbmError):
        ParticleCloud(case)
