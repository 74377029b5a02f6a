"""Both marking kernels (ibm_set_option "mark_variant": 1 = k1_mark2, the default since round 2 - radial test over the bounding
box, polar bins over the compacted list; 0 = k1_mark, the single-pass first version): bit-exact against the oracle."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import ibm_testutil as util

pytestmark = pytest.mark.gpu


VARIANT = [1]


@pytest.fixture(autouse=True, params=[1, 0])
def _variant(request):
    VARIANT[0] = request.param
    yield


def _cloud(case, **kw):
    gc = ParticleCloud(case, **kw)
    gc.set_option("mark_variant", VARIANT[0])
    gc.update()                      # the constructor marked with the default kernel: mark again with the requested version
    return gc


def _big_sphere():
    """one sphere whose shell (~5000 cells) does not fit the shared list: the direct-write path"""
    return cases.assembly(n=64, n_particles=1, seed=2, d=40.0, positions=np.array([[31.3, 32.1, 30.7]]))


def _edge_spheres():
    """spheres cut by the lattice boundary (clipped bounding boxes, PARTIAL status) and one outside reach"""
    pos = np.array([[1.2, 20.0, 20.0], [38.9, 39.2, 1.1], [20.3, 20.1, 19.8], [0.4, 0.3, 39.8]])
    return cases.assembly(n=40, n_particles=4, seed=2, positions=pos)


CASES = {"cyl30": cases.cylinder_re30, "sph48": lambda: cases.sphere_case(n=48), "asm32": lambda: util.small_assembly(32, 12),
         "col": cases.colliding_particles, "ellipse2": util.ellipse_case, "big_sphere": _big_sphere, "edge": _edge_spheres}


@pytest.mark.parametrize("name", list(CASES))
def test_second_marking_kernel_bit_exact(name):
    case = CASES[name]()
    oc, gc = OracleCloud(case), _cloud(case)
    so, sg = oc.state_arrays(), gc.state_arrays()
    assert gc.nParticles() == oc.n_instances
    for k in ("index", "copy", "active", "onMesh", "nShell", "N"):
        assert np.array_equal(so[k], sg[k]), k
    for inst in range(min(oc.n_instances, 8)):
        a, b = oc.shell(inst), gc.shell(inst)
        for k in ("cells", "jk", "Is", "Os"):
            assert np.array_equal(a[k], b[k]), (inst, k)
        assert np.allclose(a["w"], b["w"], rtol=1e-14, atol=0)
    if name == "big_sphere":
        assert so["nShell"][0] > 1024
    gc.close()


def test_forcing_and_trajectory_with_the_second_marking_kernel():
    """re-marking after every fluid step of a moving assembly, then the forcing field on the re-marked shells"""
    case = cases.assembly(n=48, n_particles=40, seed=6, moving=True, cyclic_without_walls=True)
    case.particles.v *= 10.0
    oc, gc = OracleCloud(case), _cloud(case)
    for _ in range(3):
        oc.combine_forces(); gc.combineForces()
        oc.advance(case.dt); gc.advance(case.dt)
        so, sg = oc.state_arrays(), gc.state_arrays()
        for k in ("index", "copy", "active", "onMesh", "nShell", "center"):
            assert np.array_equal(so[k], sg[k]), k
    for inst in range(oc.n_instances):
        a, b = oc.shell(inst), gc.shell(inst)
        assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["jk"], b["jk"]), inst
    f = util.caller_fields(case)
    Fo, _ = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    Fg = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    assert util.rel_err(Fg, Fo) < 1e-12
    gc.close()


def test_slab_decomposed_marking_with_the_second_kernel():
    """two slabs of one lattice in one process: per-slab shell lists of the second kernel equal the first kernel's"""
    from ibmfoam_b200.cloud import slab_ranges
    case = cases.assembly(n=32, n_particles=10, seed=8, moving=True, cyclic_without_walls=True)
    for slab in slab_ranges(case.mesh.nz, 2):
        a = ParticleCloud(case, slab=slab)
        b = _cloud(case, slab=slab)
        sa, sb = a.state_arrays(), b.state_arrays()
        assert np.array_equal(sa["nShell"], sb["nShell"])
        for inst in range(a.nParticles()):
            if sa["nShell"][inst]:
                x, y = a.shell(inst), b.shell(inst)
                assert np.array_equal(x["cells"], y["cells"]) and np.array_equal(x["jk"], y["jk"])
        a.close(); b.close()
