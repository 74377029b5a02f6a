"""EXTENSION without a reference counterpart: the fractional solid volume fraction alpha_s
(ibm_get_solid_fraction, ibm_kernels_alpha.cuh) against the numpy restatement of its definition (oracle/alpha_oracle.py)."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
import ibm_testutil as util
import ref_cases

pytestmark = pytest.mark.gpu


def _moving(case, seed=5):
    rng = np.random.default_rng(seed)
    n = case.particles.n
    case.particles.v[:] = rng.normal(0, 0.3, (n, 3))
    case.particles.omega[:] = rng.normal(0, 0.2, (n, 3))
    if case.mesh.nz == 1:
        case.particles.v[:, 2] = 0
        case.particles.omega[:, :2] = 0
    return case


CASES = {
    "asm48": lambda: _moving(util.small_assembly(48, 40)),
    "wc3d": lambda: _moving(ref_cases.wall_cyclic_3d()),              # boxes clipped by walls / periodic sides
    "cyl8": lambda: ref_cases.wall_cyclic_2d(cases.CYLINDER),         # one plane of cells, nk = 1
    "ellipse2": util.ellipse_case,
}


@pytest.mark.parametrize("name", ["asm48", "wc3d", "cyl8", "ellipse2"])
def test_solid_fraction_extension_matches_its_definition(name):
    """EXTENSION without a reference counterpart: alpha_s against the numpy restatement of its
    definition - exact for spheres / cylinders (integer sample counts); ellipses within a few samples (device sin/cos in
    the rotation matrix)."""
    from oracle import alpha_oracle
    case = CASES[name]()
    gc = ParticleCloud(case)
    a = gc.solid_fraction(8)
    p = case.particles
    rot = None
    if p.kind[0] == cases.ELLIPSE:
        th = p.theta[:, 2]
        rot = np.stack([np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1]]) for t in th])
    ref_lat = alpha_oracle.solid_fraction(case.mesh, int(p.kind[0]), p.position, d=float(p.d[0]), a=float(p.a[0]), b=float(p.b[0]), rot=rot)
    ref_caller = np.empty_like(ref_lat)
    ref_caller[case.mesh.lattice_to_cell()] = ref_lat          # lattice order -> the caller's cell order
    if p.kind[0] == cases.ELLIPSE:
        assert np.max(np.abs(a - ref_caller)) <= 3.0 / 64.0
    else:
        assert np.array_equal(a, ref_caller)
    gc.close()
