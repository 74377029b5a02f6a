"""Fifth-generation kernels (one CTA per instance, cp.async box staging; ibm_kernels_v5.cuh) on the GPU.

NOT YET RUN ON HARDWARE when this file was written (the round's GPU budget was spent): the kernels' logic is checked on
the CPU by tests/test_emu_v5.py, which executes the same source under an emulation.  The file sorts last so that a
surprise here cannot hide the results of the established parity tests under `pytest -x`; the variant is not a default.
Same bars as tests/test_gpu_forcing.py / test_gpu_stress.py: forcing 1e-12 relative, force/torque 1e-10 relative."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import golden_util as gu
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


def _aligned():
    case = _moving(cases.assembly(n=24, n_particles=2, seed=11))
    case.particles.position[0] = [12.5, 11.5, 13.5]
    case.particles.position[1] = [4.0, 19.0, 6.0]
    return case


CASES = {
    "asm48": lambda: _moving(util.small_assembly(48, 40)),            # fits the staged box: the v5 kernels run
    "wc3d": lambda: _moving(ref_cases.wall_cyclic_3d()),              # boxes clipped by walls / periodic sides
    "aligned": _aligned,                                             # degenerate bins, points on cell faces, 164 shell cells (2 chunks)
    "cyl8": lambda: ref_cases.wall_cyclic_2d(cases.CYLINDER),         # one plane of cells, nk = 1
    "sph48": lambda: _moving(cases.sphere_case(n=48)),                # d = 9.6 cells, delta = 0.3: fits
    "col": lambda: _moving(cases.colliding_particles()),              # d = 30 cells: does NOT fit -> host routes to v4/v2
    "ellipse2": util.ellipse_case,                                   # ellipses never take the table path
}


@pytest.mark.parametrize("variant", [5, 6])   # 6 = the same kernel keeping the stencils between the field phases
@pytest.mark.parametrize("name", list(CASES))
def test_forcing_v5_parity(name, variant):
    case = CASES[name]()
    f = util.caller_fields(case)
    oc, gc = OracleCloud(case), ParticleCloud(case)
    gc.set_kernel_variant(forcing=variant, stress=5)
    Fo, n_upd = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    Fg = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    assert n_upd > 0 and np.max(np.abs(Fo)) > 0
    err = util.rel_err(Fg, Fo)
    assert err < 1e-12, err
    assert np.array_equal(Fg, gc.forcing(f["S"], f["U"], f["Uold"], case.dt))
    assert gc.timings()["n_updates"] == 2 * n_upd
    gc.close()


@pytest.mark.parametrize("name", list(CASES))
def test_stress_v5_parity(name):
    case = CASES[name]()
    case.cloud.pRef, case.cloud.rhoRef = 0.37, 1.3
    f = util.caller_fields(case)
    oc, gc = OracleCloud(case), ParticleCloud(case)
    gc.set_kernel_variant(forcing=5, stress=5)
    FTo = oc.integrate_stress(f["p"], f["tau"])
    FTg = gc.integrateSurfaceStresses(f["p"], f["tau"])
    for sl in (slice(0, 3), slice(3, 6)):
        scale = np.max(np.linalg.norm(FTo[:, sl], axis=1))
        assert np.max(np.abs(FTg[:, sl] - FTo[:, sl])) / scale < 1e-10
    gc.close()


@pytest.mark.parametrize("variant", [5, 6])
def test_v5_against_reference_vectors(variant):
    case, g = gu.load_static("asm24_static")
    gc = ParticleCloud(case)
    gc.set_kernel_variant(forcing=variant, stress=5)
    f = util.caller_fields(case)
    F = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    assert util.rel_err(F, gu.dense_forcing(g, case.mesh.n_cells)) < 1e-12
    FT = gc.integrateSurfaceStresses(f["p"], f["tau"])
    ref = g["stress_FT"]
    for sl in (slice(0, 3), slice(3, 6)):
        assert np.max(np.abs(FT[:, sl] - ref[:, sl])) / np.max(np.linalg.norm(ref[:, sl], axis=1)) < 1e-10
    gc.close()


def test_cached_stencils_give_the_same_field_bit_for_bit():
    """variant 6 reuses the stencils variant 5 recomputes: identical gathers, identical sums (non-overlapping shells)"""
    case = CASES["asm48"]() if "asm48" in CASES else cases.assembly(n=48, n_particles=40, seed=2, moving=True)
    f = util.caller_fields(case)
    gc = ParticleCloud(case)
    out = {}
    for variant in (5, 6):
        gc.set_kernel_variant(forcing=variant, stress=5)
        out[variant] = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
    assert np.array_equal(out[5], out[6])
    gc.close()


def test_v5_full_step_matches_default_kernels():
    """forcing + solve with the v5 kernels against the default generation on a moving assembly (F/T feed loop 5)"""
    case = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    f = util.caller_fields(case)
    out = []
    for variant in (None, 5):
        gc = ParticleCloud(case)
        if variant is not None:
            gc.set_kernel_variant(forcing=variant, stress=variant)
        for _ in range(2):
            F = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
            gc.solve(f["p"], f["tau"], case.dt)
        out.append((F, gc.state_arrays()))
        gc.close()
    (F0, s0), (F5, s5) = out
    assert util.rel_err(F5, F0) < 1e-11
    for k in ("index", "copy", "active", "onMesh", "nShell"):
        assert np.array_equal(s0[k], s5[k]), k
    for k in ("center", "v", "omega"):
        assert np.allclose(s0[k], s5[k], rtol=1e-9, atol=1e-12), k


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
