"""Fifth-generation kernels (ibmfoam_b200/csrc/ibm_kernels_v5.cuh) under the CPU emulation of tools/cuda_emu.

The SAME source that nvcc compiles for sm_100a is compiled with g++ against a stand-in <cuda_runtime.h>; blocks run one
after another, the threads of a block are fibers, cp.async copies are only performed when the issuing thread waits for
them.  This checks the kernels' logic (boxes, staging, column hash, stencils, barrier placement, reductions) against the
oracle on a machine without a GPU.  It is NOT a parity claim for the CUDA path: that is tests/test_gpu_*.py on a B200.
Tolerances as for the GPU: forcing 1e-12 relative, force/torque 1e-10 relative."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ibmfoam_b200 import cases
from oracle.pyoracle import OracleCloud
import ibm_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_v5.cpp")
OUT = os.path.join(ROOT, "tests", "emu", "_build", "libemu_v5.so")
DEPS = [SRC] + [os.path.join(ROOT, "ibmfoam_b200", "csrc", f) for f in ("ibm_kernels_v5.cuh", "ibm_device.cuh", "ibm_tables.cuh")] + \
       [os.path.join(ROOT, "tools", "cuda_emu", f) for f in ("emu.hpp", "cuda_runtime.h", "cuda_pipeline.h")]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(ROOT, "tools", "cuda_emu"),
                               "-I" + os.path.join(ROOT, "ibmfoam_b200", "csrc"), "-shared", "-fPIC", SRC, "-o", OUT])
    return C.CDLL(OUT)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _lattice_args(mesh):
    n = np.array([mesh.nx, mesh.ny, mesh.nz], np.int32)
    return n, np.array(mesh.origin, float), np.array(mesh.delta, float)


def _instances(case, oc, seed):
    st = oc.state_arrays()
    n = oc.n_instances
    c2l = np.argsort(case.mesh.lattice_to_cell()).astype(np.int32)  # caller cell id -> lattice index
    cells, jk, off = [], [], [0]
    for i in range(n):
        sh = oc.shell(i)
        cells.append(c2l[sh["cells"]])
        jk.append((sh["jk"][:, 0] | (sh["jk"][:, 1] << 16)).astype(np.int32))
        off.append(off[-1] + len(sh["cells"]))
    soa = lambda a: np.ascontiguousarray(a.T, dtype=np.float64)
    return dict(n=n, order=np.random.default_rng(seed).permutation(n).astype(np.int32), centre=soa(st["center"]), v=soa(st["v"]),
                om=soa(st["omega"]), theta=st["theta"], onLocal=st["onMesh"].astype(np.int32), nShell=st["nShell"].astype(np.int32),
                shellOff=np.array(off[:-1], np.int32), shellCell=np.concatenate(cells).astype(np.int32) if n else np.zeros(0, np.int32),
                shellJK=np.concatenate(jk).astype(np.int32) if n else np.zeros(0, np.int32))


def _rotation(theta):
    """particleShape::setRotationMatrix (particleShape.C:186-223): Rx Ry Rz, a factor skipped when |angle| <= SMALL."""
    R = np.eye(3)
    t1, t2, t3 = theta
    if abs(t1) > 1e-15:
        R = R @ np.array([[1, 0, 0], [0, np.cos(t1), -np.sin(t1)], [0, np.sin(t1), np.cos(t1)]])
    if abs(t2) > 1e-15:
        R = R @ np.array([[np.cos(t2), 0, np.sin(t2)], [0, 1, 0], [-np.sin(t2), 0, np.cos(t2)]])
    if abs(t3) > 1e-15:
        R = R @ np.array([[np.cos(t3), -np.sin(t3), 0], [np.sin(t3), np.cos(t3), 0], [0, 0, 1]])
    return R


def _forcing(emu, case, seed=0):
    oc = OracleCloud(case)
    f_lat = cases.synthetic_fields(case.mesh)
    f = util.caller_fields(case)
    Fo, n_upd = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    ins = _instances(case, oc, seed)
    n, origin, delta = _lattice_args(case.mesh)
    p = case.particles
    soa = lambda a: np.ascontiguousarray(a.T, dtype=np.float64)
    U, Uold, S = soa(f_lat["U"]), soa(f_lat["Uold"]), soa(f_lat["S"])
    F = np.zeros_like(U)
    rc = emu.emu_forcing_v5(_ip(n), _dp(origin), _dp(delta), int(case.mesh.kz0), int(p.kind[0]), int(p.nTheta[0]), C.c_double(p.d[0]),
                            C.c_double(p.delta[0]), ins["n"], _ip(ins["order"]), _dp(ins["centre"]), _dp(ins["v"]), _dp(ins["om"]),
                            _ip(ins["onLocal"]), _ip(ins["nShell"]), _ip(ins["shellOff"]), _ip(ins["shellCell"]), _ip(ins["shellJK"]),
                            _dp(U), _dp(Uold), _dp(S), C.c_double(case.dt), _dp(F))
    assert rc == 0, "kernel reported that the class does not fit its shared-memory budget"
    Fo_lat = Fo[case.mesh.lattice_to_cell()]
    return F.T, Fo_lat, n_upd


def _moving(case, seed=5):
    rng = np.random.default_rng(seed)
    n = case.particles.n
    case.particles.v[:] = rng.normal(0, 0.3, (n, 3))
    case.particles.omega[:] = rng.normal(0, 0.2, (n, 3))
    if case.mesh.nz == 1:
        case.particles.v[:, 2] = 0
        case.particles.omega[:, :2] = 0
    return case


def _wall_box(n=24):
    """spheres next to walls and periodic sides: boxes clipped by the lattice, PARTIAL instances, missing neighbours"""
    import ref_cases
    mesh = cases.MeshSpec(n, n, n, origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0), patches=ref_cases.WALL_X_CYCLIC_YZ)
    cloud = cases.CloudSpec(coupled=True, UInf=(0.2, 0, 0))
    pos = [[4.3, 11.2, 12.7], [19.6, 3.1, 20.4], [12.4, 21.8, 1.9], [11.7, 12.2, 11.6]]
    parts = cases.ParticleTable.uniform(4, cases.SPHERE, pos, d=8.0, nTheta=50, delta=1.0, rho=1.5)
    return cases.Case("wallbox", mesh, cloud, parts, dt=0.25)


FORCING_CASES = {
    "asm32": lambda: _moving(util.small_assembly(32, 10)),
    "wallbox": lambda: _moving(_wall_box()),
    "aligned": lambda: _moving(cases.assembly(n=24, n_particles=2, seed=11)),   # centres replaced below (App. B #3 clamp)
    "cyl8": None,
    "thick_shell": None,
}


def _aligned():
    case = _moving(cases.assembly(n=24, n_particles=2, seed=11))
    case.particles.position[0] = [12.5, 11.5, 13.5]   # on a cell-centre column: a shell cell at phi = pi, degenerate bins
    case.particles.position[1] = [4.0, 19.0, 6.0]     # on cell corners: points exactly on cell faces
    return case


def _thick_shell():
    """delta = 2 cells: more shell cells than one chunk (F5_CELLS = 160) -> the chunk loop and the box limits"""
    case = _moving(cases.assembly(n=28, n_particles=3, seed=2, d=6.0))
    case.particles.delta[:] = 2.0
    return case


def _cyl8():
    import ref_cases
    return ref_cases.wall_cyclic_2d(cases.CYLINDER)   # 2-D lattice (one plane of cells), cylinders d = 8 cells, nk = 1


FORCING_CASES["cyl8"] = _cyl8
FORCING_CASES["aligned"] = _aligned
FORCING_CASES["thick_shell"] = _thick_shell


@pytest.mark.parametrize("cache", [False, True])
@pytest.mark.parametrize("name", sorted(FORCING_CASES))
def test_forcing_v5_emulated(emu, name, cache, monkeypatch):
    """cache = the instantiation that keeps the stencils between the field phases (kernel variant 6)"""
    if cache:
        monkeypatch.setenv("EMU_V5_CACHE", "1")
    else:
        monkeypatch.delenv("EMU_V5_CACHE", raising=False)
    case = FORCING_CASES[name]()
    F, Fo, n_upd = _forcing(emu, case)
    assert n_upd > 0 and np.max(np.abs(Fo)) > 0
    assert np.array_equal(Fo.any(axis=1), F.any(axis=1)), "support of the forcing field"
    err = util.rel_err(F, Fo)
    assert err < 1e-12, err


def test_forcing_with_cached_stencils_equals_recomputed_stencils_bit_for_bit(emu, monkeypatch):
    case = FORCING_CASES["asm32"]()
    monkeypatch.delenv("EMU_V5_CACHE", raising=False)
    F1, _, _ = _forcing(emu, case, seed=1)
    monkeypatch.setenv("EMU_V5_CACHE", "1")
    F2, _, _ = _forcing(emu, case, seed=1)
    assert np.array_equal(F1, F2)   # same stencils, same gathers, same sums: the cache only saves their recomputation


def test_forcing_v5_is_order_independent(emu):
    case = FORCING_CASES["asm32"]()
    F1, _, _ = _forcing(emu, case, seed=1)
    F2, _, _ = _forcing(emu, case, seed=2)
    assert np.array_equal(F1, F2)   # non-overlapping shells: one atomicAdd per cell and component


def _stress(emu, case, seed=0):
    case.cloud.pRef, case.cloud.rhoRef = 0.37, 1.3
    oc = OracleCloud(case)
    f_lat = cases.synthetic_fields(case.mesh)
    f = util.caller_fields(case)
    FTo, Tc = oc.integrate_stress(f["p"], f["tau"], corrected=True)
    ins = _instances(case, oc, seed)
    n, origin, delta = _lattice_args(case.mesh)
    p = case.particles
    P = np.ascontiguousarray(f_lat["p"], dtype=np.float64)
    Tau = np.ascontiguousarray(f_lat["tau"].T, dtype=np.float64)
    rot = np.ascontiguousarray(np.stack([_rotation(t) for t in ins["theta"]]).reshape(-1, 9))
    Fout, Tout, Tcorr = np.full((3, ins["n"]), 7.0), np.full((3, ins["n"]), 7.0), np.zeros((ins["n"], 3))
    rc = emu.emu_stress_v5(_ip(n), _dp(origin), _dp(delta), int(case.mesh.kz0), int(p.kind[0]), int(p.nTheta[0]), C.c_double(p.d[0]),
                           C.c_double(p.delta[0]), ins["n"], _ip(ins["order"]), _dp(ins["centre"]), _dp(rot), _ip(ins["onLocal"]), _dp(P),
                           _dp(Tau), C.c_double(case.cloud.pRef / case.cloud.rhoRef), C.c_double(case.cloud.rhoRef), _dp(Fout), _dp(Tout),
                           _dp(Tcorr))
    assert rc == 0
    return np.concatenate([Fout.T, Tout.T], axis=1), FTo, Tcorr, Tc


@pytest.mark.parametrize("name", ["asm32", "wallbox", "aligned", "rotated", "cyl8"])
def test_stress_v5_emulated(emu, name):
    if name == "rotated":
        case = _moving(util.small_assembly(32, 6))
        case.particles.theta[:] = np.random.default_rng(3).normal(0, 0.7, case.particles.theta.shape)
    else:
        case = FORCING_CASES[name]()
    FT, FTo, Tcorr, Tc = _stress(emu, case)
    for sl in (slice(0, 3), slice(3, 6)):
        scale = np.max(np.linalg.norm(FTo[:, sl], axis=1))
        assert np.max(np.abs(FT[:, sl] - FTo[:, sl])) / scale < 1e-10
    assert np.max(np.abs(Tcorr - Tc)) / np.max(np.linalg.norm(Tc, axis=1)) < 1e-10


def test_stress_v5_writes_zero_partials_for_instances_without_local_points(emu):
    """multi-GPU: an instance of another slab must contribute zeros, not whatever its F/T slots held"""
    case = FORCING_CASES["asm32"]()
    case.cloud.pRef, case.cloud.rhoRef = 0.0, 1.0
    oc = OracleCloud(case)
    ins = _instances(case, oc, 0)
    ins["onLocal"][::2] = 0
    f_lat = cases.synthetic_fields(case.mesh)
    n, origin, delta = _lattice_args(case.mesh)
    p = case.particles
    P = np.ascontiguousarray(f_lat["p"], dtype=np.float64)
    Tau = np.ascontiguousarray(f_lat["tau"].T, dtype=np.float64)
    rot = np.ascontiguousarray(np.tile(np.eye(3).reshape(1, 9), (ins["n"], 1)))
    Fout, Tout = np.full((3, ins["n"]), 7.0), np.full((3, ins["n"]), 7.0)
    emu.emu_stress_v5(_ip(n), _dp(origin), _dp(delta), 0, int(p.kind[0]), int(p.nTheta[0]), C.c_double(p.d[0]), C.c_double(p.delta[0]),
                      ins["n"], _ip(ins["order"]), _dp(ins["centre"]), _dp(rot), _ip(ins["onLocal"]), _dp(P), _dp(Tau), C.c_double(0.0),
                      C.c_double(1.0), _dp(Fout), _dp(Tout), None)
    assert not Fout[:, ::2].any() and not Tout[:, ::2].any()
    assert Fout[:, 1::2].any()


# ---- solid volume fraction (extension; ibm_kernels_alpha.cuh) ----
def _alpha(emu, case, n_sub=8, rot=None):
    from oracle import alpha_oracle
    p = case.particles
    n, origin, delta = _lattice_args(case.mesh)
    centres = np.ascontiguousarray(p.position.T, dtype=np.float64)
    rot = np.ascontiguousarray(rot if rot is not None else np.tile(np.eye(3).reshape(1, 9), (p.n, 1)))
    out = np.zeros(case.mesh.n_cells)
    emu.emu_solid_fraction(_ip(n), _dp(origin), _dp(delta), int(case.mesh.kz0), int(p.kind[0]), C.c_double(p.d[0]), C.c_double(p.a[0]),
                           C.c_double(p.b[0]), p.n, _dp(centres), _dp(rot), int(n_sub), _dp(out))
    ref = alpha_oracle.solid_fraction(case.mesh, int(p.kind[0]), p.position, d=float(p.d[0]), a=float(p.a[0]), b=float(p.b[0]),
                                      rot=rot.reshape(-1, 3, 3), n_sub=n_sub)
    return out, ref


@pytest.mark.parametrize("name", ["asm32", "wallbox", "aligned", "cyl8", "ellipse"])
def test_solid_fraction_emulated_is_exact_and_integrates_to_the_body_volume(emu, name):
    rot = None
    if name == "ellipse":
        case = util.ellipse_case()
        rot = np.stack([_rotation(t) for t in case.particles.theta]).reshape(-1, 9)
    else:
        case = FORCING_CASES[name]()
    a, ref = _alpha(emu, case, rot=rot)
    assert np.array_equal(a, ref)                       # integer sample counts: exact
    assert a.min() >= 0.0 and a.max() == 1.0
    p, m = case.particles, case.mesh
    cell_v = m.delta[0] * m.delta[1] * m.delta[2]
    if name in ("asm32", "cyl8", "ellipse"):            # bodies entirely inside the lattice, not overlapping
        L = m.nz * m.delta[2]
        vol = {0: np.pi * p.d[0] ** 3 / 6.0, 1: np.pi * (p.d[0] / 2) ** 2 * L, 2: np.pi * p.a[0] * p.b[0] * L}[int(p.kind[0])] * p.n
        if name == "asm32":                             # periodic box: spheres cut by the sides lose the part outside
            inside = np.all((p.position - p.d[0] / 2 >= 0) & (p.position + p.d[0] / 2 <= m.nx), axis=1)
            a2, _ = _alpha(emu, type(case)(case.name, m, case.cloud, _subset(p, inside), case.dt))
            vol, a = np.pi * p.d[0] ** 3 / 6.0 * inside.sum(), a2
        assert abs(a.sum() * cell_v - vol) < 2e-3 * vol


def _subset(p, mask):
    return cases.ParticleTable(**{k: getattr(p, k)[mask] for k in ("kind", "position", "d", "a", "b", "nTheta", "nk", "delta", "theta", "v",
                                                                    "omega", "rho")})


def test_solid_fraction_is_consistent_with_the_reference_indicator(emu):
    """I.ibm (cell centre inside) is 1 exactly where more than ~half of the cell is solid, for convex bodies; alpha = 1 deep
    inside and 0 far outside"""
    case = FORCING_CASES["asm32"]()
    a, _ = _alpha(emu, case)
    I = OracleCloud(case).indicator()[np.argsort(case.mesh.lattice_to_cell())]  # -> lattice order (single block: identity)
    assert np.all(a[I == 1] > 0.35) and np.all(a[I == 0] < 0.65)
