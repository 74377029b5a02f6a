"""ibm_lattice_from_arrays: the boundary takes what the reference has - polyMesh's raw cell centres and patch face centres
(particleShape.C:39-66) - and recognises the uniform lattice in them, in any cell numbering; graded meshes are refused.
Host-only entry point: runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from ibmfoam_b200 import capi, cases


def _call(centres, geo_d, patch_cf):
    L = capi.load()
    centres = np.ascontiguousarray(centres, dtype=np.float64)
    n_cells = centres.shape[0]
    np_ = len(patch_cf)
    cfs = [np.ascontiguousarray(a, dtype=np.float64) for a in patch_cf]
    ptrs = (C.c_void_p * max(1, np_))(*[a.ctypes.data for a in cfs])
    nf = np.array([a.shape[0] for a in cfs] or [0], dtype=np.int32)
    n = np.zeros(3, dtype=np.int32)
    origin, delta = np.zeros(3), np.zeros(3)
    side = np.full(max(1, np_), -9, dtype=np.int32)
    l2c = np.full(n_cells, -1, dtype=np.int32)
    ident = C.c_int(-1)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rc = L.ibm_lattice_from_arrays(vp(centres), n_cells, geo_d, np_, C.cast(ptrs, C.c_void_p), vp(nf), vp(n), vp(origin), vp(delta), vp(side),
                                   vp(l2c), C.cast(C.byref(ident), C.c_void_p))
    return rc, n, origin, delta, side[:np_], l2c, ident.value


def _side_faces(mesh, side):
    """face centres of one lattice side"""
    o, d, n = np.array(mesh.origin), np.array(mesh.delta), np.array([mesh.nx, mesh.ny, mesh.nz])
    ax = side // 2
    others = [a for a in range(3) if a != ax]
    g = np.meshgrid(*[o[a] + (np.arange(n[a]) + 0.5) * d[a] for a in others], indexing="ij")
    cf = np.zeros((g[0].size, 3))
    cf[:, others[0]], cf[:, others[1]] = g[0].ravel(), g[1].ravel()
    cf[:, ax] = o[ax] + (n[ax] * d[ax] if side & 1 else 0.0)
    return cf


@pytest.mark.parametrize("name", ["assembly", "cylinder_4_blocks", "shuffled"])
def test_uniform_lattice_recognised_in_any_numbering(name):
    case = {"assembly": lambda: cases.assembly(n=24, n_particles=4, seed=1),
            "cylinder_4_blocks": lambda: cases.cylinder_re30(),
            "shuffled": lambda: cases.assembly(n=12, n_particles=2, seed=1)}[name]()
    m = case.mesh
    lat_centres = m.cell_centres_lattice()
    l2c_true = m.lattice_to_cell().astype(np.int64)
    if name == "shuffled":
        l2c_true = np.random.default_rng(3).permutation(m.n_cells)
    centres = np.empty_like(lat_centres)
    centres[l2c_true] = lat_centres           # the caller's numbering
    sides = [s for (s, _, _) in m.patches]
    rc, n, origin, delta, side, l2c, ident = _call(centres, 2 if m.nz == 1 else 3, [_side_faces(m, s) for s in sides])
    assert rc == 0
    assert tuple(n) == (m.nx, m.ny, m.nz)
    assert np.allclose(origin[:2], m.origin[:2], rtol=0, atol=1e-12 * max(1.0, abs(m.origin[0])))
    assert np.allclose(delta[:2], m.delta[:2], rtol=1e-12)
    if m.nz > 1:
        assert np.isclose(origin[2], m.origin[2], atol=1e-12) and np.isclose(delta[2], m.delta[2], rtol=1e-12)
    assert np.array_equal(l2c, l2c_true)
    assert ident == int(np.array_equal(l2c_true, np.arange(m.n_cells)))
    assert list(side) == sides


def test_graded_mesh_is_refused():
    x = np.cumsum(1.0 * 1.1 ** np.arange(12))      # simpleGrading != 1
    Z, Y, X = np.meshgrid(np.arange(6) + 0.5, np.arange(6) + 0.5, x, indexing="ij")
    centres = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    rc, *_ = _call(centres, 3, [])
    assert rc == capi.load().ibm_lattice_from_arrays.restype(-4) or rc == -4   # IBM_E_UNSUPPORTED


def test_patch_off_the_sides_is_refused():
    m = cases.assembly(n=8, n_particles=1, seed=1).mesh
    centres = m.cell_centres_lattice()
    bad = _side_faces(m, 0)
    bad[:, 0] += 0.3 * m.delta[0]
    rc, *_ = _call(centres, 3, [bad])
    assert rc == -4
