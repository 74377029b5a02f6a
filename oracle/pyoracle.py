"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of the two CPU checkers:

  impl="port"       oracle/_build/libibm_oracle.so   the restatement (oracle/orc_*.hpp), always available
  impl="reference"  oracle/_ref/libibm_ref.so        the REFERENCE'S OWN SOURCES compiled against oracle/mockfoam
                                                     (oracle/ref/Makefile; needs /root/reference at build time, the
                                                     built .so travels with gpurun)

Both export the same orc_* C surface.  Import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only.  The product (ibmfoam_b200/) never imports it.  Pinning: see oracle/README.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libibm_oracle.so")
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libibm_ref.so")
REFERENCE_SRC = "/root/reference/src"


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def build_reference(force: bool = False):
    """Compile the reference's own sources (only where /root/reference exists).  Returns the path or None."""
    if os.path.isdir(REFERENCE_SRC):
        if force:
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "ref"), "-s", "clean"])
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "ref"), "-s"])
    return _REF_LIB_PATH if os.path.exists(_REF_LIB_PATH) else None


def have_reference() -> bool:
    return os.path.exists(_REF_LIB_PATH) or os.path.isdir(REFERENCE_SRC)


class ParticleInit(C.Structure):
    _fields_ = [("kind", C.c_int), ("position", C.c_double * 3), ("d", C.c_double), ("a", C.c_double),
                ("b", C.c_double), ("nTheta", C.c_int), ("nk", C.c_int), ("delta", C.c_double),
                ("theta", C.c_double * 3), ("v", C.c_double * 3), ("omega", C.c_double * 3), ("rho", C.c_double)]


class CloudCfg(C.Structure):
    _fields_ = [("moving", C.c_int), ("rotating", C.c_int), ("coupled", C.c_int), ("e", C.c_double),
                ("UInf", C.c_double * 3), ("rhoRef", C.c_double), ("pRef", C.c_double), ("g", C.c_double * 3),
                ("faithfulSearch", C.c_int), ("cyclicWithoutWalls", C.c_int)]


class InstanceState(C.Structure):
    _fields_ = [("index", C.c_int), ("copy", C.c_int), ("active", C.c_int), ("onMesh", C.c_int),
                ("center", C.c_double * 3), ("theta", C.c_double * 3), ("v", C.c_double * 3),
                ("omega", C.c_double * 3), ("rho", C.c_double), ("age", C.c_double), ("F", C.c_double * 3),
                ("T", C.c_double * 3), ("nShell", C.c_int), ("N", C.c_int)]


_libs = {}


def lib(impl: str = "port"):
    if impl not in _libs:
        if impl == "port":
            path = build()
        elif impl == "reference":
            path = build_reference()
            if path is None:
                raise RuntimeError("oracle/_ref/libibm_ref.so is missing and /root/reference is not here to build it")
        else:
            raise ValueError(impl)
        L = C.CDLL(path)
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_cloud_create.restype = C.c_void_p
        L.orc_cloud_forcing.restype = C.c_long
        L.orc_particle_mass.restype = C.c_double
        L.orc_particle_inertia.restype = C.c_double
        L.orc_cloud_max_co_num.restype = C.c_double
        if impl == "reference":
            L.orc_last_error.restype = C.c_char_p
        _libs[impl] = L
    return _libs[impl]


class ReferenceError_(RuntimeError):
    """FatalError (or a detected out-of-range subscript) raised inside the compiled reference."""


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class OracleMesh:
    def __init__(self, spec, impl="port"):
        L = self.L = lib(impl)
        self.impl = impl
        self.spec = spec
        blocks = np.asarray(spec.blocks, dtype=np.int32).reshape(-1, 6) if spec.blocks else np.zeros((0, 6), np.int32)
        pat = np.asarray(spec.patches, dtype=np.int32).reshape(-1, 3) if spec.patches else np.zeros((0, 3), np.int32)
        side, typ, nb = _c(pat[:, 0], np.int32), _c(pat[:, 1], np.int32), _c(pat[:, 2], np.int32)
        origin, delta = _c(spec.origin), _c(spec.delta)
        blocks = _c(blocks, np.int32)
        self.h = C.c_void_p(L.orc_mesh_create(spec.nx, spec.ny, spec.nz, _dp(origin), _dp(delta), blocks.shape[0],
                                              _ip(blocks), pat.shape[0], _ip(side), _ip(typ), _ip(nb), int(getattr(spec, "kz0", 0))))
        self.n_cells = L.orc_mesh_ncells(self.h)
        self.n_internal_faces = L.orc_mesh_ninternal_faces(self.h)

    # (the geometry taps below exist in the restatement only: both libraries build the mini-mesh with the same code)
    def cell_centres(self):
        out = np.empty((self.n_cells, 3))
        self.L.orc_mesh_cell_centres(self.h, _dp(out))
        return out

    def lattice_to_cell(self):
        out = np.empty(self.n_cells, dtype=np.int32)
        self.L.orc_mesh_lattice_to_cell(self.h, _ip(out))
        return out

    def faces(self):
        n = self.n_internal_faces
        own, nei = np.empty(n, np.int32), np.empty(n, np.int32)
        Cf, lam = np.empty((n, 3)), np.empty(n)
        self.L.orc_mesh_faces(self.h, _ip(own), _ip(nei), _dp(Cf), _dp(lam))
        return own, nei, Cf, lam

    def patch_faces(self, patchi):
        n = self.L.orc_mesh_patch_size(self.h, patchi)
        Cf, Sf, magSf = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
        self.L.orc_mesh_patch_faces(self.h, patchi, _dp(Cf), _dp(Sf), _dp(magSf))
        return Cf, Sf, magSf

    def find_cell(self, pt, faithful=False):
        p = _c(pt)
        return self.L.orc_mesh_find_cell(self.h, _dp(p), int(faithful))

    def __del__(self):
        try:
            self.L.orc_mesh_destroy(self.h)
        except Exception:
            pass


def particle_inits(tab):
    arr = (ParticleInit * tab.n)()
    for i in range(tab.n):
        a = arr[i]
        a.kind = int(tab.kind[i])
        a.position[:] = tab.position[i]
        a.d, a.a, a.b = float(tab.d[i]), float(tab.a[i]), float(tab.b[i])
        a.nTheta, a.nk = int(tab.nTheta[i]), int(tab.nk[i])
        a.delta = float(tab.delta[i])
        a.theta[:] = tab.theta[i]
        a.v[:] = tab.v[i]
        a.omega[:] = tab.omega[i]
        a.rho = float(tab.rho[i])
    return arr


class OracleCloud:
    """Cell-indexed fields are in CALLER (blockMesh) cell order, AoS, like OpenFOAM's."""

    def __init__(self, case, faithful=False, mesh=None, impl="port", stale_tables=False):
        """stale_tables=True: quirk B#8 as coded (what the compiled reference always does); default = DEFINED policy."""
        L = self.L = lib(impl)
        self.impl = impl
        self.case = case
        self.mesh = mesh or OracleMesh(case.mesh, impl=impl)
        assert self.mesh.impl == impl
        cs = case.cloud
        cfg = CloudCfg(int(cs.moving), int(cs.rotating), int(cs.coupled), cs.e, (C.c_double * 3)(*cs.UInf), cs.rhoRef,
                       cs.pRef, (C.c_double * 3)(*cs.g), int(bool(faithful)) | (2 if stale_tables else 0), int(cs.cyclic_without_walls))
        inits = particle_inits(case.particles)
        h = L.orc_cloud_create(self.mesh.h, C.byref(cfg), case.particles.n, inits)
        if not h:
            if impl == "reference":
                raise ReferenceError_(L.orc_last_error().decode())
            raise RuntimeError("oracle: invalid case (shape kind vs mesh dimensionality)")
        self.h = C.c_void_p(h)

    def _chk(self, rc):
        # the restatement's entry points return void (ctypes then yields an arbitrary int): only the reference reports
        if self.impl == "reference" and rc:
            raise ReferenceError_(self.L.orc_last_error().decode())

    @property
    def n_instances(self):
        return self.L.orc_cloud_n_instances(self.h)

    def state(self):
        n = self.n_instances
        arr = (InstanceState * n)()
        self.L.orc_cloud_get_state(self.h, arr)
        return arr

    def state_arrays(self):
        st = self.state()
        n = len(st)
        out = {k: np.array([getattr(s, k) for s in st], dtype=np.int32) for k in ("index", "copy", "active", "onMesh", "nShell", "N")}
        for k in ("center", "theta", "v", "omega", "F", "T"):
            out[k] = np.array([list(getattr(s, k)) for s in st], dtype=np.float64).reshape(n, 3)
        return out

    def shell(self, inst):
        st = self.state()[inst]
        n = st.nShell
        cells, jk = np.empty(n, np.int32), np.empty((n, 2), np.int32)
        Is, Os = np.empty((n, 8), np.int32), np.empty((n, 4), np.int32)
        w, W = np.empty((n, 8)), np.empty(n)
        self.L.orc_cloud_get_shell(self.h, inst, _ip(cells), _ip(jk), _ip(Is), _ip(Os), _dp(w), _dp(W))
        return dict(cells=cells, jk=jk, Is=Is, Os=Os, w=w, W=W)

    def points(self, inst):
        N = self.state()[inst].N
        pts, cell, nf = np.empty((N, 3)), np.empty(N, np.int32), np.empty(N, np.int32)
        oc, w, W = np.empty((N, 6), np.int32), np.empty((N, 6)), np.empty(N)
        self.L.orc_cloud_get_points(self.h, inst, _dp(pts), _ip(cell), _ip(nf), _ip(oc), _dp(w), _dp(W))
        return dict(pts=pts, cell=cell, nFaces=nf, otherCell=oc, w=w, W=W)

    def sf(self, inst):
        st = self.state()[inst]
        tab = self.case.particles
        i = st.index
        nk = tab.nTheta[i] // 2 if tab.kind[i] == 0 else 1
        out = np.empty((int(tab.nTheta[i]) * int(nk), 3))
        self.L.orc_cloud_get_sf(self.h, inst, _dp(out))
        return out

    def indicator(self):
        out = np.empty(self.mesh.n_cells)
        self.L.orc_cloud_indicator(self.h, _dp(out))
        return out

    def forcing(self, U, Uold, S, dt):
        U, Uold, S = _c(U), _c(Uold), _c(S)
        F = np.empty_like(U)
        n = self.L.orc_cloud_forcing(self.h, _dp(U), _dp(Uold), _dp(S), C.c_double(dt), _dp(F))
        if n < 0:
            raise ReferenceError_(self.L.orc_last_error().decode())
        return F, int(n)

    def integrate_stress(self, p, tau6, corrected=False):
        p, tau6 = _c(p), _c(tau6)
        FT = np.empty((self.n_instances, 6))
        Tc = np.empty((self.n_instances, 3)) if corrected else None
        self._chk(self.L.orc_cloud_integrate_stress(self.h, _dp(p), _dp(tau6), _dp(FT), _dp(Tc) if corrected else None))
        return (FT, Tc) if corrected else FT

    def set_force_torque(self, FT):
        FT = _c(FT)
        self.L.orc_cloud_set_force_torque(self.h, _dp(FT))

    def combine_forces(self):
        self.L.orc_cloud_combine_forces(self.h)

    def advance(self, dt):
        st = np.zeros(12)
        self.L.orc_cloud_advance(self.h, C.c_double(dt), _dp(st))
        return st

    def solve(self, p, tau6, dt):
        p, tau6 = _c(p), _c(tau6)
        st = np.zeros(12)
        self._chk(self.L.orc_cloud_solve(self.h, _dp(p), _dp(tau6), C.c_double(dt), _dp(st)))
        return st

    def substep(self, local_dt):
        self._chk(self.L.orc_cloud_substep(self.h, C.c_double(local_dt)))

    def update(self):
        self._chk(self.L.orc_cloud_update(self.h))

    def cd(self):
        out = np.empty(self.n_instances)
        self.L.orc_cloud_cd(self.h, _dp(out))
        return out

    def collision_tables(self, inst, n_particles, n_walls):
        """Dense Fc/Tc [nParticles,3] and Fw/Tw [nWalls,3] of one instance (reference build only)."""
        Fc, Tc = np.zeros((n_particles, 3)), np.zeros((n_particles, 3))
        Fw, Tw = np.zeros((max(n_walls, 1), 3)), np.zeros((max(n_walls, 1), 3))
        self.L.orc_particle_collision_tables(self.h, inst, _dp(Fc), _dp(Tc), _dp(Fw), _dp(Tw))
        return Fc, Tc, Fw[:n_walls], Tw[:n_walls]

    def written_fields(self):
        """lagrangian/IBMparticles/* as the reference's own writeFields() writes them (reference build only)."""
        self._chk(self.L.orc_cloud_write_fields(self.h))
        out = {}
        for name, width in (("positions", 3), ("index", 1), ("copy", 1), ("active", 1), ("center", 3), ("theta", 3), ("d", 1),
                            ("D", 3), ("v", 3), ("omega", 3), ("rho", 1), ("age", 1)):
            buf = np.zeros(3 * max(1, self.n_instances))
            n = self.L.orc_written_field(self.h, name.encode(), _dp(buf), buf.size)
            if n >= 0:
                out[name] = buf[:n].reshape(-1, width) if width > 1 else buf[:n].copy()
        return out

    def pair_collision(self, i1, i2):
        pt = np.zeros(3)
        hit = self.L.orc_pair_collision(self.h, int(i1), int(i2), _dp(pt))
        return bool(hit), pt

    def mass(self, inst):
        return self.L.orc_particle_mass(self.h, inst)

    def inertia(self, inst):
        return self.L.orc_particle_inertia(self.h, inst)

    def __del__(self):
        try:
            self.L.orc_cloud_destroy(self.h)
        except Exception:
            pass


def substep_schedule(dt, max_n=256):
    dts = np.zeros(max_n)
    n = lib().orc_substep_schedule(C.c_double(dt), _dp(dts), max_n)
    return dts[:n]
