"""Python mirror of the reference's particleCloud interface over the C ABI.

Method names and argument meaning follow src/particleCloud/particleCloud.H:267-358: the constructor takes the
mesh, the IBMProperties content and g; `forcing(S)` needs U and U.oldTime() too because there is no mesh
registry to look them up from (particleCloud.C:1193-1194); `solve()` needs p and devRhoReff for the same
reason (particleCloud.C:1212,1221).  Cell fields are numpy arrays in the CALLER's cell order, AoS, exactly as
OpenFOAM stores them.  Errors surface as IbmError (the reference aborts through FatalError).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import IbmError, check


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class ParticleCloud:
    def __init__(self, case=None, *, mesh=None, cloud=None, particles=None, device=0, max_instances=0,
                 slab=None, halo=0, local_group=None, nccl_id=None, rank=0, n_ranks=1):
        """slab=(k0,k1): this rank's z-planes of the (whole) mesh; halo: read-only planes of the neighbouring slabs on either
        side (ibm_set_slab_halo: the fields then cover self.local_planes); local_group / nccl_id select the transport of
        the all-reduces (multi-GPU: slab-decomposed fields, particle directory on every rank)."""
        L = capi.load()
        if case is not None:
            mesh, cloud, particles = case.mesh, case.cloud, case.particles
        self.mesh_spec, self.cloud_spec, self.particle_table = mesh, cloud, particles
        self._keep = []
        md = capi.MeshDesc()
        md.nx, md.ny, md.nz = mesh.nx, mesh.ny, mesh.nz
        md.origin[:] = mesh.origin
        md.delta[:] = mesh.delta
        if mesh.blocks:
            blocks = _c(np.asarray(mesh.blocks).reshape(-1, 6), np.int32)
            self._keep.append(blocks)
            md.nBlocks = blocks.shape[0]
            md.blocks = blocks.ctypes.data_as(capi.c_int_p)
        pats = (capi.PatchDesc * max(1, len(mesh.patches)))()
        for i, (side, typ, nb) in enumerate(mesh.patches):
            pats[i].type, pats[i].side, pats[i].neighbPatch, pats[i].nFaces = typ, side, nb, 0
        md.nPatches = len(mesh.patches)
        md.patches = pats
        self._keep.append(pats)
        cfg = capi.CloudCfg()
        cfg.moving, cfg.rotating, cfg.coupled = int(cloud.moving), int(cloud.rotating), int(cloud.coupled)
        cfg.e = cloud.e
        cfg.UInf[:] = cloud.UInf
        cfg.rhoRef, cfg.pRef = cloud.rhoRef, cloud.pRef
        cfg.g[:] = cloud.g
        cfg.cyclicWithoutWalls = int(cloud.cyclic_without_walls)
        cfg.maxInstances = int(max_instances)
        h = C.c_void_p()
        check(L.ibm_create(C.byref(md), C.byref(cfg), int(device), C.byref(h)))
        self._h = h
        self._L = L
        self.n_cells = mesh.n_cells
        self.slab = slab
        self.local_planes = (0, mesh.nz) if slab is None else (int(slab[0]), int(slab[1]))
        if slab is not None:
            if halo:
                lo, hi = C.c_int(), C.c_int()
                check(L.ibm_set_slab_halo(h, int(slab[0]), int(slab[1]), int(halo), C.byref(lo), C.byref(hi)), h)
                self.local_planes = (lo.value, hi.value)  # planes the cell fields of this rank cover (owned + halo)
            else:
                check(L.ibm_set_slab(h, int(slab[0]), int(slab[1])), h)
            self.n_cells = mesh.nx * mesh.ny * (self.local_planes[1] - self.local_planes[0])
        if local_group is not None:
            check(L.ibm_comm_init_local(h, local_group, int(rank)), h)
        elif nccl_id is not None:
            buf = C.create_string_buffer(bytes(nccl_id), L.ibm_comm_unique_id_bytes())
            check(L.ibm_comm_init(h, buf, int(rank), int(n_ranks)), h)
        n = particles.n
        inits = (capi.ParticleInit * max(1, n))()
        for i in range(n):
            a = inits[i]
            a.kind = int(particles.kind[i])
            a.position[:] = particles.position[i]
            a.d, a.a, a.b = float(particles.d[i]), float(particles.a[i]), float(particles.b[i])
            a.nTheta, a.nk = int(particles.nTheta[i]), int(particles.nk[i])
            a.delta = float(particles.delta[i])
            a.theta[:] = particles.theta[i]
            a.v[:] = particles.v[i]
            a.omega[:] = particles.omega[i]
            a.rho = float(particles.rho[i])
        check(L.ibm_set_particles(h, n, inits), h)

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self._L.ibm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- particleCloud API ----
    def nParticles(self) -> int:
        n = C.c_int()
        check(self._L.ibm_n_instances(self._h, C.byref(n)), self._h)
        return n.value

    def maxCoNum(self) -> float:
        v = C.c_double()
        check(self._L.ibm_max_co_num(self._h, C.byref(v)), self._h)
        return v.value

    def update(self):
        check(self._L.ibm_update(self._h), self._h)

    def forcing(self, S, U, Uold, dt):
        """particleCloud::forcing(S): returns F [nCells,3] (caller order)."""
        U, Uold, S = _c(U), _c(Uold), _c(S)
        F = np.empty((self.n_cells, 3))
        check(self._L.ibm_forcing(self._h, _ptr(U), _ptr(Uold), _ptr(S), float(dt), _ptr(F)), self._h)
        return F

    def integrateSurfaceStresses(self, p, devRhoReff):
        p, tau = _c(p), _c(devRhoReff)
        FT = np.empty((self.nParticles(), 6))
        check(self._L.ibm_integrate_stress(self._h, _ptr(p), _ptr(tau), _ptr(FT)), self._h)
        return FT

    def combineForces(self):
        check(self._L.ibm_combine_forces(self._h), self._h)

    def advance(self, dt):
        st = capi.SolveStats()
        check(self._L.ibm_advance(self._h, float(dt), C.byref(st)), self._h)
        return st

    def solve(self, p, devRhoReff, dt):
        p, tau = _c(p), _c(devRhoReff)
        st = capi.SolveStats()
        check(self._L.ibm_solve(self._h, _ptr(p), _ptr(tau), float(dt), C.byref(st)), self._h)
        return st

    def substep(self, local_dt):
        check(self._L.ibm_substep(self._h, float(local_dt)), self._h)

    def Cd(self):
        out = np.empty(self.nParticles())
        check(self._L.ibm_cd(self._h, _ptr(out)), self._h)
        return out

    # ---- accessors / taps ----
    def particles(self):
        n = self.nParticles()
        arr = (capi.ParticleState * max(1, n))()
        check(self._L.ibm_get_particles(self._h, arr), self._h)
        return arr[:n] if n else []

    def state_arrays(self):
        st = self.particles()
        n = len(st)
        out = {k: np.array([getattr(s, k) for s in st], dtype=np.int32) for k in ("index", "copy", "active", "onMesh", "nShell", "N")}
        for k in ("center", "theta", "v", "omega", "F", "T"):
            out[k] = np.array([list(getattr(s, k)) for s in st], dtype=np.float64).reshape(n, 3)
        return out

    def restore(self, states):
        """particleCloud::readParticles(): replace the instance list by the given ParticleState records (see ibm_restore_particles)"""
        arr = (capi.ParticleState * max(1, len(states)))(*states)
        check(self._L.ibm_restore_particles(self._h, len(states), arr), self._h)

    def set_force_torque(self, FT):
        FT = _c(FT)
        check(self._L.ibm_set_force_torque(self._h, _ptr(FT)), self._h)

    def mass_inertia(self, inst):
        m, i = C.c_double(), C.c_double()
        check(self._L.ibm_particle_mass(self._h, int(inst), C.byref(m), C.byref(i)), self._h)
        return m.value, i.value

    def shell(self, inst, n_shell=None):
        if n_shell is None:
            n_shell = self.particles()[inst].nShell
        n = int(n_shell)
        cells, jk = np.empty(n, np.int32), np.empty((n, 2), np.int32)
        Is, Os = np.empty((n, 8), np.int32), np.empty((n, 4), np.int32)
        w, W = np.empty((n, 8)), np.empty(n)
        check(self._L.ibm_get_shell(self._h, int(inst), _ptr(cells), _ptr(jk), _ptr(Is), _ptr(Os), _ptr(w), _ptr(W)), self._h)
        return dict(cells=cells, jk=jk, Is=Is, Os=Os, w=w, W=W)

    def points(self, inst, N=None):
        if N is None:
            N = self.particles()[inst].N
        N = int(N)
        pts, cell, nf = np.empty((N, 3)), np.empty(N, np.int32), np.empty(N, np.int32)
        oc, w, W = np.empty((N, 6), np.int32), np.empty((N, 6)), np.empty(N)
        check(self._L.ibm_get_points(self._h, int(inst), _ptr(pts), _ptr(cell), _ptr(nf), _ptr(oc), _ptr(w), _ptr(W)), self._h)
        return dict(pts=pts, cell=cell, nFaces=nf, otherCell=oc, w=w, W=W)

    def sf(self, inst, n):
        out = np.empty((int(n), 3))
        check(self._L.ibm_get_sf(self._h, int(inst), _ptr(out)), self._h)
        return out

    def indicator(self):
        out = np.empty(self.n_cells)
        check(self._L.ibm_get_indicator(self._h, _ptr(out)), self._h)
        return out

    def solid_fraction(self, n_sub=8):
        """EXTENSION (no reference counterpart): fractional solid volume fraction per cell, caller cell order."""
        out = np.empty(self.mesh_spec.n_cells)
        check(self._L.ibm_get_solid_fraction(self._h, int(n_sub), _ptr(out)), self._h)
        return out

    def set_kernel_variant(self, forcing=1, stress=1):
        check(self._L.ibm_set_kernel_variant(self._h, int(forcing), int(stress)), self._h)

    def set_option(self, name, value):
        check(self._L.ibm_set_option(self._h, name.encode(), int(value)), self._h)

    def set_real_option(self, name, value):
        check(self._L.ibm_set_real_option(self._h, name.encode(), float(value)), self._h)

    def soft_sphere(self, kn, gamma_n=0.0, gamma_t=0.0, mu=0.0):
        """EXTENSION (no counterpart in the reference): spring - dashpot - slider contact law for sphere / cylinder pairs
        instead of the hard-sphere impulse; see include/ibmfoam_b200.h at ibm_set_real_option."""
        for k, v in (("soft_sphere_kn", kn), ("soft_sphere_gamma_n", gamma_n), ("soft_sphere_gamma_t", gamma_t), ("soft_sphere_mu", mu)):
            self.set_real_option(k, v)
        self.set_option("collision_model", 1)

    def counter(self, name):
        v = C.c_longlong()
        check(self._L.ibm_get_counter(self._h, name.encode(), C.byref(v)), self._h)
        return v.value

    def sync(self):
        check(self._L.ibm_sync(self._h), self._h)

    def timings(self, reset=False):
        t = capi.Timing()
        check(self._L.ibm_timings(self._h, C.byref(t)), self._h)
        out = {k: getattr(t, k) for k, _ in capi.Timing._fields_}
        if reset:
            check(self._L.ibm_timings_reset(self._h), self._h)
        return out

    # ---- device-resident entry points (bench / multi-GPU harness); pointers are raw device addresses ----
    def forcing_device(self, dU, dUold, dS, dt, dF=None):
        check(self._L.ibm_forcing_device(self._h, C.c_void_p(dU), C.c_void_p(dUold), C.c_void_p(dS), float(dt),
                                         C.c_void_p(dF) if dF else None), self._h)

    def integrate_stress_device(self, dP, dTau, FT=None):
        check(self._L.ibm_integrate_stress_device(self._h, C.c_void_p(dP), C.c_void_p(dTau), _ptr(FT) if FT is not None else None), self._h)

    def solve_device(self, dP, dTau, dt):
        st = capi.SolveStats()
        check(self._L.ibm_solve_device(self._h, C.c_void_p(dP), C.c_void_p(dTau), float(dt), C.byref(st)), self._h)
        return st

    def upload_field(self, host_aos, n_comp, d_soa):
        host_aos = _c(host_aos)
        check(self._L.ibm_upload_field(self._h, _ptr(host_aos), int(n_comp), C.c_void_p(d_soa)), self._h)

    def download_field(self, d_soa, n_comp, out=None):
        if out is None:
            out = np.empty((self.n_cells, n_comp) if n_comp > 1 else (self.n_cells,))
        check(self._L.ibm_download_field(self._h, C.c_void_p(d_soa) if d_soa else None, int(n_comp), _ptr(out)), self._h)
        return out


def nccl_unique_id() -> bytes:
    L = capi.load()
    buf = C.create_string_buffer(L.ibm_comm_unique_id_bytes())
    check(L.ibm_comm_get_unique_id(buf))
    return buf.raw


class LocalGroup:
    """In-process transport for tests: ranks are host threads sharing one device."""

    def __init__(self, n):
        self.n = n
        self.handle = C.c_void_p()
        check(capi.load().ibm_comm_local_group_create(int(n), C.byref(self.handle)))

    def close(self):
        if self.handle:
            capi.load().ibm_comm_local_group_destroy(self.handle)
            self.handle = None


def slab_ranges(nz, n_ranks):
    """Contiguous z-plane ranges, as even as possible (first ranks take the remainder)."""
    base, rem = divmod(nz, n_ranks)
    out, k = [], 0
    for r in range(n_ranks):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def substep_schedule(dt, max_n=256):
    dts = np.zeros(max_n)
    n = capi.load().ibm_substep_schedule(float(dt), _ptr(dts), max_n)
    return dts[:n]
