"""OpenFOAM-free mini incompressible flow solver - VALIDATION ONLY.

The immersed-boundary path needs a caller on either side of it: something that produces U, U.oldTime() and the explicit
momentum terms S before `forcing(S)`, and p and devRhoReff before `solve()`.  In the reference that caller is
pimpleIBMFoam (applications/solvers/pimpleIBMFoam/{UEqn.H,pEqn.H}), i.e. OpenFOAM, which stays out of scope and is
not available here.  This module is a small stand-in with the same call sites so that the drag validation cases
(validation/cylinderRe30 and a periodic sphere) can be run end to end through ANY object with the particleCloud
interface (`forcing(S, U, Uold, dt)`, `solve(p, devRhoReff, dt)`, `Cd()`): the CUDA path (ibmfoam_b200.cloud.ParticleCloud)
or, in the tests, the reference's own code.  It is host-side numpy, not a hot path, and makes no accuracy claim beyond
"second-order central differences, explicit Euler, approximate projection on a collocated uniform grid".

Per time step (the reference's call order, UEqn.H:1-36, pEqn.H, pimpleIBMFoam.C:58-93):
    S   = div(phi, U) + grad(p) - nu lap(U)                     # UEqn.H:23-25 (laminar: divDevReff(U) & U = -nu lap U)
    U*  = U^n - dt S                                            # momentum predictor
    F   = particles.forcing(S) with U = U*, U.oldTime() = U^n   # UEqn.H:20-26
    U** = U^n + dt (F - S)                                      # UEqn -= F, second momentum solve
    lap(phi) = div(U**)/dt;  U^{n+1} = U** - dt grad(phi);  p += phi      # pEqn.H (one corrector)
    particles.solve() with p^{n+1}, devRhoReff = -nu dev(twoSymm(grad U^{n+1}))   # pimpleIBMFoam.C:87
Fields are cell-centred arrays in LATTICE order (x fastest); `to_caller` maps them to the caller's cell order.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only needed for the inlet/outlet Poisson problem
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
except Exception:  # pragma: no cover
    sp = spla = None


class Channel2D:
    """nx x ny uniform cells; x: velocity inlet (U_inf, 0) / pressure outlet; y: periodic.  validation/cylinderRe30."""

    def __init__(self, nx, ny, h, U_inf, nu, dt):
        self.nx, self.ny, self.h, self.Uin, self.nu, self.dt = nx, ny, float(h), float(U_inf), float(nu), float(dt)
        self.u = np.full((ny, nx), float(U_inf))
        self.v = np.zeros((ny, nx))
        self.p = np.zeros((ny, nx))
        self.uf = np.full((ny, nx + 1), float(U_inf))   # face-normal velocities at i-1/2 .. i+1/2 (divergence free)
        self.vf = np.zeros((ny, nx))                    # at j+1/2 (periodic)
        self._lu = self._poisson()

    # ---- ghost-cell helpers (x direction); y is periodic through np.roll ----
    def _xpad(self, a, kind):
        left = {"u": 2 * self.Uin - a[:, :1], "v": -a[:, :1], "p": a[:, :1], "phi": a[:, :1]}[kind]
        right = {"u": a[:, -1:], "v": a[:, -1:], "p": -a[:, -1:], "phi": -a[:, -1:]}[kind]
        return np.concatenate([left, a, right], axis=1)

    def _poisson(self):
        nx, ny, h = self.nx, self.ny, self.h
        n = nx * ny
        idx = np.arange(n).reshape(ny, nx)
        rows, cols, vals = [], [], []

        def add(r, c, v):
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(np.full(r.size, v))

        diag = np.full((ny, nx), -4.0)
        diag[:, 0] += 1.0    # Neumann at the inlet: ghost = first cell
        diag[:, -1] -= 1.0   # Dirichlet 0 at the outlet: ghost = -last cell
        rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(diag.ravel())
        add(idx[:, 1:], idx[:, :-1], 1.0)
        add(idx[:, :-1], idx[:, 1:], 1.0)
        add(idx, np.roll(idx, 1, axis=0), 1.0)
        add(idx, np.roll(idx, -1, axis=0), 1.0)
        A = sp.csc_matrix((np.concatenate(vals) / (h * h), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        return spla.splu(A)

    # ---- explicit momentum terms ----
    def explicit_terms(self):
        h, nu = self.h, self.nu
        up, vp, pp = self._xpad(self.u, "u"), self._xpad(self.v, "v"), self._xpad(self.p, "p")
        uf, vf = self.uf, self.vf

        def conv(ap):
            a = ap[:, 1:-1]
            ax_face = 0.5 * (ap[:, :-1] + ap[:, 1:])                     # value at x faces (nx+1)
            ay_face = 0.5 * (a + np.roll(a, -1, axis=0))                 # value at j+1/2
            return (uf[:, 1:] * ax_face[:, 1:] - uf[:, :-1] * ax_face[:, :-1]) / h + (vf * ay_face - np.roll(vf * ay_face, 1, axis=0)) / h

        def lap(ap):
            a = ap[:, 1:-1]
            return (ap[:, 2:] + ap[:, :-2] + np.roll(a, 1, axis=0) + np.roll(a, -1, axis=0) - 4 * a) / (h * h)

        dpdx = (pp[:, 2:] - pp[:, :-2]) / (2 * h)
        dpdy = (np.roll(self.p, -1, axis=0) - np.roll(self.p, 1, axis=0)) / (2 * h)
        Sx = conv(up) + dpdx - nu * lap(up)
        Sy = conv(vp) + dpdy - nu * lap(vp)
        return Sx, Sy

    def project(self, us, vs):
        """approximate projection: div-free face fluxes from a compact Poisson problem, central correction at the centres"""
        h, dt = self.h, self.dt
        up = self._xpad(us, "u")
        uf = 0.5 * (up[:, :-1] + up[:, 1:])
        uf[:, 0] = self.Uin
        vf = 0.5 * (vs + np.roll(vs, -1, axis=0))
        div = (uf[:, 1:] - uf[:, :-1]) / h + (vf - np.roll(vf, 1, axis=0)) / h
        phi = self._lu.solve((div / dt).ravel()).reshape(self.ny, self.nx)
        php = self._xpad(phi, "phi")
        self.uf = uf - dt * (php[:, 1:] - php[:, :-1]) / h
        self.uf[:, 0] = self.Uin
        self.vf = vf - dt * (np.roll(phi, -1, axis=0) - phi) / h
        self.u = us - dt * (php[:, 2:] - php[:, :-2]) / (2 * h)
        self.v = vs - dt * (np.roll(phi, -1, axis=0) - np.roll(phi, 1, axis=0)) / (2 * h)
        self.p = self.p + phi

    def dev_rho_reff(self, rho=1.0):
        """-rho nu dev(twoSymm(grad U)) at the cell centres, symmTensor component order xx xy xz yy yz zz"""
        h, nu = self.h, self.nu
        up, vp = self._xpad(self.u, "u"), self._xpad(self.v, "v")
        ux = (up[:, 2:] - up[:, :-2]) / (2 * h)
        vx = (vp[:, 2:] - vp[:, :-2]) / (2 * h)
        uy = (np.roll(self.u, -1, axis=0) - np.roll(self.u, 1, axis=0)) / (2 * h)
        vy = (np.roll(self.v, -1, axis=0) - np.roll(self.v, 1, axis=0)) / (2 * h)
        tr3 = 2.0 * (ux + vy) / 3.0
        z = np.zeros_like(ux)
        k = -rho * nu
        return np.stack([k * (2 * ux - tr3), k * (uy + vx), z, k * (2 * vy - tr3), z, k * (-tr3)], axis=-1).reshape(-1, 6)

    def vec(self, ax, ay):
        return np.stack([ax.ravel(), ay.ravel(), np.zeros(ax.size)], axis=1)


class PeriodicBox3D:
    """n^3 uniform cells, periodic in all directions (validation/sphere's blockMeshDict: every patch cyclic), with the mean
    velocity held at U_inf by a uniform body force.  In steady state that force balances the drag of the body:
    F_drag = rho * f_body * V_box - an independent check of the surface-stress integral."""

    def __init__(self, n, h, U_inf, nu, dt):
        self.n, self.h, self.nu, self.dt = n, float(h), float(nu), float(dt)
        self.Uinf = np.asarray(U_inf, dtype=np.float64)
        self.U = [np.full((n, n, n), float(self.Uinf[c])) for c in range(3)]   # index [k][j][i]
        self.p = np.zeros((n, n, n))
        self.Uf = [np.full((n, n, n), float(self.Uinf[c])) for c in range(3)]  # face-normal velocity at +1/2 in direction c
        k = 2.0 * np.pi * np.fft.fftfreq(n)
        ev = (2.0 * np.cos(k) - 2.0) / (h * h)
        self._eig = ev[:, None, None] + ev[None, :, None] + ev[None, None, :]
        self._eig[0, 0, 0] = 1.0
        self.body_force = np.zeros(3)

    @staticmethod
    def _ax(c):           # array axis of direction c (x = last axis)
        return 2 - c

    def explicit_terms(self):
        h, nu = self.h, self.nu
        S = []
        for c in range(3):
            a = self.U[c]
            conv = np.zeros_like(a)
            lap = -6.0 * a
            for d in range(3):
                ax = self._ax(d)
                face = 0.5 * (a + np.roll(a, -1, axis=ax)) * self.Uf[d]
                conv += (face - np.roll(face, 1, axis=ax)) / h
                lap += np.roll(a, 1, axis=ax) + np.roll(a, -1, axis=ax)
            axc = self._ax(c)
            dp = (np.roll(self.p, -1, axis=axc) - np.roll(self.p, 1, axis=axc)) / (2 * h)
            S.append(conv + dp - nu * lap / (h * h))
        return S

    def project(self, Us):
        h, dt = self.h, self.dt
        Uf = [0.5 * (Us[d] + np.roll(Us[d], -1, axis=self._ax(d))) for d in range(3)]
        div = sum((Uf[d] - np.roll(Uf[d], 1, axis=self._ax(d))) / h for d in range(3))
        ph = np.fft.fftn(div / dt) / self._eig
        ph[0, 0, 0] = 0.0
        phi = np.real(np.fft.ifftn(ph))
        for d in range(3):
            ax = self._ax(d)
            self.Uf[d] = Uf[d] - dt * (np.roll(phi, -1, axis=ax) - phi) / h
            self.U[d] = Us[d] - dt * (np.roll(phi, -1, axis=ax) - np.roll(phi, 1, axis=ax)) / (2 * h)
        self.p = self.p + phi
        # hold the mean velocity: a uniform body force f = (U_inf - <U>)/dt   (applied to centres and faces alike)
        for d in range(3):
            corr = self.Uinf[d] - self.U[d].mean()
            self.body_force[d] = corr / dt
            self.U[d] += corr
            self.Uf[d] += corr

    def dev_rho_reff(self, rho=1.0):
        h = self.h
        g = [[(np.roll(self.U[c], -1, axis=self._ax(d)) - np.roll(self.U[c], 1, axis=self._ax(d))) / (2 * h) for d in range(3)] for c in range(3)]
        tr3 = 2.0 * (g[0][0] + g[1][1] + g[2][2]) / 3.0
        k = -rho * self.nu
        comps = [k * (2 * g[0][0] - tr3), k * (g[0][1] + g[1][0]), k * (g[0][2] + g[2][0]), k * (2 * g[1][1] - tr3), k * (g[1][2] + g[2][1]),
                 k * (2 * g[2][2] - tr3)]
        return np.stack([c.ravel() for c in comps], axis=1)

    def vec(self, comps):
        return np.stack([c.ravel() for c in comps], axis=1)


def run_sphere(cloud, flow: PeriodicBox3D, n_steps, every=1, log=None):
    """Same call sites as run_cylinder for a body in the periodic box.  Returns [(step, Cd from the cloud, body force)]."""
    out = []
    dt = flow.dt
    shape = flow.p.shape
    for step in range(1, n_steps + 1):
        S = flow.explicit_terms()
        Us = [flow.U[c] - dt * S[c] for c in range(3)]
        F = cloud.forcing(flow.vec(S), flow.vec(Us), flow.vec(flow.U), dt)
        flow.project([flow.U[c] + dt * (F[:, c].reshape(shape) - S[c]) for c in range(3)])
        if step % every == 0 or step == n_steps:
            cloud.solve(flow.p.ravel(), flow.dev_rho_reff(), dt)
            out.append((step, float(np.asarray(cloud.Cd())[0]), flow.body_force.copy()))
            if log:
                log(*out[-1])
    return out


def run_moving_bodies(cloud, flow: PeriodicBox3D, n_steps, state=None):
    """validation/collidingParticles-type loop: freely moving bodies in the periodic box - forcing(S) with the rigid-body
    target velocity, then solve() = stress integral, force combination, sub-stepped collisions / integration, re-marking.
    `state(cloud)` (optional) is called after every step and its results are collected."""
    out = []
    dt = flow.dt
    shape = flow.p.shape
    for _ in range(n_steps):
        S = flow.explicit_terms()
        Us = [flow.U[c] - dt * S[c] for c in range(3)]
        F = cloud.forcing(flow.vec(S), flow.vec(Us), flow.vec(flow.U), dt)
        flow.project([flow.U[c] + dt * (F[:, c].reshape(shape) - S[c]) for c in range(3)])
        cloud.solve(flow.p.ravel(), flow.dev_rho_reff(), dt)
        if state is not None:
            out.append(state(cloud))
    return out


def run_cylinder(cloud, flow: Channel2D, n_steps, to_caller=None, from_caller=None, every=1, log=None):
    """Time loop with the reference's call sites.  `cloud` has forcing(S, U, Uold, dt) -> F [C,3], solve(p, tau6, dt),
    Cd() -> per-instance drag coefficients.  Cell fields cross the interface in the CALLER's cell order."""
    to_caller = to_caller or (lambda a: a)
    from_caller = from_caller or (lambda a: a)
    cds = []
    dt = flow.dt
    for step in range(1, n_steps + 1):
        Sx, Sy = flow.explicit_terms()
        us, vs = flow.u - dt * Sx, flow.v - dt * Sy                        # predictor
        F = from_caller(cloud.forcing(to_caller(flow.vec(Sx, Sy)), to_caller(flow.vec(us, vs)), to_caller(flow.vec(flow.u, flow.v)), dt))
        Fx, Fy = F[:, 0].reshape(flow.ny, flow.nx), F[:, 1].reshape(flow.ny, flow.nx)
        flow.project(flow.u + dt * (Fx - Sx), flow.v + dt * (Fy - Sy))
        if step % every == 0 or step == n_steps:
            cloud.solve(to_caller(flow.p.ravel()), to_caller(flow.dev_rho_reff()), dt)
            cds.append((step, float(np.asarray(cloud.Cd())[0])))
            if log:
                log(step, cds[-1][1])
    return cds
