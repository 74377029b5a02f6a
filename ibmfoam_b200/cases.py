"""Case descriptions for the immersed-boundary hot path (pure data, no compute).

Every config named in BASELINE.json / SURVEY.md §8(d) is expressed as a lattice mesh with
blockMesh cell numbering, a patch table and a particle table.  The same description feeds the
CUDA library (through the C ABI) and the CPU oracle (tests only), so both see identical inputs.

Reference sources of the numbers: validation/*/system/blockMeshDict, validation/*/constant/IBMProperties,
validation/sphere/system/initializeIBMDict; placement rule: applications/utilities/initializeIBM/
particleSizeDistributions/monodisperse/monodisperse.C:143-180.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SPHERE, CYLINDER, ELLIPSE = 0, 1, 2
PATCH, WALL, CYCLIC, EMPTY, PROCESSOR = 0, 1, 2, 3, 4
XMIN, XMAX, YMIN, YMAX, ZMIN, ZMAX = range(6)


@dataclass
class MeshSpec:
    nx: int
    ny: int
    nz: int
    origin: tuple
    delta: tuple
    blocks: list = field(default_factory=list)   # [(i0,i1,j0,j1,k0,k1)] in blockMesh order; [] = one block
    patches: list = field(default_factory=list)  # [(side, type, neighbour patch id or -1)]
    kz0: int = 0                                 # slab of a decomposed mesh: global z-plane of local plane 0

    def slab(self, k0: int, k1: int) -> "MeshSpec":
        """The processor mesh holding z-planes [k0,k1): same geometry (global plane index), z-sides that are interior
        cuts become processor patches (their faces are boundary faces, skipped by the gather like the reference's)."""
        assert not self.blocks, "slab decomposition needs a single-block lattice"
        pats = []
        for (side, typ, nb) in self.patches:
            if side == ZMIN and k0 > 0:
                pats.append((side, PROCESSOR, -1))
            elif side == ZMAX and k1 < self.nz:
                pats.append((side, PROCESSOR, -1))
            else:
                pats.append((side, typ, nb))
        return MeshSpec(self.nx, self.ny, k1 - k0, self.origin, self.delta, [], pats, kz0=self.kz0 + k0)

    @property
    def n_cells(self) -> int:
        return self.nx * self.ny * self.nz

    def cell_centres_lattice(self) -> np.ndarray:
        """Centres in lattice order (x fastest), same expression as the mini-mesh definition."""
        x = self.origin[0] + (np.arange(self.nx) + 0.5) * self.delta[0]
        y = self.origin[1] + (np.arange(self.ny) + 0.5) * self.delta[1]
        z = self.origin[2] + (np.arange(self.nz) + self.kz0 + 0.5) * self.delta[2]
        Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
        return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def lattice_to_cell(self) -> np.ndarray:
        """Caller (blockMesh) cell id of every lattice cell."""
        n = self.n_cells
        out = np.empty(n, dtype=np.int32)
        blocks = self.blocks or [(0, self.nx, 0, self.ny, 0, self.nz)]
        nxt = 0
        lat = np.arange(n, dtype=np.int64).reshape(self.nz, self.ny, self.nx)
        for (i0, i1, j0, j1, k0, k1) in blocks:
            sub = lat[k0:k1, j0:j1, i0:i1].ravel()
            out[sub] = np.arange(nxt, nxt + sub.size, dtype=np.int32)
            nxt += sub.size
        assert nxt == n, "blocks must tile the lattice"
        return out


@dataclass
class CloudSpec:
    moving: bool = False
    rotating: bool = False
    coupled: bool = True
    e: float = 1.0
    UInf: tuple = (0.0, 0.0, 0.0)
    rhoRef: float = 1.0
    pRef: float = 0.0
    g: tuple = (0.0, 0.0, 0.0)
    cyclic_without_walls: bool = False  # False = as coded


@dataclass
class ParticleTable:
    kind: np.ndarray
    position: np.ndarray
    d: np.ndarray
    a: np.ndarray
    b: np.ndarray
    nTheta: np.ndarray
    nk: np.ndarray
    delta: np.ndarray
    theta: np.ndarray
    v: np.ndarray
    omega: np.ndarray
    rho: np.ndarray

    @property
    def n(self) -> int:
        return int(self.kind.shape[0])

    @staticmethod
    def uniform(n, kind, position, d=0.0, a=0.0, b=0.0, nTheta=0, nk=1, delta=0.0, rho=1.0,
                theta=None, v=None, omega=None) -> "ParticleTable":
        z3 = np.zeros((n, 3))
        f = lambda x: np.full(n, x, dtype=np.float64)
        return ParticleTable(
            kind=np.full(n, kind, dtype=np.int32), position=np.asarray(position, dtype=np.float64).reshape(n, 3),
            d=f(d), a=f(a), b=f(b), nTheta=np.full(n, nTheta, dtype=np.int32), nk=np.full(n, nk, dtype=np.int32),
            delta=f(delta), theta=z3.copy() if theta is None else np.asarray(theta, dtype=np.float64).reshape(n, 3),
            v=z3.copy() if v is None else np.asarray(v, dtype=np.float64).reshape(n, 3),
            omega=z3.copy() if omega is None else np.asarray(omega, dtype=np.float64).reshape(n, 3), rho=f(rho))


@dataclass
class Case:
    name: str
    mesh: MeshSpec
    cloud: CloudSpec
    particles: ParticleTable
    dt: float = 1.0e-3
    note: str = ""


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def _four_blocks(nxup, nxdown, nyhalf):
    """The 4-block layout shared by cylinderRe30 and collidingParticles blockMeshDicts."""
    nx, ny = nxup + nxdown, 2 * nyhalf
    return [
        (0, nxup, 0, nyhalf, 0, 1),     # hex (0 1 8 7 ...)   lower-left
        (nxup, nx, 0, nyhalf, 0, 1),    # hex (1 2 3 8 ...)   lower-right
        (nxup, nx, nyhalf, ny, 0, 1),   # hex (8 3 4 5 ...)   upper-right
        (0, nxup, nyhalf, ny, 0, 1),    # hex (7 8 5 6 ...)   upper-left
    ]


def place_spheres(n, box_lo, box_hi, radius, seed=12345, periodic=True, max_tries=10_000_000):
    """Random sequential placement with the reference's non-overlap rule |xi-xj| > ri+rj
    (monodisperse.C:162-178), minimum-image when periodic.  PCG64 seed (the reference seeds from
    the wall clock, particleSizeDistribution.C:57 — not reproducible)."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(box_lo, dtype=np.float64)
    hi = np.asarray(box_hi, dtype=np.float64)
    L = hi - lo
    dmin = 2.0 * radius
    nb = np.maximum(1, np.floor(L / dmin).astype(int))
    bins: dict = {}
    pts = np.empty((n, 3))
    placed = 0
    tries = 0
    while placed < n:
        if tries >= max_tries:
            raise RuntimeError("sphere placement did not converge")
        batch = rng.random((4096, 3)) * L + lo
        for p in batch:
            tries += 1
            if not periodic and (np.any(p - radius < lo) or np.any(p + radius > hi)):
                continue
            ib = np.minimum(((p - lo) / L * nb).astype(int), nb - 1)
            ok = True
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        key = (ib[0] + dx, ib[1] + dy, ib[2] + dz)
                        if periodic:
                            key = (key[0] % nb[0], key[1] % nb[1], key[2] % nb[2])
                        for q in bins.get(key, ()):
                            dvec = p - pts[q]
                            if periodic:
                                dvec -= L * np.round(dvec / L)
                            if not (math.sqrt(dvec[0] ** 2 + dvec[1] ** 2 + dvec[2] ** 2) > dmin):
                                ok = False
                                break
                        if not ok:
                            break
                    if not ok:
                        break
                if not ok:
                    break
            if ok:
                pts[placed] = p
                bins.setdefault((ib[0], ib[1], ib[2]), []).append(placed)
                placed += 1
                if placed == n:
                    break
    return pts


# --------------------------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------------------------
def cylinder_re30() -> Case:
    """validation/cylinderRe30: 4-block 300x150x1 mesh, one fixed cylinder at the origin."""
    nx, ny = 300, 150
    mesh = MeshSpec(nx, ny, 1, origin=(-0.5, -0.5, -0.01), delta=(2.0 / nx, 1.0 / ny, 0.02),
                    blocks=_four_blocks(75, 225, 75),
                    patches=[(XMIN, PATCH, -1), (XMAX, PATCH, -1), (YMIN, CYCLIC, 3), (YMAX, CYCLIC, 2),
                             (ZMIN, EMPTY, -1), (ZMAX, EMPTY, -1)])
    cloud = CloudSpec(moving=False, rotating=False, coupled=True, e=1.0, UInf=(0.3, 0, 0), rhoRef=1.0)
    parts = ParticleTable.uniform(1, CYLINDER, [[0, 0, 0]], d=0.1, nTheta=100, nk=1, delta=0.01, rho=1000)
    return Case("cyl30", mesh, cloud, parts, dt=1e-3,
                note="coupled true added (missing in the case, App. B #16); rhoRef key ignored by the code")


def sphere_case(n=150, centre=(1.3e-5, -2.1e-5, 0.7e-5)) -> Case:
    """validation/sphere per its initializeIBMDict: n^3 box [-0.01,0.01]^3, d=0.004, nTheta=200, delta=1.3e-4."""
    L = 0.02
    mesh = MeshSpec(n, n, n, origin=(-0.01, -0.01, -0.01), delta=(L / n, L / n, L / n),
                    patches=[(XMIN, CYCLIC, 1), (XMAX, CYCLIC, 0), (YMIN, CYCLIC, 3), (YMAX, CYCLIC, 2),
                             (ZMIN, CYCLIC, 5), (ZMAX, CYCLIC, 4)])
    cloud = CloudSpec(coupled=True, UInf=(0.22, 0, 0), rhoRef=1.0)
    parts = ParticleTable.uniform(1, SPHERE, [list(centre)], d=0.004, nTheta=200, delta=1.3e-4, rho=1500)
    return Case("sph" if n == 150 else f"sph{n}", mesh, cloud, parts, dt=1e-4,
                note="sphere centre in generic position (the reference's own centre is undefined, App. B #1)")


def _read_colliding_positions():
    # positions of the 45 cylinders of validation/collidingParticles/constant/IBMProperties,
    # committed as a small fixture (tests/golden/colliding_positions.txt) with its generator.
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "data", "colliding_positions.txt")
    return np.loadtxt(path).reshape(-1, 3)


def colliding_particles(n_particles=None) -> Case:
    """validation/collidingParticles: 300x300x1 (4 blocks of 150x150), x walls, y cyclic, 45 moving cylinders."""
    nx = ny = 300
    mesh = MeshSpec(nx, ny, 1, origin=(-0.02, -0.02, -0.001), delta=(0.04 / nx, 0.04 / ny, 0.002),
                    blocks=_four_blocks(150, 150, 150),
                    patches=[(XMIN, WALL, -1), (XMAX, WALL, -1), (YMIN, CYCLIC, 3), (YMAX, CYCLIC, 2),
                             (ZMIN, EMPTY, -1), (ZMAX, EMPTY, -1)])
    cloud = CloudSpec(moving=True, rotating=True, coupled=True, e=1.0, UInf=(0, 0.22, 0), rhoRef=1.0,
                      g=(0.0, -9.81, 0.0))
    pos = _read_colliding_positions()
    if n_particles is not None:
        pos = pos[:n_particles]
    parts = ParticleTable.uniform(pos.shape[0], CYLINDER, pos, d=0.004, nTheta=300, nk=1, delta=0.00013, rho=1500)
    return Case("col", mesh, cloud, parts, dt=1e-4, note="g = (0 -9.81 0) supplied (no constant/g in the case)")


def two_spheres(n=64, e=1.0, speed=1.0, gap=0.5) -> Case:
    """Derived 3-D head-on pair: n^3 unit-spacing box, d = 8, closed-form answer (§8c-8)."""
    mesh = MeshSpec(n, n, n, origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0),
                    patches=[(XMIN, WALL, -1), (XMAX, WALL, -1), (YMIN, CYCLIC, 3), (YMAX, CYCLIC, 2),
                             (ZMIN, CYCLIC, 5), (ZMAX, CYCLIC, 4)])
    cloud = CloudSpec(moving=True, rotating=True, coupled=True, e=e, rhoRef=1.0)
    c = n / 2 + 0.123
    sep = 8.0 + gap
    pos = [[c - sep / 2, c + 0.211, c - 0.317], [c + sep / 2, c + 0.211, c - 0.317]]
    vel = [[speed, 0, 0], [-speed, 0, 0]]
    parts = ParticleTable.uniform(2, SPHERE, pos, d=8.0, nTheta=50, delta=1.0, rho=1.5, v=vel)
    return Case("two_spheres", mesh, cloud, parts, dt=0.25)


def assembly(n=128, phi=0.2, seed=12345, moving=False, d=8.0, n_particles=None, cyclic_without_walls=False, positions=None) -> Case:
    """Synthetic assembly (asm / big): n^3 periodic unit-spacing box of monodisperse spheres d = 8, delta = 1,
    nTheta = 50 (nk = 25).  asm: n=128, fixed, P=1565.  big: n=512, P=100000, moving+rotating."""
    mesh = MeshSpec(n, n, n, origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0),
                    patches=[(XMIN, CYCLIC, 1), (XMAX, CYCLIC, 0), (YMIN, CYCLIC, 3), (YMAX, CYCLIC, 2),
                             (ZMIN, CYCLIC, 5), (ZMAX, CYCLIC, 4)])
    vp = math.pi * d ** 3 / 6.0
    P = n_particles if n_particles is not None else int(round(phi * n ** 3 / vp))
    # positions: the result of an earlier place_spheres call with the same arguments (the placement is a pure-Python loop)
    pos = place_spheres(P, (0, 0, 0), (n, n, n), d / 2.0, seed=seed, periodic=True) if positions is None else np.asarray(positions, dtype=np.float64)
    assert pos.shape == (P, 3)
    rng = np.random.default_rng(seed + 1)
    U0 = 1.0
    if moving:
        v = rng.normal(0.0, 0.1 * U0, size=(P, 3))
        om = rng.normal(0.0, 0.1 * U0 / d, size=(P, 3))
    else:
        v = np.zeros((P, 3))
        om = np.zeros((P, 3))
    cloud = CloudSpec(moving=moving, rotating=moving, coupled=True, e=1.0, rhoRef=1.0, g=(0, 0, 0),
                      cyclic_without_walls=cyclic_without_walls)
    parts = ParticleTable.uniform(P, SPHERE, pos, d=d, nTheta=50, delta=1.0, rho=1.5, v=v, omega=om)
    return Case(f"asm{n}" if not moving else f"big{n}", mesh, cloud, parts, dt=0.25)


# --------------------------------------------------------------------------------------------
# synthetic frozen fields: Taylor-Green type, sampled at cell centres, lattice order
# --------------------------------------------------------------------------------------------
def synthetic_fields(mesh: MeshSpec, U0=1.0, nu=0.01, xp=np, centres=None):
    """Returns dict of lattice-ordered fields: U, Uold, S [C,3]; p [C]; tau [C,6] (xx xy xz yy yz zz)."""
    C = mesh.cell_centres_lattice() if centres is None else centres
    Lx = mesh.nx * mesh.delta[0]
    Ly = mesh.ny * mesh.delta[1]
    Lz = mesh.nz * mesh.delta[2]
    kx, ky, kz = 2 * math.pi / Lx, 2 * math.pi / Ly, 2 * math.pi / Lz
    x = (C[:, 0] - mesh.origin[0]) * kx
    y = (C[:, 1] - mesh.origin[1]) * ky
    z = (C[:, 2] - mesh.origin[2]) * kz if mesh.nz > 1 else xp.zeros_like(x)

    def vel(ph):
        return xp.stack([U0 * xp.sin(x + ph) * xp.cos(y) * xp.cos(z),
                         -U0 * xp.cos(x + ph) * xp.sin(y) * xp.cos(z),
                         xp.zeros_like(x)], axis=1)

    U = vel(0.0)
    Uold = vel(0.01)
    p = (U0 * U0 / 16.0) * (xp.cos(2 * x) + xp.cos(2 * y)) * (xp.cos(2 * z) + 2.0)
    # S ~ convective + pressure-gradient term of the field (smooth surrogate; frozen-field tests only)
    S = xp.stack([0.5 * U0 * U0 * kx * xp.sin(2 * x) * xp.cos(z) ** 2 - (U0 * U0 / 8.0) * kx * xp.sin(2 * x) * (xp.cos(2 * z) + 2.0),
                  0.5 * U0 * U0 * ky * xp.sin(2 * y) * xp.cos(z) ** 2 - (U0 * U0 / 8.0) * ky * xp.sin(2 * y) * (xp.cos(2 * z) + 2.0),
                  -(U0 * U0 / 8.0) * kz * (xp.cos(2 * x) + xp.cos(2 * y)) * xp.sin(2 * z)], axis=1)
    # tau = -nu dev(twoSymm(grad U)) of the analytic U (divergence-free => dev is a no-op up to round-off)
    dudx = U0 * kx * xp.cos(x) * xp.cos(y) * xp.cos(z)
    dudy = -U0 * ky * xp.sin(x) * xp.sin(y) * xp.cos(z)
    dudz = -U0 * kz * xp.sin(x) * xp.cos(y) * xp.sin(z)
    dvdx = U0 * kx * xp.sin(x) * xp.sin(y) * xp.cos(z)
    dvdy = -U0 * ky * xp.cos(x) * xp.cos(y) * xp.cos(z)
    dvdz = U0 * kz * xp.cos(x) * xp.sin(y) * xp.sin(z)
    tr = (dudx + dvdy) / 3.0
    tau = xp.stack([-nu * (2 * dudx - 2 * tr), -nu * (dudy + dvdx), -nu * dudz,
                    -nu * (2 * dvdy - 2 * tr), -nu * dvdz, -nu * (-2 * tr)], axis=1)
    return {"U": U, "Uold": Uold, "S": S, "p": p, "tau": tau}
