"""Shared helpers for the parity tests (oracle = checker, CUDA library = thing under test)."""
import numpy as np

from ibmfoam_b200 import cases


def caller_fields(case, **kw):
    """Synthetic frozen fields in CALLER cell order (AoS), like OpenFOAM internal fields."""
    f = cases.synthetic_fields(case.mesh, **kw)
    l2c = case.mesh.lattice_to_cell()
    out = {}
    for k, v in f.items():
        o = np.empty_like(v)
        o[l2c] = v
        out[k] = o
    return out


def rel_err(a, b):
    """max |a-b| / max |b|  (field-norm relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def small_assembly(n=32, n_particles=12, seed=7, moving=False, d=8.0):
    return cases.assembly(n=n, n_particles=n_particles, seed=seed, moving=moving, d=d)


def ellipse_case():
    """2-D lattice with two rotated ellipses (no reference validation case uses them; shape code only)."""
    nx = ny = 96
    mesh = cases.MeshSpec(nx, ny, 1, origin=(-0.3, -0.3, -0.01), delta=(0.6 / nx, 0.6 / ny, 0.02),
                          patches=[(cases.XMIN, cases.WALL, -1), (cases.XMAX, cases.WALL, -1), (cases.YMIN, cases.CYCLIC, 3),
                                   (cases.YMAX, cases.CYCLIC, 2), (cases.ZMIN, cases.EMPTY, -1), (cases.ZMAX, cases.EMPTY, -1)])
    cloud = cases.CloudSpec(moving=True, rotating=True, coupled=True, e=1.0, UInf=(0.3, 0, 0))
    parts = cases.ParticleTable.uniform(2, cases.ELLIPSE, [[-0.0713, 0.0131, 0.0], [0.0917, -0.0277, 0.0]], a=0.06, b=0.035,
                                        nTheta=120, nk=1, delta=0.0125, rho=1200,
                                        theta=[[0, 0, 0.37], [0, 0, 2.05]], v=[[0.1, 0.02, 0], [-0.1, 0.01, 0]],
                                        omega=[[0, 0, 0.5], [0, 0, -0.8]])
    return cases.Case("ellipse2", mesh, cloud, parts, dt=1e-3)


class OracleAsCloud:
    """The CPU checker behind the particleCloud call interface the mini flow solver drives (tests only)."""

    def __init__(self, case, impl="port", **kw):
        from oracle.pyoracle import OracleCloud
        self.oc = OracleCloud(case, impl=impl, **kw)

    def forcing(self, S, U, Uold, dt):
        return self.oc.forcing(U, Uold, S, dt)[0]

    def solve(self, p, devRhoReff, dt):
        return self.oc.solve(p, devRhoReff, dt)

    def Cd(self):
        return self.oc.cd()


def cylinder_re30_setup(dt=1e-3):
    """validation/cylinderRe30 for the mini flow solver: the case, the flow object and the lattice <-> caller cell maps"""
    import numpy as np
    from ibmfoam_b200.validation import mini_flow
    case = cases.cylinder_re30()
    case.dt = dt
    m = case.mesh
    flow = mini_flow.Channel2D(m.nx, m.ny, m.delta[0], 0.3, 1e-3, dt)   # U_inf = 0.3, nu = 1e-3: Re = U d / nu = 30
    l2c = m.lattice_to_cell()
    c2l = np.argsort(l2c)
    return case, flow, (lambda a: a[c2l]), (lambda a: a[l2c])


def sphere_box_setup(n=32, Re=20.0, nTheta=60):
    """validation/sphere on a coarse periodic box for the mini flow solver: d = 0.004 in a 0.02 box, U_inf = 0.1 along y
    (system/initializeIBMDict), delta = one cell, viscosity from the requested Reynolds number."""
    from ibmfoam_b200.validation import mini_flow
    L, d, Uinf = 0.02, 0.004, 0.1
    h = L / n
    nu = Uinf * d / Re
    case = cases.sphere_case(n=n)
    case.particles.delta[:] = h
    case.particles.nTheta[:] = nTheta
    case.cloud.UInf = (0.0, Uinf, 0.0)
    case.cloud.rhoRef = 1.0
    case.dt = min(0.2 * h / Uinf, 0.1 * h * h / nu)
    flow = mini_flow.PeriodicBox3D(n, h, (0.0, Uinf, 0.0), nu, case.dt)
    return case, flow


def colliding_spheres_setup(n=32, nu=0.05, dt=0.05):
    """two spheres d = 8 cells approaching head-on in a periodic box of fluid at rest (the 3-D counterpart of
    validation/collidingParticles that BASELINE.json's config list describes): forcing + collision + integration loop"""
    from ibmfoam_b200.validation import mini_flow
    case = cases.assembly(n=n, n_particles=2, seed=1, moving=True)
    c = n / 2 + 0.123
    sep = 8.5
    case.particles.position[:] = [[c - sep / 2, c + 0.211, c - 0.317], [c + sep / 2, c + 0.211, c - 0.317]]
    case.particles.v[:] = [[1.0, 0.05, 0.0], [-1.0, 0.0, 0.03]]
    case.particles.omega[:] = [[0.0, 0.1, 0.2], [0.1, 0.0, -0.2]]
    case.cloud.rhoRef = 1.0
    case.dt = dt
    flow = mini_flow.PeriodicBox3D(n, 1.0, (0.0, 0.0, 0.0), nu, dt)
    return case, flow


def kinematics(cloud_like):
    """centre, v, omega of every instance (works for the CUDA cloud and for OracleAsCloud)"""
    import numpy as np
    st = cloud_like.oc.state_arrays() if hasattr(cloud_like, "oc") else cloud_like.state_arrays()
    return np.concatenate([st["center"], st["v"], st["omega"], st["F"], st["T"]], axis=1)
