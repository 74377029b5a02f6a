"""Small cases used to pin the CPU restatement (oracle/orc_*.hpp) against the REFERENCE'S OWN CODE compiled against
oracle/mockfoam (oracle/_ref/libibm_ref.so), and to generate / check the committed golden vectors (tests/golden/).

Every case keeps particle centres in generic position: a centre aligned with a cell-centre column puts a cell exactly
at phi = pi, where the reference's bin index runs off the end of its table (undefined behaviour;
see test_oracle_vs_reference.py::test_reference_runs_off_its_table_for_lattice_aligned_spheres).
Meshes are small because the reference allocates a labelList(nFaces) per shape point per update (particleShape.C:54).
"""
import numpy as np

from ibmfoam_b200 import cases

WALL_X_CYCLIC_YZ = [(cases.XMIN, cases.WALL, -1), (cases.XMAX, cases.WALL, -1), (cases.YMIN, cases.CYCLIC, 3),
                    (cases.YMAX, cases.CYCLIC, 2), (cases.ZMIN, cases.CYCLIC, 5), (cases.ZMAX, cases.CYCLIC, 4)]
WALL_X_CYCLIC_Y_2D = [(cases.XMIN, cases.WALL, -1), (cases.XMAX, cases.WALL, -1), (cases.YMIN, cases.CYCLIC, 3),
                      (cases.YMAX, cases.CYCLIC, 2), (cases.ZMIN, cases.EMPTY, -1), (cases.ZMAX, cases.EMPTY, -1)]


def assembly_static(n=24, n_particles=3, seed=7):
    """Periodic box of spheres d = 8, delta = 1, nTheta = 50 with non-zero v, omega (they enter the forcing)."""
    case = cases.assembly(n=n, n_particles=n_particles, seed=seed, moving=False)
    case.particles.v[:] = np.random.default_rng(1).normal(0, 0.2, case.particles.v.shape)
    case.particles.omega[:] = np.random.default_rng(2).normal(0, 0.1, case.particles.omega.shape)
    case.cloud.UInf = (0.3, 0.1, 0.0)
    case.name = f"asm{n}_static"
    return case


def cylinder_four_blocks():
    """cylinderRe30's 4-block numbering on a coarse lattice (60 x 30), one fixed cylinder, d = 15 cells."""
    nx, ny = 60, 30
    mesh = cases.MeshSpec(nx, ny, 1, origin=(-0.5, -0.5, -0.01), delta=(2.0 / nx, 1.0 / ny, 0.02),
                          blocks=cases._four_blocks(15, 45, 15),
                          patches=[(cases.XMIN, cases.PATCH, -1), (cases.XMAX, cases.PATCH, -1), (cases.YMIN, cases.CYCLIC, 3),
                                   (cases.YMAX, cases.CYCLIC, 2), (cases.ZMIN, cases.EMPTY, -1), (cases.ZMAX, cases.EMPTY, -1)])
    cloud = cases.CloudSpec(moving=False, rotating=False, coupled=True, e=1.0, UInf=(0.3, 0, 0), rhoRef=1.0)
    parts = cases.ParticleTable.uniform(1, cases.CYLINDER, [[0.0031, -0.0017, 0]], d=0.5, nTheta=100, nk=1, delta=0.05, rho=1000,
                                        v=[[0.02, -0.01, 0]], omega=[[0, 0, 0.3]])
    return cases.Case("cyl4b", mesh, cloud, parts, dt=1e-3)


def wall_cyclic_2d(kind=cases.CYLINDER):
    """64 x 64 2-D box, x walls, y cyclic, gravity: wall hits, pair collisions, periodic copies, deletions."""
    n = 64
    mesh = cases.MeshSpec(n, n, 1, origin=(0.0, 0.0, -0.5), delta=(1.0, 1.0, 1.0), patches=WALL_X_CYCLIC_Y_2D)
    cloud = cases.CloudSpec(moving=True, rotating=True, coupled=True, e=0.9, UInf=(0.3, 0, 0), g=(0.0, -0.4, 0.0))
    pos = [[6.3, 30.2, 0], [40.1, 58.7, 0], [30.4, 9.6, 0], [20.2, 40.3, 0], [29.1, 41.1, 0], [50.3, 20.7, 0]]
    vel = [[-3.0, 0.5, 0], [0.2, 4.0, 0], [0.1, -3.5, 0], [2.0, 0.1, 0], [-2.0, -0.1, 0], [3.5, 0.3, 0]]
    om = [[0, 0, 0.3], [0, 0, -0.2], [0, 0, 0.1], [0, 0, 0.0], [0, 0, 0.5], [0, 0, -0.4]]
    if kind == cases.CYLINDER:
        parts = cases.ParticleTable.uniform(6, cases.CYLINDER, pos, d=8.0, nTheta=60, nk=1, delta=1.0, rho=1.5, v=vel, omega=om)
        name = "wc2d_cyl"
    else:
        parts = cases.ParticleTable.uniform(6, cases.ELLIPSE, pos, a=4.0, b=2.5, nTheta=60, nk=1, delta=1.0, rho=1.5, v=vel,
                                            omega=om, theta=[[0, 0, 0.3 * i] for i in range(6)])
        name = "wc2d_ell"
    return cases.Case(name, mesh, cloud, parts, dt=0.25)


def wall_cyclic_3d():
    """28^3 box, x walls, y/z cyclic: spheres d = 8 hitting a wall, each other, and crossing the periodic sides."""
    n = 28
    mesh = cases.MeshSpec(n, n, n, origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0), patches=WALL_X_CYCLIC_YZ)
    cloud = cases.CloudSpec(moving=True, rotating=True, coupled=True, e=1.0, UInf=(0.0, 0.2, 0.0), g=(0.0, 0.0, -0.3))
    pos = [[5.2, 13.3, 14.1], [14.3, 22.6, 13.7], [19.7, 8.4, 6.3], [19.1, 9.1, 15.2]]
    vel = [[-2.5, 0.3, 0.1], [0.2, 3.0, -0.4], [0.3, -0.2, 1.5], [0.1, 0.2, -1.8]]
    om = [[0.1, 0.0, 0.3], [0.0, -0.2, 0.1], [0.2, 0.1, 0.0], [-0.1, 0.3, 0.2]]
    parts = cases.ParticleTable.uniform(4, cases.SPHERE, pos, d=8.0, nTheta=50, delta=1.0, rho=1.5, v=vel, omega=om)
    return cases.Case("wc3d_sph", mesh, cloud, parts, dt=0.25)


STATIC_CASES = {
    "asm24_static": assembly_static,
    "cyl4b": cylinder_four_blocks,
    "col6": lambda: cases.colliding_particles(n_particles=6),
    "ellipse2": None,  # filled below (needs ibm_testutil)
}
DYNAMIC_CASES = {
    # name: (builder, fluid steps)
    "wc2d_cyl": (lambda: wall_cyclic_2d(cases.CYLINDER), 16),
    "wc2d_ell": (lambda: wall_cyclic_2d(cases.ELLIPSE), 12),
    "wc3d_sph": (wall_cyclic_3d, 8),
    "col8": (lambda: cases.colliding_particles(n_particles=8), 3),
}


def _ellipse2():
    import ibm_testutil as util
    return util.ellipse_case()


STATIC_CASES["ellipse2"] = _ellipse2
