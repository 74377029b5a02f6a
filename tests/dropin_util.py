"""Helpers of the drop-in demonstration (shim/dropin): case files for the driver, parsing of what it dumps."""
import os
import subprocess

import numpy as np

from ibmfoam_b200 import cases
import ref_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "shim", "dropin", "_build")
REFERENCE_BIN = os.path.join(BUILD, "dropin_reference")
SHIM_BIN = os.path.join(BUILD, "dropin_shim")


def spheres3d():
    """three spheres well inside a 28^3 box (every shape point stays on the mesh): head-on pair + a free one, gravity"""
    n = 28
    mesh = cases.MeshSpec(n, n, n, origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0), patches=ref_cases.WALL_X_CYCLIC_YZ)
    cloud = cases.CloudSpec(moving=True, rotating=True, coupled=True, e=1.0, UInf=(0.0, 0.3, 0.1), g=(0.0, 0.0, -0.2), rhoRef=1.2, pRef=0.1)
    pos = [[14.2, 9.3, 13.6], [14.2, 18.1, 13.6], [8.1, 14.4, 19.3]]
    vel = [[0.0, 1.2, 0.0], [0.0, -1.2, 0.0], [0.3, 0.1, -0.4]]
    om = [[0.1, 0.0, 0.2], [0.0, -0.2, 0.1], [0.2, 0.1, 0.0]]
    parts = cases.ParticleTable.uniform(3, cases.SPHERE, pos, d=8.0, nTheta=50, delta=1.0, rho=1.5, v=vel, omega=om)
    return cases.Case("dropin_spheres3d", mesh, cloud, parts, dt=0.25), 6, 2


def cyl2d_fixed():
    """fixed cylinders on a 2-D lattice, one clipped by a wall, one by a periodic side (PARTIAL instances), nOuter = 3"""
    n = 64
    mesh = cases.MeshSpec(n, n, 1, origin=(0.0, 0.0, -0.5), delta=(1.0, 1.0, 1.0), patches=ref_cases.WALL_X_CYCLIC_Y_2D)
    cloud = cases.CloudSpec(moving=False, rotating=False, coupled=True, e=0.9, UInf=(0.3, 0.0, 0.0))
    pos = [[3.6, 30.2, 0], [40.1, 61.9, 0], [30.4, 9.6, 0], [20.2, 40.3, 0]]
    vel = [[0.2, 0.1, 0], [0.0, -0.3, 0], [0.1, 0.1, 0], [-0.2, 0.0, 0]]
    om = [[0, 0, 0.3], [0, 0, -0.2], [0, 0, 0.1], [0, 0, 0.4]]
    parts = cases.ParticleTable.uniform(4, cases.CYLINDER, pos, d=8.0, nTheta=60, nk=1, delta=1.0, rho=1.5, v=vel, omega=om)
    return cases.Case("dropin_cyl2d_fixed", mesh, cloud, parts, dt=0.25), 3, 3


CASES = {"spheres3d": spheres3d, "cyl2d_fixed": cyl2d_fixed}


def write_case(path, case, n_steps, n_outer, n_restart=0):
    m, c, p = case.mesh, case.cloud, case.particles
    assert not m.blocks
    r = repr  # shortest round-trip representation of a double
    with open(path, "w") as f:
        f.write(f"{m.nx} {m.ny} {m.nz} {' '.join(r(float(x)) for x in m.origin)} {' '.join(r(float(x)) for x in m.delta)} {len(m.patches)}\n")
        for (side, typ, nb) in m.patches:
            f.write(f"{side} {typ} {nb}\n")
        f.write(f"{n_steps} {n_outer} {r(float(case.dt))} {int(c.moving)} {int(c.rotating)} {r(float(c.e))} "
                f"{' '.join(r(float(x)) for x in c.g)} {' '.join(r(float(x)) for x in c.UInf)} {r(float(c.rhoRef))} {r(float(c.pRef))} {p.n}\n")
        for i in range(p.n):
            vals = [p.position[i], [p.d[i], p.a[i], p.b[i]]]
            f.write(f"{int(p.kind[i])} " + " ".join(r(float(x)) for x in p.position[i]) + f" {r(float(p.d[i]))} {r(float(p.a[i]))} {r(float(p.b[i]))} "
                    f"{int(p.nTheta[i])} {r(float(p.delta[i]))} {r(float(p.rho[i]))} " + " ".join(r(float(x)) for x in p.theta[i]) + " " +
                    " ".join(r(float(x)) for x in p.v[i]) + " " + " ".join(r(float(x)) for x in p.omega[i]) + "\n")
        if n_restart:
            # after n_steps the driver writes the time directory, destroys the cloud and constructs a NEW one from it (time != 0:
            # particleCloud::readParticles), then runs n_restart more steps
            f.write(f"{int(n_restart)}\n")


def run(binary, case_path, out_path, timeout=600, no_lattice=False):
    """no_lattice: the driver leaves the optional `lattice` entry out of IBMProperties, so that the shim hands polyMesh's own
    cell / face centres to ibm_lattice_from_arrays (what a stock IBMFoam case provides)"""
    env = dict(os.environ, DROPIN_NO_LATTICE="1") if no_lattice else None
    return subprocess.run([binary, case_path, out_path], capture_output=True, text=True, timeout=timeout, env=env)


def parse(path):
    """-> header (nCells, nSteps, nOuter), list of (tag, array) in file order"""
    with open(path, "rb") as f:
        head = f.readline().split()
        assert head[0] == b"dropin"
        n_cells, n_steps, n_outer = int(head[1]), int(head[2]), int(head[3])
        recs = []
        while True:
            line = f.readline()
            if not line:
                break
            if not line.strip():
                continue
            tag, n = line.split()
            recs.append((tag.decode(), np.frombuffer(f.read(8 * int(n)), dtype=np.float64).copy()))
            f.readline()
    return (n_cells, n_steps, n_outer), recs
