"""The C++ host-side mirror of the reference's particleCloud / particleIBM interface (ibmfoam_b200/host/) over the
C ABI: builds on CPU; its driver runs the solver's call sequence on the GPU."""
import os
import subprocess

import pytest

from ibmfoam_b200 import build

EXE = os.path.join(os.path.dirname(build.LIB), "test_host_mirror")


def test_host_mirror_builds():
    build.build()
    assert os.path.exists(EXE)
    hdr = open(os.path.join(os.path.dirname(build.HERE), "ibmfoam_b200", "host", "particleCloud.H")).read()
    # the solver-facing names of src/particleCloud/particleCloud.H:267-358
    for name in ("class particleCloud", "forcing(", "integrateSurfaceStresses(", "solve(", "maxCoNum()", "nParticles()",
                 "Cd(", "UInf()", "rhoRef()", "pRef()", "ibmDict()", "mesh()"):
        assert name in hdr, name


@pytest.mark.gpu
def test_host_mirror_runs_solver_sequence():
    build.build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout
