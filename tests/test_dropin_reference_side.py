"""Drop-in demonstration, CPU part: the solver-side driver of shim/dropin compiled against the REFERENCE'S OWN
particleCloud runs and produces the quantities the GPU test compares (tests/test_zz_gpu_shim_dropin.py); the same
driver compiled against the shim refuses to run without a CUDA device (no CPU fallback behind the shim)."""
import os

import numpy as np
import pytest

import dropin_util as du


@pytest.mark.skipif(not os.path.exists(du.REFERENCE_BIN), reason="shim/dropin/_build/dropin_reference not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(du.CASES))
def test_reference_side_runs(tmp_path, name):
    case, n_steps, n_outer = du.CASES[name]()
    cp, op = str(tmp_path / "case.txt"), str(tmp_path / "out.bin")
    du.write_case(cp, case, n_steps, n_outer)
    r = du.run(du.REFERENCE_BIN, cp, op)
    assert r.returncode == 0, r.stderr
    (n_cells, ns, no), recs = du.parse(op)
    assert (n_cells, ns, no) == (case.mesh.n_cells, n_steps, n_outer)
    tags = [t for t, _ in recs]
    assert tags.count("F") == n_steps * n_outer and tags.count("Cd") == n_steps
    F = [a for t, a in recs if t == "F"]
    assert all(np.count_nonzero(a) > 0 for a in F)
    d = dict(recs)
    assert d["CoNum"][0] == 0.0                                    # App. B #9
    assert len(d["index"]) == case.particles.n and len(d["center"]) == 3 * case.particles.n
    if case.cloud.moving:
        assert not np.allclose(d["center"].reshape(-1, 3), case.particles.position)   # the particles moved
        v = d["v"].reshape(-1, 3)
        assert v[0, 1] < 0 < v[1, 1], "the head-on pair must have collided and reversed"
    else:
        assert np.array_equal(d["center"].reshape(-1, 3), case.particles.position)


@pytest.mark.skipif(not os.path.exists(du.SHIM_BIN), reason="shim/dropin/_build/dropin_shim not built")
def test_shim_side_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the shim side is exercised by tests/test_zz_gpu_shim_dropin.py")
    case, n_steps, n_outer = du.CASES["cyl2d_fixed"]()
    cp, op = str(tmp_path / "case.txt"), str(tmp_path / "out.bin")
    du.write_case(cp, case, n_steps, n_outer)
    r = du.run(du.SHIM_BIN, cp, op)
    assert r.returncode != 0 and "CUDA" in r.stderr


@pytest.mark.skipif(not os.path.exists(du.REFERENCE_BIN), reason="shim/dropin/_build/dropin_reference not built (needs /root/reference)")
def test_reference_side_restarts_from_the_written_time_directory(tmp_path):
    """the reference's own restart path (particleCloud.C:1119-1123 -> particleCloudIO.C:47-69 -> particleIBMIO.C:111-213): after
    n steps the driver writes lagrangian/IBMparticles/*, destroys the cloud and constructs a new one at time != 0"""
    case, n_steps, n_outer = du.CASES["spheres3d"]()
    cp, op = str(tmp_path / "case.txt"), str(tmp_path / "out.bin")
    du.write_case(cp, case, 3, 1, n_restart=3)
    r = du.run(du.REFERENCE_BIN, cp, op)
    assert r.returncode == 0, r.stderr
    _, recs = du.parse(op)
    tags = [t for t, _ in recs]
    assert tags.count("restarted") == 1 and tags.count("F") == 6
    d = dict(recs)
    assert d["restarted"][0] == case.particles.n          # every instance came back from the time directory
    assert len(d["index"]) == case.particles.n
