"""The CUDA path (through the C ABI) against the committed golden vectors that the REFERENCE'S OWN CODE produced
(tests/golden/*.npz; generator tests/golden/make_golden.py runs /root/reference/src compiled against oracle/mockfoam).
Nothing here touches /root/reference or oracle/: the vectors are the checker.

Bars (BASELINE.json north star): marking and particle-cell membership bit-exact; forcing field 1e-12 relative;
force/torque 1e-10 relative; sphere/cylinder trajectories, fed the reference's F/T, bit-identical (only IEEE
+,-,*,/,sqrt on that path)."""
import numpy as np
import pytest

from ibmfoam_b200.cloud import ParticleCloud
import golden_util as gu
import ibm_testutil as util
import ref_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(ref_cases.STATIC_CASES))
def test_marking_forcing_stress_against_reference_vectors(name):
    case, g = gu.load_static(name)
    gc = ParticleCloud(case)
    assert gc.nParticles() == int(g["n_instances"])
    st = gc.state_arrays()
    for k in gu.STATE_I:
        assert np.array_equal(st[k], g["state_" + k]), k
    for i in range(gc.nParticles()):
        sh, pt = gc.shell(i), gc.points(i)
        assert np.array_equal(sh["cells"], g[f"shell{i}_cells"]), (i, "shell cells")
        assert np.array_equal(sh["jk"], g[f"shell{i}_jk"]), (i, "polar bins")
        assert np.array_equal(pt["cell"], g[f"points{i}_cell"]), (i, "point -> cell membership")
        assert np.array_equal(pt["nFaces"], g[f"points{i}_nFaces"]), (i, "gather face counts")
        assert np.allclose(sh["W"], g[f"shell{i}_W"], rtol=1e-14, atol=0)
        assert np.allclose(pt["W"], g[f"points{i}_W"], rtol=1e-14, atol=0)
    assert np.array_equal(gc.indicator(), gu.indicator(g, case.mesh.n_cells))
    f = util.caller_fields(case)
    Fref = gu.dense_forcing(g, case.mesh.n_cells)
    for variant in (0, 2, 7):
        gc.set_kernel_variant(forcing=variant, stress=variant)
        F = gc.forcing(f["S"], f["U"], f["Uold"], case.dt)
        err = util.rel_err(F, Fref)
        assert err < (1e-14 if variant == 0 else 1e-12), (variant, err)          # tolerance: 1e-12 relative (fp64)
        FT = gc.integrateSurfaceStresses(f["p"], f["tau"])
        ref = g["stress_FT"]
        for sl in (slice(0, 3), slice(3, 6)):
            scale = np.max(np.linalg.norm(ref[:, sl], axis=1))
            e2 = float(np.max(np.abs(FT[:, sl] - ref[:, sl])) / scale)
            assert e2 < (1e-13 if variant == 0 else 1e-10), (variant, e2)        # tolerance: 1e-10 relative
    assert gc.timings()["n_updates"] == 3 * int(g["forcing_updates"])
    gc.combineForces()
    st = gc.state_arrays()
    ref = g["combined_FT"]
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(np.concatenate([st["F"], st["T"]], axis=1) - ref)) < 1e-10 * scale
    cd, cd_ref = gc.Cd(), g["cd"]
    ok = np.isfinite(cd_ref)
    assert np.max(np.abs(cd[ok] - cd_ref[ok])) < 1e-9 * np.max(np.abs(cd_ref[ok]))
    for i in range(gc.nParticles()):
        m, inertia = gc.mass_inertia(i)
        assert m == g["mass"][i] and inertia == g["inertia"][i]
    gc.close()


@pytest.mark.parametrize("mode", ["cluster", "graph", "plain"])
@pytest.mark.parametrize("name", ["col8", "wc2d_cyl", "wc3d_sph"])
def test_trajectories_against_reference_vectors(name, mode):
    """Each fluid step: the F/T the reference used for that step go in, ibm_advance runs the literal 100/101 sub-step loop
    (collisions, wall hits, periodic copies, integration, setPositions) and the re-marking; instance list, status and
    kinematics must equal the reference's bit for bit."""
    case, g, n_steps = gu.load_dynamic(name)
    gc = ParticleCloud(case)
    gc.set_option("active_loop5", int(mode == "cluster"))
    gc.set_option("fused_loop5", 0)
    gc.set_option("graph_loop5", int(mode == "graph"))
    for s in range(1, n_steps + 1):
        FT = g[f"s{s}_FT_used"]
        assert gc.nParticles() == FT.shape[0], s
        gc.set_force_torque(FT)
        stats = gc.advance(case.dt)
        st = gc.state_arrays()
        assert gc.nParticles() == g[f"s{s}_index"].shape[0], s
        for k in ("index", "copy", "active", "onMesh", "nShell"):
            assert np.array_equal(st[k], g[f"s{s}_{k}"]), (s, k)
        for k in ("center", "theta", "v", "omega"):
            assert np.array_equal(st[k], g[f"s{s}_{k}"]), (s, k, np.max(np.abs(st[k] - g[f"s{s}_{k}"])))
        ref = g[f"s{s}_stats"]  # nInstances, avgV, avgOmega, avgP, totK  (particleCloud.C:1288-1322)
        got = np.concatenate([[stats.nInstances], list(stats.avgV), list(stats.avgOmega), list(stats.avgP), [stats.totK]])
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-300), s
    for i in range(gc.nParticles()):
        sh = gc.shell(i)
        assert np.array_equal(sh["cells"], g[f"final_shell{i}_cells"]) and np.array_equal(sh["jk"], g[f"final_shell{i}_jk"])
    gc.close()
