"""The CPU restatement (oracle/orc_*.hpp) against the committed golden vectors that the REFERENCE'S OWN CODE produced
(tests/golden/*.npz, generator tests/golden/make_golden.py).  Bit for bit.  This runs everywhere - also where neither
/root/reference nor oracle/_ref exists - and is what pins the oracle used by the GPU parity tests."""
import numpy as np
import pytest

from oracle.pyoracle import OracleCloud
import golden_util as gu
import ibm_testutil as util
import ref_cases


@pytest.mark.parametrize("name", sorted(ref_cases.STATIC_CASES))
def test_restatement_matches_reference_vectors_static(name):
    case, g = gu.load_static(name)
    oc = OracleCloud(case)
    assert oc.n_instances == int(g["n_instances"])
    st = oc.state_arrays()
    for k in gu.STATE_I:
        assert np.array_equal(st[k], g["state_" + k]), k
    for k in ("center", "theta", "v", "omega"):
        assert np.array_equal(st[k], g["state_" + k]), k
    for i in range(oc.n_instances):
        sh, pt = oc.shell(i), oc.points(i)
        assert np.array_equal(sh["cells"], g[f"shell{i}_cells"])
        assert np.array_equal(sh["jk"], g[f"shell{i}_jk"])
        assert np.array_equal(sh["W"], g[f"shell{i}_W"])
        assert np.array_equal(pt["cell"], g[f"points{i}_cell"])
        assert np.array_equal(pt["nFaces"], g[f"points{i}_nFaces"])
        assert np.array_equal(pt["W"], g[f"points{i}_W"])
    assert np.array_equal(oc.indicator(), gu.indicator(g, case.mesh.n_cells))
    f = util.caller_fields(case)
    F, n_upd = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    assert n_upd == int(g["forcing_updates"])
    assert np.array_equal(F, gu.dense_forcing(g, case.mesh.n_cells))
    assert np.array_equal(oc.integrate_stress(f["p"], f["tau"]), g["stress_FT"])
    oc.combine_forces()
    st = oc.state_arrays()
    assert np.array_equal(np.concatenate([st["F"], st["T"]], axis=1), g["combined_FT"])
    assert np.array_equal(oc.cd(), g["cd"], equal_nan=True)
    assert np.array_equal([oc.mass(i) for i in range(oc.n_instances)], g["mass"])
    assert np.array_equal([oc.inertia(i) for i in range(oc.n_instances)], g["inertia"])


@pytest.mark.parametrize("name", sorted(ref_cases.DYNAMIC_CASES))
def test_restatement_matches_reference_vectors_dynamic(name):
    """solve() step after step with quirk B#8 as coded (like the reference): states, F/T and statistics identical."""
    case, g, n_steps = gu.load_dynamic(name)
    oc = OracleCloud(case, stale_tables=True)
    f = util.caller_fields(case)
    for s in range(1, n_steps + 1):
        stats = oc.solve(f["p"], f["tau"], case.dt)
        st = oc.state_arrays()
        for k in gu.STATE_I + gu.STATE_D:
            assert np.array_equal(st[k], g[f"s{s}_{k}"]), (s, k)
        assert np.array_equal(stats[1:], g[f"s{s}_stats"]), s
    for i in range(oc.n_instances):
        sh = oc.shell(i)
        assert np.array_equal(sh["cells"], g[f"final_shell{i}_cells"]) and np.array_equal(sh["jk"], g[f"final_shell{i}_jk"])


@pytest.mark.parametrize("name", sorted(ref_cases.DYNAMIC_CASES))
def test_defined_policy_replays_reference_kinematics(name):
    """The DEFINED B#8 policy (what the CUDA path implements) changes the stress integral of partly-off-mesh copies
    only; fed the reference's F/T, the sub-stepped motion, the copies and the re-marking are the reference's."""
    case, g, n_steps = gu.load_dynamic(name)
    oc = OracleCloud(case)
    for s in range(1, n_steps + 1):
        oc.set_force_torque(g[f"s{s}_FT_used"])
        oc.advance(case.dt)
        st = oc.state_arrays()
        for k in gu.STATE_I + ("center", "theta", "v", "omega"):
            assert np.array_equal(st[k], g[f"s{s}_{k}"]), (s, k)
