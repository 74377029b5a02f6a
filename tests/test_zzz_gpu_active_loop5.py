"""Quiet / active split of loop 5 (option active_loop5, ibm_motion.cu: k5q_*): instances that cannot interact during a
fluid step integrate all their sub-steps in one kernel, only the others go through the per-sub-step chain.  The results
must be the ones of the unsplit path - bit-identical trajectories against the oracle and against the reference vectors -
also when the displacement budget is exceeded and the step is redone from the snapshot.

Written after the round's GPU budget was spent; verified under the CPU emulation (tools/cuda_emu).  Not a default."""
import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
from oracle.pyoracle import OracleCloud
import golden_util as gu
import ibm_testutil as util

pytestmark = pytest.mark.gpu

KEYS_I = ("index", "copy", "active", "onMesh", "nShell")
KEYS_D = ("center", "theta", "v", "omega", "F", "T")


def _cloud(case, graph, safety_ppm=None, check_ppm=None):
    gc = ParticleCloud(case)
    gc.set_option("fused_loop5", 0)
    gc.set_option("graph_loop5", int(graph is True))
    gc.set_option("active_loop5", 1)
    gc.set_option("fused_active_loop5", int(graph == "fused"))   # the active list's sub-steps in one cooperative kernel
    gc.set_option("cluster_loop5", int(graph == "cluster"))      # independent groups, every sub-step of a group in one block
    if safety_ppm is not None:
        gc.set_option("active_loop5_safety_ppm", int(safety_ppm))
    if check_ppm is not None:
        gc.set_option("active_loop5_check_ppm", int(check_ppm))
    return gc


def _run_vs_oracle(case, n_steps, graph, safety_ppm=None, ft_scale=1e-3, seed=0, check_ppm=None):
    oc, gc = OracleCloud(case), _cloud(case, graph, safety_ppm, check_ppm)
    rng = np.random.default_rng(seed)
    actives = []
    for _ in range(n_steps):
        FT = rng.normal(0, ft_scale, (oc.n_instances, 6))
        if case.mesh.nz == 1:
            FT[:, 2] = 0
            FT[:, 3:5] = 0
        oc.set_force_torque(FT); gc.set_force_torque(FT)
        oc.combine_forces(); gc.combineForces()
        so = oc.advance(case.dt)
        st = gc.advance(case.dt)
        assert st.nSubSteps == int(so[0]) and st.nInstances == int(so[1])
        a, b = oc.state_arrays(), gc.state_arrays()
        for k in KEYS_I + KEYS_D:
            assert np.array_equal(a[k], b[k]), (k, len(actives))
        actives.append(gc.counter("loop5_active_instances"))
    return oc, gc, actives


@pytest.mark.parametrize("graph", ["cluster", True, False, "fused"])
def test_dilute_assembly_most_instances_are_quiet(graph):
    case = cases.assembly(n=64, n_particles=120, seed=9, moving=True, cyclic_without_walls=True)
    oc, gc, actives = _run_vs_oracle(case, 3, graph)
    assert 0 <= max(actives) < gc.nParticles(), actives     # the split did something
    assert gc.counter("loop5_fallbacks") == 0
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, False, "fused"])
def test_fast_assembly_with_periodic_copies_and_deletion(graph):
    case = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    case.particles.v *= 25.0
    case.dt = 0.5
    oc, gc, actives = _run_vs_oracle(case, 4, graph)
    assert oc.n_instances != 60
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, "fused"])
def test_two_spheres_head_on(graph):
    _, gc, _ = _run_vs_oracle(cases.two_spheres(n=40, e=0.8, speed=0.7, gap=0.3), 3, graph)
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, "fused"])
def test_colliding_cylinders_walls_and_copies(graph):
    case = cases.colliding_particles()
    rng = np.random.default_rng(2)
    case.particles.v[:, :2] = rng.normal(0, 0.8, (45, 2))
    case.particles.omega[:, 2] = rng.normal(0, 30.0, 45)
    case.dt = 2.5e-4
    oc, gc, _ = _run_vs_oracle(case, 3, graph, ft_scale=1e-7)
    assert oc.n_instances > 45
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, False, "fused"])
def test_exceeded_budget_redoes_the_step_without_the_split(graph):
    """the check allows a vanishing fraction of the budget: active instances outrun it at once -> snapshot restored, unsplit
    path, same result"""
    case = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    case.particles.v *= 25.0
    case.dt = 0.5
    oc, gc, _ = _run_vs_oracle(case, 3, graph, check_ppm=1)
    assert gc.counter("loop5_fallbacks") > 0
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, "fused"])
@pytest.mark.parametrize("name", ["col8", "wc2d_cyl", "wc3d_sph"])
def test_trajectories_against_reference_vectors_with_the_split(name, graph):
    case, g, n_steps = gu.load_dynamic(name)
    gc = _cloud(case, graph=graph)
    for s in range(1, n_steps + 1):
        gc.set_force_torque(g[f"s{s}_FT_used"])
        gc.advance(case.dt)
        st = gc.state_arrays()
        for k in ("index", "copy", "active", "onMesh", "nShell", "center", "theta", "v", "omega"):
            assert np.array_equal(st[k], g[f"s{s}_{k}"]), (s, k)
    gc.close()


@pytest.mark.parametrize("graph", ["cluster", True, False, "fused"])
def test_copy_created_after_a_rebound_next_to_a_resting_body(graph):
    """A body inside the reach of a periodic patch but moving AWAY from it has no copy; a contact turns it round in mid-step,
    the copy appears on the far side and at once overlaps a body that rests there with nothing near it at the start of the
    step: that body must not have been classified quiet (the images entered by k5q_ghosts see to it)."""
    n = 48
    pos = np.array([[2.0, 24.3, 24.1],                # g: 2 cells from the x-min face (r = 4), moving away from it
                    [2.0 + 8.05, 24.3, 24.1],         # h: hits g and hands it its velocity towards the face
                    [n - 5.7, 24.3, 24.1],            # q: 5.7 cells from the x-max face, 7.7 from where g's copy will appear
                    [24.0, 10.0, 37.0]])              # a bystander
    case = cases.assembly(n=n, n_particles=4, seed=3, moving=True, cyclic_without_walls=True, positions=pos)
    case.particles.v[:] = [[0.5, 0, 0], [-1.0, 0, 0], [0, 0, 0], [0.01, 0.02, 0.0]]
    case.particles.omega[:] = 0.0
    oc, gc, actives = _run_vs_oracle(case, 2, graph, ft_scale=1e-9)
    st = gc.state_arrays()
    q = int(np.where((st["index"] == 2) & (st["copy"] == 0))[0][0])
    assert st["v"][q, 0] < -0.5                   # q was struck by the copy
    assert oc.n_instances > 4
    assert gc.counter("loop5_fallbacks") == 0
    gc.close()


@pytest.mark.parametrize("mode", ["split", "split_fused"])
def test_two_ranks_replicated_loop5_with_the_split(mode):
    """the multi-rank design runs loop 5 redundantly on every rank: with the split (and the persistent kernel) every rank
    must still reproduce the SERIAL oracle bit for bit - the classification and the active list may be ordered differently
    on each rank, the trajectories may not differ (ranks = host threads on one device, in-process transport)"""
    import threading
    from ibmfoam_b200.cloud import LocalGroup, slab_ranges
    n_ranks = 2
    case = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    case.particles.v *= 8.0
    ranges = slab_ranges(case.mesh.nz, n_ranks)
    group = LocalGroup(n_ranks)
    oc = OracleCloud(case)
    rng = np.random.default_rng(0)
    FTs = [rng.normal(0, 1e-3, (400, 6)) for _ in range(3)]
    ref = []
    for s in range(3):
        oc.set_force_torque(FTs[s][:oc.n_instances])
        oc.combine_forces()
        oc.advance(case.dt)
        ref.append(oc.state_arrays())
    out, err = [None] * n_ranks, [None] * n_ranks

    def rank(r):
        try:
            gc = ParticleCloud(case, slab=ranges[r], local_group=group.handle, rank=r)
            gc.set_option("active_loop5", 1)
            gc.set_option("fused_active_loop5", int(mode == "split_fused"))
            for s in range(3):
                gc.set_force_torque(FTs[s][:gc.nParticles()] / n_ranks)   # the all-reduce restores the serial value exactly
                gc.combineForces()
                gc.advance(case.dt)
                sg = gc.state_arrays()
                for k in ("index", "copy", "active", "onMesh", "center", "theta", "v", "omega"):
                    assert np.array_equal(sg[k], ref[s][k]), (r, s, k)
            out[r] = (gc.nParticles(), gc.counter("loop5_active_instances"))
            gc.close()
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=rank, args=(r,)) for r in range(n_ranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    for e in err:
        if e is not None:
            raise e
    assert out[0][0] == out[1][0] == oc.n_instances
    assert 0 <= out[0][1] < out[0][0]          # the split was in force on the last step
    group.close()
