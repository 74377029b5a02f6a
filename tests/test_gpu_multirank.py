"""Multi-rank path (slab-decomposed fields, replicated particle state) on ONE GPU: ranks are host threads joined by
the in-process transport; the checker is one oracle per slab (the decomposed reference semantics: shape points
outside the local mesh contribute nothing, particleShape.C:45-47) plus the serial oracle for loop 5."""
import threading

import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import LocalGroup, ParticleCloud, slab_ranges
from oracle.pyoracle import OracleCloud, OracleMesh
import ibm_testutil as util

pytestmark = pytest.mark.gpu


def _run_ranks(n_ranks, fn):
    out, err = [None] * n_ranks, [None] * n_ranks

    def work(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(n_ranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    for e in err:
        if e is not None:
            raise e
    return out


def _slab_case(case, k0, k1):
    import copy
    c = copy.copy(case)
    c.mesh = case.mesh.slab(k0, k1)
    return c


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_loops_1_to_4_match_decomposed_oracle(n_ranks):
    case = cases.assembly(n=48, n_particles=70, seed=11, moving=True, cyclic_without_walls=True)
    ranges = slab_ranges(case.mesh.nz, n_ranks)
    f_lat = cases.synthetic_fields(case.mesh)
    nxy = case.mesh.nx * case.mesh.ny
    group = LocalGroup(n_ranks)
    oracles = [OracleCloud(_slab_case(case, *ranges[r])) for r in range(n_ranks)]
    FTo = []
    for r, oc in enumerate(oracles):
        k0, k1 = ranges[r]
        sl = slice(k0 * nxy, k1 * nxy)
        FTo.append(oc.integrate_stress(f_lat["p"][sl], f_lat["tau"][sl]))

    def rank(r):
        k0, k1 = ranges[r]
        sl = slice(k0 * nxy, k1 * nxy)
        gc = ParticleCloud(case, slab=(k0, k1), local_group=group.handle, rank=r)
        oc = oracles[r]
        sg, so = gc.state_arrays(), oc.state_arrays()
        assert gc.nParticles() == case.particles.n  # replicated: every rank holds every instance
        pos = {int(i): n for n, i in enumerate(sg["index"])}
        # marking: every instance the decomposed reference keeps on this rank has the same shell here
        for n_o, idx in enumerate(so["index"]):
            g = pos[int(idx)]
            a, b = oc.shell(n_o), gc.shell(g)
            assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["jk"], b["jk"])
        absent = set(pos) - set(int(i) for i in so["index"])
        assert all(sg["nShell"][pos[i]] == 0 for i in absent)
        # forcing on the local planes
        Fo, _ = oc.forcing(f_lat["U"][sl], f_lat["Uold"][sl], f_lat["S"][sl], case.dt)
        Fg = gc.forcing(f_lat["S"][sl], f_lat["U"][sl], f_lat["Uold"][sl], case.dt)
        assert util.rel_err(Fg, Fo) < 1e-12
        # stress partials, then the cross-rank sum
        FTg = gc.integrateSurfaceStresses(f_lat["p"][sl], f_lat["tau"][sl])
        scale = max(np.max(np.abs(x)) for x in FTo)
        for n_o, idx in enumerate(so["index"]):
            assert np.max(np.abs(FTg[pos[int(idx)]] - FTo[r][n_o])) < 1e-10 * scale
        for i in absent:
            assert not np.any(FTg[pos[i]])
        gc.combineForces()
        tot = np.zeros((case.particles.n, 6))
        for rr, o2 in enumerate(oracles):
            for n_o, idx in enumerate(o2.state_arrays()["index"]):
                tot[pos[int(idx)]] += FTo[rr][n_o]
        st = gc.state_arrays()
        got = np.concatenate([st["F"], st["T"]], axis=1)
        assert np.max(np.abs(got - tot)) < 1e-10 * scale
        # a second stress call (after the combined totals were written back): instances without local points must
        # contribute ZERO partials again, not last step's totals (they would be multiplied by the rank count)
        FTg2 = gc.integrateSurfaceStresses(f_lat["p"][sl], f_lat["tau"][sl])
        assert np.array_equal(FTg2, FTg)
        gc.combineForces()
        st2 = gc.state_arrays()
        assert np.array_equal(st2["F"], st["F"]) and np.array_equal(st2["T"], st["T"])
        res = (st["F"].copy(), st["T"].copy())
        gc.close()
        return res

    res = _run_ranks(n_ranks, rank)
    for r in range(1, n_ranks):  # identical bits on every rank (rank-ordered sum)
        assert np.array_equal(res[0][0], res[r][0]) and np.array_equal(res[0][1], res[r][1])
    group.close()


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_halo_makes_results_decomposition_independent(n_ranks):
    """EXTENSION: with a few read-only halo planes per slab every gather stencil is complete, so the shell
    lists of the ranks tile the serial ones, the forcing field on the owned planes is the SERIAL field, and the summed
    force / torque the serial force / torque - whatever the decomposition (the reference's depend on it, particleShape.C:45-47)."""
    case = cases.assembly(n=48, n_particles=70, seed=11, moving=True, cyclic_without_walls=True)
    ranges = slab_ranges(case.mesh.nz, n_ranks)
    f_lat = cases.synthetic_fields(case.mesh)
    nxy = case.mesh.nx * case.mesh.ny
    group = LocalGroup(n_ranks)
    oc = OracleCloud(case)  # the WHOLE mesh
    Fo, _ = oc.forcing(f_lat["U"], f_lat["Uold"], f_lat["S"], case.dt)
    FTo = oc.integrate_stress(f_lat["p"], f_lat["tau"])
    so = oc.state_arrays()
    shells = [oc.shell(n) for n in range(len(so["index"]))]
    # d = 8 cells: the degenerate polar bins of a sphere reference point 0 at the opposite pole (App. B #3), so the halo has to
    # span the body for bit-level agreement; 4 planes leave two cells of this case with an incomplete stencil
    halo = 10

    def rank(r):
        k0, k1 = ranges[r]
        gc = ParticleCloud(case, slab=(k0, k1), halo=halo, local_group=group.handle, rank=r)
        lo, hi = gc.local_planes
        assert lo == max(0, k0 - halo) and hi == min(case.mesh.nz, k1 + halo)
        sl = slice(lo * nxy, hi * nxy)
        sg = gc.state_arrays()
        assert np.array_equal(sg["index"], so["index"]) and np.array_equal(sg["onMesh"], so["onMesh"])
        # marking: the serial shell of every instance restricted to the owned planes (cell ids are local: + lo planes)
        n_cells_owned = 0
        for n in range(len(so["index"])):
            a, b = shells[n], gc.shell(n)
            keep = (a["cells"] // nxy >= k0) & (a["cells"] // nxy < k1)
            assert np.array_equal(a["cells"][keep] - lo * nxy, b["cells"]) and np.array_equal(a["jk"][keep], b["jk"])
            n_cells_owned += int(keep.sum())
        # forcing: the serial field on the owned planes, nothing in the halo planes
        Fg = gc.forcing(f_lat["S"][sl], f_lat["U"][sl], f_lat["Uold"][sl], case.dt)
        own = slice((k0 - lo) * nxy, (k1 - lo) * nxy)
        assert util.rel_err(Fg[own], Fo[k0 * nxy:k1 * nxy]) < 1e-12
        assert not np.any(Fg[:(k0 - lo) * nxy]) and not np.any(Fg[(k1 - lo) * nxy:])
        gc.integrateSurfaceStresses(f_lat["p"][sl], f_lat["tau"][sl])
        gc.combineForces()
        st = gc.state_arrays()
        got = np.concatenate([st["F"], st["T"]], axis=1)
        gc.close()
        return got, n_cells_owned

    res = _run_ranks(n_ranks, rank)
    scale = np.max(np.abs(FTo))
    for got, _ in res:
        assert np.max(np.abs(got - FTo)) < 1e-10 * scale  # the SERIAL force and torque of every instance
    assert sum(n for _, n in res) == sum(len(s["cells"]) for s in shells)
    group.close()


@pytest.mark.parametrize("n_ranks,owned", [(2, 1), (3, 1), (2, 0)])
def test_loop5_replicated_equals_serial_oracle(n_ranks, owned):
    """Loop 5 over several ranks: trajectories, periodic copies and deletions are those of the SERIAL reference
    (decomposition-independent), bit for bit, on every rank.  owned = 1 (default): every rank advances the instances and
    groups its slab owns and the states / new copies are exchanged once per fluid step; owned = 0: every rank advances
    everything redundantly from identical state."""
    case = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    case.particles.v *= 25.0
    case.dt = 0.5
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

    g1 = ParticleCloud(case)  # one GPU, for the number of fallbacks of the split on this case
    for s in range(3):
        g1.set_force_torque(FTs[s][:g1.nParticles()])
        g1.combineForces()
        g1.advance(case.dt)
    fallbacks_serial = g1.counter("loop5_fallbacks")
    g1.close()

    def rank(r):
        gc = ParticleCloud(case, slab=ranges[r], local_group=group.handle, rank=r)
        gc.set_option("owned_loop5", owned)
        for s in range(3):
            # hand each rank 1/n of the total so that the all-reduce restores the serial value exactly (powers of two)
            gc.set_force_torque(FTs[s][:gc.nParticles()] / n_ranks)
            gc.combineForces()
            gc.advance(case.dt)
            sg = gc.state_arrays()
            for k in ("index", "copy", "active", "onMesh"):
                assert np.array_equal(sg[k], ref[s][k]), k
            for k in ("center", "theta", "v", "omega"):
                assert np.array_equal(sg[k], ref[s][k]), k
        n = gc.nParticles()
        assert gc.counter("loop5_cluster_steps") > 0
        assert (gc.counter("loop5_owned_steps") > 0) == bool(owned)
        assert gc.counter("loop5_fallbacks") == fallbacks_serial, "the ranks must not fall back where one GPU does not"
        gc.close()
        return n

    ns = _run_ranks(n_ranks, rank)
    assert all(n == oc.n_instances for n in ns) and ns[0] != 60
    group.close()
