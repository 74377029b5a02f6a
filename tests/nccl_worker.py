"""Worker of tests/test_gpu_nccl_two_ranks.py: one process per GPU (torch.distributed.run), NCCL transport of the library.
Each rank checks its slab against the decomposed oracle (loops 1-4), the combined force / torque against the sum of the
per-slab oracles, and three fluid steps of loop 5 against the SERIAL oracle, bit for bit.  Prints NCCL_WORKER_OK."""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage(msg):
    print(f"[nccl_worker rank {os.environ.get('RANK')}] {msg}", flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)  # a rank stuck in a collective shows where, and the test ends
    import torch
    import torch.distributed as dist
    from ibmfoam_b200 import cases
    from ibmfoam_b200.cloud import ParticleCloud, nccl_unique_id, slab_ranges
    from oracle.pyoracle import OracleCloud
    import ibm_testutil as util

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().tolist())
    stage("process group up")

    # ---- loops 1-4 on slabs + combineForces over NCCL ----
    case = cases.assembly(n=48, n_particles=70, seed=11, moving=True, cyclic_without_walls=True)
    ranges = slab_ranges(case.mesh.nz, world)
    f_lat = cases.synthetic_fields(case.mesh)
    nxy = case.mesh.nx * case.mesh.ny
    oracles, FTo = [], []
    for r in range(world):
        c2 = copy.copy(case)
        c2.mesh = case.mesh.slab(*ranges[r])
        oc = OracleCloud(c2)
        sl = slice(ranges[r][0] * nxy, ranges[r][1] * nxy)
        oracles.append(oc)
        FTo.append(oc.integrate_stress(f_lat["p"][sl], f_lat["tau"][sl]))
    k0, k1 = ranges[rank]
    sl = slice(k0 * nxy, k1 * nxy)
    gc = ParticleCloud(case, device=local, slab=(k0, k1), nccl_id=nccl_id, rank=rank, n_ranks=world)
    stage("cloud 1 created")
    oc = oracles[rank]
    sg, so = gc.state_arrays(), oc.state_arrays()
    pos = {int(i): n for n, i in enumerate(sg["index"])}
    for n_o, idx in enumerate(so["index"]):
        a, b = oc.shell(n_o), gc.shell(pos[int(idx)])
        assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["jk"], b["jk"])
    Fo, _ = oc.forcing(f_lat["U"][sl], f_lat["Uold"][sl], f_lat["S"][sl], case.dt)
    Fg = gc.forcing(f_lat["S"][sl], f_lat["U"][sl], f_lat["Uold"][sl], case.dt)
    assert util.rel_err(Fg, Fo) < 1e-12
    gc.integrateSurfaceStresses(f_lat["p"][sl], f_lat["tau"][sl])
    stage("loops 1-4 done")
    gc.combineForces()  # ncclAllReduce of the partials + the same-index sum
    tot = np.zeros((case.particles.n, 6))
    for rr, o2 in enumerate(oracles):
        for n_o, idx in enumerate(o2.state_arrays()["index"]):
            tot[pos[int(idx)]] += FTo[rr][n_o]
    st = gc.state_arrays()
    stage("combined")
    got = np.concatenate([st["F"], st["T"]], axis=1)
    scale = max(np.max(np.abs(x)) for x in FTo)
    assert np.max(np.abs(got - tot)) < 1e-10 * scale, np.max(np.abs(got - tot)) / scale
    # identical bits on every rank
    t = torch.tensor(got, device=dev)
    t0 = t.clone()
    dist.broadcast(t0, 0)
    assert torch.equal(t, t0), "combined F/T differ between ranks"
    gc.close()
    stage("cloud 1 closed")

    # ---- loop 5 replicated over the ranks == the serial reference, with periodic copies and deletions ----
    fast = cases.assembly(n=48, n_particles=60, seed=4, moving=True, cyclic_without_walls=True)
    fast.particles.v *= 6.0
    fast.dt = 0.5
    ocs = OracleCloud(fast)
    gcs = ParticleCloud(fast, device=local, slab=slab_ranges(fast.mesh.nz, world)[rank], nccl_id=nccl_unique_id_shared(dist, dev, rank),
                        rank=rank, n_ranks=world)
    stage("cloud 2 created")
    rng = np.random.default_rng(0)
    for s in range(3):
        FT = rng.normal(0, 1e-3, (400, 6))
        ocs.set_force_torque(FT[:ocs.n_instances]); ocs.combine_forces(); ocs.advance(fast.dt)
        gcs.set_force_torque(FT[:gcs.nParticles()] / world); gcs.combineForces(); gcs.advance(fast.dt)
        a, b = ocs.state_arrays(), gcs.state_arrays()
        for k in ("index", "copy", "active", "onMesh", "center", "theta", "v", "omega"):
            assert np.array_equal(a[k], b[k]), (k, s)
        stage(f"step {s} equal; owned steps {gcs.counter('loop5_owned_steps')}, fallbacks {gcs.counter('loop5_fallbacks')}, flags {gcs.counter('loop5_cluster_flags')}")
    assert ocs.n_instances != 60
    assert gcs.counter("loop5_fallbacks") == 0
    assert gcs.counter("loop5_owned_steps") > 0  # every rank advanced what its slab owns; states and copies came over NCCL
    gcs.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK", world)
    dist.destroy_process_group()


def nccl_unique_id_shared(dist, dev, rank):
    import torch
    from ibmfoam_b200.cloud import nccl_unique_id
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())


if __name__ == "__main__":
    main()
