"""world_size-2 `gloo` test (CPU) of the multi-rank host logic: slab partition, per-slab (decomposed) oracle clouds,
and the all-reduce that replaces particleCloud::combineForces' cross-rank sums (particleCloud.C:201-264).  The
GPU side of the same path is tests/test_gpu_multirank.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import copy
    from ibmfoam_b200 import cases
    from ibmfoam_b200.cloud import slab_ranges
    from oracle.pyoracle import OracleCloud
    case = cases.assembly(n=32, n_particles=14, seed=21)
    k0, k1 = slab_ranges(case.mesh.nz, world)[rank]
    sc = copy.copy(case)
    sc.mesh = case.mesh.slab(k0, k1)
    oc = OracleCloud(sc)
    f = cases.synthetic_fields(case.mesh)
    nxy = case.mesh.nx * case.mesh.ny
    sl = slice(k0 * nxy, k1 * nxy)
    FT = oc.integrate_stress(f["p"][sl], f["tau"][sl])
    dense = torch.zeros((case.particles.n, 6), dtype=torch.float64)
    for n_o, idx in enumerate(oc.state_arrays()["index"]):
        dense[int(idx)] += torch.from_numpy(FT[n_o])
    local = dense.clone()
    dist.all_reduce(dense, op=dist.ReduceOp.SUM)
    F, n_upd = oc.forcing(f["U"][sl], f["Uold"][sl], f["S"][sl], case.dt)
    upd = torch.tensor([n_upd], dtype=torch.int64)
    dist.all_reduce(upd)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), local=local.numpy(), total=dense.numpy(), F=F, k0=k0, k1=k1,
             upd=upd.numpy(), n_inst=oc.n_instances)
    dist.destroy_process_group()


def test_two_rank_gloo_decomposed_oracle(tmp_path):
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # slabs tile the mesh
    assert r[0]["k0"] == 0 and r[0]["k1"] == r[1]["k0"] and r[1]["k1"] == 32
    # the all-reduced totals are the sum of the rank partials, identical on both ranks
    assert np.array_equal(r[0]["total"], r[1]["total"])
    assert np.allclose(r[0]["total"], r[0]["local"] + r[1]["local"], rtol=0, atol=0)
    # against the undecomposed oracle: particles whose three point layers stay clear of the cut see no difference,
    # the others differ (the reference's decomposition dependence, App. B #20)
    sys.path.insert(0, ROOT)
    from ibmfoam_b200 import cases
    from oracle.pyoracle import OracleCloud
    case = cases.assembly(n=32, n_particles=14, seed=21)
    oc = OracleCloud(case)
    f = cases.synthetic_fields(case.mesh)
    FT = oc.integrate_stress(f["p"], f["tau"])
    st = oc.state_arrays()
    z = st["center"][:, 2]
    cut = float(r[0]["k1"])
    clear = (np.abs(z - cut) > 4 + 1 + 1.01) & (z > 6.01) & (z < 32 - 6.01)
    assert clear.any() and (~clear).any()
    scale = np.max(np.abs(FT))
    idx = st["index"]
    assert np.max(np.abs(r[0]["total"][idx[clear]] - FT[clear])) < 1e-12 * scale
    assert np.max(np.abs(r[0]["total"][idx[~clear]] - FT[~clear])) > 1e-9 * scale
    # every shell cell is owned by exactly one slab: the update counts add up
    F, n_upd = oc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    assert int(r[0]["upd"][0]) == n_upd
