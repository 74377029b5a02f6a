"""The device-resident entry points of the C ABI (what bench.py and a multi-GPU harness call: SoA lattice fields already in
device memory) against the host entry points, which the other parity tests cover.  Buffers are torch CUDA tensors on a
GPU; under the CPU emulation of the library (IBM_TEST_EMU_LIB, tests/conftest.py) device memory is host memory and numpy
arrays stand in."""
import os

import numpy as np
import pytest

from ibmfoam_b200 import cases
from ibmfoam_b200.cloud import ParticleCloud
import ibm_testutil as util

pytestmark = pytest.mark.gpu
EMULATED = bool(os.environ.get("IBM_TEST_EMU_LIB"))


class DeviceBuffers:
    """[comp][k][j][i] SoA buffers in device memory"""

    def __init__(self, arrays):
        if EMULATED:
            self.t = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
            self.ptr = [a.ctypes.data for a in self.t]
        else:
            import torch
            self.t = [torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda") for a in arrays]
            torch.cuda.synchronize()
            self.ptr = [a.data_ptr() for a in self.t]

    def host(self, i):
        return self.t[i] if EMULATED else self.t[i].cpu().numpy()


def _soa(a):
    return np.ascontiguousarray(a.T) if a.ndim == 2 else np.ascontiguousarray(a)


@pytest.mark.parametrize("moving", [False, True])
def test_device_entry_points_equal_host_entry_points(moving):
    case = cases.assembly(n=40, n_particles=30, seed=6, moving=moving, cyclic_without_walls=True)
    f_lat = cases.synthetic_fields(case.mesh)          # lattice order
    f = util.caller_fields(case)                       # caller order (single block: the same)
    C = case.mesh.n_cells
    host, dev = ParticleCloud(case), ParticleCloud(case)
    buf = DeviceBuffers([_soa(f_lat[k]) for k in ("U", "Uold", "S", "p", "tau")] + [np.zeros((3, C))])
    dU, dUold, dS, dP, dTau, dF = buf.ptr
    for step in range(3):
        Fh = host.forcing(f["S"], f["U"], f["Uold"], case.dt)
        # (a) into a caller-provided device buffer, (b) into the library's persistent field, read back with download_field
        dev.forcing_device(dU, dUold, dS, case.dt, dF)
        dev.sync()
        Fa = buf.host(5).reshape(3, C).T
        dev.forcing_device(dU, dUold, dS, case.dt)
        Fb = dev.download_field(None, 3)
        l2c = case.mesh.lattice_to_cell()
        assert np.array_equal(Fa, Fh[np.argsort(l2c)] if case.mesh.blocks else Fh)
        assert np.array_equal(Fb, Fh)
        if moving:
            sh = host.solve(f["p"], f["tau"], case.dt)
            sd = dev.solve_device(dP, dTau, case.dt)
            assert (sh.nSubSteps, sh.nInstances) == (sd.nSubSteps, sd.nInstances)
        else:
            FTh = host.integrateSurfaceStresses(f["p"], f["tau"])
            FTd = np.empty_like(FTh)
            dev.integrate_stress_device(dP, dTau, FTd)
            assert np.array_equal(FTd, FTh)
        a, b = host.state_arrays(), dev.state_arrays()
        for k in a:
            assert np.array_equal(a[k], b[k]), (step, k)
    th, td = host.timings(), dev.timings()
    assert th["n_updates"] == td["n_updates"] // 2 and td["h2d_bytes"] < th["h2d_bytes"]
    host.close(); dev.close()


def test_upload_download_round_trip_in_caller_order():
    case = cases.colliding_particles(n_particles=3)     # four-block numbering: caller order differs from lattice order
    C = case.mesh.n_cells
    rng = np.random.default_rng(0)
    a = rng.normal(size=(C, 3))
    gc = ParticleCloud(case)
    buf = DeviceBuffers([np.zeros((3, C))])
    gc.upload_field(a, 3, buf.ptr[0])
    gc.sync()
    soa = buf.host(0).reshape(3, C)
    l2c = case.mesh.lattice_to_cell()
    assert np.array_equal(soa.T, a[l2c])                # lattice cell l holds the caller's cell l2c[l]
    assert np.array_equal(gc.download_field(buf.ptr[0], 3), a)
    gc.close()
