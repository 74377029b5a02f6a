"""The product's CUDA sources under the CPU emulation (tools/cuda_emu): `build_emu.py` compiles ibmfoam_b200/csrc/*.cu
unchanged (only `<<<...>>>` and `extern __shared__` are rewritten syntactically) against stand-in CUDA headers, and the GPU
parity tests are then run against that library in a child pytest.  This is how kernel and host logic is checked where no
GPU exists - it found the stale-graph bug of loop 5 - and it is test infrastructure only: the product never
loads the emulated library, `-m gpu` on a B200 remains the parity gate, and nothing here says anything about performance.

Default: a subset that runs in about a minute after a ~1.5 min build.  IBM_RUN_EMULATED_FULL=1 runs every GPU test file
(several minutes: loop 5 executes ~10 kernels x 100 sub-steps per fluid step, one fiber per CUDA thread)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


def _child(emu_lib, args, timeout, **extra_env):
    env = dict(os.environ, IBM_TEST_EMU_LIB=emu_lib, **extra_env)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", *args], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-2500:]
    assert r.returncode == 0, tail
    return tail


SUBSET = [
    "tests/test_gpu_marking.py::test_marking_bit_exact",
    "tests/test_gpu_zz_reference_vectors.py::test_marking_forcing_stress_against_reference_vectors",
    "tests/test_gpu_zz_reference_vectors.py::test_trajectories_against_reference_vectors[col8-graph]",
    "tests/test_gpu_v7_kernels.py",
    "tests/test_zzz_gpu_solid_fraction.py",
    "tests/test_zzz_gpu_active_loop5.py::test_copy_created_after_a_rebound_next_to_a_resting_body[fused]",
    "tests/test_zzz_gpu_active_loop5.py::test_copy_created_after_a_rebound_next_to_a_resting_body[cluster]",
    "tests/test_zzz_gpu_active_loop5.py::test_two_spheres_head_on[fused]",
    "tests/test_zzz_gpu_active_loop5.py::test_two_spheres_head_on[cluster]",
    "tests/test_gpu_zz_reference_vectors.py::test_trajectories_against_reference_vectors[wc2d_cyl-cluster]",
    "tests/test_gpu_motion.py::test_two_spheres_head_on_closed_form[fused]",
    "tests/test_zzz_gpu_soft_sphere.py",
    "tests/test_zzz_gpu_mark_variant.py",
    # several ranks (in-process transport): near-slab marking list, compacted launches, halo
    "tests/test_gpu_multirank.py::test_loops_1_to_4_match_decomposed_oracle[2]",
    "tests/test_gpu_multirank.py::test_halo_makes_results_decomposition_independent[2]",
    # loop 5 owned by slab: two ranks, states and new periodic copies exchanged once per fluid step, == the serial oracle
    "tests/test_gpu_multirank.py::test_loop5_replicated_equals_serial_oracle[2-1]",
]


def test_gpu_parity_subset_under_emulation(emu_lib):
    out = _child(emu_lib, SUBSET, timeout=1500)
    assert " passed" in out and "failed" not in out


def test_refused_cooperative_launch_falls_back_to_the_unfused_chain(emu_lib):
    """fused_active_loop5 on a device / context that refuses cooperative launches: same results through the unfused chain"""
    out = _child(emu_lib, ["tests/test_zzz_gpu_active_loop5.py::test_two_spheres_head_on[fused]",
                           "tests/test_zzz_gpu_active_loop5.py::test_copy_created_after_a_rebound_next_to_a_resting_body[fused]"],
                 timeout=1500, EMU_REFUSE_COOP="1")
    assert " passed" in out and "failed" not in out


@pytest.mark.parametrize("no_lattice", [False, True])
def test_shim_drop_in_under_emulation(emu_lib, tmp_path, no_lattice):
    """shim/particleCloud.C + the emulated library against the reference's own class through the same solver-side driver;
    no_lattice: without the optional `lattice` dictionary entry - the shim describes the mesh from polyMesh's cell and face
    centres alone (ibm_lattice_from_arrays), as it would in a stock IBMFoam case"""
    import numpy as np
    import dropin_util as du
    if not os.path.exists(du.REFERENCE_BIN):
        pytest.skip("dropin_reference not built (needs /root/reference)")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "shim", "dropin"), "-s", "shim_emu"])
    case, n_steps, n_outer = du.CASES["cyl2d_fixed"]()
    cp = str(tmp_path / "case.txt")
    du.write_case(cp, case, n_steps, n_outer)
    outs = {}
    for side, binary in (("reference", du.REFERENCE_BIN), ("shim", os.path.join(du.BUILD, "dropin_shim_emu"))):
        op = str(tmp_path / f"{side}.bin")
        r = du.run(binary, cp, op, no_lattice=no_lattice)
        assert r.returncode == 0, (side, r.stderr)
        outs[side] = du.parse(op)
    (h_ref, rec_ref), (h_shim, rec_shim) = outs["reference"], outs["shim"]
    assert h_ref == h_shim and [t for t, _ in rec_ref] == [t for t, _ in rec_shim]
    for (tag, a), (_, b) in zip(rec_ref, rec_shim):
        if tag in ("index", "copy", "active", "CoNum"):
            assert np.array_equal(a, b), tag
        else:
            assert np.max(np.abs(a - b)) <= (1e-12 if tag == "F" else 1e-9) * max(np.max(np.abs(a)), 1e-300), tag


def test_shim_restart_under_emulation(emu_lib, tmp_path):
    """write -> destroy -> construct at time != 0 (readParticles) -> continue: shim + emulated library against the reference's own
    restart path (particleCloudIO.C:47-69, particleIBMIO.C:111-213), moving spheres; also the particleIBM views at the end"""
    import numpy as np
    import dropin_util as du
    if not os.path.exists(du.REFERENCE_BIN):
        pytest.skip("dropin_reference not built (needs /root/reference)")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "shim", "dropin"), "-s", "shim_emu"])
    case, _, _ = du.CASES["spheres3d"]()
    cp = str(tmp_path / "case.txt")
    du.write_case(cp, case, 3, 1, n_restart=3)
    outs = {}
    for side, binary in (("reference", du.REFERENCE_BIN), ("shim", os.path.join(du.BUILD, "dropin_shim_emu"))):
        op = str(tmp_path / f"{side}.bin")
        r = du.run(binary, cp, op)
        assert r.returncode == 0, (side, r.stderr)
        outs[side] = du.parse(op)
    (h_ref, rec_ref), (h_shim, rec_shim) = outs["reference"], outs["shim"]
    assert h_ref == h_shim and [t for t, _ in rec_ref] == [t for t, _ in rec_shim]
    assert dict(rec_shim)["restarted"][0] == case.particles.n
    for (tag, a), (_, b) in zip(rec_ref, rec_shim):
        if tag in ("index", "copy", "active", "CoNum", "restarted", "view_index"):
            assert np.array_equal(a, b), tag
        else:
            assert np.max(np.abs(a - b)) <= (1e-12 if tag == "F" else 1e-9) * max(np.max(np.abs(a)), 1e-300), tag


@pytest.mark.skipif(not os.environ.get("IBM_RUN_EMULATED_FULL"), reason="set IBM_RUN_EMULATED_FULL=1 (several minutes)")
@pytest.mark.parametrize("path", ["tests/test_gpu_forcing.py", "tests/test_gpu_marking.py", "tests/test_gpu_stress.py", "tests/test_gpu_motion.py",
                                  "tests/test_gpu_multirank.py", "tests/test_gpu_zz_reference_vectors.py", "tests/test_gpu_zz_device_api.py", "tests/test_gpu_v7_kernels.py", "tests/test_zzz_gpu_solid_fraction.py", "tests/test_zzz_gpu_active_loop5.py",
                                  "tests/test_zz_gpu_validation_cases.py", "tests/test_zzz_gpu_soft_sphere.py", "tests/test_zzz_gpu_mark_variant.py"])
def test_gpu_parity_file_under_emulation(emu_lib, path):
    _child(emu_lib, [path], timeout=6000)


@pytest.mark.skipif(not os.environ.get("IBM_RUN_EMULATED_FULL"), reason="set IBM_RUN_EMULATED_FULL=1 (about two minutes)")
def test_bench_control_flow_dry_run(emu_lib):
    """bench.py's logic (probe, selection, full-size re-verification, timing, end-to-end leg, JSON assembly) on a 32^3
    stand-in workload against the emulated library; the numbers mean nothing, the keys and the flow are what is checked"""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cuda_emu", "bench_dryrun.py"), "--steps", "2", "--warmup", "1",
                        "--e2e-steps", "1"], cwd=ROOT, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
    kg = line["config"]["kernel_generations"]
    assert kg["forcing"] == 7 and kg["stress"] == 7
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
