"""Generates tests/golden/*.npz from the REFERENCE'S OWN CODE (oracle/_ref/libibm_ref.so = /root/reference/src compiled
against oracle/mockfoam by oracle/ref/Makefile).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The vectors are committed; tests/test_golden_oracle.py (CPU) checks the restatement against them and
tests/test_gpu_golden.py (GPU box, where /root/reference does not exist) checks the CUDA path against them.
Inputs are not stored: every case is rebuilt from its seeded builder in tests/ref_cases.py, and a digest of the
inputs is stored so that a drifting builder is noticed.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ibm_testutil as util  # noqa: E402
import ref_cases  # noqa: E402
from oracle.pyoracle import OracleCloud  # noqa: E402

STATE_I = ("index", "copy", "active", "onMesh", "nShell", "N")
STATE_D = ("center", "theta", "v", "omega", "F", "T")


def input_digest(case):
    h = hashlib.sha256()
    p = case.particles
    for a in (p.kind, p.position, p.d, p.a, p.b, p.nTheta, p.nk, p.delta, p.theta, p.v, p.omega, p.rho):
        h.update(np.ascontiguousarray(a).tobytes())
    m = case.mesh
    h.update(repr((m.nx, m.ny, m.nz, tuple(m.origin), tuple(m.delta), m.blocks, m.patches, case.dt)).encode())
    c = case.cloud
    h.update(repr((c.moving, c.rotating, c.coupled, c.e, tuple(c.UInf), c.rhoRef, c.pRef, tuple(c.g))).encode())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def put_state(out, prefix, st):
    for k in STATE_I:
        out[prefix + k] = st[k].astype(np.int32)
    for k in STATE_D:
        out[prefix + k] = st[k]


def static_vectors(case):
    rc = OracleCloud(case, impl="reference")
    out = {"digest": input_digest(case), "n_instances": np.int32(rc.n_instances)}
    put_state(out, "state_", rc.state_arrays())
    for i in range(rc.n_instances):
        sh, pt = rc.shell(i), rc.points(i)
        out[f"shell{i}_cells"] = sh["cells"]
        out[f"shell{i}_jk"] = sh["jk"]
        out[f"shell{i}_W"] = sh["W"]          # total IDW weight of the 8 neighbour points
        out[f"points{i}_cell"] = pt["cell"]   # point -> cell membership
        out[f"points{i}_nFaces"] = pt["nFaces"].astype(np.int8)
        out[f"points{i}_W"] = pt["W"]         # total IDW weight of the face gather
    out["indicator"] = np.packbits(rc.indicator() != 0)
    f = util.caller_fields(case)
    F, n_upd = rc.forcing(f["U"], f["Uold"], f["S"], case.dt)
    nz = np.flatnonzero(np.any(F != 0, axis=1)).astype(np.int32)
    out["forcing_cells"] = nz
    out["forcing_F"] = F[nz]
    out["forcing_updates"] = np.int64(n_upd)
    out["stress_FT"] = rc.integrate_stress(f["p"], f["tau"])
    rc.combine_forces()
    st = rc.state_arrays()
    out["combined_FT"] = np.concatenate([st["F"], st["T"]], axis=1)
    out["cd"] = rc.cd()
    out["mass"] = np.array([rc.mass(i) for i in range(rc.n_instances)])
    out["inertia"] = np.array([rc.inertia(i) for i in range(rc.n_instances)])
    return out


def dynamic_vectors(case, n_steps):
    rc = OracleCloud(case, impl="reference")
    out = {"digest": input_digest(case), "n_steps": np.int32(n_steps)}
    put_state(out, "s0_", rc.state_arrays())
    f = util.caller_fields(case)
    for s in range(1, n_steps + 1):
        # the F/T the sub-steps of this solve() will use: stress integral + combineForces on the pre-step instances
        rc.integrate_stress(f["p"], f["tau"])
        rc.combine_forces()
        st = rc.state_arrays()
        out[f"s{s}_FT_used"] = np.concatenate([st["F"], st["T"]], axis=1)
        stats = rc.solve(f["p"], f["tau"], case.dt)
        put_state(out, f"s{s}_", rc.state_arrays())
        out[f"s{s}_stats"] = stats[1:]
    for i in range(rc.n_instances):
        sh = rc.shell(i)
        out[f"final_shell{i}_cells"] = sh["cells"]
        out[f"final_shell{i}_jk"] = sh["jk"]
    return out


def main():
    for name, builder in sorted(ref_cases.STATIC_CASES.items()):
        vec = static_vectors(builder())
        path = os.path.join(HERE, f"static_{name}.npz")
        np.savez_compressed(path, **vec)
        print(f"{path}: {os.path.getsize(path)} bytes, {int(vec['n_instances'])} instances, {int(vec['forcing_updates'])} updates")
    for name, (builder, n_steps) in sorted(ref_cases.DYNAMIC_CASES.items()):
        vec = dynamic_vectors(builder(), n_steps)
        path = os.path.join(HERE, f"dynamic_{name}.npz")
        np.savez_compressed(path, **vec)
        print(f"{path}: {os.path.getsize(path)} bytes, {n_steps} steps")


if __name__ == "__main__":
    main()
