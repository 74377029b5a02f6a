"""Loading of the committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
reference's own code).  The case inputs are rebuilt from tests/ref_cases.py and checked against the stored digest."""
import importlib.util
import os

import numpy as np

import ref_cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

STATE_I = make_golden.STATE_I
STATE_D = make_golden.STATE_D


def load_static(name):
    case = ref_cases.STATIC_CASES[name]()
    g = np.load(os.path.join(GOLDEN, f"static_{name}.npz"))
    assert np.array_equal(g["digest"], make_golden.input_digest(case)), "case builder drifted from the golden inputs"
    return case, g


def load_dynamic(name):
    builder, n_steps = ref_cases.DYNAMIC_CASES[name]
    case = builder()
    g = np.load(os.path.join(GOLDEN, f"dynamic_{name}.npz"))
    assert np.array_equal(g["digest"], make_golden.input_digest(case)), "case builder drifted from the golden inputs"
    assert int(g["n_steps"]) == n_steps
    return case, g, n_steps


def dense_forcing(g, n_cells):
    F = np.zeros((n_cells, 3))
    F[g["forcing_cells"]] = g["forcing_F"]
    return F


def indicator(g, n_cells):
    return np.unpackbits(g["indicator"])[:n_cells].astype(np.float64)
