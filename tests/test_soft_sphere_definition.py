"""The numpy definition of the soft-sphere extension (oracle/soft_sphere_oracle.py) against closed forms - CPU only.
The GPU path is compared with this definition bit for bit in tests/test_zzz_gpu_soft_sphere.py."""
import numpy as np

from oracle import soft_sphere_oracle as sso


def test_schedule_is_the_literal_loop():
    dts = sso.substep_schedule(0.25)
    assert len(dts) in (100, 101) and abs(sum(dts) - 0.25) < 1e-15 + 0.25 / 100   # App. B #22: 100 or 101 sub-steps


def test_undamped_law_is_elastic_and_clipping_only_matters_with_damping():
    m_eff, kn = 0.7, 40.0
    assert abs(sso.restitution_clipped(kn, 0.0, m_eff) - 1.0) < 1e-6
    for e in (0.9, 0.7):
        ln_e = np.log(e)
        gn = -2.0 * ln_e * np.sqrt(kn * m_eff / (np.pi ** 2 + ln_e ** 2))
        assert abs(sso.restitution(kn, gn, m_eff) - e) < 1e-12
        assert e <= sso.restitution_clipped(kn, gn, m_eff) < e + 0.03


def test_pair_force_is_antisymmetric_and_capped():
    x1, x2 = np.array([0.0, 0.3, 0.0]), np.array([7.5, 0.0, 0.1])
    v1, v2 = np.array([1.0, 0.2, 0.0]), np.array([-1.0, 0.0, 0.3])
    w1, w2 = np.array([0.0, 0.0, 0.4]), np.array([0.1, 0.0, 0.0])
    F1, T1, F2, T2 = sso.pair_force(x1, x2, v1, v2, w1, w2, 4.0, 4.0, kn=50.0, gn=1.0, gt=1e9, mu=0.25)
    assert np.array_equal(F1, -F2)
    n = (x1 - x2) / np.linalg.norm(x1 - x2)
    Fn = F1 @ n
    Ft = np.linalg.norm(F1 - Fn * n)
    assert Fn > 0 and abs(Ft - 0.25 * Fn) < 1e-12 * Fn           # sliding: the Coulomb cap binds
    assert sso.pair_force(x1, x2 + np.array([1.0, 0, 0]), v1, v2, w1, w2, 4.0, 4.0, 50.0, 1.0, 1.0, 0.2) is None   # no overlap


def test_free_flight_and_momentum():
    st = {"center": [[0.0, 0, 0], [7.9, 0.5, 0]], "v": [[1.0, 0, 0], [-1.0, 0.1, 0]], "omega": np.zeros((2, 3)), "theta": np.zeros((2, 3)),
          "F": np.zeros((2, 3)), "T": np.zeros((2, 3))}
    mass, inertia = [2.0, 3.0], [6.4, 6.4]
    out = sso.advance(st, 0.25, [4.0, 4.0], mass, inertia, kn=80.0, gn=0.5, gt=0.5, mu=0.3)
    p0 = 2.0 * np.array(st["v"][0]) + 3.0 * np.array(st["v"][1])
    p1 = 2.0 * out["v"][0] + 3.0 * out["v"][1]
    assert np.allclose(p0, p1, atol=1e-13)
    assert np.abs(out["omega"]).max() > 0            # oblique contact with friction
