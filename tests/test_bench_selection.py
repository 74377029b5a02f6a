"""bench.py's choice among kernel generations / loop-5 modes / marking kernels from a probe result (CPU: the probe result
is a cached JSON file, no child process, no GPU): a variant is used only if the probe showed parity AND a lower time."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture()
def bench(tmp_path, monkeypatch):
    import bench as b
    monkeypatch.setattr(b, "PROBE_CACHE", str(tmp_path / "probe.json"))
    return b


def _choose(bench, info):
    with open(bench.PROBE_CACHE, "w") as fh:
        json.dump(info, fh)
    fv, sv, split, mark, got = bench.choose_variants(0)
    assert got.get("cached")
    return fv, sv, split, mark


BASE = {"ok": True, "kernels_ok": True, "forcing6_ok": False, "mark2_ok": True, "loop5_split_ok": True, "loop5_fused_active_ok": True,
        "forcing_ms": {"2": 13.0, "5": 3.0}, "stress_ms": {"4": 9.0, "5": 2.5},
        "advance_ms": {"0": 11.0, "1": 8.0, "2": 5.0}, "update_ms": {"0": 3.0, "1": 2.0}}


def test_everything_verified_and_faster_is_used(bench):
    assert _choose(bench, BASE) == (5, 5, 2, 1)


def test_slower_variants_are_not_used(bench):
    info = dict(BASE, forcing_ms={"2": 13.0, "5": 14.0}, stress_ms={"4": 9.0, "5": 9.5}, advance_ms={"0": 11.0, "1": 12.0, "2": 11.5},
                update_ms={"0": 3.0, "1": 3.1})
    assert _choose(bench, info) == (bench.DEFAULT_VARIANTS["forcing"], bench.DEFAULT_VARIANTS["stress"], 0, 0)


def test_parity_failures_veto_whatever_the_time(bench):
    info = dict(BASE, kernels_ok=False, mark2_ok=False, loop5_split_ok=False, loop5_fused_active_ok=False)
    assert _choose(bench, info) == (bench.DEFAULT_VARIANTS["forcing"], bench.DEFAULT_VARIANTS["stress"], 0, 0)
    info = dict(BASE, loop5_fused_active_ok=False)                      # the persistent kernel failed or hung: plain split
    assert _choose(bench, info) == (5, 5, 1, 1)
    info = dict(BASE, loop5_split_ok=True, loop5_fused_active_ok=True, advance_ms={"0": 11.0, "1": 8.0})   # no timing of mode 2
    assert _choose(bench, info)[2] == 1


def test_failed_probe_keeps_the_defaults(bench):
    assert _choose(bench, {"ok": False, "error": "timeout"}) == (bench.DEFAULT_VARIANTS["forcing"], bench.DEFAULT_VARIANTS["stress"], 0, 0)
    assert _choose(bench, {"ok": True}) == (bench.DEFAULT_VARIANTS["forcing"], bench.DEFAULT_VARIANTS["stress"], 0, 0)


def test_cached_stencil_forcing_is_chosen_only_when_verified_and_fastest(bench):
    info = dict(BASE, forcing6_ok=True, forcing_ms={"2": 13.0, "5": 3.0, "6": 2.4})
    assert _choose(bench, info)[0] == 6
    info = dict(BASE, forcing6_ok=True, forcing_ms={"2": 13.0, "5": 3.0, "6": 3.4})
    assert _choose(bench, info)[0] == 5
    info = dict(BASE, forcing6_ok=False, forcing_ms={"2": 13.0, "5": 3.0, "6": 2.4})
    assert _choose(bench, info)[0] == 5
    info = dict(BASE, kernels_ok=False, forcing6_ok=True, forcing_ms={"2": 13.0, "5": 3.0, "6": 2.4})   # stress v5 disagreed, forcing 6 is fine
    assert _choose(bench, info)[:2] == (6, bench.DEFAULT_VARIANTS["stress"])


def test_failed_run_with_probed_variants_is_retried_once_with_the_defaults(bench, monkeypatch):
    calls = {}
    monkeypatch.setattr(bench, "main", lambda: (_ for _ in ()).throw(RuntimeError("kernel written without hardware failed")))
    monkeypatch.setattr(bench.os, "execv", lambda exe, argv: calls.setdefault("argv", argv))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "5"])
    bench._guarded_main()
    assert calls["argv"][-1] == "--no-probe" and "--steps" in calls["argv"]
    for argv in (["bench.py", "--no-probe"], ["bench.py", "--variant", "5"], ["bench.py", "--impl", "reference"]):
        calls.clear()
        monkeypatch.setattr(sys, "argv", argv)
        with pytest.raises(RuntimeError):
            bench._guarded_main()                 # nothing to fall back to: the error surfaces
        assert not calls
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    with pytest.raises(RuntimeError):
        bench._guarded_main()                     # multi-rank runs are not retried
