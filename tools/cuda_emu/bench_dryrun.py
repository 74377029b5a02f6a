#!/usr/bin/env python
"""TEST INFRASTRUCTURE — a DRY RUN of bench.py's control flow on a machine without a GPU: the emulated library
(build_emu.py) stands in for libibmfoam_b200.so, torch's CUDA entry points are replaced by CPU stand-ins, and every
workload is shrunk to a 32^3 box.  It exists to catch mistakes in bench.py's logic (timing bookkeeping, end-to-end leg,
JSON assembly) before GPU time is spent on them.  The numbers it
prints are meaningless and are marked as such ("dry_run": true); bench.py itself contains no such switch.

    python tools/cuda_emu/bench_dryrun.py [bench.py arguments]"""
import contextlib
import importlib.util
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main(argv):
    import build_emu
    lib = build_emu.build()
    from ibmfoam_b200 import capi
    capi.LIB_PATH = lib
    import torch

    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 0.0

    cpu = torch.device("cpu")
    real_device, real_empty, real_tensor, real_zeros, real_arange = torch.device, torch.empty, torch.tensor, torch.zeros, torch.arange
    strip = lambda k: {a: (cpu if a == "device" else v) for a, v in k.items() if a != "pin_memory"}
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = FakeEvent
    torch.device = lambda *a, **k: cpu
    torch.empty = lambda *a, **k: real_empty(*a, **strip(k))
    torch.tensor = lambda *a, **k: real_tensor(*a, **strip(k))
    torch.zeros = lambda *a, **k: real_zeros(*a, **strip(k))
    torch.arange = lambda *a, **k: real_arange(*a, **strip(k))

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for k in list(bench.WORKLOADS):
        bench.WORKLOADS[k] = (32, bench.WORKLOADS[k][1])          # every workload: a 32^3 box (24 spheres)
    real_make_case = bench.make_case

    def make_case(workload):
        from ibmfoam_b200 import cases
        n, moving = bench.WORKLOADS[workload]
        return cases.assembly(n=n, phi=0.2, moving=moving, cyclic_without_walls=True)

    bench.make_case = make_case
    bench.cpu_sample = lambda *a, **k: (1.0, 1, 1.0, 24)            # the CPU baseline legs are not what is being dry-run
    bench.all_cores_sample = lambda *a, **k: (1.0, 1, 1.0, 24, 1)
    os.environ.setdefault("IBM_BENCH_WATCHDOG_S", "7200")   # emulated kernels are slow; the watchdog's retry is for hardware runs
    sys.argv = ["bench.py"] + argv
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    line = json.loads([ln for ln in buf.getvalue().splitlines() if ln.startswith("{")][-1])
    line["dry_run"] = True
    print(json.dumps(line))
    return line


if __name__ == "__main__":
    main(sys.argv[1:])
