#!/usr/bin/env python
"""Turns what tools/gpu_first_call.sh left in gpurun_out/first_call/ into one markdown summary (stdout): the legs and their
exit codes, the micro-benchmark, one row per bench line (step time, phase breakdown, kernel generations, roofline
fraction), and the failed tests of each pytest leg.  Copy the result into profiles/ as the round's summary.

    python tools/summarize_first_call.py [gpurun_out/first_call]
"""
import json
import os
import re
import sys


def main(d):
    out = []
    summ = os.path.join(d, "summary.txt")
    if os.path.exists(summ):
        out.append("## legs\n\n| leg | exit code | seconds |\n|---|---|---|")
        for ln in open(summ):
            m = re.match(r"(\S+) rc=(\d+) (\d+)s", ln.strip())
            if m:
                out.append(f"| {m.group(1)} | {m.group(2)} | {m.group(3)} |")
    mb = os.path.join(d, "microbench.log")
    if os.path.exists(mb):
        for ln in open(mb):
            if ln.startswith("{"):
                j = json.loads(ln)
                out.append("\n## micro-benchmark\n")
                out.append(f"fp64 {j['fp64_tflops']:.1f} TFLOP/s (DFMA), rsqrt {j['fp64_rsqrt_per_s']:.3g}/s, launch {j['launch_us_stream']:.2f} us "
                           f"(stream) / {j['launch_us_graph']:.2f} us (graph), grid barrier "
                           + ", ".join(f"{k} blocks {v:.2f} us" for k, v in j["grid_sync_us"].items()) + f", copy {j['hbm_copy_GBs']:.0f} GB/s")
    rows = []
    for name in sorted(os.listdir(d)) if os.path.isdir(d) else []:
        if not (name.startswith("bench_") and name.endswith(".log")):
            continue
        for ln in open(os.path.join(d, name), errors="ignore"):
            if ln.startswith("{"):
                try:
                    j = json.loads(ln)
                except ValueError:
                    continue
                kg = (j.get("config") or {}).get("kernel_generations") or {}
                br = j.get("breakdown_ms_per_step") or {}
                rf = j.get("roofline") or {}
                rows.append(f"| {name[6:-4]} | {j.get('impl', 'b200')} | {j.get('value') or 0:.3g} | {j.get('ms_per_step') or 0:.2f} | "
                            f"{br.get('update_ms', 0):.2f} / {br.get('forcing_ms', 0):.2f} / {br.get('stress_ms', 0):.2f} / {br.get('advance_ms', 0):.2f} | "
                            f"f{kg.get('forcing')} s{kg.get('stress')} m{kg.get('marking')} split={kg.get('loop5_quiet_active_split')} "
                            f"fused={kg.get('loop5_active_list_fused')} active={kg.get('loop5_active_instances')} fallbacks={kg.get('loop5_fallbacks')} | "
                            f"{100.0 * (rf.get('frac') or 0):.1f} % ({rf.get('kernel')}) | {(j.get('e2e') or {}).get('value') or 0:.3g} |")
    if rows:
        out.append("\n## bench lines\n\n| leg | impl | updates/s | ms/step | update / forcing / stress / advance ms | variants | roofline | e2e updates/s |\n|---|---|---|---|---|---|---|---|")
        out += rows
    fails = []
    for name in sorted(os.listdir(d)) if os.path.isdir(d) else []:
        if name.startswith("pytest_") and name.endswith(".log"):
            txt = open(os.path.join(d, name), errors="ignore").read()
            tail = [ln for ln in txt.splitlines() if ln.startswith("FAILED") or ln.startswith("ERROR") or " passed" in ln or " failed" in ln]
            fails.append(f"* `{name[7:-4]}`: " + ("; ".join(tail[-6:]) if tail else "no summary line (timeout?)"))
    if fails:
        out.append("\n## parity on hardware\n")
        out += fails
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join("gpurun_out", "first_call"))
