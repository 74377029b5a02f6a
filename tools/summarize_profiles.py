"""Summaries of ncu captures for profiles/: the headline raw metrics of each `--set full` report and its per-line hot spots.
usage: python tools/summarize_profiles.py <dir with *.ncu-rep> <prefix written under profiles/>   (e.g. gpurun_out/r2p r2)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("time_ms", "gpu__time_duration.sum"), ("dram_read_GB", "dram__bytes_read.sum"), ("dram_write_GB", "dram__bytes_write.sum"),
    ("regs", "launch__registers_per_thread"), ("block", "launch__block_size"), ("grid", "launch__grid_size"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"), ("occ_limit_regs", "launch__occupancy_limit_registers"),
    ("occ_limit_smem", "launch__occupancy_limit_shared_mem"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("l1tex_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("dram_throughput_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("issue_active_pct", "smsp__issue_active.avg.pct"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), ("warp_inst", "smsp__inst_executed.sum"),
    ("shared_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
]
SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
TIME = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    res = []
    for v in rows[2:]:
        d = {"kernel": v[h.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for name, key in KEYS:
            if key not in h:
                continue
            i = h.index(key)
            try:
                x = float(v[i].replace(",", ""))
            except ValueError:
                continue
            if name.endswith("_GB"):
                x *= SCALE.get(u[i], 1.0)
            elif name == "time_ms":
                x *= TIME.get(u[i], 1.0)
            elif name == "smem_dyn_KB":
                x *= {"Kbyte/block": 1.0, "byte/block": 1e-3}.get(u[i], 1.0)
            d[name] = x
        res.append(d)
    return res


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    prof = os.path.join(ROOT, "profiles")
    allk = {}
    for f in sorted(os.listdir(src)):
        if not f.endswith(".ncu-rep"):
            continue
        name = f[:-8]
        rep = os.path.join(src, f)
        ks = raw(rep)
        allk[name] = ks
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, name, "30"], capture_output=True, text=True).stdout
        with open(os.path.join(prof, f"{prefix}_lines_{name}.txt"), "w") as o:
            o.write(f"# {name}: ncu --set full --clock-control none --import-source on (one launch at the bench's full size)\n")
            for k in ks:
                o.write("# " + json.dumps(k) + "\n")
            o.write(lines)
    with open(os.path.join(prof, f"{prefix}_ncu_summary.json"), "w") as o:
        json.dump(allk, o, indent=1)
    for n, ks in allk.items():
        for k in ks:
            print(n, json.dumps(k))


if __name__ == "__main__":
    main()
