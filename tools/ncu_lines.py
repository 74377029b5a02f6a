"""Per-source-line hot spots from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
SORTKEY = sys.argv[4] if len(sys.argv) > 4 else "samples"   # or "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr = None, None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":   # aggregated source-line row
        d = dict(zip(hdr, r))
        key = (fname, int(r[0]))
        s = int(d["# Samples"] or 0)
        if s == 0: continue
        a = agg.setdefault(key, dict(src=r[1].strip(), samples=0, inst=0, tags=0, shw=0, shi=0, stalls=collections.Counter()))
        a["samples"] += s; a["inst"] += int(d["Instructions Executed"] or 0); a["tags"] += int(d["L1 Tag Requests Global"] or 0)
        a["shw"] += int(d["L1 Wavefronts Shared"] or 0); a["shi"] += int(d["L1 Wavefronts Shared Ideal"] or 0)
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"): a["stalls"][k[6:]] += int(v)
tot = sum(a["samples"] for a in agg.values())
ti = sum(a["inst"] for a in agg.values())
print(f"total samples {tot}, warp instructions {ti}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:top]:
    st = ",".join(f"{k}:{v}" for k, v in a["stalls"].most_common(3))
    print(f"{a['samples']/tot:6.3f} inst {a['inst']/ti:5.3f} tags {a['tags']:9d} shw {a['shw']:8d}/{a['shi']:8d} {f}:{ln:4d} [{st}] {a['src'][:90]}")
