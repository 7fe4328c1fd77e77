"""Aggregate the warp-stall samples / instructions of an ncu report by source-line ranges (phases of a long kernel).
usage: ncu_phases.py report.ncu-rep source.cu 'name:regex-of-first-line' ...   (phases in source order)"""
import csv, re, subprocess, sys
rep, src = sys.argv[1], sys.argv[2]
marks = [a.split(":", 1) for a in sys.argv[3:]]
text = open(src).read().split("\n")
bounds = []
for name, rx in marks:
    ln = next(i + 1 for i, l in enumerate(text) if re.search(rx, l))
    bounds.append((ln, name))
bounds.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
iw, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
acc = {name: [0, 0, 0, 0] for _, name in bounds}
acc["(before)"] = [0, 0, 0, 0]
for r in rows[hi + 1:]:
    try:
        ln = int(r[0]); n = int(r[ie]); sm = int(r[isamp]); w = int(r[iw] or 0); wi = int(r[ii] or 0)
    except Exception:
        continue
    name = "(before)"
    for b, nm in bounds:
        if ln >= b:
            name = nm
    a = acc[name]; a[0] += n; a[1] += sm; a[2] += w; a[3] += wi
ti = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
print(f"{'phase':28s} {'inst':>10s} {'%':>6s} {'samples':>9s} {'%':>6s} {'smem wavefronts':>16s} {'ideal':>12s}")
for name, a in acc.items():
    print(f"{name:28s} {a[0]/1e6:9.1f}M {100*a[0]/max(ti,1):5.1f}% {a[1]:9d} {100*a[1]/max(ts,1):5.1f}% {a[2]/1e6:15.1f}M {a[3]/1e6:11.1f}M")
