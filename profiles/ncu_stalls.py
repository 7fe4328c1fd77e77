"""Per-opcode stall breakdown of one kernel from an ncu report (SASS page): which instructions the warps wait on and why.
usage: ncu_stalls.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
isamp, iexe, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
by_op = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hi + 1:]:
    try:
        n = int(r[isamp])
    except Exception:
        continue
    op = r[isrc].split()[0] if r[isrc].split() else "?"
    if op.startswith("@"):
        op = r[isrc].split()[1]
    op = op.split(".")[0]
    by_op[op]["samples"] += n
    by_op[op]["inst"] += int(r[iexe] or 0)
    for h in reasons:
        v = int(r[hdr.index(h)] or 0)
        by_op[op][h] += v
        tot[h] += v
allsamp = sum(c["samples"] for c in by_op.values())
print("total samples", allsamp, {k: v for k, v in tot.most_common(8)})
for op, c in sorted(by_op.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    rs = ", ".join(f"{h[6:]} {c[h]}" for h in sorted(reasons, key=lambda h: -c[h])[:4] if c[h])
    print(f"{op:10s} samples {c['samples']:7d} ({100*c['samples']/allsamp:4.1f}%)  inst {c['inst']/1e6:8.1f}M  | {rs}")
