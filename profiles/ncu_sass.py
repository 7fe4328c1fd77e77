"""Hot SASS instructions of one kernel of an ncu report: python ncu_sass.py rep.ncu-rep <kernel substring> [min Minst]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) * 1e6 if len(sys.argv) > 3 else 5e6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i, done = 0, False
while i < len(rows) and not done:
    r = rows[i]
    if r and r[0] == "Kernel Name" and pat in r[1]:
        hdr = rows[i + 1]
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        j, tot = i + 2, 0
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            try:
                body.append((int(rows[j][ie]), int(rows[j][isamp]), rows[j][1]))
                tot += int(rows[j][ie])
            except Exception:
                pass
            j += 1
        print("==", r[1], "total warp inst", tot)
        for n, sm, s in body:
            if n >= thr:
                print(f"{n/1e6:8.1f}M s{sm:6d} {s[:110]}")
        done = True
    i += 1
