import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]
ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
thr=float(sys.argv[2])
for r in rows[hi+1:]:
    try: n=int(r[ie])
    except: continue
    if n>=thr: print(f"{n/1e6:8.1f}M s{int(r[isamp]):6d} {r[1][:100]}")
