import csv,sys,subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Line No"][0]
hdr=rows[hi]
ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
tot=0; lines=[]
for r in rows[hi+1:]:
    if r[0]!="" :
        try:
            n=int(r[ie]); lines.append((n,int(r[isamp]),r[0],r[1])); tot+=n
        except: pass
print("total inst", tot)
for n,sm,ln,src in sorted(lines,reverse=True)[:topn]:
    print(f"{n/tot*100:5.1f}% {n/1e6:8.1f}M  samples {sm:7d}  L{ln}: {src[:100]}")
