"""Top stalled instructions of a kernel from `ncu --page source --csv` output."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ci={h:i for i,h in enumerate(hdr)}
S=ci['Warp Stall Sampling (All Samples)']; X=ci['Instructions Executed']; A=ci['Avg. Threads Executed']
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data=[]; tot=0
for r in rows[2:]:
    try: v=float(r[S])
    except: continue
    tot+=v
    top=sorted(((float(r[ci[s]] or 0),s) for s in stalls),reverse=True)[:2]
    data.append((v,r[ci['Source']].strip()[:70],r[X],r[A],top))
n=int(sys.argv[2]) if len(sys.argv)>2 else 25
for v,sx,x,a,top in sorted(data,key=lambda t:-t[0])[:n]:
    print("%5.1f%% %-70s exec=%9s thr=%5s %s"%(100*v/tot,sx,x,a," ".join("%s=%d"%(s[6:],c) for c,s in top if c)))
