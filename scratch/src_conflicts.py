"""Per-opcode shared-memory wavefront / bank-conflict / stall summary of an `ncu --page source --csv` export."""
import csv, sys
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f)))
    hdr=None; idx=None
    tot_w=tot_e=tot_i=0
    items=[]; stalls={}; nk=0
    for r in rows:
        if r and r[0]=="Kernel Name":
            nk+=1
            if nk>1: break
            continue
        if r and r[0]=="Address":
            hdr=r; idx={h:i for i,h in enumerate(hdr)}; continue
        if hdr is None or len(r)<len(hdr): continue
        gi=lambda k:int(float(r[idx[k]] or 0)) if r[idx[k]] not in ("-","") else 0
        w=gi("L1 Wavefronts Shared"); e=gi("L1 Wavefronts Shared Excessive"); ide=gi("L1 Wavefronts Shared Ideal"); n=gi("Instructions Executed")
        tot_w+=w; tot_e+=e; tot_i+=n
        s=r[idx["Source"]].strip()
        if w: items.append((e,w,ide,n,s))
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                stalls[k]=stalls.get(k,0)+gi(k)
    print(f, "wavefronts",tot_w,"excessive",tot_e,"inst",tot_i)
    agg={}
    for e,w,ide,n,s in items:
        t=s.split(); op=t[1] if t[0].startswith("@") else t[0]
        a=agg.setdefault(op,[0,0,0,0]); a[0]+=e;a[1]+=w;a[2]+=ide;a[3]+=n
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("  %-28s excess %9d wave %9d ideal %9d inst %9d"%(k,*v))
    items.sort(reverse=True)
    for e,w,ide,n,s in items[:10]: print("     ",e,w,ide,n,s)
    tot=sum(stalls.values()) or 1
    print("  stalls %:", {k[6:]:round(100*v/tot,1) for k,v in sorted(stalls.items(), key=lambda kv:-kv[1]) if v/tot>0.01})
