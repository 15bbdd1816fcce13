"""Splits the SASS page of an ncu capture of gaussblur_kernel into the kernel's phases (developer tool).
  ncu -i gpurun_out/<tag>_gaussblur.ncu-rep --page source --csv --print-source sass > sass.csv
  python tools/ncu_regions.py sass.csv <total SMSP cycles of the launch, smsp__cycles_active.sum>
Per phase: share of the warp-state samples (= warp time), of the executed warp instructions, FMA-pipe cycles
(FMUL2 / FFMA2 count two) as a fraction of the given cycle count, top stall reasons. The phases are found from
the positions of the FMUL2 instructions (two dense runs: the horizontal and the vertical tap loop)."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
tot=sum(float(r[ix["# Samples"]]) for r in data)
toti=sum(float(r[ix["Instructions Executed"]]) for r in data)
def op(r):
    o=r[ix["Source"]].split(); return o[1] if o[0].startswith("@") else o[0]
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
f2=[i for i,r in enumerate(data) if op(r)=="FMUL2"]
gaps=[(f2[i],f2[i+1]) for i in range(len(f2)-1) if f2[i+1]-f2[i]>60]
print("n",len(data),"samples",tot,"inst",toti,"gaps",gaps)
def region(a,b,name):
    n=sum(float(data[i][ix["# Samples"]]) for i in range(a,b)); ins=sum(float(data[i][ix["Instructions Executed"]]) for i in range(a,b))
    st={k[6:]:sum(float(data[i][ix[k]] or 0) for i in range(a,b)) for k in stalls}
    top=" ".join("%s:%.1f"%(k,100*v/tot) for k,v in sorted(st.items(),key=lambda x:-x[1])[:6])
    fma=sum(float(data[i][ix["Instructions Executed"]])*(2 if op(data[i]) in("FMUL2","FFMA2") else 1) for i in range(a,b) if op(data[i]).split('.')[0] in ("FMUL2","FFMA2","FFMA","FMUL","FADD","IMAD","HFMA2"))
    print("%-14s %4d-%4d samp %5.1f%% inst %5.1f%% fma-cycles %5.1f%% of all cycles | %s"%(name,a,b,100*n/tot,100*ins/toti,100*fma/float(sys.argv[2]),top))
hl0=f2[0]; hl1=gaps[0][0]; vl0=gaps[-1][1]; vl1=f2[-1]
region(0,hl0-40,"setup/sync")
region(hl0-40,hl0,"h prologue")
region(hl0,hl1+10,"h loop")
region(hl1+10,vl0-60,"h epi+sync")
region(vl0-60,vl0,"v prologue")
region(vl0,vl1+12,"v loop")
region(vl1+12,len(data),"v epi+rest")
