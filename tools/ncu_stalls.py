"""Per-CUDA-source-line warp-stall samples of one kernel from an .ncu-rep captured with --import-source on:
   python tools/ncu_stalls.py report.ncu-rep [top_n]   (reads the report here with `ncu -i`, no GPU needed)."""
import csv, sys, subprocess
rep=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 30
txt=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
cur=None; tot=0; out=[]; hdr=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if len(r)==2 and r[0]=="Function Name": print(r[1]); continue
    if len(r)>2 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<10: continue
    if r[0]!='' and r[2]=='-':
        k=int(r[6]) if r[6].isdigit() else 0
        stalls={}
        for i,h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h and i<len(r) and r[i].isdigit() and int(r[i])>0:
                stalls[h[6:]]=int(r[i])
        out.append((k,cur,int(r[0]),r[1].strip()[:95],sorted(stalls.items(), key=lambda kv:-kv[1])[:2], r[7]))
        tot+=k
print('total',tot)
for o in sorted(out, key=lambda o:-o[0])[:n]:
    print(f"{o[0]:5d} {100*o[0]/tot:4.1f}% {o[1]}:{o[2]} x{o[5]} | {o[3]} | {o[4]}")
