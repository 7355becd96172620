import csv, sys, subprocess
kern=sys.argv[1]; skip=sys.argv[2] if len(sys.argv)>2 else '0'
out=subprocess.run(['ncu','-i','r01_full.ncu-rep' if len(sys.argv)<4 else sys.argv[3],'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern,'--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[h]; 
ci=hdr.index('Instructions Executed'); cs=hdr.index('# Samples'); 
tot_i=0; tot_s=0; lines=[]
for r in rows[h+1:]:
    if r[0]!='':   # cuda source line aggregate
        try: lines.append((int(r[0]), r[1], int(r[ci]), int(r[cs])))
        except: pass
tot_i=sum(l[2] for l in lines); tot_s=sum(l[3] for l in lines)
print('total instr',tot_i,'samples',tot_s)
for l in sorted(lines,key=lambda x:-x[3])[:int(sys.argv[4]) if len(sys.argv)>4 else 30]:
    print(f"{l[0]:5d} {100*l[2]/tot_i:5.1f}%i {100*l[3]/tot_s:5.1f}%s  {l[1][:110]}")
