import csv, collections, subprocess, sys
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass']+sys.argv[3:],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur_file=None; cur_line=None
agg=collections.defaultdict(lambda:[0,0,0,''])
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; ie=hdr.index('Instructions Executed'); it=hdr.index('Thread Instructions Executed'); isamp=hdr.index('# Samples'); continue
    if r[0]!='':
        cur_line=(cur_file,int(r[0])); agg[cur_line][3]=r[1]; continue
    if r[2] in ('...',''): continue
    try:
        a=agg[cur_line]; a[0]+=int(r[ie]); a[1]+=int(r[it]); a[2]+=int(r[isamp])
    except: pass
tot=sum(a[0] for a in agg.values()); ts=sum(a[2] for a in agg.values())
print('total warp inst',tot,'samples',ts)
byfile=collections.Counter()
for k,a in agg.items(): byfile[k[0]]+=a[0]
print(dict(byfile))
for k,a in sorted(agg.items(), key=lambda x:-x[1][2])[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[2]/max(ts,1):5.1f}% smp  thr/inst {a[1]/max(a[0],1):5.1f}  {k[0]}:{k[1]}  {a[3].strip()[:100]}")
