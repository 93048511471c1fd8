import csv, sys, subprocess
rep, fname, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file=None; hdr=None
agg=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; ii=[i for i,h in enumerate(hdr) if h=='Instructions Executed'][0]; isamp=[i for i,h in enumerate(hdr) if h=='# Samples'][0]; continue
    if r[0]!='' and hdr:
        try: agg.append((cur_file,int(r[0]),r[1][:95],int(r[ii]),int(r[isamp])))
        except: pass
tot=sum(a[3] for a in agg); ts=sum(a[4] for a in agg); print(tot, ts)
for a in sorted(agg,key=lambda a:-a[4])[:n]:
    print(f"{a[0]:14s} {a[1]:4d} inst={a[3]:9d} ({100*a[3]/tot:4.1f}%) samp={a[4]:5d} ({100*a[4]/ts:4.1f}%)  {a[2]}")
