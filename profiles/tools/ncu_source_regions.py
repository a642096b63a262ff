import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}
data = rows[2:]
# segment by change of exec count > 15%
segs=[]; cur=None
for n,r in enumerate(data):
    ex=int(r[ix["Instructions Executed"]]); s=int(r[ix["# Samples"]])
    if cur and abs(ex-cur['ex'])<=0.15*max(cur['ex'],1):
        cur['n']+=1; cur['tot']+=ex; cur['s']+=s; cur['end']=n
    else:
        cur=dict(start=n,end=n,ex=ex,n=1,tot=ex,s=s); segs.append(cur)
tot=sum(s['tot'] for s in segs); ts=sum(s['s'] for s in segs)
for s in segs:
    if s['tot']>0.002*tot:
        print("lines %4d-%4d n=%3d exec/inst %8.3fM total %6.2f%% samples %6.2f%%  %s"%(s['start'],s['end'],s['n'],s['ex']/1e6,100*s['tot']/tot,100*s['s']/ts, data[s['start']][1].strip()[:50]))
