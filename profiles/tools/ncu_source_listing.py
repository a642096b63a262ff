import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h:i for i,h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
samp = sum(int(r[ix["# Samples"]]) for r in data)
print("total inst", tot, "samples", samp)
# print compact listing: idx, addr offset, exec count (M), samples, top stall, source
base = int(data[0][0],16)
lo = int(sys.argv[2]) if len(sys.argv)>2 else 0
hi = int(sys.argv[3]) if len(sys.argv)>3 else len(data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n,r in enumerate(data[lo:hi], lo):
    ex = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[h]]),h[6:]) for h in stalls), reverse=True)[:2]
    print("%4d %5x %8.2fM %5d %-26s %s" % (n, int(r[0],16)-base, ex/1e6, s, ",".join("%s:%d"%(b,a) for a,b in st if a>0), r[1].strip()[:90]))
