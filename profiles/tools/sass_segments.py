import csv, sys
path = sys.argv[1]; units = float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iSt = hdr.index("Warp Stall Sampling (All Samples)")
# print segments: consecutive instructions with similar exec count
seg=[]; cur=None
out=[]
for n,r in enumerate(rows[2:]):
    if len(r)<=iI: continue
    c=int(r[iI]); st=int(r[iSt] or 0)
    out.append((n, r[iS].strip()[:60], c/units, st))
tot_st=sum(o[3] for o in out)
# group by exec count bucket
i=0
while i < len(out):
    j=i; c=out[i][2]
    while j < len(out) and abs(out[j][2]-c) <= 0.02*max(c,0.01)+0.002: j+=1
    n=j-i; s=sum(o[2] for o in out[i:j]); st=sum(o[3] for o in out[i:j])
    if s>0.5 or st/tot_st>0.005: print("%5d..%5d  n=%4d  exec/unit each=%.3f  sum=%.1f  stall%%=%.1f   first: %s" % (out[i][0], out[j-1][0], n, c, s, 100*st/tot_st, out[i][1]))
    i=j
