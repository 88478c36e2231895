import csv, sys, collections, re
path = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 3e7
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed")
iW = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
mix = collections.Counter(); wf = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iI: continue
    src = r[iS].strip()
    toks = src.split()
    if not toks: continue
    op = toks[0]
    if op.startswith("@"): op = toks[1]
    base = op.split(".")[0]
    if base in ("SHFL","LDS","STS","LDG","F2F","F2I","I2F","MUFU","DADD","DFMA","DMUL","DSETP","FSETP","ISETP"): pass
    n = int(r[iI]); mix[base] += n; tot += n
    if iW is not None:
        try: wf[base] += int(r[iW])
        except: pass
print("total warp-instr per unit: %.1f" % (tot/units))
for k, v in mix.most_common(40):
    print("%-10s %8.1f   wavefronts %.1f" % (k, v/units, wf[k]/units))
