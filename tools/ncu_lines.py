"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (development aid).
usage: python tools/ncu_lines.py export.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ia = hdr.index("Address")
        ii = hdr.index("Instructions Executed")
        it = hdr.index("Thread Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[ia] == "-":   # the source-line row (metrics aggregated over its SASS)
        key = (cur_file, line)
        try:
            inst = int(r[ii]); thr = int(r[it]); samp = int(r[isamp])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, 0, r[1]])
        a[0] += inst; a[1] += thr; a[2] += samp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp-instr {tot_i}  samples {tot_s}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f:22s}:{l:5d} inst {a[0]:11d} {100*a[0]/max(tot_i,1):5.1f}%  lanes {a[1]/max(a[0],1):5.1f}  samp {100*a[2]/max(tot_s,1):5.1f}%  | {a[3].strip()[:110]}")
