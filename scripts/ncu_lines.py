"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by (file, source line)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
cur = None; fname = ''; h = None; tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': h = r; ie = h.index('Instructions Executed'); ts = h.index('Thread Instructions Executed'); smp = h.index('# Samples'); continue
    if h is None or len(r) <= ie: continue
    if r[0].strip():
        try: cur = (fname, int(r[0])); agg[cur][3] = r[1].strip()[:100]
        except ValueError: continue
    try: v = int(r[ie]); t = int(r[ts]); s = int(r[smp])
    except ValueError: continue
    if cur is None: continue
    agg[cur][0] += v; agg[cur][1] += t; agg[cur][2] += s; tot += v
print('total inst', tot)
tsamp = sum(v[2] for v in agg.values()) or 1
for (f, l), (v, t, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:<4d} {100*v/tot:5.1f}% inst {100*s/tsamp:5.1f}% samp lanes {t/max(v,1):5.1f}  {src}")
