import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name'): continue
    if r[0] != '' and r[2] == '-':
        try: agg.append((cur, int(r[0]), r[1].strip()[:110], int(r[6]), int(r[7])))
        except: pass
ti = sum(a[4] for a in agg); ts = sum(a[3] for a in agg)
print("total instr", ti, "samples", ts)
bf = collections.Counter()
for a in agg: bf[a[0]] += a[4]
print([(k, round(100*v/ti,1)) for k, v in bf.most_common()])
agg.sort(key=lambda a: -a[3])
for a in agg[:n]: print("%-20s %5d instr %5.1f%% samp %5.1f%%  %s" % (a[0][:20], a[1], 100*a[4]/ti, 100*a[3]/ts, a[2]))
