import csv, collections, sys
base = sys.argv[1]
rows = list(csv.reader(open(base + '_raw.csv')))
hdr = rows[0]; units = rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__grid_size','sm__cycles_elapsed.max','smsp__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_active.avg']
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')][:70])
    for k in keys:
        if k in hdr: print("  %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
rd = csv.reader(open(base + '_src.csv'))
name = next(rd); hdr = next(rd); idx = {h:i for i,h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); top = []; n = 0; ops = collections.Counter()
for r in rd:
    if len(r) < len(hdr): continue
    try: ns = int(r[idx['# Samples']])
    except: continue
    n += 1
    for s in stalls:
        try: tot[s] += int(r[idx[s]])
        except: pass
    src = r[idx['Source']].strip()
    try: ex = int(r[idx['Instructions Executed']])
    except: ex = 0
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    ops[op.split('.')[0]] += ex
    top.append((ns, src[:90], ex))
S = sum(tot.values())
print("SASS instrs", n, "samples", S)
for s, v in tot.most_common(8): print("  %-26s %5.1f%%" % (s, 100 * v / S))
print("executed by opcode:", [(k, v) for k, v in ops.most_common(14)])
top.sort(reverse=True)
for t in top[:14]: print(t)
