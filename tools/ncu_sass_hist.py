"""Groups the SASS of one kernel by execution count (= loop level: per chunk, per tile, per CTA) with the opcode mix
and stall samples of each group.  Input as for ncu_sass_top.py.  usage: ncu_sass_hist.py x.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h:i for i,h in enumerate(hdr)}
data = rows[2:]
ex = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in data:
    e = int(r[idx['Instructions Executed']] or 0)
    ex[e] += 1; samp[e] += int(r[idx['# Samples']] or 0)
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
    op = m.group(2).split('.')[0] if m else '?'
    ops[e][op] += 1
tot_ex = sum(e*c for e,c in ex.items())
print('total warp instr', tot_ex)
for e,c in sorted(ex.items(), key=lambda kv: -kv[0]*kv[1])[:12]:
    print('ex=%8d  n_instr=%5d  share_of_executed=%5.1f%%  samples=%5d  top ops: %s' % (e, c, 100*e*c/tot_ex, samp[e], dict(ops[e].most_common(12))))
