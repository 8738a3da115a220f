"""Top stalled SASS instructions of one kernel.  Input: `ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv`.
usage: ncu_sass_top.py x.csv [N]"""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
idx = {h:i for i,h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = rows[2:]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print('total samples', tot, 'instr', len(data))
# aggregate stall reasons overall
agg = collections.Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[idx[s]] or 0)
print({k:v for k,v in agg.most_common(10)})
# top instructions
top = sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for r in top:
    n = int(r[idx['# Samples']])
    why = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print('%6d %5.1f%% ex=%8s %s  | %s' % (n, 100*n/tot, r[idx['Instructions Executed']], r[idx['Source']].strip()[:70], why))
