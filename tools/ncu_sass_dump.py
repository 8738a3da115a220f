"""Prints, in address order, the SASS instructions of one kernel whose execution count is in a given set, with their stall
samples — i.e. one loop level of the kernel.  usage: ncu_sass_dump.py x.csv 75648[,18912]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h:i for i,h in enumerate(hdr)}
want = set(int(x) for x in sys.argv[2].split(','))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for n, r in enumerate(rows[2:]):
    e = int(r[idx['Instructions Executed']] or 0)
    if e in want:
        s = int(r[idx['# Samples']] or 0)
        why = max(((int(r[idx[x]] or 0), x[6:]) for x in stalls))
        print('%5d ex=%7d s=%4d %-14s %s' % (n, e, s, why[1] if s else '', r[idx['Source']].strip()[:100]))
