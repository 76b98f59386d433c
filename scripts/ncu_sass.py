"""Executed-instruction histogram by SASS opcode from an ncu source-page CSV (cuda,sass)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
ops = collections.Counter(); tot = 0
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); continue
    if hdr and len(r) == len(hdr):
        try: n = int(r[iI])
        except ValueError: continue
        sass = r[3].strip()
        toks = sass.split()
        if not toks or toks[0] == '-': continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        op = '.'.join(op.split('.')[:2]) if op.startswith(('IMAD','LDS','STS','LDG','STG','SHFL')) else op.split('.')[0]
        ops[op] += n; tot += n
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print('total', tot)
for op, n in ops.most_common(40):
    print(f"{op:14s} {100*n/tot:5.1f}%  {n/scale:10.2f}")
