"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': []}
        sections.append(cur)
    elif len(r) >= 2 and r[0] == 'Function Name':
        cur['fn'] = r[1]
    elif r and r[0] == 'Line No':
        cur['hdr'] = r
    elif cur is not None and 'hdr' in cur and len(r) == len(cur['hdr']):
        cur['rows'].append(r)


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg, tot, tots = {}, 0, 0
for s in sections:
    h = s['hdr']
    iL, iI, iS = h.index('Line No'), h.index('Instructions Executed'), h.index('# Samples')
    for r in s['rows']:
        if r[iL] in ('', '-'):
            continue
        k = (s['file'].split('/')[-1], r[iL], r[1][:100])
        a = agg.setdefault(k, [0, 0, 0])
        a[0] += num(r[iI]); a[1] += num(r[iS]); a[2] += 1
        tot += num(r[iI]); tots += num(r[iS])
print('total warp-inst', tot, 'samples', tots)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:12]:>12}:{k[1]:>4} {100 * a[0] / tot:5.1f}% inst {100 * a[1] / max(tots, 1):5.1f}% smp {a[2]:3d} sass  {k[2]}")
