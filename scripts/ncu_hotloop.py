"""Per-SASS-instruction executed counts and stall samples of the hot loop, from `ncu --page source --print-source sass` CSV.
usage: ncu_hotloop.py file.csv rounds  -> prints instructions with executed/rounds ratio >= thr, and totals"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) != len(hdr): continue
    data.append((r[iS].strip(), int(r[iI]), int(r[iSm]), {h: int(r[i]) for i, h in stall_cols if r[i] not in ("", "0")}))
tot_i = sum(d[1] for d in data); tot_s = sum(d[2] for d in data)
mx = max(d[1] for d in data)
# the hot loop = instructions executed at least 2% as often as the most executed one
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
print(f"total warp-inst {tot_i:.4e}, samples {tot_s}, max per-instruction count {mx:.4e}")
acc_i = acc_s = 0
for k, (src, n, sm, st) in enumerate(data):
    if n >= thr * mx:
        acc_i += n; acc_s += sm
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{k:5d} {n / mx:6.3f} {100 * sm / tot_s:5.2f}%  {src[:70]:70s} {' '.join(f'{h[6:]}={v}' for h, v in top)}")
print(f"listed: {100 * acc_i / tot_i:.1f}% of instructions, {100 * acc_s / tot_s:.1f}% of samples")
