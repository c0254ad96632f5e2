"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions and stall samples per opcode,
and the top instructions by stall samples.   python tools/ncu_source_hist.py src.csv [ntiles]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); sm = collections.Counter(); tot_ex = tot_s = 0
items = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    op = src.split()[0]
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0].rstrip(";")
    e = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    ex[op] += e; sm[op] += s; tot_ex += e; tot_s += s
    items.append((s, e, r[ci["Address"]][-5:], src))
print(f"total executed warp-instr {tot_ex}  per tile {tot_ex/ntiles:.0f}   samples {tot_s}")
print(f"{'op':10s} {'exec/tile':>10s} {'exec%':>7s} {'samples%':>9s}")
for op, e in ex.most_common(22):
    print(f"{op:10s} {e/ntiles:10.1f} {100*e/tot_ex:6.1f}% {100*sm[op]/max(tot_s,1):8.1f}%")
print("top stall sites:")
for s, e, a, src in sorted(items, reverse=True)[:25]:
    print(f"  {100*s/max(tot_s,1):5.2f}%  exec/tile {e/ntiles:7.1f}  {a}  {src[:90]}")
