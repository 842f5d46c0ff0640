"""Hottest SASS instructions of an ncu source-page CSV with their dominant stall reasons.  Usage: ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]; col = {k: i for i, k in enumerate(h)}; data = rows[2:]
tot = sum(int(r[col['# Samples']] or 0) for r in data)
stall = [k for k in h if k.startswith('stall_') and 'Not' not in k]
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col['# Samples']] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    s = int(r[col['# Samples']] or 0)
    st = sorted(((int(r[col[k]] or 0), k[6:]) for k in stall), reverse=True)[:3]
    print(f"{i:5d} {100*s/tot:5.2f}% x{int(r[col['Instructions Executed']] or 0)/1e6:8.1f}M  {r[col['Source']].strip()[:70]:70s} " + ' '.join(f'{k}:{100*v/tot:.1f}' for v, k in st if v))
