#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total time / share for the LAST proof."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
recs = [(r[kn].split('(')[0].replace('void ', ''), float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
n_proofs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
# a proof starts at the first dft_tile launch after a query_kernel
starts = [0] + [i + 1 for i, (k, _) in enumerate(recs) if k.startswith('map_kernel') and i + 1 < len(recs)
                and any(t in recs[i + 1][0] for t in ('dft_tile', 'trace_expand', 'wl_chunk'))]   # ... or at its converter
seq = recs[starts[-1]:] if len(starts) > 1 else recs
tot = sum(c for _, c in seq)
print(f"launches in last proof: {len(seq)}  total {tot / 1000:.1f} us")
agg = {}
for k, c in seq:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += c
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:70]}` | {v[0]} | {v[1] / 1000:.1f} | {100 * v[1] / tot:.1f}% |")
