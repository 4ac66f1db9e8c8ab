#!/usr/bin/env python
"""Condense `ncu --page source --csv` output (can be 100+ MB) into the top-N instructions by stall samples per kernel.
usage: ncu -i rep.ncu-rep --page source --csv | python profiles/ncu_top_stalls.py [N] > top.txt"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
heads = [i for i, r in enumerate(rows) if len(r) > 2 and r[0] == 'Address' and r[1] == 'Source']
for hn, hi in enumerate(heads):
    end = heads[hn + 1] if hn + 1 < len(heads) else len(rows)
    hdr = rows[hi]
    sc = hdr.index('# Samples')
    stall_cols = [j for j, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[hi + 1:end] if len(r) > sc and r[sc].isdigit()]
    tot = sum(int(r[sc]) for r in body) or 1
    print(f"=== kernel instance {hn}: {len(body)} instructions, {tot} samples")
    agg = {}
    for r in body:
        for j in stall_cols:
            if r[j] not in ('', '0'):
                agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(body, key=lambda r: -int(r[sc]))[:n]:
        st = sorted([(int(r[j] or 0), hdr[j]) for j in stall_cols if r[j] not in ('', '0')], reverse=True)[:2]
        print(f"{int(r[sc]):7d} {100 * int(r[sc]) / tot:5.1f}%  {r[1][:100]:100s} {st}")
