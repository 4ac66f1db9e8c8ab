#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel of ONE step
(the launches between the last two set_diag_kernel launches belong to the e2e step; we take the last complete step)."""
import csv, collections, sys, re

def main(path, nsteps_in_file=6):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    recs = [(re.sub(r"\(.*", "", r[kn]), float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
    # step boundaries: scatter/set_diag launches start a factorisation
    starts = [i for i, (k, _) in enumerate(recs) if k.startswith("set_diag_kernel")]
    # several set_diag launches per factorisation are adjacent; keep the first of each run
    first = [s for j, s in enumerate(starts) if j == 0 or s - starts[j - 1] > 8]
    if len(first) >= 2:
        lo, hi2 = first[-2], first[-1]
    else:
        lo, hi2 = 0, len(recs)
    step = recs[lo:hi2]
    agg = collections.OrderedDict()
    for k, t in step:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{path}: {len(recs)} launches in file, last complete step = launches [{lo},{hi2}) = {len(step)} launches, "
          f"sum of kernel time {tot/1e6:.3f} ms (cold-cache, serialised)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:45s} n={n:5d}  {t/1e6:9.3f} ms  {100*t/tot:5.1f}%  avg {t/n/1e3:9.1f} us")

if __name__ == "__main__":
    main(sys.argv[1])
