#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share.
usage: launch_summary.py launches.csv [skip-kernel-name ...]   (named kernels are listed but left out of the share)"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = set(sys.argv[2:])
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.defaultdict(list)
for r in data:
    if len(r) > mv and r[mn] == 'gpu__time_duration.sum':
        agg[r[kn].split('(')[0]].append(float(r[mv].replace(',', '')))
tot = sum(sum(v) for k, v in agg.items() if k not in skip)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = "setup (excluded)" if k in skip else "share of step %.1f%%" % (100 * sum(v) / tot)
    print("%-40s n=%4d total %.3f ms  avg %.1f us  %s" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, share))
