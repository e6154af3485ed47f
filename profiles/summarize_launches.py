"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
tot, cnt, seq = collections.defaultdict(float), collections.Counter(), []
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0].replace('ldm::', '')
    t = float(r[mv].replace(',', ''))
    tot[name] += t
    cnt[name] += 1
    seq.append((name, t, r[gi]))
T = sum(tot.values())
print(f"launches {len(seq)}  total {T / 1e3:.1f} us (per-launch times are cold-cache and serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k:28s} n={cnt[k]:4d} total={v / 1e3:9.1f} us {100 * v / T:5.1f}%  avg={v / cnt[k] / 1e3:7.1f} us")
if len(sys.argv) > 2:
    g = [(i, s) for i, s in enumerate(seq) if sys.argv[2] in s[0]]
    print("top launches of", sys.argv[2], [(i, round(s[1] / 1e3, 1), s[2]) for i, s in sorted(g, key=lambda x: -x[1][1])[:24]])
