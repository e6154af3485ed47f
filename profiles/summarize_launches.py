"""Summarises an `ncu --csv` launch list (any subset of gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum): per-kernel launch counts, total time, share of the step, DRAM bytes.
    python profiles/summarize_launches.py profiles/r1_dram_unet_step.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mn, mv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0].replace('ldm::', '').replace('void ', '')
    agg[name][r[mn]] += float(r[mv].replace(',', ''))
    if r[mn] == 'gpu__time_duration.sum':
        cnt[name] += 1
T = sum(d['gpu__time_duration.sum'] for d in agg.values())
print(f"launches {sum(cnt.values())}  total {T / 1e3:.1f} us (ncu: cold-cache, serialised launches: compare shares, not absolutes)")
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    t = d['gpu__time_duration.sum']
    line = f"{k:34s} n={cnt[k]:4d} time={t / 1e3:8.1f} us {100 * t / T:5.1f}%"
    if 'dram__bytes_read.sum' in d:
        line += f"  dram_read={d['dram__bytes_read.sum'] / 1e6:8.1f} MB  dram_write={d['dram__bytes_write.sum'] / 1e6:7.1f} MB"
    print(line)
