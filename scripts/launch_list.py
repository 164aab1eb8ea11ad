"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv <cmd>`) ->
profiles/<name>.md: one row per kernel launch and the share of every kernel in the summed kernel time.

    python scripts/launch_list.py gpurun_out/r02_launches.csv profiles/r02_bench_launch_list.md "title / command"
"""
import collections
import csv
import sys

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        ms = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else (v if u == 'ms' else v * 1e3))
        name = r['Kernel Name'].split('(')[0].replace('void ', '')
        rows.append((name, r['Grid Size'], r['Block Size'], r['Stream'], ms))
tot = sum(x[4] for x in rows)
by = collections.defaultdict(float)
for x in rows:
    by[x[0]] += x[4]
md = ['# ' + title, '', 'ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are serialised and cold-cache: compare shares, not absolutes).', '',
      '| kernel | launches | summed duration (ms) | share of the kernel time |', '|---|---|---|---|']
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    md.append('| %s | %d | %.2f | %.1f %% |' % (k, sum(1 for x in rows if x[0] == k), v, 100 * v / tot if tot else 0))
md += ['', '| # | kernel | grid | block | stream | duration (ms) |', '|---|---|---|---|---|---|']
for i, x in enumerate(rows):
    md.append('| %d | %s | %s | %s | %s | %.2f |' % (i, x[0], x[1], x[2], x[3], x[4]))
md.append('')
md.append('Sum over %d launches: %.1f ms.' % (len(rows), tot))
open(out, 'w').write('\n'.join(md) + '\n')
print('wrote', out, len(rows), 'launches')
