"""Summarise an ncu report into profiles/<name>.md: key metrics, derived per-tick figures, hottest source lines.

    python scripts/summarize_ncu.py gpurun_out/x.ncu-rep profiles/r01_x.md "description" [replica_ticks]
"""
import collections, csv, io, os, subprocess, sys

rep, out, desc = sys.argv[1], sys.argv[2], sys.argv[3]
replica_ticks = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__thread_inst_executed.sum', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_inst0.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic']
lines = ['# ' + desc, '', 'source report: `%s` (%s)' % (rep, os.environ.get('NCU_HOW', 'ncu --set full --clock-control none --import-source on')), '']
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    lines.append('## launch: %s' % d.get('Kernel Name', '?'))
    lines.append('')
    lines.append('| metric | value | unit |')
    lines.append('|---|---|---|')
    for k in WANT[1:]:
        if k in d:
            lines.append('| %s | %s | %s |' % (k, d[k], u.get(k, '')))
    try:
        inst = float(d['sm__inst_executed.sum'].replace(',', ''))
        if replica_ticks:
            lines.append('| derived: warp instructions per replica-tick | %.0f | |' % (inst / replica_ticks))
    except Exception:
        pass
    lines.append('')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, ''])
cur, h, total = None, None, 0
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        h = r; continue
    if h is None or len(r) < 8:
        continue
    d = dict(zip(h, r))
    try:
        ln = int(d['Line No']); ie = int(d['Instructions Executed'] or 0); sm = int(d['# Samples'] or 0)
    except Exception:
        continue
    agg[(cur, ln)][0] += ie; agg[(cur, ln)][1] += sm; agg[(cur, ln)][2] = r[1][:100].replace('|', '/')
    total += ie
if total:
    lines += ['## hottest source lines (share of executed warp instructions, stall samples)', '', '| file:line | inst % | samples | source |', '|---|---|---|---|']
    for (f, ln), (ie, sm, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append('| %s:%d | %.1f | %d | `%s` |' % (f, ln, 100.0 * ie / total, sm, s))
open(out, 'w').write('\n'.join(lines) + '\n')
print('wrote', out)
