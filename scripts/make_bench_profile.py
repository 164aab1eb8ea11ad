"""Reads ncu reports of the CURRENT kernels and (re)writes the entries of profiles/r02_bench_profile.json that bench.py's
roofline block uses: warp instructions per replica-tick (fifo) and DRAM bytes per unit of work, with their source.

    python scripts/make_bench_profile.py <workload> <report.ncu-rep> <units> <profiles/summary.md> "<description>"

units = replica-ticks of the captured launch (fifo workloads, printed by scripts/ncu_probe_grp.py) or swept runnable jobs
(sjf / dlas: replicas x summary.sum_queued, printed by scripts/ncu_probe_legacy.py).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, rep, units, md, desc = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4], sys.argv[5]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
d = dict(zip(hdr, rows[2]))


def f(name):
    return float(d[name].replace(',', ''))


def scaled(name):
    """value in base units (ncu prints Gbyte / Mbyte / Kbyte per row)"""
    u = dict(zip(hdr, rows[1]))[name].lower()
    m = {'gbyte': 1e9, 'mbyte': 1e6, 'kbyte': 1e3, 'byte': 1.0}.get(u, 1.0)
    return f(name) * m


inst = f('smsp__inst_executed.sum') if 'smsp__inst_executed.sum' in d else f('sm__inst_executed.sum')
dram = scaled('dram__bytes_read.sum') + scaled('dram__bytes_write.sum')
entry = {
    'kernel': d.get('Kernel Name'), 'source': md, 'description': desc, 'units_in_capture': units,
    'gpu_time_ms': f('gpu__time_duration.sum') if dict(zip(hdr, rows[1]))['gpu__time_duration.sum'] == 'ms' else f('gpu__time_duration.sum') / 1e3,
    'warp_inst_executed': inst, 'dram_bytes': dram, 'dram_bytes_per_unit': dram / units,
    'issue_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
    'warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active'),
    'threads_per_inst': f('smsp__thread_inst_executed_per_inst_executed.ratio') if 'smsp__thread_inst_executed_per_inst_executed.ratio' in d else None,
    'registers_per_thread': f('launch__registers_per_thread'),
}
if workload.startswith('fifo') or workload.startswith('env'):
    entry['warp_inst_per_replica_tick'] = inst / units
path = os.path.join(ROOT, 'profiles', 'r02_bench_profile.json')
allp = json.load(open(path)) if os.path.exists(path) else {}
allp[workload] = entry
json.dump(allp, open(path, 'w'), indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))
subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'summarize_ncu.py'), rep, os.path.join(ROOT, md), desc, str(units)])
