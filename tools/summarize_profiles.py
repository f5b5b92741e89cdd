#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r1        # tag = round / iteration name
"""
import collections
import csv
import gzip
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
os.makedirs(P, exist_ok=True)

# ---- launch list -----------------------------------------------------------------------------
def src(name):
    """gpurun_out/<tag>_<name> (r2 scripts) or gpurun_out/<name> (r1 scripts)"""
    p = os.path.join(G, f'{tag}_{name}')
    return p if os.path.exists(p) else os.path.join(G, name)


rows = list(csv.reader(open(src('launches.csv'))))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mn, mu = (hdr.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit'))
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    v = float(r[mv].replace(',', '')) * {'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(r[mu], 1.0)
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f'{tag}_launches_summary.txt'), 'w') as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host` (default preconditioner)\n")
    f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    f.write(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:58s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:9.2f} {100 * v[1] / tot:6.1f}%\n")
with open(src('launches.csv'), 'rb') as fsrc, gzip.open(os.path.join(P, f'{tag}_launches.csv.gz'), 'wb') as dst:
    shutil.copyfileobj(fsrc, dst)

# ---- full capture of the PCG kernels -----------------------------------------------------------
rep = src('prof.ncu-rep') if os.path.exists(src('prof.ncu-rep')) else src('prof_cg.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, body = rr[0], rr[1], rr[2:]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__sass_inst_executed_op_shared_ld.sum', 'smsp__inst_executed.sum']
idx = [h.index(k) for k in keep if k in h]
traffic = {}
with open(os.path.join(P, f'{tag}_ncu_cg_metrics.csv'), 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['# ncu --set full --clock-control none --import-source on -k regex:k_cg_A|k_cg_B|k_mg_res|k_mg_line|k_mg_fused|k_diag|k_factor -s <skip> -c <n>: consecutive launches of the hot kernels inside one time step (cold cache per replay: the L2 window does not show here)'])
    w.writerow([h[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in body:
        w.writerow([r[i] for i in idx])
        kname = r[h.index('Kernel Name')]
        name = next((v for k, v in (('k_cg_A', 'cg_A'), ('k_cg_B', 'cg_B'), ('k_mg_res', 'mg_res'), ('k_mg_fused', 'mg_res'), ('k_mg_line', 'mg_line'), ('k_diag', 'diag'), ('k_factor', 'factor')) if k in kname), 'other')

        def tobytes(col):
            v, u = float(r[h.index(col)].replace(',', '')), units[h.index(col)]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        traffic.setdefault(name, []).append(tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'))
# the capture also holds coarse-level launches of the V-cycle kernels: the fine level is the one with the most traffic
json.dump({k: max(v) for k, v in traffic.items()}, open(os.path.join(P, 'ncu_traffic.json'), 'w'), indent=1)
for name in ('bench.json', 'bench_ref.json', 'bench_rline.json', 'ensemble_rline.json', 'ensemble_mgline.json'):
    if os.path.exists(src(name)):
        shutil.copy(src(name), os.path.join(P, f'{tag}_{name}'))
print(open(os.path.join(P, f'{tag}_launches_summary.txt')).read())
print(json.load(open(os.path.join(P, 'ncu_traffic.json'))))
