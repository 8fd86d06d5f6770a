#!/usr/bin/env python
"""Summarise ncu outputs: `launches` (csv from --metrics gpu__time_duration.sum) or `raw` (.ncu-rep)."""
import collections
import csv
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'launch__grid_size', 'launch__block_size']


def launches(path, steps):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{a[1]/tot*100:6.2f}%  avg {a[1]/a[0]:9.1f} us  x{a[0]:4d}  {k[:110]}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('----', r[idx['Kernel Name']][:120])
        for w in WANT:
            if w in idx:
                print(f"   {w} [{units[idx[w]]}] = {r[idx[w]]}")


def traffic(path, key, out='profiles/ncu_traffic.json'):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first profiled launch -> profiles/ncu_traffic.json[key]
    (bench.py reports it as roofline.traffic)."""
    import json
    import os

    res = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(res.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        tot += float(r[i].replace(',', '')) * mult[units[i]]
    d = json.load(open(out)) if os.path.exists(out) else {}
    d[key] = int(tot)
    d[key + ':source'] = os.path.basename(path)
    json.dump(d, open(out, 'w'), indent=1, sort_keys=True)
    print(key, int(tot))


if __name__ == '__main__':
    if sys.argv[1] == 'traffic':
        traffic(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'launches':
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    else:
        raw(sys.argv[2])
