#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel key metrics + top SASS stall lines.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, collections, io, re
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_active.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    short = name.split('(')[0][-60:]
    if pat and not re.search(pat, name): continue
    if short in seen: continue
    seen.add(short)
    print("====", short)
    for w in want:
        if w in idx: print("   %-62s %s %s" % (w, r[idx[w]], units[idx[w]]))
    st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print("   stall %-40s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
