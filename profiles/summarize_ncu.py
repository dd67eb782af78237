"""Condense ncu output into the small CSVs committed under profiles/.

  python profiles/summarize_ncu.py launches <raw launch csv> <out csv> "<command line>"
  python profiles/summarize_ncu.py full <file.ncu-rep> <out csv> "<note>"
"""
import csv, io, subprocess, sys
from collections import OrderedDict

mode, src, dst, note = sys.argv[1:5]
if mode == 'launches':
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    h = rows[0]
    ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        v = v / 1e3 if r[iu] in ('ns', 'nsecond') else (v * 1e3 if r[iu] in ('ms', 'msecond') else v)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('# %s\n' % note)
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'avg_us', 'share_pct'])
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, c, '%.2f' % (t / c), '%.2f' % (100 * t / tot)])
else:
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
    idx = [h.index(m) for m in want]
    with open(dst, 'w') as f:
        f.write('# %s\n' % note)
        w = csv.writer(f)
        w.writerow(['ID', 'Kernel Name'] + want)
        w.writerow(['', ''] + [units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[h.index('ID')], r[h.index('Kernel Name')]] + [r[i] for i in idx])
