#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` (+ optional `--page source --csv`) exports: per kernel the time,
DRAM traffic, throughputs, issue utilisation, occupancy, top stall reasons and — with a source page —
the dynamic instruction mix.  Usage: python tools/ncu_summary.py raw.csv [src.csv]"""
import collections
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(d['Kernel Name'][:100])
        for w in WANT:
            if w in d:
                print('   %-62s %s %s' % (w, d[w], u.get(w, '')))
        st = []
        for k, v in d.items():
            if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
                try:
                    st.append((float(v), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('   stalls (warps per issue):', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:5]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        kern, data, i = None, {}, 0
        while i < len(rows):
            r = rows[i]
            if len(r) >= 2 and r[0] == 'Kernel Name':
                kern = r[1][:80] + '#%d' % len(data)
                data[kern] = (rows[i + 1], [])
                i += 2
                continue
            if kern and len(r) == len(data[kern][0]):
                data[kern][1].append(r)
            i += 1
        for k, (h, rs) in data.items():
            ia, isrc = h.index('Instructions Executed'), h.index('Source')
            tot = sum(int(r[ia]) for r in rs)
            warps = max(1, max(int(r[ia]) for r in rs[:5]))
            mix = collections.Counter()
            for r in rs:
                t = r[isrc].strip().split()
                op = t[1] if t[0].startswith('@') else t[0]
                mix[op.split('.')[0]] += int(r[ia])
            print(k)
            print('   static %d, dynamic %.1fM warp-instr, %.0f per warp' % (len(rs), tot / 1e6, tot / warps))
            print('   ' + ' '.join('%s:%.0f' % (o, c / warps) for o, c in mix.most_common(24)))


if __name__ == '__main__':
    main()
