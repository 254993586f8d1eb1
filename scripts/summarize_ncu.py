"""Turn the ncu artefacts brought back in gpurun_out/ into the text summaries committed under
profiles/ (run here, no GPU needed):

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.txt
    python scripts/summarize_ncu.py full gpurun_out/prof_stream.ncu-rep profiles/r01_lhs_stream.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (KeyError, ValueError):
            continue
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        k = re.sub(r'\(.*', '', row['Kernel Name'])[:78]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, '
                'serialised): compare SHARES, not absolutes\n')
        f.write('# source: %s ; %d launches, total %.1f us\n' % (src, n, tot))
        f.write('%-80s %6s %12s %10s %7s\n' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-80s %6d %12.1f %10.1f %6.1f%%\n' % (k, a[0], a[1], a[1] / a[0],
                                                          100 * a[1] / tot))


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, dst):
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on ; source: %s\n' % rep)
        for r in rows[2:]:
            f.write('\n== %s\n' % r[idx['Kernel Name']])
            for k in KEYS:
                if k in idx:
                    f.write('%-64s %16s %s\n' % (k, r[idx[k]], units[idx[k]]))
            stalls = [(h, r[idx[h]]) for h in hdr if 'smsp__average_warps_issue_stalled' in h]
            stalls.sort(key=lambda t: -float(t[1].replace(',', '') or 0))
            f.write('-- warp stall reasons (warps per issue-active cycle)\n')
            for h, v in stalls[:8]:
                f.write('%-64s %16s\n' % (h.replace('smsp__average_warps_issue_stalled_', '')
                                          .replace('_per_issue_active.ratio', ''), v))
        src = ncu_csv(rep, 'source')
        if len(src) > 2:
            h = src[1]
            ix = {c: i for i, c in enumerate(h)}
            need = max(ix['Instructions Executed'], ix['Source'])
            # several kernels in one report: section / repeated header rows are skipped
            data = [r for r in src[2:] if len(r) > need and
                    (r[ix['Instructions Executed']] or '0').replace(',', '').isdigit()]
            tot = sum(int((r[ix['Instructions Executed']] or '0').replace(',', '')) for r in data)
            ops = collections.Counter()
            for r in data:
                m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
                ops[m.group(2).split('.')[0] if m else '?'] += int((r[ix['Instructions Executed']] or '0').replace(',', ''))
            f.write('\n-- executed warp instructions by opcode (all captured kernels): total %d\n' % tot)
            for op, c in ops.most_common(24):
                f.write('%-10s %12d %5.1f%%\n' % (op, c, 100.0 * c / max(tot, 1)))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
