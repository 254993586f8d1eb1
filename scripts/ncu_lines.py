"""Per-source-line share of executed warp instructions from an ncu report (source page):
   python scripts/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ''
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'cuda,sass'], capture_output=True, text=True).stdout
    fn = fp = None
    data = {}
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == 'File Path':
            fp = r[1].split('/')[-1]
        elif len(r) == 2 and r[0] == 'Function Name':
            fn = r[1][:60]
        elif len(r) > 8 and r[0].isdigit() and fn:
            try:
                data.setdefault(fn, []).append((int(r[7].replace(',', '')), fp, int(r[0]),
                                                r[1][:100]))
            except ValueError:
                pass
    for k, v in data.items():
        if want not in k:
            continue
        tot = sum(a for a, *_ in v)
        print('==', k, 'executed warp instructions:', tot)
        for a, f, l, s in sorted(v, reverse=True)[:top]:
            print('%5.1f%% %s:%d %s' % (100 * a / max(tot, 1), f, l, s))


if __name__ == '__main__':
    main()
