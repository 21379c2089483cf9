"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python tools/ncu_lines.py dump.csv n_units [top]"""
import csv
import sys


def main():
    path, units = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    tot_s = tot_i = 0
    lines = []
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Name':
            cur = r[1].split('/')[-1]
        if len(r) >= 8 and r[0].isdigit():
            try:
                s, n = int(r[4]), int(r[7])
            except ValueError:
                continue
            lines.append((cur, int(r[0]), r[1].strip()[:100], s, n))
            tot_s += s
            tot_i += n
    print('total samples %d, warp instructions %d (%.1f per unit)' % (tot_s, tot_i, tot_i / units))
    for f, l, src, s, n in sorted(lines, key=lambda x: -x[3])[:top]:
        print('%-14s %4d %6.2f%% smp %8.1f inst/unit  %s' % (f, l, 100.0 * s / tot_s, n / units, src))


if __name__ == "__main__":
    main()
