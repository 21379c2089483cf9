"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.

usage: python tools/ncu_lines.py dump.csv n_units[,n_units2,...] [top] [inst|smp]

A dump of several kernel launches repeats its per-file sections once per launch; give one n_units per launch
(pairs, instances ... the launch processed) to get warp instructions per unit.  Sorted by stall samples (smp,
default) or by executed warp instructions (inst)."""
import csv
import sys


def main():
    path = sys.argv[1]
    units = [float(u) for u in sys.argv[2].split(',')]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    by = sys.argv[4] if len(sys.argv) > 4 else 'smp'
    rows = list(csv.reader(open(path)))
    launches = []
    names = []
    seen = set()
    cur_file = None
    cur_fn = None
    lines = None
    for r in rows:
        if len(r) >= 2 and r[0] in ('File Path', 'File Name'):
            cur_file = r[1].split('/')[-1]
            if lines is None or cur_file in seen:          # the same file again: next launch
                lines = None
        elif len(r) >= 2 and r[0] == 'Function Name':
            if lines is None or r[1] != cur_fn:            # ... or another kernel
                lines = []
                launches.append(lines)
                names.append(r[1].split('(')[0][-60:])
                seen = set()
            cur_fn = r[1]
            seen.add(cur_file)
        elif lines is not None and len(r) >= 8 and r[0].isdigit():
            try:
                s, n = int(r[4]), int(r[7])
            except ValueError:
                continue
            lines.append((cur_file, int(r[0]), r[1].strip()[:100], s, n))
    for k, lines in enumerate(launches):
        u = units[min(k, len(units) - 1)]
        tot_s = sum(x[3] for x in lines) or 1
        tot_i = sum(x[4] for x in lines)
        print('=== launch %d %s: total samples %d, warp instructions %d (%.1f per unit)' % (k, names[k], tot_s, tot_i, tot_i / u))
        key = (lambda x: -x[4]) if by == 'inst' else (lambda x: -x[3])
        for f, l, src, s, n in sorted(lines, key=key)[:top]:
            print('%-18s %4d %6.2f%% smp %8.1f inst/unit  %s' % (f, l, 100.0 * s / tot_s, n / u, src))


if __name__ == "__main__":
    main()
