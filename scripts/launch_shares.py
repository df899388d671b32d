#!/usr/bin/env python
"""Share of a step by kernel from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`,
the format of profiles/*launches*.csv): launches, total ms and share per kernel name (template arguments kept, function
arguments dropped), heaviest first.   python scripts/launch_shares.py profiles/r02_launches_full_scale.csv [--skip N] [--top K]"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--skip', type=int, default=0, help='launches to drop at the start (warm-up step)')
    ap.add_argument('--top', type=int, default=25)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    header = rows[start]
    name_i, metric_i, unit_i, value_i = (header.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value'))
    scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
    ms, count = collections.defaultdict(float), collections.Counter()
    seen = 0
    for r in rows[start + 1:]:
        if len(r) != len(header) or r[metric_i] != 'gpu__time_duration.sum':
            continue
        seen += 1
        if seen <= args.skip:
            continue
        name = re.sub(r'\(.*$', '', r[name_i]).replace('void ', '').strip()
        ms[name] += float(r[value_i].replace(',', '')) * scale.get(r[unit_i], 1.0)
        count[name] += 1
    total = sum(ms.values())
    print(f'{seen - args.skip} launches, {total:.2f} ms in total\n\n| share | launches | ms | kernel |\n|---|---|---|---|')
    for name, t in sorted(ms.items(), key=lambda kv: -kv[1])[:args.top]:
        print(f'| {100 * t / total:.1f} % | {count[name]} | {t:.3f} | `{name[:110]}` |')


if __name__ == '__main__':
    main()
