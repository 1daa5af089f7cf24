"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel table.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip-torch] > profiles/launches_rNN.md
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip_torch = '--skip-torch' in sys.argv
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for row in rows:
        name = row['Kernel Name']
        if skip_torch and name.startswith('void at::'):
            continue
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        ms = v/1e6 if unit in ('ns', 'nsecond') else v/1e3 if unit in ('us', 'usecond') else v
        agg.setdefault(name, []).append(ms)
    total = sum(sum(v) for v in agg.values())
    print('| launches | avg ms | total ms | share | kernel |')
    print('|---:|---:|---:|---:|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'| {len(v)} | {sum(v)/len(v):.4f} | {sum(v):.3f} | {100*sum(v)/total:.1f}% | `{k[:110]}` |')


if __name__ == '__main__':
    main()
