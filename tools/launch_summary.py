"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', ''))) for x in csv.DictReader(lines)]
    agg = collections.OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n[:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    print(f'{len(rows)} launches, total {tot / 1e3:.1f} us')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{t / 1e3:10.1f} us {c:5d}x {t / c / 1e3:8.1f} us/launch {100 * t / tot:5.1f}%  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
