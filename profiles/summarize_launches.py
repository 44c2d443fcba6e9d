"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel and grid."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1000, 'us': v, 'ms': v * 1000}.get(row['Metric Unit'], v)
        key = (name, row['Grid Size'])
        tot[key] += v
        cnt[key] += 1
    T = sum(tot.values())
    print("total %.1f us over %d launches (cold-cache, serialised: compare SHARES)" % (T, sum(cnt.values())))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print("%-46s %-16s n=%4d avg=%8.1f us share=%5.1f%%" % (k[0][:46], k[1], cnt[k], v / cnt[k], 100 * v / T))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
