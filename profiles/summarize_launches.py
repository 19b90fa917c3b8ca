"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        try:
            v = float(x['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        u = x['Metric Unit']
        v = {'nsecond': v / 1e3, 'ns': v / 1e3, 'usecond': v, 'us': v, 'msecond': v * 1e3, 'ms': v * 1e3,
             'second': v * 1e6, 's': v * 1e6}.get(u, v)
        agg.setdefault(x['Kernel Name'], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print('%d launches, %.1f us total (cold-cache, serialised under ncu: compare SHARES)' % (n, tot))
    print('%-90s %6s %10s %9s %6s' % ('kernel', 'n', 'sum_us', 'mean_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print('%-90s %6d %10.1f %9.2f %5.1f%%' % (k[:90], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    mine = {k: v for k, v in agg.items() if 'ojdf' in k or 'conv_tc' in k}
    print('\nown kernels (libojdf.so): %.1f us = %.2f%% of the captured step time' %
          (sum(sum(v) for v in mine.values()), 100 * sum(sum(v) for v in mine.values()) / tot))
    for k, v in mine.items():
        print('  %-88s n=%d mean=%.2f us  [%s]' % (k[:88], len(v), sum(v) / len(v), ' '.join('%.1f' % t for t in v[:8])))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
