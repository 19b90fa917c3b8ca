"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
per kernel name: launches, time (sum / mean / share) and, when the DRAM counters were collected in the same pass, the
DRAM bytes per launch (read + write) -- the `traffic` figure of bench.py's roofline objects."""
import collections
import csv
import json
import sys

_US = {'nsecond': 1e-3, 'ns': 1e-3, 'usecond': 1.0, 'us': 1.0, 'msecond': 1e3, 'ms': 1e3, 'second': 1e6, 's': 1e6}
_B = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()          # kernel -> {'t': [us...], 'rd': [bytes...], 'wr': [...]}
    for x in csv.DictReader(lines):
        try:
            v = float(x['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        m, u = x['Metric Name'], x['Metric Unit']
        a = agg.setdefault(x['Kernel Name'], {'t': [], 'rd': [], 'wr': []})
        if m == 'gpu__time_duration.sum':
            a['t'].append(v * _US.get(u, 1.0))
        elif m == 'dram__bytes_read.sum':
            a['rd'].append(v * _B.get(u, 1.0))
        elif m == 'dram__bytes_write.sum':
            a['wr'].append(v * _B.get(u, 1.0))
    return agg


def main(path, top=40, json_out=None):
    agg = load(path)
    tot = sum(sum(a['t']) for a in agg.values())
    n = sum(len(a['t']) for a in agg.values())
    print('%d launches, %.1f us total (cold-cache, serialised under ncu: compare SHARES)' % (n, tot))
    print('%-84s %6s %10s %9s %6s %12s' % ('kernel', 'n', 'sum_us', 'mean_us', 'share', 'dram_MB/launch'))
    summary = {}
    for k, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]['t'])):
        t = a['t']
        if not t:
            continue
        dram = (sum(a['rd']) + sum(a['wr'])) / len(t) if a['rd'] else None
        summary[k] = {'n': len(t), 'sum_us': sum(t), 'mean_us': sum(t) / len(t), 'share': sum(t) / tot,
                      'dram_bytes_per_launch': dram}
    for k, s in list(summary.items())[:top]:
        print('%-84s %6d %10.1f %9.2f %5.1f%% %12s' % (k[:84], s['n'], s['sum_us'], s['mean_us'], 100 * s['share'],
                                                      '-' if s['dram_bytes_per_launch'] is None else '%.3f' % (s['dram_bytes_per_launch'] / 1e6)))
    mine = {k: s for k, s in summary.items() if any(ns in k for ns in ('ojdf::', 'tc::', 'ss::', 'wt::', 'chain::'))}
    own = sum(s['sum_us'] for s in mine.values())
    print('\nown kernels (libojdf.so): %.1f us = %.2f%% of the captured time' % (own, 100 * own / tot))
    if json_out:
        json.dump({'total_us': tot, 'launches': n, 'own_share': own / tot, 'kernels': summary}, open(json_out, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else None)
