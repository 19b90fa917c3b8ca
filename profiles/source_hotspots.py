"""Per-CUDA-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/source_hotspots.py <report.ncu-rep> <kernel-name-regex> [top]

Runs `ncu -i ... --page source --print-source cuda,sass --csv` and folds the SASS rows onto the
CUDA line that precedes them."""
import csv
import io
import subprocess
import sys


def main(rep, kernel, top=25):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv',
                          '--kernel-name', 'regex:' + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = None
    agg = {}
    for r in rows:
        if '# Samples' in r and 'Line No' in r:
            hdr = r
            si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
            continue
        if hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit():
            continue
        s = int(r[si]) if r[si].isdigit() else 0
        n = int(r[ii]) if r[ii].isdigit() else 0
        a = agg.setdefault((int(r[0]), r[1]), [0, 0])
        a[0] += s
        a[1] += n
    tot = sum(v[0] for v in agg.values()) or 1
    print('%s: %d samples' % (kernel, tot))
    for (ln, src), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%% %9d inst  L%-5s %s' % (100.0 * s / tot, n, ln, src.strip()[:120]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
