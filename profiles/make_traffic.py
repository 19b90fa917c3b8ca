"""DRAM traffic per frame of each roofline object of bench.py, from the ncu launch list of the bench step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... bench.py --ncu-range`):
  extract   = the ray-record and gather kernels (ojdf::extract_kernel / rays)
  integrate = count + offsets + scatter + rank + apply (+ reset)
  fusionnet = every libojdf launch between the gather and the apply kernel of a step that is not one of the above
              (convolutions, chains, pools, the pooled-branch bias kernels on the side stream)
Writes profiles/r2_traffic.json (read by bench.py for the `traffic` field of its roofline objects)."""
import csv
import json
import sys

INTEGRATE = ('count_kernel', 'offsets_kernel', 'scatter_kernel', 'rank_kernel', 'apply_kernel', 'reset_ctrl_kernel', 'sort_long')
EXTRACT = ('extract_kernel', 'ray_setup_kernel', 'rays_kernel')


def main(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    launches = {}
    for x in csv.DictReader(lines):
        d = launches.setdefault(int(x['ID']), {'name': x['Kernel Name'], 't': 0.0, 'b': 0.0})
        v = float(x['Metric Value'].replace(',', ''))
        if x['Metric Name'] == 'gpu__time_duration.sum':
            d['t'] = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(x['Metric Unit'], 1.0)
        elif x['Metric Name'].startswith('dram__bytes'):
            d['b'] += v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(x['Metric Unit'], 1.0)
    seq = [launches[k] for k in sorted(launches)]
    steps = sum(1 for l in seq if 'apply_kernel' in l['name'])
    acc = {'extract': [0.0, 0.0, 0], 'integrate': [0.0, 0.0, 0], 'fusionnet': [0.0, 0.0, 0], 'adapnet': [0.0, 0.0, 0]}
    in_fusion = False
    n_extract = 0
    for l in seq:
        n = l['name']
        own = 'ojdf::' in n or 'tc::' in n or 'ss::' in n or 'wt::' in n or 'chain::' in n
        if any(k in n for k in INTEGRATE):
            key = 'integrate'
            if 'apply_kernel' in n:
                in_fusion = False
        elif any(k in n for k in EXTRACT):
            key = 'extract'
            if 'extract_kernel' in n:                               # two launches per frame: per-ray records (before AdapNet++),
                n_extract += 1                                      # then the gather (right before FusionNet)
                in_fusion = n_extract % 2 == 0
        elif own:
            key = 'fusionnet' if in_fusion else 'adapnet'
        else:
            continue
        acc[key][0] += l['b']; acc[key][1] += l['t']; acc[key][2] += 1
    res = {k: {'dram_bytes_per_frame': v[0] / steps, 'ncu_us_per_frame': v[1] / steps, 'launches_per_frame': v[2] / steps} for k, v in acc.items()}
    res['steps_captured'] = steps
    res['source'] = path
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'profiles/r2_traffic.json')
