"""Text summary of an `ncu --set full` report: one block per launch with the metrics DESIGN.md / bench.py quote.
Usage: python profiles/summarize_full.py report.ncu-rep > profiles/<name>_details.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.max', 'SM cycles'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2 -> SM bytes'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts (LSU)'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print('%s: %d launches' % (path, len(data)))
    for r in data:
        print('\n== %s  grid %s  block %s' % (r[idx['Kernel Name']][:100], r[idx['Grid Size']], r[idx['Block Size']]))
        for key, label in WANT:
            if key in idx:
                print('   %-34s %16s %s' % (label, r[idx[key]], units[idx[key]]))


if __name__ == '__main__':
    main(sys.argv[1])
