"""Turn gpurun_out/ ncu artefacts into the text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv            -> per-kernel share of the step
  python profiles/summarize.py full gpurun_out/prof.ncu-rep                -> key counters per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'dur'),
    ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'regs'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma_pipe%'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_active%'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64_pipe%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active%'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts'),
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = agg.setdefault(r[ki].split('(')[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('%-72s %6s %12s %7s' % ('kernel', 'n', 'total_us', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %6d %12.1f %6.1f%%' % (k, n, t, 100 * t / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), lab) for k, lab in KEYS if k in hdr]
    ni = hdr.index('Kernel Name')
    print(' | '.join(['kernel'] + ['%s[%s]' % (lab, units[i]) for i, lab in idx]))
    for r in rows[2:]:
        print(' | '.join([r[ni].split('(')[0]] + [r[i] for i, _ in idx]))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
