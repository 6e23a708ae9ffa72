"""Read an `ncu --set full` capture of tools/ncu_workload.py and write the summaries bench.py and the judge use:

    python tools/ncu_traffic.py gpurun_out/r02_full.ncu-rep --nobs 4096 --chains 16 --tag r02a

  profiles/<tag>_ncu_kernels.txt   one row per kernel class: launches, total time, DRAM bytes, DRAM GB/s, pipe utilisation
  profiles/r02_traffic.json        DRAM bytes per eval of the update kernel (dram__bytes_read.sum + dram__bytes_write.sum
                                   summed over the update launches of ONE factorisation / chains) -> roofline.traffic
"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = {
    'dur': 'gpu__time_duration.sum',
    'rd': 'dram__bytes_read.sum',
    'wr': 'dram__bytes_write.sum',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dmma_pct': 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'fp64_pct': 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'issue_pct': 'sm__inst_issued.avg.pct_of_peak_sustained_active',
    'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'wr_per_s': 'dram__bytes_write.sum.per_second',
    'rd_per_s': 'dram__bytes_read.sum.per_second',
    'regs': 'launch__registers_per_thread',
    'warps_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
         'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0,
         'byte/s': 1.0, 'Kbyte/s': 1e3, 'Mbyte/s': 1e6, 'Gbyte/s': 1e9, 'Tbyte/s': 1e12,
         'byte/second': 1.0, 'Kbyte/second': 1e3, 'Mbyte/second': 1e6, 'Gbyte/second': 1e9, 'Tbyte/second': 1e12}


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return float('nan')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--nobs', type=int, default=4096)
    ap.add_argument('--chains', type=int, default=16)
    ap.add_argument('--panel', type=int, default=128)
    ap.add_argument('--tag', default='r02')
    ap.add_argument('--no-json', action='store_true')
    ap.add_argument('--how', default='ncu --set full --clock-control none')
    args = ap.parse_args()
    if args.rep.endswith('.ncu-rep'):
        out = subprocess.run(['ncu', '-i', args.rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    else:
        out = open(args.rep).read()
    rows = [r for r in csv.reader(out.splitlines()) if r]
    while rows and rows[0][0] != 'ID':           # ncu log files start with ==PROF== lines
        rows.pop(0)
    if 'Metric Name' in rows[0]:
        # long format of `ncu --metrics ... --csv --log-file`: one row per (launch, metric) -> pivot to the wide format
        h = rows[0]
        iid, ik, imn, imu, imv = h.index('ID'), h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
        names, unit_of, per_id, kname = [], {}, collections.OrderedDict(), {}
        for r in rows[1:]:
            if len(r) <= imv:
                continue
            if r[imn] not in unit_of:
                names.append(r[imn]); unit_of[r[imn]] = r[imu]
            per_id.setdefault(r[iid], {})[r[imn]] = r[imv]
            kname[r[iid]] = r[ik]
        hdr = ['ID', 'Kernel Name'] + names
        units = ['', ''] + [unit_of[n] for n in names]
        rows = [hdr, units] + [[i, kname[i]] + [d.get(n, 'nan') for n in names] for i, d in per_id.items()]
    hdr, units = rows[0], rows[1]
    ni = hdr.index('Kernel Name')
    ix = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}

    def val(r, k):
        if k not in ix:
            return float('nan')
        return num(r[ix[k]]) * SCALE.get(units[ix[k]], 1.0)

    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) <= ni:
            continue
        name = r[ni].split('(')[0]
        a = agg.setdefault(name, {'n': 0, 'dur': 0.0, 'rd': 0.0, 'wr': 0.0, 'dmma': [], 'fp64': [], 'issue': [], 'dramp': [], 'regs': None})
        a['n'] += 1
        a['dur'] += val(r, 'dur')
        a['rd'] += val(r, 'rd')
        a['wr'] += val(r, 'wr')
        a['dmma'].append((val(r, 'dmma_pct'), val(r, 'dur')))
        a['fp64'].append((val(r, 'fp64_pct'), val(r, 'dur')))
        a['issue'].append((val(r, 'issue_active_pct') if 'issue_active_pct' in ix else val(r, 'issue_pct'), val(r, 'dur')))
        a['dramp'].append((val(r, 'dram_pct'), val(r, 'dur')))
        a['regs'] = r[ix['regs']] if 'regs' in ix else None

    def wavg(pairs):
        pairs = [(p, w) for p, w in pairs if p == p and w == w]
        tw = sum(w for _, w in pairs)
        return sum(p * w for p, w in pairs) / tw if tw > 0 else float('nan')

    lines = ['%s, one log-lik pass, N=%d, %d matrices (tools/ncu_workload.py); per kernel: launches, '
             'summed duration, DRAM bytes read/written, achieved DRAM GB/s over the summed duration, duration-weighted pipe figures' % (args.how, args.nobs, args.chains),
             '%-44s %5s %10s %12s %12s %9s %8s %8s %8s %8s %5s' % ('kernel', 'n', 'total_us', 'dram_rd_MB', 'dram_wr_MB', 'dram_GB/s', 'dram%', 'dmma%', 'fp64%', 'issue%', 'regs')]
    for name, a in agg.items():
        gbs = (a['rd'] + a['wr']) / a['dur'] / 1e9 if a['dur'] > 0 else float('nan')
        lines.append('%-44s %5d %10.1f %12.2f %12.2f %9.1f %8.1f %8.1f %8.1f %8.1f %5s' % (
            name[:44], a['n'], a['dur'] * 1e6, a['rd'] / 1e6, a['wr'] / 1e6, gbs, wavg(a['dramp']), wavg(a['dmma']), wavg(a['fp64']), wavg(a['issue']), a['regs']))
    text = '\n'.join(lines)
    print(text)
    open(os.path.join(ROOT, 'profiles', '%s_ncu_kernels.txt' % args.tag), 'w').write(text + '\n')
    if args.no_json:
        return
    upd = [a for name, a in agg.items() if 'gemm_dmma_tma_kernel' in name]
    asm = [a for name, a in agg.items() if 'cov_assemble_kernel' in name]
    d = {'source': 'profiles/%s_ncu_kernels.txt (%s of tools/ncu_workload.py --nobs %d --chains %d; '
                   'dram__bytes_read.sum + dram__bytes_write.sum over all update launches of one pass / chains)' % (args.tag, args.how, args.nobs, args.chains),
         'update_kernel': [], 'assemble_kernel': []}
    path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.isfile(path):
        old = json.load(open(path))
        d['update_kernel'] = [r for r in old.get('update_kernel', []) if not (r['n'] == args.nobs and r['panel_width'] == args.panel)]
        d['assemble_kernel'] = [r for r in old.get('assemble_kernel', []) if r['n'] != args.nobs]
    if upd:
        tot = sum(a['rd'] + a['wr'] for a in upd)
        d['update_kernel'].append({'n': args.nobs, 'panel_width': args.panel, 'chains': args.chains, 'launches': sum(a['n'] for a in upd),
                                   'dram_bytes_per_eval': tot / args.chains, 'dmma_pipe_pct': wavg(sum((a['dmma'] for a in upd), []))})
    if asm:
        a = asm[0]
        d['assemble_kernel'].append({'n': args.nobs, 'chains': args.chains, 'dram_bytes_written_per_eval': a['wr'] / args.chains,
                                     'dram_bytes_read_per_eval': a['rd'] / args.chains, 'duration_us': a['dur'] * 1e6,
                                     'dram_write_gbs': a['wr'] / a['dur'] / 1e9 if a['dur'] > 0 else None,
                                     'fp64_pipe_pct': wavg(a['fp64']), 'issue_active_pct': wavg(a['issue']), 'dram_pct_of_peak': wavg(a['dramp'])})
    json.dump(d, open(path, 'w'), indent=1)
    print('wrote', path)


if __name__ == '__main__':
    main()
