#!/usr/bin/env python
"""bench.py -- GP log-lik evals/sec (N=4096, batched chains) on N B200s, with the Cholesky's fraction of FP64 peak.

One "step" = one pass of the hot path over one batch: every chain of this rank evaluates one
log-marginal likelihood  log N(g_c; 0, K(theta_c) + S(theta_c))  (assemble K+S, Cholesky, forward
substitution, quadratic form, log-det; SURVEY 8a rows a2+a4+a5) and, with more than one rank, the
per-chain results are all-gathered once (the "one all-gather of samples per sweep" of SURVEY 8e).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path
  python bench.py --impl reference [...]                            the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` is timed with the inputs resident in HBM; `e2e` goes through
the C-ABI host-buffer call (pinned H2D of x/g/theta and D2H of loglik/info inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ncu-measured DRAM bytes per eval of the update kernel, keyed by (N, panel width); see profiles/
TRAFFIC_BYTES_PER_EVAL = {(4096, 128): 827.98e6}
METRIC = 'GP log-lik evals/sec (N=4096, batched chains)'
UNIT = 'evals/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nobs', dest='n', type=int, default=4096, help='observations per chain (BASELINE: 4096)')
    ap.add_argument('--chains-per-gpu', type=int, default=1024, help='BASELINE config 5: 8192 chains over 8 GPUs')
    ap.add_argument('--cpu-sample', type=int, default=4, help='evals in the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-peaks', action='store_true')
    ap.add_argument('--gemm-cfg', type=int, default=None, help='experiment: DMMA tile kernel variant (0: 8 warps, 1: 16 warps)')
    return ap.parse_args()


def workload_config(n, B, world):
    """`config` of the JSON line: identical for the CUDA arm and the reference arm."""
    return {'workload': 'BASELINE config 5 (8xB200 chain-parallel: N=4096, 8192 chains sharded by GPU): N=%d, %d chains '
                        'per GPU, one log-lik eval per chain per step, SE+noise kernel on a unit-spaced 1-D grid '
                        '(IH45-shaped)' % (n, B),
            'n': n, 'chains_per_gpu': B, 'evals_per_step': world * B, 'kernel': 'SE iso + noise'}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 200 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------ CPU (oracle)
def cpu_unit_evals_per_s(n, sample, form):
    """Time the oracle's restatement of the unit on the host cores (OpenBLAS threads = all cores)."""
    from oracle import sds_oracle as so
    import gpmc_b200 as gp
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
    G, H = gp.synthetic.loglik_batch(sample, n)
    so.loglik_unit(x[:256], G[0, :256], H[0], form=form)        # warm the BLAS threads
    t0 = time.perf_counter()
    vals = [so.loglik_unit(x, G[i], H[i], form=form) for i in range(sample)]
    dt = time.perf_counter() - t0
    return sample / dt, dt, vals


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure
    Python with its kcGP primitives missing, so the arm is the oracle port (oracle/sds_oracle.py),
    in the form the reference writes it: dense inv for the quadratic form (sliceSample.py:147)."""
    if rank != 0:
        return
    cores = os.cpu_count()
    sample = max(1, min(2, 12 // max(1, args.steps)))      # ~5 s per eval at N=4096 on 8 cores
    for _ in range(min(args.warmup, 1)):
        cpu_unit_evals_per_s(args.n, 1, 'inv')
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        cpu_unit_evals_per_s(args.n, sample, 'inv')
        total += sample
    dt = time.perf_counter() - t0
    v = total / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args.n, args.chains_per_gpu, max(1, args.gpus)),
                       reference_sample='each timed step is a bounded sample of that workload: %d evals, one chain after '
                                        'another on the host (the reference has no multi-chain mode, framework.py:68-75), '
                                        'in the form the reference writes: sliceSample.py:136-137,183-190,196,147 '
                                        '(dense inv)' % sample),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d steps x %d evals at N=%d, numpy/scipy OpenBLAS on %d threads' % (args.steps, sample, args.n, cores)},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------- GPU path
def measured_peaks(torch, gp):
    """FP64 peaks measured in this run: cuBLAS DGEMM 8192^3 (the practical dense-FP64 ceiling) and the
    register-resident DMMA / DFMA issue rates of our own probes."""
    out = {}
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device='cuda')
    b = torch.randn((n, n), dtype=torch.float64, device='cuda')
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out['cublas_dgemm_8192_tflops'] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    del a, b, c
    torch.cuda.empty_cache()
    out['dmma_issue_tflops'] = gp.ops.fp64_peak('dmma', 8192)[0]
    out['dfma_issue_tflops'] = gp.ops.fp64_peak('dfma', 8192)[0]
    return out


def run_b200(args):
    import torch
    import gpmc_b200 as gp
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG=VERSION (the image default); keep stdout for the JSON line
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if args.gemm_cfg is not None:
        gp.ops.set_tuning(0, args.gemm_cfg)
    n, B = args.n, args.chains_per_gpu
    x_h = np.arange(n, dtype=np.float64).reshape(n, 1)
    # chains are keyed by GLOBAL id, so the job's inputs do not depend on how it is sharded
    G_h, H_h = gp.synthetic.loglik_batch(B, n, first_chain=rank * B)
    x = torch.tensor(x_h).cuda()
    G = torch.tensor(G_h).cuda()
    H = torch.tensor(H_h).cuda()
    gathered = torch.empty((world * B,), dtype=torch.float64, device='cuda') if world > 1 else None

    def step():
        ll, info = gp.ops.loglik_batched(x, G, H)
        if world > 1:
            dist.all_gather_into_tensor(gathered, ll)
        return ll, info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    if rank == 0 and not args.no_peaks:
        peaks = measured_peaks(torch, gp)

    for _ in range(max(1, args.warmup)):                  # at least one: first-touch allocations, lazy module load
        ll, info = step()
    barrier()
    assert int((info != 0).sum().item()) == 0, 'a factorisation failed on the synthetic workload'

    sampler = ClockSampler(local)
    sampler.start()
    gp.ops.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ll, info = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    prof = gp.ops.profile_read()
    gp.ops.profile(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned H2D + D2H every step)
    e2e = None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 2)):
            gp.ops.loglik_host(x_h, G_h, H_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ll_h, info_h = gp.ops.loglik_host(x_h, G_h, H_h)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': world * B * args.steps / dt, 'unit': UNIT,
               'h2d_bytes_per_step': int(x_h.nbytes + G_h.nbytes + H_h.nbytes),
               'd2h_bytes_per_step': int(B * 8 + B * 4)}
        assert np.array_equal(ll_h, ll.cpu().numpy()), 'host-buffer path and device path disagree'

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family: the blocked Cholesky (DMMA update + panel kernels)
    evals_timed = B * args.steps
    chol_flops = n ** 3 / 3.0                                    # algorithmic flops per eval
    chol_ms = prof['gemm_update'][0] + prof['potf2'][0] + prof['panel_trsm'][0]
    gemm_ms, gemm_launches = prof['gemm_update']
    achieved = evals_timed * chol_flops / (chol_ms * 1e-3) / 1e12 if chol_ms > 0 else None
    # executed MACs of the update kernel: full 128-row tiles of every block column
    nb = gp._lib.load().gpmc_panel_width()                          # block-column width of this build (64 or 128)
    nt = (n + nb - 1) // nb
    # DMMA flops the update kernel really issues per factorisation: full 128-row tiles below the diagonal block, 8.5 of
    # the 16 32x32 sub-tiles of the diagonal block (6 skipped, 4 at 10/16), one 8-row fragment for the border row
    exec_flops = sum(2.0 * (j * nb) * nb * (((n - j * nb - nb + 127) // 128 * 128) + 0.53125 * nb + 8) for j in range(1, nt))
    peak = peaks.get('cublas_dgemm_8192_tflops')
    roofline = {
        'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
        'frac': (achieved / peak) if (achieved and peak) else None,
        # DRAM bytes per eval of the dominant kernel (all update launches of one factorisation) from an ncu capture
        # (dram__bytes_read.sum + dram__bytes_write.sum; profiles/), next to the algorithmic bytes (every L row read
        # once per block column + the block column read and written once)
        'traffic': TRAFFIC_BYTES_PER_EVAL.get((n, nb)), 'traffic_unit': 'bytes per eval (update kernel)',
        'traffic_algorithmic': sum((n - j * nb) * (j * nb) * 8 + 2 * (n - j * nb) * nb * 8 for j in range(1, nt)),
        'peak_source': 'measured in this run: cuBLAS DGEMM 8192^3 FP64 (MEASURED_PEAKS.json has no FP64 figure); '
                       'nominal B200 FP64 tensor 40 TFLOP/s',
        'what': 'batched blocked Cholesky (TMA-staged DMMA update kernel + potf2 + panel solve launches), N^3/3 flop per eval '
                'over the summed CUDA-event durations of those launches on their stream',
        'frac_of_nominal_40tf': (achieved / 40.0) if achieved else None,
        'update_kernel': {'ms_total': gemm_ms, 'launches': gemm_launches,
                          'executed_tflops': evals_timed * exec_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None},
        'kernel_ms': {k: v[0] for k, v in prof.items()},
        'assemble': {'bytes_per_eval': 8.0 * n * (n + 64) / 2, 'ms_total': prof['assemble'][0],
                     'achieved_gbs': evals_timed * 8.0 * n * (n + 64) / 2 / (prof['assemble'][0] * 1e-3) / 1e9 if prof['assemble'][0] > 0 else None},
        'measured_fp64': peaks,
    }
    try:
        mp = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        roofline['assemble']['hbm_peak_gbs'] = mp.get('hbm_gbs')
    except Exception:
        roofline['assemble']['hbm_peak_gbs'] = 6650.0

    cpu = None
    if not args.no_cpu_baseline and world == 1:        # the CPU arm is reported at N=1 only
        v_inv, dt_inv, vals = cpu_unit_evals_per_s(n, args.cpu_sample, 'inv')
        v_chol, dt_chol, vals_c = cpu_unit_evals_per_s(n, args.cpu_sample, 'trsv')
        got = ll.cpu().numpy()[:args.cpu_sample]
        relerr = float(np.max(np.abs(got - np.asarray(vals_c)) / np.abs(np.asarray(vals_c))))
        cpu = {'value': v_inv, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '%d evals at N=%d of the same workload (rank 0 chains 0..%d): oracle port of sliceSample.py:136-147 '
                         'as written (dense inv), numpy/scipy OpenBLAS on all host threads' % (args.cpu_sample, n, args.cpu_sample - 1),
               'restated_unit_value': v_chol, 'restated_unit': 'cdist+exp, dpotrf, one solve_triangular, log-diag',
               'gpu_vs_oracle_max_rel_err': relerr}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(n, B, world),
                       l2='per-step working set %d matrices x %.0f MiB >> 126 MB L2 (no flush needed)' % (B, n * n * 8 / 2 ** 20),
                       collective='one all_gather of loglik per step' if world > 1 else 'none (single rank)'),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(sum(v[1] for v in prof.values())),
        'roofline': roofline, 'cpu_baseline': cpu,
    }
    sys.stdout.flush()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def protect_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on stdout
    at NCCL_DEBUG=VERSION/WARN), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to
    a private duplicate of the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    args = parse()
    protect_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
