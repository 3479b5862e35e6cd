#!/usr/bin/env python
"""Benchmark of the GPdoemd design-evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4|C5] [--impl ours|reference]

A "step" = one pass of predict + Taylor marginal + criteria (+ argmax) over one batch of N synthetic
candidate designs for all M rival models.  Prints ONE JSON line (rank 0).  See DESIGN.md section
"Measurement" for the definition of every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the reference arm runs on rank 0 alone and may use every host core; torchrun presets OMP_NUM_THREADS=1
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'candidate designs scored/sec (predict+marginal+criterion, FP64)'
UNIT = 'designs/s'


def workload_name(cfg, name):
    kind = '%s GPs' % cfg['kernel'] if not cfg.get('inducing') else '%s sparse GPs (%d inducing inputs)' % (cfg['kernel'], cfg['inducing'])
    return ('%s: synthetic M=%d models, E=%d outputs, D=%d, P=%d, n_train=%d %s, N=%g candidates per GPU, '
            'Taylor order %d + %s' % (name, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], kind, cfg['N'],
                                      cfg['order'], '/'.join(cfg['criteria'])))


def algorithmic_flops_per_candidate(cfg):
    """per GP n^2 (first order) or (P+2) n^2 (second order) for the triangular contraction,
    the dominant kernel; the evaluation/criteria terms are reported separately in DESIGN.md."""
    G = cfg['M'] * cfg['E']
    n, P = cfg.get('inducing') or cfg['n'], cfg['P']
    return G * (n * n if cfg['order'] == 1 else (P + 2) * n * n)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_run(cfg, seed, n_sample, steps, warmup):
    """The reference's CPU algorithm (oracle port: GPy arithmetic restated in NumPy/LAPACK + the
    reference's Taylor/criteria code) on a bounded sample of the workload, all host threads."""
    from gpdoemd_b200 import synthetic
    from oracle import criteria as ocrit
    from oracle import gpy_kernels
    from oracle import marginal as omarg
    from oracle.model import OracleGPModel, OracleSparseGPModel
    sparse = dict(num_inducing=cfg['inducing'], gp_noise=1e-4) if cfg.get('inducing') else {}
    oms, ds = synthetic.build_models(OracleSparseGPModel if sparse else OracleGPModel, seed, cfg['M'], cfg['E'], cfg['D'],
                                     cfg['P'], cfg['n'], cfg['kernel'], kernel_lookup=gpy_kernels.KERNELS, **sparse)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(7, n_sample, cfg['D'], lo, hi)
    marg = omarg.taylor_first_order if cfg['order'] == 1 else omarg.taylor_second_order

    def step():
        mu = np.zeros((n_sample, cfg['M'], cfg['E']))
        S = np.zeros((n_sample, cfg['M'], cfg['E'], cfg['E']))
        for i, om in enumerate(oms):
            mu[:, i], S[:, i] = marg(om, X)
        out = {}
        for cname in cfg['criteria']:
            dc = ocrit.CRITERIA[cname](mu, S, 0.01, None)
            out[cname] = int(np.argmax(dc))
        return out

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_sample * steps / dt, dt / steps


def blas_threads():
    """threads NumPy's BLAS/LAPACK actually uses (the `cores` of cpu_baseline)"""
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') in ('blas', 'openmp')]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_size(cfg):
    # ~10-30 s of CPU work per run: the oracle does ~700 designs/s on C2 (8 cores) and far less on n=4000
    per = {'C2': 4000, 'C3': 48, 'C4': 64, 'C5': 256, 'S2': 4000}
    return per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='C2')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--candidates', type=float, default=None, help='override N per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-stepwise', action='store_true', help='skip the informational drop-in-functions leg')
    ap.add_argument('--tri-variant', type=int, default=None)
    ap.add_argument('--tri-cps', type=int, default=None)
    ap.add_argument('--tri-trim', type=int, default=None)
    ap.add_argument('--tail-overlap', type=int, default=None)
    ap.add_argument('--gp-build', choices=('host', 'device'), default=None,
                    help='where the exact-GP inference (K, Cholesky, alpha) of the setup runs; default: device for '
                         'n_train >= 2000 (16x faster setup at n = 4000), host LAPACK otherwise; outside the timed region')
    args = ap.parse_args()
    # ONE JSON line on stdout: libraries (NCCL prints its version banner there) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()

    from gpdoemd_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.config])
    if args.candidates:
        cfg['N'] = int(args.candidates)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    seed = 2
    config = {'workload': workload_name(cfg, args.config), 'config_id': args.config,
              'global_candidates': cfg['N'] * world, 'parallelism': 'candidates sharded x%d, GP factors replicated' % world}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        ns = cpu_sample_size(cfg)[args.config]
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        v, sps = cpu_reference_run(cfg, seed, ns, steps, warm)
        line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
                'warmup': warm, 'ms_per_step': sps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
                                 'sample': '%d candidates per step of the same workload (all models, all criteria)' % ns},
                'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return 0

    import torch
    import gpdoemd_b200 as g
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    g.set_default_device(local_rank)
    ctx = g.get_context(local_rank)
    if args.tri_variant is not None:
        ctx.set_option('tri_variant', args.tri_variant)
    if args.tri_cps is not None:
        ctx.set_option('tri_cps', args.tri_cps)
    if args.tail_overlap is not None:
        ctx.set_option('tail_overlap', args.tail_overlap)
    if args.tri_trim is not None:
        ctx.set_option('tri_trim', args.tri_trim)
    if (cfg.get('inducing') or cfg['n']) >= 2000:
        ctx.set_scratch_bytes(64 << 30)

    # ---- problem: identical GP state on every rank, candidates = this rank's contiguous shard ----------
    sparse = dict(num_inducing=cfg['inducing'], gp_noise=1e-4) if cfg.get('inducing') else {}
    models, ds = synthetic.build_models(g.SparseGPModel if sparse else g.GPModel, seed, cfg['M'], cfg['E'], cfg['D'],
                                        cfg['P'], cfg['n'], cfg['kernel'], device=local_rank,
                                        build_on_device=(args.gp_build or ('device' if cfg['n'] >= 2000 else 'host')) == 'device',
                                        **sparse)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    N, D = cfg['N'], cfg['D']
    first_row = rank * N
    X_dev = ctx.dev_alloc(N * D * 8)
    ctx.fill_uniform(X_dev, N, D, seed=20250 + 2, first_row=first_row, lo=lo, hi=hi)
    X_host = ctx.pinned_empty((N, D)) if N * D * 8 <= (4 << 30) else None
    if X_host is not None:
        X_host[:] = ctx.dev_download(X_dev, (N, D))
    crits = list(cfg['criteria'])
    noise, order = 0.01, cfg['order']

    def step_resident():
        return g.score_designs_multi(models, None, crits, order=order, noise_var=noise, return_dc=False,
                                     idx_offset=first_row, x_dev=X_dev, N=N)

    dc_host = ctx.pinned_empty((len(crits), N)) if X_host is not None else None     # page-locked result buffer

    def step_e2e():
        return g.score_designs_multi(models, X_host, crits, order=order, noise_var=noise, return_dc=True,
                                     idx_offset=first_row, dc_out=dc_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile):
        """K steps, L2 flushed before each, device time per step from CUDA events on the library stream."""
        total = 0.0
        res = None
        if profile:
            ctx.profile_enable(True)
        for _ in range(steps):
            ctx.flush_l2()
            ctx.timer_start()
            res = fn()
            total += ctx.timer_stop()
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile_enable(False)
        return total, res, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = step_resident()
    barrier()
    l0 = ctx.launch_count()
    t_ms, res, prof = timed(step_resident, args.steps, True)
    launches = ctx.launch_count() - l0
    barrier()

    e2e_ms = None
    if X_host is not None:
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_dev_ms, res_e2e, _ = timed(step_e2e, args.steps, False)
        barrier()
        e2e_ms = e2e_dev_ms
        for c in crits:
            assert res_e2e[c][0] == res[c][0], 'resident and end-to-end paths disagree on the argmax'

    # the reference's own call pattern (demos/case_study.py:284-293) through the drop-in functions: one
    # taylor_* call per model (host arrays in and out), then one call per criterion -- informational
    stepwise_ms = None
    if X_host is not None and N <= 2000000 and not args.no_stepwise:
        marg = g.taylor_first_order if order == 1 else g.taylor_second_order
        E, Mn = cfg['E'], cfg['M']

        def step_stepwise():
            mu = np.empty((N, Mn, E))
            S = np.empty((N, Mn, E, E))
            for i, mo in enumerate(models):
                mu[:, i], S[:, i] = marg(mo, X_host)
            return {c: int(np.argmax(getattr(g, c)(mu, S, noise, None))) for c in crits}
        r_sw = step_stepwise()
        barrier()
        t0 = time.perf_counter()
        nsw = max(1, min(3, args.steps))
        for _ in range(nsw):
            r_sw = step_stepwise()
        stepwise_ms = (time.perf_counter() - t0) * 1e3 / nsw
        for c in crits:
            assert r_sw[c] + first_row == res[c][0], 'drop-in functions and fused path disagree on the argmax'
    clocks = sampler.stop()      # sampled over warm-up + timed steps + end-to-end steps (all under load)

    # ---- multi-GPU: the only collective is the argmax all-gather; time = max over ranks -----------------
    best = {c: (res[c][1], res[c][0]) for c in crits}
    if dist is not None:
        sc = g.ShardedScorer()
        torch.cuda.set_device(local_rank)     # the library and torch share the CUDA runtime's current device
        t = torch.tensor([t_ms, e2e_ms if e2e_ms is not None else 0.0], dtype=torch.float64,
                         device=torch.device('cuda', local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, e2e_max = float(t[0]), float(t[1])
        e2e_ms = e2e_max if e2e_ms is not None else None
        best = {c: sc.argmax_allgather(res[c][1], res[c][0]) for c in crits}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak = ctx.probe_fp64_peak()
    ms_per_step = t_ms / args.steps
    value = N * world / (ms_per_step * 1e-3)
    tri = prof['tri']
    tri_ms = tri['ms'] / max(1, tri['launches'])
    achieved = tri['flops'] / max(tri['launches'], 1) / (tri_ms * 1e-3) / 1e12 if tri['launches'] else None
    # with "tail_overlap" the step has TWO tri_kernel launches: the main one over complete waves (family 'tri', the
    # dominant kernel reported here) and a small one over the head tiles ('tri_head') that runs concurrently with
    # the evaluation kernel on a second stream -- the overlapped families' times add up to more than the step
    tri_step_ms = tri['ms'] / args.steps
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))
        if args.config in tj and not args.candidates:
            traffic = tj[args.config]['dram_bytes_read'] + tj[args.config]['dram_bytes_write']
    except Exception:
        pass
    roofline = {'bound': 'tensor', 'kernel': 'tri_kernel (FP64 DMMA triangular contraction V = L^-1 K*)',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if achieved else None,
                'traffic': traffic,
                'peak_source': 'FP64 mma.sync m8n8k4 register-resident probe run in this process (MEASURED_PEAKS.json '
                               'has no FP64 entry; cuBLAS DGEMM on this pool: 36.1 TFLOP/s, profiles/r01_fp64_peak.jsonl)',
                'algorithmic_flops_per_launch': tri['flops'] / max(tri['launches'], 1),
                'launch_ms': tri_ms, 'share_of_step': tri_step_ms / ms_per_step,
                'pipeline_frac': algorithmic_flops_per_candidate(cfg) * N / (ms_per_step * 1e-3) / 1e12 / peak,
                'kernel_ms_per_step': {k: v['ms'] / args.steps for k, v in prof.items() if v['launches']}}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(config, l2='flushed before every timed step (256 MiB write); panels working set >> L2',
                           timing='CUDA events on the library stream per step, summed; max over ranks'),
            'roofline': roofline, 'gpu_launches': launches, 'clocks': clocks,
            'argmax': {c: int(best[c][1]) for c in crits}}
    if e2e_ms is not None:
        nk = len(crits)
        line['e2e'] = {'value': N * world / (e2e_ms / args.steps * 1e-3), 'unit': UNIT,
                       'h2d_bytes_per_step': world * N * D * 8, 'd2h_bytes_per_step': world * nk * (N * 8 + 16),
                       'api': 'gpdoemd_b200.score_designs_multi(models, X_host_pinned, criteria, dc_out=pinned) -> dc arrays + argmax'}
    if stepwise_ms is not None:
        line['e2e_stepwise'] = {'value': N / (stepwise_ms * 1e-3), 'unit': UNIT, 'per_rank': True,
                                'api': 'taylor_*_order(model, X) per model + HR/BH/...(mu, S, noise) per criterion, '
                                       'NumPy arrays in and out, host wall clock'}
    if not args.no_cpu_baseline and world >= 1:
        ns = cpu_sample_size(cfg)[args.config]
        v, sps = cpu_reference_run(cfg, seed, ns, 1, 0)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
                                'sample': '%d candidates of the same workload, 1 step, NumPy/LAPACK threads = all cores' % ns}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
