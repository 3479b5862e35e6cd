#!/usr/bin/env python
"""Benchmark of the GPdoemd design-evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4|C5|S2] [--impl ours|reference]

A "step" = one pass of predict + Taylor marginal + criteria (+ argmax, + the NCCL all-gather of the per-rank winners
when N > 1) over one batch of synthetic candidate designs for all M rival models.  Prints ONE JSON line (rank 0).

The headline (`value`, `e2e`, `roofline`) is BASELINE configs[1] (C2) with N candidates per GPU (weak scaling).
The same line carries, nested inside keys the driver keeps whole:
  roofline.configs       C3 / C4 / C5 at a reduced N of >= 40 complete persistent-kernel waves each: throughput,
                         tri_kernel fraction of the FP64 peak, pipeline fraction, parity against the CPU oracle and a
                         small CPU baseline of their own (N = 1 runs only);
  roofline.scaling_sweep BASELINE configs[4] (C5) at a FIXED GLOBAL candidate count cut into contiguous blocks over
                         the N ranks (strong scaling), the all-gather inside every timed step (every N);
  cpu_baseline.parity    the CPU leg's own candidate sample scored by the GPU in the same process.
See DESIGN.md section "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun presets OMP_NUM_THREADS=1 for every rank; the CPU legs run on rank 0 alone and may use every host core
if int(os.environ.get('RANK', '0')) == 0 and (int(os.environ.get('WORLD_SIZE', '1')) > 1 or 'reference' in sys.argv):
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'candidate designs scored/sec (predict+marginal+criterion, FP64)'
UNIT = 'designs/s'
SEED = 2
NUM_SMS = 148
NOISE = 0.01
# reduced candidate counts of the extra configs: tiles x GPs x RHS = a whole number (>= 40) of 148-CTA waves
EXTRA_N = {'C3': 148 * 128, 'C4': 148 * 128, 'C5': 370 * 128}
SWEEP_GLOBAL_N = 1 << 18            # C5 strong-scaling sweep: fixed global candidate count
CPU_SAMPLE = {'C2': 2000, 'C3': 48, 'C4': 128, 'C5': 256, 'S2': 2000}


def workload_name(cfg, name):
    kind = '%s GPs' % cfg['kernel'] if not cfg.get('inducing') else '%s sparse GPs (%d inducing inputs)' % (cfg['kernel'], cfg['inducing'])
    return ('%s: synthetic M=%d models, E=%d outputs, D=%d, P=%d, n_train=%d %s, N=%g candidates per GPU, '
            'Taylor order %d + %s' % (name, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], kind, cfg['N'],
                                      cfg['order'], '/'.join(cfg['criteria'])))


def config_dict(cfg, name, world):
    """Identical in both arms (the driver compares it)."""
    return {'workload': workload_name(cfg, name), 'config_id': name, 'global_candidates': cfg['N'] * world,
            'parallelism': 'candidates sharded x%d (contiguous blocks), GP factors replicated' % world,
            'l2': 'flushed before every timed step (256 MiB write); panels working set >> L2',
            'timing': 'value: CUDA events on the library stream around every step, summed, max over ranks; '
                      'e2e: host perf_counter around the public API call after a device sync, max over ranks'}


def tri_flops_per_candidate(cfg):
    """The dominant kernel's algorithmic work, SURVEY 8(d): n^2 per GP (first order) or (P+2) n^2 (second order)."""
    G = cfg['M'] * cfg['E']
    n, P = cfg.get('inducing') or cfg['n'], cfg['P']
    return G * (n * n if cfg['order'] == 1 else (P + 2) * n * n)


def survey_flops_per_candidate(cfg):
    """SURVEY 8(d) F1 / F2 in full: triangular contraction + mean/derivative contractions + kernel evaluation
    (c_k = 3 D_active + 45 flops per (candidate, training row), one software exp counted as 40)."""
    G = cfg['M'] * cfg['E']
    n, P, D = cfg.get('inducing') or cfg['n'], cfg['P'], cfg['D']
    ck = 3 * D + 45
    if cfg['order'] == 1:
        return G * (n * n + 2 * n * (1 + P) + ck * n)
    return G * ((P + 2) * n * n + 2 * n * (1 + P + P * (P + 1) // 2) * 2 + ck * n)


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


def blas_threads():
    """threads NumPy's BLAS/LAPACK actually uses (the `cores` of cpu_baseline)"""
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') in ('blas', 'openmp')]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


# ---- the CPU leg: the reference's algorithm (oracle port) on a bounded sample --------------------------------------
def oracle_models(cfg, seed, adopt_from=None):
    """Oracle (NumPy/LAPACK) models of the workload.  adopt_from: GPU-side models whose posterior (alpha, L) is shared
    instead of refactorised (n = 4000: one host Cholesky per GP would take minutes for 64 GPs); the hyper-parameters,
    training data and transforms are the same either way."""
    from gpdoemd_b200 import synthetic
    from oracle import gpy_kernels
    from oracle.model import OracleGPModel, OracleSparseGPModel, adopt_exact_model
    if adopt_from is not None:
        oms = [adopt_exact_model(m) for m in adopt_from]
        for om in oms:
            om.want_var_jac = cfg['order'] == 2   # first order: skip the variance Jacobian GPy computes and the reference
        return oms                                # discards (it needs K^-1 = an n^3 dpotri per GP)
    sparse = dict(num_inducing=cfg['inducing'], gp_noise=1e-4) if cfg.get('inducing') else {}
    oms, _ = synthetic.build_models(OracleSparseGPModel if sparse else OracleGPModel, seed, cfg['M'], cfg['E'], cfg['D'],
                                    cfg['P'], cfg['n'], cfg['kernel'], kernel_lookup=gpy_kernels.KERNELS, **sparse)
    return oms


def oracle_step(cfg, oms, X):
    from oracle import criteria as ocrit
    from oracle import marginal as omarg
    marg = omarg.taylor_first_order if cfg['order'] == 1 else omarg.taylor_second_order
    mu = np.zeros((len(X), cfg['M'], cfg['E']))
    S = np.zeros((len(X), cfg['M'], cfg['E'], cfg['E']))
    for i, om in enumerate(oms):
        mu[:, i], S[:, i] = marg(om, X)
    return {c: ocrit.CRITERIA[c](mu, S.copy(), NOISE, None) for c in cfg['criteria']}


def cpu_leg(cfg, oms, X, steps, warmup):
    """-> (designs/s, seconds per step, {criterion: dc}) on the candidates X."""
    dc = None
    for _ in range(warmup):
        dc = oracle_step(cfg, oms, X)
    t0 = time.perf_counter()
    for _ in range(steps):
        dc = oracle_step(cfg, oms, X)
    dt = time.perf_counter() - t0
    return len(X) * steps / dt, dt / steps, dc


def parity_block(dc_gpu, dc_cpu, criteria):
    """GPU criterion values against the CPU oracle on the same candidates: gate 1e-7 relative (north star), argmax."""
    worst, match = 0.0, True
    for c in criteria:
        a, b = np.asarray(dc_gpu[c], dtype=float), np.asarray(dc_cpu[c], dtype=float)
        scale = np.maximum(np.abs(b), 1e-6 * max(float(np.median(np.abs(b))), 1e-300))
        worst = max(worst, float(np.max(np.abs(a - b) / scale)))
        match = match and int(np.argmax(a)) == int(np.argmax(b))
    return {'max_rel_err': worst, 'argmax_match': bool(match), 'n': int(len(next(iter(dc_cpu.values())))),
            'criteria': list(criteria), 'gate': 1e-7, 'ok': bool(worst < 1e-7 and match),
            'against': 'oracle port (oracle/) on the same candidates, criterion values per candidate'}


def sample_rows(N, k, must=()):
    idx = np.unique(np.r_[np.linspace(0, N - 1, max(2, k - len(must))).astype(np.int64), np.asarray(must, dtype=np.int64)])
    return idx


# ---- the GPU side -------------------------------------------------------------------------------------------------
class Problem:
    """Models on this rank's GPU + a resident candidate block + its pinned host copy."""

    def __init__(self, g, ctx, cfg, local_rank, gp_build, N, first_row, want_host=True):
        from gpdoemd_b200 import synthetic
        self.g, self.ctx, self.cfg, self.N, self.first_row = g, ctx, cfg, int(N), int(first_row)
        sparse = dict(num_inducing=cfg['inducing'], gp_noise=1e-4) if cfg.get('inducing') else {}
        on_dev = (gp_build or ('device' if cfg['n'] >= 2000 else 'host')) == 'device'
        self.models, ds = synthetic.build_models(g.SparseGPModel if sparse else g.GPModel, SEED, cfg['M'], cfg['E'], cfg['D'],
                                                 cfg['P'], cfg['n'], cfg['kernel'], device=local_rank,
                                                 build_on_device=on_dev, **sparse)
        for mo in self.models:
            mo.device_handle()            # GP state to the device now (setup, outside every timed region)
        self.lo = np.max([d['xlo'] for d in ds], axis=0)
        self.hi = np.min([d['xhi'] for d in ds], axis=0)
        D = cfg['D']
        self.X_dev = ctx.dev_alloc(max(self.N, 1) * D * 8)
        if self.N:
            ctx.fill_uniform(self.X_dev, self.N, D, seed=20250 + SEED, first_row=self.first_row, lo=self.lo, hi=self.hi)
        self.X_host = self.dc_host = None
        if want_host and self.N and self.N * D * 8 <= (4 << 30):
            self.X_host = ctx.pinned_empty((self.N, D))
            self.X_host[:] = ctx.dev_download(self.X_dev, (self.N, D))
            self.dc_host = ctx.pinned_empty((len(cfg['criteria']), self.N))
        self.crits = list(cfg['criteria'])

    def step_resident(self, scorer=None):
        kw = dict(order=self.cfg['order'], noise_var=NOISE, return_dc=False, idx_offset=self.first_row, x_dev=self.X_dev, N=self.N)
        if scorer is not None:
            return scorer.score_designs_multi(self.models, None, self.crits, **kw)
        return self.g.score_designs_multi(self.models, None, self.crits, **kw)

    def step_e2e(self, scorer=None):
        kw = dict(order=self.cfg['order'], noise_var=NOISE, return_dc=True, idx_offset=self.first_row, dc_out=self.dc_host)
        if scorer is not None:
            return scorer.score_designs_multi(self.models, self.X_host, self.crits, **kw)
        return self.g.score_designs_multi(self.models, self.X_host, self.crits, **kw)

    def close(self):
        self.ctx.dev_free(self.X_dev)
        for m in self.models:
            m._invalidate_device()
        self.models = []


def timed_device(ctx, fn, steps, profile=True):
    """K steps, L2 flushed before each, device time per step from CUDA events on the library stream."""
    total, res = 0.0, None
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


def timed_host(ctx, fn, steps):
    """End-to-end clock: the L2 flush has COMPLETED (device sync) before the host stopwatch starts; the public API call
    returns after its own final stream synchronisation, so perf_counter brackets H2D + kernels + D2H."""
    total, res = 0.0, None
    for _ in range(steps):
        ctx.flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        res = fn()
        total += time.perf_counter() - t0
    return total * 1e3, res


def measure(prob, steps, warmup, scorer, barrier, reduce_max, want_e2e=True):
    """Timed resident steps (+ end-to-end steps) of one problem on this rank -> dict (times are max over ranks)."""
    ctx = prob.ctx
    res = None
    for _ in range(warmup):
        res = prob.step_resident(scorer)
    barrier()
    l0 = ctx.launch_count()
    t_ms, res, prof = timed_device(ctx, lambda: prob.step_resident(scorer), steps)
    launches = ctx.launch_count() - l0
    barrier()
    e2e_ms = 0.0
    if want_e2e and (prob.X_host is not None or prob.N == 0):
        for _ in range(2):
            prob.step_e2e(scorer)
        barrier()
        e2e_ms, res_e2e = timed_host(ctx, lambda: prob.step_e2e(scorer), steps)
        barrier()
        for c in prob.crits:
            assert res_e2e[c][0] == res[c][0], 'resident and end-to-end paths disagree on the argmax'
    nccl_ms = prof['comm']['ms'] if prof else 0.0
    t_ms, e2e_ms, nccl_ms = reduce_max([t_ms, e2e_ms, nccl_ms])
    return {'t_ms': t_ms, 'e2e_ms': e2e_ms if want_e2e else None, 'nccl_ms': nccl_ms, 'prof': prof, 'res': res,
            'launches': launches}


def roofline_of(cfg, N_local, steps, ms_per_step, prof, peak, traffic):
    tri = prof['tri']
    tri_ms = tri['ms'] / max(1, tri['launches'])
    achieved = tri['flops'] / max(tri['launches'], 1) / (tri_ms * 1e-3) / 1e12 if tri['launches'] else None
    # with "tail_overlap" the step has TWO tri_kernel launches: the main one over complete waves (family 'tri', the
    # dominant kernel reported here) and a small one over the head tiles ('tri_head') that runs concurrently with
    # the evaluation kernel on a second stream -- the overlapped families' times add up to more than the step
    return {'bound': 'tensor', 'kernel': 'tri_kernel (FP64 DMMA triangular contraction V = L^-1 K*)',
            'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if achieved else None,
            'traffic': traffic,
            'algorithmic_flops_per_launch': tri['flops'] / max(tri['launches'], 1),
            'launch_ms': tri_ms, 'share_of_step': (tri['ms'] / steps) / ms_per_step,
            'pipeline_frac': tri_flops_per_candidate(cfg) * N_local / (ms_per_step * 1e-3) / 1e12 / peak,
            'pipeline_frac_survey_flops': survey_flops_per_candidate(cfg) * N_local / (ms_per_step * 1e-3) / 1e12 / peak,
            'kernel_ms_per_step': {k: v['ms'] / steps for k, v in prof.items() if v['launches']}}


def hbm_secondary(traffic, launch_ms):
    """achieved HBM GB/s of the dominant launch (ncu DRAM bytes / event-timed launch) against the measured copy
    bandwidth of this pool -- the secondary roofline, expected far below its peak because the kernel is FP64-pipe bound."""
    if not traffic or not launch_ms:
        return None
    peak, src = 6552.6, 'SURVEY 8(d) hbm_gbs (MEASURED_PEAKS.json absent)'
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        src = 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        pass
    ach = traffic / (launch_ms * 1e-3) / 1e9
    return {'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_source': src}


def committed_traffic(name, N):
    """DRAM bytes (read + write) of the dominant launch from the committed ncu capture of the same command, else None."""
    for f in ('r02_traffic.json', 'r01_traffic.json'):
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', f)))
            e = tj.get(name)
            if e and int(e.get('N', 100000 if name == 'C2' else -1)) == int(N):
                return e['dram_bytes_read'] + e['dram_bytes_write'], 'profiles/' + f
        except Exception:
            pass
    return None, None


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
    ap.add_argument('--no-extra', action='store_true', help='headline only: skip the C3/C4/C5 blocks and the C5 sweep')
    ap.add_argument('--extra-steps', type=int, default=3, help='timed steps of each C3/C4/C5 block and of the sweep')
    ap.add_argument('--sweep-candidates', type=float, default=SWEEP_GLOBAL_N)
    ap.add_argument('--tri-variant', type=int, default=None)
    ap.add_argument('--tri-cps', type=int, default=None)
    ap.add_argument('--tri-trim', type=int, default=None)
    ap.add_argument('--tail-overlap', type=int, default=None)
    ap.add_argument('--option', action='append', default=[], help='name=value, passed to gpd_set_option')
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
    config = config_dict(cfg, args.config, world)
    warmup = max(args.warmup, 3)

    if args.impl == 'reference':
        # the reference's own CPU implementation of the path (oracle port: GPy cannot be installed here), all host
        # threads, rank 0 alone; every step scores a bounded sample of the workload; K and W are honoured
        if rank != 0:
            return 0
        ns = CPU_SAMPLE[args.config]
        oms = oracle_models(cfg, SEED)
        lo = np.max([om.xmin for om in oms], axis=0)
        hi = np.min([om.xmax for om in oms], axis=0)
        X = synthetic.candidates(7, ns, cfg['D'], lo, hi)
        v, sps, _ = cpu_leg(cfg, oms, X, args.steps, args.warmup)
        line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': sps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
                                 'sample': '%d candidates per step of the same workload (all models, all criteria), '
                                           'NumPy/LAPACK threads = %d' % (ns, blas_threads())},
                'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return 0

    import torch
    import gpdoemd_b200 as g
    torch.cuda.set_device(local_rank)
    dist = scorer = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    g.set_default_device(local_rank)
    ctx = g.get_context(local_rank)
    if world > 1:
        scorer = g.ShardedScorer(device=local_rank)       # the library's own NCCL communicator (gpd_comm_init)
        torch.cuda.set_device(local_rank)
    opts = dict(o.split('=') for o in args.option)
    for k, v in (('tri_variant', args.tri_variant), ('tri_cps', args.tri_cps), ('tail_overlap', args.tail_overlap),
                 ('tri_trim', args.tri_trim)):
        if v is not None:
            opts[k] = v
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    big = (cfg.get('inducing') or cfg['n']) >= 2000
    if big:
        ctx.set_scratch_bytes(64 << 30)

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def reduce_max(vals):
        if dist is None:
            return [float(v) for v in vals]
        torch.cuda.set_device(local_rank)     # the library and torch share the CUDA runtime's current device
        t = torch.tensor(vals, dtype=torch.float64, device=torch.device('cuda', local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    # ---- headline: this rank's contiguous block of N candidates (weak scaling), collective inside every step ----------
    N, D = cfg['N'], cfg['D']
    prob = Problem(g, ctx, cfg, local_rank, args.gp_build, N, rank * N)
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure(prob, args.steps, warmup, scorer, barrier, reduce_max)
    res, prof = m['res'], m['prof']
    crits = prob.crits

    # ---- BASELINE configs[4]: C5 strong-scaling sweep, fixed global N, all-gather inside the step (every rank) ---------
    sweep = None
    if not args.no_extra and args.config == 'C2' and not args.candidates:
        c5 = dict(synthetic.CONFIGS['C5'])
        Ng = int(args.sweep_candidates)
        lo_i, hi_i = g.scoring.shard_bounds(Ng, rank, world)
        ctx.set_scratch_bytes(64 << 30)
        p5 = Problem(g, ctx, c5, local_rank, args.gp_build, hi_i - lo_i, lo_i, want_host=False)
        m5 = measure(p5, args.extra_steps, 1, scorer, barrier, reduce_max, want_e2e=False)
        ms5 = m5['t_ms'] / args.extra_steps
        tiles = -(-(hi_i - lo_i) // 128)
        sweep = {'config_id': 'C5', 'workload': 'BASELINE configs[4]: M=4, E=4, D=3, P=4, n_train=4000 Matern52, Taylor '
                                                'order 1 + BH; N=1e8 is cut to a fixed GLOBAL count so that every N finishes in seconds',
                 'scaling': 'strong', 'global_candidates': Ng, 'n_gpus': world, 'candidates_per_rank': hi_i - lo_i,
                 'value': Ng / (ms5 * 1e-3), 'unit': UNIT, 'ms_per_step': ms5, 'steps': args.extra_steps,
                 'nccl_ms_per_step': m5['nccl_ms'] / args.extra_steps,
                 'collective': ('gpd_score_sharded: pack + ncclAllGather(%d x 16 B) + fold on the library stream, inside every '
                                'timed step' % world) if world > 1 else 'none (one rank)',
                 'waves_per_rank': tiles * c5['M'] * c5['E'] / NUM_SMS,
                 'tri_ms_per_step': m5['prof']['tri']['ms'] / args.extra_steps,
                 'argmax': {c: int(m5['res'][c][0]) for c in p5.crits}}
        p5.close()
        if not big:
            ctx.set_scratch_bytes(6 << 30)
    clocks = sampler.stop()      # sampled over warm-up + timed steps + end-to-end steps + the sweep (all under load)

    if rank != 0:
        if scorer is not None:
            scorer.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rank 0 from here on: derived numbers, CPU legs, parity, the single-GPU extra configs ----------------------------
    peak = ctx.probe_fp64_peak()
    ms_per_step = m['t_ms'] / args.steps
    value = N * world / (ms_per_step * 1e-3)
    traffic, traffic_src = committed_traffic(args.config, N)
    roofline = roofline_of(cfg, N, args.steps, ms_per_step, prof, peak, traffic)
    roofline['traffic_source'] = traffic_src
    roofline['hbm'] = hbm_secondary(traffic, roofline['launch_ms'])
    roofline['peak_source'] = ('FP64 mma.sync m8n8k4 register-resident probe run in this process (MEASURED_PEAKS.json '
                               'has no FP64 entry; cuBLAS DGEMM on this pool: 36.1 TFLOP/s, profiles/r01_fp64_peak.jsonl)')
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
            'roofline': roofline, 'gpu_launches': m['launches'], 'clocks': clocks,
            'argmax': {c: int(res[c][0]) for c in crits}}
    if world > 1:
        line['nccl_ms_per_step'] = m['nccl_ms'] / args.steps
        # (kept out of `config`, which is identical in both arms)
        line['collective'] = line['roofline']['collective'] = (
            'gpd_score_sharded: pack + ncclAllGather(%d x %d x 16 B) + fold on the library stream, inside every timed step, '
            '%.4f ms per step' % (world, len(crits), m['nccl_ms'] / args.steps))
    if m['e2e_ms']:
        nk = len(crits)
        line['e2e'] = {'value': N * world / (m['e2e_ms'] / args.steps * 1e-3), 'unit': UNIT,
                       'h2d_bytes_per_step': world * N * D * 8, 'd2h_bytes_per_step': world * nk * (N * 8 + 16),
                       'clock': 'host perf_counter around the call, device idle (synchronised) at the start',
                       'api': 'gpdoemd_b200.score_designs_multi(models, X_host_pinned, criteria, dc_out=pinned) -> dc arrays + argmax'
                              + (' through ShardedScorer (gpd_score_sharded)' if world > 1 else '')}

    # the reference's own call pattern (demos/case_study.py:284-293) through the drop-in functions: one
    # taylor_* call per model (host arrays in and out), then one call per criterion -- informational
    if prob.X_host is not None and N <= 2000000 and not args.no_stepwise and world == 1:
        marg = g.taylor_first_order if cfg['order'] == 1 else g.taylor_second_order
        E, Mn = cfg['E'], cfg['M']

        def step_stepwise():
            mu = np.empty((N, Mn, E))
            S = np.empty((N, Mn, E, E))
            for i, mo in enumerate(prob.models):
                mu[:, i], S[:, i] = marg(mo, prob.X_host)
            return {c: int(np.argmax(getattr(g, c)(mu, S, NOISE, None))) for c in crits}
        for _ in range(3):               # steady state: the page-locked result buffers cycle through a small pool
            r_sw = step_stepwise()
        ctx.sync()
        t0 = time.perf_counter()
        nsw = max(1, min(5, args.steps))
        for _ in range(nsw):
            r_sw = step_stepwise()
        stepwise_ms = (time.perf_counter() - t0) * 1e3 / nsw
        for c in crits:
            assert r_sw[c] + prob.first_row == res[c][0], 'drop-in functions and fused path disagree on the argmax'
        line['e2e_stepwise'] = {'value': N / (stepwise_ms * 1e-3), 'unit': UNIT, 'per_rank': True,
                                'api': 'taylor_*_order(model, X) per model + the caller\'s mu[:, i], S[:, i] = ... + '
                                       'HR/BH/...(mu, S, noise) per criterion, NumPy arrays in and out, host wall clock'}

        # the same pattern with ONE call for all models: mu / S come back as read-only arrays that stay resident on
        # the device, the criteria consume the device copy (gpd_marginals + gpd_criterion_resident)
        def step_resident_dropin():
            mu, S = g.taylor_multi(prob.models, prob.X_host, order=cfg['order'])
            return {c: int(np.argmax(getattr(g, c)(mu, S, NOISE, None))) for c in crits}
        for _ in range(3):
            r_rd = step_resident_dropin()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(nsw):
            r_rd = step_resident_dropin()
        rd_ms = (time.perf_counter() - t0) * 1e3 / nsw
        for c in crits:
            assert r_rd[c] + prob.first_row == res[c][0], 'taylor_multi + criteria and the fused path disagree on the argmax'
        line['e2e_stepwise_resident'] = {'value': N / (rd_ms * 1e-3), 'unit': UNIT, 'per_rank': True,
                                         'api': 'mu, S = taylor_multi(models, X) (all models, one pass, arrays stay '
                                                'resident) + HR/BH/...(mu, S, noise) per criterion + np.argmax, host wall clock'}

    def cpu_and_parity(cfgk, name, p, steps_cpu=1):
        """CPU leg on a sample of p's OWN candidates (rank 0's block) + the GPU's criterion values on the same rows."""
        ns = CPU_SAMPLE[name] if world == 1 else min(CPU_SAMPLE[name], 512)
        res_k = p.step_e2e(None)                      # dc of every candidate of this rank's block, host side
        must = [res_k[c][0] - p.first_row for c in p.crits]
        idx = sample_rows(p.N, ns, must)
        Xs = np.ascontiguousarray(p.X_host[idx])
        adopt = (cfgk.get('inducing') or cfgk['n']) >= 2000 and not cfgk.get('inducing')
        oms = oracle_models(cfgk, SEED, adopt_from=p.models if adopt else None)
        v, sps, dc_cpu = cpu_leg(cfgk, oms, Xs, steps_cpu, 1 if adopt and cfgk['order'] == 2 else 0)
        par = parity_block({c: res_k[c][2][idx] for c in p.crits}, dc_cpu, p.crits)
        cb = {'value': v, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
              'sample': '%d of this run\'s own candidates (evenly spaced + the GPU argmax rows), %d step, NumPy/LAPACK threads = %d%s'
                        % (len(idx), steps_cpu, blas_threads(),
                           '; posterior (alpha, L) shared with the GPU models instead of a host Cholesky per GP' +
                           ('; GPy\'s discarded variance Jacobian skipped' if cfgk['order'] == 1 else '') if adopt else ''),
              'parity': par}
        return cb

    if not args.no_cpu_baseline and prob.X_host is not None:
        cb = cpu_and_parity(cfg, args.config, prob)
        line['cpu_baseline'] = cb
        line['parity'] = cb['parity']
    prob.close()

    if sweep is not None:
        line['roofline']['scaling_sweep'] = sweep          # nested in a key the driver keeps whole; `config` stays the
        line['scaling_sweep'] = sweep                      # same dict in both arms

    # ---- C3 / C4 / C5 at a reduced N (single-GPU runs only): driver-visible roofline + parity for the big configs ------
    if not args.no_extra and args.config == 'C2' and not args.candidates and world == 1:
        ctx.set_scratch_bytes(64 << 30)
        extra = {}
        for name in ('C3', 'C4', 'C5'):
            ck = dict(synthetic.CONFIGS[name])
            ck['N'] = EXTRA_N[name]
            t_setup = time.perf_counter()
            pk = Problem(g, ctx, ck, local_rank, args.gp_build, ck['N'], 0)
            ctx.sync()
            t_setup = time.perf_counter() - t_setup
            mk = measure(pk, args.extra_steps, 1, None, barrier, reduce_max)
            msk = mk['t_ms'] / args.extra_steps
            tr, tr_src = committed_traffic(name, ck['N'])
            rk = roofline_of(ck, ck['N'], args.extra_steps, msk, mk['prof'], peak, tr)
            rk['traffic_source'] = tr_src
            rk['hbm'] = hbm_secondary(tr, rk['launch_ms'])
            entry = {'workload': workload_name(ck, name), 'value': ck['N'] / (msk * 1e-3), 'unit': UNIT, 'ms_per_step': msk,
                     'steps': args.extra_steps, 'warmup': 1, 'gpu_launches': mk['launches'], 'roofline': rk,
                     'e2e': {'value': ck['N'] / (mk['e2e_ms'] / args.extra_steps * 1e-3), 'unit': UNIT},
                     'setup_s': t_setup, 'argmax': {c: int(mk['res'][c][0]) for c in pk.crits}}
            if not args.no_cpu_baseline:
                entry['cpu_baseline'] = cpu_and_parity(ck, name, pk)
            pk.close()
            extra[name] = entry
        line['roofline']['configs'] = extra
        line['configs'] = extra
    emit(line)
    if scorer is not None:
        scorer.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
