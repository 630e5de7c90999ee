#!/usr/bin/env python
"""Headline benchmark: GP posterior + EI evaluations / s (FP64, batched) on N B200s - BASELINE.json metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg4]
  (N > 1: launched by torchrun, one rank per GPU)

--workload cfg3 (default, the configuration BASELINE.json's metric is quoted on, configs[2]: "synthetic batched GP inference:
    65,536 independent GPs, n=64, d=3, 4,096 EI candidates each", per GPU - weak scaling, GPs are independent units).
    A step = one pass of the hot path over the batch: for every (GP, candidate) pair the fused kernel K4 produces the
    posterior mean, variance and the EI score plus the per-GP arg-min (and, for N > 1, the NCCL all-gather of the per-GP
    (score, index) records).
--workload cfg4 (configs[3]: "batched NLML hyperparameter fit: 16,384 GPs x 64 Sobol multistarts, n=128, d=3, SE-ARD", per
    GPU).  A step = K3 over all (GP, start) pairs + per-GP arg-min over the starts (utilities_2.py:166) + the gather.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's OWN functions (oracle/ref_cpu_arm.py: the
unmodified sub_uts/utilities_2.py from the staged baseline/_ref) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
    os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def _baseline_metric():
    """The metric name is BASELINE.json's, verbatim."""
    try:
        with open(os.path.join(REPO, 'BASELINE.json')) as fh:
            return json.load(fh)['metric']
    except Exception:
        return 'GP posterior+EI evals/sec (FP64, batched) at 1/2/4/8 B200; % of roofline'


METRIC = _baseline_metric()
METRIC_CFG4 = 'NLML evals/sec (FP64, batched Cholesky over Sobol multistarts) at 1/2/4/8 B200; % of roofline'
UNIT = 'evals/s'


def flops_eval(n, d):
    """Algorithmic flops per (GP, candidate) evaluation (SURVEY.md par. 8d): dense quadratic form as the reference
    writes it (utilities_2.py:267-268) + cross-covariance + mean + epilogue; transcendentals not counted."""
    return 2 * n * n + n * (3 * d + 6) + 2 * d + 30


def flops_nlml(n, d):
    return n ** 3 / 3 + 2 * n * n + (n * n / 2) * (3 * d + 2) + 4 * n


def k3_kernel_name(n, d):
    """Instance the dispatcher launches for n > 32 (gprto_api.cu: nlml_device -> k3_nlml_lc.cuh)."""
    nt = (n + 15) // 16 * 2
    tiles = nt * (nt + 1) // 2
    slots = (tiles - nt + 6) // 7
    rs = 7 if nt >= 12 else (4 if nt == 10 else slots)
    return 'k3_nlml_lc_kernel<%d,%d,%d>' % (nt, d if d in (2, 3) else 0, rs) if n > 32 else 'k3_nlml_tile_kernel<%d,%d>' % (nt, d if d in (2, 3) else 0)


def fp64_peak():
    """FP64 roofline denominator: MEASURED_PEAKS.json carries HBM and bf16 only, so the FP64 figure is the one
    measured on this pool's B200 by tools/fp64_peaks.cu (DMMA.8x8x4 and DFMA share one pipe: 37.1 / 36.7 TFLOP/s)."""
    path = os.path.join(REPO, 'profiles', 'r01_fp64_peaks.json')
    try:
        with open(path) as fh:
            j = json.load(fh)
        return max(j['dmma884_tflops_sustained'], j['dfma_tflops_sustained']), 'measured (tools/fp64_peaks.cu -> profiles/r01_fp64_peaks.json)'
    except Exception:
        return 37.2, 'nominal (148 SM x 64 FP64 FMA/clk x 1.965 GHz)'


def hbm_peak():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as fh:
            return json.load(fh)['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(summary, kernel_prefix, units):
    """DRAM bytes per launch from the tracked ncu summary, scaled from the captured slice to `units` of this launch; None if
    the summary is missing or belongs to another kernel (so that the number cannot silently go stale)."""
    try:
        with open(os.path.join(REPO, 'profiles', summary)) as fh:
            j = json.load(fh)
        if kernel_prefix not in j['kernel']:
            return None, None
        return j['dram_bytes_per_launch'] * units / j['units_in_capture'], summary
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        load = [c for c, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm      # samples taken under load
        load.sort()
        sm = load
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU baseline
def _port_worker_cfg3(args):
    """Fallback when no copy of the reference is present: the oracle's restatement of the same call structure (kind 'port')."""
    seed, n_gp, n, d, m = args
    from threadpoolctl import threadpool_limits
    from gp_rto_ei_b200 import workloads
    from oracle import gp_oracle as O
    with threadpool_limits(1):
        w = workloads.cfg3_numpy(seed=seed, B=n_gp, n=n, d=d, m=m)
        gps = [O.make_gp(w['X'][b], w['y'][b], w['theta'][b]) for b in range(n_gp)]
        t0 = time.perf_counter()
        acc = 0.0
        for b in range(n_gp):
            _, _, s = O.posterior_acq_scalar(w['cand'][b], gps[b], None, w['f_best'][b], 3)
            acc += float(s.min())
        dt = time.perf_counter() - t0
    return dt, acc


def _port_worker_cfg4(args):
    seed, n_gp, n, d, S = args
    from threadpoolctl import threadpool_limits
    from gp_rto_ei_b200 import workloads
    from oracle import gp_oracle as O
    with threadpool_limits(1):
        w = workloads.cfg4_numpy(seed=seed, B=n_gp, n=n, d=d, S=S)
        nz = [O.normalise(w['X'][b], w['y'][b].reshape(-1, 1)) for b in range(n_gp)]
        t0 = time.perf_counter()
        ok = 0
        for b in range(n_gp):
            v, _ = O.nlml_batch(w['theta'][b], nz[b][4], nz[b][5][:, 0])
            ok += int((v == v).sum())
        dt = time.perf_counter() - t0
    return dt, ok


def cpu_baseline_run(workload, n, d, per_core, inner, cores=None, pool=None):
    """Bounded sample of the same workload on all host cores: `cores` processes x per_core GPs x inner (candidates | starts).
    The timed code is the reference's own (oracle/ref_cpu_arm.py, kind 'reference') when a copy of the reference is present."""
    from oracle import ref_cpu_arm as R
    cores = cores or len(os.sched_getaffinity(0))
    use_ref = R.available()
    if workload == 'cfg3':
        worker = R.cfg3_worker if use_ref else _port_worker_cfg3
        what = ('ITR_GP_RTO.GP_obj_f -> GP_model.GP_inference_np of the UNMODIFIED reference (sub_uts/utilities_2.py:477-507, 235-278), one '
                'candidate per call' if use_ref else 'oracle/gp_oracle.posterior_acq_scalar (restatement of GP_obj_f -> GP_inference_np)')
        unit_name = 'candidates'
    else:
        worker = R.cfg4_worker if use_ref else _port_worker_cfg4
        what = ('GP_model.negative_loglikelihood of the UNMODIFIED reference (sub_uts/utilities_2.py:92-114), one hyper-parameter vector per '
                'call' if use_ref else 'oracle/gp_oracle.nlml (restatement of negative_loglikelihood)')
        unit_name = 'Sobol starts'
    res, wall = R.run_pool(worker, [(1000 + i, per_core, n, d, inner) for i in range(cores)], cores, pool)
    evals = cores * per_core * inner
    t = max(r[0] for r in res)
    return {'value': evals / t, 'unit': UNIT, 'cores': cores, 'kind': 'reference' if use_ref else 'port',
            'sample': '%d processes x %d GP x %d %s (n=%d, d=%d), %s, 1 BLAS thread per process' % (cores, per_core, inner, unit_name, n, d, what),
            'per_core': evals / t / cores, 'pool_wall_s': wall, 'reference_root': os.path.relpath(R.ref_harness.REFERENCE_ROOT, REPO)
            if use_ref and R.ref_harness.REFERENCE_ROOT.startswith(REPO) else (R.ref_harness.REFERENCE_ROOT if use_ref else None)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cfg4 = args.workload == 'cfg4'
    n, d = (128, 3) if cfg4 else (64, 3)
    inner = 64 if cfg4 else 4096
    per_core = 2 if cfg4 else 1
    vals = []
    from oracle import ref_cpu_arm as R
    with R.make_pool() as pool:                 # one pool of worker processes for all steps (imports paid once, in the warm-up)
        for i in range(args.warmup + args.steps):
            r = cpu_baseline_run(args.workload, n, d, per_core, inner, pool=pool)
            if i >= args.warmup:
                vals.append(r)
    v = sum(x['value'] for x in vals) / len(vals)
    evals_per_step = vals[0]['cores'] * per_core * inner
    desc = ('cfg4 (BASELINE.json configs[3]): batched NLML hyper-parameter fit, 16,384 GPs x 64 Sobol multistarts, n=128, d=3' if cfg4 else
            'cfg3 (BASELINE.json configs[2]): synthetic batched GP inference, 65,536 GPs x 4,096 EI candidates, n=64, d=3')
    line = {'metric': METRIC_CFG4 if cfg4 else METRIC, 'value': v, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * evals_per_step / v, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': desc + ' (bounded sample per step, see cpu_baseline.sample)'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': vals[0]['cores'], 'kind': vals[0]['kind'], 'sample': vals[0]['sample'],
                             'reference_root': vals[0]['reference_root']},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ helpers of the GPU arm
class Ctx:
    pass


def make_ctx(args):
    import torch
    from gp_rto_ei_b200 import api, parallel
    c = Ctx()
    c.rank, c.local_rank, c.world = parallel.init()
    if c.world != args.gpus and c.rank == 0:
        print('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, c.world), file=sys.stderr)
    torch.cuda.set_device(c.local_rank)
    c.dev = torch.device('cuda', c.local_rank)
    c.h = api.Handle(c.local_rank)
    return c


def sync_all(c):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if c.world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(c, x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(c, x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def timed_steps(c, step, warmup, steps):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max
    over ranks.  Returns (ms_total over ranks, per-step device ms on this rank, launches)."""
    import torch
    for _ in range(warmup):
        step()
    sync_all(c)
    launches0 = c.h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all(c)
    e0.record()
    for i in range(steps):
        evk[i][0].record()
        step()
        evk[i][1].record()
    e1.record()
    sync_all(c)
    launches = c.h.launch_count - launches0
    ms_total = max_over_ranks(c, e0.elapsed_time(e1))
    k_ms = sum(a.elapsed_time(b) for a, b in evk) / steps
    return ms_total, k_ms, launches


def pcie_peak(c, nbytes=1 << 30):
    """Pinned-memory copy bandwidth of this box, each direction alone and both at once (two streams) - the e2e legs'
    own roofline.  GB/s per direction."""
    import torch
    hb_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hb_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=c.dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=c.dev)
    s1, s2 = torch.cuda.Stream(device=c.dev), torch.cuda.Stream(device=c.dev)

    def run(h2d, d2h, reps=3):
        best = 1e30
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(hb_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    hb_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return nbytes / best / 1e9
    out = {'h2d_alone_gbs': run(True, False), 'd2h_alone_gbs': run(False, True), 'bidirectional_each_way_gbs': run(True, True),
           'how': 'torch pinned <-> device copies of 1 GiB, best of 3, wall clock around synchronize'}
    del hb_in, hb_out, d_in, d_out
    return out


def host_mem_ok(c, need_bytes):
    try:
        import psutil
        return psutil.virtual_memory().available > 2.5 * need_bytes * c.world + (8 << 30)
    except Exception:
        return False


# ------------------------------------------------------------------------------------------ cfg 4 (NLML sweep)
def cfg4_setup(c, B, S, n, d, seed):
    import torch
    from gp_rto_ei_b200 import workloads
    w = workloads.cfg4_torch(c.dev, seed=seed, B=B, n=n, d=d, S=S)
    nz = c.h.normalise(w['X'], w['y'])
    st = dict(w=w, nz=nz, nll=torch.empty(B, S, dtype=torch.float64, device=c.dev), status=torch.empty(B, S, dtype=torch.int32, device=c.dev),
              best_v=torch.empty(B, dtype=torch.float64, device=c.dev), best_i=torch.empty(B, dtype=torch.int64, device=c.dev),
              rec=torch.empty(B, 2, dtype=torch.float64, device=c.dev))
    return st


def cfg4_step(c, st, B, S, n, d, lo, gather=True):
    """K3 over all (GP, start) pairs, per-GP arg-min over the starts (utilities_2.py:166), and the gather of (nll_best, idx).
    gather=False: rank-local call (the parity sample runs on rank 0 alone and must not enter a collective)."""
    from gp_rto_ei_b200 import api, parallel
    lib, hp, P, h = c.h.lib, c.h._h, api._ptr, c.h
    nz, w = st['nz'], st['w']
    api._lib.check(lib.gprto_nlml(hp, B, S, n, d, P(nz['Xn']), P(nz['Yn']), P(w['theta']), 1e-8, P(st['nll']), P(st['status']), h._stream()), 'gprto_nlml')
    api._lib.check(lib.gprto_argmin(hp, B, S, P(st['nll']), P(st['best_v']), P(st['best_i']), h._stream()), 'gprto_argmin')
    if c.world > 1 and gather:
        api._lib.check(lib.gprto_pack_best(hp, B, P(st['best_v']), P(st['best_i']), 0, P(st['rec']), h._stream()), 'gprto_pack_best')
        return parallel.allgather_records(st['rec'])
    return None


def nlml_block(c, B, S, n, d, steps, warmup, label, with_e2e=True):
    """Timed cfg-4 block (used as the headline of --workload cfg4 and as `secondary` of cfg3): dict with value / roofline / e2e."""
    import torch
    st = cfg4_setup(c, B, S, n, d, seed=1 + 1000 * c.rank)
    ms_total, k_ms, launches = timed_steps(c, lambda: cfg4_step(c, st, B, S, n, d, 0), warmup, steps)
    ms = ms_total / steps
    evals = float(B) * S * c.world
    peak, peak_src = fp64_peak()
    F = flops_nlml(n, d)
    tf = float(B) * S * F / (k_ms * 1e-3) / 1e12
    failed = sum_over_ranks(c, float((st['status'] != 0).sum()))
    out = {'metric': METRIC_CFG4, 'workload': label, 'value': evals / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps, 'n_gpus': c.world,
           'gps_per_gpu': B, 'starts': S, 'n': n, 'd': d, 'flops_per_eval': F, 'gpu_launches': launches,
           'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': tf / peak, 'kernel': k3_kernel_name(n, d),
                        'kernel_ms': k_ms, 'peak_source': 'FP64 ' + peak_src},
           'failed_factorisations': int(failed),
           'note_failures': 'corners of the reference hyper-parameter box (sn2 = e^-16, sf2 = e^10) are indefinite in FP64; the reference would raise there (utilities_2.py:107)'}
    tr, src = ncu_traffic('k3_ncu_summary.json', 'k3_nlml_lc_kernel', float(B) * S)
    out['roofline']['traffic'] = tr
    out['roofline']['traffic_source'] = src
    if with_e2e:
        # e2e: theta from pinned host memory every step, (nll_best, idx) per GP back to pinned host memory
        th_h = st['w']['theta'].cpu().pin_memory()
        bv_h = torch.empty(B, dtype=torch.float64).pin_memory()
        bi_h = torch.empty(B, dtype=torch.int64).pin_memory()

        def e2e_step():
            st['w']['theta'].copy_(th_h, non_blocking=True)
            cfg4_step(c, st, B, S, n, d, 0)
            bv_h.copy_(st['best_v'], non_blocking=True)
            bi_h.copy_(st['best_i'], non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        sync_all(c)
        k = max(2, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        out['e2e'] = {'value': evals * k / dt, 'unit': UNIT, 'h2d_bytes_per_step': B * S * (d + 2) * 8, 'd2h_bytes_per_step': B * 16,
                      'sample': 'theta[%d,%d,%d] H2D from pinned memory, gprto_nlml + gprto_argmin, (nll_best, idx) D2H, %d steps' % (B, S, d + 2, k)}
    return out


def run_cfg4(args, c):
    B, S, n, d = args.B4, args.S, args.n4, args.d
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
        time.sleep(0.3)
    blk = nlml_block(c, B, S, n, d, args.steps, args.warmup, 'cfg4')
    clocks = sampler.stop() if c.rank == 0 else None
    if c.rank != 0:
        return
    cpu = None if args.skip_extras else cpu_baseline_run('cfg4', n, d, 2, S)
    # parity sample: 4 GPs x all starts against the oracle (the reference raises where the oracle reports failure)
    from oracle import gp_oracle as O
    import numpy as np
    st = cfg4_setup(c, 4, S, n, d, seed=1 + 1000 * c.rank)
    cfg4_step(c, st, 4, S, n, d, 0, gather=False)
    nll = st['nll'].cpu().numpy(); stat = st['status'].cpu().numpy()
    Xh, yh, th = st['w']['X'].cpu().numpy(), st['w']['y'].cpu().numpy(), st['w']['theta'].cpu().numpy()
    worst, mism, worst_cond_ok = 0.0, 0, 0.0
    gpu_t, ora_t = 0.0, 0.0
    for b in range(4):
        _, _, _, _, Xn, Yn = O.normalise(Xh[b], yh[b].reshape(-1, 1))
        ref, rst = O.nlml_batch(th[b], Xn, Yn[:, 0])
        ok = (rst == 0) & (stat[b] == 0)
        mism += int(((rst == 0) != (stat[b] == 0)).sum())
        rel = np.abs(nll[b] - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(np.max(rel[ok])))
        if b < 2:                               # 80-bit arbiter on every start of two GPs: which side is closer to the exact value
            for s_ in np.flatnonzero(ok):
                tv = O.truth_ld(Xn, Yn[:, 0], th[b][s_], jitter=1e-8)['nll']
                if np.isfinite(tv):
                    gpu_t = max(gpu_t, abs(nll[b][s_] - tv) / max(1.0, abs(tv)))
                    ora_t = max(ora_t, abs(ref[s_] - tv) / max(1.0, abs(tv)))
    line = {'metric': blk['metric'], 'value': blk['value'], 'unit': UNIT, 'n_gpus': c.world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': blk['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'cfg4 (BASELINE.json configs[3]): batched NLML hyper-parameter fit, %d GPs per GPU x %d Sobol multistarts, n=%d, d=%d, '
                                   'SE-ARD; per step: K3 sweep + per-GP arg-min over the starts%s' % (B, S, n, d, ' + NCCL all-gather of (nll_best, idx)' if c.world > 1 else ''),
                       'gps_per_gpu': B, 'starts': S, 'n': n, 'd': d, 'parallelism': 'gp-sharded x%d' % c.world,
                       'l2': 'inputs are 4 KB per GP and stay L2-resident; the kernel is FP64-compute bound (no HBM stream to flush)'},
            'e2e': blk['e2e'], 'gpu_launches': blk['gpu_launches'], 'clocks': clocks, 'roofline': blk['roofline'], 'cpu_baseline': cpu,
            'failed_factorisations': blk['failed_factorisations'],
            'parity': {'sample': '4 GPs x %d starts vs oracle.nlml (restatement of negative_loglikelihood, pinned bit-for-bit to the reference)' % S,
                       'max_rel_err_where_both_succeed': worst, 'failure_flag_mismatches': mism,
                       'vs_long_double_truth': {'gpu': gpu_t, 'oracle': ora_t, 'sample': '2 GPs x every start both sides factorise; 80-bit Cholesky (oracle.truth_ld)'},
                       'note': 'the Sobol box reaches sn2 = e^-16, sf2 = e^10: cond(K) up to ~1e16 there, where two correct algorithms differ by cond*eps; '
                               'the long-double arbiter shows which side carries the error'}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ cfg 3 (posterior + EI)
def parity_block(c, w, gd, mean, var, score, best_i, cand, f_best, npar, n_truth, d, m):
    """Parity of the GPU results on a sample of the bench batch, reported four ways per quantity (SURVEY.md par. 7.3-2):
    pointwise max relative error vs the oracle, error normalised by the natural scale of the quantity, the sample's
    max cond(K), and GPU-vs-long-double-truth beside oracle-vs-long-double-truth (which of the two is closer to the exact value)."""
    import numpy as np
    from scipy.stats import norm as _norm
    from oracle import gp_oracle as O
    Xh, yh, th = w['X'][:npar].cpu().numpy(), w['y'][:npar].cpu().numpy(), w['theta'][:npar].cpu().numpy()
    ch, fh = cand[:npar].cpu().numpy(), f_best[:npar].cpu().numpy()
    mg, vg, sg = mean[:npar].cpu().numpy(), var[:npar].cpu().numpy(), score[:npar].cpu().numpy()
    pw = {'mean': 0.0, 'var': 0.0, 'score': 0.0}
    sc = {'mean': 0.0, 'var': 0.0, 'score': 0.0}
    truth = {'gpu_mean': 0.0, 'oracle_mean': 0.0, 'gpu_var': 0.0, 'oracle_var': 0.0}
    conds, mismatch, compared = [], 0, 0
    tiny = 1e-300
    for b in range(npar):
        gp = O.make_gp(Xh[b], yh[b], th[b])
        conds.append(float(np.linalg.cond(gp['invK'])))
        mv, vv, sv = O.posterior_acq_vectorised(ch[b], gp, None, fh[b], 3)
        sf2 = math.exp(2 * th[b][d])
        sig = np.sqrt(vv); dl = fh[b] - mv
        with np.errstate(divide='ignore', invalid='ignore'):
            z = np.where(sig == 0, 0.0, dl / sig)
        ei_scale = float(np.max(sig * _norm.pdf(z) + np.abs(dl) * _norm.cdf(z)))
        pw['mean'] = max(pw['mean'], float(np.max(np.abs(mg[b] - mv) / np.maximum(np.abs(mv), tiny))))
        pw['var'] = max(pw['var'], float(np.max(np.abs(vg[b] - vv) / np.maximum(np.abs(vv), tiny))))
        pw['score'] = max(pw['score'], float(np.max(np.abs(sg[b] - sv) / np.maximum(np.abs(sv), tiny))))
        sc['mean'] = max(sc['mean'], float(np.max(np.abs(mg[b] - mv)) / np.max(np.abs(mv))))
        sc['var'] = max(sc['var'], float(np.max(np.abs(vg[b] - vv)) / (sf2 * gp['Y_std'] ** 2)))
        sc['score'] = max(sc['score'], float(np.max(np.abs(sg[b] - sv)) / max(ei_scale, tiny)))
        srt = np.sort(sv)
        if srt[1] - srt[0] > 1e-9 * ei_scale:                       # selected candidate: compared where the gap exceeds the tolerance
            compared += 1
            mismatch += int(int(best_i[b]) != int(np.argmin(sv)))
        if b < n_truth:
            xn = (ch[b] - gp['X_mean']) / gp['X_std']
            tr = O.truth_ld(gp['X_norm'], gp['y_norm'], th[b], xn)
            m_true = tr['mean_norm'] * gp['Y_std'] + gp['Y_mean']
            v_true = tr['var_norm'] * gp['Y_std'] ** 2
            ms_, vs_ = float(np.max(np.abs(m_true))), sf2 * gp['Y_std'] ** 2
            truth['gpu_mean'] = max(truth['gpu_mean'], float(np.max(np.abs(mg[b] - m_true))) / ms_)
            truth['oracle_mean'] = max(truth['oracle_mean'], float(np.max(np.abs(mv - m_true))) / ms_)
            truth['gpu_var'] = max(truth['gpu_var'], float(np.max(np.abs(vg[b] - v_true))) / vs_)
            truth['oracle_var'] = max(truth['oracle_var'], float(np.max(np.abs(vv - v_true))) / vs_)
    return {'sample': '%d GPs x %d candidates of the bench batch vs oracle (vectorised NumPy restatement of GP_inference_np + GP_obj_f)' % (npar, m),
            'pointwise_max_rel_err': pw,
            'pointwise_note': 'pointwise relative error is dominated by candidates where the quantity itself cancels (mean crossing zero, EI -> 0 '
                              'in the tail, var -> 0 next to data): |err| / |value| with |value| << scale',
            'scale_normalised_max_err': sc,
            'scale_note': 'mean by max|mean| of the GP, var by the prior variance sf2*y_std^2, score by max(sigma*phi + |Delta|*Phi)',
            'max_cond_K': max(conds), 'median_cond_K': sorted(conds)[len(conds) // 2],
            'vs_long_double_truth': dict(truth, sample='%d GPs; 80-bit Cholesky-form evaluation of the same formulas (oracle.truth_ld); scale-normalised' % n_truth),
            'argmin_compared': compared, 'argmin_mismatch': mismatch,
            'tolerance': 'north_star 1e-9 relative: met scale-normalised; pointwise it is met wherever |value| > cond(K)*eps*scale '
                         '(tests/test_gpu_parity.py asserts |gpu-ref| <= 1e-9|ref| + first-order rounding floor of the reference formula)'}


def run_cfg3(args, c):
    import torch
    import torch.distributed as dist
    from gp_rto_ei_b200 import api, parallel, workloads
    rank, world, dev, h = c.rank, c.world, c.dev, c.h
    n, d = args.n, args.d
    B, m = args.B, args.m
    w = workloads.cfg3_torch(dev, seed=1000 * rank, B=B, n=n, d=d, m=m)
    gd = h.fit_fixed(w['X'], w['y'], w['theta'])            # K2 (outside the timed region: the metric is K4)
    assert int(gd['status'].abs().sum()) == 0
    cand, f_best = w['cand'], w['f_best']
    mean = torch.empty(B, m, dtype=torch.float64, device=dev)
    var = torch.empty_like(mean); score = torch.empty_like(mean)
    best_s = torch.empty(B, dtype=torch.float64, device=dev); best_i = torch.empty(B, dtype=torch.int64, device=dev)
    rec = torch.empty(B, 2, dtype=torch.float64, device=dev)
    lib, hp, P = h.lib, h._h, api._ptr

    def step():
        api._lib.check(lib.gprto_posterior_acq(hp, B, n, d, m, P(gd['Xn']), P(gd['invK']), P(gd['alpha']), P(gd['theta']), P(gd['x_mean']),
                                               P(gd['x_std']), P(gd['y_mean']), P(gd['y_std']), P(cand), None, P(f_best), 3, P(mean), P(var),
                                               P(score), P(best_s), P(best_i), h._stream()), 'gprto_posterior_acq')
        if world > 1:
            api._lib.check(lib.gprto_pack_best(hp, B, P(best_s), P(best_i), 0, P(rec), h._stream()), 'gprto_pack_best')
            return parallel.allgather_records(rec)
        return None

    sampler = ClockSampler(c.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                     # let nvidia-smi come up; its first samples are idle ones
    ms_total, k4_ms, launches = timed_steps(c, step, args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    evals_step = float(B) * m * world
    value = evals_step / (ms_step * 1e-3)

    if args.skip_extras:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': value, 'unit': UNIT, 'ms_per_step': ms_step, 'kernel_ms': k4_ms, 'gpu_launches': launches}))
        return

    # ---- e2e: host buffers through the C-ABI host entry (H2D of candidates, D2H of mean/var/score/best inside the timed region),
    # on the FULL batch when the host has the memory for the pinned buffers (12.9 GB per rank), else on a quarter of it
    need = B * m * (d + 3) * 8
    Be = B if (args.e2e_B <= 0 and host_mem_ok(c, need)) else min(B, args.e2e_B if args.e2e_B > 0 else B // 4)
    pcie = pcie_peak(c)
    cand_h = torch.empty(Be, m, d, dtype=torch.float64).pin_memory()
    cand_h.copy_(cand[:Be])
    outs = dict(mean=torch.empty(Be, m, dtype=torch.float64).pin_memory(), var=torch.empty(Be, m, dtype=torch.float64).pin_memory(),
                score=torch.empty(Be, m, dtype=torch.float64).pin_memory(), best_score=torch.empty(Be, dtype=torch.float64).pin_memory(),
                best_idx=torch.empty(Be, dtype=torch.int64).pin_memory())
    gsub = {k: v[:Be] for k, v in gd.items()}
    fb = f_best[:Be]
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_run(out_dict):
        h.posterior_acq_host(gsub, cand_h, None, fb, 3, out_dict)
        sync_all(c)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h.posterior_acq_host(gsub, cand_h, None, fb, 3, out_dict)       # returns after the last D2H completed
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        return float(Be) * m * world * e2e_steps / dt, dt / e2e_steps
    e2e_value, e2e_s = e2e_run(outs)
    assert torch.equal(outs['score'][:4], score[:4].cpu())
    h2d = Be * m * d * 8
    d2h = Be * m * 3 * 8 + Be * 16
    # variant that returns only what the RTO loop consumes (per-GP best score + index): same H2D, 16 B per GP back
    e2e_best_value, e2e_best_s = e2e_run({k: outs[k] for k in ('best_score', 'best_idx')})
    del cand_h, outs
    # variant with on-device candidates (gprto_posterior_acq_sampled): per GP only (centre, radius) go in (H2D from pinned
    # memory each step) and (best score, best index, best point) come back; full batch of B GPs
    ctr_h = w['X'][:, 0, :].contiguous().cpu().pin_memory()
    rad_h = torch.full((B,), 0.25, dtype=torch.float64).pin_memory()
    ctr_d = torch.empty(B, d, dtype=torch.float64, device=dev); rad_d = torch.empty(B, dtype=torch.float64, device=dev)
    so = dict(best_score=torch.empty(B, dtype=torch.float64, device=dev), best_idx=torch.empty(B, dtype=torch.int64, device=dev),
              best_x=torch.empty(B, d, dtype=torch.float64, device=dev))
    so_h = {k: torch.empty_like(v, device='cpu').pin_memory() for k, v in so.items()}

    def sampled_step(seed):
        ctr_d.copy_(ctr_h, non_blocking=True); rad_d.copy_(rad_h, non_blocking=True)
        h.posterior_acq_sampled(gd, ctr_d, rad_d, m, shape=1, seed=seed, f_best=f_best, acq=3, out=so)
        for k in so:
            so_h[k].copy_(so[k], non_blocking=True)
        torch.cuda.synchronize()
    sampled_step(0)
    sync_all(c)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        sampled_step(i + 1)
    e2e_sampled_value = float(B) * m * world * e2e_steps / max_over_ranks(c, time.perf_counter() - t0)

    # ---- secondary metric on every N: cfg-4 NLML sweep, 1/8 of the per-GPU batch (the rate is size-independent: one CTA per (GP, start))
    nlml = None
    if not args.no_nlml:
        try:
            nlml = nlml_block(c, 2048, 64, 128, d, 3, 1, 'cfg4 slice (1/8 of the 16,384 x 64 sweep per GPU)', with_e2e=False)
        except Exception as exc:          # secondary line must never take the headline down
            nlml = {'error': repr(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    parity = parity_block(c, w, gd, mean, var, score, best_i, cand, f_best, 64, 8, d, m)
    cpu = cpu_baseline_run('cfg3', n, d, 4, 4096)      # ~20-30 s of CPU work in total
    peak, peak_src = fp64_peak()
    F = flops_eval(n, d)
    achieved = (float(B) * m * F) / (k4_ms * 1e-3) / 1e12
    hbm, hbm_src = hbm_peak()
    alg_bytes = 8.0 * (d + 3) + 8.0 * (n * n + n * d + n + 3 * d + 4) / m
    traffic, traffic_src = ncu_traffic('k4_ncu_summary.json', 'k4_dmma_kernel', float(B) * m)
    bidir = pcie['bidirectional_each_way_gbs']
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'cfg3 (BASELINE.json configs[2]): synthetic batched GP inference, %d independent GPs per GPU, n=%d, d=%d, '
                               '%d EI candidates each; outputs mean+var+score per candidate and per-GP arg-min' % (B, n, d, m),
                   'gps_per_gpu': B, 'n': n, 'd': d, 'candidates_per_gp': m, 'acq': 'EI', 'parallelism': 'gp-sharded x%d' % world,
                   'l2': 'inputs per step (%.1f GB candidates + %.1f GB invK) exceed the 126 MB L2; no flush needed' %
                         (B * m * d * 8 / 1e9, B * n * n * 8 / 1e9),
                   'collective': 'NCCL all-gather of per-GP (score, idx) records each step' if world > 1 else 'none (single GPU)'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'sample': 'gprto_posterior_acq_host on %d of %d GPs x %d candidates per GPU (%s), pinned host buffers, %d steps' %
                          (Be, B, m, 'full config' if Be == B else 'host memory bounded', e2e_steps),
                'pcie_gbs_each_way': {'h2d': h2d / e2e_s / 1e9, 'd2h': d2h / e2e_s / 1e9},
                'pcie_peak': pcie, 'pcie_frac': max(h2d, d2h) / e2e_s / 1e9 / bidir,
                'bound': 'PCIe: 24 B in + 24 B out per evaluation; pcie_frac = achieved GB/s of the busier direction / measured bidirectional pinned-copy peak',
                'best_only': {'value': e2e_best_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': Be * 16,
                              'pcie_frac': h2d / e2e_best_s / 1e9 / pcie['h2d_alone_gbs'],
                              'note': 'same call returning only the per-GP arg-min - what the RTO loop consumes (utilities_2.py:991); '
                                      'this is the variant the drop-in ITR_GP_RTO uses for host-generated SLSQP iterates'},
                'sampled_on_device': {'value': e2e_sampled_value, 'unit': UNIT, 'h2d_bytes_per_step': B * (d + 1) * 8, 'd2h_bytes_per_step': B * (d + 2) * 8,
                                      'note': 'gprto_posterior_acq_sampled: the %d trust-region (ball) candidates of each of %d GPs are generated inside K4 from '
                                              '(centre, radius, seed) - the device form of the multistart seeding of utilities_2.py:970-975; returns best score / '
                                              'index / point per GP; not PCIe-bound' % (m, B)}},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                     'traffic_source': traffic_src, 'kernel': 'k4_dmma_kernel<8,3>', 'kernel_ms': k4_ms, 'flops_per_eval': F,
                     'executed_note': 'algorithmic flops (the reference\'s dense quadratic form); the kernel folds invK into a triangular matrix and executes '
                                      '72 of 128 DMMA fragments per tile, so the executed pipe utilisation is the ncu figure (DMMA 46.3 % + FP64 39.1 % = 85 %, profiles/k4_ncu_summary.json)',
                     'peak_source': 'FP64 ' + peak_src + '; FP64 DMMA and DFMA share one pipe on B200, MEASURED_PEAKS.json has no FP64 entry',
                     'hbm_check': {'algorithmic_bytes_per_eval': alg_bytes, 'achieved_gbs': float(B) * m * alg_bytes / (k4_ms * 1e-3) / 1e9,
                                   'peak_gbs': hbm, 'peak_source': hbm_src}},
        'cpu_baseline': cpu,
        'secondary': nlml,
        'parity': parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg4'])
    ap.add_argument('--B', type=int, default=65536, help='cfg3: GPs per GPU')
    ap.add_argument('--m', type=int, default=4096)
    ap.add_argument('--n', type=int, default=64)
    ap.add_argument('--d', type=int, default=3)
    ap.add_argument('--B4', type=int, default=16384, help='cfg4: GPs per GPU')
    ap.add_argument('--S', type=int, default=64)
    ap.add_argument('--n4', type=int, default=128)
    ap.add_argument('--e2e-B', dest='e2e_B', type=int, default=0, help='GPs of the e2e leg (0: the full batch if host memory allows, else a quarter)')
    ap.add_argument('--no-nlml', action='store_true')
    ap.add_argument('--skip-extras', action='store_true', help='profiling runs: no e2e / CPU baseline / parity legs')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        c = make_ctx(args)
        if args.workload == 'cfg4':
            run_cfg4(args, c)
        else:
            run_cfg3(args, c)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == '__main__':
    main()
