"""Monte-Carlo replica driver (BASELINE.json configs[4] / SURVEY.md par. 8f-2): R seeded replicas of one RTO episode,
sharded over the GPUs of one box, gathered with one NCCL all-gather of (objective, replica id) records.

The reference runs its 30 "Monte-Carlo" episodes sequentially in one Python process sharing one RNG stream
(run_simple/main_simple.py:174, run_WO/main_WO_prior_EI.py:105).  Here replica r is an independent unit seeded with
``np.random.seed(r); random.seed(r)`` (both RNGs: the reference leaves the stdlib one unseeded, utilities_2.py:684), so
any replica can be reproduced on its own.  Rank k of K owns replicas [lo, hi) = parallel.shard_range(R, k, K).

--mode lockstep (default): the rank's replicas advance TOGETHER (gp_rto_ei_b200/lockstep.py): every SLSQP round of every
replica's fit / acquisition is one batched kernel launch, the host only runs SciPy's SLSQP C routine per instance.  A
lock-step replica is bit-identical to the same seed run on its own (tests/test_gpu_lockstep.py).  `--workers W` splits the
rank's block over W processes sharing its GPU, each in lock step on its part: the per-instance SLSQP calls (~6 us each,
GIL-bound) are the bottleneck, so W = host cores / GPUs scales almost linearly; the kernels of W processes time-slice the
device, which is idle most of the time anyway.
--mode sequential: one episode after the other through the drop-in's ITR_GP_RTO (0.87 s per toy episode; reference ~26 s).
The only collective is the final gather of (objective, replica id).

    python -m torch.distributed.run --nproc-per-node 8 -m gp_rto_ei_b200.replicas --replicas 8192 --problem wo --workers 4
    python -m gp_rto_ei_b200.replicas --replicas 64 --problem simple            (single GPU)
"""
import argparse
import contextlib
import io
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def problem_setup(name):
    """Arguments of the published experiments: run_simple/main_simple.py:179-206 (toy problem, EI, prior, known
    noise) and run_WO/main_WO_prior_EI.py (Williams-Otto, EI, model prior, noisy plant with known noise)."""
    from gp_rto_ei_b200.sub_uts import systems as S
    if name == 'simple':
        return dict(obj_model=S.Benoit_Model, obj_system=S.Benoit_System, cons_model=[S.con1_model], cons_system=[S.con1_system_tight],
                    x0=np.array([1.1, -0.1]), Xtrain=np.array([[1.2, 0.], [1.4, 0.1], [1.3, -0.1]]), bounds=np.array([[-.6, 1.5], [-1., 1.]]),
                    noise=[1e-3] * 2, scale_inputs=False, truth=S.Benoit_System_noiseless,
                    truth_cons=[S.con1_system_tight_noiseless])
    if name == 'wo':
        model, plant = S.WO_model(), S.WO_system()
        return dict(obj_model=model.WO_obj_ca, obj_system=plant.WO_obj_sys_ca, cons_model=[model.WO_con1_model_ca, model.WO_con2_model_ca],
                    cons_system=[plant.WO_con1_sys_ca, plant.WO_con2_sys_ca], x0=np.array([6.9, 83.]),
                    Xtrain=np.array([[5.7, 74.], [6.35, 74.9], [6.6, 75.], [6.75, 79.]]), bounds=np.array([[4., 7.], [70., 100.]]),
                    noise=[0.5 ** 2, 5e-8, 5e-8], scale_inputs=True, truth=plant.WO_obj_sys_ca_noise_less,
                    truth_cons=[plant.WO_con1_sys_ca_noise_less, plant.WO_con2_sys_ca_noise_less])
    raise ValueError(name)


def run_replica(seed, problem='simple', obj_setting=3, n_iter=20, multi_opt=30, multi_hyper=15, trajectory=False):
    """One seeded episode through the sequential drop-in (ITR_GP_RTO.RTO_routine); returns the noise-free objective of the
    final iterate, its feasibility and the iterate (trajectory=True: also X_opt, y_opt, TR_l, backtrack)."""
    from gp_rto_ei_b200.sub_uts import utilities_2 as u2
    p = problem_setup(problem)
    np.random.seed(seed)
    random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        rto = u2.ITR_GP_RTO(p['obj_model'], p['obj_system'], p['cons_model'], p['cons_system'], p['x0'].copy(), 0.25, 0.7, 0.2, 0.8, 0.8,
                            1.2, n_iter, ['data0', p['Xtrain'].copy()], p['bounds'], obj_setting=obj_setting, noise=p['noise'],
                            multi_opt=multi_opt, multi_hyper=multi_hyper, TR_scaling=False, TR_curvature=False, store_data=True,
                            inner_TR=False, scale_inputs=p['scale_inputs'])
        X_opt, y_opt, TR_l, xnew, backtrack = rto.RTO_routine()
    x = np.asarray(xnew, dtype=float).flatten()
    f = float(np.asarray(p['truth'](x)).reshape(-1)[0])
    feas = all(float(np.asarray(c(x)).reshape(-1)[0]) > 0 for c in p['truth_cons'])
    out = dict(seed=seed, objective=f, feasible=bool(feas), x=x.tolist(), backtracks=int(np.sum(backtrack)), TR_final=float(TR_l[-1]))
    if trajectory:
        out.update(X_opt=np.asarray(X_opt), y_opt=np.asarray(y_opt), TR_l=np.asarray(TR_l, dtype=float), backtrack=np.asarray(backtrack, dtype=bool))
    return out


def _worker(args):
    seeds, device, kw = args
    import torch
    torch.cuda.set_device(device)
    return [run_replica(s, **kw) for s in seeds]


def _lockstep_worker(args):
    """One process: its seeds in lock step (gp_rto_ei_b200/lockstep.py), in batches of at most `batch` replicas."""
    seeds, device, kw, batch = args
    import torch
    torch.cuda.set_device(device)
    from gp_rto_ei_b200 import api, lockstep
    h = api.default_handle(device)
    out, stats = [], {}
    t0 = time.perf_counter()
    for lo in range(0, len(seeds), batch):
        chunk = list(seeds[lo:lo + batch])
        ls = lockstep.LockstepRTO(kw['problem'], chunk, obj_setting=kw['obj_setting'], n_iter=kw['n_iter'], multi_opt=kw['multi_opt'],
                                  multi_hyper=kw['multi_hyper'], handle=h)
        res = ls.run()
        f, feas = ls.outcome(res)
        for j, sd in enumerate(chunk):
            out.append(dict(seed=sd, objective=float(f[j]), feasible=bool(feas[j]), x=res['xnew'][j].tolist(),
                            backtracks=int(res['backtrack'][j].sum()), TR_final=float(res['TR_l'][j, -1]), dead=bool(res['dead'][j])))
        for k, v in ls.stats.items():
            stats[k] = stats.get(k, 0) + v
    stats['seconds'] = time.perf_counter() - t0
    stats['gpu_launches'] = h.launch_count
    return out, stats


def run_shard(seeds, device, workers, kw, mode='lockstep', batch=512):
    seeds = list(seeds)
    if mode == 'sequential':
        if workers <= 1 or len(seeds) <= 1:
            return _worker((seeds, device, kw)), {}
        import multiprocessing as mp
        chunks = [seeds[i::workers] for i in range(workers)]
        with mp.get_context('spawn').Pool(workers) as pool:
            parts = pool.map(_worker, [(c, device, kw) for c in chunks if c])
        out = [r for part in parts for r in part]
        out.sort(key=lambda r: r['seed'])
        return out, {}
    if workers <= 1 or len(seeds) <= 1:
        return _lockstep_worker((seeds, device, kw, batch))
    import multiprocessing as mp
    per = (len(seeds) + workers - 1) // workers
    chunks = [seeds[i * per:(i + 1) * per] for i in range(workers)]
    with mp.get_context('spawn').Pool(workers) as pool:
        parts = pool.map(_lockstep_worker, [(c, device, kw, batch) for c in chunks if c])
    out = [r for part, _ in parts for r in part]
    stats = {}
    for _, st in parts:
        for k, v in st.items():
            stats[k] = max(stats.get(k, 0), v) if k == 'seconds' else stats.get(k, 0) + v
    return out, stats


class _UtilSampler:
    """nvidia-smi utilization.gpu of one device while the replicas run (GPU-busy fraction of a host-bound driver)."""

    def __init__(self, gpu):
        import subprocess
        import threading
        self.vals = []
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu), '--query-gpu=utilization.gpu', '--format=csv,noheader,nounits', '-lms', '250'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            try:
                self.vals.append(float(line.strip()))
            except ValueError:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        return (sum(self.vals) / len(self.vals) / 100.0) if self.vals else None


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--replicas', type=int, default=8)
    ap.add_argument('--problem', default='simple', choices=['simple', 'wo'])
    ap.add_argument('--obj', type=int, default=3)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--multi-opt', type=int, default=30)
    ap.add_argument('--multi-hyper', type=int, default=15)
    ap.add_argument('--mode', default='lockstep', choices=['lockstep', 'sequential'],
                    help='lockstep: all replicas of a worker advance together, one batched launch per SLSQP round; sequential: one episode after the other')
    ap.add_argument('--workers', type=int, default=1, help='worker processes per rank sharing its GPU (each runs its replicas in lock step); 0 = host cores / ranks')
    ap.add_argument('--batch', type=int, default=512, help='replicas per lock-step batch inside a worker')
    a = ap.parse_args(argv)
    import torch
    import torch.distributed as dist
    from gp_rto_ei_b200 import api, parallel
    rank, local_rank, world = parallel.init()
    torch.cuda.set_device(local_rank)
    if a.workers <= 0:
        a.workers = max(1, len(os.sched_getaffinity(0)) // world)
    lo, hi = parallel.shard_range(a.replicas, rank, world)
    kw = dict(problem=a.problem, obj_setting=a.obj, n_iter=a.iters, multi_opt=a.multi_opt, multi_hyper=a.multi_hyper)
    util = _UtilSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    res, stats = run_shard(range(lo, hi), local_rank, a.workers, kw, a.mode, a.batch)
    dt = time.perf_counter() - t0
    busy = util.stop() if util else None
    # records: (objective of the final iterate, +inf when infeasible; global replica id) -> one all-gather -> arg-min
    h = api.Handle(local_rank)
    per = (a.replicas + world - 1) // world                              # equal record counts per rank (pad with +inf)
    score = torch.full((per,), float('inf'), dtype=torch.float64, device=h.device)
    idx = torch.full((per,), -1, dtype=torch.int64, device=h.device)
    allv = torch.full((per,), float('nan'), dtype=torch.float64, device=h.device)
    for j, r in enumerate(res):
        score[j] = r['objective'] if r['feasible'] else float('inf')
        idx[j] = r['seed']
        allv[j] = r['objective']
    rec = parallel.allgather_records(h.pack_best(score, idx))
    best_v, best_seed, _ = parallel.global_argmin(rec)
    tmax = torch.tensor([dt], dtype=torch.float64, device=h.device)
    tot = torch.tensor([float(stats.get(k, 0)) for k in ('slsqp_calls_fit', 'slsqp_calls_acq', 'nlml_evals', 'k4_points', 'gpu_launches')] +
                       [float(sum(1 for r in res if r.get('dead')))], dtype=torch.float64, device=h.device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        s, i = parallel.unpack_records(rec)
        ok = (i >= 0) & torch.isfinite(s)
        sec = float(tmax[0])
        sign = -1.0 if a.problem == 'wo' else 1.0            # Williams-Otto: objective = -profit
        line = {'metric': 'RTO replicas/s', 'value': a.replicas / sec, 'replicas': a.replicas, 'n_gpus': world, 'mode': a.mode,
                'workers_per_gpu': a.workers, 'host_cores': len(os.sched_getaffinity(0)), 'lockstep_batch': a.batch if a.mode == 'lockstep' else None, 'problem': a.problem, 'iters': a.iters,
                'multi_opt': a.multi_opt, 'multi_hyper': a.multi_hyper, 'seconds': sec, 'feasible_fraction': float(ok.sum()) / a.replicas,
                'best_objective': best_v, 'best_replica': best_seed,
                'mean_feasible_objective': float(s[ok].mean()) if bool(ok.any()) else None,
                'best_profit' if a.problem == 'wo' else 'best_f': sign * best_v,
                'mean_feasible_profit' if a.problem == 'wo' else 'mean_feasible_f': (sign * float(s[ok].mean())) if bool(ok.any()) else None,
                'optimum': 75.81 if a.problem == 'wo' else 0.145,
                'slsqp_calls': {'fit': int(tot[0]), 'acquisition': int(tot[1])}, 'nlml_evals': int(tot[2]), 'k4_point_evals': int(tot[3]),
                'gpu_launches': int(tot[4]), 'aborted_replicas': int(tot[5]), 'gpu_busy_fraction_rank0': busy,
                'collective': 'all-gather of (objective, replica id) records + redundant arg-min' if world > 1 else 'none'}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
