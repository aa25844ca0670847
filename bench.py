#!/usr/bin/env python
"""Benchmark of the mcpy hot path on B200: trial-move delta-E evaluations per second.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

Workload (BASELINE.json configs[1], SURVEY.md 8d "S2"): a 2 000-atom Lennard-Jones
fluid (eps = sigma = 1, rc = 3, rho* = 0.6, periodic cubic box) and a batch of
pre-generated GCMC proposals against it -- insertions, deletions and displacements
in the ratio 1:1:1.  One "step" scores the whole batch once.

  value      whole-job delta-E evaluations / s with the configuration and the
             proposals already resident in HBM (the five launches of the trial
             pipeline per step, CUDA events on the launching stream, L2 flushed
             between steps).
  e2e        the same metric through the plug-in with HOST buffers: every step
             uploads the configuration (cell list rebuilt) and the proposals and reads
             the delta-E array back (B200Calculator.set_configuration +
             trial_delta_energies -> mcpyb_upload_host + mcpyb_trial_delta_e_host).
  dropin     what an unmodified mcpy run gets today: one
             calculator.get_potential_energy(atoms) per trial, sequential, host buffers
             (the engine answers with its resident base + one single-launch delta).
  stages_ms  device time of each launch of the pipeline (CUDA events between them).
  roofline   the dominant kernel (the row sums) against the FP64 pipe: algorithmic flop
             per launch / its own event-timed duration, over the DFMA rate measured in
             this process (MEASURED_PEAKS.json has no FP64 figure).  Tensor cores are
             not used: nothing on this path is a dense contraction.
  cpu_baseline / --impl reference
             the reference's cost of one trial: one full-energy evaluation
             (grand_canonical_ensemble.py:283-284) by the port of ASE's
             LennardJones + NeighborList cost structure in oracle/lj.py, one process
             per host core.  ASE itself (third-party, not vendored) is not installable
             offline, hence kind = "port".

With N > 1 (torchrun, one rank per GPU) every rank owns its own replica and its own
proposal batch (weak scaling); there is no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'trial-move dE evals/sec'
UNIT = 'dE evals/s'
N_ATOMS = 2000
RC = 3.0
RHO = 0.6
ALG_FLOP_PER_CANDIDATE = 17          # 3 sub + MIC + r2 + compare
ALG_FLOP_PER_PAIR = 8                # extra work of an in-cutoff pair
ALG_CANDIDATES_PER_SITE = 27 * RHO * RC ** 3      # 27-cell stencil at cell edge rc


def workload_config(trials):
    return {'workload': 'S2: 2000-atom LJ GCMC (eps=sigma=1, rc=3, rho*=0.6, cubic pbc), '
                        f'{trials} proposals/step insert:delete:displace = 1:1:1',
            'n_atoms': N_ATOMS, 'trials_per_step': trials, 'potential': 'ase.LennardJones(rc=3)',
            'l2': 'flushed between timed steps (256 MiB write); working set itself is L2-resident'}


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One trial = one full-energy evaluation of the trial configuration."""
    positions, cell, pbc = args
    from oracle.lj import AseStyleLJ
    return AseStyleLJ(1.0, 1.0, RC).energy(positions, cell, pbc)


def cpu_reference_rate(cfg, tb, n_sample, pool, cores):
    from mcpy_b200.utils import synthetic
    tasks = []
    for t in range(n_sample):
        p1, _ = synthetic.apply_trial(cfg, tb, t)
        tasks.append((p1, cfg['cell'], cfg['pbc']))
    t0 = time.perf_counter()
    pool.map(_cpu_worker, tasks, chunksize=max(1, n_sample // (cores * 2)))
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from mcpy_b200.utils import synthetic
    cfg = synthetic.s2_lj(N_ATOMS, RC, RHO)
    cores = os.cpu_count() or 1
    n_sample = 4 * cores
    tb = synthetic.trial_batch(cfg, max(n_sample, 64), seed=7)
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_rate(cfg, tb, min(cores, n_sample), pool, cores)
        total_t, total_n = 0.0, 0
        for _ in range(args.steps):
            _, dt = cpu_reference_rate(cfg, tb, n_sample, pool, cores)
            total_t += dt
            total_n += n_sample
    value = total_n / total_t
    sample = (f'{n_sample} of the batch proposals per step, each one full-energy evaluation '
              f'(ASE-style neighbour-list rebuild + per-atom loop, numpy), {cores} processes')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args.trials),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from mcpy_b200.calculators import B200Calculator
    from mcpy_b200.utils import synthetic

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the mcpy_b200 engine has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    T = args.trials
    cfg = synthetic.s2_lj(N_ATOMS, RC, RHO, seed=2 + rank)          # one replica per rank
    tb = synthetic.trial_batch(cfg, T, seed=7 + 1000 * rank)
    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=RC, device=local)
    eng = calc.engine
    stream = torch.cuda.Stream(device=dev)          # a real (non-legacy) stream: events and the
    torch.cuda.set_stream(stream)                    # engine's launches share it
    eng.set_stream(stream.cuda_stream)

    class _Cfg:                       # minimal ASE-shaped view for the plug-in calls
        def __init__(self, c):
            self.positions, self.numbers, self.cell, self.pbc = c['positions'], c['numbers'], c['cell'], c['pbc']
    atoms = _Cfg(cfg)

    # ---- resident inputs for the kernel-only number
    calc.set_configuration(atoms)
    d = {k: torch.from_numpy(np.ascontiguousarray(tb[k])).to(dev) for k in
         ('old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')}
    d_dE = torch.empty(T, dtype=torch.float64, device=dev)
    d_pairs = torch.empty(T, dtype=torch.int32, device=dev)
    n_old, n_new = int(tb['old_off'][-1]), int(tb['new_off'][-1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def launch(pairs_ptr=0):
        eng.trial_delta_e_dev(T, 0, d['old_off'].data_ptr(), d['old_idx'].data_ptr(),
                              d['new_off'].data_ptr(), d['new_pos'].data_ptr(), d['new_Z'].data_ptr(),
                              n_old, n_new, d_dE.data_ptr(), pairs_ptr)

    launch(d_pairs.data_ptr())
    torch.cuda.synchronize()
    pairs_per_step = int(d_pairs.sum().item())
    dE_ref = d_dE.clone()

    for _ in range(args.warmup):
        flush.zero_()
        launch()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = eng.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                                 # evict the working set from L2
        starts[k].record(stream)
        launch()
        ends[k].record(stream)
    barrier()
    gpu_launches = eng.launch_count - launches0
    # per-stage device times of the same step (separate, untimed pass: events between the launches)
    eng.stage_timing(True)
    for _ in range(max(3, args.steps // 2)):
        flush.zero_()
        launch()
        stage_ms, stage_calls = eng.stage_timing(True)
    eng.stage_timing(False)
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    kernel_ms = max_over_ranks(kernel_ms)
    assert torch.equal(d_dE, dE_ref), 'trial kernel is not deterministic'
    value = world * T * args.steps / (kernel_ms * 1e-3)
    ms_per_step = kernel_ms / args.steps

    # ---- end to end through the plug-in with host buffers
    # the step's inputs and its result live in page-locked host arrays (mcpy_b200.pinned_copy): the engine
    # hands them to the copy engine as they are; pageable numpy arrays would be staged first ('e2e_pageable')
    from mcpy_b200 import pinned_copy, pinned_empty
    trials_pageable = {k: np.ascontiguousarray(tb[k]) for k in ('old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')}
    trials_host = {k: pinned_copy(v) for k, v in trials_pageable.items()}
    dE_host = pinned_empty(T, np.float64)

    def e2e_step(tr=trials_host, out=dE_host):
        calc.set_configuration(atoms)                                 # H2D config + cell list
        return eng.trial_delta_e(tr['old_off'], tr['old_idx'], tr['new_off'],
                                 tr['new_pos'], tr['new_Z'], out=out)            # H2D trials, D2H dE

    for _ in range(args.warmup):
        out = e2e_step()
    assert np.array_equal(out, dE_ref.cpu().numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * T * args.steps / e2e_s
    # the same call with ordinary (pageable) numpy arrays, for the record
    for _ in range(args.warmup):
        out = e2e_step(trials_pageable, None)
    assert np.array_equal(out, dE_ref.cpu().numpy())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(trials_pageable, None)
    torch.cuda.synchronize()
    e2e_pageable_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d = (N_ATOMS * (24 + 4) + 256) + sum(int(v.nbytes) for v in trials_host.values()) + 4 * T
    d2h = 8 * T

    # ---- drop-in chain: one get_potential_energy(atoms) per trial, sequential
    n_chain = min(T, args.chain)
    chain_cfgs = [synthetic.apply_trial(cfg, tb, t) for t in range(min(n_chain, 256))]

    class _A:
        pass
    a = _A()
    a.cell, a.pbc = cfg['cell'], cfg['pbc']
    for p1, z1 in chain_cfgs[:8]:
        a.positions, a.numbers = p1, z1
        calc.get_potential_energy(a)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_chain):
        a.positions, a.numbers = chain_cfgs[i % len(chain_cfgs)]
        calc.get_potential_energy(a)
    chain_s = max_over_ranks(time.perf_counter() - t0)
    dropin_value = world * n_chain / chain_s

    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary workloads of the same path (rank 0 only, untimed by the driver's metric):
    # EMT energies / trial dE on S3 (Ag140 + O) and S5 (~10.8k-atom CuPd + CO), free volume on both
    extra = {}
    if rank == 0 and not args.no_extra:
        extra = secondary_workloads(local)

    # ---- roofline of the trial kernel (FP64 pipe)
    sites = n_old + n_new
    alg_flop = sites * ALG_CANDIDATES_PER_SITE * ALG_FLOP_PER_CANDIDATE + pairs_per_step * ALG_FLOP_PER_PAIR
    peak = eng.fp64_peak_tflops()
    rows_ms = stage_ms['row_sums']
    achieved = alg_flop / (rows_ms * 1e-3) / 1e12
    pair_rate = sum_over_ranks(pairs_per_step) / (ms_per_step * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        n_sample = 4 * cores
        with mp.get_context('fork').Pool(cores) as pool:
            cpu_reference_rate(cfg, tb, cores, pool, cores)          # warm the workers
            rate, dt = cpu_reference_rate(cfg, tb, n_sample, pool, cores)
        cpu_baseline = {
            'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{n_sample} of the batch proposals, one full-energy evaluation each '
                      f'(ASE-style neighbour-list rebuild + per-atom loop, numpy), {cores} processes, '
                      f'{dt:.1f} s wall'}

    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('lj_trial_rows_block_kernel_dram_bytes_per_launch')
        except Exception:
            traffic = None

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(T), replicas_per_gpu=1,
                       parallelism=f'replica-parallel x{world}, no data-path collective'),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'api': 'B200Calculator.set_configuration + Engine.trial_delta_e '
                       '(mcpyb_upload_host + mcpyb_trial_delta_e_host), host numpy arrays in page-locked memory '
                       '(mcpy_b200.pinned_copy / pinned_empty)',
                'ms_per_step': 1e3 * e2e_s / args.steps,
                'pageable': {'value': world * T * args.steps / e2e_pageable_s,
                             'ms_per_step': 1e3 * e2e_pageable_s / args.steps,
                             'note': 'same calls with ordinary numpy arrays: staged through pinned memory inside the engine'}},
        'dropin': {'value': dropin_value, 'unit': UNIT, 'calls': n_chain,
                   'api': 'calculator.get_potential_energy(atoms) per trial, sequential '
                          '(mcpyb_tracked_energy_host: resident base + one delta; the trial rides in the '
                          'kernel parameters, the result lands in pinned host memory)',
                   'paths': {'full': calc.path_counts[0], 'incremental': calc.path_counts[1],
                             'identical': calc.path_counts[2]},
                   'us_per_call': 1e6 * chain_s / n_chain},
        'stages_ms': {k: round(v, 5) for k, v in stage_ms.items()},
        'pair_interactions_per_s': pair_rate,
        'pairs_per_step': pairs_per_step,
        'gpu_launches': int(gpu_launches),
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak if peak else None, 'traffic': traffic,
                     'kernel': 'lj_trial_rows_block_kernel<4,12>',
                     'kernel_ms': rows_ms, 'kernel_share_of_step': rows_ms / sum(stage_ms.values()),
                     'algorithmic_flop_per_launch': alg_flop,
                     'step_achieved': alg_flop / (ms_per_step * 1e-3) / 1e12,
                     'peak_source': 'measured here: DFMA microbenchmark (mcpyb_fp64_peak_tflops); '
                                    'MEASURED_PEAKS.json holds only HBM and bf16 figures',
                     'note': 'HBM is not the bound: the 64 KB configuration is L1/L2 resident; algorithmic flop '
                             'follow SURVEY 8d (27-cell stencil at cell edge rc: 437 candidates x 17 flop per site '
                             '+ 8 per in-cutoff pair); the kernel itself visits ~200 candidates per site after '
                             'pruning, see profiles/ for the FP64-pipe counters'},
        'clocks': clocks,
    }
    if cpu_baseline is not None:
        line['cpu_baseline'] = cpu_baseline
    if extra:
        line['secondary'] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _time(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def secondary_workloads(device):
    """EMT and free-volume numbers next to their CPU counterparts (small, bounded samples)."""
    from scipy.spatial import cKDTree

    from mcpy_b200 import Engine
    from mcpy_b200.utils import synthetic
    out = {}
    eng = Engine(device)
    eng.set_emt()
    for name, cfg, T in (('S3_Ag140_O20', synthetic.s3_ag_o(), 4096), ('S5_CuPd_CO_10825', synthetic.s5_cupd_co(), 4096)):
        pos, num, cell, pbc = cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']
        t_e = _time(lambda: eng.potential_energy(pos, num, cell, pbc), 5)
        eng.upload([pos], [num], cell.reshape(1, 9), pbc.reshape(1, 3))
        eng.energies()
        tb = synthetic.trial_batch(cfg, T, seed=11, insert_Z=8)
        args = (tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'])
        t_t = _time(lambda: eng.trial_delta_e(*args), 3)
        # the sequential one-call-per-trial chain of an unmodified run (resident base + one delta)
        chain = [synthetic.apply_trial(cfg, tb, t) for t in range(64)]
        for p1, z1 in chain[:4]:
            eng.tracked_energy(p1, z1, cell, pbc)
        t0 = time.perf_counter()
        for rep_ in range(4):
            for p1, z1 in chain:
                eng.tracked_energy(p1, z1, cell, pbc)
        t_chain = (time.perf_counter() - t0) / (4 * len(chain))
        # free volume: host points uploaded vs PCG64 regenerated on the device, vs scipy cKDTree
        P = int(cfg['mc_sample_points'])
        rng = np.random.default_rng(3)
        t_dev = _time(lambda: eng.sphere_free_volume_pcg64(rng, P, cfg['sphere_radius'], np.zeros(3), pos, cfg['radii']), 3)

        def cpu_fv():
            pts = (np.random.default_rng(3).random((int(1.5 * P) + 1, 3)) - 0.5) * 2.0 * cfg['sphere_radius']
            pts = pts[np.einsum('ij,ij->i', pts, pts) <= cfg['sphere_radius'] ** 2][:P]
            cov = np.zeros(len(pts), bool)
            for r in np.unique(cfg['radii']):
                if r > 0:
                    d, _ = cKDTree(pos[cfg['radii'] == r]).query(pts, k=1, distance_upper_bound=float(r))
                    cov |= np.isfinite(d)
            return cov.sum()
        t_cpu = _time(cpu_fv, 1)
        out[name] = {
            'n_atoms': int(len(pos)),
            'emt_energy_calls_per_s_host_buffers': 1.0 / t_e,
            'emt_trial_dE_evals_per_s_host_buffers': T / t_t,
            'emt_dropin_chain_us_per_call': 1e6 * t_chain,
            'free_volume_points': P,
            'free_volume_points_per_s_device_pcg64': P / t_dev,
            'free_volume_points_per_s_cpu_ckdtree_1core': P / t_cpu,
        }
    eng.close()
    # S4: 64 replicas x 2000-atom LJ through the batched plug-in call.  Each call carries ONE trial move
    # per replica, as BatchedReplicaExchange._batched_single_move issues them
    # (batched_replica_exchange.py:235-302): full recomputation vs resident bases + one batched delta.
    from mcpy_b200.calculators import B200Calculator

    class _A:
        pass
    base, calls = [], []
    for k in range(64):
        c = synthetic.s2_lj(N_ATOMS, RC, RHO, seed=100 + k)
        tbk = synthetic.trial_batch(c, 8, seed=300 + k)
        base.append(c)
        calls.append([synthetic.apply_trial(c, tbk, t) for t in range(8)])

    def batch_of(t):
        reps = []
        for k in range(64):
            a = _A()
            a.positions, a.numbers = calls[k][t]
            a.cell, a.pbc = base[k]['cell'], base[k]['pbc']
            reps.append(a)
        return reps
    batches = [batch_of(t) for t in range(8)]
    res = {}
    for label, inc in (('full_recompute', False), ('resident_bases', True)):
        calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=RC, device=device, incremental=inc)
        for bt in batches[:2]:
            calc.get_potential_energies(bt)
        t0 = time.perf_counter()
        for rep_ in range(3):
            for bt in batches:
                res[label + '_E'] = calc.get_potential_energies(bt)
        res[label] = (time.perf_counter() - t0) / (3 * len(batches))
        if inc:
            res['paths'] = list(calc.path_counts)
            t_1 = _time(lambda: [calc.get_potential_energy(a) for a in batches[0][:8]], 3) / 8
        calc.engine.close()
    scale = np.maximum(1.0, np.abs(res['full_recompute_E']))
    assert np.all(np.abs(res['resident_bases_E'] - res['full_recompute_E']) <= 1e-10 * scale)
    out['S4_64_replicas_x_2000_LJ'] = {
        'batched_call_ms_full_recompute': 1e3 * res['full_recompute'],
        'batched_call_ms_resident_bases': 1e3 * res['resident_bases'],
        'replica_energies_per_s_batched': 64 / res['resident_bases'],
        'replica_energies_per_s_batched_full_recompute': 64 / res['full_recompute'],
        'replica_energies_per_s_one_by_one': 1.0 / t_1,
        'paths_full_incremental_identical': res['paths'],
    }
    # relax-then-energy (BaseCalculator.get_potential_energy, base_calculator.py:14-22): LBFGS steps per second with
    # the loop inside the engine, next to the same force evaluations issued one Python call at a time
    relax = {}
    for name, cfg, pot in (('S1_argon_500_LJ', synthetic.s1_argon(), 'lj'), ('S3_Ag140_O20_EMT', synthetic.s3_ag_o(), 'emt')):
        eng = Engine(device)
        if pot == 'lj':
            eng.set_lj(cfg['epsilon'], cfg['sigma'], cfg['rc'])
        else:
            eng.set_emt()
        pos, num, cell, pbc = cfg['positions'], cfg['numbers'], cfg['cell'], cfg['pbc']
        eng.relax_lbfgs(pos, num, cell, pbc, fmax=0.0, steps=5)
        t0 = time.perf_counter()
        _, _, nst, _ = eng.relax_lbfgs(pos, num, cell, pbc, fmax=0.0, steps=40)
        t_loop = (time.perf_counter() - t0) / (nst + 1)
        t_call = _time(lambda: eng.energy_forces(pos, num, cell, pbc), 20)
        relax[name] = {'n_atoms': int(len(pos)), 'lbfgs_us_per_step_engine_loop': 1e6 * t_loop,
                       'energy_forces_us_per_python_call': 1e6 * t_call}
        eng.close()
    out['relax_then_energy'] = relax
    return out



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--trials', type=int, default=65536, help='proposals per step')
    ap.add_argument('--chain', type=int, default=2000, help='sequential drop-in calls to time')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary EMT / free-volume numbers')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        return run_reference(args)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
