#!/usr/bin/env python
"""Benchmark of the mcpy hot path on B200: trial-move delta-E evaluations per second.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

Workload, identical at every N (weak scaling): the block of BASELINE.json configs[3] ("64 replicas x
~2 000-atom LJ, sharded over 8 B200": SURVEY.md 8d "S4") that one GPU owns -- 8 replicas of configs[1]
("S2": 2 000-atom Lennard-Jones fluid, eps = sigma = 1, rc = 3, rho* = 0.6, periodic cubic box) on a
chemical-potential ladder.  One SWEEP scores 65 536 pre-generated GCMC proposals per replica (524 288 dE
evaluations: insertions, deletions, displacements 1:1:1) against the rank's resident configurations in ONE
batched launch of the trial pipeline.  Every 10 sweeps the ladder attempts an exchange: the records
[E, beta, N, mu, Lambda, ...] of all replicas are all_gathered (NCCL over NVLink; the path's only collective),
every rank takes the same Metropolis decisions, accepted pairs swap configurations -- inside a rank by
relabelling device rows, across ranks point to point -- and the changed cell lists are rebuilt.  One STEP =
`--sweeps` sweeps (default 80) and their exchange attempts, so that K = 20 steps time more than half a second.

  value      whole-job dE evaluations / s over all ranks, configurations and proposals resident in HBM
             (ResidentLadderShard + Engine.trial_delta_e_dev); K steps between two barriers, CUDA events on
             the launching stream, max over ranks.  Inputs are larger than L2: 16 distinct proposal batches
             (18 MB each) and 35 MB of per-configuration stencil lists cycle through the 126 MB L2.
  e2e        the same steps through the plug-in API with HOST buffers: every sweep uploads the 8
             configurations from their ``Atoms`` (B200Calculator.set_configurations), the proposals from
             page-locked numpy arrays, and reads the dE array back; every 10 sweeps the replicas -- unmodified
             reference GrandCanonicalEnsemble objects -- go through ShardedReplicaExchange's exchange step
             (host records, NCCL all_gather, pickled state transport).
  driver     what an unmodified mcpy replica-exchange run gets: ShardedReplicaExchange.run() over the same
             replicas, ONE trial per replica per batched get_potential_energies call, exchanges every 10 steps.
  dropin     one calculator.get_potential_energy(atoms) per trial, sequential (a single GCMC chain).
  roofline   the dominant kernel (the row sums of the trial pipeline) against the FP64 pipe: algorithmic flop
             per launch / its event-timed duration over the DFMA rate measured in this process; further
             kernels (cell-list build, EMT trial, free volume) under ``secondary.*.roofline``.
  cpu_baseline / --impl reference
             the reference's cost of one trial: one full-energy evaluation
             (grand_canonical_ensemble.py:283-284) by ASE's LennardJones when a real ``ase`` is importable,
             else by the port of its cost structure (oracle/lj.py::AseStyleLJ), one process per host core.
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
REPLICAS_PER_GPU = 8
TRIALS_PER_REPLICA = 65536
EXCHANGE_INTERVAL = 10
MU_LADDER = (-3.0, -1.8)             # SURVEY 8d S4, over the whole ladder
T_STAR = 2.0
ALG_FLOP_PER_CANDIDATE = 17          # 3 sub + MIC + r2 + compare
ALG_FLOP_PER_PAIR = 8                # extra work of an in-cutoff pair
ALG_CANDIDATES_PER_SITE = 27 * RHO * RC ** 3      # 27-cell stencil at cell edge rc


def workload_config(args):
    """One dictionary for both arms (the driver compares them)."""
    trials = REPLICAS_PER_GPU * TRIALS_PER_REPLICA
    return {'workload': f'S4 shard: {REPLICAS_PER_GPU} replicas per GPU x S2 (2000-atom LJ GCMC, eps=sigma=1, rc=3, '
                        f'rho*=0.6, cubic pbc) on a mu ladder; sweep = {TRIALS_PER_REPLICA} proposals per replica '
                        f'({trials} dE evals, insert:delete:displace = 1:1:1); step = {args.sweeps} sweeps + an '
                        f'exchange attempt every {EXCHANGE_INTERVAL} sweeps',
            'n_atoms': N_ATOMS, 'replicas_per_gpu': REPLICAS_PER_GPU, 'trials_per_sweep': trials,
            'sweeps_per_step': args.sweeps, 'exchange_interval_sweeps': EXCHANGE_INTERVAL,
            'potential': 'ase.LennardJones(rc=3)',
            'parallelism': 'replica-sharded, 8 replicas per GPU; one all_gather of the replica records per '
                           'exchange attempt, point-to-point transport of accepted cross-rank swaps',
            'l2': f'inputs larger than L2: {args.batches} distinct proposal batches '
                  f'({args.batches * 2.2 * TRIALS_PER_REPLICA / 8192:.0f} MB) and 35 MB of stencil lists cycle through the '
                  f'126 MB L2',
            'host_buffers': 'e2e: page-locked numpy arrays (mcpy_b200.pinned_copy); pageable arrays reported '
                            'under e2e.pageable'}


def local_configs(rank):
    from mcpy_b200.utils import synthetic
    return [synthetic.s2_lj(N_ATOMS, RC, RHO, seed=100 + rank * REPLICAS_PER_GPU + k) for k in range(REPLICAS_PER_GPU)]


def proposal_batch(cfgs, seed):
    """One sweep's proposals: TRIALS_PER_REPLICA per local replica, concatenated (``rep`` = replica of each trial)."""
    from mcpy_b200.utils import synthetic
    parts = [synthetic.trial_batch_periodic(c, TRIALS_PER_REPLICA, seed=seed * 64 + k) for k, c in enumerate(cfgs)]
    old_off = np.concatenate([[0]] + [p['old_off'][1:] + sum(int(q['old_off'][-1]) for q in parts[:k])
                                      for k, p in enumerate(parts)]).astype(np.int32)
    new_off = np.concatenate([[0]] + [p['new_off'][1:] + sum(int(q['new_off'][-1]) for q in parts[:k])
                                      for k, p in enumerate(parts)]).astype(np.int32)
    return dict(rep=np.repeat(np.arange(len(cfgs), dtype=np.int32), TRIALS_PER_REPLICA),
                old_off=old_off, new_off=new_off,
                old_idx=np.concatenate([p['old_idx'] for p in parts]),
                new_pos=np.concatenate([p['new_pos'] for p in parts]),
                new_Z=np.concatenate([p['new_Z'] for p in parts]),
                kind=np.concatenate([p['kind'] for p in parts]))


def trial_config(cfgs, tb, t):
    """(positions, numbers, cell, pbc) of the configuration after trial ``t`` of a batch."""
    c = cfgs[int(tb['rep'][t])]
    old = tb['old_idx'][tb['old_off'][t]:tb['old_off'][t + 1]]
    ns, ne = tb['new_off'][t], tb['new_off'][t + 1]
    keep = np.ones(len(c['positions']), bool)
    keep[old] = False
    return (np.concatenate([c['positions'][keep], tb['new_pos'][ns:ne]]),
            np.concatenate([c['numbers'][keep], tb['new_Z'][ns:ne]]), c['cell'], c['pbc'])


# --------------------------------------------------------------------------- CPU arm
def real_ase():
    """The real ASE if one is importable (never the stand-in under baseline/ase_shim)."""
    import importlib.util
    spec = importlib.util.find_spec('ase')
    if spec is None or 'ase_shim' in (spec.origin or ''):
        return None
    try:
        import ase
        from ase.calculators.lj import LennardJones  # noqa: F401
        return ase
    except Exception:
        return None


def _cpu_worker(task):
    """One trial = one full-energy evaluation of the trial configuration, as compute_energy issues it."""
    positions, numbers, cell, pbc, use_ase = task
    if use_ase:
        from ase import Atoms
        from ase.calculators.lj import LennardJones
        atoms = Atoms(numbers=numbers, positions=positions, cell=cell, pbc=pbc)
        return LennardJones(sigma=1.0, epsilon=1.0, rc=RC).get_potential_energy(atoms)
    from oracle.lj import AseStyleLJ
    return AseStyleLJ(1.0, 1.0, RC).energy(positions, cell, pbc)


def cpu_reference_rate(cfgs, tb, n_sample, pool, cores, use_ase, first=0):
    tasks = [trial_config(cfgs, tb, (first + t) % len(tb['rep'])) + (use_ase,) for t in range(n_sample)]
    t0 = time.perf_counter()
    pool.map(_cpu_worker, tasks, chunksize=max(1, n_sample // (cores * 2)))
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def cpu_kind(use_ase):
    return ('reference', 'ase.calculators.lj.LennardJones (real ASE)') if use_ase else \
        ('port', 'oracle/lj.py::AseStyleLJ (port of ASE LennardJones + NeighborList: kd-tree neighbour-list '
                 'rebuild + interpreted per-atom loop; ASE is not installable offline)')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import multiprocessing as mp
    use_ase = real_ase() is not None
    cfgs = local_configs(0)
    cores = os.cpu_count() or 1
    n_sample = 4 * cores
    tb = proposal_batch(cfgs, seed=7)
    with mp.get_context('fork').Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_reference_rate(cfgs, tb, min(cores, n_sample), pool, cores, use_ase, first=17 * w)
        total_t, total_n = 0.0, 0
        for k in range(args.steps):
            # samples stride across the replicas of the shard
            _, dt = cpu_reference_rate(cfgs, tb, n_sample, pool, cores, use_ase, first=k * 8209)
            total_t += dt
            total_n += n_sample
    value = total_n / total_t
    kind, how = cpu_kind(use_ase)
    sample = (f'{n_sample} proposals of the shard per step, each one full-energy evaluation of its 2000-atom trial '
              f'configuration (the reference recomputes the total for every trial), {how}, {cores} processes')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
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
                 '--format=csv,noheader,nounits', '-lms', '50'],
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


def _profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', name)))
    except Exception:
        return {}


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from mcpy_b200 import pinned_copy, pinned_empty
    from mcpy_b200.calculators import B200Calculator
    from mcpy_b200.ensembles import ResidentLadderShard

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

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())
    max_over_ranks = lambda x: reduce(x, dist.ReduceOp.MAX) if world > 1 else float(x)   # noqa: E731
    sum_over_ranks = lambda x: reduce(x, dist.ReduceOp.SUM) if world > 1 else float(x)   # noqa: E731

    R = REPLICAS_PER_GPU
    T = R * TRIALS_PER_REPLICA
    SW, XI, NB, K, W = args.sweeps, EXCHANGE_INTERVAL, args.batches, args.steps, args.warmup
    n_total = R * world
    cfgs = local_configs(rank)
    mus = np.linspace(MU_LADDER[0], MU_LADDER[1], n_total)
    betas = np.full(n_total, 1.0 / T_STAR)
    batches = [proposal_batch(cfgs, seed=1000 * rank + b) for b in range(NB)]

    calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=RC, device=local)
    eng = calc.engine
    stream = torch.cuda.Stream(device=dev)          # a real (non-legacy) stream: events, torch copies and the
    torch.cuda.set_stream(stream)                    # engine's launches share it
    eng.set_stream(stream.cuda_stream)

    # ---- kernel-only: configurations and proposals resident in HBM
    shard = ResidentLadderShard(eng, [(c['positions'], c['numbers'], c['cell'], c['pbc']) for c in cfgs],
                                betas=betas, mus=mus, seed=31, rank=rank, world_size=world, device=f'cuda:{local}')
    keys = ('rep', 'old_off', 'old_idx', 'new_off', 'new_pos', 'new_Z')
    d_batches = [{k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in keys} for b in batches]
    n_old = [int(b['old_off'][-1]) for b in batches]
    n_new = [int(b['new_off'][-1]) for b in batches]
    d_dE = torch.empty(T, dtype=torch.float64, device=dev)
    d_pairs = torch.empty(T, dtype=torch.int32, device=dev)

    def sweep(b, pairs_ptr=0):
        d = d_batches[b]
        eng.trial_delta_e_dev(T, d['rep'].data_ptr(), d['old_off'].data_ptr(), d['old_idx'].data_ptr(),
                              d['new_off'].data_ptr(), d['new_pos'].data_ptr(), d['new_Z'].data_ptr(),
                              n_old[b], n_new[b], d_dE.data_ptr(), pairs_ptr)

    # sanity before anything is timed: against the CPU oracle on a few trials, and bit-reproducible
    sweep(0, d_pairs.data_ptr())
    torch.cuda.synchronize()
    pairs_per_sweep = int(d_pairs.sum().item())
    dE_ref = d_dE.clone()
    if rank == 0:
        from oracle import cfast
        h = dE_ref.cpu().numpy()
        for t in (0, T // 3 + 1, T - 1):
            p1, _, cell, pbc = trial_config(cfgs, batches[0], t)
            c0 = cfgs[int(batches[0]['rep'][t])]
            e0 = cfast.lj_energy(c0['positions'], cell, pbc, 1.0, 1.0, RC)
            ref = cfast.lj_energy(p1, cell, pbc, 1.0, 1.0, RC) - e0
            assert abs(h[t] - ref) <= 1e-10 * max(1.0, abs(ref)) + 1e-13 * abs(e0), (t, h[t], ref)
    sweep(0)
    torch.cuda.synchronize()
    assert torch.equal(d_dE, dE_ref), 'trial kernel is not deterministic'

    counter = [0]

    def step():
        for s in range(SW):
            sweep(counter[0] % NB)
            counter[0] += 1
            if (s + 1) % XI == 0:
                shard.attempt_exchanges()

    for _ in range(W):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    comm = shard.comm
    c0, t0c, g0, x0 = comm.collective_seconds, comm.transport_seconds, comm.gathers, comm.transports
    rounds0, swaps0 = shard.rounds, sum(1 for *_, ok in shard.swap_log if ok)
    barrier()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(stream)
    for k in range(K):
        step()
        ev[k + 1].record(stream)
    barrier()
    gpu_launches = eng.launch_count - launches0
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    value = world * T * SW * K / (total_ms * 1e-3)
    ms_per_step = total_ms / K
    rounds = shard.rounds - rounds0
    exchange = {
        'attempts': rounds, 'accepted_swaps': sum(1 for *_, ok in shard.swap_log if ok) - swaps0,
        'cross_rank_swaps': comm.transports - x0,
    }
    # an exchange attempt on its own, untimed by the metric: host wall clock from an idle stream to an idle stream
    # (inside the timed loop the attempt's first synchronisation also drains the sweeps queued before it, so the
    # tallies of the loop measure queueing, not the collective)
    c1, g1 = comm.collective_seconds, comm.gathers
    tx_total = 0.0
    for _ in range(8):
        barrier()
        tx = time.perf_counter()
        shard.attempt_exchanges()
        torch.cuda.synchronize()
        tx_total += time.perf_counter() - tx
    exchange['exchange_us_per_attempt'] = 1e6 * max_over_ranks(tx_total) / 8
    exchange['collective_us_per_exchange'] = 1e6 * (comm.collective_seconds - c1) / max(1, comm.gathers - g1) if world > 1 else 0.0
    exchange['collective_note'] = ('one all_gather_into_tensor of the records and the two boundary configurations of every '
                                   'rank (host -> device, NCCL, device -> host of the records), measured from an idle stream')
    exchange['share_of_step'] = exchange['exchange_us_per_attempt'] * (SW // XI) / (ms_per_step * 1e3)
    exchange['collective_share_of_step'] = exchange['collective_us_per_exchange'] * (SW // XI) / (ms_per_step * 1e3)

    # per-stage device times of a sweep (separate, untimed pass: events between the launches)
    eng.stage_timing(True)
    for b in range(max(8, min(NB, 24))):
        sweep(b % NB)
        stage_ms, stage_calls = eng.stage_timing(True)
    eng.stage_timing(False)
    sweep(0)
    torch.cuda.synchronize()

    # ---- end to end through the plug-in with host buffers
    from baseline import refpath
    ase_kind, mcpy_root = refpath.enable()
    NBH = min(NB, 4)
    host_pageable = [{k: np.ascontiguousarray(batches[b][k]) for k in keys} for b in range(NBH)]
    host_pinned = [{k: pinned_copy(v) for k, v in hb.items()} for hb in host_pageable]
    dE_host = pinned_empty(T, np.float64)
    re = None
    e2e_note = None
    try:
        from ase import Atoms
        from mcpy.cell import Cell
        from mcpy.ensembles import GrandCanonicalEnsemble
        from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove, MoveSelector
        from mcpy_b200.ensembles import ShardedReplicaExchange

        def factory(T=None, mu=None, rank=0):
            c = local_configs(rank // R)[rank % R]
            atoms = Atoms(numbers=c['numbers'], positions=c['positions'].copy(), cell=c['cell'], pbc=c['pbc'])
            cell = Cell(atoms, seed=500 + rank)
            moves = [InsertionMove(cell, ['H'], 600 + rank), DeletionMove(cell, ['H'], 700 + rank),
                     DisplacementMove(['H'], 800 + rank, max_displacement=0.3)]
            ms = MoveSelector([1, 1, 1], moves, seed=900 + rank, n_moves=1)
            return GrandCanonicalEnsemble(atoms=atoms, cells=[cell], units_type='LJ', calculator=calc, mu=mu,
                                          species=['H'], temperature=T_STAR, move_selector=ms,
                                          random_seed=1000 + rank, traj_file=None, outfile=None)
        re = ShardedReplicaExchange(factory, calc, mus=[{'H': float(m)} for m in mus], gcmc_steps=0,
                                    exchange_interval=XI, seed=31, rank=rank, world_size=world,
                                    device=f'cuda:{local}', outfile=None, global_minimum_file=None)
        re._rebatch_initial_energies()
        replicas_atoms = lambda: [r.atoms for r in re.replicas]                      # noqa: E731
    except ImportError as exc:                    # the reference is not installed: no host-side exchange step
        e2e_note = f'exchange step skipped: {exc}'

        class _A:
            pass
        fixed = []
        for c in cfgs:
            a = _A()
            a.positions, a.numbers, a.cell, a.pbc = c['positions'], c['numbers'], c['cell'], c['pbc']
            fixed.append(a)
        replicas_atoms = lambda: fixed                                                # noqa: E731

    def e2e_sweep(hb, out):
        calc.set_configurations(replicas_atoms())                                     # H2D configurations + cell lists
        return eng.trial_delta_e(hb['old_off'], hb['old_idx'], hb['new_off'], hb['new_pos'], hb['new_Z'],
                                 rep=hb['rep'], out=out)                              # H2D proposals, D2H dE

    e2e_counter = [0]

    def e2e_step(bufs, out):
        for s in range(SW):
            e2e_sweep(bufs[e2e_counter[0] % NBH], out)
            e2e_counter[0] += 1
            if re is not None and (s + 1) % XI == 0:
                re._attempt_exchanges()

    out = e2e_sweep(host_pinned[0], dE_host)
    assert np.array_equal(out, dE_ref.cpu().numpy()), 'host-buffer path disagrees with the device path'
    for _ in range(min(W, 3)):
        e2e_step(host_pinned, dE_host)
    if re is not None:
        rc0, rt0, rg0 = re.comm.collective_seconds, re.comm.transport_seconds, re.comm.gathers
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step(host_pinned, dE_host)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * T * SW * K / e2e_s
    e2e_exchange = None
    if re is not None:
        e2e_exchange = {'attempts': K * (SW // XI),
                        'collective_us_per_exchange': 1e6 * (re.comm.collective_seconds - rc0) / max(1, re.comm.gathers - rg0),
                        'transport_s_total': re.comm.transport_seconds - rt0}
    # the same sweeps double-buffered by the CALLER: two B200Calculator instances (two engines, each with its own streams
    # and staging) driven from two host threads, so that the upload of sweep k+1 runs while the kernels of sweep k do
    # (ctypes releases the GIL during the C-ABI calls).  Exchange attempts stay on the main thread between groups.
    e2e_db = None
    try:
        if world > 1:
            # informational arm with collectives inside: kept to the single-GPU run, where no rank can be left waiting
            raise RuntimeError('measured at one GPU only')
        from concurrent.futures import ThreadPoolExecutor
        calc_b = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=RC, device=local)
        calcs = (calc, calc_b)
        outs = (dE_host, pinned_empty(T, np.float64))

        def db_sweep(which, hb):
            calcs[which].set_configurations(replicas_atoms())
            return calcs[which].engine.trial_delta_e(hb['old_off'], hb['old_idx'], hb['new_off'], hb['new_pos'], hb['new_Z'],
                                                     rep=hb['rep'], out=outs[which])
        with ThreadPoolExecutor(max_workers=2) as pool:
            def db_step():
                pending = []
                for s in range(SW):
                    if len(pending) == 2:
                        pending.pop(0).result()
                    pending.append(pool.submit(db_sweep, s & 1, host_pinned[e2e_counter[0] % NBH]))
                    e2e_counter[0] += 1
                    if re is not None and (s + 1) % XI == 0:
                        while pending:
                            pending.pop(0).result()
                        re._attempt_exchanges()
                while pending:
                    pending.pop(0).result()
            first = db_sweep(0, host_pinned[0]).copy()                # (the replicas have been through exchanges since dE_ref)
            assert np.array_equal(db_sweep(1, host_pinned[0]), first), 'second engine disagrees'
            db_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                db_step()
            torch.cuda.synchronize()
            db_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
        e2e_db = {'value': world * T * SW * K / db_s, 'ms_per_sweep': 1e3 * db_s / (K * SW),
                  'note': 'the same calls and copies per sweep, issued from two host threads over two B200Calculator '
                          'instances: H2D of one sweep overlaps the kernels of the other'}
        calc_b.engine.close()
    except Exception as exc:                                          # informational: never take the headline down
        e2e_db = {'note': str(exc)} if world > 1 else {'error': f'{type(exc).__name__}: {exc}'}
    # the same sweeps with ordinary (pageable) numpy arrays, a shorter sample
    KP = max(2, K // 5)
    e2e_step(host_pageable, None)
    barrier()
    t0 = time.perf_counter()
    for _ in range(KP):
        e2e_step(host_pageable, None)
    torch.cuda.synchronize()
    e2e_pageable_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d_sweep = R * (N_ATOMS * (24 + 4)) + 256 + sum(int(v.nbytes) for v in host_pinned[0].values())
    d2h_sweep = 8 * T

    # ---- driver mode: the unmodified replica-exchange loop, one trial per replica per batched call
    driver = None
    if re is not None and not args.no_driver:
        n_steps = args.driver_steps
        for r in re.replicas:
            r.initialize_run()
        re._rebatch_initial_energies()
        g0, c0d, t0d, s0 = re.comm.gathers, re.comm.collective_seconds, re.comm.transport_seconds, re.comm.transports
        paths0 = list(calc.path_counts)
        barrier()
        t0 = time.perf_counter()
        for st in range(n_steps):
            re._batched_gcmc_step()
            if st > 0 and st % XI == 0:
                re._attempt_exchanges()
            re._re_step += 1
        torch.cuda.synchronize()
        drv_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        for r in re.replicas:
            r.finalize_run()
        driver = {'value': world * R * n_steps / drv_s, 'unit': 'replica trial-move dE evals/s',
                  'steps': n_steps, 'ms_per_batched_call': 1e3 * drv_s / n_steps,
                  'collective_us_per_exchange': 1e6 * (re.comm.collective_seconds - c0d) / max(1, re.comm.gathers - g0),
                  'transport_us_per_cross_rank_swap': 1e6 * (re.comm.transport_seconds - t0d) / max(1, re.comm.transports - s0),
                  'exchange_attempts': re.comm.gathers - g0 if world > 1 else (n_steps - 1) // XI,
                  'paths_full_incremental_identical': [a - b for a, b in zip(calc.path_counts, paths0)],
                  'api': 'ShardedReplicaExchange (subclass of the unmodified BatchedReplicaExchange) over unmodified '
                         'GrandCanonicalEnsemble replicas + B200Calculator.get_potential_energies; host time is the '
                         "reference's own Python (snapshots, moves, acceptance, wrap)"}

    # ---- resident mode: the same replica-exchange run with every replica's trials on the device (row f2)
    resident = None
    if re is not None and not args.no_resident:
        try:
            from mcpy_b200.ensembles import ResidentGCMC, ResidentReplicaExchange
            n_res = args.resident_steps
            rre = ResidentReplicaExchange(factory, calc, mus=[{'H': float(m)} for m in mus], gcmc_steps=10 ** 9,
                                          exchange_interval=XI, write_out_interval=10 ** 9, seed=31, rank=rank,
                                          world_size=world, device=f'cuda:{local}', outfile=None,
                                          global_minimum_file=None, engine=eng)
            for r in rre.replicas:
                r.initialize_run()
            rre._rebatch_initial_energies()

            def resident_steps(n):
                xi = rre.exchange_interval
                for _ in range(n):
                    rre._batched_gcmc_step()
                    if rre._re_step > 0 and rre._re_step % xi == 0:
                        rre._attempt_exchanges()
                    rre._re_step += 1
            resident_steps(5 * XI)                                   # warm-up: first upload, first exchanges
            g0r, c0r = rre.comm.gathers, rre.comm.collective_seconds
            dev0, blocks0 = rre.device_ms, rre.blocks
            rre.resident.sync_host()                                 # (the selector's counters come back with a full sync)
            acc0 = sum(sum(r.move_selector.move_acceptance_total) for r in rre.replicas)
            swaps0 = sum(rre.global_successes)
            barrier()
            t0 = time.perf_counter()
            resident_steps(n_res)
            torch.cuda.synchronize()
            res_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
            launches_resident = rre.blocks - blocks0
            dev_res = rre.device_ms - dev0
            swaps_res = sum(rre.global_successes) - swaps0
            rre.resident.sync_host()
            moves_res = sum(sum(r.move_selector.move_acceptance_total) for r in rre.replicas) - acc0
            gathers_res, coll_res = rre.comm.gathers - g0r, rre.comm.collective_seconds - c0r
            # the same run with an exchange attempt every 100 steps: blocks of 100 trials per launch
            rre.exchange_interval = 10 * XI
            rre._re_step = 0
            resident_steps(20 * XI)
            dev1 = rre.device_ms
            barrier()
            t0 = time.perf_counter()
            resident_steps(n_res)
            torch.cuda.synchronize()
            res100_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
            dev100 = rre.device_ms - dev1
            rre.resident.sync_host()
            # one chain alone: what a single GrandCanonicalEnsemble run gets (ResidentGCMC.run_trials)
            one = ResidentGCMC([factory(mu={'H': float(mus[0])}, rank=rank * R)], engine=eng)
            one.advance(500, sync='scalars')
            t0 = time.perf_counter()
            one_ms = one.advance(args.resident_chain_trials, sync='scalars')
            one_wall = time.perf_counter() - t0
            cyc = one.cycles(0)
            one.sync_host()
            one.close()
            # BASELINE config 1 as a sweep: one chain per SM, each at its own chemical potential, ONE launch per block
            sweep = None
            if rank == 0 and not args.no_extra:
                n_ch = 148
                sweep_mus = np.linspace(MU_LADDER[0], MU_LADDER[1], n_ch)
                many = ResidentGCMC([factory(mu={'H': float(m)}, rank=k % R) for k, m in enumerate(sweep_mus)], engine=eng)
                many.advance(300, sync='scalars', counters=False)
                t0 = time.perf_counter()
                many_ms = many.advance(args.resident_sweep_trials, sync='scalars', counters=False)
                many_wall = time.perf_counter() - t0
                many.sync_host()
                sweep = {'chains': n_ch, 'trials_per_chain': args.resident_sweep_trials,
                         'device_trials_per_s': n_ch * args.resident_sweep_trials / (1e-3 * many_ms),
                         'wall_trials_per_s': n_ch * args.resident_sweep_trials / many_wall,
                         'mean_atoms': float(np.mean([len(g.atoms) for g in many.ensembles])),
                         'note': 'density-vs-mu sweep of the 2 000-atom LJ box (BASELINE configs[1]): 148 independent '
                                 'GrandCanonicalEnsemble chains, a CTA (an SM) each'}
                many.close()
            resident = {
                'value': world * R * n_res / res_s, 'unit': 'replica trial-move dE evals/s (whole trials: proposal, dE, '
                                                             'acceptance, commit)',
                'steps': n_res, 'trials_per_replica': n_res, 'wall_s': res_s,
                'device_ms': dev_res, 'launches': launches_resident,
                'us_per_trial_per_replica_device': 1e3 * dev_res / n_res,
                'host_share': 1.0 - 1e-3 * dev_res / res_s,
                'exchange_every_100_steps': {'value': world * R * n_res / res100_s, 'wall_s': res100_s, 'device_ms': dev100,
                                             'us_per_trial_per_replica_device': 1e3 * dev100 / n_res,
                                             'host_share': 1.0 - 1e-3 * dev100 / res100_s},
                'exchange_attempts': gathers_res if world > 1 else n_res // XI,
                'accepted_swaps': swaps_res,
                'collective_us_per_exchange': 1e6 * coll_res / max(1, gathers_res),
                'accepted_moves': moves_res,
                'speedup_vs_driver': (world * R * n_res / res_s) / driver['value'] if driver else None,
                'single_chain': {'trials': args.resident_chain_trials, 'device_trials_per_s': args.resident_chain_trials / (1e-3 * one_ms),
                                 'wall_trials_per_s': args.resident_chain_trials / one_wall,
                                 'device_us_per_trial': 1e3 * one_ms / args.resident_chain_trials,
                                 'serial_share_of_cycles': cyc['serial'] / max(1, cyc['total']),
                                 'speedup_vs_dropin_calls': None},
                'mu_sweep': sweep,
                'api': 'ResidentReplicaExchange (subclass of the unmodified BatchedReplicaExchange): the unmodified '
                       'GrandCanonicalEnsemble replicas hand their state to one CTA each (csrc/gcmc_chain.cuh: MT19937 / '
                       'PCG64 draws, proposal, LJ dE, acceptance, wrap, commit on the device, bit-identical decisions), '
                       'one launch per block of steps between two exchange attempts, swaps of local replicas on the '
                       'device, records all_gathered as in the driver arm'}
            rre.resident.close()
        except Exception as exc:                                     # the arm must not take the headline down with it
            resident = {'error': f'{type(exc).__name__}: {exc}'}

    # ---- drop-in chain: one get_potential_energy(atoms) per trial, sequential (a single GCMC run)
    n_chain = args.chain
    chain_cfgs = [trial_config(cfgs[:1], {k: (v[:TRIALS_PER_REPLICA + 1] if k.endswith('off') else v) for k, v in batches[0].items()}, t)
                  for t in range(min(n_chain, 256))]

    class _A1:
        pass
    a = _A1()
    a.cell, a.pbc = cfgs[0]['cell'], cfgs[0]['pbc']
    calc.path_counts = [0, 0, 0]
    for p1, z1, _, _ in chain_cfgs[:8]:
        a.positions, a.numbers = p1, z1
        calc.get_potential_energy(a)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_chain):
        a.positions, a.numbers = chain_cfgs[i % len(chain_cfgs)][:2]
        calc.get_potential_energy(a)
    chain_s = max_over_ranks(time.perf_counter() - t0)
    dropin_value = world * n_chain / chain_s
    if resident and 'single_chain' in resident:
        resident['single_chain']['speedup_vs_dropin_calls'] = resident['single_chain']['wall_trials_per_s'] / (n_chain / chain_s)

    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary workloads of the same path (rank 0 only, untimed by the driver's metric)
    extra = {}
    if rank == 0 and not args.no_extra:
        extra = secondary_workloads(local)

    # ---- roofline of the trial kernel (FP64 pipe)
    sites = n_old[0] + n_new[0]
    alg_flop = sites * ALG_CANDIDATES_PER_SITE * ALG_FLOP_PER_CANDIDATE + pairs_per_sweep * ALG_FLOP_PER_PAIR
    peak = eng.fp64_peak_tflops()
    rows_ms = stage_ms['row_sums']
    achieved = alg_flop / (rows_ms * 1e-3) / 1e12
    pair_rate = sum_over_ranks(pairs_per_sweep) * SW / (ms_per_step * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        use_ase = real_ase() is not None
        cores = os.cpu_count() or 1
        n_sample = 4 * cores
        with mp.get_context('fork').Pool(cores) as pool:
            cpu_reference_rate(cfgs, batches[0], cores, pool, cores, use_ase)          # warm the workers
            rate, dt = cpu_reference_rate(cfgs, batches[0], n_sample, pool, cores, use_ase, first=8209)
        kind, how = cpu_kind(use_ase)
        cpu_baseline = {
            'value': rate, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': f'{n_sample} proposals of the shard, one full-energy evaluation each ({how}), '
                      f'{cores} processes, {dt:.1f} s wall'}

    prof = _profile_json('traffic.json')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'ms_per_step_median_rank0': float(np.median(step_ms)), 'ms_per_sweep': ms_per_step / SW,
        'timed_region_s': total_ms * 1e-3,
        'exchange': exchange,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_sweep * SW, 'd2h_bytes_per_step': d2h_sweep * SW,
                'api': 'B200Calculator.set_configurations(list of Atoms) + Engine.trial_delta_e (mcpyb_upload_host + '
                       'mcpyb_trial_delta_e_host) per sweep, proposals and results in page-locked numpy arrays '
                       '(mcpy_b200.pinned_copy / pinned_empty); ShardedReplicaExchange._attempt_exchanges over the '
                       'unmodified GrandCanonicalEnsemble replicas every 10 sweeps',
                'ms_per_step': 1e3 * e2e_s / K, 'ms_per_sweep': 1e3 * e2e_s / (K * SW), 'timed_region_s': e2e_s,
                'exchange': e2e_exchange, 'note': e2e_note,
                'double_buffered': e2e_db,
                'pageable': {'value': world * T * SW * KP / e2e_pageable_s,
                             'ms_per_sweep': 1e3 * e2e_pageable_s / (KP * SW), 'steps': KP,
                             'note': 'same calls with ordinary numpy arrays: staged through pinned memory inside the engine'}},
        'driver': driver,
        'resident': resident,
        'dropin': {'value': dropin_value, 'unit': UNIT, 'calls': n_chain,
                   'api': 'calculator.get_potential_energy(atoms) per trial, sequential '
                          '(mcpyb_tracked_energy_host: resident base + one delta; the trial rides in the '
                          'kernel parameters, the result lands in pinned host memory)',
                   'paths': {'full': calc.path_counts[0], 'incremental': calc.path_counts[1],
                             'identical': calc.path_counts[2]},
                   'us_per_call': 1e6 * chain_s / n_chain},
        'stages_ms': {k: round(v, 5) for k, v in stage_ms.items()},
        'pair_interactions_per_s': pair_rate,
        'pairs_per_sweep': pairs_per_sweep,
        'gpu_launches': int(gpu_launches),
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak if peak else None,
                     'traffic': prof.get('lj_trial_rows_direct_kernel_dram_bytes_per_launch',
                                         prof.get('lj_trial_rows_warp_kernel_dram_bytes_per_launch')),
                     'pipe_fp64_pct': prof.get('lj_trial_rows_direct_kernel_pipe_fp64_pct',
                                               prof.get('lj_trial_rows_warp_kernel_pipe_fp64_pct')),
                     'kernel': 'lj_trial_rows_direct_kernel<4> (the warp-per-item row sums of rows_warp.cuh, work items '
                               'announced by trial_sites_direct_kernel)',
                     'kernel_ms': rows_ms, 'kernel_share_of_sweep': rows_ms / sum(stage_ms.values()),
                     'algorithmic_flop_per_launch': alg_flop,
                     'algorithmic_bytes_per_launch': int(sites * 48 + sites * 2 * 12 + R * N_ATOMS * 32),
                     'traffic_note': 'site records in (48 B) + two partial sums per site out (12 B each) + the configurations; '
                                     'the ncu figure (cold caches, serialised launches) adds the 35 MB of per-configuration '
                                     'stencil lists, which stay in the 126 MB L2 between the launches of a live run; either '
                                     'way DRAM runs at < 10 % of its peak under this FP64-bound kernel',
                     'step_achieved': alg_flop * SW / (ms_per_step * 1e-3) / 1e12,
                     'peak_source': 'measured here: DFMA microbenchmark (mcpyb_fp64_peak_tflops); '
                                    'MEASURED_PEAKS.json holds only HBM and bf16 figures',
                     'note': 'HBM is not the bound: each 64 KB configuration is L1/L2 resident; algorithmic flop '
                             'follow SURVEY 8d (27-cell stencil at cell edge rc: 437 candidates x 17 flop per site '
                             '+ 8 per in-cutoff pair); the kernel itself visits fewer candidates after pruning -- '
                             'pipe_fp64_pct is the executed FP64-pipe utilisation of the committed ncu capture '
                             '(profiles/)'},
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
    bts = [batch_of(t) for t in range(8)]
    res = {}
    for label, inc in (('full_recompute', False), ('resident_bases', True)):
        calc = B200Calculator('lj', epsilon=1.0, sigma=1.0, rc=RC, device=device, incremental=inc)
        for bt in bts[:2]:
            calc.get_potential_energies(bt)
        t0 = time.perf_counter()
        for rep_ in range(3):
            for bt in bts:
                res[label + '_E'] = calc.get_potential_energies(bt)
        res[label] = (time.perf_counter() - t0) / (3 * len(bts))
        if inc:
            res['paths'] = list(calc.path_counts)
            t_1 = _time(lambda: [calc.get_potential_energy(a) for a in bts[0][:8]], 3) / 8
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
    out['kernel_rooflines'] = kernel_rooflines(device)
    out['fp32_pair_math'] = fp32_mode(device)
    return out


def fp32_mode(device):
    """north_star's second precision for the batched LJ row sums (mcpyb_set_precision): the same sweep in fp64 and
    with fp32 pair math + fp64 accumulation; error of dE relative to the fp64 mode and the row-sum stage time."""
    from mcpy_b200 import Engine
    cfgs = local_configs(0)
    tb = proposal_batch(cfgs, seed=5)
    res = {}
    dE = {}
    for prec in ('fp64', 'fp32'):
        eng = Engine(device)
        eng.set_lj(1.0, 1.0, RC)
        eng.set_precision(prec)
        eng.upload([c['positions'] for c in cfgs], [c['numbers'] for c in cfgs],
                   np.stack([c['cell'].reshape(9) for c in cfgs]), np.stack([c['pbc'] for c in cfgs]))
        call = lambda: eng.trial_delta_e(tb['old_off'], tb['old_idx'], tb['new_off'], tb['new_pos'], tb['new_Z'], rep=tb['rep'])  # noqa: E731
        for _ in range(3):
            dE[prec] = call()
        eng.stage_timing(True)
        for _ in range(6):
            call()
            st, _ = eng.stage_timing(True)
        res[prec + '_row_sums_ms'] = st['row_sums']
        eng.close()
    rel = np.abs(dE['fp32'] - dE['fp64']) / np.maximum(1.0, np.abs(dE['fp64']))
    res['max_rel_error_vs_fp64'] = float(rel.max())
    res['tolerance'] = 1e-5
    res['note'] = ('fp32 r2 / reciprocal / c6 from an fp64 distance vector, class sums added in fp32 per trip and accumulated '
                   'in fp64; membership bit-exact.  On B200 the mode is SLOWER than fp64: the three f64->f32 conversions per '
                   'candidate issue at a quarter of the DFMA rate, and the FP64 pipe is only 2x narrower than the FP32 one')
    return res


def kernel_rooflines(device):
    """The other kernels north_star names, each against the bound that applies, timed with CUDA events
    on the engine's stream with their inputs resident in HBM; peaks: HBM copy bandwidth from MEASURED_PEAKS.json
    (fallback: the profiling recipe's 6 551.7 GB/s), FP64 FMA rate measured in this process."""
    import torch

    from mcpy_b200 import Engine
    from mcpy_b200.utils import synthetic
    try:
        hbm = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        hbm_src = 'MEASURED_PEAKS.json (burst: the kernel is timed alone)'
    except Exception:
        hbm, hbm_src = 6551.7, 'fallback of /opt/skills/guides/B200_PROFILING.md'
    dev = torch.device('cuda', device)
    stream = torch.cuda.Stream(device=dev)
    res = {}

    def gpu_ms(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(reps):
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            times.append(a.elapsed_time(b))
        return float(np.median(times))

    with torch.cuda.stream(stream):
        # K1: counting-sort cell list of a 64 x 2 000-atom replica batch (72 B / atom: SURVEY 8d)
        eng = Engine(device)
        eng.set_stream(stream.cuda_stream)
        eng.set_lj(1.0, 1.0, RC)
        cfgs = [synthetic.s2_lj(N_ATOMS, RC, RHO, seed=100 + k) for k in range(64)]
        pos = torch.from_numpy(np.concatenate([c['positions'] for c in cfgs])).to(dev)
        num = torch.from_numpy(np.concatenate([c['numbers'] for c in cfgs]).astype(np.int32)).to(dev)
        off = np.arange(65, dtype=np.int64) * N_ATOMS
        cells = np.stack([c['cell'].reshape(9) for c in cfgs])
        pbcs = np.stack([c['pbc'] for c in cfgs])
        ms = gpu_ms(lambda: eng.upload_dev(off, pos.data_ptr(), num.data_ptr(), cells, pbcs))
        nbytes = 64 * N_ATOMS * 72
        res['K1_build_cell_list_64x2000'] = {
            'bound': 'hbm', 'achieved': nbytes / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
            'frac': nbytes / (ms * 1e-3) / 1e9 / hbm, 'kernel_ms': ms, 'algorithmic_bytes_per_launch': nbytes,
            'peak_source': hbm_src,
            'note': 'one CTA per replica (64 CTAs on 148 SMs): the launch is latency-bound at this size, 9 MB cannot '
                    'fill HBM; timed as Engine.upload_dev (48-byte header H2D + the build launch)'}
        # K2: total energies of the same batch (FP64 pipe; 17 flop per stencil candidate + 8 per in-cutoff pair)
        d_E = torch.empty(64, dtype=torch.float64, device=dev)
        ms = gpu_ms(lambda: eng.energies_dev(d_E.data_ptr()))
        E, pairs = eng.energies(return_pairs=True)
        peak64 = eng.fp64_peak_tflops()
        flop = 64 * N_ATOMS * ALG_CANDIDATES_PER_SITE * ALG_FLOP_PER_CANDIDATE + 2 * int(pairs.sum()) * ALG_FLOP_PER_PAIR
        res['K2_lj_energies_64x2000'] = {
            'bound': 'fp64', 'achieved': flop / (ms * 1e-3) / 1e12, 'peak': peak64, 'unit': 'TFLOP/s',
            'frac': flop / (ms * 1e-3) / 1e12 / peak64, 'kernel_ms': ms, 'algorithmic_flop_per_launch': flop,
            'pairs': int(pairs.sum())}
        eng.close()
        # K4 / K5: EMT pair + embedding pass on the 10 825-atom CuPd + CO particle (FP64 + XU: 55 flop + 6
        # transcendentals per in-cutoff pair, 15 flop + 3 per atom; the stencil walk 17 flop per candidate)
        eng = Engine(device)
        eng.set_stream(stream.cuda_stream)
        eng.set_emt()
        c5 = synthetic.s5_cupd_co()
        n5 = len(c5['positions'])
        p5 = torch.from_numpy(c5['positions']).to(dev)
        z5 = torch.from_numpy(c5['numbers'].astype(np.int32)).to(dev)
        eng.upload_dev(np.array([0, n5], dtype=np.int64), p5.data_ptr(), z5.data_ptr(), c5['cell'].reshape(1, 9),
                       c5['pbc'].reshape(1, 3))
        d_E1 = torch.empty(1, dtype=torch.float64, device=dev)
        ms = gpu_ms(lambda: eng.energies_dev(d_E1.data_ptr()))
        _, pairs5 = eng.energies(return_pairs=True)
        npair = int(pairs5[0])
        rho5 = n5 / (4.0 / 3.0 * np.pi * (c5['sphere_radius'] - 3.5) ** 3)
        cand = 27 * rho5 * eng.cutoff ** 3
        flop = n5 * cand * 17 + 2 * npair * (55 + 6 * 20) + n5 * (15 + 3 * 20)
        res['K4_K5_emt_energy_S5'] = {
            'bound': 'fp64', 'achieved': flop / (ms * 1e-3) / 1e12, 'peak': peak64, 'unit': 'TFLOP/s',
            'frac': flop / (ms * 1e-3) / 1e12 / peak64, 'kernel_ms': ms, 'algorithmic_flop_per_launch': flop,
            'pairs': npair,
            'note': 'exp / log / sqrt counted at 20 FP64-pipe flop each (their polynomial bodies run on the FP64 pipe; '
                    'only the seed comes from the XU): emt_pair_kernel + emt_embed_kernel + reduce'}
        # K6: EMT trial-move dE on the same particle, the trial batch resident in HBM (list-driven warp-per-trial
        # kernel, emt_rows.cuh).  Work per trial, counted from the pair counter the kernel returns: every directed
        # pair term is 2 exp + a shared cutoff exp (~55 flop + 3 x 20), every touched neighbour one re-embedding
        # (log + 2 exp, ~15 flop + 3 x 20); the list walk is 9 flop per candidate.
        res['K6_emt_trial_dE_S5'] = {}
        for T in (8192, 65536):
            tb = synthetic.trial_batch(c5, T, seed=11, insert_Z=8)
            dt = {k: torch.from_numpy(np.ascontiguousarray(tb[k], dtype=np.int32)).to(dev) for k in ('old_off', 'old_idx', 'new_off', 'new_Z')}
            dnp = torch.from_numpy(np.ascontiguousarray(tb['new_pos'], dtype=np.float64)).to(dev)
            d_dE = torch.empty(T + 1, dtype=torch.float64, device=dev)
            d_np = torch.empty(T, dtype=torch.int32, device=dev)
            n_old, n_new = int(tb['old_off'][-1]), int(tb['new_off'][-1])

            def tr():
                eng.trial_delta_e_dev(T, 0, dt['old_off'].data_ptr(), dt['old_idx'].data_ptr(), dt['new_off'].data_ptr(),
                                      dnp.data_ptr(), dt['new_Z'].data_ptr(), n_old, n_new, d_dE.data_ptr(), d_np.data_ptr())
            ms = gpu_ms(tr)
            terms = int(d_np.sum().item())
            sites = n_old + n_new
            flop = terms * (55 + 3 * 20) + sites * 54 * (15 + 3 * 20) + sites * 200 * 9
            res['K6_emt_trial_dE_S5'][f'T{T}'] = {
                'bound': 'fp64', 'achieved': flop / (ms * 1e-3) / 1e12, 'peak': peak64, 'unit': 'TFLOP/s',
                'frac': flop / (ms * 1e-3) / 1e12 / peak64, 'kernel_ms': ms, 'dE_evals_per_s': T / (ms * 1e-3),
                'pair_terms': terms, 'algorithmic_flop_per_launch': flop}
        res['K6_emt_trial_dE_S5']['note'] = (
            'memset + emt_trial_rows_kernel + emt_trial_redo_kernel per call; transcendentals at 20 FP64-pipe flop '
            'each; the first-generation kernel (MCPYB_EMT_LISTS=0) is the A/B baseline in profiles/README.md')
        # K7: covered test with the sample points streamed from HBM (24 B per point), S5 geometry
        P = 1_000_000
        rng = np.random.default_rng(3)
        pts = torch.from_numpy(synthetic.sphere_points(rng, P, c5['sphere_radius'])).to(dev)
        rad = torch.from_numpy(c5['radii']).to(dev)
        cnt = torch.zeros(2, dtype=torch.int64, device=dev)

        def fv():
            cnt.zero_()
            eng.free_volume_count_dev(pts.data_ptr(), P, p5.data_ptr(), rad.data_ptr(), n5, cnt.data_ptr())
        ms = gpu_ms(fv)
        nbytes = 24 * P
        res['K7_free_volume_count_S5_1M_points_from_hbm'] = {
            'bound': 'hbm', 'achieved': nbytes / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
            'frac': nbytes / (ms * 1e-3) / 1e9 / hbm, 'kernel_ms': ms, 'algorithmic_bytes_per_launch': nbytes,
            'points_per_s': P / (ms * 1e-3), 'peak_source': hbm_src,
            'note': 'fv_build (spheres binned) + fv_count per call; 24 MB of points against ~0.36 kflop of gather + '
                    'compare work per point (SURVEY 8d): the kernel is bound by the L1/L2 gather of the sphere cells, '
                    'not by streaming the points'}
        # K7 + K8: the same count with the points regenerated from the PCG64 stream on the device (no HBM traffic:
        # integer-ALU bound, 3 x 128-bit LCG steps per candidate point)
        rng = np.random.default_rng(3)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            eng.sphere_free_volume_pcg64(rng, P, c5['sphere_radius'], np.zeros(3), c5['positions'], c5['radii'])
        dt = (time.perf_counter() - t0) / reps
        res['K7_K8_free_volume_pcg64_S5_1M_points'] = {
            'bound': 'int-alu', 'points_per_s': P / dt, 'pcg64_steps_per_s': 1.5 * 3 * P * 2 / dt, 'call_ms': 1e3 * dt,
            'note': 'host call incl. sphere upload and count readback; candidates = 1.5 P per round, two rounds '
                    '(count + evaluate) regenerate the stream instead of storing 24 MB of points'}
        eng.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--sweeps', type=int, default=80, help='sweeps (batched launches of 8 x 65 536 proposals) per step')
    ap.add_argument('--batches', type=int, default=16, help='distinct proposal batches cycled through (18 MB each)')
    ap.add_argument('--chain', type=int, default=2000, help='sequential drop-in calls to time')
    ap.add_argument('--driver-steps', type=int, default=60, help='steps of the unmodified replica-exchange loop to time')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-driver', action='store_true')
    ap.add_argument('--resident-steps', type=int, default=4000, help='steps of the device-resident replica-exchange run to time')
    ap.add_argument('--resident-chain-trials', type=int, default=50000, help='trials of the single device-resident chain')
    ap.add_argument('--resident-sweep-trials', type=int, default=5000, help='trials per chain of the 148-chain mu sweep')
    ap.add_argument('--no-resident', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary EMT / free-volume numbers')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.sweeps % EXCHANGE_INTERVAL:
        args.sweeps += EXCHANGE_INTERVAL - args.sweeps % EXCHANGE_INTERVAL
    if args.impl == 'reference':
        return run_reference(args)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
