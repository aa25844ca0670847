"""``ResidentGCMC`` -- whole Markov chains of the reference's ``GrandCanonicalEnsemble`` on the device.

The reference advances a chain one trial at a time from Python
(``mcpy/ensembles/grand_canonical_ensemble.py:244-308``): snapshot, ``MoveSelector.do_trial_move``, one calculator
call, ``_acceptance_condition``, wrap / record or roll back.  Here the UNMODIFIED ensemble objects stay the owners of the
run -- atoms, moves, cells, generators, counters, files -- and hand their state to ``csrc/gcmc_chain.cuh`` for a block of
trials: ONE CTA per chain keeps the configuration, its cell list and the five Mersenne-Twister streams in shared memory
and runs proposal, dE, acceptance and commit without a host round trip.  Afterwards the state is written back into the
same objects, so a chain can alternate between :meth:`ResidentGCMC.run` and the reference's own ``gcmc.run`` and sees
one draw sequence, one accept / reject sequence and one trajectory either way (``tests/test_gpu_resident_gcmc.py``).

Scope (``supports`` says why a chain falls outside): one atomic species, Lennard-Jones, an orthorhombic box periodic on
all three axes with every edge > 2 rc, a box ``Cell`` for insertion / deletion, atomic ``InsertionMove`` (no
``min_insert``) / ``DeletionMove`` / ``DisplacementMove`` (``n_steps=1``, nothing constrained), no ``minima_file``.
Everything else keeps going through ``B200Calculator`` one call per trial.

Several ensembles (the replicas of a ladder, a mu sweep) run in one launch, a CTA each.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import GCMC_MAX_MOVES, GCMC_MT_WORDS, GCMC_TRIAL_DTYPE, GcmcDesc, GcmcScalars
from ..engine import Engine

KIND_INSERT, KIND_DELETE, KIND_DISPLACE = 0, 1, 2
STATUS = {0: 'done', 1: 'cell list full', 2: 'atom storage full', 3: 'trace full', 4: 'displacement of an empty configuration'}


def _lj_parameters(calculator, lj):
    if lj is not None:
        eps, sigma, rc = lj
        return float(eps), float(sigma), float(rc)
    p = getattr(calculator, 'lj_parameters', None)                     # B200Calculator('lj')
    if p is None and hasattr(calculator, 'inner'):
        p = getattr(calculator.inner, 'lj_parameters', None)
    if p is None:
        raise ValueError("ResidentGCMC needs the Lennard-Jones parameters: pass lj=(epsilon, sigma, rc) or use a "
                         "B200Calculator('lj') as the ensemble's calculator")
    return float(p['epsilon']), float(p['sigma']), float(p['rc'])


def supports(gcmc, lj=None):
    """``None`` when ``gcmc`` (a reference ``GrandCanonicalEnsemble``) can run resident, else the reason it cannot."""
    from mcpy.cell import Cell
    from mcpy.moves import DeletionMove, DisplacementMove, InsertionMove
    atoms = gcmc.atoms
    cell = np.asarray(atoms.cell, dtype=float)
    if not np.all(np.asarray(atoms.pbc)):
        return 'the box must be periodic along all three axes'
    if np.any(cell - np.diag(np.diag(cell)) != 0.0) or np.any(np.diag(cell) <= 0.0):
        return 'the box must be orthorhombic with a diagonal cell matrix'
    if atoms.constraints:
        return 'constraints are not handled on the device'
    if set(atoms.arrays) - {'numbers', 'positions'}:
        return f'extra per-atom arrays {sorted(set(atoms.arrays) - {"numbers", "positions"})} are not carried'
    if len(gcmc._mu) != 1:
        return 'exactly one chemical potential (one species) is handled'
    species = next(iter(gcmc._mu))
    from ase.data import atomic_numbers
    if species not in atomic_numbers or species in gcmc.units.molecules:
        return f'{species!r} is not an atomic species'
    if np.any(atoms.numbers != atomic_numbers[species]):
        return 'every atom must be of the exchanged species'
    try:
        _, _, rc = _lj_parameters(gcmc._calculator if hasattr(gcmc, '_calculator') else None, lj)
    except ValueError as exc:
        return str(exc)
    if np.any(np.diag(cell) <= 2.0 * rc):
        return 'every box edge must exceed 2 rc (one image per neighbour)'
    if gcmc._minima_file is not None:
        return 'a minima_file wants one write per improvement'
    moves = gcmc.move_selector.move_list
    if not 1 <= len(moves) <= GCMC_MAX_MOVES:
        return f'1..{GCMC_MAX_MOVES} moves'
    for m in moves:
        if type(m) is InsertionMove:
            if m.min_insert is not None:
                return 'InsertionMove.min_insert is not handled on the device'
        elif type(m) is DeletionMove:
            pass
        elif type(m) is DisplacementMove:
            if m.n_steps != 1 or m.constraints.size:
                return 'DisplacementMove with n_steps > 1 or constrained indices is not handled on the device'
            continue
        else:
            return f'{type(m).__name__} is not handled on the device'
        if list(m.species) != [species]:
            return 'moves must exchange the one species of the run'
        if type(m.cell) is not Cell:
            return f'{type(m.cell).__name__}: insertion / deletion must use the box Cell'
        if not np.array_equal(np.asarray(m.cell.dimensions, dtype=float), cell):
            return "the move's Cell must span the atoms' box"
    return None


class ResidentGCMC:
    """Device-resident trials for a list of reference ``GrandCanonicalEnsemble`` objects.

    ``lj``: ``(epsilon, sigma, rc)`` when the ensembles' calculator is not a ``B200Calculator('lj')``.
    ``atom_capacity``: atoms a chain may grow to (default: twice the current count, at least 256).
    ``trace``: keep one record per trial of the last :meth:`run_trials` (``mcpyb_gcmc_trial``), at most this many.
    ``resync_energy``: at every full hand-back, replace the running ``E_old`` by a full energy of the configuration
    through the ensemble's own calculator (the kernel already re-sums it on the device when its rounding bound asks).

    :meth:`run_trials` / :meth:`advance` ``(n, sync='full')`` advance every chain and write the whole state back;
    ``advance(n, sync='scalars')`` brings back only ``E_old``, ``n_atoms``, counters and minimum scores and leaves the chains
    the owners of the state (host ``positions`` are NaN until :meth:`sync_host`, or :meth:`sync_positions` for a frame);
    :meth:`run` is ``gcmc.run(steps)`` with one launch per block between file writes; :func:`install` routes
    ``GrandCanonicalEnsemble.run`` here without a change to the calling script.
    """

    def __init__(self, ensembles, engine: Engine | None = None, lj=None, atom_capacity=None, cell_capacity=0,
                 trace=0, resync_energy=False):
        if not isinstance(ensembles, (list, tuple)):
            ensembles = [ensembles]
        if not ensembles:
            raise ValueError('ResidentGCMC needs at least one ensemble')
        self.ensembles = list(ensembles)
        for k, g in enumerate(self.ensembles):
            why = supports(g, lj)
            if why is not None:
                raise ValueError(f'ensemble {k} cannot run device-resident: {why}')
        self._lib = _lib.load()
        if engine is None:
            calc = self.ensembles[0]._calculator
            engine = getattr(calc, 'engine', None) or getattr(getattr(calc, 'inner', None), 'engine', None) or Engine(0)
        self.engine = engine
        self.resync_energy = bool(resync_energy)
        self.trace_capacity = int(trace)
        n = len(self.ensembles)
        self._desc = (GcmcDesc * n)()
        self._species = []
        self._pcg_cells = []                       # per chain: the Cell objects whose PCG64 streams the chain draws from
        for k, g in enumerate(self.ensembles):
            self._fill_desc(k, g, lj, atom_capacity, cell_capacity)
        h = C.c_void_p()
        self.engine._check(self._lib.mcpyb_gcmc_create(self.engine._h, n, self._desc, self.trace_capacity, C.byref(h)))
        self._h = h
        from ase.data import atomic_numbers
        self._Z = [atomic_numbers[sp] for sp in self._species]
        self._host_fresh = [True] * n              # the ensemble objects hold the current state
        self._device_fresh = [False] * n           # the chain holds the current state
        self._counted = [[[0] * GCMC_MAX_MOVES for _ in range(3)] for _ in range(n)]   # device counters already added on the host
        self._trials_seen = [0] * n
        self._gauss = {}
        self._scalars = GcmcScalars()              # scratch of _apply_scalars
        self._scalars_ref = C.byref(self._scalars)
        self._storage_sig = None
        self._recv_block = None
        self._blk = None
        self.last_status = [0] * n
        self.trials_run = 0

    # ------------------------------------------------------------------ description of a chain
    def _fill_desc(self, k, g, lj, atom_capacity, cell_capacity):
        from mcpy.moves import DeletionMove, InsertionMove
        d = self._desc[k]
        atoms = g.atoms
        species = next(iter(g._mu))
        self._species.append(species)
        box = np.diag(np.asarray(atoms.cell, dtype=float))
        for a in range(3):
            d.box[a] = box[a]
        d.beta = float(g.units.beta)
        d.mu = float(g._mu[species])
        d.lambda3 = float(g.units.lambda_dbs[species] ** 3)
        d.lj_epsilon, d.lj_sigma, d.lj_rc = _lj_parameters(getattr(g, '_calculator', None), lj)
        ms = g.move_selector
        cells = []
        d.n_move_types = len(ms.move_list)
        for m, mv in enumerate(ms.move_list):
            d.move_cum_weight[m] = float(ms.rho[m])
            d.move_limit[m] = -1
            d.move_pcg[m] = 0
            d.move_volume[m] = 0.0
            if type(mv) is InsertionMove:
                d.move_kind[m] = KIND_INSERT
                d.move_volume[m] = float(mv.cell._box_volume)
                if mv.max_atoms is not None:
                    d.move_limit[m] = int(mv.max_atoms)
                if not any(mv.cell is c for c in cells):
                    cells.append(mv.cell)
                d.move_pcg[m] = next(i for i, c in enumerate(cells) if c is mv.cell)
            elif type(mv) is DeletionMove:
                d.move_kind[m] = KIND_DELETE
                d.move_volume[m] = float(mv.cell._box_volume)
                if mv.min_atoms is not None:
                    d.move_limit[m] = int(mv.min_atoms)
            else:
                d.move_kind[m] = KIND_DISPLACE
                d.move_max_displacement[m] = float(mv.max_displacement)
        self._pcg_cells.append(cells)
        n = len(atoms)
        cap = int(atom_capacity) if atom_capacity else max(256, 2 * n)
        d.atom_capacity = min(max(cap, n + 2), 65535)
        d.cell_capacity = int(cell_capacity)

    # ------------------------------------------------------------------ host objects -> device
    def _gens(self, g):
        ms = g.move_selector
        return [ms.rng.random] + [mv.rng.random for mv in ms.move_list] + [g.rng_acceptance.random]

    def push(self, chains=None):
        """Upload the current state of the ensembles (configuration, generators, energy, best) to their chains."""
        for k in (range(len(self.ensembles)) if chains is None else chains):
            g = self.ensembles[k]
            if not self._host_fresh[k]:
                raise RuntimeError(f'ResidentGCMC: chain {k} was advanced on the device and not synchronised back; '
                                   'call sync_host() before touching the ensemble on the host')
            atoms = g.atoms
            ms = g.move_selector
            s = GcmcScalars()
            s.energy = float(g.E_old)
            s.best_score = float(g._best_score)
            s.best_energy = float(g._best_energy)
            s.n_atoms = len(atoms)
            s.best_n = -1
            s.last_move = int(ms.to_use)
            for c, cell in enumerate(self._pcg_cells[k]):
                st = cell._rng.bit_generator.state
                if st.get('bit_generator') != 'PCG64' or st.get('has_uint32', 0):
                    raise ValueError("the Cell's generator must be numpy PCG64 with no buffered 32-bit half")
                hi_lo, inc = Engine._pcg_words(cell._rng)
                s.pcg[c][0], s.pcg[c][1], s.pcg[c][2], s.pcg[c][3] = int(hi_lo[0]), int(hi_lo[1]), int(inc[0]), int(inc[1])
            gens = self._gens(g)
            mt = np.empty((len(gens), GCMC_MT_WORDS), dtype=np.uint32)
            for j, r in enumerate(gens):
                version, words, gauss_next = r.getstate()
                if version != 3 or len(words) != GCMC_MT_WORDS:
                    raise ValueError('unexpected random.Random state layout')
                mt[j] = words
                self._gauss[(k, j)] = gauss_next
            pos = np.ascontiguousarray(atoms.positions, dtype=np.float64)
            self.engine._check(self._lib.mcpyb_gcmc_set_state(self._h, k, pos.ctypes.data if len(pos) else None,
                                                              C.byref(s), mt.ctypes.data))
            self._device_fresh[k] = True
            self._counted[k] = [[0] * GCMC_MAX_MOVES for _ in range(3)]
            self._trials_seen[k] = 0

    # ------------------------------------------------------------------ device -> host objects
    def _apply_scalars(self, k, counters=True):
        """What the drivers read between blocks: ``E_old``, ``n_atoms`` (and an ``Atoms`` of that length), the
        selector's counters and ``to_use``, the running-minimum scores.  Positions and generator states stay on the
        device until :meth:`sync_host`; the host ``positions`` are NaN meanwhile.  ``counters=False`` leaves the
        selector's counters for later (the device's are cumulative since the upload, so nothing is lost)."""
        g = self.ensembles[k]
        s = self._scalars
        self.engine._check(self._lib.mcpyb_gcmc_get_scalars(self._h, k, self._scalars_ref))
        n = int(s.n_atoms)
        atoms = g.atoms
        if self._host_fresh[k] or len(atoms) != n:
            atoms.arrays = {'numbers': np.full(n, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                            'positions': np.full((n, 3), np.nan)}
        self._host_fresh[k] = False
        g.n_atoms = n
        g.E_old = s.energy
        if s.best_n >= 0:                                              # improved since the upload
            g._best_score = s.best_score
            g._best_energy = s.best_energy
        if not counters:
            return s
        ms = g.move_selector
        ms.to_use = int(s.last_move)
        seen = self._counted[k]
        for m in range(self._desc[k].n_move_types):
            for c, (src, interval, total) in enumerate(((s.attempts, ms.move_counter, ms.move_counter_total),
                                                        (s.accepted, ms.move_acceptance, ms.move_acceptance_total),
                                                        (s.failed, ms.move_failed_counter, ms.move_failed_counter_total))):
                d = int(src[m]) - seen[c][m]
                if d:
                    interval[m] += d
                    total[m] += d
                    seen[c][m] += d
        return s

    def _apply_block_scalars(self):
        """:meth:`_apply_scalars` without the counters, for all chains from ONE call: ``E_old``, ``n_atoms``, the
        running-minimum scores of chains that improved.  Returns the trials each chain did in the last launch."""
        b = self._blk
        if b is None:
            n = len(self.ensembles)
            b = self._blk = (np.empty(n), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64), np.empty(2 * n))
            self._blk_ptrs = tuple(a.ctypes.data for a in b)
        self.engine._check(self._lib.mcpyb_gcmc_get_block_scalars(self._h, *self._blk_ptrs))
        energy, n_atoms, trials, best = b[0].tolist(), b[1].tolist(), b[2].tolist(), b[3].tolist()
        seen = self._trials_seen
        done = np.empty(len(energy), dtype=np.int64)
        fresh = self._host_fresh
        for k, g in enumerate(self.ensembles):
            n = n_atoms[k]
            if fresh[k] or g.n_atoms != n or len(g.atoms) != n:
                atoms = g.atoms
                atoms.arrays = {'numbers': np.full(n, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                                'positions': np.full((n, 3), np.nan)}
                fresh[k] = False
                g.n_atoms = n
            g.E_old = energy[k]
            if best[2 * k] < g._best_score:                            # improved on the device (never worse than the host's)
                g._best_score, g._best_energy = best[2 * k], best[2 * k + 1]
            done[k] = trials[k] - seen[k]
            seen[k] = trials[k]
        return done

    def sync_host(self, chains=None):
        """Bring the ensembles fully up to date with their chains: configuration, the five Mersenne-Twister states and
        the cells' PCG64 states, the best configuration.  The ensembles own the state afterwards: whatever the host does
        to them (a block of the reference's own loop, a swap) is uploaded again before the next device block."""
        for k in (range(len(self.ensembles)) if chains is None else chains):
            if self._host_fresh[k]:
                continue
            g = self.ensembles[k]
            d = self._desc[k]
            cap = d.atom_capacity
            nm = d.n_move_types
            pos = np.empty((cap, 3), dtype=np.float64)
            best = np.empty((cap, 3), dtype=np.float64)
            mt = np.empty((nm + 2, GCMC_MT_WORDS), dtype=np.uint32)
            s = GcmcScalars()
            self.engine._check(self._lib.mcpyb_gcmc_get_state(self._h, k, pos.ctypes.data, C.byref(s), mt.ctypes.data,
                                                              best.ctypes.data))
            self._apply_scalars(k)
            n = int(s.n_atoms)
            atoms = g.atoms
            # same object, new arrays: what the reference's in-place moves leave behind
            atoms.arrays = {'numbers': np.full(n, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                            'positions': pos[:n].copy()}
            for j, r in enumerate(self._gens(g)):
                r.setstate((3, tuple(mt[j].tolist()), self._gauss[(k, j)]))
            for c, cell in enumerate(self._pcg_cells[k]):
                st = cell._rng.bit_generator.state
                st['state']['state'] = (int(s.pcg[c][0]) << 64) | int(s.pcg[c][1])
                cell._rng.bit_generator.state = st
            if s.best_n >= 0:
                bn = int(s.best_n)
                ba = atoms.copy()                                      # cell, pbc, info ... as `atoms.copy()` in _record_minimum
                ba.arrays = {'numbers': np.full(bn, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                             'positions': best[:bn].copy()}
                g._best_atoms = ba
            if self.resync_energy:
                g.E_old = float(g.compute_energy(atoms))
            self._host_fresh[k] = True
            self._device_fresh[k] = False          # the ensemble owns the state again (the host may change it)

    pull = sync_host

    def sync_positions(self, chains=None):
        """Only what a trajectory frame needs: the configurations (and the scalars / counters) of the chains, which stay
        the owners of the state -- generator states and the best configuration are left on the device, nothing is
        uploaded again before the next block."""
        for k in (range(len(self.ensembles)) if chains is None else chains):
            if self._host_fresh[k]:
                continue
            g = self.ensembles[k]
            pos = np.empty((self._desc[k].atom_capacity, 3), dtype=np.float64)
            s = GcmcScalars()
            self.engine._check(self._lib.mcpyb_gcmc_get_state(self._h, k, pos.ctypes.data, C.byref(s), None, None))
            self._apply_scalars(k)
            n = int(s.n_atoms)
            atoms = g.atoms
            atoms.arrays = {'numbers': np.full(n, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                            'positions': pos[:n].copy()}

    def swap_configurations(self, i, j):
        """``_swap_states`` of the replica drivers between chains ``i`` and ``j`` on the device (configuration, energy,
        particle number); the caller swaps the ensembles' ``atoms`` / ``E_old`` / ``n_atoms``.  Chains that are not both
        device-fresh, or whose storage differs, are synchronised and re-uploaded instead."""
        if self._device_fresh[i] and self._device_fresh[j] and self._same_storage(i, j):
            self.engine._check(self._lib.mcpyb_gcmc_swap_configurations(self._h, int(i), int(j)))
            return True
        self.sync_host([i, j])
        return False

    # ------------------------------------------------------------------ swaps across two processes, device to device
    class _DeviceBytes:
        """``__cuda_array_interface__`` over a device address (what ``torch.as_tensor`` needs to wrap it)."""

        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                             'version': 2, 'strides': None}

    def config_block(self, k):
        """(device address, bytes, the five scalars that go with it) of chain ``k``'s configuration block."""
        ptr, nbytes = C.c_void_p(), C.c_int64(0)
        sc = (C.c_double * 5)()
        self.engine._check(self._lib.mcpyb_gcmc_config_ptr(self._h, int(k), C.byref(ptr), C.byref(nbytes), sc))
        return ptr.value, nbytes.value, list(sc)

    def remote_swap(self, k, comm, peer):
        """Swap chain ``k``'s configuration with a chain of rank ``peer`` (which calls this too) without passing
        through the host: a 12-number header each way (can this side do it, storage layout, the scalars of the
        configuration), then the configuration blocks themselves, device to device over ``comm`` (NCCL).  Returns
        False -- on BOTH sides, both see both headers -- when either side cannot (its chain is not device-fresh, its
        ``Atoms`` carries ``info`` the block does not hold, the layouts differ): the caller then ships the ``Atoms``."""
        import torch
        g = self.ensembles[k]
        dev = torch.device(comm.device)
        able = bool(self._device_fresh[k]) and not g.atoms.info
        ptr, nbytes, sc = self.config_block(k)
        d = self._desc[k]
        info = self.info(k)
        mine = torch.tensor([1.0 if able else 0.0, float(nbytes), float(d.atom_capacity), float(info['cell_capacity']),
                             float(d.box[0]), float(d.box[1]), float(d.box[2])] + sc, dtype=torch.float64, device=dev)
        theirs = torch.zeros_like(mine)
        comm.exchange([(mine, peer)], [(theirs, peer)])
        mine_h, theirs_h = mine.cpu().numpy(), theirs.cpu().numpy()
        if not (mine_h[0] == 1.0 and theirs_h[0] == 1.0 and np.array_equal(mine_h[1:7], theirs_h[1:7])):
            comm.transports -= 1
            return False
        block = torch.as_tensor(self._DeviceBytes(ptr, nbytes), device=dev)
        if self._recv_block is None or self._recv_block.numel() != nbytes:
            self._recv_block = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        comm.exchange([(block, peer)], [(self._recv_block, peer)])
        comm.transports -= 1                                   # one state transport, two grouped calls
        block.copy_(self._recv_block)
        torch.cuda.current_stream(dev).synchronize()
        got = (C.c_double * 5)(*theirs_h[7:12].tolist())
        self.engine._check(self._lib.mcpyb_gcmc_set_config_scalars(self._h, int(k), got))
        # the ensemble: what ShardedReplicaExchange._swap_states sets from the peer's payload
        n = int(theirs_h[8])
        atoms = g.atoms
        atoms.arrays = {'numbers': np.full(n, self._Z[k], dtype=atoms.arrays['numbers'].dtype),
                        'positions': np.full((n, 3), np.nan)}
        g.E_old, g.n_atoms = float(theirs_h[7]), n
        self._host_fresh[k] = False
        return True

    def _same_storage(self, i, j):
        sig = self._storage_sig
        if sig is None:                                                # per chain: what must agree for a pointer swap
            sig = self._storage_sig = [(d.atom_capacity, tuple(d.box), d.lj_rc, d.n_move_types, tuple(sorted(self.info(k).items())))
                                       for k, d in enumerate(self._desc)]
        return sig[i] == sig[j]

    # ------------------------------------------------------------------ running
    def advance(self, n_trials, sync='full', counters=True):
        """Advance every chain by ``n_trials`` trials (an int, or one count per chain) in as few launches as the
        capacities allow.  ``sync='full'`` writes the whole state back into the ensembles; ``sync='scalars'`` only what
        the drivers read between blocks (:meth:`_apply_scalars`) and leaves the rest on the device for the next block.
        Returns the device time of the launches in milliseconds."""
        n = len(self.ensembles)
        want = np.full(n, int(n_trials), dtype=np.int64) if np.isscalar(n_trials) else np.asarray(n_trials, dtype=np.int64).copy()
        if want.shape != (n,) or np.any(want < 0):
            raise ValueError('n_trials: one non-negative count per chain')
        status = np.zeros(n, dtype=np.int32)
        total_ms = 0.0
        ms = C.c_double(0.0)
        self.traces = [np.empty(0, dtype=GCMC_TRIAL_DTYPE) for _ in range(n)] if self.trace_capacity else None
        for _ in range(64):
            stale = [k for k in range(n) if not self._device_fresh[k]]
            if stale:
                self.push(stale)
            self.engine._check(self._lib.mcpyb_gcmc_run(self._h, want.ctypes.data, status.ctypes.data))
            self._lib.mcpyb_gcmc_last_run_ms(self._h, C.byref(ms))
            total_ms += ms.value
            if counters:
                done = np.empty(n, dtype=np.int64)
                for k in range(n):
                    s = self._apply_scalars(k, True)
                    done[k] = int(s.trials_done) - self._trials_seen[k]
                    self._trials_seen[k] = int(s.trials_done)
            else:
                done = self._apply_block_scalars()
            if self.trace_capacity:
                for k in range(n):
                    self.traces[k] = np.concatenate([self.traces[k], self._get_trace(k)])
            want -= done
            self.trials_run += int(done.sum())
            self.last_status = status.tolist()
            if not np.any(want > 0):
                break
            for k in np.nonzero(want > 0)[0]:
                st = int(status[k])
                if st == 4:
                    self.sync_host()
                    raise ValueError('No atoms to displace.')                  # displacement_move.py:60
            need = {int(status[k]) for k in np.nonzero(want > 0)[0]} & {1, 2}
            if need:
                self._grow(need)
        else:
            raise RuntimeError(f'ResidentGCMC: chains keep stopping early ({[STATUS.get(s, s) for s in self.last_status]})')
        if sync == 'full':
            self.sync_host()
        return total_ms

    def run_trials(self, n_trials):
        """:meth:`advance` with the whole state written back."""
        return self.advance(n_trials, sync='full')

    def run(self, steps):
        """``gcmc.run(steps)`` of every ensemble (``base_ensemble.py:257-264``) with the trials on the device: blocks of
        steps between two outfile / trajectory writes are one launch each."""
        for g in self.ensembles:
            g.initialize_run()
        try:
            done = 0
            while done < steps:
                block = steps - done
                need_positions = False
                for g in self.ensembles:
                    for iv in (g._outfile_write_interval if g._outfile is not None else None,
                               g._trajectory_write_interval if g._traj_file is not None else None):
                        if iv:
                            block = min(block, iv - (g._step % iv))
                for g in self.ensembles:
                    if g._traj_file is not None and (g._step + block) % g._trajectory_write_interval == 0:
                        need_positions = True
                last = done + block >= steps
                self.advance([block * g.move_selector.n_moves for g in self.ensembles], sync='full' if last else 'scalars')
                if need_positions and not last:
                    self.sync_positions([k for k, g in enumerate(self.ensembles) if g._traj_file is not None and
                                         (g._step + block) % g._trajectory_write_interval == 0])
                done += block
                for g in self.ensembles:
                    g._step += block
                    if g._outfile is not None and g._step % g._outfile_write_interval == 0:
                        g.write_outfile()
                    if g._traj_file is not None and g._step % g._trajectory_write_interval == 0:
                        g.write_coordinates(g.atoms, g.E_old)
        finally:
            self.sync_host()
            for g in self.ensembles:
                g.finalize_run()

    # ------------------------------------------------------------------ helpers
    def _get_trace(self, k):
        buf = np.empty(self.trace_capacity, dtype=GCMC_TRIAL_DTYPE)
        n = C.c_int64(0)
        self.engine._check(self._lib.mcpyb_gcmc_get_trace(self._h, k, buf.ctypes.data, self.trace_capacity, C.byref(n)))
        return buf[:n.value].copy()

    def _grow(self, statuses):
        """Chains ran into a capacity (1: a cell of the cell list, 2: the atom storage): rebuild the chain set with
        more room -- for every chain alike, so that replica swaps keep finding identical storage."""
        self.sync_host()
        ccap = max(self.info(k)['cell_capacity'] for k in range(len(self.ensembles)))
        for d in self._desc:
            if 2 in statuses:
                if d.atom_capacity >= 65535:
                    raise RuntimeError('ResidentGCMC: a chain outgrew 65 535 atoms')
                d.atom_capacity = min(65535, 2 * d.atom_capacity)
            d.cell_capacity = 2 * ccap if 1 in statuses else ccap
        self._lib.mcpyb_gcmc_destroy(self._h)
        self._h = None
        h = C.c_void_p()
        self.engine._check(self._lib.mcpyb_gcmc_create(self.engine._h, len(self.ensembles), self._desc,
                                                       self.trace_capacity, C.byref(h)))
        self._h = h
        self._device_fresh = [False] * len(self.ensembles)
        self._trials_seen = [0] * len(self.ensembles)
        self._storage_sig = None

    def info(self, k=0):
        out = (C.c_int32 * 6)()
        self.engine._check(self._lib.mcpyb_gcmc_info(self._h, k, out))
        return {'cells': (out[0], out[1], out[2]), 'cell_capacity': out[3], 'state_bytes': out[4],
                'in_shared_memory': bool(out[5])}

    def cycles(self, k=0):
        """SM cycles of chain ``k``'s last launch: serial section of thread 0, whole loop, loop trips."""
        out = (C.c_int64 * 8)()
        self.engine._check(self._lib.mcpyb_gcmc_cycles(self._h, k, out))
        return {'serial': out[0], 'total': out[1], 'rounds': out[2], 'energy_resyncs': out[3], 'look_ahead': out[4],
                'row_sums': out[5], 'look_ahead_to_serial': out[6], 'serial_to_look_ahead': out[7]}

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mcpyb_gcmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def install():
    """Route ``GrandCanonicalEnsemble.run(steps)`` through the device-resident chains whenever it can be: afterwards an
    UNMODIFIED mcpy script whose ensemble has a ``B200Calculator('lj')`` and is in scope (:func:`supports`) runs its
    trials on the device -- same draws, decisions, files -- and every other ensemble runs the reference's loop as before.
    Returns the original method (pass it to :func:`uninstall`).  This is the three-line change a maintainer would make in
    ``BaseEnsemble.run`` (``mcpy/ensembles/base_ensemble.py:257-264``), done from outside."""
    from mcpy.ensembles.grand_canonical_ensemble import GrandCanonicalEnsemble
    original = GrandCanonicalEnsemble.run
    if getattr(original, '_b200_resident', False):
        return original._b200_original

    def run(self, steps):
        calc = getattr(self, '_calculator', None)
        if getattr(calc, 'lj_parameters', None) is not None and getattr(calc, 'engine', None) is not None \
                and type(self) is GrandCanonicalEnsemble and supports(self) is None:
            chains = ResidentGCMC(self, engine=calc.engine)
            try:
                return chains.run(steps)
            finally:
                chains.close()
        return original(self, steps)
    run._b200_resident = True
    run._b200_original = original
    GrandCanonicalEnsemble.run = run
    return original


def uninstall(original=None):
    """Undo :func:`install`."""
    from mcpy.ensembles.grand_canonical_ensemble import GrandCanonicalEnsemble
    current = GrandCanonicalEnsemble.__dict__.get('run')
    if current is not None and getattr(current, '_b200_resident', False):
        if 'run' in GrandCanonicalEnsemble.__dict__:
            del GrandCanonicalEnsemble.run               # BaseEnsemble.run is inherited again
        if original is not None and original is not getattr(GrandCanonicalEnsemble, 'run'):
            GrandCanonicalEnsemble.run = original
