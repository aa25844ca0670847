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
    ``resync_energy``: after a block, replace the accumulated ``E_old`` by a full energy of the final configuration
    through the ensemble's own calculator (one call per chain and block).
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
        self._resident = False
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
    def push(self):
        """Upload every ensemble's current state (configuration, generators, energy, best) to its chain."""
        from random import Random  # noqa: F401  (the states handled here are random.Random.getstate() tuples)
        for k, g in enumerate(self.ensembles):
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
            gens = [ms.rng.random] + [mv.rng.random for mv in ms.move_list] + [g.rng_acceptance.random]
            mt = np.empty((len(gens), GCMC_MT_WORDS), dtype=np.uint32)
            self._gauss = getattr(self, '_gauss', {})
            for j, r in enumerate(gens):
                version, words, gauss_next = r.getstate()
                if version != 3 or len(words) != GCMC_MT_WORDS:
                    raise ValueError('unexpected random.Random state layout')
                mt[j] = words
                self._gauss[(k, j)] = gauss_next
            pos = np.ascontiguousarray(atoms.positions, dtype=np.float64)
            self.engine._check(self._lib.mcpyb_gcmc_set_state(self._h, k, pos.ctypes.data if len(pos) else None,
                                                              C.byref(s), mt.ctypes.data))
        self._resident = True

    # ------------------------------------------------------------------ device -> host objects
    def pull(self):
        """Write every chain's state back into its ensemble: atoms, ``E_old``, ``n_atoms``, the generator states,
        the selector's counters and ``to_use``, the running minimum."""
        from ase import Atoms
        from ase.data import atomic_numbers
        for k, g in enumerate(self.ensembles):
            d = self._desc[k]
            cap = d.atom_capacity
            nm = d.n_move_types
            pos = np.empty((cap, 3), dtype=np.float64)
            best = np.empty((cap, 3), dtype=np.float64)
            mt = np.empty((nm + 2, GCMC_MT_WORDS), dtype=np.uint32)
            s = GcmcScalars()
            self.engine._check(self._lib.mcpyb_gcmc_get_state(self._h, k, pos.ctypes.data, C.byref(s), mt.ctypes.data,
                                                              best.ctypes.data))
            n = int(s.n_atoms)
            Z = atomic_numbers[self._species[k]]
            atoms = g.atoms
            # same object, new arrays: what the reference's in-place moves leave behind
            atoms.arrays = {'numbers': np.full(n, Z, dtype=atoms.arrays['numbers'].dtype),
                            'positions': pos[:n].copy()}
            g.n_atoms = n
            g.E_old = float(s.energy)
            ms = g.move_selector
            ms.to_use = int(s.last_move)
            for m in range(nm):
                a, acc, f = int(s.attempts[m]), int(s.accepted[m]), int(s.failed[m])
                ms.move_counter[m] += a
                ms.move_counter_total[m] += a
                ms.move_acceptance[m] += acc
                ms.move_acceptance_total[m] += acc
                ms.move_failed_counter[m] += f
                ms.move_failed_counter_total[m] += f
            gens = [ms.rng.random] + [mv.rng.random for mv in ms.move_list] + [g.rng_acceptance.random]
            for j, r in enumerate(gens):
                r.setstate((3, tuple(int(w) for w in mt[j]), self._gauss[(k, j)]))
            for c, cell in enumerate(self._pcg_cells[k]):
                st = cell._rng.bit_generator.state
                st['state']['state'] = (int(s.pcg[c][0]) << 64) | int(s.pcg[c][1])
                cell._rng.bit_generator.state = st
            if s.best_n >= 0:                                          # improved during the block
                g._best_score = float(s.best_score)
                g._best_energy = float(s.best_energy)
                bn = int(s.best_n)
                g._best_atoms = Atoms(numbers=np.full(bn, Z), positions=best[:bn].copy(), cell=atoms.cell, pbc=atoms.pbc)
            if self.resync_energy:
                g.E_old = float(g.compute_energy(atoms))
        # the counters were added to the host ones: the next push starts the device's from zero
        self._resident = False

    # ------------------------------------------------------------------ running
    def run_trials(self, n_trials):
        """Advance every chain by ``n_trials`` trials (an int, or one count per chain) in as few launches as the
        capacities allow, and write the states back.  Returns the device time of the launches in milliseconds."""
        n = len(self.ensembles)
        want = np.full(n, int(n_trials), dtype=np.int64) if np.isscalar(n_trials) else np.asarray(n_trials, dtype=np.int64).copy()
        if want.shape != (n,) or np.any(want < 0):
            raise ValueError('n_trials: one non-negative count per chain')
        status = np.zeros(n, dtype=np.int32)
        total_ms = 0.0
        self.traces = [np.empty(0, dtype=GCMC_TRIAL_DTYPE) for _ in range(n)] if self.trace_capacity else None
        for _ in range(64):
            self.push()
            self.engine._check(self._lib.mcpyb_gcmc_run(self._h, want.ctypes.data, status.ctypes.data))
            ms = C.c_double(0.0)
            self._lib.mcpyb_gcmc_last_run_ms(self._h, C.byref(ms))
            total_ms += ms.value
            done = self._trials_done()
            if self.trace_capacity:
                for k in range(n):
                    self.traces[k] = np.concatenate([self.traces[k], self._get_trace(k)])
            self.pull()
            want -= done
            self.trials_run += int(done.sum())
            self.last_status = status.tolist()
            if not np.any(want > 0):
                break
            for k in np.nonzero(want > 0)[0]:
                st = int(status[k])
                if st == 4:
                    raise ValueError('No atoms to displace.')                  # displacement_move.py:60
                if st in (1, 2):
                    self._grow(int(k), st)
        else:
            raise RuntimeError(f'ResidentGCMC: chains keep stopping early ({[STATUS.get(s, s) for s in self.last_status]})')
        return total_ms

    def run(self, steps):
        """``gcmc.run(steps)`` of every ensemble (``base_ensemble.py:257-264``) with the trials on the device: blocks of
        steps between two outfile / trajectory writes are one launch each."""
        for g in self.ensembles:
            g.initialize_run()
        try:
            done = 0
            while done < steps:
                block = steps - done
                for g in self.ensembles:
                    for iv in (g._outfile_write_interval if g._outfile is not None else None,
                               g._trajectory_write_interval if g._traj_file is not None else None):
                        if iv:
                            block = min(block, iv - (g._step % iv))
                self.run_trials([block * g.move_selector.n_moves for g in self.ensembles])
                done += block
                for g in self.ensembles:
                    g._step += block
                    if g._outfile is not None and g._step % g._outfile_write_interval == 0:
                        g.write_outfile()
                    if g._traj_file is not None and g._step % g._trajectory_write_interval == 0:
                        g.write_coordinates(g.atoms, g.E_old)
        finally:
            for g in self.ensembles:
                g.finalize_run()

    # ------------------------------------------------------------------ helpers
    def _trials_done(self):
        out = np.zeros(len(self.ensembles), dtype=np.int64)
        s = GcmcScalars()
        for k in range(len(self.ensembles)):
            self.engine._check(self._lib.mcpyb_gcmc_get_state(self._h, k, None, C.byref(s), None, None))
            out[k] = s.trials_done
        return out

    def _get_trace(self, k):
        buf = np.empty(self.trace_capacity, dtype=GCMC_TRIAL_DTYPE)
        n = C.c_int64(0)
        self.engine._check(self._lib.mcpyb_gcmc_get_trace(self._h, k, buf.ctypes.data, self.trace_capacity, C.byref(n)))
        return buf[:n.value].copy()

    def _grow(self, k, status):
        """A chain ran into a capacity: rebuild the chain set with more room for it."""
        d = self._desc[k]
        info = self.info(k)
        if status == 2:
            if d.atom_capacity >= 65535:
                raise RuntimeError('ResidentGCMC: a chain outgrew 65 535 atoms')
            d.atom_capacity = min(65535, 2 * d.atom_capacity)
        else:
            d.cell_capacity = 2 * info['cell_capacity']
        self._lib.mcpyb_gcmc_destroy(self._h)
        self._h = None
        h = C.c_void_p()
        self.engine._check(self._lib.mcpyb_gcmc_create(self.engine._h, len(self.ensembles), self._desc,
                                                       self.trace_capacity, C.byref(h)))
        self._h = h

    def info(self, k=0):
        out = (C.c_int32 * 6)()
        self.engine._check(self._lib.mcpyb_gcmc_info(self._h, k, out))
        return {'cells': (out[0], out[1], out[2]), 'cell_capacity': out[3], 'state_bytes': out[4],
                'in_shared_memory': bool(out[5])}

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mcpyb_gcmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
