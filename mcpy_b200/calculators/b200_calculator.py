"""``B200Calculator`` -- LJ / EMT energies on a B200 behind mcpy's calculator API.

The drop-in boundary:

* ``get_potential_energy(atoms) -> float`` -- the sole call the ensembles make
  (``mcpy/ensembles/base_ensemble.py:190-191``, used at
  ``grand_canonical_ensemble.py:56,283``);
* ``get_potential_energies(atoms_list, chunk_size=None) -> ndarray`` -- the batched
  contract of ``BatchedReplicaExchange`` (checked at
  ``batched_replica_exchange.py:98-102``, called at ``:202,278``; reference
  implementation ``alchemi_calculator.py:118-155``: empty list -> ``np.empty(0)``,
  ``chunk_size`` caps the structures per pass);
* the ASE ``Calculator`` protocol (``calculate`` / ``results`` /
  ``implemented_properties``) so ``atoms.calc = calc; atoms.get_potential_energy()``
  works (``mcpy/calculators/base_calculator.py:19-22``,
  ``canonical_ensemble.py:117-121``), forces included, so ASE's optimisers can drive it.
  :class:`B200RelaxCalculator` is the relax-then-energy adapter of
  ``base_calculator.py:6-22`` with the LBFGS loop inside the engine.

Beyond the reference interface, :meth:`trial_delta_energies` exposes the
batched trial-move kernel (many proposed insertions / deletions / displacements
scored against one resident configuration in one launch).

``atoms`` is never mutated.  Non-finite energies are returned, not raised
(``batched_replica_exchange.py:354-362`` handles them).  Nothing here computes
on the CPU: without the CUDA library or a GPU, construction raises.
"""
from __future__ import annotations

import numpy as np

from ..engine import Engine
from ..utils.chunking import chunk_ranges

try:                                     # real ASE (or the test shim) when importable
    from ase.calculators.calculator import Calculator as _AseCalculator
    from ase.calculators.calculator import all_changes as _all_changes
except Exception:                        # pragma: no cover - plain duck-typed use
    _AseCalculator = None
    _all_changes = ['positions', 'numbers', 'cell', 'pbc', 'initial_charges', 'initial_magmoms']


def _config_of(atoms):
    """(positions, numbers, cell 3x3, pbc) views of an ASE-like ``Atoms``."""
    pos = np.asarray(atoms.positions, dtype=np.float64)
    num = np.asarray(atoms.numbers, dtype=np.int32)
    cell = np.asarray(atoms.cell, dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(atoms.pbc, dtype=bool).reshape(3)
    return pos, num, cell, pbc


class _B200Core:
    """Plug-in methods shared by the ASE-derived and the plain class."""

    implemented_properties = ['energy', 'free_energy', 'forces']

    def _init_engine(self, potential, epsilon, sigma, rc, device, chunk_size, incremental=True, precision='fp64'):
        potential = str(potential).lower()
        if potential not in ('lj', 'lennardjones', 'emt'):
            raise ValueError(f"potential must be 'lj' or 'emt', got {potential!r}")
        if chunk_size is not None and chunk_size <= 0:
            raise ValueError(f'chunk_size must be a positive int or None, got {chunk_size}')
        self.potential = 'emt' if potential == 'emt' else 'lj'
        self.chunk_size = chunk_size
        self.engine = Engine(device)
        if self.potential == 'lj':
            self.engine.set_lj(epsilon=epsilon, sigma=sigma, rc=rc)
            self.lj_parameters = {'epsilon': float(epsilon), 'sigma': float(sigma),
                                  'rc': self.engine.cutoff}
        else:
            self.engine.set_emt()
        if precision != 'fp64':
            if self.potential != 'lj':
                raise ValueError("precision='fp32' is available for the LJ trial batches only")
            self.engine.set_precision(precision)
        self.precision = precision
        self.n_energy_calls = 0
        self.incremental = bool(incremental)
        self.path_counts = [0, 0, 0]          # full evaluations, incremental, identical-to-base

    # -- mcpy calculator interface -------------------------------------------
    def get_potential_energy(self, atoms=None, force_consistent=False) -> float:
        """Total potential energy of ``atoms`` (not mutated)."""
        if atoms is None:
            atoms = getattr(self, 'atoms', None)
            if atoms is None:
                raise ValueError('get_potential_energy() needs an Atoms object')
        self.n_energy_calls += 1
        if self.incremental:
            # one resident base + a delta per call; same numbers as the full path to rounding
            res = self.engine.tracked_energy_atoms(atoms)
            if res is not None:
                self.path_counts[res[1]] += 1
                return res[0]
        cell = np.asarray(atoms.cell, dtype=np.float64)
        if self.incremental:
            e, path = self.engine.tracked_energy(atoms.positions, atoms.numbers, cell, atoms.pbc)
            self.path_counts[path] += 1
            return e
        return self.engine.potential_energy(atoms.positions, atoms.numbers, cell, atoms.pbc)

    def get_forces(self, atoms=None) -> np.ndarray:
        """Forces ``(n, 3)`` of ``atoms`` (ASE ``LennardJones`` / ``EMT`` ``results['forces']``): what the
        optimisers behind ``BaseCalculator.get_potential_energy`` (``base_calculator.py:14-22``) and
        ``CanonicalEnsemble.relax`` (``canonical_ensemble.py:115-123``) ask for."""
        return self.get_energy_and_forces(atoms)[1]

    def get_energy_and_forces(self, atoms=None):
        if atoms is None:
            atoms = getattr(self, 'atoms', None)
            if atoms is None:
                raise ValueError('get_forces() needs an Atoms object')
        self.n_energy_calls += 1
        cell = np.asarray(atoms.cell, dtype=np.float64)
        return self.engine.energy_forces(atoms.positions, atoms.numbers, cell, atoms.pbc)

    def get_potential_energies(self, atoms_list, chunk_size=None) -> np.ndarray:
        """Energies of several (ragged) structures, one batched launch per chunk."""
        if not atoms_list:
            return np.empty(0, dtype=np.float64)
        cs = self.chunk_size if chunk_size is None else chunk_size
        out = []
        if self.incremental and (cs is None or cs >= len(atoms_list)):
            # the whole batch in one piece: resident bases + one batched delta launch, the list walked in C
            res = self.engine.tracked_energies_atoms(atoms_list)
            if res is not None:
                en, paths = res
                c = np.bincount(paths, minlength=3)
                pc = self.path_counts
                pc[0] += int(c[0]); pc[1] += int(c[1]); pc[2] += int(c[2])
                self.n_energy_calls += len(atoms_list)
                return en
        for start, stop in chunk_ranges(len(atoms_list), cs):
            chunk = atoms_list[start:stop]
            # marshalling dominates this call (64 x 2000 atoms: ~0.75 ms of a 1.1 ms call when every
            # Atoms is converted on its own), so convert once per batch, not once per replica
            pos_list = [a.positions for a in chunk]
            num_list = [a.numbers for a in chunk]
            cells = np.empty((len(chunk), 3, 3))
            pbcs = np.empty((len(chunk), 3), dtype=bool)
            for r, a in enumerate(chunk):
                cells[r] = a.cell
                pbcs[r] = a.pbc
            cells = cells.reshape(len(chunk), 9)
            if self.incremental and start == 0 and stop == len(atoms_list):
                # the whole batch in one piece: resident bases + one batched delta launch
                en, paths = self.engine.tracked_energies(pos_list, num_list, cells, pbcs)
                for pth in paths:
                    self.path_counts[int(pth)] += 1
                out.append(en)
            else:
                out.append(self.engine.potential_energies(pos_list, num_list, cells, pbcs))
            self.n_energy_calls += stop - start
        return np.concatenate(out)

    # -- batched trial moves (beyond the reference interface) ------------------
    def set_configuration(self, atoms):
        """Make ``atoms`` the resident configuration trial moves are scored against."""
        pos, num, cell, pbc = _config_of(atoms)
        self.engine.upload([pos], [num], cell.reshape(1, 9), pbc.reshape(1, 3))

    def set_configurations(self, atoms_list):
        cfg = [_config_of(a) for a in atoms_list]
        self.engine.upload([c[0] for c in cfg], [c[1] for c in cfg],
                           np.stack([c[2].reshape(9) for c in cfg]), np.stack([c[3] for c in cfg]))

    def trial_delta_energies(self, trials, rep=None, return_pairs=False):
        """dE of each trial against the resident configuration(s).

        ``trials`` is a sequence of ``(old_indices, new_positions, new_numbers)``:
        displacement -> both, insertion -> ``([], pos, Z)``, deletion -> ``(idx, [], [])``.
        """
        old_off, new_off = [0], [0]
        old_idx, new_pos, new_Z = [], [], []
        for old, pos, Z in trials:
            old = np.atleast_1d(np.asarray(old, dtype=np.int32)) if np.size(old) else np.zeros(0, np.int32)
            pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
            Z = np.atleast_1d(np.asarray(Z, dtype=np.int32)) if np.size(Z) else np.zeros(0, np.int32)
            if len(Z) != len(pos):
                raise ValueError('one atomic number per new position is required')
            old_idx.append(old); new_pos.append(pos); new_Z.append(Z)
            old_off.append(old_off[-1] + len(old)); new_off.append(new_off[-1] + len(pos))
        cat = lambda xs, shape, dt: (np.concatenate(xs) if xs else np.zeros(shape, dt))
        return self.engine.trial_delta_e(old_off, cat(old_idx, 0, np.int32), new_off,
                                         cat(new_pos, (0, 3), np.float64), cat(new_Z, 0, np.int32),
                                         rep=rep, return_pairs=return_pairs)


if _AseCalculator is not None:

    class B200Calculator(_B200Core, _AseCalculator):
        __doc__ = __doc__
        implemented_properties = ['energy', 'free_energy', 'forces']
        nolabel = True

        def __init__(self, potential='lj', epsilon=1.0, sigma=1.0, rc=None, device=0,
                     chunk_size=None, incremental=True, precision='fp64', **kwargs):
            _AseCalculator.__init__(self, **kwargs)
            self._init_engine(potential, epsilon, sigma, rc, device, chunk_size, incremental, precision)

        def calculate(self, atoms=None, properties=('energy',), system_changes=_all_changes):
            _AseCalculator.calculate(self, atoms, properties, system_changes)
            if 'forces' in properties:
                e, f = _B200Core.get_energy_and_forces(self, self.atoms)
                self.results['forces'] = f
            else:
                e = _B200Core.get_potential_energy(self, self.atoms)
            self.results['energy'] = e
            self.results['free_energy'] = e

        def get_potential_energy(self, atoms=None, force_consistent=False):
            # Straight to the engine: mcpy calls this once per trial on a mutated
            # Atoms, so ASE's result cache (one atoms.copy() + compare per call)
            # would only add host overhead.
            return _B200Core.get_potential_energy(self, atoms)

else:

    class B200Calculator(_B200Core):
        __doc__ = __doc__

        def __init__(self, potential='lj', epsilon=1.0, sigma=1.0, rc=None, device=0,
                     chunk_size=None, incremental=True, precision='fp64', **kwargs):
            self.results = {}
            self.atoms = None
            self._init_engine(potential, epsilon, sigma, rc, device, chunk_size, incremental, precision)

        def calculate(self, atoms=None, properties=('energy',), system_changes=_all_changes):
            self.atoms = atoms
            if 'forces' in properties:
                e, f = self.get_energy_and_forces(atoms)
                self.results = {'energy': e, 'free_energy': e, 'forces': f}
            else:
                e = self.get_potential_energy(atoms)
                self.results = {'energy': e, 'free_energy': e}


class B200RelaxCalculator:
    """Relax-then-energy adapter: the drop-in for ``mcpy.calculators.BaseCalculator(calculator, steps, fmax)``
    (``mcpy/calculators/base_calculator.py:6-22``; ADR ``docs/adr/0002-trial-energy-is-lbfgs-relaxed.md``) with an
    ASE ``LennardJones`` / ``EMT`` calculator inside.

    ``get_potential_energy(atoms)`` relaxes ``atoms`` IN PLACE with ASE's LBFGS iteration (no line search, ASE
    defaults ``maxstep=0.2``, ``memory=100``, ``alpha=70``, ``damping=1``) for at most ``steps`` steps or until
    ``max |f| < fmax``, honouring ``FixAtoms`` constraints, and returns the energy of the relaxed positions --
    what ``LBFGS(atoms).run(steps=steps, fmax=fmax); atoms.get_potential_energy()`` does, as one
    ``mcpyb_relax_lbfgs_host`` call instead of one Python round trip per optimiser step.  ``steps=0`` is the
    single-point energy (``atoms`` untouched), as in the reference.
    """

    def __init__(self, potential='lj', steps=100, fmax=0.05, epsilon=1.0, sigma=1.0, rc=None, device=0,
                 maxstep=0.2, memory=100, alpha=70.0, damping=1.0, calculator=None):
        if steps < 0:
            raise ValueError(f'steps must be >= 0, got {steps}')
        if not fmax >= 0.0:
            raise ValueError(f'fmax must be >= 0, got {fmax}')
        # an existing B200Calculator (and its engine) may be shared
        self.calculator = calculator if calculator is not None else B200Calculator(
            potential, epsilon=epsilon, sigma=sigma, rc=rc, device=device, incremental=False)
        self.engine = self.calculator.engine
        self.steps, self.fmax = int(steps), float(fmax)
        self.maxstep, self.memory, self.alpha, self.damping = float(maxstep), int(memory), float(alpha), float(damping)
        self.last_steps = 0               # optimiser steps of the last call
        self.last_fmax = float('nan')     # max |f| at its final positions

    @staticmethod
    def _fixed_indices(atoms):
        idx = []
        for c in getattr(atoms, 'constraints', None) or []:
            if type(c).__name__ != 'FixAtoms':
                raise ValueError(f'B200RelaxCalculator supports FixAtoms constraints only, got {type(c).__name__}')
            idx.extend(int(i) for i in c.get_indices())
        return idx or None

    def get_potential_energy(self, atoms) -> float:
        cell = np.asarray(atoms.cell, dtype=np.float64)
        if self.steps == 0 or len(atoms) == 0:
            self.last_steps, self.last_fmax = 0, float('nan')
            return self.engine.potential_energy(atoms.positions, atoms.numbers, cell, atoms.pbc)
        e, pos, self.last_steps, self.last_fmax = self.engine.relax_lbfgs(
            atoms.positions, atoms.numbers, cell, atoms.pbc, fmax=self.fmax, steps=self.steps,
            fixed=self._fixed_indices(atoms), maxstep=self.maxstep, memory=self.memory, alpha=self.alpha,
            damping=self.damping)
        atoms.positions[...] = pos        # the optimiser writes the relaxed geometry back (base_calculator.py:19-21)
        return e


class B200LBFGS:
    """Drop-in for ``ase.optimize.LBFGS`` wherever mcpy takes an optimiser CLASS -- ``CanonicalEnsemble(optimizer=...)``
    builds ``optimizer(atoms, logfile=None)`` and calls ``run(fmax=...)`` on every trial configuration
    (``mcpy/ensembles/canonical_ensemble.py:115-123``).  ``atoms.calc`` must be a :class:`B200Calculator`; the
    whole relaxation is ONE ``mcpyb_relax_lbfgs_host`` call (ASE's iteration without line search and ASE's
    defaults, ``FixAtoms`` honoured), the relaxed positions are written back in place."""

    def __init__(self, atoms, logfile=None, trajectory=None, maxstep=0.2, memory=100, damping=1.0, alpha=70.0,
                 **kwargs):
        self.atoms = atoms
        self.maxstep, self.memory, self.damping, self.alpha = float(maxstep), int(memory), float(damping), float(alpha)
        self.nsteps = 0
        self.fmax = float('nan')

    def run(self, fmax=0.05, steps=100_000_000):
        atoms = self.atoms
        engine = getattr(getattr(atoms, 'calc', None), 'engine', None)
        if engine is None:
            raise ValueError('B200LBFGS needs atoms.calc to be a B200Calculator')
        if steps == 0 or len(atoms) == 0:
            return True
        e, pos, self.nsteps, self.fmax = engine.relax_lbfgs(
            atoms.positions, atoms.numbers, np.asarray(atoms.cell, dtype=np.float64), atoms.pbc, fmax=fmax,
            steps=steps, fixed=B200RelaxCalculator._fixed_indices(atoms), maxstep=self.maxstep, memory=self.memory,
            alpha=self.alpha, damping=self.damping)
        atoms.positions[...] = pos
        return bool(self.fmax < fmax)

    def get_number_of_steps(self):
        return self.nsteps
