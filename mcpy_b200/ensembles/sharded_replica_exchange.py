"""Replica exchange sharded over GPUs: one process per GPU, replicas block-partitioned.

``ShardedReplicaExchange`` IS the reference's ``BatchedReplicaExchange``
(``mcpy/ensembles/batched_replica_exchange.py``) run on the local shard of the ladder: it
subclasses it, builds only the replicas of ``shard_range(R, rank, world)`` and inherits ``run``,
``_rebatch_initial_energies``, ``_batched_gcmc_step``, ``_batched_single_move`` (ONE batched
``calculator.get_potential_energies`` call per sub-move, ``:235-302``) and the per-replica /
per-run output unchanged.  There is no data-path collective.  Only the exchange step is
overridden:

* every ``exchange_interval`` steps the ranks ``all_gather`` one fixed-size fp64 record per
  replica ``[E_old, beta, N_s.., mu_s.., Lambda_s.., has_mu_s..]`` (a few hundred bytes over
  NCCL / NVLink; latency-bound) -- the only collective;
* every rank then evaluates the SAME pairing and Metropolis test with the reference's
  master-seeded generator (``self.rng``, ``:110,311,370``: identical on every rank) and the
  reference's exponent (``_accept_swap`` ``:320-370``), so all ranks agree without further
  communication;
* an accepted pair that straddles two ranks ships its state point-to-point -- the ``Atoms``
  object as a whole (positions, numbers, per-atom arrays, cell, pbc, constraints, info), ``E_old``
  and ``n_atoms``, what ``_swap_states`` (``:401-410``) exchanges and what the MPI driver's
  ``sendrecv`` carries (``replica_exchange.py:188-192``).  Contiguous sharding makes only the
  G-1 boundary pairs cross-rank.  Temperatures / mu stay with the slot and both slots refresh
  their free volume after a swap.

Per-rank files: ``outfile`` and ``global_minimum_file`` get a ``.rank<g>`` suffix when
``world_size > 1`` (each rank logs its own replicas; indices in those files are local).
"""
from __future__ import annotations

import logging
import pickle
from typing import Callable, Dict, List, Optional

import numpy as np

try:
    from mcpy.ensembles.batched_replica_exchange import BatchedReplicaExchange
except ImportError as exc:                                   # pragma: no cover
    raise ImportError('mcpy_b200.ensembles.ShardedReplicaExchange extends BatchedReplicaExchange of '
                      'farrisric/mcpy: the `mcpy` package (and its dependency `ase`) must be importable') from exc

from .acceptance import swap_exponent
from .ladder_comm import LadderComm

logger = logging.getLogger(__name__)


class ShardedReplicaExchange(BatchedReplicaExchange):
    def __init__(self, gcmc_factory: Callable, calculator, temperatures: Optional[List[float]] = None,
                 mus: Optional[List[Dict[str, float]]] = None, gcmc_steps: int = 100,
                 exchange_interval: int = 10, outfile: Optional[str] = None, write_out_interval: int = 20,
                 seed: int = 31, global_minimum_file: Optional[str] = None, consolidate_logging: bool = True,
                 rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                 device: str = 'cpu') -> None:
        if temperatures is None and mus is None:
            raise ValueError('Provide either temperatures or mus (one per replica).')
        if temperatures is not None and mus is not None:
            raise ValueError('Pass temperatures OR mus, not both.')
        ladder = list(temperatures) if temperatures is not None else list(mus)
        self.comm = LadderComm(len(ladder), rank=rank, world_size=world_size, group=group, device=device)
        self.rank, self.world_size, self.group, self.device = self.comm.rank, self.comm.world_size, group, device
        self.n_replicas_total = len(ladder)
        self.lo, self.hi = self.comm.lo, self.comm.hi
        lo = self.lo

        def local_factory(rank=0, **kw):          # the reference numbers replicas 0..R-1 through `rank`
            return gcmc_factory(rank=lo + rank, **kw)

        def per_rank(name):
            return name if name is None or self.world_size == 1 else f'{name}.rank{self.rank}'

        key = 'temperatures' if temperatures is not None else 'mus'
        super().__init__(local_factory, calculator, **{key: ladder[self.lo:self.hi]}, gcmc_steps=gcmc_steps,
                         exchange_interval=exchange_interval, outfile=per_rank(outfile),
                         write_out_interval=write_out_interval, seed=seed,
                         global_minimum_file=per_rank(global_minimum_file),
                         consolidate_logging=consolidate_logging)
        self.global_attempts = [0] * self.n_replicas_total
        self.global_successes = [0] * self.n_replicas_total
        self.swap_log = []                             # (step, i, j, delta, accepted), identical on every rank
        self.species = None
        self.exchange_rounds = 0

    @property
    def collective_seconds(self):                      # all_gather of the records (+ the host<->device hops around it)
        return self.comm.collective_seconds

    @property
    def transport_seconds(self):                       # point-to-point state transport of cross-rank swaps
        return self.comm.transport_seconds

    # ------------------------------------------------------------------ helpers
    def owner(self, i: int) -> int:
        return self.comm.owner(i)

    def local(self, i: int):
        return self.replicas[i - self.lo]

    def _species(self) -> List[str]:
        """Union of the replicas' ``_mu`` keys in first-seen order (the reference sums ``mu_s N_s`` in dict
        order, ``grand_canonical_ensemble.py:188-195``)."""
        keys: List[str] = []
        for r in self.replicas:
            for k in (getattr(r, '_mu', None) or {}):
                if k not in keys:
                    keys.append(k)
        merged: List[str] = []
        for ks in self.comm.all_gather_object(keys):
            for k in ks:
                if k not in merged:
                    merged.append(k)
        return merged

    def _batched_gcmc_step(self) -> None:
        if self.replicas:                              # a rank may own no replica (R < world size)
            super()._batched_gcmc_step()

    # --------------------------------------------------------- exchange
    def _record(self, r) -> np.ndarray:
        S = len(self.species)
        rec = np.zeros(2 + 4 * S)
        rec[0], rec[1] = r.E_old, r.units.beta
        mu = getattr(r, '_mu', None) or {}
        for k, s in enumerate(self.species):
            # counted for EVERY species of the ladder: the cross terms score this configuration with the
            # neighbour's chemical potentials (`ri._minimum_score(rj.atoms, ...)`, :347-348)
            rec[2 + k] = self._count(r, s)
            rec[2 + 2 * S + k] = r.units.lambda_dbs.get(s, np.nan)
            if s in mu:
                rec[2 + S + k] = mu[s]
                rec[2 + 3 * S + k] = 1.0
        return rec

    @staticmethod
    def _count(r, s) -> int:
        """``r._species_count(r.atoms, s)`` (``grand_canonical_ensemble.py:176-186``); an atomic species is counted on
        the atomic numbers instead of a list of 2 000 chemical symbols."""
        if s in r.units.molecules:
            return r._species_count(r.atoms, s)
        try:
            from ase.data import atomic_numbers
            return int(np.count_nonzero(r.atoms.numbers == atomic_numbers[s]))
        except (ImportError, KeyError):
            return r._species_count(r.atoms, s)

    def _gather_records(self) -> np.ndarray:
        """The only collective: all_gather of the per-replica records (``LadderComm.gather_records``)."""
        width = 2 + 4 * len(self.species)
        local = np.array([self._record(r) for r in self.replicas]).reshape(-1, width)
        return self.comm.gather_records(local)

    def _attempt_exchanges(self) -> None:
        if self.species is None:
            self.species = self._species()
        step = self._re_step
        offset = 0 if self.rng.get_uniform() < 0.5 else 1
        records = self._gather_records()
        self.exchange_rounds += 1
        S = len(self.species)
        for i in range(offset, self.n_replicas_total - 1, 2):
            j = i + 1
            self.global_attempts[i] += 1
            self.global_attempts[j] += 1
            delta = swap_exponent(records[i], records[j], S)
            if not np.isfinite(delta):
                self.logger.warning('swap %d<->%d: non-finite acceptance exponent (delta=%r); rejecting. A NaN/inf '
                                    'replica energy usually means a diverged relaxation.', i, j, delta)
                accepted = False
            else:
                p = 1.0 if delta >= 0.0 else float(np.exp(delta))
                accepted = self.rng.get_uniform() < p
            self.swap_log.append((step, i, j, float(delta), bool(accepted)))
            if accepted:
                self._swap_states(i, j)
                self.global_successes[i] += 1
                self.global_successes[j] += 1
        # the inherited output / post-mortem code reads the local slice of the tallies
        self.exchange_attempts = self.global_attempts[self.lo:self.hi]
        self.exchange_successes = self.global_successes[self.lo:self.hi]

    # ------------------------------------------------------- state transport
    @staticmethod
    def _pack(r) -> bytes:
        atoms = r.atoms
        calc = getattr(atoms, 'calc', None)
        atoms.calc = None                              # a calculator holds device handles: it stays with the slot
        try:
            return pickle.dumps((atoms, float(r.E_old), int(r.n_atoms)), protocol=pickle.HIGHEST_PROTOCOL)
        finally:
            atoms.calc = calc

    def _swap_states(self, i: int, j: int) -> None:
        oi, oj = self.owner(i), self.owner(j)
        if oi == self.rank and oj == self.rank:
            ri, rj = self.local(i), self.local(j)
            ri.atoms, rj.atoms = rj.atoms, ri.atoms
            ri.E_old, rj.E_old = rj.E_old, ri.E_old
            ri.n_atoms, rj.n_atoms = rj.n_atoms, ri.n_atoms
            ri.calculate_cells_volume(ri.atoms)
            rj.calculate_cells_volume(rj.atoms)
            return
        if self.rank not in (oi, oj):
            return
        mine, peer = (i, oj) if self.rank == oi else (j, oi)
        r = self.local(mine)
        calc = getattr(r.atoms, 'calc', None)
        atoms, e_old, n_atoms = pickle.loads(self.comm.exchange_bytes(self._pack(r), peer))
        atoms.calc = calc
        r.atoms, r.E_old, r.n_atoms = atoms, e_old, n_atoms
        r.calculate_cells_volume(r.atoms)
