"""Replica exchange sharded over GPUs: one process per GPU, replicas block-partitioned.

Multi-GPU counterpart of ``BatchedReplicaExchange``
(``mcpy/ensembles/batched_replica_exchange.py``), SURVEY.md section 8e:

* rank g owns the replicas ``shard_range(R, g, G)`` and runs their trial moves locally --
  propose on the host, ONE batched ``calculator.get_potential_energies`` call per
  sub-move (``_batched_single_move``, reference ``:235-302``), per-replica acceptance;
  there is no data-path collective;
* every ``exchange_interval`` steps the ranks ``all_gather`` one fixed-size fp64 record
  per replica ``[E_old, beta, N_s.., mu_s.., Lambda_s..]`` (a few hundred bytes over
  NCCL / NVLink; latency-bound) -- the only collective;
* every rank then evaluates the SAME deterministic pairing and Metropolis test with a
  generator seeded identically on all ranks (the reference's ``self.rng``, ``:110,311,370``),
  using the reference's exponent (``_accept_swap`` ``:320-370``), so all ranks agree
  without further communication;
* an accepted pair that straddles two ranks ships its configuration point-to-point
  (positions, numbers, extra per-atom arrays, E_old, n_atoms); contiguous sharding makes
  only the G-1 boundary pairs cross-rank.  Like the reference (``_swap_states``
  ``:401-410``) temperatures / mu stay with the slot and both slots refresh their free
  volume after a swap.

Replica objects are the reference's ``GrandCanonicalEnsemble`` (or anything with the
same attributes); they are built by ``gcmc_factory(T=..., rank=i)`` / ``(mu=..., rank=i)``
exactly as for the reference class, but only for the local indices.
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional

import numpy as np

from ..utils.chunking import shard_range

logger = logging.getLogger(__name__)


class _SharedRandom:
    """``mcpy.utils.RandomNumberGenerator`` draws: ``random.Random(seed).uniform(0, 1)``."""

    def __init__(self, seed):
        from random import Random
        self.random = Random(seed)

    def get_uniform(self, a=0.0, b=1.0):
        return self.random.uniform(a, b)


def swap_exponent(rec_i, rec_j, n_species):
    """ln(w'/w) of exchanging the configurations of slots i and j.

    ``rec = [E, beta, N_1..N_S, mu_1..mu_S, Lambda_1..Lambda_S]``.  Restates
    ``BatchedReplicaExchange._accept_swap`` / ``_de_broglie_cross_term``
    (``batched_replica_exchange.py:339-353, 372-386``) in the same operation order:
    ``beta_i Phi_ii + beta_j Phi_jj - beta_i Phi_ji - beta_j Phi_ij`` with
    ``Phi_Z^(k) = E_Z - sum_s mu_k,s N_Z,s`` accumulated species by species, plus
    ``3 sum_s (N_X,s - N_Y,s) ln(Lambda_i,s / Lambda_j,s)`` when the betas differ.
    """
    S = n_species
    Ei, bi, Ni, mui, Li = rec_i[0], rec_i[1], rec_i[2:2 + S], rec_i[2 + S:2 + 2 * S], rec_i[2 + 2 * S:2 + 3 * S]
    Ej, bj, Nj, muj, Lj = rec_j[0], rec_j[1], rec_j[2:2 + S], rec_j[2 + S:2 + 2 * S], rec_j[2 + 2 * S:2 + 3 * S]

    def phi(E, N, mu):
        score = E
        for s in range(S):
            score -= mu[s] * N[s]
        return score

    phi_ii, phi_jj = phi(Ei, Ni, mui), phi(Ej, Nj, muj)
    phi_ji, phi_ij = phi(Ej, Nj, mui), phi(Ei, Ni, muj)
    delta = bi * phi_ii + bj * phi_jj - bi * phi_ji - bj * phi_ij
    if bi != bj:
        term = 0.0
        for s in range(S):
            dn = int(Ni[s]) - int(Nj[s])
            if dn:
                term += 3.0 * dn * np.log(Li[s] / Lj[s])
        delta += term
    return delta


class ShardedReplicaExchange:
    def __init__(self, gcmc_factory: Callable, calculator, temperatures: Optional[List[float]] = None,
                 mus: Optional[List[Dict[str, float]]] = None, gcmc_steps: int = 100,
                 exchange_interval: int = 10, seed: int = 31, rank: Optional[int] = None,
                 world_size: Optional[int] = None, group=None, device: str = 'cpu') -> None:
        if temperatures is None and mus is None:
            raise ValueError('Provide either temperatures or mus (one per replica).')
        if temperatures is not None and mus is not None:
            raise ValueError('Pass temperatures OR mus, not both.')
        if not hasattr(calculator, 'get_potential_energies'):
            raise TypeError('calculator must implement get_potential_energies(atoms_list). '
                            'Use B200Calculator or wrap your calculator.')
        self._dist = None
        if world_size is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    self._dist = dist
                    rank, world_size = dist.get_rank(group), dist.get_world_size(group)
            except ImportError:
                pass
            if world_size is None:
                rank, world_size = 0, 1
        elif world_size > 1:
            import torch.distributed as dist
            self._dist = dist
        self.rank, self.world_size, self.group, self.device = int(rank), int(world_size), group, device
        ladder = list(temperatures) if temperatures is not None else list(mus)
        self.temperatures = list(temperatures) if temperatures is not None else None
        self.mus = list(mus) if mus is not None else None
        self.n_replicas = len(ladder)
        self.lo, self.hi = shard_range(self.n_replicas, self.rank, self.world_size)
        key = 'T' if temperatures is not None else 'mu'
        self.replicas = [gcmc_factory(**{key: ladder[i]}, rank=i) for i in range(self.lo, self.hi)]
        shared = [a for a in ('atoms', 'move_selector')
                  if len({id(getattr(r, a, None)) for r in self.replicas}) < len(self.replicas)]
        if shared:
            raise ValueError(f"replicas share {', '.join(shared)}; gcmc_factory must build a fresh "
                             'ensemble with its own atoms, cells and move_selector per replica')
        self.calculator = calculator
        self.gcmc_steps = gcmc_steps
        self.exchange_interval = exchange_interval
        self.rng = _SharedRandom(seed)                 # identical stream on every rank
        self.exchange_attempts = [0] * self.n_replicas
        self.exchange_successes = [0] * self.n_replicas
        self.swap_log = []                             # (step, i, j, delta, accepted) on every rank
        self._re_step = 0

    # ------------------------------------------------------------------ helpers
    def owner(self, i: int) -> int:
        for g in range(self.world_size):
            lo, hi = shard_range(self.n_replicas, g, self.world_size)
            if lo <= i < hi:
                return g
        raise IndexError(i)

    def local(self, i: int):
        return self.replicas[i - self.lo]

    def _species(self) -> List[str]:
        """Species in the order the replicas' ``_mu`` dicts list them (the reference sums
        ``mu_s N_s`` in that order, ``grand_canonical_ensemble.py:188-195``)."""
        keys: List[str] = []
        for r in self.replicas:
            for k in (getattr(r, '_mu', None) or {}):
                if k not in keys:
                    keys.append(k)
        if self._dist is not None and self.world_size > 1:
            gathered = [None] * self.world_size
            self._dist.all_gather_object(gathered, keys, group=self.group)
            keys = []
            for ks in gathered:
                for k in ks:
                    if k not in keys:
                        keys.append(k)
        return keys

    # ------------------------------------------------------------------ run
    def run(self) -> None:
        for r in self.replicas:
            getattr(r, 'initialize_run', lambda: None)()
        self.species = self._species()
        self._rebatch_initial_energies()
        try:
            for step in range(self.gcmc_steps):
                self._batched_gcmc_step()
                if step > 0 and step % self.exchange_interval == 0:
                    self._attempt_exchanges(step)
                self._re_step += 1
        finally:
            for r in self.replicas:
                getattr(r, 'finalize_run', lambda: None)()

    def _rebatch_initial_energies(self) -> None:
        if not self.replicas:
            return
        energies = self.calculator.get_potential_energies([r.atoms for r in self.replicas])
        for r, E in zip(self.replicas, energies):
            r.E_old = float(E)

    # ----------------------------------------------------------- batched MC
    def _batched_gcmc_step(self) -> None:
        if not self.replicas:
            return
        n_sub = max(r.move_selector.n_moves for r in self.replicas)
        for k in range(n_sub):
            self._batched_single_move([i for i, r in enumerate(self.replicas)
                                       if k < r.move_selector.n_moves])
        for r in self.replicas:
            r._step = getattr(r, '_step', 0) + 1

    def _batched_single_move(self, active: List[int]) -> None:
        """Propose on each active local replica, ONE batched energy call, per-replica acceptance
        (reference ``_batched_single_move``, ``batched_replica_exchange.py:235-302``)."""
        snapshots, constraint_snapshots, trial_meta = {}, {}, {}
        for i in active:
            r = self.replicas[i]
            snapshots[i] = {k: v.copy() for k, v in r.atoms.arrays.items()}
            constraint_snapshots[i] = [c.copy() for c in getattr(r.atoms, 'constraints', [])]
            result = r.move_selector.do_trial_move(r.atoms)
            atoms_new, delta_particles, species = result if isinstance(result, tuple) else (result, 0, None)
            if atoms_new is False or atoms_new is None:
                r.atoms.arrays = snapshots[i]
                if hasattr(r.atoms, 'set_constraint'):
                    r.atoms.set_constraint(constraint_snapshots[i])
                continue
            if atoms_new is not r.atoms:
                raise RuntimeError('GCMC moves must mutate the passed atoms in place')
            trial_meta[i] = (delta_particles, species)
        viable = [i for i in active if i in trial_meta]
        if not viable:
            return
        energies = self.calculator.get_potential_energies([self.replicas[i].atoms for i in viable])
        for i, E_new in zip(viable, energies):
            r = self.replicas[i]
            delta_particles, species = trial_meta[i]
            delta_E = float(E_new) - r.E_old
            volume = r.move_selector.get_volume()
            n_exchange = r.move_selector.get_exchange_count()
            if n_exchange is None:
                n_exchange = r.n_atoms
            if r._acceptance_condition(delta_E, delta_particles, volume, species, n_exchange):
                if getattr(r, '_wrap_on_accept', False):
                    r.atoms.wrap()
                r.n_atoms = len(r.atoms)
                r.E_old = float(E_new)
                r.move_selector.acceptance_counter()
                r.calculate_cells_volume(r.atoms)
                getattr(r, '_record_minimum', lambda a, e: None)(r.atoms, r.E_old)
            else:
                r.atoms.arrays = snapshots[i]
                if hasattr(r.atoms, 'set_constraint'):
                    r.atoms.set_constraint(constraint_snapshots[i])

    # --------------------------------------------------------- exchange
    def _record(self, r) -> np.ndarray:
        S = len(self.species)
        rec = np.zeros(2 + 3 * S)
        rec[0], rec[1] = r.E_old, r.units.beta
        mu = getattr(r, '_mu', None) or {}
        for k, s in enumerate(self.species):
            if s in mu:
                rec[2 + k] = r._species_count(r.atoms, s)
                rec[2 + S + k] = mu[s]
                rec[2 + 2 * S + k] = r.units.lambda_dbs[s]
            else:
                rec[2 + 2 * S + k] = 1.0
        return rec

    def _gather_records(self) -> np.ndarray:
        """The only collective: all_gather of the per-replica records (padded to equal shards)."""
        S = len(self.species)
        width = 2 + 3 * S
        local = np.array([self._record(r) for r in self.replicas]).reshape(-1, width)
        if self.world_size == 1 or self._dist is None:
            return local
        import torch
        per = max(shard_range(self.n_replicas, g, self.world_size)[1]
                  - shard_range(self.n_replicas, g, self.world_size)[0] for g in range(self.world_size))
        buf = torch.zeros(per, width, dtype=torch.float64, device=self.device)
        if len(local):
            buf[:len(local)] = torch.from_numpy(local).to(self.device)
        out = [torch.empty_like(buf) for _ in range(self.world_size)]
        self._dist.all_gather(out, buf, group=self.group)
        rows = []
        for g in range(self.world_size):
            lo, hi = shard_range(self.n_replicas, g, self.world_size)
            rows.append(out[g][:hi - lo].cpu().numpy())
        return np.concatenate(rows)

    def _attempt_exchanges(self, step: int = 0) -> None:
        offset = 0 if self.rng.get_uniform() < 0.5 else 1
        records = self._gather_records()
        S = len(self.species)
        for i in range(offset, self.n_replicas - 1, 2):
            j = i + 1
            self.exchange_attempts[i] += 1
            self.exchange_attempts[j] += 1
            delta = swap_exponent(records[i], records[j], S)
            if not np.isfinite(delta):
                logger.warning('swap %d<->%d: non-finite acceptance exponent (delta=%r); rejecting.',
                               i, j, delta)
                accepted = False
            else:
                p = 1.0 if delta >= 0.0 else float(np.exp(delta))
                accepted = self.rng.get_uniform() < p
            self.swap_log.append((step, i, j, float(delta), bool(accepted)))
            if accepted:
                self._swap_states(i, j)
                self.exchange_successes[i] += 1
                self.exchange_successes[j] += 1

    # ------------------------------------------------------- state transport
    @staticmethod
    def _pack(r):
        atoms = r.atoms
        extra = sorted(k for k in atoms.arrays if k not in ('positions', 'numbers'))
        n = len(atoms)
        parts = [np.asarray(atoms.positions, dtype=np.float64).reshape(-1),
                 np.asarray(atoms.numbers, dtype=np.float64)]
        for k in extra:
            parts.append(np.asarray(atoms.arrays[k], dtype=np.float64).reshape(-1))
        head = np.array([n, r.E_old, r.n_atoms, len(extra)], dtype=np.float64)
        return head, np.concatenate(parts) if n else np.zeros(0), extra

    @staticmethod
    def _unpack(r, head, payload, extra_names, template_atoms):
        n = int(head[0])
        pos = payload[:3 * n].reshape(n, 3).copy()
        num = payload[3 * n:4 * n].astype(int)
        atoms = type(template_atoms)(numbers=num, positions=pos, cell=np.asarray(template_atoms.cell),
                                     pbc=template_atoms.pbc)
        cur = 4 * n
        for k in extra_names:
            old = template_atoms.arrays[k]
            width = int(np.prod(old.shape[1:])) if old.ndim > 1 else 1
            vals = payload[cur:cur + n * width].reshape((n,) + old.shape[1:]).astype(old.dtype)
            atoms.arrays[k] = vals
            cur += n * width
        r.atoms = atoms
        r.E_old = float(head[1])
        r.n_atoms = int(head[2])

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
        import torch
        mine, peer = (i, oj) if self.rank == oi else (j, oi)
        r = self.local(mine)
        head, payload, extra = self._pack(r)
        dist = self._dist
        t_head = torch.from_numpy(head).to(self.device)
        r_head = torch.empty_like(t_head)
        first = self.rank < peer
        # fixed-size header first (sizes), then the payload; lower rank sends first
        if first:
            dist.send(t_head, peer, group=self.group); dist.recv(r_head, peer, group=self.group)
        else:
            dist.recv(r_head, peer, group=self.group); dist.send(t_head, peer, group=self.group)
        n_peer = int(r_head[0].item())
        if int(r_head[3].item()) != len(extra):
            raise RuntimeError('replicas exchange configurations with different per-atom arrays')
        widths = 4
        for k in extra:
            a = r.atoms.arrays[k]
            widths += int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        t_pay = torch.from_numpy(np.ascontiguousarray(payload)).to(self.device)
        r_pay = torch.empty(n_peer * widths, dtype=torch.float64, device=self.device)
        if first:
            if t_pay.numel():
                dist.send(t_pay, peer, group=self.group)
            if r_pay.numel():
                dist.recv(r_pay, peer, group=self.group)
        else:
            if r_pay.numel():
                dist.recv(r_pay, peer, group=self.group)
            if t_pay.numel():
                dist.send(t_pay, peer, group=self.group)
        self._unpack(r, r_head.cpu().numpy(), r_pay.cpu().numpy(), extra, r.atoms)
        r.calculate_cells_volume(r.atoms)
