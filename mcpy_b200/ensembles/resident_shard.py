"""A shard of a replica ladder kept RESIDENT in HBM, for batched trial scoring between exchange steps.

``ShardedReplicaExchange`` drives the reference's host-side replicas one trial per replica per call.
This class is the device-resident form of the same ladder for the batched trial kernel (many proposals
per replica per launch, ``Engine.trial_delta_e_dev``): the configurations of the local replicas live in
one device array, the engine's cell lists are built from it in place, and an exchange attempt moves
configurations between SLOTS without a host round trip of the atoms:

* ``attempt_exchanges()`` gathers one record per replica ``[E, beta, N, mu, Lambda, has, n_atoms, cell, pbc]`` with the
  ladder's only collective (``LadderComm.gather_records``), takes the even / odd pairing and the Metropolis
  decisions from a generator seeded identically on every rank with the reference's exponent
  (``swap_exponent`` = ``BatchedReplicaExchange._accept_swap``, ``batched_replica_exchange.py:320-370``),
  swaps accepted local pairs inside the device array and rebuilds the engine's cell lists from the device
  array (``Engine.upload_dev``).  The configurations of a rank's two BOUNDARY slots ride in the same
  all_gather as the records (2 x 64 KB for 2 000 atoms; NVLink moves that in microseconds, a second
  round of point-to-point calls costs more in launch latency than the bytes do), so an accepted cross-rank
  swap finds its incoming configuration already on the device: one collective per exchange attempt.  As in the reference, beta / mu stay with the slot,
  E and the configuration travel together (``_swap_states``, ``:401-410``).

One exchanged species (every atom of the configuration counts), one box per replica.
"""
from __future__ import annotations

from random import Random

import numpy as np

from .ladder_comm import LadderComm
from .acceptance import swap_exponent

_WIDTH = 19                  # E, beta, N, mu, Lambda, has_mu, n_atoms, cell (9), pbc (3)


def _pair_exponent(ri, rj):
    """``swap_exponent(rec_i, rec_j, 1)`` for records ``[E, beta, N, mu, Lambda, has]`` held as Python floats."""
    Ei, bi, Ni, mui, Li, hi = ri
    Ej, bj, Nj, muj, Lj, hj = rj
    phi_ii = Ei - mui * Ni if hi else Ei
    phi_jj = Ej - muj * Nj if hj else Ej
    phi_ji = Ej - mui * Nj if hi else Ej
    phi_ij = Ei - muj * Ni if hj else Ei
    delta = bi * phi_ii + bj * phi_jj - bi * phi_ji - bj * phi_ij
    if bi != bj and hi:
        dn = int(Ni) - int(Nj)
        if dn:
            if Li != Li or Lj != Lj:
                raise KeyError('a replica has no de Broglie wavelength for the exchanged species')
            delta += 0.0 + 3.0 * dn * np.log(Li / Lj)
    return delta


class ResidentLadderShard:
    def __init__(self, engine, configs, betas, mus, lambdas=None, seed=31, rank=None, world_size=None,
                 group=None, device=None):
        """``configs``: the LOCAL replicas, each ``(positions (n, 3), numbers (n,), cell (3, 3), pbc (3,))``;
        ``betas`` / ``mus`` (/ ``lambdas``): the GLOBAL ladder, one value per slot."""
        import torch
        self.torch = torch
        self.engine = engine
        self.n_total = len(betas)
        if len(mus) != self.n_total:
            raise ValueError('betas and mus must list every slot of the ladder')
        self.device = device if device is not None else f'cuda:{engine.device}'
        self.comm = LadderComm(self.n_total, rank=rank, world_size=world_size, group=group, device=self.device)
        self.lo, self.hi = self.comm.lo, self.comm.hi
        if len(configs) != self.hi - self.lo:
            raise ValueError(f'rank {self.comm.rank} owns slots [{self.lo}, {self.hi}) but got {len(configs)} configurations')
        self.betas = np.asarray(betas, dtype=float)
        self.mus = np.asarray(mus, dtype=float)
        self.lambdas = np.ones(self.n_total) if lambdas is None else np.asarray(lambdas, dtype=float)
        self.rng = Random(seed)                        # mcpy.utils.RandomNumberGenerator(seed).get_uniform()
        self.n_of = [len(c[0]) for c in configs]
        self.cells = np.stack([np.asarray(c[2], dtype=float).reshape(9) for c in configs]) if configs else np.zeros((0, 9))
        self.pbcs = np.stack([np.asarray(c[3], dtype=bool).reshape(3) for c in configs]) if configs else np.zeros((0, 3), bool)
        dev = self.device
        cat = np.concatenate([np.asarray(c[0], dtype=np.float64).reshape(-1, 3) for c in configs]) if configs else np.zeros((0, 3))
        num = np.concatenate([np.asarray(c[1], dtype=np.int32).reshape(-1) for c in configs]) if configs else np.zeros(0, np.int32)
        self.d_pos = torch.from_numpy(np.ascontiguousarray(cat)).to(dev)
        self.d_num = torch.from_numpy(np.ascontiguousarray(num)).to(dev)
        if str(dev) != 'cpu':
            # the engine builds its cell lists from tensors torch has just written: one stream for both
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        # capacity of a boundary block of the exchange buffer: the largest replica of the whole ladder
        self.cap = max(max(self.comm.all_gather_object(max(self.n_of, default=0))), 1)
        self._xbuf = None
        self._xnp = None
        self.E = np.zeros(len(configs))
        self.swap_log = []
        self.rounds = 0
        self.upload(energies=True)

    # ------------------------------------------------------------------ residency
    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.n_of)]).astype(np.int64)

    def upload(self, energies=False):
        """(Re)build the engine's cell lists from the device array; optionally read the replica energies."""
        self.engine.upload_dev(self.offsets(), self.d_pos.data_ptr(), self.d_num.data_ptr(), self.cells, self.pbcs)
        if energies and len(self.n_of):
            self.E = self.engine.energies().copy()

    def slot_rows(self, k):
        off = self.offsets()
        return int(off[k]), int(off[k + 1])

    # ------------------------------------------------------------------ exchange
    def records(self):
        nloc = self.hi - self.lo
        rec = np.zeros((nloc, _WIDTH))
        if nloc:
            sl = slice(self.lo, self.hi)
            rec[:, 0] = self.E
            rec[:, 1] = self.betas[sl]
            rec[:, 2] = self.n_of
            rec[:, 3] = self.mus[sl]
            rec[:, 4] = self.lambdas[sl]
            rec[:, 5] = 1.0
            rec[:, 6] = self.n_of
            rec[:, 7:16] = self.cells
            rec[:, 16:19] = self.pbcs
        return rec

    def _gather(self):
        """Records of every replica (host) and, for world > 1, the device buffer holding every rank's two boundary
        configurations: ONE all_gather, packed by ONE launch (the records are read from page-locked memory where
        they are; ``Engine.copy_segments_dev``)."""
        rec = self.records()
        comm = self.comm
        if comm.world_size == 1 or comm._dist is None:
            return rec, None
        import time
        torch = self.torch
        t0 = time.perf_counter()
        nloc, per, cap = self.hi - self.lo, comm.per, self.cap
        blk = 4 * cap                                      # positions (3 cap) + numbers (cap), as float64
        width = per * _WIDTH + 2 * blk
        if self._xbuf is None or self._xbuf[0].shape[0] != width:
            self._xbuf = (torch.zeros(width, dtype=torch.float64, device=self.device),
                          torch.zeros(comm.world_size * width, dtype=torch.float64, device=self.device),
                          torch.zeros(per * _WIDTH, dtype=torch.float64, pin_memory=comm.on_gpu),
                          torch.zeros(comm.world_size, per * _WIDTH, dtype=torch.float64, pin_memory=comm.on_gpu))
            self._xnp = (self._xbuf[2].numpy(), self._xbuf[3].numpy())
        d_in, d_out, h_in, h_out = self._xbuf
        np_in, np_out = self._xnp
        np_in[:nloc * _WIDTH] = rec.reshape(-1)
        np_in[nloc * _WIDTH:] = 0.0
        din = d_in.data_ptr()
        segs = [(h_in.data_ptr(), din, per * _WIDTH, 0)]
        if nloc:
            p0, z0 = self.d_pos.data_ptr(), self.d_num.data_ptr()
            off = self.offsets()
            for which, k in ((0, 0), (1, nloc - 1)):
                a, n = int(off[k]), int(off[k + 1] - off[k])
                if n > cap:
                    raise RuntimeError(f'replica with {n} atoms exceeds the exchange capacity {cap}')
                base = din + 8 * (per * _WIDTH + which * blk)
                segs.append((p0 + 24 * a, base, 3 * n, 0))
                segs.append((z0 + 4 * a, base + 8 * 3 * cap, n, 3))
        self.engine.copy_segments_dev(segs)
        comm._dist.all_gather_into_tensor(d_out, d_in, group=comm.group)
        h_out.copy_(d_out.view(comm.world_size, width)[:, :per * _WIDTH], non_blocking=True)
        comm._sync()
        out = np_out.reshape(comm.world_size, per, _WIDTH)
        rows = np.concatenate([out[g, :hi - lo] for g, (lo, hi) in enumerate(comm._bounds)])
        comm.collective_seconds += time.perf_counter() - t0
        comm.gathers += 1
        return rows, d_out.view(comm.world_size, width)

    def attempt_exchanges(self):
        """One exchange round; returns the number of accepted swaps (identical on every rank)."""
        torch = self.torch
        offset = 0 if self.rng.uniform(0.0, 1.0) < 0.5 else 1
        rec, gathered = self._gather()
        self.rounds += 1
        accepted_pairs = []
        # the six leading columns as Python floats: the same IEEE operations as swap_exponent on numpy scalars, in
        # the same order, at a twentieth of the cost (every rank evaluates every pair of the ladder: 32 per round
        # at 64 replicas)
        cols = rec[:, :6].tolist()
        for i in range(offset, self.n_total - 1, 2):
            j = i + 1
            delta = _pair_exponent(cols[i], cols[j])
            if not np.isfinite(delta):
                ok = False
            else:
                p = 1.0 if delta >= 0.0 else float(np.exp(delta))
                ok = self.rng.uniform(0.0, 1.0) < p
            self.swap_log.append((self.rounds, i, j, float(delta), bool(ok)))
            if ok:
                accepted_pairs.append((i, j))
        if not accepted_pairs:
            return 0
        me, comm = self.comm.rank, self.comm
        off = self.offsets()
        # every slot's block as (address of its rows, address of its numbers, copy kind of the numbers, rows): blocks
        # that change slot are re-pointed, then ONE launch writes the new device arrays (Engine.copy_segments_dev)
        p0, z0 = self.d_pos.data_ptr(), self.d_num.data_ptr()
        blocks = [[p0 + 24 * int(off[k]), z0 + 4 * int(off[k]), 1, int(off[k + 1] - off[k])] for k in range(len(self.n_of))]
        touched = False
        for i, j in accepted_pairs:
            oi, oj = comm.owner(i), comm.owner(j)
            if oi == me and oj == me:
                a, b = i - self.lo, j - self.lo
                blocks[a], blocks[b] = blocks[b], blocks[a]
                self.n_of[a], self.n_of[b] = self.n_of[b], self.n_of[a]
                self.E[a], self.E[b] = self.E[b], self.E[a]
                self.cells[[a, b]] = self.cells[[b, a]]
                self.pbcs[[a, b]] = self.pbcs[[b, a]]
                touched = True
            elif me in (oi, oj):
                # the partner's configuration is its rank's boundary block of the gathered buffer: slot i is the LAST
                # slot of the lower rank, slot j the FIRST slot of the upper one
                mine, theirs, peer, which = (i, j, oj, 0) if me == oi else (j, i, oi, 1)
                k = mine - self.lo
                n_in = int(rec[theirs, 6])
                cap, per = self.cap, comm.per
                base = per * _WIDTH + which * 4 * cap
                row0 = gathered.data_ptr() + 8 * (peer * gathered.shape[1] + base)
                blocks[k] = [row0, row0 + 8 * 3 * cap, 2, n_in]
                self.n_of[k] = n_in
                self.E[k] = rec[theirs, 0]
                self.cells[k] = rec[theirs, 7:16]
                self.pbcs[k] = rec[theirs, 16:19] != 0
                comm.transports += 1
                touched = True
        if touched:
            total = sum(b[3] for b in blocks)
            new_pos = torch.empty((total, 3), dtype=torch.float64, device=self.device)
            new_num = torch.empty(total, dtype=torch.int32, device=self.device)
            segs, at = [], 0
            dp, dz = new_pos.data_ptr(), new_num.data_ptr()
            for src_p, src_z, kind, n in blocks:
                if n:
                    segs.append((src_p, dp + 24 * at, 3 * n, 0))
                    segs.append((src_z, dz + 4 * at, n, kind))
                at += n
            self.engine.copy_segments_dev(segs)
            self.d_pos, self.d_num = new_pos, new_num
            # the cell lists are rebuilt in place from the new arrays
            self.upload()
        return len(accepted_pairs)

    def positions_host(self):
        """The local configurations back on the host (tests): list of ``(positions, numbers)``."""
        off = self.offsets()
        p, z = self.d_pos.cpu().numpy(), self.d_num.cpu().numpy()
        return [(p[off[k]:off[k + 1]].copy(), z[off[k]:off[k + 1]].copy()) for k in range(len(self.n_of))]
