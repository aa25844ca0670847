"""``ResidentReplicaExchange`` -- the reference's batched replica-exchange run with every replica's trials on the device.

``BatchedReplicaExchange.run`` (``mcpy/ensembles/batched_replica_exchange.py:133-174``) advances all replicas one
sub-move at a time: propose on the host, ONE batched calculator call, accept / reject on the host (``:235-302``).  The
replicas' trial sequences do not depend on each other between two exchange attempts, so this driver -- a subclass of
``ShardedReplicaExchange``, hence of the reference class, whose ``run`` / exchange rule / output it inherits unchanged
-- hands the local replicas to :class:`ResidentGCMC` and advances all of them in ONE launch (a CTA per replica) up to
the next step at which the host looks at them: an exchange attempt, a row of the replica-exchange outfile, a
replica's own outfile row or trajectory frame, the end of the run.  Between two such steps only ``E_old``,
``n_atoms`` and the selector's counters come back; positions and generator states stay on the device, and an accepted
swap of two local replicas exchanges their configurations there (``mcpyb_gcmc_swap_configurations``).  A swap across
two ranks sends the boundary replica's configuration block device to device over NCCL (``ResidentGCMC.remote_swap``)
when both sides hold the same storage layout; otherwise the replica comes back to the host and travels as a whole
``Atoms`` as in ``ShardedReplicaExchange``.

Each replica draws, accepts and rejects exactly as under the reference driver (``tests/test_gpu_resident_gcmc.py``
runs the two side by side).  The replicas must be in ``ResidentGCMC``'s scope (``resident_gcmc.supports``).
"""
from __future__ import annotations

from .resident_gcmc import ResidentGCMC
from .sharded_replica_exchange import ShardedReplicaExchange


class ResidentReplicaExchange(ShardedReplicaExchange):
    chains_class = ResidentGCMC            # what advances the local replicas (the CPU tests put a host stand-in here)

    def __init__(self, *args, lj=None, engine=None, atom_capacity=None, cell_capacity=0, **kw):
        super().__init__(*args, **kw)
        self.resident = None
        if self.replicas:
            if engine is None:
                engine = getattr(self.calculator, 'engine', None)
            self.resident = self.chains_class(self.replicas, engine=engine, lj=lj, atom_capacity=atom_capacity,
                                              cell_capacity=cell_capacity)
        import numpy as np
        self._trials_per_step = np.array([r.move_selector.n_moves for r in self.replicas], dtype=np.int64)
        self._ahead = 0                    # steps the device is ahead of the run loop's step counter
        self.device_ms = 0.0               # device time of the chain launches
        self.device_swaps = 0              # cross-rank swaps that went device to device
        self._rec_static = None            # the slot-bound columns of the exchange records
        self.blocks = 0

    # ------------------------------------------------------------------ blocks of steps
    @staticmethod
    def _file_due(r, step_after):
        return ((r._outfile is not None and step_after % r._outfile_write_interval == 0),
                (r._traj_file is not None and step_after % r._trajectory_write_interval == 0))

    def _block(self):
        """(steps, needs positions): how far the replicas can run before the host looks at them.  The run loop is at
        step ``self._re_step``; after its step ``s`` it attempts exchanges when ``s > 0 and s % exchange_interval == 0``
        and writes a row when ``s % write_out_interval == 0`` (``batched_replica_exchange.py:153-160``)."""
        step = self._re_step
        files = [r for r in self.replicas if r._outfile is not None or r._traj_file is not None]
        b = 1
        while True:
            s = step + b - 1
            frames = [self._file_due(r, r._step + b) for r in files]
            last = s >= self.gcmc_steps - 1
            if last or (s > 0 and s % self.exchange_interval == 0) or s % self.write_out_interval == 0 \
                    or any(o or t for o, t in frames):
                return b, last or any(t for _, t in frames)
            b += 1

    def _batched_gcmc_step(self) -> None:
        if not self.replicas:                          # a rank may own no replica (R < world size)
            return
        if self._ahead == 0:
            b, frames = self._block()
            last = self._re_step + b - 1 >= self.gcmc_steps - 1
            # the selector's counters are only read by a replica's own outfile rows (and after the run)
            counters = last or any(r._outfile is not None for r in self.replicas)
            self.device_ms += self.resident.advance(self._trials_per_step * b if b > 1 else self._trials_per_step,
                                                    sync='full' if last else 'scalars', counters=counters)
            if frames and not last:                    # a trajectory frame is due: positions only, the chains keep the state
                due = [k for k, r in enumerate(self.replicas) if self._file_due(r, r._step + b)[1]]
                if hasattr(self.resident, 'sync_positions'):
                    self.resident.sync_positions(due)
                else:
                    self.resident.sync_host(due)
            self._ahead = b
            self.blocks += 1
        self._ahead -= 1
        for r in self.replicas:                        # the tail of the reference's step (:226-231)
            r._step += 1
            if r._outfile is not None or r._traj_file is not None:
                out_due, traj_due = self._file_due(r, r._step)
                if out_due:
                    r.write_outfile()
                if traj_due:
                    r.write_coordinates(r.atoms, r.E_old)

    # ------------------------------------------------------------------ records
    def _gather_records(self):
        """The records of ``ShardedReplicaExchange._record`` without a Python pass over every replica's atoms: in the
        resident chains' scope a replica holds atoms of its one species only, so N_s is ``n_atoms``; beta, mu and Lambda
        belong to the slot and are read once."""
        import numpy as np
        if self.resident is None or len(self.species) != 1 or not hasattr(self.resident, 'remote_swap'):
            return super()._gather_records()
        if self._rec_static is None:
            self._rec_static = np.array([self._record(r) for r in self.replicas]).reshape(-1, 6)
        rec = self._rec_static.copy()
        rec[:, 0] = [r.E_old for r in self.replicas]
        rec[:, 2] = [r.n_atoms for r in self.replicas]
        return self.comm.gather_records(rec)

    # ------------------------------------------------------------------ swaps
    def _swap_states(self, i: int, j: int) -> None:
        oi, oj = self.owner(i), self.owner(j)
        if self.resident is not None:
            if oi == self.rank and oj == self.rank:
                self.resident.swap_configurations(i - self.lo, j - self.lo)
            elif self.rank in (oi, oj):
                mine, peer = (i, oj) if self.rank == oi else (j, oi)
                if self.comm.on_gpu and hasattr(self.resident, 'remote_swap') and \
                        self.resident.remote_swap(mine - self.lo, self.comm, peer):
                    self.device_swaps += 1
                    r = self.local(mine)
                    r.calculate_cells_volume(r.atoms)
                    return
                # the boundary replica travels as a whole Atoms: bring it back first; it is uploaded again afterwards
                self.resident.sync_host([mine - self.lo])
        super()._swap_states(i, j)

    def run(self) -> None:
        try:
            super().run()
        finally:
            if self.resident is not None:
                self.resident.sync_host()
