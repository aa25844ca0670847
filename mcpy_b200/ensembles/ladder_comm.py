"""The communication of a replica ladder sharded over GPUs.

Two primitives, both tiny and latency-bound, both over ``torch.distributed`` (NCCL on device tensors over
NVLink / NVSwitch, gloo on CPU tensors in the tests):

* :meth:`LadderComm.gather_records` -- the ONLY collective of the path: one ``all_gather`` of the fixed-size
  fp64 record of every replica (what ``BatchedReplicaExchange._accept_swap`` reads from the two replicas of a
  pair, ``mcpy/ensembles/batched_replica_exchange.py:344-353,382-386``), padded to equal shards, through
  buffers allocated once;
* :meth:`LadderComm.exchange` -- point-to-point transport for an accepted swap whose two slots live on
  different ranks (what the MPI driver's ``sendrecv`` carries, ``replica_exchange.py:188-192``): every tensor a
  rank sends and receives in one exchange round travels in ONE grouped ``batch_isend_irecv``.

Both keep wall-clock tallies (``collective_seconds``, ``transport_seconds``) so the bench can name their share.
"""
from __future__ import annotations

import time

import numpy as np

from ..utils.chunking import shard_range


class LadderComm:
    def __init__(self, n_replicas_total, rank=None, world_size=None, group=None, device='cpu'):
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
        self.on_gpu = str(device) != 'cpu'
        self.n_total = int(n_replicas_total)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world_size)
        self.per = -(-self.n_total // self.world_size) if self.n_total else 0
        self._bounds = [shard_range(self.n_total, g, self.world_size) for g in range(self.world_size)]
        self._buf = None
        self.collective_seconds = 0.0
        self.transport_seconds = 0.0
        self.gathers = 0
        self.transports = 0

    # ------------------------------------------------------------------ layout
    def owner(self, i: int) -> int:
        for g, (lo, hi) in enumerate(self._bounds):
            if lo <= i < hi:
                return g
        raise IndexError(i)

    def all_gather_object(self, obj):
        if self._dist is None or self.world_size == 1:
            return [obj]
        out = [None] * self.world_size
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    def _sync(self):
        if self.on_gpu:
            import torch
            torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ the collective
    def gather_records(self, local: np.ndarray) -> np.ndarray:
        """``local``: (hi - lo, width) float64 -> (n_total, width): every replica's record on every rank."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.world_size == 1 or self._dist is None:
            return local
        import torch
        t0 = time.perf_counter()
        width = local.shape[1]
        if self._buf is None or self._buf[0].shape != (self.per, width):
            pin = self.on_gpu
            self._buf = (torch.zeros(self.per, width, dtype=torch.float64, pin_memory=pin),
                         torch.zeros(self.per, width, dtype=torch.float64, device=self.device),
                         torch.zeros(self.world_size * self.per, width, dtype=torch.float64, device=self.device),
                         torch.zeros(self.world_size * self.per, width, dtype=torch.float64, pin_memory=pin))
        h_in, d_in, d_out, h_out = self._buf
        if len(local):
            h_in[:len(local)] = torch.from_numpy(local)
        d_in.copy_(h_in, non_blocking=True)
        self._dist.all_gather_into_tensor(d_out, d_in, group=self.group)
        h_out.copy_(d_out, non_blocking=True)
        self._sync()
        out = h_out.numpy().reshape(self.world_size, self.per, width)
        rows = np.concatenate([out[g, :hi - lo] for g, (lo, hi) in enumerate(self._bounds)])
        self.collective_seconds += time.perf_counter() - t0
        self.gathers += 1
        return rows

    # ------------------------------------------------------------------ point to point
    def exchange(self, sends, recvs):
        """``sends`` / ``recvs``: lists of ``(tensor, peer_rank)``; all of them in one grouped call.  Tensors of a
        pair of ranks must be listed in the same order on both sides.  Returns after completion."""
        if not sends and not recvs:
            return
        t0 = time.perf_counter()
        dist = self._dist
        ops = [dist.P2POp(dist.isend, t, p, group=self.group) for t, p in sends] + \
              [dist.P2POp(dist.irecv, t, p, group=self.group) for t, p in recvs]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self._sync()
        self.transport_seconds += time.perf_counter() - t0
        self.transports += 1

    def exchange_bytes(self, payload: bytes, peer: int) -> bytes:
        """Swap byte strings with ``peer`` (sizes first, then the payloads)."""
        import torch
        dev = self.device
        n_mine = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
        n_peer = torch.zeros(1, dtype=torch.int64, device=dev)
        self.exchange([(n_mine, peer)], [(n_peer, peer)])
        t_pay = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
        r_pay = torch.empty(int(n_peer.item()), dtype=torch.uint8, device=dev)
        self.exchange([(t_pay, peer)], [(r_pay, peer)])
        self.transports -= 1                                   # one state transport, two grouped calls
        return r_pay.cpu().numpy().tobytes()
