"""Host-side mirrors of the scalar acceptance arithmetic and the sharded replica drivers.

``acceptance`` needs nothing but numpy.  The replica drivers extend classes of farrisric/mcpy and are imported
lazily, so that ``mcpy_b200.ensembles.acceptance`` stays usable where ``mcpy`` is not installed."""
from .acceptance import Units, acceptance_probability, accept_trial, swap_exponent  # noqa: F401

_LAZY = {'ShardedReplicaExchange': 'sharded_replica_exchange',
         'ResidentLadderShard': 'resident_shard', 'LadderComm': 'ladder_comm',
         'ResidentGCMC': 'resident_gcmc', 'ResidentReplicaExchange': 'resident_replica_exchange'}


def __getattr__(name):
    if name in _LAZY:
        from importlib import import_module
        return getattr(import_module(f'.{_LAZY[name]}', __name__), name)
    raise AttributeError(name)
