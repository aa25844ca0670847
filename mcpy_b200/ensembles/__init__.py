"""Host-side mirrors of the scalar acceptance arithmetic and the sharded replica driver."""
from .acceptance import Units, acceptance_probability, accept_trial  # noqa: F401
from .sharded_replica_exchange import ShardedReplicaExchange, swap_exponent  # noqa: F401
