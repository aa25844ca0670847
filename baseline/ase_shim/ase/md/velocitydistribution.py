def MaxwellBoltzmannDistribution(*args, **kwargs):
    raise NotImplementedError('MD is outside the hot path (shim)')
