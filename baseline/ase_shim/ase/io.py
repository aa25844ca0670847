def read(*args, **kwargs):
    raise NotImplementedError('ase.io.read is outside the shim')


def write(*args, **kwargs):
    raise NotImplementedError('ase.io.write is outside the shim')
