"""Build recipe for ``libmcpy_b200.so`` (hand-written CUDA, sm_100a only).

``python -m mcpy_b200.build`` or ``__graft_entry__.build()`` runs one
``nvcc -shared`` over ``csrc/engine.cu`` (which includes every kernel).  The
library is built IN-TREE next to this file so it travels with the repository
snapshot; there is no JIT cache and no CPU fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmcpy_b200.so')
SOURCES = ['engine.cu']
HEADERS = ['common.cuh', 'cell_list.cuh', 'stencil.cuh', 'lj_kernels.cuh', 'emt_kernels.cuh',
           'free_volume.cuh', 'pcg64.cuh', 'site_kernels.cuh', 'rows_block.cuh', 'rows_warp.cuh', 'chain.cuh', 'emt_rows.cuh', 'gcmc_chain.cuh', os.path.join('..', '..', 'include', 'mcpy_b200.h')]

NVCC_FLAGS = [
    '-O3', '-std=c++17',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '--extended-lambda',
    '-shared', '-Xcompiler', '-fPIC',
]


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: the mcpy_b200 engine is CUDA-only and cannot be built')
    return nvcc


def marshal_path() -> str:
    import sysconfig
    return os.path.join(HERE, '_marshal' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))


def build_marshal(force: bool = False) -> str:
    """The CPython helper that walks an atoms list for the batched call (csrc/marshal.cpp; plain gcc, no CUDA)."""
    import sysconfig
    src, out = os.path.join(CSRC, 'marshal.cpp'), marshal_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cc = shutil.which('g++') or shutil.which('c++')
    if not cc:
        raise RuntimeError('g++ not found: cannot build the marshalling helper')
    import numpy
    cmd = [cc, '-O2', '-std=c++17', '-shared', '-fPIC', '-Wall', '-I' + sysconfig.get_paths()['include'], '-I' + numpy.get_include(),
           '-o', out, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('gcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


LIB_BOUNDS = os.path.join(HERE, 'libmcpy_b200_bounds.so')


def build(force: bool = False, verbose: bool = False, bounds: bool = False) -> str:
    """Compile the engine for sm_100a; returns the path of the shared library.  ``bounds=True`` builds the checking
    variant (``-DMCPYB_BOUNDS``: device-side asserts on the hand-computed indices, csrc/common.cuh) next to the
    product library; ``MCPYB_LIBRARY=<path>`` makes ``mcpy_b200`` load it."""
    build_marshal(force)
    if bounds:
        cmd = [find_nvcc()] + NVCC_FLAGS + ['-DMCPYB_BOUNDS=1', '-o', LIB_BOUNDS] + [os.path.join(CSRC, s) for s in SOURCES]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
        return LIB_BOUNDS
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, bounds='--bounds' in sys.argv))
