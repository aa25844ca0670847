"""ctypes binding of the C ABI in ``include/mcpy_b200.h``.

This is the reference-side stub INTEGRATION.md describes: mcpy is pure Python
(``pyproject.toml:29-34`` has no extension modules), so the binding a maintainer
adds is a ctypes loader.  There is deliberately no fallback: if the CUDA
library is missing or no GPU is present the import of the library / creation
of an engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCPYB_LIBRARY: another build of the SAME engine (the bounds-checking variant of `python -m mcpy_b200.build --bounds`)
LIB_PATH = os.environ.get('MCPYB_LIBRARY') or os.path.join(_HERE, 'libmcpy_b200.so')

# Every symbol include/mcpy_b200.h declares (tests check the library exports all of them).
SYMBOLS = [
    'mcpyb_create', 'mcpyb_destroy', 'mcpyb_last_error', 'mcpyb_version', 'mcpyb_set_stream',
    'mcpyb_synchronize', 'mcpyb_set_lj', 'mcpyb_set_emt', 'mcpyb_cutoff', 'mcpyb_emt_species',
    'mcpyb_upload_host', 'mcpyb_upload_dev', 'mcpyb_energies_host', 'mcpyb_energies_dev',
    'mcpyb_potential_energies_host', 'mcpyb_tracked_energy_host', 'mcpyb_tracked_energy_z64_host', 'mcpyb_atom_energies_host',
    'mcpyb_tracked_energies_host', 'mcpyb_tracked_energies_rows_host', 'mcpyb_forces_host', 'mcpyb_energy_forces_host', 'mcpyb_relax_lbfgs_host',
    'mcpyb_trial_delta_e_host', 'mcpyb_trial_delta_e_dev',
    'mcpyb_neighbor_pairs_host',
    'mcpyb_free_volume_count_host', 'mcpyb_free_volume_count_dev',
    'mcpyb_sphere_free_volume_pcg64', 'mcpyb_dome_free_volume_pcg64', 'mcpyb_box_free_volume_pcg64',
    'mcpyb_alloc_pinned', 'mcpyb_free_pinned',
    'mcpyb_min_dist2_host', 'mcpyb_pcg64_host_doubles', 'mcpyb_fp64_peak_tflops', 'mcpyb_stage_timing', 'mcpyb_launch_count',
    'mcpyb_resident_chain', 'mcpyb_set_precision', 'mcpyb_copy_segments_dev',
    'mcpyb_gcmc_create', 'mcpyb_gcmc_destroy', 'mcpyb_gcmc_set_state', 'mcpyb_gcmc_get_state', 'mcpyb_gcmc_run',
    'mcpyb_gcmc_get_trace', 'mcpyb_gcmc_info', 'mcpyb_gcmc_last_run_ms', 'mcpyb_gcmc_get_scalars',
    'mcpyb_gcmc_swap_configurations', 'mcpyb_gcmc_cycles', 'mcpyb_gcmc_config_ptr', 'mcpyb_gcmc_set_config_scalars',
    'mcpyb_gcmc_get_block_scalars', 'mcpyb_region_mask_host',
]

GCMC_MAX_MOVES = 8          # MCPYB_GCMC_MAX_MOVES
GCMC_MT_WORDS = 625         # MCPYB_GCMC_MT_WORDS


class GcmcDesc(C.Structure):
    """``mcpyb_gcmc_desc``."""
    _fields_ = [('box', C.c_double * 3), ('beta', C.c_double), ('mu', C.c_double), ('lambda3', C.c_double),
                ('lj_epsilon', C.c_double), ('lj_sigma', C.c_double), ('lj_rc', C.c_double),
                ('move_cum_weight', C.c_double * GCMC_MAX_MOVES), ('move_volume', C.c_double * GCMC_MAX_MOVES),
                ('move_max_displacement', C.c_double * GCMC_MAX_MOVES),
                ('move_kind', C.c_int32 * GCMC_MAX_MOVES), ('move_limit', C.c_int32 * GCMC_MAX_MOVES),
                ('move_pcg', C.c_int32 * GCMC_MAX_MOVES),
                ('n_move_types', C.c_int32), ('atom_capacity', C.c_int32), ('cell_capacity', C.c_int32)]


class GcmcScalars(C.Structure):
    """``mcpyb_gcmc_scalars``."""
    _fields_ = [('energy', C.c_double), ('best_score', C.c_double), ('best_energy', C.c_double),
                ('attempts', C.c_int64 * GCMC_MAX_MOVES), ('accepted', C.c_int64 * GCMC_MAX_MOVES),
                ('failed', C.c_int64 * GCMC_MAX_MOVES), ('trials_done', C.c_int64),
                ('pcg', (C.c_uint64 * 4) * GCMC_MAX_MOVES),
                ('n_atoms', C.c_int32), ('best_n', C.c_int32), ('last_move', C.c_int32), ('status', C.c_int32)]


GCMC_TRIAL_DTYPE = np.dtype([('dE', np.float64), ('move', np.int32), ('flags', np.int32), ('n_atoms', np.int32),
                             ('index', np.int32)])          # mcpyb_gcmc_trial

_p = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)

_lib = None


class EngineError(RuntimeError):
    """A call into libmcpy_b200 failed (CUDA error or internal capacity)."""


def load():
    """Load ``libmcpy_b200.so``; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing. The mcpy_b200 engine is CUDA-only: build it with '
            '`python -m mcpy_b200.build` (needs nvcc); there is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.mcpyb_create.argtypes = [C.c_int, C.POINTER(_p)]
    lib.mcpyb_destroy.argtypes = [_p]
    lib.mcpyb_destroy.restype = None
    lib.mcpyb_last_error.argtypes = [_p]
    lib.mcpyb_last_error.restype = C.c_char_p
    lib.mcpyb_version.restype = C.c_char_p
    lib.mcpyb_set_stream.argtypes = [_p, _p]
    lib.mcpyb_synchronize.argtypes = [_p]
    lib.mcpyb_set_lj.argtypes = [_p, C.c_double, C.c_double, C.c_double]
    lib.mcpyb_set_emt.argtypes = [_p]
    lib.mcpyb_cutoff.argtypes = [_p]
    lib.mcpyb_cutoff.restype = C.c_double
    lib.mcpyb_emt_species.argtypes = [_p, C.c_int, _f64p]
    lib.mcpyb_upload_host.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p]
    lib.mcpyb_upload_dev.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p]
    lib.mcpyb_copy_segments_dev.argtypes = [_p, C.c_int, _p, _p, _p, _p]
    lib.mcpyb_energies_host.argtypes = [_p, _p, _p]
    lib.mcpyb_energies_dev.argtypes = [_p, _p]
    lib.mcpyb_potential_energies_host.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p, _p]
    lib.mcpyb_tracked_energy_host.argtypes = [_p, _p, _p, C.c_int, _p, _p, _p, _p]
    lib.mcpyb_tracked_energy_z64_host.argtypes = [_p, _p, _p, C.c_int, _p, _p, _p, _p]
    lib.mcpyb_atom_energies_host.argtypes = [_p, _p, _p]
    lib.mcpyb_tracked_energies_host.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p, _p, _p]
    lib.mcpyb_tracked_energies_rows_host.argtypes = [_p, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p, _p]
    lib.mcpyb_forces_host.argtypes = [_p, _p]
    lib.mcpyb_energy_forces_host.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p, _p, _p]
    lib.mcpyb_relax_lbfgs_host.argtypes = [_p, _p, _p, C.c_int, _p, _p, _p, C.c_double, C.c_int, C.c_double,
                                           C.c_int, C.c_double, C.c_double, _p, _p, _p]
    lib.mcpyb_trial_delta_e_host.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p]
    lib.mcpyb_trial_delta_e_dev.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p]
    lib.mcpyb_neighbor_pairs_host.argtypes = [_p, C.c_int, C.c_double, C.c_int, C.c_int64,
                                              _p, _p, _p, _p, _i64p]
    lib.mcpyb_free_volume_count_host.argtypes = [_p, _p, C.c_int64, _p, _p, C.c_int64, _p, _p,
                                                 C.c_int, _p, _p]
    lib.mcpyb_free_volume_count_dev.argtypes = [_p, _p, C.c_int64, _p, _p, C.c_int64, _p, _p,
                                                C.c_int, _p]
    lib.mcpyb_sphere_free_volume_pcg64.argtypes = [_p, _p, _p, C.c_int64, C.c_double, _p, _p, _p,
                                                   C.c_int64, _p, _p]
    lib.mcpyb_dome_free_volume_pcg64.argtypes = [_p, _p, _p, C.c_int64, C.c_double, _p, C.c_double, _p, _p,
                                                 C.c_int64, _p, _p]
    lib.mcpyb_box_free_volume_pcg64.argtypes = [_p, _p, _p, C.c_int64, _p, _p, _p, C.c_int64, _p, _p,
                                                C.c_int, _p]
    lib.mcpyb_alloc_pinned.argtypes = [C.c_size_t]
    lib.mcpyb_alloc_pinned.restype = C.c_void_p
    lib.mcpyb_free_pinned.argtypes = [_p]
    lib.mcpyb_free_pinned.restype = None
    lib.mcpyb_min_dist2_host.argtypes = [_p, C.c_int, _p, C.c_int64, _p, _p]
    lib.mcpyb_pcg64_host_doubles.argtypes = [_p, _p, C.c_uint64, C.c_int64, _p]
    lib.mcpyb_fp64_peak_tflops.argtypes = [_p, _f64p]
    lib.mcpyb_stage_timing.argtypes = [_p, C.c_int, _p, _p]
    lib.mcpyb_launch_count.argtypes = [_p]
    lib.mcpyb_launch_count.restype = C.c_int64
    lib.mcpyb_resident_chain.argtypes = [_p, C.c_int, _i64p, _i64p]
    lib.mcpyb_set_precision.argtypes = [_p, C.c_int]
    lib.mcpyb_gcmc_create.argtypes = [_p, C.c_int, C.POINTER(GcmcDesc), C.c_int64, C.POINTER(_p)]
    lib.mcpyb_gcmc_destroy.argtypes = [_p]
    lib.mcpyb_gcmc_destroy.restype = None
    lib.mcpyb_gcmc_set_state.argtypes = [_p, C.c_int, _p, C.POINTER(GcmcScalars), _p]
    lib.mcpyb_gcmc_get_state.argtypes = [_p, C.c_int, _p, C.POINTER(GcmcScalars), _p, _p]
    lib.mcpyb_gcmc_run.argtypes = [_p, _p, _p]
    lib.mcpyb_gcmc_get_scalars.argtypes = [_p, C.c_int, C.POINTER(GcmcScalars)]
    lib.mcpyb_gcmc_swap_configurations.argtypes = [_p, C.c_int, C.c_int]
    lib.mcpyb_gcmc_cycles.argtypes = [_p, C.c_int, _i64p]
    lib.mcpyb_gcmc_config_ptr.argtypes = [_p, C.c_int, C.POINTER(_p), _i64p, _f64p]
    lib.mcpyb_gcmc_set_config_scalars.argtypes = [_p, C.c_int, _f64p]
    lib.mcpyb_gcmc_get_block_scalars.argtypes = [_p, _p, _p, _p, _p]
    lib.mcpyb_region_mask_host.argtypes = [_p, _p, _p, C.c_int, _p, C.c_double, C.c_int, C.c_double, _p, C.c_int, _p]
    lib.mcpyb_gcmc_get_trace.argtypes = [_p, C.c_int, _p, C.c_int64, _i64p]
    lib.mcpyb_gcmc_info.argtypes = [_p, C.c_int, _p]
    lib.mcpyb_gcmc_last_run_ms.argtypes = [_p, _f64p]
    _lib = lib
    return lib


def pinned_empty(shape, dtype=np.float64):
    """An uninitialised numpy array in page-locked host memory (``mcpyb_alloc_pinned``): the ``_host`` calls
    DMA such arrays where they are instead of staging them.  Freed when the array is collected."""
    import weakref
    lib = load()
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    addr = lib.mcpyb_alloc_pinned(max(nbytes, 1))
    if not addr:
        raise MemoryError(f'could not page-lock {nbytes} bytes of host memory')
    buf = (C.c_char * max(nbytes, 1)).from_address(addr)
    weakref.finalize(buf, lib.mcpyb_free_pinned, addr)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pinned_copy(a, dtype=None):
    """``a`` copied into a page-locked array of the same shape."""
    a = np.asarray(a)
    out = pinned_empty(a.shape, dtype or a.dtype)
    out[...] = a
    return out


def ptr(a):
    """Raw address of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)
