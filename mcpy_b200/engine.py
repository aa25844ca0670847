"""Thin Python object around one ``mcpyb_engine`` handle (include/mcpy_b200.h).

Holds no arithmetic: every method marshals numpy buffers into the C ABI and
raises on a non-zero status.  The calculators and cells in this package are
written on top of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EngineError, as_f64, as_i32, as_i64, as_u8, ptr

try:                                      # host-side marshalling helper (csrc/marshal.cpp, built by mcpy_b200.build)
    from . import _marshal
except ImportError:                       # not built: the numpy marshalling below does the same work, slower
    _marshal = None

_F64, _I32, _I64 = np.dtype(np.float64), np.dtype(np.int32), np.dtype(np.int64)


def _address(a):
    """Address of a C-contiguous numpy array's data (None for an empty one)."""
    if a.size == 0:
        return None
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError):                            # read-only array: no writable buffer export
        return a.ctypes.data

_ERR_VALUE = (-2,)            # MCPYB_ERR_ARG -> ValueError, like the reference's bad-argument errors


class Engine:
    """One engine = one CUDA device, one stream, one potential, one resident batch."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.mcpyb_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise EngineError(
                f'mcpyb_create(device={device}) failed (status {rc}): no usable CUDA device. '
                'The mcpy_b200 engine has no CPU fallback.')
        self._h = h
        self.device = int(device)
        self.potential = None
        self._off1 = np.zeros(2, dtype=np.int64)      # scratch of the single-configuration fast path
        self._e1 = np.zeros(1, dtype=np.float64)
        self._path = C.c_int(0)
        self._e1_ptr = self._e1.ctypes.data
        self._path_ref = C.byref(self._path)
        self._tables = None                            # pointer / size tables of tracked_energies_atoms
        self._tracked_fns = None                       # (fn int32, fn int64, handle) addresses for _marshal.tracked_energy
        self._table_ptrs = None

    # -- plumbing ------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None):
            self._lib.mcpyb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._lib.mcpyb_last_error(self._h).decode()
        if rc in _ERR_VALUE:
            raise ValueError(msg)
        raise EngineError(f'{msg} (status {rc})')

    def set_stream(self, cuda_stream: int | None):
        """Enqueue on an existing CUDA stream, e.g. ``torch.cuda.current_stream().cuda_stream``."""
        self._check(self._lib.mcpyb_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self._lib.mcpyb_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.mcpyb_launch_count(self._h))

    def resident_chain(self, enable=None):
        """Switch the resident chain grid of the one-call-per-trial path (``mcpyb_resident_chain``) and return
        ``(grids launched, trials answered through the mailbox)``."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.mcpyb_resident_chain(self._h, -1 if enable is None else int(bool(enable)), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    STAGES = ('trial_sites', 'trial_scan', 'trial_scatter', 'row_sums', 'combine')

    def stage_timing(self, enable=True):
        """Fold the last trial launch into the per-stage running means; returns ``({stage: ms}, calls)``."""
        out = np.zeros(5)
        calls = C.c_int64(0)
        self._check(self._lib.mcpyb_stage_timing(self._h, 1 if enable else 0, out.ctypes.data, C.byref(calls)))
        return dict(zip(self.STAGES, out.tolist())), int(calls.value)

    def fp64_peak_tflops(self) -> float:
        """Measured DFMA throughput of this GPU (roofline denominator of the pair kernels)."""
        out = C.c_double(0.0)
        self._check(self._lib.mcpyb_fp64_peak_tflops(self._h, C.byref(out)))
        return float(out.value)

    @property
    def cutoff(self) -> float:
        return float(self._lib.mcpyb_cutoff(self._h))

    # -- potentials ----------------------------------------------------------
    def set_lj(self, epsilon=1.0, sigma=1.0, rc=None):
        self._check(self._lib.mcpyb_set_lj(self._h, float(epsilon), float(sigma),
                                           -1.0 if rc is None else float(rc)))
        self.potential = 'lj'

    def set_precision(self, precision='fp64'):
        """'fp64' (default) or 'fp32': fp32 pair math with fp64 accumulation in the batched LJ row sums
        (``mcpyb_set_precision``; 1e-5 relative instead of 1e-10)."""
        if precision not in ('fp64', 'fp32'):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        self._check(self._lib.mcpyb_set_precision(self._h, 1 if precision == 'fp32' else 0))
        self.precision = precision

    def set_emt(self):
        self._check(self._lib.mcpyb_set_emt(self._h))
        self.potential = 'emt'

    def emt_species(self, Z: int) -> np.ndarray:
        out = np.zeros(9)
        self._check(self._lib.mcpyb_emt_species(self._h, int(Z), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    # -- batches -------------------------------------------------------------
    @staticmethod
    def _batch_args(positions_list, numbers_list, cells, pbcs):
        R = len(positions_list)
        off = np.zeros(R + 1, dtype=np.int64)
        for r, p in enumerate(positions_list):
            off[r + 1] = off[r] + len(p)
        if R and off[-1]:
            # one concatenation and one dtype conversion for the whole batch
            pos = np.concatenate(positions_list).reshape(-1, 3)
            num = np.concatenate(numbers_list).reshape(-1)
        else:
            pos = np.zeros((0, 3))
            num = np.zeros(0, dtype=np.int32)
        cells = as_f64(cells, (R, 9)) if R else np.zeros((0, 9))
        pbcs = as_u8(np.asarray(pbcs, dtype=bool).reshape(R, 3)) if R else np.zeros((0, 3), np.uint8)
        return R, off, as_f64(pos), as_i32(num), cells, pbcs

    def upload(self, positions_list, numbers_list, cells, pbcs):
        """Make R configurations resident in HBM (host buffers -> cell lists)."""
        R, off, pos, num, cells, pbcs = self._batch_args(positions_list, numbers_list, cells, pbcs)
        self._keep = (off, pos, num, cells, pbcs)
        self._check(self._lib.mcpyb_upload_host(self._h, R, ptr(off), ptr(pos), ptr(num),
                                                ptr(cells), ptr(pbcs)))
        self._R = R
        self._ntot = int(off[-1])

    def upload_dev(self, atom_off, d_positions: int, d_numbers: int, cells, pbcs):
        """Same with device pointers (integers) for positions / numbers."""
        off = as_i64(atom_off)
        R = len(off) - 1
        cells = as_f64(cells, (R, 9))
        pbcs = as_u8(np.asarray(pbcs, dtype=bool).reshape(R, 3))
        self._keep = (off, cells, pbcs)
        self._check(self._lib.mcpyb_upload_dev(self._h, R, ptr(off), C.c_void_p(d_positions),
                                               C.c_void_p(d_numbers), ptr(cells), ptr(pbcs)))
        self._R = R
        self._ntot = int(off[-1])

    def copy_segments_dev(self, segments):
        """``segments``: iterable of ``(src address, dst address, count, kind)`` with kind 0 f64->f64, 1 i32->i32,
        2 f64->i32, 3 i32->f64; all of them in one launch on the engine's stream (``mcpyb_copy_segments_dev``)."""
        seg = np.array(segments, dtype=np.uint64).reshape(-1, 4)
        src, dst = np.ascontiguousarray(seg[:, 0]), np.ascontiguousarray(seg[:, 1])
        cnt, kind = seg[:, 2].astype(np.int64), seg[:, 3].astype(np.uint8)
        self._check(self._lib.mcpyb_copy_segments_dev(self._h, len(seg), ptr(src), ptr(dst), ptr(cnt), ptr(kind)))

    def energies(self, return_pairs=False):
        E = np.empty(self._R, dtype=np.float64)
        pairs = np.empty(self._R, dtype=np.int64) if return_pairs else None
        self._check(self._lib.mcpyb_energies_host(self._h, ptr(E), ptr(pairs)))
        return (E, pairs) if return_pairs else E

    def energies_dev(self, d_out: int):
        self._check(self._lib.mcpyb_energies_dev(self._h, C.c_void_p(d_out)))

    def potential_energies(self, positions_list, numbers_list, cells, pbcs) -> np.ndarray:
        """upload + energies in one C call (the whole plug-in call)."""
        R, off, pos, num, cells, pbcs = self._batch_args(positions_list, numbers_list, cells, pbcs)
        E = np.empty(R, dtype=np.float64)
        self._check(self._lib.mcpyb_potential_energies_host(
            self._h, R, ptr(off), ptr(pos), ptr(num), ptr(cells), ptr(pbcs), ptr(E)))
        self._R = R
        self._ntot = int(off[-1])
        return E

    def tracked_energies(self, positions_list, numbers_list, cells, pbcs):
        """Batch through the incremental path (resident bases + one batched delta launch); returns
        ``(energies, paths)`` with paths 0 = full evaluation, 1 = incremental, 2 = identical to its base."""
        R = len(positions_list)
        rows = self._row_pointers(positions_list, numbers_list) if R else None
        if rows is not None:
            # every configuration is already a C-contiguous float64 / int32-or-int64 array (what ase.Atoms
            # holds): the engine reads them where they are -- concatenating 64 x 2 000 atoms costs ~0.4 ms,
            # more than the rest of the call
            pos_ptrs, num_ptrs, is64, n_of = rows
            cells = as_f64(cells, (R, 9))
            pbcs = as_u8(np.asarray(pbcs, dtype=bool).reshape(R, 3))
            E = np.empty(R, dtype=np.float64)
            paths = np.zeros(R, dtype=np.int32)
            self._check(self._lib.mcpyb_tracked_energies_rows_host(
                self._h, R, ptr(pos_ptrs), ptr(num_ptrs), is64, ptr(n_of), ptr(cells), ptr(pbcs), ptr(E), ptr(paths)))
            self._R = R
            self._ntot = int(n_of.sum())
            return E, paths
        R, off, pos, num, cells, pbcs = self._batch_args(positions_list, numbers_list, cells, pbcs)
        E = np.empty(R, dtype=np.float64)
        paths = np.zeros(R, dtype=np.int32)
        self._check(self._lib.mcpyb_tracked_energies_host(
            self._h, R, ptr(off), ptr(pos), ptr(num), ptr(cells), ptr(pbcs), ptr(E), ptr(paths)))
        self._R = R
        self._ntot = int(off[-1])
        return E, paths

    def tracked_energies_atoms(self, atoms_list):
        """``tracked_energies`` of a list of ``ase.Atoms``-like configurations, the walk over the list done by the
        CPython helper (csrc/marshal.cpp: ~10 us for 64 configurations, against ~135 us from Python).  Returns None
        when some configuration cannot be read in place (the caller then marshals with numpy)."""
        if _marshal is None:
            return None
        R = len(atoms_list)
        tab = self._tables
        if tab is None or tab[0] != R:
            tab = self._tables = (R, np.zeros(R, np.uintp), np.zeros(R, np.uintp), np.zeros(R, np.int64),
                                  np.zeros((R, 9)), np.zeros((R, 3), np.uint8), np.empty(R), np.zeros(R, np.int32))
            self._table_ptrs = tuple(ptr(t) for t in tab[1:])
        _, pp, zp, n_of, cells, pbcs, E, paths = tab
        is64 = _marshal.rows(atoms_list, pp, zp, n_of, cells, pbcs)
        if is64 < 0:
            return None
        tp = self._table_ptrs
        rc = self._lib.mcpyb_tracked_energies_rows_host(self._h, R, tp[0], tp[1], is64, tp[2], tp[3], tp[4], tp[5], tp[6])
        if rc != 0:
            self._check(rc)
        self._R = R
        self._ntot = int(n_of.sum())
        return E.copy(), paths

    @staticmethod
    def _row_pointers(positions_list, numbers_list):
        """(position pointers, number pointers, numbers-are-int64 flag, rows per configuration) when every array can
        be read in place, else None (the caller concatenates)."""
        R = len(positions_list)
        pp, zp, n_of = [0] * R, [0] * R, [0] * R
        nd, f64, i32, i64 = np.ndarray, _F64, _I32, _I64
        addr, cbuf = C.addressof, C.c_char.from_buffer       # 0.45 us an address; .ctypes.data costs 1 us
        zdt = None
        for r in range(R):
            p, z = positions_list[r], numbers_list[r]
            if type(p) is not nd or type(z) is not nd or p.dtype is not f64 or p.ndim != 2:
                return None
            n = p.shape[0]
            if p.shape[1] != 3 or z.shape != (n,) or not p.flags.c_contiguous or not z.flags.c_contiguous:
                return None
            if zdt is None:
                zdt = z.dtype
                if zdt is not i32 and zdt is not i64:
                    return None
            elif z.dtype is not zdt:
                return None
            try:
                pp[r] = addr(cbuf(p)) if n else 0
                zp[r] = addr(cbuf(z)) if n else 0
            except (TypeError, ValueError):                    # read-only array: no writable buffer export
                pp[r] = p.__array_interface__['data'][0]
                zp[r] = z.__array_interface__['data'][0]
            n_of[r] = n
        pp = np.array(pp, dtype=np.uintp)
        zp = np.array(zp, dtype=np.uintp)
        n_of = np.array(n_of, dtype=np.int64)
        return pp, zp, int(zdt is _I64), n_of

    def potential_energy(self, positions, numbers, cell, pbc) -> float:
        """Single configuration, minimal marshalling (the per-trial drop-in call)."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        num = np.ascontiguousarray(numbers, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.float64)
        pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
        n = pos.shape[0]
        if pos.size != 3 * n or num.size != n or cell.size != 9 or pbc.size != 3:
            raise ValueError('positions (n,3), numbers (n,), cell (3,3), pbc (3,) expected')
        off = self._off1
        off[1] = n
        rc = self._lib.mcpyb_potential_energies_host(
            self._h, 1, off.ctypes.data, pos.ctypes.data, num.ctypes.data, cell.ctypes.data,
            pbc.ctypes.data, self._e1.ctypes.data)
        if rc != 0:
            self._check(rc)
        self._R = 1
        self._ntot = n
        return float(self._e1[0])

    def tracked_energy_atoms(self, atoms):
        """``tracked_energy`` of an ``ase.Atoms``-like object with the attribute walk and the C call made by the
        CPython helper (csrc/marshal.cpp); None when ``atoms`` cannot be read in place (the caller marshals with
        numpy)."""
        if _marshal is None:
            return None
        fns = self._tracked_fns
        if fns is None:
            cast = lambda f: C.cast(f, C.c_void_p).value                      # noqa: E731
            fns = self._tracked_fns = (cast(self._lib.mcpyb_tracked_energy_host),
                                       cast(self._lib.mcpyb_tracked_energy_z64_host), self._h.value)
        res = _marshal.tracked_energy(fns[0], fns[1], fns[2], atoms)
        if res is None:
            return None
        if res[0] != 0:
            self._check(res[0])
        self._R = 1
        self._ntot = res[3]
        return res[1], res[2]

    def tracked_energy(self, positions, numbers, cell, pbc):
        """Single configuration through the incremental path (base + one delta); returns
        ``(energy, path)`` with path 0 = full evaluation, 1 = incremental, 2 = identical to the base."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        num = np.ascontiguousarray(numbers, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.float64)
        pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
        n = pos.shape[0]
        if pos.size != 3 * n or num.size != n or cell.size != 9 or pbc.size != 3:
            raise ValueError('positions (n,3), numbers (n,), cell (3,3), pbc (3,) expected')
        # this call runs once per Monte-Carlo trial: addresses through the buffer protocol (0.45 us each, .ctypes.data
        # costs 1 us) and the two output pointers kept from construction
        rc = self._lib.mcpyb_tracked_energy_host(self._h, _address(pos), _address(num), n, _address(cell),
                                                 _address(pbc), self._e1_ptr, self._path_ref)
        if rc != 0:
            self._check(rc)
        self._R, self._ntot = 1, n
        return float(self._e1[0]), self._path.value

    def forces(self) -> np.ndarray:
        """Forces of the resident batch, ``(sum N, 3)`` (ASE's ``results['forces']``)."""
        f = np.empty((self._ntot, 3))
        self._check(self._lib.mcpyb_forces_host(self._h, ptr(f)))
        return f

    def energy_forces(self, positions, numbers, cell, pbc):
        """Single configuration: ``(energy, forces (n, 3))`` in one call."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        num = np.ascontiguousarray(numbers, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.float64)
        pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
        n = pos.shape[0]
        if pos.size != 3 * n or num.size != n or cell.size != 9 or pbc.size != 3:
            raise ValueError('positions (n,3), numbers (n,), cell (3,3), pbc (3,) expected')
        off = self._off1
        off[1] = n
        f = np.empty((n, 3))
        self._check(self._lib.mcpyb_energy_forces_host(
            self._h, 1, off.ctypes.data, pos.ctypes.data, num.ctypes.data, cell.ctypes.data,
            pbc.ctypes.data, self._e1.ctypes.data, f.ctypes.data))
        self._R, self._ntot = 1, n
        return float(self._e1[0]), f

    def relax_lbfgs(self, positions, numbers, cell, pbc, fmax=0.05, steps=100_000_000, fixed=None,
                    maxstep=0.2, memory=100, alpha=70.0, damping=1.0):
        """``ase.optimize.LBFGS(atoms).run(fmax, steps)`` followed by the energy, in one C call
        (``mcpyb_relax_lbfgs_host``).  Returns ``(energy, relaxed positions (n, 3), steps taken, final max |f|)``;
        ``positions`` itself is not modified."""
        pos = np.array(positions, dtype=np.float64, order='C', copy=True)
        num = np.ascontiguousarray(numbers, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.float64)
        pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
        n = pos.shape[0]
        if pos.size != 3 * n or num.size != n or cell.size != 9 or pbc.size != 3:
            raise ValueError('positions (n,3), numbers (n,), cell (3,3), pbc (3,) expected')
        fx = None
        if fixed is not None:
            fx = np.zeros(n, dtype=np.uint8)
            fx[np.asarray(fixed, dtype=int)] = 1
        nsteps, fm = C.c_int(0), C.c_double(0.0)
        self._check(self._lib.mcpyb_relax_lbfgs_host(
            self._h, pos.ctypes.data, num.ctypes.data, n, cell.ctypes.data, pbc.ctypes.data, ptr(fx),
            float(fmax), int(min(steps, 2**31 - 1)), float(maxstep), int(memory), float(alpha), float(damping),
            self._e1.ctypes.data, C.byref(nsteps), C.byref(fm)))
        self._R, self._ntot = 1, n
        return float(self._e1[0]), pos, int(nsteps.value), float(fm.value)

    def atom_energies(self, want_sigma1=False):
        e = np.empty(self._ntot)
        s = np.empty(self._ntot) if want_sigma1 else None
        self._check(self._lib.mcpyb_atom_energies_host(self._h, ptr(e), ptr(s)))
        return (e, s) if want_sigma1 else e

    # -- trial moves ---------------------------------------------------------
    def trial_delta_e(self, old_off, old_idx, new_off, new_pos, new_Z=None, rep=None,
                      return_pairs=False, out=None):
        """dE of T trial moves against the resident batch (``mcpyb_trial_delta_e_host``).  Arrays from
        ``pinned_empty`` / ``pinned_copy`` (and an ``out`` of that kind) are DMA'd where they are."""
        old_off = as_i32(old_off)
        new_off = as_i32(new_off)
        T = len(old_off) - 1
        if len(new_off) - 1 != T:
            raise ValueError('old_off and new_off must describe the same number of trials')
        old_idx = as_i32(old_idx)
        new_pos = as_f64(new_pos).reshape(-1, 3)
        new_Z = None if new_Z is None else as_i32(new_Z)
        rep = None if rep is None else as_i32(rep)
        if T and (old_off[-1] != len(old_idx) or new_off[-1] != len(new_pos)):
            raise ValueError('trial offsets do not match the index / position arrays')
        if out is not None and (out.dtype != np.float64 or out.shape != (T,) or not out.flags.c_contiguous):
            raise ValueError('out must be a C-contiguous float64 array of one value per trial')
        dE = np.empty(T, dtype=np.float64) if out is None else out
        pairs = np.empty(T, dtype=np.int32) if return_pairs else None
        self._check(self._lib.mcpyb_trial_delta_e_host(
            self._h, T, ptr(rep), ptr(old_off), ptr(old_idx), ptr(new_off), ptr(new_pos),
            ptr(new_Z), ptr(dE), ptr(pairs)))
        return (dE, pairs) if return_pairs else dE

    def trial_delta_e_dev(self, T, d_rep, d_old_off, d_old_idx, d_new_off, d_new_pos, d_new_Z,
                          n_old_total, n_new_total, d_dE, d_pairs=0):
        v = C.c_void_p
        self._check(self._lib.mcpyb_trial_delta_e_dev(
            self._h, int(T), v(d_rep or 0), v(d_old_off), v(d_old_idx or 0), v(d_new_off),
            v(d_new_pos or 0), v(d_new_Z or 0), int(n_old_total), int(n_new_total),
            v(d_dE), v(d_pairs or 0)))

    # -- neighbour pairs -----------------------------------------------------
    def neighbor_pairs(self, rep=0, cutoff=None, mode=0, max_pairs=None):
        """(i, j, S, r2) of every image within ``cutoff``, sorted like oracle.neighbors."""
        cutoff = self.cutoff if cutoff is None else float(cutoff)
        if max_pairs is None:
            max_pairs = 1 << 22
        while True:
            i = np.empty(max_pairs, np.int32)
            j = np.empty(max_pairs, np.int32)
            S = np.empty((max_pairs, 3), np.int32)
            r2 = np.empty(max_pairs, np.float64)
            n = C.c_int64(0)
            self._check(self._lib.mcpyb_neighbor_pairs_host(
                self._h, int(rep), cutoff, int(mode), int(max_pairs), ptr(i), ptr(j), ptr(S),
                ptr(r2), C.byref(n)))
            if n.value <= max_pairs:
                break
            max_pairs = int(n.value)
        k = n.value
        i, j, S, r2 = i[:k].astype(np.int64), j[:k].astype(np.int64), S[:k].astype(np.int64), r2[:k]
        order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
        return i[order], j[order], S[order], r2[order]

    # -- free volume ---------------------------------------------------------
    def free_volume_count(self, points, positions, radii, offset=None, shifts=None,
                          return_mask=False):
        points = as_f64(points).reshape(-1, 3)
        positions = as_f64(positions).reshape(-1, 3)
        radii = as_f64(radii).reshape(-1)
        if len(radii) != len(positions):
            raise ValueError('one radius per atom is required')
        offset = as_f64(np.zeros(3) if offset is None else offset).reshape(3)
        shifts = as_f64(np.zeros((1, 3)) if shifts is None else shifts).reshape(-1, 3)
        counts = np.zeros(2, dtype=np.int64)
        mask = np.empty(len(points), dtype=np.uint8) if return_mask else None
        self._check(self._lib.mcpyb_free_volume_count_host(
            self._h, ptr(points), len(points), ptr(positions), ptr(radii), len(positions),
            ptr(offset), ptr(shifts), len(shifts), ptr(counts), ptr(mask)))
        if return_mask:
            return int(counts[0]), int(counts[1]), mask.astype(bool)
        return int(counts[0]), int(counts[1])

    def free_volume_count_dev(self, d_points: int, n_points: int, d_positions: int, d_radii: int, n_atoms: int,
                              d_counts2: int, offset=None, shifts=None):
        """Covered test with everything resident in HBM (device pointers); ``d_counts2``: two uint64 (free, covered),
        zeroed by the caller.  No synchronisation."""
        offset = as_f64(np.zeros(3) if offset is None else offset).reshape(3)
        shifts = as_f64(np.zeros((1, 3)) if shifts is None else shifts).reshape(-1, 3)
        v = C.c_void_p
        self._check(self._lib.mcpyb_free_volume_count_dev(
            self._h, v(d_points), int(n_points), v(d_positions), v(d_radii), int(n_atoms), ptr(offset), ptr(shifts),
            len(shifts), v(d_counts2)))

    # -- free volume with device-regenerated PCG64 points -------------------------
    @staticmethod
    def _pcg_words(rng):
        """(state, inc) of a numpy ``Generator(PCG64)`` as {high, low} uint64 pairs."""
        st = rng.bit_generator.state
        if st.get('bit_generator') != 'PCG64':
            raise ValueError('the cell generator must be numpy PCG64 (np.random.default_rng)')
        s, i = int(st['state']['state']), int(st['state']['inc'])
        m = (1 << 64) - 1
        return (np.array([s >> 64, s & m], dtype=np.uint64), np.array([i >> 64, i & m], dtype=np.uint64))

    def sphere_free_volume_pcg64(self, rng, n_points, radius, center, positions, radii):
        """SphericalCell sampling + covered test on the device; advances ``rng`` exactly as
        ``_sample_sphere_points`` would (spherical_cell.py:88-106)."""
        state, inc = self._pcg_words(rng)
        positions = as_f64(positions).reshape(-1, 3)
        radii = as_f64(radii).reshape(-1)
        center = as_f64(center).reshape(3)
        counts = np.zeros(2, dtype=np.int64)
        draws = C.c_int64(0)
        self._check(self._lib.mcpyb_sphere_free_volume_pcg64(
            self._h, ptr(state), ptr(inc), int(n_points), float(radius), ptr(center), ptr(positions),
            ptr(radii), len(positions), ptr(counts), C.byref(draws)))
        rng.bit_generator.advance(int(draws.value))
        return int(counts[0]), int(counts[1])

    def dome_free_volume_pcg64(self, rng, n_points, radius, center, bottom_z, positions, radii):
        """DomeCell sampling + covered test on the device (dome_cell.py:106-138): the full-ball
        stream of ``_sample_sphere_points``, points below ``bottom_z`` ignored.  Returns
        (free, covered) among the points at or above ``bottom_z``; advances ``rng`` like the host."""
        state, inc = self._pcg_words(rng)
        positions = as_f64(positions).reshape(-1, 3)
        radii = as_f64(radii).reshape(-1)
        center = as_f64(center).reshape(3)
        counts = np.zeros(2, dtype=np.int64)
        draws = C.c_int64(0)
        self._check(self._lib.mcpyb_dome_free_volume_pcg64(
            self._h, ptr(state), ptr(inc), int(n_points), float(radius), ptr(center), float(bottom_z),
            ptr(positions), ptr(radii), len(positions), ptr(counts), C.byref(draws)))
        rng.bit_generator.advance(int(draws.value))
        return int(counts[0]), int(counts[1])

    def box_free_volume_pcg64(self, rng, n_points, dimensions, positions, radii, offset, shifts):
        """CustomCell sampling + covered test on the device; advances ``rng`` by 3 P draws."""
        state, inc = self._pcg_words(rng)
        positions = as_f64(positions).reshape(-1, 3)
        radii = as_f64(radii).reshape(-1)
        dims = as_f64(dimensions).reshape(9)
        offset = as_f64(offset).reshape(3)
        shifts = as_f64(shifts).reshape(-1, 3)
        counts = np.zeros(2, dtype=np.int64)
        self._check(self._lib.mcpyb_box_free_volume_pcg64(
            self._h, ptr(state), ptr(inc), int(n_points), ptr(dims), ptr(positions), ptr(radii),
            len(positions), ptr(offset), ptr(shifts), len(shifts), ptr(counts)))
        rng.bit_generator.advance(3 * int(n_points))
        return int(counts[0]), int(counts[1])

    # -- min_insert distance test ---------------------------------------------------
    def region_mask(self, positions, numbers, center, radius, species_Z, bottom_z=None) -> np.ndarray:
        """K10: bool mask ``numbers in species_Z and |positions - center|^2 <= radius ** 2`` (and ``z >= bottom_z``):
        the inside test of ``SphericalCell`` / ``DomeCell.get_atoms_specie_inside_cell`` with numpy's bits."""
        pos, num = as_f64(positions).reshape(-1, 3), as_i32(numbers)
        sp = as_i32(np.asarray(species_Z, dtype=np.int32).reshape(-1))
        c = as_f64(np.asarray(center, dtype=np.float64).reshape(3))
        out = np.empty(len(num), dtype=np.uint8)
        self._check(self._lib.mcpyb_region_mask_host(self._h, ptr(pos), ptr(num), len(num), ptr(c), float(radius) ** 2,
                                                     int(bottom_z is not None), float(bottom_z or 0.0), ptr(sp), len(sp),
                                                     ptr(out)))
        return out.view(np.bool_)

    def min_dist2(self, points, rep=0, mask=None) -> np.ndarray:
        """Smallest squared image distance from each point to the (masked) atoms of the
        resident replica ``rep`` within the cutoff; ``inf`` where no atom is that close."""
        points = as_f64(points).reshape(-1, 3)
        mask = None if mask is None else as_u8(np.asarray(mask, dtype=bool))
        out = np.empty(len(points), dtype=np.float64)
        self._check(self._lib.mcpyb_min_dist2_host(self._h, int(rep), ptr(points), len(points),
                                                   ptr(mask), ptr(out)))
        return out
