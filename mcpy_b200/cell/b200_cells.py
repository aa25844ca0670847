"""``mcpy.cell`` regions whose Monte-Carlo free volume is counted on a B200.

``B200SphericalCell``, ``B200DomeCell`` and ``B200CustomCell`` ARE the reference's
``SphericalCell`` / ``DomeCell`` / ``CustomCell`` (``mcpy/cell/spherical_cell.py``,
``dome_cell.py``, ``custom_cell.py``): they subclass them and inherit the constructors, the
region predicates, ``get_random_point`` and the numpy PCG64 stream ``np.random.default_rng(seed)``
of ``cell.py:25`` unchanged, so a run seeded like the reference proposes the same points.
Three things are overridden:

* ``calculate_volume(atoms)`` -- the cKDTree loops of ``spherical_cell.py:127-140``,
  ``dome_cell.py:127-136`` and ``custom_cell.py:80-89`` become one integer count on the GPU
  (``csrc/free_volume.cuh``), the same integer.  ``points='host'`` draws the sample points with
  numpy exactly as the reference does and uploads them; ``points='device'`` regenerates the
  identical PCG64 stream on the GPU (``csrc/pcg64.cuh``) and advances the host generator by the
  same number of draws;
* ``_radii_for(atoms)`` (``cell.py:69-89``) and the species half of
  ``get_atoms_specie_inside_cell`` -- one table lookup by atomic number instead of a Python walk
  over the list of chemical symbols (milliseconds per call at 10 000 atoms, several times the GPU
  work it feeds).  Same values, same error for a species without a radius.
"""
from __future__ import annotations

import numpy as np

try:
    from mcpy.cell import CustomCell, DomeCell, SphericalCell
except ImportError as exc:                                   # pragma: no cover
    raise ImportError('mcpy_b200.cell extends the cells of farrisric/mcpy: the `mcpy` package (and its '
                      'dependency `ase`) must be importable') from exc

from ..engine import Engine


class _B200CellMixin:
    """Engine attachment and the by-atomic-number lookups shared by the three regions."""

    def _attach(self, engine, device, points):
        if points not in ('host', 'device'):
            raise ValueError("points must be 'host' or 'device'")
        self.engine = engine if engine is not None else Engine(device)
        self.points = points
        self.last_counts = None           # (free, covered) of the last device count

    def _radius_table(self):
        """radius by atomic number (NaN = no radius given), rebuilt when ``species_radii`` changes."""
        key = tuple(sorted(self.species_radii.items()))
        if getattr(self, '_radius_key', None) != key:
            from ase.data import atomic_numbers
            table = np.full(max(len(atomic_numbers), 120), np.nan)
            for sym, r in self.species_radii.items():
                table[atomic_numbers[sym]] = float(r)
            self._radius_key, self._radius_by_Z = key, table
        return self._radius_by_Z

    def _radii_for(self, atoms):
        radii = self._radius_table()[np.asarray(atoms.numbers)]
        if np.isnan(radii).any():
            return super()._radii_for(atoms)                 # raises the reference's ValueError
        return radii

    def _species_mask(self, atoms, specie):
        return np.isin(np.asarray(atoms.numbers), self._species_numbers(specie))

    @staticmethod
    def _species_numbers(specie):
        from ase.data import atomic_numbers
        names = [specie] if isinstance(specie, str) else list(specie)
        return [atomic_numbers[s] for s in names if s in atomic_numbers]

    # K10: from this many atoms on the `species AND inside` mask of the sphere / dome cells is taken on the device
    # (below it numpy's O(N) pass is cheaper than a launch and two copies); same bits either way
    mask_on_device_from = 4096

    def _inside_sphere(self, atoms, specie, bottom_z=None):
        if len(atoms) >= self.mask_on_device_from:
            return np.where(self.engine.region_mask(atoms.positions, atoms.numbers, self.center, self.radius,
                                                    self._species_numbers(specie), bottom_z=bottom_z))[0]
        d = atoms.positions - self.center
        mask = self._species_mask(atoms, specie) & (np.einsum('ij,ij->i', d, d) <= self.radius ** 2)
        if bottom_z is not None:
            mask &= atoms.positions[:, 2] >= bottom_z
        return np.where(mask)[0]


class B200SphericalCell(_B200CellMixin, SphericalCell):
    """``mcpy.cell.SphericalCell`` with the covered test on the GPU."""

    def __init__(self, atoms, vacuum, species_radii, mc_sample_points=100_000, seed=None,
                 engine=None, device=0, points='host'):
        SphericalCell.__init__(self, atoms, vacuum, species_radii, mc_sample_points=mc_sample_points, seed=seed)
        self._attach(engine, device, points)

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        return self._inside_sphere(atoms, specie)

    def calculate_volume(self, atoms):
        if len(atoms) == 0:
            self.volume = self.sphere_volume
            return
        radii = self._radii_for(atoms)
        if self.points == 'device':
            n_free, n_cov = self.engine.sphere_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.radius, self.center, atoms.positions, radii)
        else:
            pts = self._sample_sphere_points(self.mc_sample_points)
            n_free, n_cov = self.engine.free_volume_count(pts, atoms.positions, radii)
        self.last_counts = (n_free, n_cov)
        self.volume = self._clamp_free_volume(float(n_free) / self.mc_sample_points * self.sphere_volume,
                                              self.sphere_volume)


class B200DomeCell(_B200CellMixin, DomeCell):
    """``mcpy.cell.DomeCell`` with the covered test on the GPU (free fraction over the FULL-ball
    sample count, ``dome_cell.py:136``)."""

    def __init__(self, atoms, particle_species, bottom_z, vacuum, species_radii,
                 mc_sample_points: int = 100_000, seed=None, engine=None, device=0, points='host'):
        DomeCell.__init__(self, atoms, particle_species, bottom_z, vacuum, species_radii,
                          mc_sample_points=mc_sample_points, seed=seed)
        self._attach(engine, device, points)

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        return self._inside_sphere(atoms, specie, bottom_z=self.bottom_z)

    def calculate_volume(self, atoms):
        n_atoms = len(atoms)
        radii = self._radii_for(atoms) if n_atoms else np.empty(0)
        if self.points == 'device':
            n_free, n_cov = self.engine.dome_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.radius, self.center, self.bottom_z,
                atoms.positions.reshape(-1, 3), radii)
        else:
            pts = self._sample_sphere_points(self.mc_sample_points)
            pts = pts[pts[:, 2] >= self.bottom_z]
            n_free, n_cov = (self.engine.free_volume_count(pts, atoms.positions, radii) if n_atoms
                             else (len(pts), 0))
        self.last_counts = (n_free, n_cov)
        free = float(n_free) / self.mc_sample_points * self.sphere_volume
        self.volume = free if n_atoms == 0 else self._clamp_free_volume(free, self.sphere_volume)


class B200CustomCell(_B200CellMixin, CustomCell):
    """``mcpy.cell.CustomCell`` with the covered test on the GPU; the -1/0/+1 images along the
    periodic axes of the ORIGINAL box (``custom_cell.py:93-107``) are applied inside the kernel."""

    def __init__(self, atoms, custom_height=None, bottom_z=None, species_radii=None,
                 mc_sample_points: int = 100_000, seed=None, engine=None, device=0, points='host'):
        CustomCell.__init__(self, atoms, custom_height=custom_height, bottom_z=bottom_z,
                            species_radii=species_radii, mc_sample_points=mc_sample_points, seed=seed)
        self._attach(engine, device, points)

    def get_atoms_specie_inside_cell(self, atoms, specie):
        if len(atoms) == 0:
            return np.empty(0, dtype=int)
        f = (atoms.positions - self.offset) @ self._dim_inv
        inside = np.all((f[:, :2] >= 0.0) & (f[:, :2] < 1.0), axis=1) & (f[:, 2] >= 0.0)
        return np.where(self._species_mask(atoms, specie) & inside)[0]

    def _image_shifts(self, atoms):
        """Lattice shifts in the order ``_periodic_images`` concatenates its images."""
        pbc = np.asarray(atoms.pbc, dtype=bool)
        if not pbc.any():
            return np.zeros((1, 3))
        ranges = [(-1, 0, 1) if p else (0,) for p in pbc]
        return np.array([np.array([i, j, k], dtype=float) @ self.original_dimensions
                         for i in ranges[0] for j in ranges[1] for k in ranges[2]])

    def calculate_volume(self, atoms):
        if len(atoms) == 0:
            self.volume = self.cell_volume
            return
        radii = self._radii_for(atoms)
        shifts = self._image_shifts(atoms)
        if self.points == 'device':
            n_free, n_cov = self.engine.box_free_volume_pcg64(
                self._rng, self.mc_sample_points, self.dimensions, atoms.positions, radii, self.offset, shifts)
        else:
            cart = self._rng.random((self.mc_sample_points, 3)) @ self.dimensions
            n_free, n_cov = self.engine.free_volume_count(cart, atoms.positions, radii,
                                                          offset=self.offset, shifts=shifts)
        self.last_counts = (n_free, n_cov)
        self.volume = self._clamp_free_volume(
            self.cell_volume * (1.0 - float(n_cov) / self.mc_sample_points), self.cell_volume)
