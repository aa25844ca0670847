"""Lennard-Jones energy oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Three restatements of the arithmetic mcpy reaches through
``BaseEnsemble.compute_energy`` (``mcpy/ensembles/base_ensemble.py:190-191``)
when the calculator wraps ``ase.calculators.lj.LennardJones`` (ASE 3.23.0,
third-party, not vendored; call sites ``examples/gcmc_molecule.py:27-31``,
``benchmark/lammps_gcmc_parity.py:197,294``):

* :func:`lj_energy` -- the definition: ASE's per-atom half-sum over a
  both-ways neighbour list, shifted at ``rc`` (``smooth=False``).
* :class:`LJCalcRestated` -- the reference's own numpy shortcut for cubic
  boxes (``benchmark/lammps_gcmc_parity.py:168-187``), which the reference
  asserts equal to ASE within 1e-9 relative (``:190-203``).
* :class:`AseStyleLJ` -- a port of the *cost structure* of the ASE path
  (kd-tree neighbour-list build, then an interpreted loop over atoms with
  numpy on each atom's neighbours); this is what ``bench.py`` times as the CPU
  baseline, because it is what one mcpy trial move costs today
  (``mcpy/ensembles/grand_canonical_ensemble.py:283-284``: every trial is one
  full ``get_potential_energy``).

The trial-move energy change itself has no separate formula in the reference:
``delta_E = E_new - self.E_old`` (``grand_canonical_ensemble.py:284``).
:func:`lj_delta_e` restates exactly that.
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy.spatial import cKDTree

from .neighbors import neighbor_pairs, complete_cell


def lj_e0(sigma, epsilon, rc):
    """Pair energy at the cutoff, ``e0 = 4 eps ((sigma/rc)^12 - (sigma/rc)^6)``
    (ASE ``lj.py``: "potential value at rc")."""
    return 4 * epsilon * ((sigma / rc) ** 12 - (sigma / rc) ** 6)


def lj_pair_energy(r2, sigma, epsilon, rc):
    """ASE's pairwise energy for squared distances ``r2`` (all within rc):
    ``c6 = (sigma^2/r2)^3``; ``c12 = c6^2``; ``4 eps (c12 - c6) - e0``."""
    c6 = (sigma ** 2 / r2) ** 3
    c12 = c6 ** 2
    return 4 * epsilon * (c12 - c6) - lj_e0(sigma, epsilon, rc)


def lj_energy(positions, cell, pbc, sigma=1.0, epsilon=1.0, rc=None,
              return_pairs=False):
    """Total LJ energy the way ASE LennardJones(smooth=False) defines it.

    ``energies[i] = 0.5 * sum_j e_pair(i, j)`` over the both-ways list, pairs
    kept unless ``r2 > rc**2``; ``E = energies.sum()``.
    """
    if rc is None:
        rc = 3 * sigma
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    if n == 0:
        return (0.0, 0) if return_pairs else 0.0
    i, j, S, r2 = neighbor_pairs(pos, cell, pbc, rc, inclusive=True)
    e_pair = lj_pair_energy(r2, sigma, epsilon, rc) if len(r2) else np.zeros(0)
    energies = 0.5 * np.bincount(i, weights=e_pair, minlength=n)
    E = float(energies.sum())
    if return_pairs:
        return E, len(i) // 2
    return E


def lj_forces(positions, cell, pbc, sigma=1.0, epsilon=1.0, rc=None):
    """Forces of ASE LennardJones(smooth=False) (``ase/calculators/lj.py``, ASE 3.23.0):
    ``pairwise_forces = (-24 eps (2 c12 - c6) / r2)[:, None] * distance_vectors`` with
    ``distance_vectors = pos[j] + S @ cell - pos[i]`` summed over the both-ways list of atom i
    (``c6`` zeroed where ``r2 > rc**2``).  Needed by the relax-then-energy contract of
    ``mcpy/calculators/base_calculator.py:14-22`` and ``CanonicalEnsemble.relax``
    (``canonical_ensemble.py:115-123``)."""
    if rc is None:
        rc = 3 * sigma
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    F = np.zeros((n, 3))
    if n == 0:
        return F
    cell = np.array(cell, dtype=np.float64).reshape(3, 3)
    i, j, S, r2 = neighbor_pairs(pos, cell, pbc, rc, inclusive=True)
    if len(i) == 0:
        return F
    d = (pos[j] + S @ complete_cell(cell, pbc)) - pos[i]
    c6 = (sigma ** 2 / r2) ** 3
    c12 = c6 ** 2
    w = -24 * epsilon * (2 * c12 - c6) / r2
    np.add.at(F, i, w[:, None] * d)
    return F


def lj_delta_e(pos_old, pos_new, cell, pbc, sigma=1.0, epsilon=1.0, rc=None):
    """``E(new) - E(old)`` by two full evaluations
    (``grand_canonical_ensemble.py:283-284``)."""
    return (lj_energy(pos_new, cell, pbc, sigma, epsilon, rc)
            - lj_energy(pos_old, cell, pbc, sigma, epsilon, rc))


class LJCalcRestated:
    """Restatement of ``LJCalc`` (``benchmark/lammps_gcmc_parity.py:168-187``):
    sigma = eps = 1, cubic box of side ``L``, minimum image, ``r2 < rc^2``,
    upper triangle, ``sum(4 (inv6^2 - inv6) - E_SHIFT)``."""

    def __init__(self, L, rc=3.0):
        self.L = float(L)
        self.rc = float(rc)
        self.e_shift = 4.0 * (self.rc ** -12 - self.rc ** -6)

    def energy(self, positions):
        pos = np.asarray(positions, dtype=np.float64)
        n = len(pos)
        if n < 2:
            return 0.0
        d = pos[:, None, :] - pos[None, :, :]
        d -= self.L * np.round(d / self.L)
        r2 = np.einsum('ijk,ijk->ij', d, d)
        r2 = r2[np.triu_indices(n, k=1)]
        r2 = r2[r2 < self.rc * self.rc]
        inv6 = 1.0 / r2 ** 3
        return float(np.sum(4.0 * (inv6 * inv6 - inv6) - self.e_shift))

    def get_potential_energy(self, atoms):
        return self.energy(atoms.positions)


class AseStyleLJ:
    """Port of the cost structure of ASE 3.23.0 ``LennardJones.calculate``.

    ``NeighborList([rc/2]*N, skin=0.3, self_interaction=False, bothways=True)``
    is rebuilt from scratch (ASE does so whenever ``'numbers'`` is in
    ``system_changes``, i.e. on every insertion / deletion trial, and whenever
    an atom moved further than the skin): positions wrapped, one
    ``cKDTree.query_ball_point`` per periodic image in the half space, an
    interpreted loop over atoms inside each image, then the both-ways
    expansion; afterwards the interpreted ``for ii in range(natoms)`` energy
    loop with numpy on each neighbour row.  Used only as the timed CPU
    baseline and cross-checked against :func:`lj_energy` in the tests.
    """

    skin = 0.3

    def __init__(self, sigma=1.0, epsilon=1.0, rc=None):
        self.sigma = float(sigma)
        self.epsilon = float(epsilon)
        self.rc = 3 * self.sigma if rc is None else float(rc)

    # -- neighbour list ----------------------------------------------------
    def build(self, positions, cell, pbc):
        pos = np.asarray(positions, dtype=np.float64)
        pbc = np.asarray(pbc, dtype=bool).reshape(3)
        cell = np.array(cell, dtype=np.float64).reshape(3, 3)
        natoms = len(pos)
        cut = self.rc / 2 + self.skin
        rcmax = cut
        cc = complete_cell(cell, pbc)
        frac = np.linalg.solve(cc.T, pos.T).T
        wrap = np.zeros_like(frac)
        wrap[:, pbc] = np.floor(frac[:, pbc])
        positions0 = (frac - wrap) @ cc
        offsets = -wrap.astype(int)          # positions0 = pos + offsets @ cell
        icell = np.linalg.pinv(cc)
        N = []
        for i in range(3):
            if pbc[i]:
                v = icell[:, i]
                h = 1 / np.sqrt(np.dot(v, v))
                N.append(int(2 * rcmax / h) + 1)
            else:
                N.append(0)
        tree = cKDTree(positions0, copy_data=True)
        neighbors = [np.empty(0, int) for _ in range(natoms)]
        displacements = [np.empty((0, 3), int) for _ in range(natoms)]
        for n1, n2, n3 in itertools.product(range(0, N[0] + 1),
                                            range(-N[1], N[1] + 1),
                                            range(-N[2], N[2] + 1)):
            if n1 == 0 and (n2 < 0 or n2 == 0 and n3 < 0):
                continue
            shift0 = np.array((n1, n2, n3))
            displacement = shift0 @ cc
            indices_all = tree.query_ball_point(positions0 - displacement,
                                                r=2 * rcmax)
            for a in range(natoms):
                indices = indices_all[a]
                if not len(indices):
                    continue
                indices = np.array(indices)
                delta = positions0[indices] + displacement - positions0[a]
                i = indices[np.linalg.norm(delta, axis=1) < 2 * cut]
                if n1 == 0 and n2 == 0 and n3 == 0:
                    i = i[i > a]
                neighbors[a] = np.concatenate((neighbors[a], i))
                disp = shift0 + offsets[i] - offsets[a]
                displacements[a] = np.concatenate((displacements[a], disp))
        # bothways=True expansion
        nb2 = [[] for _ in range(natoms)]
        dp2 = [[] for _ in range(natoms)]
        for a in range(natoms):
            for b, disp in zip(neighbors[a], displacements[a]):
                nb2[b].append(a)
                dp2[b].append(-disp)
        for a in range(natoms):
            nbrs = np.concatenate((neighbors[a], nb2[a])).astype(int)
            disp = np.array(list(displacements[a]) + dp2[a], dtype=int).reshape(-1, 3)
            neighbors[a] = nbrs
            displacements[a] = disp
        self._nl = (neighbors, displacements)
        return self._nl

    # -- energy loop -------------------------------------------------------
    def energy(self, positions, cell, pbc, rebuild=True):
        pos = np.asarray(positions, dtype=np.float64)
        cell = np.array(cell, dtype=np.float64).reshape(3, 3)
        natoms = len(pos)
        if natoms == 0:
            return 0.0
        if rebuild or getattr(self, '_nl', None) is None:
            self.build(pos, cell, pbc)
        neighbors_all, offsets_all = self._nl
        sigma, epsilon, rc = self.sigma, self.epsilon, self.rc
        e0 = 4 * epsilon * ((sigma / rc) ** 12 - (sigma / rc) ** 6)
        energies = np.zeros(natoms)
        for ii in range(natoms):
            neighbors, offsets = neighbors_all[ii], offsets_all[ii]
            cells = np.dot(offsets, cell)
            distance_vectors = pos[neighbors] + cells - pos[ii]
            r2 = (distance_vectors ** 2).sum(1)
            c6 = (sigma ** 2 / r2) ** 3
            c6[r2 > rc ** 2] = 0.0
            c12 = c6 ** 2
            pairwise_energies = 4 * epsilon * (c12 - c6)
            pairwise_energies -= e0 * (c6 != 0.0)
            energies[ii] += 0.5 * pairwise_energies.sum()
        return float(energies.sum())

    def get_potential_energy(self, atoms):
        return self.energy(atoms.positions, np.asarray(atoms.cell), atoms.pbc)
