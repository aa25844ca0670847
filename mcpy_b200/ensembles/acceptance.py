"""Scalar acceptance arithmetic of the GCMC loop, restated bit for bit.

Mirrors ``SetUnits`` (``mcpy/utils/set_unit_constant.py:55-96``) and
``GrandCanonicalEnsemble._acceptance_condition``
(``mcpy/ensembles/grand_canonical_ensemble.py:197-242``).  The reference keeps this
on the host and so does this package: the engine only delivers energies, and the
accept / reject sequence must match the reference exactly when it is fed the same
proposals and the same uniform draws -- including WHEN a draw is consumed (only on
the uphill branch for dN = 0, only when p <= 1 for insertions / deletions).
"""
from __future__ import annotations

import numpy as np

# constants hard-coded by the reference (set_unit_constant.py:66-69)
_PLANCK = 4.13567e-15           # eV s
_BOLTZMANN = 8.617333e-5        # eV / K
_AMU_KG = 1.66053906660e-27
_LAMBDA_CONV = np.sqrt(1.60218e-19) * 1e10


class Units:
    """beta and thermal de Broglie wavelengths per species ('LJ' or 'metal' units)."""

    def __init__(self, unit_type, temperature, masses):
        """``masses``: dict species -> mass in amu (ignored for 'LJ')."""
        self.unit_type = unit_type
        self.temperature = temperature
        if unit_type == 'LJ':
            self.beta = 1 / (temperature * 1.0)
            self.lambda_dbs = {s: 1 for s in masses}
        elif unit_type == 'metal':
            self.beta = 1 / (temperature * _BOLTZMANN)
            self.lambda_dbs = {
                s: (_PLANCK / np.sqrt(2 * np.pi * m * _AMU_KG * (1 / self.beta))) * _LAMBDA_CONV
                for s, m in masses.items()}
        else:
            raise ValueError("Invalid unit type. Choose 'LJ' or 'metal'.")

    def de_broglie_insertion(self, volume, n_atoms, specie):
        if n_atoms < 0:
            raise ValueError(f"n_atoms={n_atoms} < 0 for insertion of '{specie}'")
        return volume / ((n_atoms + 1) * self.lambda_dbs[specie] ** 3)

    def de_broglie_deletion(self, volume, n_atoms, specie):
        return self.lambda_dbs[specie] ** 3 * n_atoms / volume


def acceptance_probability(units, mu, potential_diff, delta_particles, volume, species, n_atoms):
    """The ``p`` of the reference rule; ``None`` for a downhill dN = 0 move (always accepted,
    no probability formed)."""
    if delta_particles == 0:
        if potential_diff <= 0:
            return None
        return np.exp(-potential_diff * units.beta)
    if delta_particles == 1:
        return (units.de_broglie_insertion(volume, n_atoms, species)
                * np.exp(-units.beta * (potential_diff - mu[species])))
    if delta_particles == -1:
        return (units.de_broglie_deletion(volume, n_atoms, species)
                * np.exp(-units.beta * (potential_diff + mu[species])))
    raise ValueError(f'unexpected delta_particles={delta_particles}')


def accept_trial(units, mu, potential_diff, delta_particles, volume, species, n_atoms, draw):
    """Accept / reject exactly like ``_acceptance_condition``; ``draw()`` is called only where
    the reference consumes a uniform."""
    p = acceptance_probability(units, mu, potential_diff, delta_particles, volume, species, n_atoms)
    if p is None:
        return True
    if delta_particles != 0 and p > 1:
        return True
    return bool(p > draw())


def swap_exponent(rec_i, rec_j, n_species):
    """ln(w'/w) of exchanging the configurations of slots i and j.

    ``rec = [E, beta, N_1..N_S, mu_1..mu_S, Lambda_1..Lambda_S, has_1..has_S]`` (``has_s`` = slot
    has a chemical potential for species s; a record without the ``has`` block means "all"; a
    ``Lambda`` the replica's units do not define travels as NaN).
    Same operation order as ``BatchedReplicaExchange._accept_swap`` / ``_de_broglie_cross_term``
    (``batched_replica_exchange.py:339-353, 372-386``):
    ``beta_i Phi_ii + beta_j Phi_jj - beta_i Phi_ji - beta_j Phi_ij`` with
    ``Phi_Z^(k) = E_Z - sum_{s in mu_k} mu_k,s N_Z,s`` accumulated species by species, plus
    ``3 sum_{s in mu_i} (N_X,s - N_Y,s) ln(Lambda_i,s / Lambda_j,s)`` when the betas differ.
    """
    S = n_species
    Ei, bi, Ni, mui, Li = rec_i[0], rec_i[1], rec_i[2:2 + S], rec_i[2 + S:2 + 2 * S], rec_i[2 + 2 * S:2 + 3 * S]
    Ej, bj, Nj, muj, Lj = rec_j[0], rec_j[1], rec_j[2:2 + S], rec_j[2 + S:2 + 2 * S], rec_j[2 + 2 * S:2 + 3 * S]
    has_i = rec_i[2 + 3 * S:2 + 4 * S] if len(rec_i) >= 2 + 4 * S else np.ones(S)
    has_j = rec_j[2 + 3 * S:2 + 4 * S] if len(rec_j) >= 2 + 4 * S else np.ones(S)

    def phi(E, N, mu, has):
        score = E
        for s in range(S):
            if has[s]:
                score -= mu[s] * N[s]
        return score

    phi_ii, phi_jj = phi(Ei, Ni, mui, has_i), phi(Ej, Nj, muj, has_j)
    phi_ji, phi_ij = phi(Ej, Nj, mui, has_i), phi(Ei, Ni, muj, has_j)
    delta = bi * phi_ii + bj * phi_jj - bi * phi_ji - bj * phi_ij
    if bi != bj:
        term = 0.0
        for s in range(S):
            if not has_i[s]:
                continue
            dn = int(Ni[s]) - int(Nj[s])
            if dn:
                if np.isnan(Li[s]) or np.isnan(Lj[s]):      # the reference's lambda_dbs[specie] lookup fails here
                    raise KeyError(f'a replica has no de Broglie wavelength for species #{s} of the ladder')
                term += 3.0 * dn * np.log(Li[s] / Lj[s])
        delta += term
    return delta
