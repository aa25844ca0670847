"""Effective-medium-theory energy oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the published algorithm of ``ase.calculators.emt.EMT`` (ASE 3.23.0,
third-party dependency of the reference, ``uv.lock:9-20``; not vendored in
``/root/reference``).  Reference call sites: ``tests/test_canonical_nvt.py:26``,
``tests/test_audit_regressions.py:287``, ``docs/make_figures.py:214,228,312,346``,
``docs/examples/gcmc_simulations.rst:31,52``.

Pins (``tests/test_oracle_emt.py``): the numbers ASE's own documentation prints for its EMT
tutorials -- N2 on Cu(111) (``e_slab`` 11.509056, slab + N2 11.689927, fmax 1.0797, ``e_N2``
0.440344), H2O (2.769633, fmax 8.6091), Au on Al(100) (3.323870, fmax 0.2462), bulk Cu
(-0.00568151), the Ag equation of state (B = 100.14189241 GPa) -- all reproduced to the last
printed digit, energies AND forces, like and unlike pairs, mixed pbc; the reference's prose
value ``E_EMT(O2, d=1.245956 A)/2 = +0.46 eV`` (``docs/make_figures.py:334``); and for the
metals no recalled output covers (Ni, Pd, Pt) the identities of the published algorithm
(ideal fcc at beta*s0: sigma1 = 12, E = 0; isolated atom -E0; lattice constant and bulk
modulus of the table row).  ASE itself is not installable here, so these are the pins there are.
"""
from __future__ import annotations

from math import exp, log, sqrt

import numpy as np

from .neighbors import neighbor_pairs

BOHR = 0.52917721067           # ase.units.Bohr (CODATA 2014, ASE default)
BETA = 1.809                   # (16 pi / 3)^(1/3) / sqrt(2), historical rounding

#            E0     s0    V0     eta2    kappa   lambda  n0
#            eV     bohr  eV     bohr^-1 bohr^-1 bohr^-1 bohr^-3
PARAMETERS = {
    'Al': (-3.28, 3.00, 1.493, 1.240, 2.000, 1.169, 0.00700),
    'Cu': (-3.51, 2.67, 2.476, 1.652, 2.740, 1.906, 0.00910),
    'Ag': (-2.96, 3.01, 2.132, 1.652, 2.790, 1.892, 0.00547),
    'Au': (-3.80, 3.00, 2.321, 1.674, 2.873, 2.182, 0.00703),
    'Ni': (-4.44, 2.60, 3.673, 1.669, 2.757, 1.948, 0.01030),
    'Pd': (-3.90, 2.87, 2.773, 1.818, 3.107, 2.155, 0.00688),
    'Pt': (-5.85, 2.90, 4.067, 1.812, 3.145, 2.192, 0.00802),
    'H': (-3.21, 1.31, 0.132, 2.652, 2.790, 3.892, 0.00547),
    'C': (-3.50, 1.81, 0.332, 1.652, 2.790, 1.892, 0.01322),
    'N': (-5.10, 1.88, 0.132, 1.652, 2.790, 1.892, 0.01222),
    'O': (-4.60, 1.95, 0.332, 1.652, 2.790, 1.892, 0.00850),
}
ATOMIC_NUMBERS = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'Al': 13, 'Ni': 28, 'Cu': 29,
                  'Pd': 46, 'Ag': 47, 'Pt': 78, 'Au': 79}
SYMBOL_OF_Z = {z: s for s, z in ATOMIC_NUMBERS.items()}


def cutoffs():
    """``rc``, ``rc_list``, ``acut`` of EMT with ``asap_cutoff=False``: the
    cutoff is taken over ALL table species, so it does not depend on which
    elements are present."""
    maxseq = max(par[1] for par in PARAMETERS.values()) * BOHR
    rc = BETA * maxseq * 0.5 * (sqrt(3) + sqrt(4))
    rr = rc * 2 * sqrt(4) / (sqrt(3) + sqrt(4))
    acut = log(9999.0) / (rr - rc)
    rc_list = rc + 0.5
    return rc, rc_list, acut


def species_parameters(Z):
    """Derived per-species constants (``EMT.initialize``)."""
    rc, rc_list, acut = cutoffs()
    p = PARAMETERS[SYMBOL_OF_Z[int(Z)]]
    s0 = p[1] * BOHR
    eta2 = p[3] / BOHR
    kappa = p[4] / BOHR
    gamma1 = 0.0
    gamma2 = 0.0
    for i, n in enumerate([12, 6, 24]):
        r = s0 * BETA * sqrt(i + 1)
        x = n / (12 * (1.0 + exp(acut * (r - rc))))
        gamma1 += x * exp(-eta2 * (r - BETA * s0))
        gamma2 += x * exp(-kappa / BETA * (r - BETA * s0))
    return {'E0': p[0], 's0': s0, 'V0': p[2], 'eta2': eta2, 'kappa': kappa,
            'lambda': p[5] / BOHR, 'n0': p[6] / BOHR ** 3,
            'gamma1': gamma1, 'gamma2': gamma2}


def parameter_table():
    """Dense table for every EMT species, in ascending-Z order.

    Returns ``(Zs, table)`` with ``table[k] = (E0, s0, V0, eta2, kappa, lambda,
    n0, gamma1, gamma2)`` -- the layout the CUDA engine is handed by the host
    side, so tests can check that both sides derive identical constants."""
    Zs = sorted(SYMBOL_OF_Z)
    rows = []
    for Z in Zs:
        p = species_parameters(Z)
        rows.append([p['E0'], p['s0'], p['V0'], p['eta2'], p['kappa'],
                     p['lambda'], p['n0'], p['gamma1'], p['gamma2']])
    return np.array(Zs), np.array(rows)


def emt_energy(positions, numbers, cell, pbc, return_details=False):
    """Total EMT energy (pair pass over the half list with ``r < rc_list``,
    then the per-atom embedding pass), SURVEY.md 8a row a9."""
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    numbers = np.asarray(numbers).reshape(-1)
    n = len(pos)
    if n == 0:
        return (0.0, np.zeros(0), 0) if return_details else 0.0
    rc, rc_list, acut = cutoffs()
    par = {int(Z): species_parameters(Z) for Z in np.unique(numbers)}
    i, j, S, r2 = neighbor_pairs(pos, cell, pbc, rc_list, use_sqrt=True, half=True)
    sigma1 = np.zeros(n)
    energy = 0.0
    if len(i):
        r = np.sqrt(r2)
        P = lambda key, idx: np.array([par[int(z)][key] for z in numbers[idx]])
        V0a, V0b = P('V0', i), P('V0', j)
        ka, kb = P('kappa', i), P('kappa', j)
        s0a, s0b = P('s0', i), P('s0', j)
        e2a, e2b = P('eta2', i), P('eta2', j)
        g1a, g1b = P('gamma1', i), P('gamma1', j)
        g2a, g2b = P('gamma2', i), P('gamma2', j)
        ksi = P('n0', j) / P('n0', i)
        x = np.exp(acut * (r - rc))
        theta = 1.0 / (1.0 + x)
        y1 = 0.5 * V0a * np.exp(-kb * (r / BETA - s0b)) * ksi / g2a * theta
        y2 = 0.5 * V0b * np.exp(-ka * (r / BETA - s0a)) / ksi / g2b * theta
        energy -= float(np.sum(y1 + y2))
        np.add.at(sigma1, i, np.exp(-e2b * (r - BETA * s0b)) * ksi * theta / g1a)
        np.add.at(sigma1, j, np.exp(-e2a * (r - BETA * s0a)) / ksi * theta / g1b)
    for a in range(n):
        p = par[int(numbers[a])]
        if sigma1[a] <= 0.0:
            # ASE: log(0) raises -> "energy -= E0" (isolated atom)
            energy -= p['E0']
            continue
        ds = -log(sigma1[a] / 12) / (BETA * p['eta2'])
        x = p['lambda'] * ds
        y = exp(-x)
        z = 6 * p['V0'] * exp(-p['kappa'] * ds)
        energy += p['E0'] * ((1 + x) * y - 1) + z
    if return_details:
        return float(energy), sigma1, len(i)
    return float(energy)


def emt_forces(positions, numbers, cell, pbc):
    """Forces of ``ase.calculators.emt.EMT`` (ASE 3.23.0): ``interact1`` adds the pair-energy force
    ``((y1 kappa_b + y2 kappa_a) / beta + (y1 + y2) acut theta x) d / r`` to atom a and its negative
    to b; after the embedding pass has set ``deds[a] = dE_c/dsigma1`` (0 for an isolated atom),
    ``interact2`` subtracts ``((y1' eta2_b + y2' eta2_a) + (y1' + y2') acut theta x) d / r`` with
    ``y1' = dsigma1_a(pair) * deds[a]``, ``y2' = dsigma1_b(pair) * deds[b]``; ``d = pos_b + S @ cell - pos_a``
    over the half list.  Checked against finite differences of ``emt_energy`` in
    ``tests/test_oracle_forces.py`` (the reference's tests assert no EMT force value)."""
    from .neighbors import complete_cell
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    numbers = np.asarray(numbers).reshape(-1)
    n = len(pos)
    F = np.zeros((n, 3))
    if n == 0:
        return F
    rc, rc_list, acut = cutoffs()
    par = {int(Z): species_parameters(Z) for Z in np.unique(numbers)}
    cellc = complete_cell(np.array(cell, dtype=np.float64).reshape(3, 3), pbc)
    i, j, S, r2 = neighbor_pairs(pos, cell, pbc, rc_list, use_sqrt=True, half=True)
    if len(i) == 0:
        return F
    _, sigma1, _ = emt_energy(pos, numbers, cell, pbc, return_details=True)
    deds = np.zeros(n)
    for a in range(n):
        p = par[int(numbers[a])]
        if sigma1[a] <= 0.0:
            continue
        ds = -log(sigma1[a] / 12) / (BETA * p['eta2'])
        x = p['lambda'] * ds
        y = exp(-x)
        z = 6 * p['V0'] * exp(-p['kappa'] * ds)
        deds[a] = (x * y * p['E0'] * p['lambda'] + p['kappa'] * z) / (sigma1[a] * BETA * p['eta2'])
    r = np.sqrt(r2)
    d = (pos[j] + S @ cellc) - pos[i]
    P = lambda key, idx: np.array([par[int(z)][key] for z in numbers[idx]])
    V0a, V0b = P('V0', i), P('V0', j)
    ka, kb = P('kappa', i), P('kappa', j)
    s0a, s0b = P('s0', i), P('s0', j)
    e2a, e2b = P('eta2', i), P('eta2', j)
    g1a, g1b = P('gamma1', i), P('gamma1', j)
    g2a, g2b = P('gamma2', i), P('gamma2', j)
    ksi = P('n0', j) / P('n0', i)
    x = np.exp(acut * (r - rc))
    theta = 1.0 / (1.0 + x)
    y1 = 0.5 * V0a * np.exp(-kb * (r / BETA - s0b)) * ksi / g2a * theta
    y2 = 0.5 * V0b * np.exp(-ka * (r / BETA - s0a)) / ksi / g2b * theta
    f1 = (y1 * kb + y2 * ka) / BETA + (y1 + y2) * acut * theta * x
    z1 = np.exp(-e2b * (r - BETA * s0b)) * ksi / g1a * theta * deds[i]
    z2 = np.exp(-e2a * (r - BETA * s0a)) / ksi / g1b * theta * deds[j]
    f2 = (z1 * e2b + z2 * e2a) + (z1 + z2) * acut * theta * x
    f = ((f1 - f2) / r)[:, None] * d
    np.add.at(F, i, f)
    np.add.at(F, j, -f)
    return F


def emt_delta_e(pos_old, num_old, pos_new, num_new, cell, pbc):
    """``E(new) - E(old)`` (``grand_canonical_ensemble.py:283-284``)."""
    return emt_energy(pos_new, num_new, cell, pbc) - emt_energy(pos_old, num_old, cell, pbc)
