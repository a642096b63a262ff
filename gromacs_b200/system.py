"""Synthetic water boxes: the reference's own benchmark system, rebuilt without the reference.

BenchmarkSystem(k) (src/gromacs/nbnxm/benchmark/bench_system.cpp:95-217) stacks an equilibrated box of
1000 SPC/E waters (3000 atoms, 3.10736 nm) k times, doubling x, y, z cyclically.  data/spce1000.npz holds
that unit box (written by tests/golden/make_golden.py from the reference's coordinates).  The helpers
below restate the few interaction-constant formulas the benchmark configurations need."""
import math
import os
from dataclasses import dataclass

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spce1000.npz")

Q_O, Q_H = -0.8476, 0.4238                       # bench_system.cpp:78-80
C6_O, C12_O = 0.0026173456, 2.634129e-06         # bench_system.cpp:82-84
ONE_4PI_EPS0 = 138.935458                        # kJ mol^-1 nm e^-2 (math/units.h)


@dataclass
class WaterBox:
    x: np.ndarray            # [n,3] float32, inside the box
    q: np.ndarray            # [n]
    type: np.ndarray         # [n] 0 = O, 1 = H
    box: np.ndarray          # [3]
    excl_index: np.ndarray   # CSR exclusions (each water excludes its own 3 atoms)
    excl_atoms: np.ndarray
    factors: tuple

    @property
    def natoms(self):
        return self.x.shape[0]


def stacking_factors(k):
    """generateCoordinates, bench_system.cpp:95-125: k must be a power of two."""
    if k < 1 or (k & (k - 1)):
        raise ValueError("The size factor has to be a power of 2")
    f = [1, 1, 1]
    d = 0
    while k > 1:
        f[d] *= 2
        k //= 2
        d = (d + 1) % 3
    return tuple(f)


def benchmark_system(k):
    unit = np.load(_DATA)
    x0, box0 = unit["x"].astype(np.float32), unit["box"].astype(np.float32)
    fx, fy, fz = stacking_factors(k)
    n0 = x0.shape[0]
    x = np.empty((fx * fy * fz * n0, 3), np.float32)
    i = 0
    for ix in range(fx):
        for iy in range(fy):
            for iz in range(fz):
                shift = np.array([ix * box0[0], iy * box0[1], iz * box0[2]], np.float32)
                x[i:i + n0] = x0 + shift
                i += n0
    box = box0 * np.array([fx, fy, fz], np.float32)
    for d in range(3):  # put_atoms_in_box
        x[:, d] = np.where(x[:, d] >= box[d], x[:, d] - box[d], x[:, d])
        x[:, d] = np.where(x[:, d] < 0, x[:, d] + box[d], x[:, d])
    n = x.shape[0]
    t = np.ones(n, np.int32)
    t[0::3] = 0
    q = np.where(t == 0, Q_O, Q_H).astype(np.float32)
    excl_index = (np.arange(n + 1, dtype=np.int64) * 3).astype(np.int32)
    first = (np.arange(n, dtype=np.int32) // 3) * 3
    excl_atoms = (first[:, None] + np.arange(3, dtype=np.int32)[None, :]).reshape(-1).astype(np.int32)
    return WaterBox(x=x, q=q, type=t, box=box, excl_index=excl_index, excl_atoms=excl_atoms, factors=(fx, fy, fz))


def spce_nbfp():
    """nbfp table in the nbat convention (6*C6, 12*C12) for types O, H plus the all-zero filler type
    (atomdata.cpp:505-520, 579-588); hydrogens carry no LJ (bench_system.cpp:85)."""
    nt = 3
    nbfp = np.zeros((nt * nt, 2), np.float32)
    nbfp[0] = (6.0 * C6_O, 12.0 * C12_O)
    return nbfp, nt


def geometric_comb_params(nbfp, nt):
    """per-type sqrt(6*C6), sqrt(12*C12) (LJCombinationRule::Geometric, atomdata.cpp:326-400)."""
    c = np.zeros((nt, 2), np.float32)
    for t in range(nt):
        c[t] = (math.sqrt(nbfp[t * nt + t, 0]), math.sqrt(nbfp[t * nt + t, 1]))
    return c


def _bisect_coeff(f, rtol):
    beta, i = 5.0, 0
    while f(beta) > rtol:
        i += 1
        beta *= 2
    low, high = 0.0, beta
    for _ in range(i + 60):
        beta = (low + high) / 2
        if f(beta) > rtol:
            low = beta
        else:
            high = beta
    return beta


def ewald_beta(rc, rtol=1e-5):
    """calc_ewaldcoeff_q (src/gromacs/ewald/ewald_utils.cpp:57-85): erfc(beta*rc) = rtol."""
    return _bisect_coeff(lambda b: math.erfc(b * rc), rtol)


def ewald_beta_lj(rc, rtol=1e-3):
    """calc_ewaldcoeff_lj (ewald_utils.cpp:93-120): exp(-x^2)(1 + x^2 + x^4/2) = rtol with x = beta*rc."""
    def f(b):
        x2 = (b * rc) ** 2
        return math.exp(-x2) * (1 + x2 + x2 * x2 / 2)
    return _bisect_coeff(f, rtol)


def force_switch_constants(p, rsw, rc):
    """force_switch_constants (src/gromacs/mdtypes/interaction_const.cpp:176-190): (c2, c3, cpot)."""
    c2 = ((p + 1) * rsw - (p + 4) * rc) / (rc ** (p + 2) * (rc - rsw) ** 2)
    c3 = -((p + 1) * rsw - (p + 3) * rc) / (rc ** (p + 2) * (rc - rsw) ** 3)
    cpot = -rc ** (-p) + p * c2 / 3 * (rc - rsw) ** 3 + p * c3 / 4 * (rc - rsw) ** 4
    return c2, c3, cpot


def potential_switch_constants(rsw, rc):
    """potential_switch_constants (interaction_const.cpp:192-205): (c3, c4, c5)."""
    d = rc - rsw
    return -10.0 / d ** 3, 15.0 / d ** 4, -6.0 / d ** 5


def potential_shift_constants(rc):
    """PotShift: dispersion cpot = -rc^-6, repulsion cpot = -rc^-12 (interaction_const.cpp:255-262)."""
    return -rc ** -6.0, -rc ** -12.0
