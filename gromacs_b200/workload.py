"""Benchmark / test workloads: BASELINE.json's configurations assembled from the synthetic water box, the
host pair search and the interaction-constant formulas — everything the reference's
`gmx nonbonded-benchmark` set-up (src/gromacs/nbnxm/benchmark/bench_setup.cpp:178-294) and mdrun's
pair-list tuning would hand to the GPU path."""
import math
import time
from dataclasses import dataclass

import numpy as np

from . import system as S
from .interaction import make_params
from .nbnxm import PairlistGpu
from .pairsearch import Grid

# flops per computed pair: the reference's own table, src/gromacs/gmxlib/nrnb.cpp:92-112, as applied by
# accountFlops (src/gromacs/nbnxm/kerneldispatch.cpp:396-452); (F-only, F+E)
FLOPS_PER_PAIR = {"cut": (66, 107), "fswitch": (78, 129), "pswitch": (93, 127), "ljpme": (102, 140)}

# BASELINE.json configs -> (size factor k, rc, VdW flavor, energy every step, rlistOuter, rlistInner, nstlistPrune)
# The list radii and the pruning interval of the benchmark configurations are the REFERENCE's: its own pair-list tuning
# (increaseNstlist + setupDynamicPairlistPruning, nbnxm/pairlist_tuning.cpp:455-700) run on SPC/E water at 300 K, dt 2 fs,
# nstlist 100, default verlet-buffer-tolerance, for the GPU list - tests/golden/make_pairlist_tuning.sh asks a build of the
# reference and tests/golden/pairlist_tuning.json keeps its answer (tests/test_hostplan.py checks this table against it):
#   rc 1.0: outer 1.172, inner 1.003, pruned every 10 steps;  rc 1.2: outer 1.358, inner 1.201, every 10 steps.
# numRollingPruningParts = nstlistPrune / 2 (pairlist_tuning.cpp:685).  The *_test entries are small cases for the tests.
CONFIGS = {
    "bench3k": dict(k=1, rc=0.9, vdw="cut", energy=False, rlist_outer=0.9, rlist_inner=0.9, dynamic_pruning=False),
    "water48k_test": dict(k=16, rc=0.9, vdw="cut", energy=True, rlist_outer=0.95, rlist_inner=0.95, dynamic_pruning=False),
    "water24k_test": dict(k=8, rc=0.9, vdw="cut", energy=True, rlist_outer=0.95, rlist_inner=0.95, dynamic_pruning=False),
    "water384k_test": dict(k=128, rc=0.9, vdw="cut", energy=True, rlist_outer=0.95, rlist_inner=0.95, dynamic_pruning=False),
    "water96k_fswitch": dict(k=32, rc=1.0, vdw="fswitch", energy=True, rlist_outer=1.172, rlist_inner=1.003, nstlist_prune=10, dynamic_pruning=True),
    "water384k_ljpme": dict(k=128, rc=1.0, vdw="ljpme", energy=False, rlist_outer=1.172, rlist_inner=1.003, nstlist_prune=10, dynamic_pruning=True),
    "water384k_pswitch": dict(k=128, rc=1.0, vdw="pswitch", energy=False, rlist_outer=1.172, rlist_inner=1.003, nstlist_prune=10, dynamic_pruning=True),
    "water1536k": dict(k=512, rc=1.0, vdw="cut", energy=False, rlist_outer=1.172, rlist_inner=1.003, nstlist_prune=10, dynamic_pruning=True),
    "water12m": dict(k=4096, rc=1.2, vdw="cut", energy=False, rlist_outer=1.358, rlist_inner=1.201, nstlist_prune=10, dynamic_pruning=True),
}


@dataclass
class Workload:
    name: str
    cfg: dict
    box: S.WaterBox
    grid: Grid
    nbat: object
    params: object
    useful_pairs: float       # N/2 (rho 4/3 pi rc^3 + 1), bench_setup.cpp:332-337
    flops_per_pair: int
    grid_seconds: float = 0.0  # wall time of the host gridder (nbnxm_b200_grid_create)

    def pairlist(self, min_sci=0, **kw) -> PairlistGpu:
        return self.grid.pairlist(self.cfg["rlist_outer"], self.box.excl_index, self.box.excl_atoms, min_sci=min_sci, **kw)


def make_interaction_params(vdw, rc, rlist_outer, rlist_inner, dynamic_pruning, ewald_rtol=1e-5, ewald_rtol_lj=1e-3):
    """What init_interaction_const + initNbparam give for Ewald electrostatics and the chosen LJ treatment
    (src/gromacs/mdtypes/interaction_const.cpp:118-131, 151-166, 229-250)."""
    beta = S.ewald_beta(rc, ewald_rtol)
    kw = dict(epsfac=S.ONE_4PI_EPS0, rcoulomb=rc, rvdw=rc, rlist_outer=rlist_outer, rlist_inner=rlist_inner,
              ewald_beta=beta, sh_ewald=math.erfc(beta * rc) / rc, use_dynamic_pruning=dynamic_pruning)
    if vdw == "cut":
        cd, cr = S.potential_shift_constants(rc)
        return make_params("EwaldAna", "CutCombGeom", disp=(0.0, 0.0, cd), rep=(0.0, 0.0, cr), **kw)
    if vdw == "fswitch":
        rsw = rc - 0.2
        return make_params("EwaldAna", "FSwitch", rvdw_switch=rsw, disp=S.force_switch_constants(6.0, rsw, rc),
                           rep=S.force_switch_constants(12.0, rsw, rc), **kw)
    if vdw == "pswitch":
        rsw = rc - 0.2
        return make_params("EwaldAna", "PSwitch", rvdw_switch=rsw, sw=S.potential_switch_constants(rsw, rc), **kw)
    if vdw == "ljpme":
        cd, cr = S.potential_shift_constants(rc)
        blj = S.ewald_beta_lj(rc, ewald_rtol_lj)
        crc2 = (blj * rc) ** 2
        sh_lj = (math.exp(-crc2) * (1 + crc2 + 0.5 * crc2 * crc2) - 1) / rc ** 6
        return make_params("EwaldAna", "EwaldGeom", disp=(0.0, 0.0, cd), rep=(0.0, 0.0, cr), ewaldcoeff_lj=blj,
                           sh_lj_ewald=sh_lj, **kw)
    raise ValueError(vdw)


def rolling_prune_parts(cfg):
    """numRollingPruningParts = nstlistPrune / c_nbnxmGpuRollingListPruningInterval (pairlist_tuning.cpp:685)"""
    return max(1, cfg.get("nstlist_prune", 6) // 2)


def make_workload(name, nthreads=None, energy=None, nslabs=1, host_grid=True) -> Workload:
    """host_grid=False: no gridding on the host (grid is None, nbat carries the parameter tables only): for callers that put
    the atoms on the grid on the device (GpuPairSearch.put_atoms_on_grid)."""
    cfg = dict(CONFIGS[name])
    if energy is not None:
        cfg["energy"] = energy
    box = S.benchmark_system(cfg["k"])
    nbfp, nt = S.spce_nbfp()
    comb = S.geometric_comb_params(nbfp, nt)
    t_grid = time.perf_counter()
    if host_grid:
        grid = Grid(box.box, box.x, nthreads=nthreads, nslabs=nslabs)
        nbat = grid.atomdata(box.x, box.q, box.type, nbfp, nt, nbfp_comb=comb, lj_comb_per_type=comb)
    else:
        from .nbnxm import AtomData
        from .pairsearch import shift_vectors
        grid = None
        nbat = AtomData(xq=np.zeros((0, 4), np.float32), nbfp=nbfp, nbfp_comb=comb, numTypes=nt, shift_vec=shift_vectors(box.box))
    t_grid = time.perf_counter() - t_grid
    params = make_interaction_params(cfg["vdw"], cfg["rc"], cfg["rlist_outer"], cfg["rlist_inner"], cfg["dynamic_pruning"])
    n = box.natoms
    density = n / float(np.prod(box.box.astype(np.float64)))
    useful = n * 0.5 * (density * 4.0 / 3.0 * math.pi * cfg["rc"] ** 3 + 1.0)
    fl = FLOPS_PER_PAIR[cfg["vdw"]][1 if cfg["energy"] else 0]
    return Workload(name=name, cfg=cfg, box=box, grid=grid, nbat=nbat, params=params, useful_pairs=useful, flops_per_pair=fl,
                    grid_seconds=t_grid)
