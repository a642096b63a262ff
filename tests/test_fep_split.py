"""CPU test of the perturbed-pair split of the host builder (nbnxm_b200_pairlist_split_fep, make_fep_list of the
reference): with the perturbed atoms masked in the cluster list's atom data, cluster-list forces (oracle) + perturbed
atom-pair-list forces (FEP oracle) must give the brute-force result of the A-state system at lambda = 0 and of the
B-state system at lambda = 1 — every pair with a perturbed atom moved exactly once, exclusions and self pairs included."""
import numpy as np
import pytest

from util import load_golden, oracle_params, relrms


@pytest.mark.parametrize("case", ["bench1_ewald_cutnone", "bench1_rf_cutnone_split"])
def test_cluster_list_plus_fep_list_equals_brute_force_at_both_end_states(oracle, case):
    from gromacs_b200.pairsearch import Grid, split_fep_pairlist
    from oracle.nbfe_oracle import FepParams, nbfe_forces
    d = load_golden(case)
    x, box = d["sys_x"], d["sys_box"]
    n = x.shape[0]
    nt = int(d["nbat_ntypes"][0])
    q_a, t_a = d["sys_q"].astype(np.float64), d["sys_type"].astype(np.int32)
    rng = np.random.default_rng(4)
    mols = rng.choice(n // 3, size=45, replace=False)
    perturbed = np.zeros(n, np.uint8)
    q_b, t_b = q_a.copy(), t_a.copy()
    for k, m in enumerate(mols):
        atoms = np.arange(3 * m, 3 * m + 3)
        perturbed[atoms] = 1
        q_b[atoms] *= (0.0 if k % 3 == 0 else 0.5)            # decharged or half-charged in state B
        if k % 2 == 0:
            t_b[atoms] = nt - 1                               # and without LJ in state B
    rlist = 1.0
    grid = Grid(box, x, nthreads=2)
    ai = grid.atom_index
    real = ai >= 0
    # cluster list: perturbed atoms masked (nbnxm_atomdata_mask_fep)
    q_m, t_m = np.where(perturbed, 0.0, q_a), np.where(perturbed, nt - 1, t_a)
    nbat = grid.atomdata(x, q_m, t_m, d["nbat_nbfp"], nt, nbfp_comb=d["nbat_nbfp_comb"])
    grid.pairlist(rlist, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=200)
    plist, fep = split_fep_pairlist(grid, perturbed)
    assert fep["iinr"].shape[0] > 0 and (fep["excl_fep"] == 0).sum() > 0
    p = oracle_params(oracle, d)
    f_cl, _, e_cl, _ = oracle.forces(p, plist.sci, plist.cjPacked, plist.excl, nbat.xq, nbat.type,
                                     np.zeros((nbat.numAtoms(), 2), np.float32), nbat.nbfp, nbat.nbfp_comb, nbat.shift_vec)
    # perturbed atom-pair list, nbat order
    def nbat_order(a, fill):
        out = np.full(ai.shape[0], fill, np.asarray(a).dtype)
        out[real] = np.asarray(a)[ai[real]]
        return out
    g = lambda k: float(d[k][0])
    elec = "rf" if "rf" in case else "ewald"
    fp = FepParams(nbfp=d["nbat_nbfp"].reshape(-1, 2).astype(np.float64), ntypes=nt, elec=elec, vdw="cut", epsfac=g("ic_epsfac"),
                   c_rf=g("ic_c_rf"), two_k_rf=2.0 * g("ic_k_rf"), ewald_beta=g("ic_ewald_beta"), sh_ewald=g("ic_sh_ewald"),
                   rcoulomb_sq=g("ic_rcoulomb") ** 2, rvdw_sq=g("ic_rvdw") ** 2, disp=(0.0, 0.0, g("ic_disp_cpot")),
                   rep=(0.0, 0.0, g("ic_rep_cpot")))
    xn = nbat.xq[:, :3].astype(np.float64)
    zeros2 = np.zeros((ai.shape[0], 2))
    for lam, q_end, t_end in ((0.0, q_a, t_a), (1.0, q_b, t_b)):
        fp.lambda_coul = fp.lambda_vdw = lam
        f_fep, _, e_lj, e_el, _, _ = nbfe_forces(fp, xn, nbat_order(q_a, 0.0), nbat_order(q_b, 0.0), nbat_order(t_a, nt - 1),
                                                 nbat_order(t_b, nt - 1), zeros2, zeros2, nbat.shift_vec, **fep)
        f = oracle.nbat_to_atom_order(f_cl + f_fep, ai, n)
        fb, eb = oracle.brute_force(p, x, q_end.astype(np.float32), t_end, d["nbat_nbfp"], d["nbat_nbfp_comb"], box,
                                    d["sys_excl_index"], d["sys_excl_atoms"])
        assert relrms(f, fb) < 2e-6, (lam, relrms(f, fb))
        assert abs(e_cl[0] + e_lj - eb[0]) <= 2e-6 * abs(eb[0]) + 1e-5, (lam, e_cl[0] + e_lj, eb[0])
        assert abs(e_cl[1] + e_el - eb[1]) <= 2e-6 * abs(eb[1]), (lam, e_cl[1] + e_el, eb[1])


def test_split_runs_once_per_list():
    """the split clears the moved pairs' bits in the cluster list, so a second pass over the same list would re-list them as
    excluded pairs: it is refused until a new list has been built"""
    import ctypes as C
    from gromacs_b200.nbnxm import load_library
    from gromacs_b200.pairsearch import split_fep_pairlist as split_fep_list
    from gromacs_b200.workload import make_workload
    wl = make_workload("bench3k", nthreads=2)
    wl.pairlist(min_sci=0)
    pert = np.zeros(wl.box.natoms, np.uint8)
    pert[:30] = 1
    _, fep = split_fep_list(wl.grid, pert)
    assert fep["jjnr"].size > 0
    lib = load_library()
    assert lib.nbnxm_b200_pairlist_split_fep(wl.grid._g, pert.ctypes.data_as(C.POINTER(C.c_ubyte))) != 0
    wl.pairlist(min_sci=0)          # a new list: the split is allowed again and gives the same perturbed list
    _, fep2 = split_fep_list(wl.grid, pert)
    assert np.array_equal(fep2["jjnr"], fep["jjnr"]) and np.array_equal(fep2["excl_fep"], fep["excl_fep"])
