"""GPU parity of the perturbed (FEP) pair kernel through the C ABI (nbnxm_b200_copy_fepparams / _init_fep_atomdata /
_init_feppairlist / _launch_free_energy_kernel) against the reference's golden data for its GPU FEP kernel and the
pinned oracle.  The kernel body is also checked on the CPU (tests/test_fep_emu.py)."""
import numpy as np
import pytest

from test_oracle_fep import cases, fep_test_system

pytestmark = pytest.mark.gpu

ELEC = {"cut": "Cut", "rf": "RF", "ewald": "EwaldAna"}
VDW = {"cut": "Cut", "cutgeom": "CutCombGeom", "cutlb": "CutCombLB", "fswitch": "FSwitch", "pswitch": "PSwitch"}


def run_gpu(p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, shift_vec, lst):
    from gromacs_b200 import LOCAL, AtomData, NbnxmGpu, StepWorkload, make_params
    n = len(x)
    npad = (n + 7) // 8 * 8
    xq = np.full((npad, 4), -1.0e6, np.float32)
    xq[:, 3] = 0
    xq[:n, :3] = x
    pad = lambda a, fill, dt: np.concatenate([np.asarray(a, dt), np.full((npad - n,) + np.asarray(a).shape[1:], fill, dt)])
    params = make_params(ELEC[p.elec] + ("Twin" if p.twin and p.elec == "ewald" else ""), VDW[p.vdw], epsfac=p.epsfac,
                         rcoulomb=p.rcoulomb_sq ** 0.5, rvdw=p.rvdw_sq ** 0.5, rlist_outer=max(p.rcoulomb_sq, p.rvdw_sq) ** 0.5,
                         ewald_beta=p.ewald_beta, sh_ewald=p.sh_ewald, k_rf=0.5 * p.two_k_rf, c_rf=p.c_rf,
                         rvdw_switch=p.rvdw_switch, disp=p.disp, rep=p.rep, sw=p.sw)
    comb = p.vdw in ("cutgeom", "cutlb")
    nbat = AtomData(xq=xq, type=pad(type_a, 0, np.int32), lj_comb=pad(lj_a, 0, np.float32) if comb else None,
                    nbfp=np.ascontiguousarray(p.nbfp, np.float32), nbfp_comb=None, numTypes=p.ntypes,
                    shift_vec=np.ascontiguousarray(shift_vec, np.float32))
    nb = NbnxmGpu(params, nbat)
    try:
        sw = StepWorkload(computeEnergy=True, computeVirial=True)
        nb.gpu_init_atomdata(nbat)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.copy_gpu_fepparams(True, p.alpha_coul, p.alpha_vdw, p.lambda_power, p.sigma6_with_invalid_sigma, p.sigma6_minimum,
                              p.lambda_coul, p.lambda_vdw)
        nb.gpu_init_fep_atomdata(pad(q_a, 0, np.float32), pad(q_b, 0, np.float32), pad(type_a, 0, np.int32), pad(type_b, 0, np.int32),
                                 pad(lj_a, 0, np.float32) if comb else None, pad(lj_b, 0, np.float32) if comb else None)
        nb.gpu_init_feppairlist(lst["iinr"], lst["jindex"], lst["jjnr"], lst["shift"], lst["excl_fep"], LOCAL)
        nb.gpu_clear_outputs(True)
        nb.gpu_launch_free_energy_kernel(sw, LOCAL)
        nb.gpu_launch_cpyback(nbat, sw, LOCAL)
        fshift = np.zeros((45, 3), np.float32)
        e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL, shiftForces=fshift)
        dvdl_lj, dvdl_el = nb.gpu_get_fep_dvdl()
    finally:
        nb.gpu_free()
    return nbat.f[:n].astype(np.float64), fshift.astype(np.float64), e_lj, e_el, dvdl_lj, dvdl_el


@pytest.mark.parametrize("name,ref", cases()[::13], ids=[c[0] for c in cases()[::13]])
def test_fep_kernel_matches_reference_gpu_refdata(name, ref):
    p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst = fep_test_system(name)
    f, fshift, e_lj, e_el, dvdl_lj, dvdl_el = run_gpu(p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((45, 3)), lst)
    scale = np.abs(ref[4:16]).max()
    assert np.abs(f.reshape(-1) - ref[4:16]).max() <= 5e-6 * scale
    assert np.abs(fshift[0] - ref[16:19]).max() <= 5e-6 * scale
    for got, want, floor in ((e_lj, ref[0], 1.0), (e_el, ref[1], 100.0), (dvdl_el, ref[2], 100.0), (dvdl_lj, ref[3], 1.0)):
        assert abs(got - want) <= 5e-6 * max(abs(want), floor), (name, got, want)


def test_foreign_lambda_energies_equal_the_full_evaluation_at_those_lambdas():
    """nbnxm_b200_launch_foreign_energy_kernel against the oracle evaluated at each foreign lambda"""
    import copy
    from gromacs_b200 import LOCAL, AtomData, NbnxmGpu, StepWorkload, make_params
    from oracle.nbfe_oracle import nbfe_forces
    name = [c[0] for c in cases() if "Reaction_Field" in c[0] and "0_5_0_3_scCoulomb_Yes" in c[0] and "ljrule_None" in c[0]][0]
    p, x, q_a, q_b, type_a, type_b, lj_a, lj_b, lst = fep_test_system(name)
    lambdas = np.array([0.0, 0.2, 0.5, 0.9, 1.0], np.float32)
    want = []
    for lam in lambdas:
        q = copy.copy(p)
        q.lambda_coul = q.lambda_vdw = float(lam)
        want.append(nbfe_forces(q, x, q_a, q_b, type_a, type_b, lj_a, lj_b, np.zeros((1, 3)), **lst)[2:])
    want = np.array(want)
    n, npad = 4, 8
    xq = np.full((npad, 4), -1.0e6, np.float32)
    xq[:, 3] = 0
    xq[:n, :3] = x
    pad = lambda a, dt: np.concatenate([np.asarray(a, dt), np.zeros(npad - n, dt)])
    params = make_params("RF", "Cut", epsfac=p.epsfac, rcoulomb=1.0, rvdw=1.0, rlist_outer=1.0, k_rf=0.0, c_rf=p.c_rf, disp=p.disp,
                         rep=p.rep)
    nbat = AtomData(xq=xq, type=pad(type_a, np.int32), nbfp=np.ascontiguousarray(p.nbfp, np.float32), numTypes=p.ntypes,
                    shift_vec=np.zeros((45, 3), np.float32))
    nb = NbnxmGpu(params, nbat)
    try:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.copy_gpu_fepparams(True, p.alpha_coul, p.alpha_vdw, p.lambda_power, p.sigma6_with_invalid_sigma, p.sigma6_minimum, 0.5, 0.5)
        nb.gpu_init_fep_atomdata(pad(q_a, np.float32), pad(q_b, np.float32), pad(type_a, np.int32), pad(type_b, np.int32))
        nb.gpu_init_feppairlist(lst["iinr"], lst["jindex"], lst["jjnr"], lst["shift"], lst["excl_fep"], LOCAL)
        got = nb.gpu_launch_foreign_energy_kernel(lambdas, lambdas, LOCAL)
    finally:
        nb.gpu_free()
    assert np.abs(got - want).max() <= 5e-6 * np.abs(want).max()


def test_cluster_and_perturbed_kernels_together_equal_brute_force(oracle):
    """the whole perturbed path on the GPU: host builder + split, masked cluster kernel + perturbed kernel adding into the
    same forces / energies, against brute force of the A-state system at lambda = 0 and of the B-state at lambda = 1"""
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.pairsearch import Grid, split_fep_pairlist
    from util import load_golden, oracle_params, product_params, relrms
    d = load_golden("bench1_ewald_cutnone")
    x, box = d["sys_x"], d["sys_box"]
    n = x.shape[0]
    nt = int(d["nbat_ntypes"][0])
    q_a, t_a = d["sys_q"].astype(np.float32), d["sys_type"].astype(np.int32)
    rng = np.random.default_rng(4)
    perturbed = np.zeros(n, np.uint8)
    q_b, t_b = q_a.copy(), t_a.copy()
    for k, m in enumerate(rng.choice(n // 3, size=45, replace=False)):
        atoms = np.arange(3 * m, 3 * m + 3)
        perturbed[atoms] = 1
        q_b[atoms] *= (0.0 if k % 3 == 0 else 0.5)
        if k % 2 == 0:
            t_b[atoms] = nt - 1
    grid = Grid(box, x, nthreads=2)
    ai = grid.atom_index
    real = ai >= 0
    nbat = grid.atomdata(x, np.where(perturbed, 0.0, q_a), np.where(perturbed, nt - 1, t_a), d["nbat_nbfp"], nt,
                         nbfp_comb=d["nbat_nbfp_comb"])
    grid.pairlist(1.0, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=200)
    plist, fep = split_fep_pairlist(grid, perturbed)

    def nbat_order(a, fill):
        out = np.full(ai.shape[0], fill, np.asarray(a).dtype)
        out[real] = np.asarray(a)[ai[real]]
        return out
    p = oracle_params(oracle, d)
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        sw = StepWorkload(computeEnergy=True, computeVirial=True)
        nb.gpu_init_atomdata(nbat)
        nb.gpu_init_pairlist(plist, LOCAL)
        nb.setupGpuShortRangeWork(LOCAL)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        nb.gpu_init_fep_atomdata(nbat_order(q_a, 0.0), nbat_order(q_b, 0.0), nbat_order(t_a, nt - 1), nbat_order(t_b, nt - 1))
        nb.gpu_init_feppairlist(fep["iinr"], fep["jindex"], fep["jjnr"], fep["shift"], fep["excl_fep"], LOCAL)
        for lam, q_end, t_end in ((0.0, q_a, t_a), (1.0, q_b, t_b)):
            nb.copy_gpu_fepparams(True, 0.0, 0.0, 1, 0.3 ** 6, 0.3 ** 6, lam, lam)
            nb.gpu_clear_outputs(True)
            nb.gpu_launch_kernel(sw, LOCAL)
            nb.gpu_launch_free_energy_kernel(sw, LOCAL)
            nb.gpu_launch_cpyback(nbat, sw, LOCAL)
            e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL)
            f = oracle.nbat_to_atom_order(nbat.f.astype(np.float64), ai, n)
            fb, eb = oracle.brute_force(p, x, q_end, t_end, d["nbat_nbfp"], d["nbat_nbfp_comb"], box, d["sys_excl_index"],
                                        d["sys_excl_atoms"])
            assert relrms(f, fb) <= 5e-6, (lam, relrms(f, fb))
            assert abs(e_lj - eb[0]) <= 2e-6 * abs(eb[0]) + 2e-6 and abs(e_el - eb[1]) <= 2e-6 * abs(eb[1]), (lam, e_lj, e_el, eb)
    finally:
        nb.gpu_free()


def test_device_built_and_split_lists_equal_the_hosts_and_brute_force(oracle):
    """the perturbed path with nothing of the search on the host: list built on the device, perturbed pairs split off on the device
    (nbnxm_b200_gpu_search_set_perturbed) and installed device to device - the same cluster and perturbed lists as the host builder +
    nbnxm_b200_pairlist_split_fep, entry for entry, and masked cluster kernel + perturbed kernel against brute force of the A-state
    system at lambda = 0 and of the B-state system at lambda = 1; then the same from coordinates gridded on the device"""
    import torch
    from gromacs_b200 import LOCAL, NbnxmGpu, StepWorkload
    from gromacs_b200.pairsearch import Grid, GpuPairSearch, split_fep_pairlist
    from test_gpusearch_emu import assert_same_list
    from util import load_golden, oracle_params, product_params, relrms
    d = load_golden("bench1_ewald_cutnone")
    x, box = d["sys_x"], d["sys_box"]
    n = x.shape[0]
    nt = int(d["nbat_ntypes"][0])
    q_a, t_a = d["sys_q"].astype(np.float32), d["sys_type"].astype(np.int32)
    rng = np.random.default_rng(4)
    perturbed = np.zeros(n, np.uint8)
    q_b, t_b = q_a.copy(), t_a.copy()
    for k, m in enumerate(rng.choice(n // 3, size=45, replace=False)):
        atoms = np.arange(3 * m, 3 * m + 3)
        perturbed[atoms] = 1
        q_b[atoms] *= (0.0 if k % 3 == 0 else 0.5)
        if k % 2 == 0:
            t_b[atoms] = nt - 1
    grid = Grid(box, x, nthreads=2)
    ai = grid.atom_index
    real = ai >= 0
    nbat = grid.atomdata(x, np.where(perturbed, 0.0, q_a), np.where(perturbed, nt - 1, t_a), d["nbat_nbfp"], nt,
                         nbfp_comb=d["nbat_nbfp_comb"])
    grid.pairlist(1.0, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=200)
    ref, fep_ref = split_fep_pairlist(grid, perturbed)

    def nbat_order(a, fill):
        out = np.full(ai.shape[0], fill, np.asarray(a).dtype)
        out[real] = np.asarray(a)[ai[real]]
        return out
    p = oracle_params(oracle, d)
    sw = StepWorkload(computeEnergy=True, computeVirial=True)

    def end_states(nb):
        nb.gpu_init_fep_atomdata(nbat_order(q_a, 0.0), nbat_order(q_b, 0.0), nbat_order(t_a, nt - 1), nbat_order(t_b, nt - 1))
        for lam, q_end, t_end in ((0.0, q_a, t_a), (1.0, q_b, t_b)):
            nb.copy_gpu_fepparams(True, 0.0, 0.0, 1, 0.3 ** 6, 0.3 ** 6, lam, lam)
            nb.gpu_clear_outputs(True)
            nb.gpu_launch_kernel(sw, LOCAL)
            nb.gpu_launch_free_energy_kernel(sw, LOCAL)
            nb.gpu_launch_cpyback(nbat, sw, LOCAL)
            e_lj, e_el = nb.gpu_wait_finish_task(sw, LOCAL)
            f = oracle.nbat_to_atom_order(nbat.f.astype(np.float64), ai, n)
            fb, eb = oracle.brute_force(p, x, q_end, t_end, d["nbat_nbfp"], d["nbat_nbfp_comb"], box, d["sys_excl_index"],
                                        d["sys_excl_atoms"])
            assert relrms(f, fb) <= 5e-6, (lam, relrms(f, fb))
            assert abs(e_lj - eb[0]) <= 2e-6 * abs(eb[0]) + 2e-6 and abs(e_el - eb[1]) <= 2e-6 * abs(eb[1]), (lam, e_lj, e_el, eb)

    def same_fep(a, b):
        for key in ("iinr", "jindex", "jjnr", "shift"):
            assert np.array_equal(a[key], b[key]), key
        assert np.array_equal(a["excl_fep"] != 0, b["excl_fep"] != 0)

    # (1) host grid, list and split on the device
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        nb.gpu_init_atomdata(nbat)
        nb.gpu_upload_shiftvec(nbat)
        nb.gpu_copy_xq_to_gpu(nbat, LOCAL)
        search = GpuPairSearch(nb, grid, d["sys_excl_index"], d["sys_excl_atoms"])
        search.set_perturbed(perturbed)
        search.build(1.0, LOCAL, min_sci=200)
        got = search.download()
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
        same_fep(search.fep_download(), fep_ref)
        nb.setupGpuShortRangeWork(LOCAL)
        end_states(nb)
        # switched off again: the plain list, no perturbed pairs
        search.set_perturbed(None)
        search.build(1.0, LOCAL, min_sci=200)
        plain = grid.pairlist(1.0, d["sys_excl_index"], d["sys_excl_atoms"], min_sci=200)
        got = search.download()
        assert_same_list((got.sci, got.cjPacked, got.excl), (plain.sci, plain.cjPacked, plain.excl))
        assert search.fep_download()["jjnr"].size == 0
        search.free()
    finally:
        nb.gpu_free()
    # (2) gridding on the device as well: the perturbed atoms are masked out of the cluster kernels' atom data there
    nb = NbnxmGpu(product_params(d, vdw="Cut"), nbat)
    try:
        nb.gpu_upload_shiftvec(nbat)
        x_dev = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
        torch.cuda.synchronize()
        search = GpuPairSearch(nb)
        search.set_atoms(q_a, t_a, nt, None, d["sys_excl_index"], d["sys_excl_atoms"])
        search.set_perturbed(perturbed)
        search.put_atoms_on_grid(box, x_dev.data_ptr())
        atom_index, _, _ = search.get_order()
        assert np.array_equal(atom_index, ai)
        search.build(1.0, LOCAL, min_sci=200)
        got = search.download()
        assert_same_list((got.sci, got.cjPacked, got.excl), (ref.sci, ref.cjPacked, ref.excl))
        same_fep(search.fep_download(), fep_ref)
        nb.setupGpuShortRangeWork(LOCAL)
        end_states(nb)
        search.free()
    finally:
        nb.gpu_free()
