"""CPU tests: pin the oracle (oracle/nbnxm_oracle.c) against the reference.

 - the reference's own golden XML values for the 243-atom TestSystem (all 18 Coulomb x VdW files that
   the GPU flavors cover), parsed into tests/golden/test243_*.npz by tests/golden/make_golden.py;
 - forces / energies / shift forces computed by the reference's SIMD 4xM kernel (and its plain-C
   GPU-layout kernel) in the dev container on the same inputs.
"""
import numpy as np
import pytest

from util import golden_cases, load_golden, maxrel, oracle_forces, oracle_params, relrms


@pytest.mark.parametrize("case", golden_cases("test243_"))
def test_oracle_matches_reference_xml_refdata(oracle, case):
    d = load_golden(case)
    if "split" in case:
        pytest.skip("non-default cut-off: not the refdata configuration")
    p = oracle_params(oracle, d)
    f, fsh, e, _ = oracle_forces(oracle, d, p)
    fa = oracle.nbat_to_atom_order(f, d["nbat_atom_index"], d["sys_x"].shape[0])
    # refdata was generated in double from double coordinates; our inputs are the float coordinates
    assert relrms(fa, d["refdata_f"]) < 2e-6
    assert maxrel(fa, d["refdata_f"]) < 1e-5
    assert abs(e[0] - d["refdata_vvdw"][0]) < 2e-6 * abs(d["refdata_vvdw"][0]) + 1e-5
    assert abs(e[1] - d["refdata_vcoul"][0]) < 2e-6 * abs(d["refdata_vcoul"][0])


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_simd_kernel(oracle, case):
    d = load_golden(case)
    p = oracle_params(oracle, d)
    f, fsh, e, npairs = oracle_forces(oracle, d, p)
    fa = oracle.nbat_to_atom_order(f, d["nbat_atom_index"], d["sys_x"].shape[0])
    ref = d["ref_simd4xm_f"].astype(np.float64)
    assert relrms(fa, ref) < 3e-6          # the SIMD kernel computes in float32
    assert maxrel(fa, ref) < 2e-5
    assert abs(e[0] - d["ref_simd4xm_vvdw"][0]) < 1e-5 * abs(e[0]) + 1e-4
    assert abs(e[1] - d["ref_simd4xm_vcoul"][0]) < 5e-5 * abs(e[1])   # float32 accumulation in the reference
    # shift forces: every non-central entry (the GPU path skips the central one, whose shift vector is 0)
    fs_ref = d["ref_simd4xm_fshift"].astype(np.float64)
    mask = np.ones(45, bool)
    mask[22] = False
    assert np.abs(fsh[mask] - fs_ref[mask]).max() < 2e-5 * np.abs(fs_ref[mask]).max() + 1e-3
    # the pair count is 32 per set mask bit
    cj = d["pl_cjPacked"]
    bits = sum(bin(int(v)).count("1") for v in np.concatenate([cj[:, 4], cj[:, 6]]))
    assert npairs == 32 * bits


@pytest.mark.parametrize("case", ["test243_ewald_cutnone", "bench1_ewald_cutnone", "test243_rf_cutnone"])
def test_oracle_tabulated_ewald_matches_reference_plainc_gpu_layout_kernel(oracle, case):
    """nbnxm_kernel_gpu_ref walks the identical sci/cjPacked/excl arrays (RF or tabulated Ewald + plain LJ)."""
    d = load_golden(case)
    elec = "RF" if "rf" in case else "EwaldTab"
    p = oracle_params(oracle, d, elec=elec)
    f, fsh, e, _ = oracle_forces(oracle, d, p)
    fa = oracle.nbat_to_atom_order(f, d["nbat_atom_index"], d["sys_x"].shape[0])
    ref = d["ref_gpulayout_plainc_f"].astype(np.float64)
    assert relrms(fa, ref) < 2e-5   # the reference's table has ~1e-6 interpolation error per pair
    assert abs(e[0] - d["ref_gpulayout_plainc_vvdw"][0]) < 1e-5 * abs(e[0]) + 1e-4


@pytest.mark.parametrize("case", ["test243_ewald_cutnone", "test243_ewald_ljpmegeom", "test243_rf_fswitch",
                                  "test243_ewaldtwin_pswitch", "bench1_ewald_cutgeom"])
def test_list_walk_equals_brute_force(oracle, case):
    """The reference's pair list + exclusion masks cover exactly the pairs within the cut-off."""
    d = load_golden(case)
    p = oracle_params(oracle, d)
    f, _, e, _ = oracle_forces(oracle, d, p)
    fa = oracle.nbat_to_atom_order(f, d["nbat_atom_index"], d["sys_x"].shape[0])
    fb, eb = oracle.brute_force(p, d["sys_x"], d["sys_q"], d["sys_type"], d["nbat_nbfp"], d["nbat_nbfp_comb"],
                                d["sys_box"], d["sys_excl_index"], d["sys_excl_atoms"])
    # the list walk rounds x+shift to float like the kernels do; brute force uses minimum image in double
    assert relrms(fa, fb) < 1e-6
    assert abs(e[0] - eb[0]) < 1e-6 * abs(eb[0]) + 1e-6
    assert abs(e[1] - eb[1]) < 1e-6 * abs(eb[1])


@pytest.mark.parametrize("case", ["test243_ewald_cutnone_rl1.0_split", "bench1_ewald_cutgeom", "bench1_rf_cutnone_split"])
def test_prune_restatement_properties(oracle, case):
    d = load_golden(case)
    rl = float(d["rlist"][0])
    rc = float(d["ic_rcoulomb"][0])
    rin = 0.5 * (rl + rc)
    p = oracle_params(oracle, d, rlist_inner=rin)
    cj0 = d["pl_cjPacked"].copy()
    cj = cj0.copy()
    outer = np.zeros(2 * cj.shape[0], np.uint32)
    cnt = oracle.prune(p, d["pl_sci"], cj, outer, d["nbat_xq"], d["shift_vec"], fresh=True)
    orig = np.stack([cj0[:, 4], cj0[:, 6]], 1).reshape(-1)
    inner = np.stack([cj[:, 4], cj[:, 6]], 1).reshape(-1)
    assert np.all((outer & ~orig) == 0) and np.all((inner & ~outer) == 0)      # inner in outer in original
    assert inner.sum() < orig.sum()                                            # pruning removed something
    assert np.all(cnt >= 0) and np.all(cnt < 8192)
    # physics is unchanged by pruning: pruned pairs lie beyond rlistInner >= rc
    f0, _, e0, n0 = oracle_forces(oracle, d, p, cjp=cj0)
    f1, _, e1, n1 = oracle_forces(oracle, d, p, cjp=cj)
    assert n1 < n0
    assert np.array_equal(f0, f1) and np.array_equal(e0, e1)
    # rolling prune on unchanged coordinates changes nothing and only ever adds bits
    cj2 = cj.copy()
    for part in range(2):
        oracle.prune(p, d["pl_sci"], cj2, outer, d["nbat_xq"], d["shift_vec"], fresh=False, part=part, nparts=2)
    assert np.array_equal(cj2, cj)
    xq = d["nbat_xq"].copy()
    rng = np.random.default_rng(1)
    xq[:, :3] += rng.normal(0, 0.02, (xq.shape[0], 3)).astype(np.float32)
    cj3 = cj.copy()
    oracle.prune(p, d["pl_sci"], cj3, outer, xq, d["shift_vec"], fresh=False, part=0, nparts=1)
    inner3 = np.stack([cj3[:, 4], cj3[:, 6]], 1).reshape(-1)
    assert np.all((inner & ~inner3) == 0) and np.all((inner3 & ~outer) == 0) and inner3.sum() > inner.sum()


def test_f32_port_matches_double_oracle(oracle):
    d = load_golden("bench1_ewald_cutgeom")
    p = oracle_params(oracle, d)
    f, _, e, n = oracle_forces(oracle, d, p)
    f32, e32, n32 = oracle.forces_f32_omp(p, d["pl_sci"], d["pl_cjPacked"], d["pl_excl"], d["nbat_xq"], d["nbat_type"],
                                          d["nbat_lj_comb"], d["nbat_nbfp"], d["nbat_nbfp_comb"], d["shift_vec"],
                                          calc_energy=True, nthreads=2)
    assert n32 == n
    assert relrms(f32.astype(np.float64), f) < 5e-6
    assert abs(e32[1] - e[1]) < 1e-4 * abs(e[1])


@pytest.mark.parametrize("vdw", ["cutnone", "cutgeom", "cutlb", "fswitch", "pswitch", "ljpmegeom"])
def test_oracle_lj_only_matches_reference_coulomb_none_refdata(oracle, vdw):
    """ElecType::None (kernels without NBNxM electrostatics): the reference's golden XML for CoulombKernelType::None
    (tests/golden/refdata/coulombnone.npz) against the oracle's plain cut-off flavor with the charges set to zero —
    the oracle configuration the GPU's ElecNone flavor is checked against (tests/test_gpu_zz_elec_none.py)."""
    import os
    from util import GOLDEN
    ref = np.load(os.path.join(GOLDEN, "refdata", "coulombnone.npz"))
    d = load_golden("test243_ewald_" + vdw)
    d0 = dict(d)
    d0["nbat_xq"] = d["nbat_xq"].copy()
    d0["nbat_xq"][:, 3] = 0
    f, _, e, _ = oracle_forces(oracle, d0, oracle_params(oracle, d0, elec="Cut"))
    fa = oracle.nbat_to_atom_order(f, d["nbat_atom_index"], d["sys_x"].shape[0])
    assert ref["vcoul_" + vdw][0] == 0 and e[1] == 0
    # refdata comes from double coordinates, ours are their float roundings: 3e-6 of the (smaller) LJ-only forces
    assert relrms(fa, ref["f_" + vdw]) < 4e-6
    assert maxrel(fa, ref["f_" + vdw]) < 1e-5
    assert abs(e[0] - ref["vvdw_" + vdw][0]) < 2e-6 * abs(ref["vvdw_" + vdw][0]) + 1e-5


@pytest.mark.parametrize("case", ["test243_ewald_cutnone_rl1.0_split", "bench1_ewald_cutgeom"])
def test_prune_restatement_against_numpy_distances(oracle, case):
    """The masks the prune restatement keeps, re-derived independently in numpy (double precision): bit jm*8+i of half w
    survives iff some atom pair of i-cluster i and the 4 j-atoms of half w of j-cluster jm lies within the list radius
    (nbnxm_cuda_kernel_pruneonly.cuh:223-321).  Pairs within 1e-6 (relative) of the radius may fall either way in
    float32; everything else must agree, for the outer and the inner mask."""
    d = load_golden(case)
    rl = float(d["rlist"][0])
    rin = 0.5 * (rl + float(d["ic_rcoulomb"][0]))
    p = oracle_params(oracle, d, rlist_inner=rin)
    cj0 = d["pl_cjPacked"].copy()
    cj = cj0.copy()
    outer = np.zeros(2 * cj.shape[0], np.uint32)
    oracle.prune(p, d["pl_sci"], cj, outer, d["nbat_xq"], d["shift_vec"], fresh=True)
    x = d["nbat_xq"][:, :3].astype(np.float64)
    sv = d["shift_vec"].astype(np.float64)
    checked = 0
    for s in d["pl_sci"]:
        xi = x[s[0] * 64:(s[0] + 1) * 64].reshape(8, 8, 3) + sv[s[1]]                  # [i-cluster, atom]
        for g in range(s[2], s[3]):
            for jm in range(4):
                xj = x[int(cj0[g, jm]) * 8:int(cj0[g, jm]) * 8 + 8]
                r2 = ((xi[:, :, None, :] - xj[None, None, :, :]) ** 2).sum(-1)            # [i-cluster, i-atom, j-atom]
                for w in range(2):
                    r2min = r2[:, :, 4 * w:4 * w + 4].min(axis=(1, 2))                      # per i-cluster
                    for i in range(8):
                        bit = 1 << (jm * 8 + i)
                        if not (int(cj0[g, 4 + 2 * w]) & bit):
                            assert not (int(outer[2 * g + w]) & bit) and not (int(cj[g, 4 + 2 * w]) & bit)
                            continue
                        for radius, kept in ((rl, int(outer[2 * g + w]) & bit), (rin, int(cj[g, 4 + 2 * w]) & bit)):
                            if r2min[i] < radius * radius * (1 - 1e-6):
                                assert kept, (g, jm, w, i, r2min[i], radius)
                            elif r2min[i] > radius * radius * (1 + 1e-6):
                                assert not kept, (g, jm, w, i, r2min[i], radius)
                            checked += 1
    assert checked > 1000
