"""Shared helpers for the test-suite: golden fixtures -> product inputs / oracle inputs."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ELEC_OF_CASE = {"ewald": "EwaldAna", "ewaldtwin": "EwaldAnaTwin", "rf": "RF", "cut": "Cut",
                "ewaldtab": "EwaldTab", "ewaldtabtwin": "EwaldTabTwin"}
VDW_OF_CASE = {"cutnone": "Cut", "cutgeom": "CutCombGeom", "cutlb": "CutCombLB", "fswitch": "FSwitch",
               "pswitch": "PSwitch", "ljpmegeom": "EwaldGeom", "ljpmelb": "EwaldLB"}


def golden_cases(prefix=""):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def maxrel(a, b):
    """max-component error relative to the largest force magnitude"""
    return float(np.abs(a - b).max() / np.abs(b).max())


def product_params(d, elec=None, vdw=None, rlist_inner=None, dynamic_pruning=False):
    from gromacs_b200 import make_params
    g = lambda k: float(d[k][0])
    return make_params(
        elec or ELEC_OF_CASE[str(d["case_coulomb"])], vdw or VDW_OF_CASE[str(d["case_vdw"])],
        epsfac=g("ic_epsfac"), rcoulomb=g("ic_rcoulomb"), rvdw=g("ic_rvdw"), rlist_outer=g("rlist"),
        rlist_inner=rlist_inner, ewald_beta=g("ic_ewald_beta"), sh_ewald=g("ic_sh_ewald"), k_rf=g("ic_k_rf"),
        c_rf=g("ic_c_rf"), rvdw_switch=g("ic_rvdw_switch"),
        disp=(g("ic_disp_c2"), g("ic_disp_c3"), g("ic_disp_cpot")),
        rep=(g("ic_rep_c2"), g("ic_rep_c3"), g("ic_rep_cpot")),
        sw=(g("ic_sw_c3"), g("ic_sw_c4"), g("ic_sw_c5")), ewaldcoeff_lj=g("ic_ewaldcoeff_lj"),
        sh_lj_ewald=g("ic_sh_lj_ewald"),
        coulomb_tab_scale=g("ic_coulomb_tab_scale") if "ic_coulomb_tab_scale" in d else 0.0,
        use_dynamic_pruning=dynamic_pruning)


def product_inputs(d):
    from gromacs_b200 import AtomData, PairlistGpu
    nbat = AtomData(xq=d["nbat_xq"], type=d["nbat_type"], lj_comb=d["nbat_lj_comb"], nbfp=d["nbat_nbfp"],
                    nbfp_comb=d["nbat_nbfp_comb"], numTypes=int(d["nbat_ntypes"][0]), shift_vec=d["shift_vec"])
    plist = PairlistGpu(sci=d["pl_sci"], cjPacked=d["pl_cjPacked"], excl=d["pl_excl"], na_ci=int(d["pl_na_ci"][0]),
                        rlist=float(d["rlist"][0]))
    return nbat, plist


def oracle_params(O, d, elec=None, vdw=None, rlist_inner=None):
    p = O.params_from_golden(d, rlist_inner=rlist_inner)
    from gromacs_b200.nbnxm import ELEC_TYPES, VDW_TYPES
    if elec is not None:
        p.elec_type = ELEC_TYPES[elec]
    if vdw is not None:
        p.vdw_type = VDW_TYPES[vdw]
    return p


def oracle_forces(O, d, p, sci=None, cjp=None, calc_energy=True):
    return O.forces(p, d["pl_sci"] if sci is None else sci, d["pl_cjPacked"] if cjp is None else cjp, d["pl_excl"],
                    d["nbat_xq"], d["nbat_type"], d["nbat_lj_comb"], d["nbat_nbfp"], d["nbat_nbfp_comb"],
                    d["shift_vec"], tab=d.get("ic_coulomb_tab_F"), calc_energy=calc_energy)
