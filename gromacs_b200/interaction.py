"""Kernel parameters from interaction constants, following initNbparam / set_cutoff_parameters and the
flavor pickers of the reference (src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:168-216, 218-240, 368-460)."""
from .nbnxm import ELEC_TYPES, VDW_TYPES, Params


def pick_vdw_type(vdw_type, vdw_modifier, lj_comb_rule, ljpme_comb_rule="Geom"):
    """nbnxmGpuPickVdwKernelType (nbnxm_gpu_data_mgmt.cpp:368-420).
    vdw_type: "Cut" | "Pme"; vdw_modifier: "None" | "PotShift" | "ForceSwitch" | "PotSwitch";
    lj_comb_rule: "None" | "Geometric" | "LorentzBerthelot"."""
    if vdw_type == "Cut":
        if vdw_modifier in ("None", "PotShift"):
            return {"None": "Cut", "Geometric": "CutCombGeom", "LorentzBerthelot": "CutCombLB"}[lj_comb_rule]
        if vdw_modifier == "ForceSwitch":
            return "FSwitch"
        if vdw_modifier == "PotSwitch":
            return "PSwitch"
        raise ValueError("The requested VdW interaction modifier %s is not implemented in the GPU kernels" % vdw_modifier)
    if vdw_type == "Pme":
        return "EwaldGeom" if ljpme_comb_rule == "Geom" else "EwaldLB"
    raise ValueError("The requested VdW type %s is not implemented in the GPU kernels" % vdw_type)


def pick_elec_type(coulomb_type, rcoulomb, rvdw, analytical=True):
    """nbnxmGpuPickElectrostaticsKernelType (nbnxm_gpu_data_mgmt.cpp:168-216, 422-460).
    Analytical Ewald is the reference's default on every NVIDIA device except CC 7.0 / 8.0."""
    if coulomb_type == "Cut":
        return "Cut"
    if coulomb_type == "Fmm":
        return "None"           # FMM computes its own direct part (nbnxmIsDirectCoulombProvider false)
    if coulomb_type == "RF":
        return "RF"
    if coulomb_type in ("Pme", "Ewald"):
        twin = rcoulomb != rvdw
        if analytical:
            return "EwaldAnaTwin" if twin else "EwaldAna"
        return "EwaldTabTwin" if twin else "EwaldTab"
    raise ValueError("The requested electrostatics type %s is not implemented in the GPU kernels" % coulomb_type)


def make_params(elec, vdw, *, epsfac, rcoulomb, rvdw, rlist_outer, rlist_inner=None, ewald_beta=0.0,
                sh_ewald=0.0, k_rf=0.0, c_rf=0.0, rvdw_switch=0.0, disp=(0.0, 0.0, 0.0), rep=(0.0, 0.0, 0.0),
                sw=(0.0, 0.0, 0.0), ewaldcoeff_lj=0.0, sh_lj_ewald=0.0, coulomb_tab_scale=0.0,
                use_dynamic_pruning=False):
    """set_cutoff_parameters (nbnxm_gpu_data_mgmt.cpp:218-240)."""
    p = Params()
    p.elec_type = ELEC_TYPES[elec] if isinstance(elec, str) else int(elec)
    p.vdw_type = VDW_TYPES[vdw] if isinstance(vdw, str) else int(vdw)
    p.epsfac = epsfac
    p.c_rf = c_rf
    p.two_k_rf = 2.0 * k_rf
    p.ewald_beta = ewald_beta
    p.sh_ewald = sh_ewald
    p.sh_lj_ewald = sh_lj_ewald
    p.ewaldcoeff_lj = ewaldcoeff_lj
    p.rcoulomb_sq = rcoulomb * rcoulomb
    p.rvdw_sq = rvdw * rvdw
    p.rvdw_switch = rvdw_switch
    p.rlist_outer_sq = rlist_outer * rlist_outer
    ri = rlist_outer if rlist_inner is None else rlist_inner
    p.rlist_inner_sq = ri * ri
    p.disp_c2, p.disp_c3, p.disp_cpot = disp
    p.rep_c2, p.rep_c3, p.rep_cpot = rep
    p.sw_c3, p.sw_c4, p.sw_c5 = sw
    p.coulomb_tab_scale = coulomb_tab_scale
    p.use_dynamic_pruning = int(use_dynamic_pruning)
    return p
