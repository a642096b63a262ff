"""GPU parity of ElecType::None (no NBNxM electrostatics; the reference's ElecNone kernels, cuda/nbnxm_cuda.cu:213):
LJ only, whatever the charges are.  Oracle: the plain cut-off flavor on the same list with all charges set to zero."""
import numpy as np
import pytest

from test_gpu_parity import E_REL, check_forces, run_step
from util import load_golden, oracle_forces, oracle_params, product_inputs, product_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["test243_ewald_cutnone", "bench1_ewald_fswitch", "bench1_ewald_cutgeom"])
def test_elec_none_is_lj_only(oracle, case, monkeypatch):
    from gromacs_b200 import NbnxmGpu
    d = load_golden(case)
    nbat, plist = product_inputs(d)
    d0 = dict(d)
    d0["nbat_xq"] = d["nbat_xq"].copy()
    d0["nbat_xq"][:, 3] = 0
    f_ref, _, e_ref, _ = oracle_forces(oracle, d0, oracle_params(oracle, d0, elec="Cut"))
    assert e_ref[1] == 0 and np.abs(nbat.xq[:, 3]).max() > 0.4
    nb = NbnxmGpu(product_params(d, elec="None"), nbat)
    try:
        f, e_lj, e_el, _ = run_step(nb, nbat, plist, energy=True, virial=True)
        check_forces(f, f_ref)
        assert e_el == 0.0
        assert abs(e_lj - e_ref[0]) <= E_REL * abs(e_ref[0]) + 2e-6
        for scalar in (False, True):
            if scalar:
                monkeypatch.setenv("NBNXM_B200_SCALAR_KERNEL", "1")
            f, _, _, _ = run_step(nb, nbat, plist, energy=False, virial=False, fresh_list=False)
            check_forces(f, f_ref)
    finally:
        nb.gpu_free()
