"""What float32 pair arithmetic can deliver on the golden fixtures: the error of the REFERENCE's own single-precision
SIMD 4xM kernel (outputs stored in tests/golden by oracle/ref_harness/dump_nbnxm.cpp) against the double-precision
oracle, quantity by quantity.  The north star asks for energies and virial within 1e-6 relative; the reference's float
kernel itself is at 2e-6 ... 4e-6 on the virial and up to 1.6e-5 on the Coulomb energy of these boxes, so the GPU tests
(tests/test_gpu_parity.py) hold the CUDA kernels to `max(1.5e-6, 1.5 x the reference SIMD kernel's own error on that
fixture)` for the virial instead of a flat number, and to 1e-6 for the energies (which the CUDA kernels accumulate in
double; measured 6e-9 ... 1.3e-6, profiles/r02d_parity_errors.jsonl)."""
import numpy as np
import pytest

from util import golden_cases, load_golden, oracle_forces, oracle_params


def virial(shift_vec, fshift):
    return -0.5 * np.einsum("si,sj->ij", shift_vec.astype(np.float64), fshift)


def reference_simd_errors(oracle, d):
    """(virial, E_lj, E_el) relative errors of the reference's SIMD float kernel against the double oracle"""
    _, fsh_ref, e_ref, _ = oracle_forces(oracle, d, oracle_params(oracle, d))
    fs = d["ref_simd4xm_fshift"].astype(np.float64).reshape(-1, 3)
    v, vr = virial(d["shift_vec"], fs), virial(d["shift_vec"], fsh_ref)
    e_lj = float(d["ref_simd4xm_vvdw"].reshape(-1)[0])
    e_el = float(d["ref_simd4xm_vcoul"].reshape(-1)[0])
    return (float(np.abs(v - vr).max() / np.abs(vr).max()), abs(e_lj - e_ref[0]) / abs(e_ref[0]),
            abs(e_el - e_ref[1]) / abs(e_ref[1]))


@pytest.mark.parametrize("case", golden_cases())
def test_reference_float_kernel_is_not_within_1e6_of_double(oracle, case):
    d = load_golden(case)
    vir, e_lj, e_el = reference_simd_errors(oracle, d)
    # sanity of the fixture: the reference kernel is a float kernel of the same physics
    assert vir < 1e-5 and e_lj < 5e-6 and e_el < 3e-5
    # ... and its float accumulation of the shift forces sits above 1e-6 on every fixture but the split RF list
    if case != "bench1_rf_cutnone_split":
        assert vir > 1e-6, vir
