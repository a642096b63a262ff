"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/nbnxm_b200.h declares; calling into it without a device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import load_golden, product_inputs, product_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nbnxm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nbnxm_b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from gromacs_b200.nbnxm import EXPORTED_SYMBOLS
    assert declared_symbols() == sorted(EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from gromacs_b200 import load_library
    lib = load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_layouts_match_reference_pairlist_structs():
    # nbnxm_sci_t 16 B, nbnxm_cj_packed_t 32 B, nbnxm_excl_t 128 B (pairlist.h:189-287)
    d = load_golden("test243_ewald_cutnone")
    assert d["pl_sci"].dtype == np.int32 and d["pl_sci"].shape[1] * 4 == 16
    assert d["pl_cjPacked"].shape[1] * 4 == 32 and d["pl_excl"].shape[1] * 4 == 128
    from gromacs_b200.nbnxm import Params
    assert C.sizeof(Params) == 25 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gromacs_b200 import NbnxmError, NbnxmGpu
    d = load_golden("test243_ewald_cutnone")
    nbat, _ = product_inputs(d)
    with pytest.raises(NbnxmError, match="no CUDA device"):
        NbnxmGpu(product_params(d), nbat)
