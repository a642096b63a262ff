"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/nbnxm_b200.h declares; calling into it without a device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import load_golden, product_inputs, product_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="nbnxm_b200.h"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nbnxm_b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from gromacs_b200.nbnxm import EXPORTED_SYMBOLS
    assert declared_symbols() == sorted(EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from gromacs_b200 import load_library
    lib = load_library()
    for name in declared_symbols() + declared_symbols("nbnxm_b200_search.h"):
        assert hasattr(lib, name), name


def test_search_header_and_binding_agree():
    from gromacs_b200.pairsearch import SEARCH_SYMBOLS
    assert declared_symbols("nbnxm_b200_search.h") == sorted(SEARCH_SYMBOLS)


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


def test_sm100a_kernels_keep_their_register_budget():
    """The force kernels' occupancy is part of the design: the packed force-only Ewald kernels must fit 96
    registers (20 warps per SM) without spilling, the energy kernels 128; and the library must carry sm_100a code only."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "gromacs_b200", "libnbnxm_b200.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
    packed = {k: v for k, v in usage.items() if "nbnxm_force_kernel_packed" in k}
    assert len(packed) >= 4 * 6 * 2
    for k, (reg, stack) in packed.items():
        energy = "Lb1EEE" in k
        assert reg <= (128 if energy else 96), (k, reg)
        assert stack <= 32, (k, stack)
    # the headline flavors do not spill at all: Ewald analytical + LJ cut (type table) F and F+E, + force switch F+E
    for vdw, e in ((1, 0), (1, 1), (3, 1)):
        k = [n for n in packed if "packedILi4ELi%dELb%dEEE" % (vdw, e) in n]
        assert k and packed[k[0]][1] == 0, (vdw, e, k and packed[k[0]])


def test_public_headers_are_plain_c_and_link():
    """include/*.h compile as C99 (the boundary is a C ABI: no C++ types in the signatures) and a C program referencing an
    entry point of every group links against the library"""
    import shutil
    import subprocess
    import tempfile
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = """
#include "nbnxm_b200.h"
#include "nbnxm_b200_search.h"
int main(void)
{
    void* fn[] = { (void*)nbnxm_b200_init, (void*)nbnxm_b200_launch_kernel, (void*)nbnxm_b200_launch_kernel_pruneonly,
                   (void*)nbnxm_b200_pairlist_build, (void*)nbnxm_b200_gpu_search_build, (void*)nbnxm_b200_reduce_f,
                   (void*)nbnxm_b200_launch_free_energy_kernel, (void*)nbnxm_b200_chunk_plan };
    return nbnxm_b200_last_error() == 0 || fn[0] == 0;
}
"""
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "abi.c")
        with open(path, "w") as fh:
            fh.write(src)
        subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-pedantic", "-fsyntax-only",
                        "-I" + os.path.join(ROOT, "include"), path], check=True)
        lib_dir = os.path.join(ROOT, "gromacs_b200")
        subprocess.run([cc, "-std=c99", "-I" + os.path.join(ROOT, "include"), path, "-L" + lib_dir, "-l:libnbnxm_b200.so",
                        "-Wl,-rpath," + lib_dir, "-o", os.path.join(tmp, "abi")], check=True)


def test_shim_compiles_against_the_reference_headers():
    """the reference-side binding (gromacs_b200/gmx_shim/nbnxm_b200_shim.cpp) is a translation unit of the reference tree:
    syntax-checked against the reference's own headers where they and a configured build tree exist (dev container)"""
    import subprocess
    if not (os.path.isdir("/root/reference/src/gromacs") and os.path.exists("/tmp/gmxbuild/src/include/config.h")):
        pytest.skip("needs /root/reference and a configured CPU build tree of it")
    out = subprocess.run([os.path.join(ROOT, "gromacs_b200", "gmx_shim", "check_shim.sh")], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
