#!/usr/bin/env python
"""TEST INFRASTRUCTURE: tests/golden/refdata/coulombnone.npz from the reference's golden XML files for kernels
without NBNxM electrostatics (src/gromacs/nbnxm/tests/refdata/Combinations_NbnxmKernelTest_CoulombNone_Vdw*.xml,
CoulombKernelType::None in nbnxm/tests/kernel_test.cpp:305-357: the 243-atom TestSystem, rvdw = rcoulomb = the
pair-list cut-off, charges present but not interacting).  Dev container only (reads /root/reference)."""
import os

import numpy as np

from make_golden import REFDATA, parse_refdata_xml

XML_VDW = {"cutgeom": "CutCombGeom", "cutlb": "CutCombLB", "cutnone": "CutCombNone", "fswitch": "ForceSwitch",
           "pswitch": "PotSwitch", "ljpmegeom": "EwaldCombGeom"}

out = {}
for key, name in XML_VDW.items():
    f, vdw, coul = parse_refdata_xml(os.path.join(REFDATA, "Combinations_NbnxmKernelTest_CoulombNone_Vdw%s.xml" % name))
    out["f_" + key], out["vvdw_" + key], out["vcoul_" + key] = f, np.array([vdw]), np.array([coul])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refdata", "coulombnone.npz"), **out)
print({k: v.shape for k, v in out.items()})
