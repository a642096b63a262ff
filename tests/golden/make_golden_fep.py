#!/usr/bin/env python
"""TEST INFRASTRUCTURE: tests/golden/refdata/fep_gpu.npz from the reference's golden XML files of its GPU perturbed-pair
kernel test (src/gromacs/nbnxm/tests/refdata/NBInteraction_NonbondedFepGpuTest_testGpuKernel_*.xml, 312 files:
nbnxm/tests/freeenergygpukernel.cpp:690-840).  Dev container only (reads /root/reference)."""
import glob
import os
import re

import numpy as np

REFDATA = "/root/reference/src/gromacs/nbnxm/tests/refdata"
PREFIX = "NBInteraction_NonbondedFepGpuTest_testGpuKernel_softcore_beutler_list_"

names, rows = [], []
for path in sorted(glob.glob(os.path.join(REFDATA, PREFIX + "*.xml"))):
    txt = open(path).read()
    real = lambda key: float(re.search(r'<Real Name="%s">([^<]+)</Real>' % key, txt).group(1))
    forces = txt.split('<Sequence Name="Forces">')[1].split("</Sequence>")[0]
    f = [float(v) for v in re.findall(r'<Real Name="[XYZ]">([^<]+)</Real>', forces)]
    sh = txt.split("<Shift-Forces")[1]
    fs = [float(v) for v in re.findall(r'<Real Name="[XYZ]">([^<]+)</Real>', sh)]
    assert len(f) == 12 and len(fs) == 3
    names.append(os.path.basename(path)[len(PREFIX):-4])
    rows.append([real("EVdw "), real("ECoul "), real("dVdlCoul "), real("dVdlVdw ")] + f + fs)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refdata", "fep_gpu.npz")
np.savez_compressed(out, names=np.array(names), values=np.array(rows, np.float64))
print(len(names), "cases ->", out)
