"""gromacs_b200: Blackwell-native (sm_100a) NBNXM short-range nonbonded path.

Only what the path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (include/nbnxm_b200.h) -> libnbnxm_b200.so
  nbnxm.py         host-side mirror of the reference's GPU-backend interface (nbnxm_gpu.h, gpu_data_mgmt.h)
  interaction.py   interaction_const_t / PairlistParams -> kernel parameters (initNbparam)
"""
from .nbnxm import (ALL, LOCAL, NONLOCAL, AtomData, NbnxmError, NbnxmGpu, PairlistGpu, Params, StepWorkload,  # noqa: F401
                    load_library)
from .interaction import make_params  # noqa: F401
