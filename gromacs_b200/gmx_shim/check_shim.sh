#!/bin/bash
# Syntax-checks the reference-side binding against the reference's own headers (dev container only:
# needs /root/reference and a configured build tree for config.h).  A GMX_GPU=CUDA flavour of config.h is
# derived from the CPU build's so that the GPU declarations in nbnxm_gpu.h / gpu_data_mgmt.h are the real ones.
set -euo pipefail
REF=${REF:-/root/reference}
BLD=${BLD:-/tmp/gmxbuild}
HERE=$(cd "$(dirname "$0")" && pwd)
CFG=$(mktemp -d)
sed 's/#define GMX_GPU_CUDA 0/#define GMX_GPU_CUDA 1/; s/#define GMX_GPU 0/#define GMX_GPU 1/; s/#define GMX_GPU_NB_CLUSTER_SIZE $/#define GMX_GPU_NB_CLUSTER_SIZE 8/' \
    "$BLD/src/include/config.h" > "$CFG/config.h"
INC="-I$CFG -I$REF/src/include -I$BLD/src/include -I$REF/src -I$REF/api/legacy/include -I$BLD/api/legacy/include"
for m in math timing utility pbcutil topology serialization simd taskassignment gpu_utils hardware mdtypes; do
    [ -d "$REF/src/gromacs/$m/include" ] && INC="$INC -I$REF/src/gromacs/$m/include"
done
INC="$INC -isystem $REF/src/external/thread_mpi/include -isystem $REF/src/external -I/usr/local/cuda/include -I$HERE/../../include"
/usr/bin/g++ -std=c++17 -fsyntax-only -DGMX_DOUBLE=0 -DHAVE_CONFIG_H $INC "$HERE/nbnxm_b200_shim.cpp"
rm -rf "$CFG"
echo "shim syntax check against $REF: ok"
