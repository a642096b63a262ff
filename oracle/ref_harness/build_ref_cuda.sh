#!/bin/bash
# TEST / BENCHMARK INFRASTRUCTURE.  Dev container only (needs /root/reference).
#
# Builds the reference with GMX_GPU=CUDA for sm_100 (its own CMake build, -use_fast_math, as BASELINE.md section 1
# names the competitor) and puts under oracle/_ref/cuda/ (git-ignored, travels with gpurun):
#   bench_ref_gpu            oracle/ref_harness/bench_ref_gpu.cpp, linked against libgromacs
#   lib/libgromacs.so.12     the UNMODIFIED reference, stock CUDA nbnxm backend
#   lib_shim/libgromacs.so.12  the same objects with the nbnxm GPU backend objects
#                            (nbnxm_gpu_data_mgmt.cpp, nbnxm_gpu_buffer_ops.cpp, cuda/nbnxm_cuda*.cu, cuda/nbfe_*.cu,
#                            cuda/nbnxm_gpu_buffer_ops_internal.cu) replaced by gromacs_b200/gmx_shim/nbnxm_b200_shim.cpp,
#                            linked against gromacs_b200/libnbnxm_b200.so — INTEGRATION.md step 2, executed
# Nothing of the reference's sources is copied; only built binaries land in oracle/_ref.
#
#   REF=/root/reference BLD=/tmp/gmxcuda oracle/ref_harness/build_ref_cuda.sh
set -euo pipefail
REF=${REF:-/root/reference}
BLD=${BLD:-/tmp/gmxcuda}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$HERE/../_ref/cuda
mkdir -p "$OUT/lib" "$OUT/lib_shim"
if [ ! -f "$BLD/lib/libgromacs.so" ]; then
    mkdir -p "$BLD"
    (cd "$BLD" && CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake "$REF" -G Ninja -DCMAKE_BUILD_TYPE=Release \
        -DGMX_GPU=CUDA -DCMAKE_CUDA_ARCHITECTURES=100-real -DGMX_MPI=OFF -DGMX_THREAD_MPI=ON -DGMX_OPENMP=ON \
        -DGMX_FFT_LIBRARY=fftpack -DGMX_BUILD_OWN_FFTW=OFF -DGMX_EXTERNAL_BLAS=OFF -DGMX_EXTERNAL_LAPACK=OFF \
        -DGMX_HWLOC=OFF -DGMX_SIMD=AVX2_256 -DGMX_BUILD_HELP=OFF -DGMX_INSTALL_LEGACY_API=OFF -DGMXAPI=OFF \
        -DGMX_USE_COLVARS=NONE -DGMX_USE_PLUMED=OFF -DREGRESSIONTEST_DOWNLOAD=OFF \
        && ninja -j"$(nproc)" gmx)
fi
INC="-I$REF/src/include -I$BLD/src/include -I$REF/src -I$REF/api/legacy/include -I$BLD/api/legacy/include"
for m in math timing utility pbcutil pulling topology serialization linearalgebra simd taskassignment gpu_utils hardware mdtypes; do
    [ -d "$REF/src/gromacs/$m/include" ] && INC="$INC -I$REF/src/gromacs/$m/include"
done
INC="$INC -isystem $REF/src/external/thread_mpi/include -isystem $REF/src/external -isystem /usr/local/cuda/include"
CXXFLAGS="-O2 -std=c++17 -mavx2 -mfma -fopenmp -fPIC -DGMX_DOUBLE=0 -DHAVE_CONFIG_H"

# 1. the harness; libgromacs is found next to it: LD_LIBRARY_PATH selects lib/ (stock) or lib_shim/ (drop-in)
/usr/bin/g++ $CXXFLAGS $INC -g -rdynamic "$HERE/bench_ref_gpu.cpp" -L"$BLD/lib" -lgromacs -o "$OUT/bench_ref_gpu"

# 2. the unmodified reference
cp -L "$BLD/lib/libgromacs.so.12" "$BLD/lib/libmuparser.so.2" "$OUT/lib/"
strip --strip-unneeded "$OUT/lib/libgromacs.so.12" || true

# 3. the drop-in: the reference's own link line minus its nbnxm GPU backend objects, plus the shim and libnbnxm_b200
make -s -j"$(nproc)" -C "$ROOT/gromacs_b200/csrc"
/usr/bin/g++ $CXXFLAGS -Dlibgromacs_EXPORTS $INC -I"$ROOT/include" -c "$ROOT/gromacs_b200/gmx_shim/nbnxm_b200_shim.cpp" -o "$OUT/nbnxm_b200_shim.o"
LINK=$(cd "$BLD" && ninja -t commands lib/libgromacs.so.12.0.0 | tail -1 | sed 's/^.*&& \(\/usr\/bin\/g++ .*\) && :$/\1/')
LINK=$(echo "$LINK" | tr ' ' '\n' \
    | grep -v -E 'libgromacs.dir/nbnxm/(nbnxm_gpu_data_mgmt|nbnxm_gpu_buffer_ops)\.cpp\.o$' \
    | grep -v -E 'libgromacs.dir/nbnxm/cuda/.*\.cu\.o$' \
    | grep -v -- '--dependency-file' | tr '\n' ' ')
LINK=${LINK/-o lib\/libgromacs.so.12.0.0/-o $OUT/lib_shim/libgromacs.so.12}
(cd "$BLD" && eval "$LINK" "$OUT/nbnxm_b200_shim.o" -L"$ROOT/gromacs_b200" -lnbnxm_b200 \
    "-Wl,-rpath,'\$ORIGIN/../../../../gromacs_b200'" -Wl,-z,defs)
cp -L "$BLD/lib/libmuparser.so.2" "$OUT/lib_shim/"
strip --strip-unneeded "$OUT/lib_shim/libgromacs.so.12" || true
rm -f "$OUT/nbnxm_b200_shim.o"
echo "built $OUT/bench_ref_gpu, lib/ (stock CUDA backend), lib_shim/ (nbnxm_b200 backend)"
