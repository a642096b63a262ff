#!/bin/bash
# TEST INFRASTRUCTURE. Builds the fixture generator against a CPU-only build of the
# reference. Only usable where /root/reference exists (the dev container).
#
#   REF=/root/reference BLD=/tmp/gmxbuild oracle/ref_harness/build_ref.sh
#
# If $BLD/lib/libgromacs.so is missing it is configured + built first
# (CPU-only, AVX2_256, ~6 min on 8 cores; see SURVEY.md section 0 for the flags).
# Outputs go to oracle/_ref/ (git-ignored).
set -euo pipefail
REF=${REF:-/root/reference}
BLD=${BLD:-/tmp/gmxbuild}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../_ref
mkdir -p "$OUT"
if [ ! -f "$BLD/lib/libgromacs.so" ]; then
    mkdir -p "$BLD"
    (cd "$BLD" && CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake "$REF" -G Ninja -DCMAKE_BUILD_TYPE=Release \
        -DGMX_GPU=OFF -DGMX_MPI=OFF -DGMX_THREAD_MPI=ON -DGMX_OPENMP=ON -DGMX_FFT_LIBRARY=fftpack \
        -DGMX_BUILD_OWN_FFTW=OFF -DGMX_EXTERNAL_BLAS=OFF -DGMX_EXTERNAL_LAPACK=OFF -DGMX_HWLOC=OFF \
        -DGMX_SIMD=AVX2_256 -DBUILD_TESTING=OFF -DGMX_BUILD_HELP=OFF -DGMX_INSTALL_LEGACY_API=OFF \
        -DGMXAPI=OFF -DGMX_USE_COLVARS=NONE -DGMX_USE_PLUMED=OFF -DREGRESSIONTEST_DOWNLOAD=OFF \
        && ninja -j"$(nproc)" gmx)
fi
INC="-I$REF/src/include -I$BLD/src/include -I$REF/src -I$REF/api/legacy/include -I$BLD/api/legacy/include"
for m in math timing utility pbcutil topology serialization simd taskassignment; do
    INC="$INC -I$REF/src/gromacs/$m/include"
done
INC="$INC -isystem $REF/src/external/thread_mpi/include -isystem $REF/src/external"
/usr/bin/g++ -O2 -std=c++17 -mavx2 -mfma -fopenmp -DGMX_DOUBLE=0 -DHAVE_CONFIG_H $INC \
    "$HERE/dump_nbnxm.cpp" "$REF/src/gromacs/nbnxm/tests/testsystem.cpp" \
    -L"$BLD/lib" -lgromacs -Wl,-rpath,"$BLD/lib" -o "$OUT/dump_nbnxm"
/usr/bin/g++ -O2 -std=c++17 -mavx2 -mfma -fopenmp -DGMX_DOUBLE=0 -DHAVE_CONFIG_H $INC \
    "$HERE/bench_ref.cpp" -L"$BLD/lib" -lgromacs -Wl,-rpath,'$ORIGIN/lib' -o "$OUT/bench_ref"
# the timing harness travels to the GPU box with the reference's shared libraries beside it
mkdir -p "$OUT/lib"
cp -L "$BLD/lib/libgromacs.so.12" "$BLD/lib/libmuparser.so.2" "$OUT/lib/"
strip --strip-unneeded "$OUT/lib/libgromacs.so.12" || true
echo "built $OUT/dump_nbnxm $OUT/bench_ref"
# optional second flavour of the timing harness: the same sources against an AVX_512 build of the reference
# (BLD512, configured like BLD but with -DGMX_SIMD=AVX_512); bench.py times both where the host supports AVX-512 and
# reports the faster one, as SURVEY.md section 8d asks
BLD512=${BLD512:-/tmp/gmxbuild512}
if [ -f "$BLD512/lib/libgromacs.so" ]; then
    INC512="-I$REF/src/include -I$BLD512/src/include -I$REF/src -I$REF/api/legacy/include -I$BLD512/api/legacy/include"
    for m in math timing utility pbcutil topology serialization simd taskassignment; do
        INC512="$INC512 -I$REF/src/gromacs/$m/include"
    done
    INC512="$INC512 -isystem $REF/src/external/thread_mpi/include -isystem $REF/src/external"
    /usr/bin/g++ -O2 -std=c++17 -march=skylake-avx512 -fopenmp -DGMX_DOUBLE=0 -DHAVE_CONFIG_H $INC512 \
        "$HERE/bench_ref.cpp" -L"$BLD512/lib" -lgromacs -Wl,-rpath,'$ORIGIN/lib512' -o "$OUT/bench_ref_avx512"
    mkdir -p "$OUT/lib512"
    cp -L "$BLD512/lib/libgromacs.so.12" "$BLD512/lib/libmuparser.so.2" "$OUT/lib512/"
    strip --strip-unneeded "$OUT/lib512/libgromacs.so.12" || true
    echo "built $OUT/bench_ref_avx512"
fi
