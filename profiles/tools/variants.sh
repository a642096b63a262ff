set -e
mkdir -p gpurun_out
run() { # name extra
  make -s -j32 -C gromacs_b200/csrc EXTRA="$2" >/dev/null 2>&1 || { echo "build failed $1"; return; }
  python bench.py --steps 10 --warmup 3 --workload water1536k --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/var_$1.json
  python -c "
import json;d=json.load(open('gpurun_out/var_$1.json'));print('$1','kernel_us',round(d['roofline']['kernel_us'],1),'frac',round(d['roofline']['frac'],3))"
}
touch gromacs_b200/csrc/*.cuh
run fi1_mb14 "-DNBNXM_PACKED_FI=1 -DNBNXM_PACKED_MIN_BLOCKS=14"
touch gromacs_b200/csrc/*.cuh
run fi1_mb18 "-DNBNXM_PACKED_FI=1 -DNBNXM_PACKED_MIN_BLOCKS=18"
touch gromacs_b200/csrc/*.cuh
run fi0_mb14 "-DNBNXM_PACKED_FI=0 -DNBNXM_PACKED_MIN_BLOCKS=14"
touch gromacs_b200/csrc/*.cuh
run fi0_mb18 "-DNBNXM_PACKED_FI=0 -DNBNXM_PACKED_MIN_BLOCKS=18"
touch gromacs_b200/csrc/*.cuh
run fi0_mb21 "-DNBNXM_PACKED_FI=0 -DNBNXM_PACKED_MIN_BLOCKS=21"
touch gromacs_b200/csrc/*.cuh
run fi1_mb12 "-DNBNXM_PACKED_FI=1 -DNBNXM_PACKED_MIN_BLOCKS=12"
