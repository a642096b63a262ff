#!/bin/bash
# TEST INFRASTRUCTURE, dev container only.  Asks the REFERENCE for the pair-list radii of the benchmark workloads instead
# of guessing them: a CPU build of the reference (oracle/ref_harness/build_ref.sh) runs grompp + `mdrun -nstlist 100` with
# GMX_EMULATE_GPU=1 on SPC/E water at 300 K, dt 2 fs, default verlet-buffer-tolerance, and its log states the dual
# pair-list set-up (increaseNstlist + setupDynamicPairlistPruning, nbnxm/pairlist_tuning.cpp:455-700) for the GPU list.
# Output: tests/golden/pairlist_tuning.json (committed).
set -euo pipefail
GMX=${GMX:-/tmp/gmxbuild/bin/gmx}
export GMXLIB=${GMXLIB:-/root/reference/share/top}
OUT=$(cd "$(dirname "$0")" && pwd)/pairlist_tuning.json
W=$(mktemp -d)
cd "$W"
cat > topol.top <<'TOP'
#include "oplsaa.ff/forcefield.itp"
#include "oplsaa.ff/spce.itp"
[ system ]
water
[ molecules ]
TOP
$GMX -quiet solvate -cs spc216.gro -box 6.2 6.2 6.2 -o conf.gro -p topol.top > /dev/null 2>&1
echo "{" > "$OUT"
first=1
for rc in 0.9 1.0 1.2; do
cat > g.mdp <<MDP
integrator = md
dt = 0.002
nsteps = 0
nstlist = 10
cutoff-scheme = Verlet
coulombtype = PME
rcoulomb = $rc
rvdw = $rc
tcoupl = v-rescale
tc-grps = System
tau-t = 0.1
ref-t = 300
gen-vel = yes
gen-temp = 300
constraints = h-bonds
MDP
    $GMX -quiet grompp -f g.mdp -c conf.gro -p topol.top -o t.tpr -maxwarn 5 > /dev/null 2>&1
    GMX_EMULATE_GPU=1 $GMX -quiet mdrun -s t.tpr -nstlist 100 -ntmpi 1 -ntomp 4 -nsteps 0 -g md.log > /dev/null 2>&1 || true
    outer=$(grep -A2 "Using a dual" md.log | sed -n 2p | sed 's/.*rlist \([0-9.]*\) nm.*/\1/')
    inner=$(grep -A2 "Using a dual" md.log | sed -n 3p | sed 's/.*rlist \([0-9.]*\) nm.*/\1/')
    prune=$(grep -A2 "Using a dual" md.log | sed -n 3p | sed 's/.*updated every *\([0-9]*\) steps.*/\1/')
    [ $first = 1 ] || echo "," >> "$OUT"
    first=0
    printf ' "%s": {"nstlist": 100, "rlist_outer": %s, "rlist_inner": %s, "nstlist_prune": %s}' "$rc" "$outer" "$inner" "$prune" >> "$OUT"
done
printf '\n}\n' >> "$OUT"
cat "$OUT"
rm -rf "$W"
