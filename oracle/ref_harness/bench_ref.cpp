/* TEST / BENCHMARK INFRASTRUCTURE — times the UNMODIFIED reference's own CPU SIMD nonbonded kernel
 * (Cpu4xN_Simd_4xN via nonbonded_verlet_t::dispatchNonbondedKernel, exactly what
 * `gmx nonbonded-benchmark` times, src/gromacs/nbnxm/benchmark/bench_setup.cpp:408-440) on
 * BenchmarkSystem(size), but with a complete interaction_const_t so that LJ force-switch,
 * potential-switch and LJ-PME can be selected (the stock tool cannot, bench_setup.cpp:183-188).
 * Links libgromacs.so of a CPU-only reference build; prints one JSON line.
 */
#include "gmxpre.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "gromacs/gpu_utils/hostallocator.h"
#include "gromacs/mdlib/forcerec.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/inputrec.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/locality.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/benchmark/bench_system.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_geometry.h"
#include "gromacs/nbnxm/pairlistparams.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/topology/topology.h"
#include "gromacs/utility/logger.h"

using namespace gmx;

int main(int argc, char** argv)
{
    int         size = 1, nt = 1, iters = 10, warmup = 2, energy = 0;
    double      rc = 0.9;
    std::string vdw = "cut", kernel = "4xm";
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--size") size = std::atoi(v.c_str());
        else if (k == "--nt") nt = std::atoi(v.c_str());
        else if (k == "--iter") iters = std::atoi(v.c_str());
        else if (k == "--warmup") warmup = std::atoi(v.c_str());
        else if (k == "--energy") energy = std::atoi(v.c_str());
        else if (k == "--rc") rc = std::atof(v.c_str());
        else if (k == "--vdw") vdw = v;
        else if (k == "--kernel") kernel = v;
    }
    gmx_omp_nthreads_set(ModuleMultiThread::Pairsearch, nt);
    gmx_omp_nthreads_set(ModuleMultiThread::Nonbonded, nt);

    BenchmarkSystem sys(size, "");

    t_inputrec ir;
    ir.vdwtype      = (vdw == "ljpme") ? VanDerWaalsType::Pme : VanDerWaalsType::Cut;
    ir.vdw_modifier = (vdw == "fswitch")   ? InteractionModifiers::ForceSwitch
                      : (vdw == "pswitch") ? InteractionModifiers::PotSwitch
                                           : InteractionModifiers::PotShift;
    ir.rvdw         = rc;
    ir.rvdw_switch  = rc - 0.2;
    if (vdw == "ljpme")
    {
        ir.ljpme_combination_rule = LongRangeVdW::Geom;
        ir.ewald_rtol_lj          = 1e-3;
    }
    ir.coulombtype      = CoulombInteractionType::Pme;
    ir.coulomb_modifier = InteractionModifiers::PotShift;
    ir.rcoulomb         = rc;
    ir.ewald_rtol       = 1e-5;
    ir.epsilon_r        = 1;
    ir.epsilon_rf       = 0;
    gmx_mtop_t mtop;
    mtop.ffparams.reppow = 12;
    mtop.ffparams.functype.resize(1);
    mtop.ffparams.functype[0] = InteractionFunction::LennardJonesShortRange;
    interaction_const_t ic    = init_interaction_const(nullptr, ir, mtop, false, std::nullopt);
    init_interaction_const_tables(nullptr, &ic, rc, 0);

    const HostAllocationPolicy pol{};
    NbnxmKernelSetup           ks;
    ks.kernelType         = (kernel == "2xmm") ? NbnxmKernelType::Cpu4xN_Simd_2xNN : NbnxmKernelType::Cpu4xN_Simd_4xN;
    ks.ewaldExclusionType = EwaldExclusionType::Analytical;
    PairlistParams plp(ks.kernelType, {}, false, rc, false);
    auto           sets   = std::make_unique<PairlistSets>(plp, false, 0, pol);
    auto           search = std::make_unique<PairSearch>(PbcType::Xyz, false, nullptr, nullptr, plp.pairlistType, false, false, nt, pol);
    const bool     ljpme  = (vdw == "ljpme");
    auto           nbat   = std::make_unique<nbnxm_atomdata_t>(pol, MDLogger(), ks.kernelType,
                                                   (vdw == "cut") ? gmx::LJCombinationRule::Geometric : gmx::LJCombinationRule::None,
                                                   ljpme ? gmx::LJCombinationRule::Geometric : gmx::LJCombinationRule::None,
                                                   sys.nonbondedParameters, true, 1, nt);
    auto nbv = std::make_unique<nonbonded_verlet_t>(std::move(sets), std::move(search), std::move(nbat), ks, nullptr);
    const rvec lo = { 0, 0, 0 };
    const rvec hi = { sys.box[XX][XX], sys.box[YY][YY], sys.box[ZZ][ZZ] };
    nbv->putAtomsOnGrid(sys.box, 0, lo, hi, nullptr, { 0, int(sys.coordinates.size()) }, sys.coordinates.size(),
                        sys.coordinates.size() / det(sys.box), sys.atomInfoAllVdw, sys.coordinates, nullptr);
    nbv->constructPairlist(InteractionLocality::Local, sys.excls, false, 0, nullptr);
    nbv->setAtomProperties(sys.atomTypes, sys.charges, sys.atomInfoAllVdw);

    StepWorkload sw;
    sw.computeForces = true;
    sw.computeEnergy = energy != 0;
    sw.computeVirial = energy != 0;
    std::vector<real> vVdw(1, 0), vCoul(1, 0);
    for (int i = 0; i < warmup; i++)
    {
        nbv->dispatchNonbondedKernel(InteractionLocality::Local, ic, sw, enbvClearFYes, sys.forceRec.shift_vec, vVdw, vCoul, nullptr);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++)
    {
        nbv->dispatchNonbondedKernel(InteractionLocality::Local, ic, sw, enbvClearFNo, sys.forceRec.shift_vec, vVdw, vCoul, nullptr);
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const PairlistSet& ps       = nbv->pairlistSets().pairlistSet(InteractionLocality::Local);
    const double       listPairs = double(ps.natpair_ljq_) + double(ps.natpair_lj_) + double(ps.natpair_q_);
    const double       n         = double(sys.coordinates.size());
    const double       density   = n / det(sys.box);
    const double       useful    = n * 0.5 * (density * 4.0 / 3.0 * M_PI * rc * rc * rc + 1.0);
    std::printf("{\"natoms\": %.0f, \"threads\": %d, \"iters\": %d, \"sec_per_iter\": %.6e, \"useful_pairs\": %.6e, "
                "\"list_pairs\": %.6e, \"kernel\": \"%s\", \"vdw\": \"%s\", \"energy\": %d, \"rc\": %.3f}\n",
                n, nt, iters, sec / iters, useful, listPairs, kernel.c_str(), vdw.c_str(), energy, rc);
    return 0;
}
