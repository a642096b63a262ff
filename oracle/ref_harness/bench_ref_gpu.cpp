/* TEST / BENCHMARK INFRASTRUCTURE — drives the reference's GPU nonbonded path through the reference's own callers
 * (nonbonded_verlet_t::putAtomsOnGrid / constructPairlist / dispatchNonbondedKernel / dispatchPruneKernelGpu and the
 * gmx::gpu_* boundary functions, in the order do_force calls them, src/gromacs/mdlib/sim_util.cpp:863-2482) on
 * BenchmarkSystem(size), the water box of `gmx nonbonded-benchmark`.
 *
 * The program is linked against a GMX_GPU=CUDA build of libgromacs.  Which backend sits behind gmx::gpu_* is decided
 * by the libgromacs.so that is loaded at run time:
 *   oracle/_ref/cuda/lib       the UNMODIFIED reference (its CUDA kernels compiled for sm_100) — the competitor number
 *   oracle/_ref/cuda/lib_shim  the same build with the backend objects replaced by gromacs_b200/gmx_shim/nbnxm_b200_shim.cpp
 *                              + libnbnxm_b200.so — the drop-in, linked and executed
 * The set-up follows api/nblib/nbnxmsetuphelpers.cpp:342-400 (createNbnxmGPU) with a complete interaction_const_t
 * (force switch, potential switch and LJ-PME selectable) and the dynamic-pruning fields of PairlistParams filled the
 * way pairlist_tuning.cpp:685 does.  Prints one JSON line; --dump writes the atom-order forces (float32 N x 3).
 */
#include "gmxpre.h"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "gromacs/gpu_utils/device_stream_manager.h"
#include "gromacs/gpu_utils/gpu_utils.h"
#include "gromacs/gmxlib/nrnb.h"
#include "gromacs/gpu_utils/hostallocator.h"
#include "gromacs/hardware/device_information.h"
#include "gromacs/hardware/device_management.h"
#include "gromacs/mdlib/forcerec.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/inputrec.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/locality.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/benchmark/bench_system.h"
#include "gromacs/nbnxm/gpu_data_mgmt.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_geometry.h"
#include "gromacs/nbnxm/nbnxm_gpu.h"
#include "gromacs/nbnxm/pairlistparams.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/timing/gpu_timing.h"
#include "gromacs/topology/topology.h"
#include "gromacs/utility/logger.h"

using namespace gmx;

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void onCrash(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "bench_ref_gpu: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

#define STAGE(name) std::fprintf(stderr, "[bench_ref_gpu] %s\n", name)

int main(int argc, char** argv)
{
    signal(SIGSEGV, onCrash);
    signal(SIGABRT, onCrash);
    int         size = 1, nt = 1, iters = 10, warmup = 2, energy = 0, nstlistPrune = 6, dynamicPruning = 0;
    double      rc = 0.9, rlistOuter = 0, rlistInner = 0;
    std::string vdw = "cut", dump;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--size") size = std::atoi(v.c_str());
        else if (k == "--nt") nt = std::atoi(v.c_str());
        else if (k == "--iter") iters = std::atoi(v.c_str());
        else if (k == "--warmup") warmup = std::atoi(v.c_str());
        else if (k == "--energy") energy = std::atoi(v.c_str());
        else if (k == "--rc") rc = std::atof(v.c_str());
        else if (k == "--rlist-outer") rlistOuter = std::atof(v.c_str());
        else if (k == "--rlist-inner") rlistInner = std::atof(v.c_str());
        else if (k == "--nstlist-prune") nstlistPrune = std::atoi(v.c_str());
        else if (k == "--dynamic-pruning") dynamicPruning = std::atoi(v.c_str());
        else if (k == "--vdw") vdw = v;
        else if (k == "--dump") dump = v;
    }
    if (rlistOuter <= 0) rlistOuter = rc;
    if (rlistInner <= 0) rlistInner = rlistOuter;
    gmx_omp_nthreads_set(ModuleMultiThread::Pairsearch, nt);
    gmx_omp_nthreads_set(ModuleMultiThread::Nonbonded, nt);

    STAGE("system");
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

    STAGE("devices");
    /* the device, as the test hardware environment sets it up (src/testutils/test_hardware_environment.cpp:97) */
    std::string detectionError;
    if (!canPerformDeviceDetection(&detectionError))
    {
        std::printf("{\"error\": \"device detection not possible: %s\"}\n", detectionError.c_str());
        return 1;
    }
    std::vector<std::unique_ptr<DeviceInformation>> devices = findDevices();
    if (devices.empty())
    {
        std::printf("{\"error\": \"no GPU found\"}\n");
        return 1;
    }
    const DeviceInformation& deviceInfo = *devices[0];
    setActiveDevice(deviceInfo);
    SimulationWorkload simulationWork;
    simulationWork.computeNonbonded = true;
    simulationWork.useGpuNonbonded  = true;
    DeviceStreamManager deviceStreamManager(deviceInfo, simulationWork, false);

    STAGE("atomdata + gpu_init");
    const HostAllocationPolicy pol{ deviceStreamManager.context(), PinningPolicy::PinnedIfSupported };
    NbnxmKernelSetup           ks;
    ks.kernelType         = NbnxmKernelType::Gpu8x8x8;
    ks.ewaldExclusionType = EwaldExclusionType::Analytical;
    PairlistParams plp(ks.kernelType, PairlistType::Hierarchical8x8x8, false, rlistOuter, false);
    plp.lifetime = 99;
    /* not set by the constructor (pairlistparams.cpp:56-70); init_nb_verlet assigns it (nbnxm_setup.cpp:519) */
    plp.haveNonbondedFEGpu_ = false;
    if (dynamicPruning)
    {
        /* pairlist_tuning.cpp:570-590, :685: inner radius, pruning interval, rolling parts = nstlistPrune / 2 */
        plp.useDynamicPruning      = true;
        plp.rlistInner             = rlistInner;
        plp.nstlistPrune           = nstlistPrune;
        plp.numRollingPruningParts = nstlistPrune / 2;
    }
    const bool ljpme = (vdw == "ljpme");
    auto       nbat  = std::make_unique<nbnxm_atomdata_t>(pol, MDLogger(), ks.kernelType,
                                                   (vdw == "cut")     ? gmx::LJCombinationRule::Geometric
                                                   : (vdw == "cutlb") ? gmx::LJCombinationRule::LorentzBerthelot
                                                                      : gmx::LJCombinationRule::None,
                                                   ljpme ? gmx::LJCombinationRule::Geometric : gmx::LJCombinationRule::None,
                                                   sys.nonbondedParameters, true, 1, 1);
    NbnxmGpu*  nbnxmGpu = gpu_init(deviceStreamManager, &ic, plp, nbat.get(), false, std::nullopt);
    auto       sets     = std::make_unique<PairlistSets>(plp, false, gpu_min_ci_balanced(nbnxmGpu), pol);
    auto       search   = std::make_unique<PairSearch>(PbcType::Xyz, false, nullptr, nullptr, plp.pairlistType, false, false, nt, pol);
    auto       nbv      = std::make_unique<nonbonded_verlet_t>(std::move(sets), std::move(search), std::move(nbat), ks, nbnxmGpu);

    STAGE("search step");
    /* the search step (sim_util.cpp:1388-1500) */
    const rvec lo = { 0, 0, 0 };
    const rvec hi = { sys.box[XX][XX], sys.box[YY][YY], sys.box[ZZ][ZZ] };
    double     t0 = now();
    nbv->putAtomsOnGrid(sys.box, 0, lo, hi, nullptr, { 0, int(sys.coordinates.size()) }, sys.coordinates.size(),
                        sys.coordinates.size() / det(sys.box), sys.atomInfoAllVdw, sys.coordinates, nullptr);
    const double tGrid = now() - t0;
    nbv->setAtomProperties(sys.atomTypes, sys.charges, sys.atomInfoAllVdw);
    gpu_init_atomdata(nbv->gpuNbv(), &nbv->nbat());
    t0 = now();
    t_nrnb nrnb; /* the GPU branch of constructPairlist counts the search in it */
    nbv->constructPairlist(InteractionLocality::Local, sys.excls, false, 0, &nrnb);
    const double tList = now() - t0;
    nbv->setupGpuShortRangeWork(nullptr, InteractionLocality::Local);
    /* the periodic shift vectors into the atom data (do_force: nbnxm_atomdata_copy_shiftvec, sim_util.cpp), then to the device */
    nbnxm_atomdata_copy_shiftvec(false, sys.forceRec.shift_vec, &nbv->nbat());
    gpu_upload_shiftvec(nbv->gpuNbv(), &nbv->nbat());

    STAGE("steps");
    StepWorkload sw;
    sw.computeForces          = true;
    sw.computeNonbondedForces = true;
    sw.computeEnergy          = energy != 0;
    sw.computeVirial          = energy != 0;
    std::vector<real> vVdw(1, 0), vCoul(1, 0);
    std::vector<RVec> fshift(c_numShiftVectors, RVec{ 0, 0, 0 });
    real              eLJ = 0, eEl = 0;

    auto doStep = [&](int64_t step) {
        /* sim_util.cpp order: rolling prune (launchGpuEndOfStepTasks of the previous step, :972) -> copy xq -> clear ->
         * kernel -> copy back -> wait */
        if (step > 0 && nbv->isDynamicPruningStepGpu(step))
        {
            nbv->dispatchPruneKernelGpu(step);
        }
        gpu_clear_outputs(nbv->gpuNbv(), sw.computeVirial);
        gpu_copy_xq_to_gpu(nbv->gpuNbv(), &nbv->nbat(), AtomLocality::Local);
        nbv->dispatchNonbondedKernel(InteractionLocality::Local, ic, sw, enbvClearFNo, sys.forceRec.shift_vec, vVdw, vCoul, nullptr);
        gpu_launch_cpyback(nbv->gpuNbv(), &nbv->nbat(), sw, AtomLocality::Local);
        eLJ = 0;
        eEl = 0;
        for (auto& v : fshift)
        {
            v = { 0, 0, 0 };
        }
        double dvdlLj = 0, dvdlEl = 0; /* gpu_reduce_staged_outputs adds to them on every energy step (gpu_common.h:151-161) */
        gpu_try_finish_task(nbv->gpuNbv(), sw, AtomLocality::Local, &eLJ, &eEl, &dvdlLj, &dvdlEl, fshift, nullptr, GpuTaskCompletion::Wait);
    };

    int64_t step = 0;
    for (int i = 0; i < warmup; i++)
    {
        doStep(step++);
    }
    gpu_reset_timings(nbv.get());
    t0 = now();
    for (int i = 0; i < iters; i++)
    {
        doStep(step++);
    }
    const double sec = now() - t0;

    STAGE("timings");
    /* kernel times from the backend's own timers (GMX_ENABLE_GPU_TIMING=1 for the stock CUDA backend) */
    double kForce = 0, kPrune = 0, kRoll = 0, h2d = 0, d2h = 0;
    int    cForce = 0, cPrune = 0, cRoll = 0;
    if (gmx_wallclock_gpu_nbnxm_t* t = gpu_get_timings(nbv->gpuNbv()))
    {
        for (int p = 0; p < 2; p++)
        {
            for (int e = 0; e < 2; e++)
            {
                kForce += t->ktime[p][e].t;
                cForce += t->ktime[p][e].c;
            }
        }
        kPrune = t->pruneTime.t;
        cPrune = t->pruneTime.c;
        kRoll  = t->dynamicPruneTime.t;
        cRoll  = t->dynamicPruneTime.c;
        h2d    = t->nb_h2d_t;
        d2h    = t->nb_d2h_t;
    }

    STAGE("forces to atom order");
    /* forces back to atom order (kernel_test.cpp:628-661) */
    std::vector<RVec> f(sys.coordinates.size(), RVec{ 0, 0, 0 });
    nbv->atomdata_add_nbat_f_to_f(AtomLocality::Local, f);
    double fsum2 = 0, fs[3] = { 0, 0, 0 };
    for (const RVec& v : f)
    {
        fsum2 += double(v[0]) * v[0] + double(v[1]) * v[1] + double(v[2]) * v[2];
        fs[0] += v[0];
        fs[1] += v[1];
        fs[2] += v[2];
    }
    double vir = 0; /* -1/2 sum shift . fshift, the part of the virial this path owns */
    for (int s = 0; s < c_numShiftVectors; s++)
    {
        vir += -0.5 * (double(sys.forceRec.shift_vec[s][0]) * fshift[s][0] + double(sys.forceRec.shift_vec[s][1]) * fshift[s][1]
                       + double(sys.forceRec.shift_vec[s][2]) * fshift[s][2]);
    }
    if (!dump.empty())
    {
        if (FILE* fp = std::fopen(dump.c_str(), "wb"))
        {
            std::fwrite(f.data(), sizeof(RVec), f.size(), fp);
            std::fwrite(fshift.data(), sizeof(RVec), fshift.size(), fp);
            std::fclose(fp);
        }
    }
    const double n       = double(sys.coordinates.size());
    const double density = n / det(sys.box);
    const double useful  = n * 0.5 * (density * 4.0 / 3.0 * M_PI * rc * rc * rc + 1.0);
    std::printf("{\"natoms\": %.0f, \"vdw\": \"%s\", \"energy\": %d, \"rc\": %.3f, \"rlist_outer\": %.3f, \"rlist_inner\": %.3f, "
                "\"dynamic_pruning\": %d, \"iters\": %d, \"sec_per_step_host_buffers\": %.6e, \"useful_pairs\": %.6e, "
                "\"force_kernel_ms\": %.6f, \"force_kernel_count\": %d, \"first_prune_ms\": %.6f, \"first_prune_count\": %d, "
                "\"rolling_prune_ms\": %.6f, \"rolling_prune_count\": %d, \"h2d_ms\": %.6f, \"d2h_ms\": %.6f, "
                "\"host_grid_s\": %.4f, \"host_list_s\": %.4f, \"e_lj\": %.9e, \"e_el\": %.9e, \"virial_shift_part\": %.9e, "
                "\"f_rms\": %.9e, \"f_sum\": [%.4e, %.4e, %.4e], \"device\": \"%s\"}\n",
                n, vdw.c_str(), energy, rc, rlistOuter, rlistInner, dynamicPruning, iters, sec / iters, useful,
                cForce ? kForce / cForce : 0.0, cForce, cPrune ? kPrune / cPrune : 0.0, cPrune, cRoll ? kRoll / cRoll : 0.0, cRoll,
                iters ? h2d / iters : 0.0, iters ? d2h / iters : 0.0, tGrid, tList, double(eLJ), double(eEl), vir,
                std::sqrt(fsum2 / n), fs[0], fs[1], fs[2], deviceInfo.prop.name);
    return 0;
}
