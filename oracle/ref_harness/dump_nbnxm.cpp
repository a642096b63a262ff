/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * Fixture generator that links against a CPU-only build of the reference
 * (libgromacs.so, see build_ref.sh) and dumps, for one synthetic system:
 *   - the GPU-layout pair list the reference builds (sci / cjPacked / excl,
 *     nbnxm/pairlist.h:189-287) and the matching nbnxm_atomdata_t contents
 *     (xq, types, lj_comb, nbfp, nbfp_comb, shift_vec; nbnxm/atomdata.h:184-375),
 *   - the interaction constants the GPU kernels consume
 *     (mdtypes/interaction_const.h:109),
 *   - the forces / energies / shift forces the reference's own SIMD 4xM kernel
 *     (nbnxm/simd_kernel.h) and plain-C GPU-layout kernel
 *     (nbnxm/kernels_reference/kernel_gpu_ref.cpp:64) compute for it.
 *
 * The call sequence follows nbnxm/benchmark/bench_setup.cpp:224-294 and
 * nbnxm/tests/kernel_test.cpp:194-266,314-364,614-661 (no code is taken from
 * there; this file only *calls* the reference's public classes).
 *
 * Output: a flat tagged binary ("NBXD"), converted to .npz by
 * tests/golden/make_golden.py.
 */
#include "gmxpre.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "gromacs/gpu_utils/hostallocator.h"
#include "gromacs/math/functions.h"
#include "gromacs/math/units.h"
#include "gromacs/mdlib/forcerec.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/atominfo.h"
#include "gromacs/mdtypes/inputrec.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/locality.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/benchmark/bench_system.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_geometry.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/nbnxm/pairlistparams.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/nbnxm/tests/testsystem.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/topology/topology.h"
#include "gromacs/utility/logger.h"
#include "gromacs/utility/vec.h"

using namespace gmx;

namespace
{

struct Writer
{
    FILE* fp;
    explicit Writer(const char* fn) : fp(std::fopen(fn, "wb"))
    {
        if (!fp)
        {
            std::perror(fn);
            std::exit(1);
        }
        std::fwrite("NBXD", 1, 4, fp);
    }
    ~Writer() { std::fclose(fp); }
    // dtype: 0 = f32, 1 = i32, 2 = u32, 3 = f64
    void put(const char* name, int dtype, const void* data, size_t n0, size_t n1 = 1)
    {
        char nm[32] = { 0 };
        std::strncpy(nm, name, 31);
        std::fwrite(nm, 1, 32, fp);
        int32_t  hdr[2] = { dtype, 0 };
        uint64_t dims[2] = { n0, n1 };
        std::fwrite(hdr, 4, 2, fp);
        std::fwrite(dims, 8, 2, fp);
        const size_t es = (dtype == 3 ? 8 : 4);
        if (n0 * n1 > 0)
        {
            std::fwrite(data, es, n0 * n1, fp);
        }
    }
    void putf(const char* name, double v)
    {
        put(name, 3, &v, 1);
    }
};

struct Sys
{
    int                  numAtomTypes;
    std::vector<real>    nbfp;
    std::vector<int>     atomTypes;
    std::vector<real>    charges;
    std::vector<int32_t> atomInfo;
    ListOfLists<int>     excls;
    std::vector<RVec>    x;
    matrix               box;
};

struct Opt
{
    std::string system   = "test243";
    std::string coulomb  = "ewald";   // ewald | ewaldtwin | rf
    std::string vdw      = "cutgeom"; // cutgeom cutlb cutnone fswitch pswitch ljpmegeom
    real        rc       = 0.9;
    real        rlist    = 0.9;
    int         minSci   = 0;
    int         nthreads = 1;
    std::string out      = "dump.bin";
};

interaction_const_t makeIc(const Opt& o)
{
    t_inputrec ir;
    const bool ljpme = (o.vdw == "ljpmegeom");
    ir.vdwtype       = ljpme ? VanDerWaalsType::Pme : VanDerWaalsType::Cut;
    ir.vdw_modifier  = (o.vdw == "fswitch")   ? InteractionModifiers::ForceSwitch
                       : (o.vdw == "pswitch") ? InteractionModifiers::PotSwitch
                                              : InteractionModifiers::PotShift;
    ir.rvdw          = (o.coulomb == "ewaldtwin") ? o.rc - 0.2 : o.rc;
    ir.rvdw_switch   = ir.rvdw - 0.2;
    if (ljpme)
    {
        ir.ljpme_combination_rule = LongRangeVdW::Geom;
        ir.ewald_rtol_lj          = 1e-4;
    }
    ir.coulombtype = (o.coulomb == "rf") ? CoulombInteractionType::RF : CoulombInteractionType::Pme;
    ir.coulomb_modifier = InteractionModifiers::PotShift;
    ir.rcoulomb         = o.rc;
    ir.ewald_rtol       = (o.system == "test243") ? 1e-6 : 1e-5;
    ir.epsilon_r        = 1;
    ir.epsilon_rf       = 0;

    gmx_mtop_t mtop;
    mtop.ffparams.reppow = 12;
    mtop.ffparams.functype.resize(1);
    mtop.ffparams.functype[0] = InteractionFunction::LennardJonesShortRange;

    interaction_const_t ic = init_interaction_const(nullptr, ir, mtop, false, std::nullopt);
    init_interaction_const_tables(nullptr, &ic, o.rc, 0);
    return ic;
}

gmx::LJCombinationRule pairCombRule(const Opt& o)
{
    if (o.vdw == "cutgeom")
    {
        return gmx::LJCombinationRule::Geometric;
    }
    if (o.vdw == "cutlb")
    {
        return gmx::LJCombinationRule::LorentzBerthelot;
    }
    return gmx::LJCombinationRule::None;
}

std::unique_ptr<nonbonded_verlet_t> makeNbv(const Opt& o, const Sys& s, NbnxmKernelType kt, real rlist, int minSci)
{
    gmx_omp_nthreads_set(ModuleMultiThread::Pairsearch, o.nthreads);
    gmx_omp_nthreads_set(ModuleMultiThread::Nonbonded, o.nthreads);
    const HostAllocationPolicy pol{};

    NbnxmKernelSetup ks;
    ks.kernelType         = kt;
    ks.ewaldExclusionType = kernelTypeIsPlainC(kt) ? EwaldExclusionType::Table : EwaldExclusionType::Analytical;

    PairlistParams plp(kt, sc_layoutType, false, rlist, false);
    auto           sets   = std::make_unique<PairlistSets>(plp, false, minSci, pol);
    auto           search = std::make_unique<PairSearch>(
            PbcType::Xyz, false, nullptr, nullptr, plp.pairlistType, false, false, o.nthreads, pol);
    const bool ljpme = (o.vdw == "ljpmegeom");
    auto       nbat  = std::make_unique<nbnxm_atomdata_t>(pol,
                                                   MDLogger(),
                                                   kt,
                                                   ljpme ? gmx::LJCombinationRule::None : pairCombRule(o),
                                                   ljpme ? gmx::LJCombinationRule::Geometric : gmx::LJCombinationRule::None,
                                                   s.nbfp,
                                                   true,
                                                   1,
                                                   o.nthreads);
    auto nbv = std::make_unique<nonbonded_verlet_t>(std::move(sets), std::move(search), std::move(nbat), ks, nullptr);

    const rvec lo = { 0, 0, 0 };
    const rvec hi = { s.box[XX][XX], s.box[YY][YY], s.box[ZZ][ZZ] };
    nbv->putAtomsOnGrid(s.box, 0, lo, hi, nullptr, { 0, int(s.x.size()) }, s.x.size(),
                        s.x.size() / det(s.box), s.atomInfo, s.x, nullptr);
    nbv->constructPairlist(InteractionLocality::Local, s.excls, false, 0, nullptr);
    nbv->setAtomProperties(s.atomTypes, s.charges, s.atomInfo);
    return nbv;
}

void runAndDump(Writer& w, const char* tag, nonbonded_verlet_t& nbv, const interaction_const_t& ic, const Sys& s, bool gpuLayout)
{
    std::vector<RVec> shiftVecs(c_numShiftVectors);
    calc_shifts(s.box, shiftVecs);
    StepWorkload sw;
    sw.computeForces = true;
    sw.computeEnergy = true;
    sw.computeVirial = true;
    std::vector<real> vVdw(1, 0), vCoul(1, 0);
    nbv.dispatchNonbondedKernel(InteractionLocality::Local, ic, sw, enbvClearFYes, shiftVecs, vVdw, vCoul, nullptr);
    std::vector<RVec> f(s.x.size(), { 0, 0, 0 });
    if (gpuLayout)
    {
        // nonbonded_verlet_t::atomdata_add_nbat_f_to_f (nbnxm.cpp:190-198) skips non-simple lists when there is no
        // GPU object, so map nbat order -> atom order here with the grid's atom index array.
        auto        order = nbv.getLocalAtomOrder();
        const auto& fo    = nbv.nbat().outputBuffer(0).f;
        for (size_t i = 0; i < order.size(); i++)
        {
            if (order[i] >= 0)
            {
                for (int d = 0; d < 3; d++) f[order[i]][d] += fo[3 * i + d];
            }
        }
    }
    else
    {
        nbv.atomdata_add_nbat_f_to_f(AtomLocality::All, f);
    }
    std::vector<float> ff(3 * f.size());
    for (size_t i = 0; i < f.size(); i++)
    {
        for (int d = 0; d < 3; d++)
        {
            ff[3 * i + d] = f[i][d];
        }
    }
    std::string n = std::string(tag);
    w.put((n + "_f").c_str(), 0, ff.data(), f.size(), 3);
    const auto&        out = nbv.nbat().outputBuffer(0);
    std::vector<float> fs(out.fshift.begin(), out.fshift.end());
    w.put((n + "_fshift").c_str(), 0, fs.data(), fs.size() / 3, 3);
    w.putf((n + "_vvdw").c_str(), vVdw[0]);
    w.putf((n + "_vcoul").c_str(), vCoul[0]);
}

} // namespace

int main(int argc, char** argv)
{
    Opt o;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--system") o.system = v;
        else if (k == "--coulomb") o.coulomb = v;
        else if (k == "--vdw") o.vdw = v;
        else if (k == "--rc") o.rc = std::atof(v.c_str());
        else if (k == "--rlist") o.rlist = std::atof(v.c_str());
        else if (k == "--minsci") o.minSci = std::atoi(v.c_str());
        else if (k == "--nt") o.nthreads = std::atoi(v.c_str());
        else if (k == "--out") o.out = v;
        else
        {
            std::fprintf(stderr, "unknown option %s\n", k.c_str());
            return 2;
        }
    }

    Sys s;
    if (o.system == "test243")
    {
        // nbnxm/tests/kernel_test.cpp:560-562: geometric params only for CutCombGeom
        test::TestSystem ts(o.vdw == "cutgeom" ? gmx::LJCombinationRule::Geometric : gmx::LJCombinationRule::LorentzBerthelot,
                            true);
        s.numAtomTypes = ts.numAtomTypes;
        s.nbfp         = ts.nonbondedParameters;
        s.atomTypes    = ts.atomTypes;
        s.charges      = ts.charges;
        s.atomInfo     = ts.atomInfo;
        s.excls        = ts.excls;
        s.x            = ts.coordinates;
        copy_mat(ts.box, s.box);
    }
    else if (o.system.rfind("bench", 0) == 0)
    {
        const int       k = std::atoi(o.system.c_str() + 5);
        BenchmarkSystem bs(k, "");
        s.numAtomTypes = bs.numAtomTypes;
        s.nbfp         = bs.nonbondedParameters;
        s.atomTypes    = bs.atomTypes;
        s.charges      = bs.charges;
        s.atomInfo     = bs.atomInfoAllVdw;
        s.excls        = bs.excls;
        s.x            = bs.coordinates;
        copy_mat(bs.box, s.box);
    }
    else
    {
        std::fprintf(stderr, "unknown system\n");
        return 2;
    }

    const interaction_const_t ic = makeIc(o);

    Writer w(o.out.c_str());

    // ---- system, atom order
    {
        std::vector<float> x(3 * s.x.size());
        for (size_t i = 0; i < s.x.size(); i++)
        {
            for (int d = 0; d < 3; d++) x[3 * i + d] = s.x[i][d];
        }
        w.put("sys_x", 0, x.data(), s.x.size(), 3);
        std::vector<float> q(s.charges.begin(), s.charges.end());
        w.put("sys_q", 0, q.data(), q.size());
        w.put("sys_type", 1, s.atomTypes.data(), s.atomTypes.size());
        std::vector<float> nb(s.nbfp.begin(), s.nbfp.end());
        w.put("sys_nbfp_c6c12", 0, nb.data(), s.numAtomTypes * s.numAtomTypes, 2);
        float box[3] = { s.box[XX][XX], s.box[YY][YY], s.box[ZZ][ZZ] };
        w.put("sys_box", 0, box, 3);
        std::vector<int> exclIdx(1, 0), exclA;
        for (Index i = 0; i < s.excls.ssize(); i++)
        {
            for (int a : s.excls[i]) exclA.push_back(a);
            exclIdx.push_back(int(exclA.size()));
        }
        w.put("sys_excl_index", 1, exclIdx.data(), exclIdx.size());
        w.put("sys_excl_atoms", 1, exclA.data(), exclA.size());
    }

    // ---- interaction constants (what initNbparam consumes, nbnxm_gpu_data_mgmt.cpp:218-320,462)
    w.putf("ic_epsfac", ic.coulomb.epsfac);
    w.putf("ic_rcoulomb", ic.coulomb.cutoff);
    w.putf("ic_ewald_beta", ic.coulomb.ewaldCoeff);
    w.putf("ic_sh_ewald", ic.coulomb.ewaldShift);
    w.putf("ic_k_rf", ic.coulomb.reactionFieldCoefficient);
    w.putf("ic_c_rf", ic.coulomb.reactionFieldShift);
    w.putf("ic_rvdw", ic.vdw.cutoff);
    w.putf("ic_rvdw_switch", ic.vdw.switchDistance);
    w.putf("ic_disp_c2", ic.vdw.dispersionShift.c2);
    w.putf("ic_disp_c3", ic.vdw.dispersionShift.c3);
    w.putf("ic_disp_cpot", ic.vdw.dispersionShift.cpot);
    w.putf("ic_rep_c2", ic.vdw.repulsionShift.c2);
    w.putf("ic_rep_c3", ic.vdw.repulsionShift.c3);
    w.putf("ic_rep_cpot", ic.vdw.repulsionShift.cpot);
    w.putf("ic_sw_c3", ic.vdw.switchConstants.c3);
    w.putf("ic_sw_c4", ic.vdw.switchConstants.c4);
    w.putf("ic_sw_c5", ic.vdw.switchConstants.c5);
    w.putf("ic_ewaldcoeff_lj", ic.vdw.ewaldCoeff);
    w.putf("ic_sh_lj_ewald", ic.vdw.ewaldShift);
    w.putf("ic_vdw_modifier", int(ic.vdw.modifier));
    w.putf("ic_vdw_type", int(ic.vdw.type));
    w.putf("ic_coulomb_type", int(ic.coulomb.type));
    w.putf("rlist", o.rlist);
    if (ic.coulombEwaldTables)
    {
        const auto&        t = *ic.coulombEwaldTables;
        std::vector<float> tf(t.tableF.begin(), t.tableF.end());
        w.put("ic_coulomb_tab_F", 0, tf.data(), tf.size());
        w.putf("ic_coulomb_tab_scale", t.scale);
    }

    // ---- GPU-layout list + nbat (kernel type Cpu8x8x8_PlainC = GMX_EMULATE_GPU layout)
    {
        auto nbv = makeNbv(o, s, NbnxmKernelType::Cpu8x8x8_PlainC, o.rlist, o.minSci);
        const NbnxmPairlistGpu* pl = nbv->pairlistSets().pairlistSet(InteractionLocality::Local).gpuList();
        static_assert(sizeof(nbnxm_sci_t) == 16 && sizeof(nbnxm_cj_packed_t) == 32 && sizeof(nbnxm_excl_t) == 128);
        w.put("pl_sci", 1, pl->sci.data(), pl->sci.size(), 4);
        w.put("pl_cjPacked", 2, pl->cjPacked.list_.data(), pl->cjPacked.list_.size(), 8);
        w.put("pl_excl", 2, pl->excl.data(), pl->excl.size(), 32);
        w.putf("pl_nci_tot", pl->nci_tot);
        w.putf("pl_na_ci", pl->na_ci);

        const nbnxm_atomdata_t& nbat = nbv->nbat();
        const int               n    = nbat.numAtoms();
        w.put("nbat_xq", 0, nbat.x().data(), n, 4);
        w.put("nbat_type", 1, nbat.params().type.data(), nbat.params().type.size());
        std::vector<float> ljc(nbat.params().lj_comb.begin(), nbat.params().lj_comb.end());
        w.put("nbat_lj_comb", 0, ljc.data(), ljc.size() / 2, 2);
        std::vector<float> nbfp(nbat.params().nbfp.begin(), nbat.params().nbfp.end());
        w.put("nbat_nbfp", 0, nbfp.data(), nbfp.size() / 2, 2);
        std::vector<float> nbfpc(nbat.params().nbfp_comb.begin(), nbat.params().nbfp_comb.end());
        w.put("nbat_nbfp_comb", 0, nbfpc.data(), nbfpc.size() / 2, 2);
        w.putf("nbat_ntypes", nbat.params().numTypes);
        w.putf("nbat_comb_rule", int(nbat.params().ljCombinationRule));
        std::vector<RVec> shiftVecs(c_numShiftVectors);
        calc_shifts(s.box, shiftVecs);
        std::vector<float> sv(3 * c_numShiftVectors);
        for (int i = 0; i < c_numShiftVectors; i++)
        {
            for (int d = 0; d < 3; d++) sv[3 * i + d] = shiftVecs[i][d];
        }
        w.put("shift_vec", 0, sv.data(), c_numShiftVectors, 3);
        auto order = nbv->getLocalAtomOrder();
        w.put("nbat_atom_index", 1, order.data(), order.size());

        // Plain-C kernel on the GPU list layout: only RF / tabulated Ewald with plain LJ cut.
        if (ic.vdw.type == VanDerWaalsType::Cut && ic.vdw.modifier == InteractionModifiers::PotShift)
        {
            runAndDump(w, "ref_gpulayout_plainc", *nbv, ic, s, true);
        }
    }

    // ---- the reference's SIMD 4xM kernel (analytical Ewald) on its own CPU list
    {
        auto nbv = makeNbv(o, s, NbnxmKernelType::Cpu4xN_Simd_4xN, o.rc, 0);
        runAndDump(w, "ref_simd4xm", *nbv, ic, s, false);
    }
    return 0;
}
