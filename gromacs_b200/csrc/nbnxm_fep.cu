/* Perturbed (free-energy) nonbonded pair kernels behind the reference's FEP entry points:
 *   nbnxm_b200_copy_fepparams            = copy_gpu_fepparams            (src/gromacs/nbnxm/gpu_data_mgmt.h:75)
 *   nbnxm_b200_init_fep_atomdata         = the q4 / atomTypes4 / ljComb4 part of gpu_init_atomdata
 *                                          (nbnxm_gpu_data_mgmt.cpp:1006, NBAtomDataGpu, gpu_types_common.h:165)
 *   nbnxm_b200_init_feppairlist          = gpu_init_feppairlist          (nbnxm_gpu_data_mgmt.cpp:880)
 *   nbnxm_b200_launch_free_energy_kernel = gpu_launch_free_energy_kernel (cuda/nbfe_cuda.cu, kernel nbfe_cuda_kernel.cuh)
 * The kernel is one thread per pair of the atom-pair list around the host+device body in nbfe_bodies.h, which the CPU
 * tests run pair by pair against the pinned oracle (oracle/nbfe_oracle.py).  It adds into the same force accumulator,
 * shift forces and energies as the cluster-pair kernels, on the list's stream; dV/dlambda has its own accumulator.
 *   nbnxm_b200_launch_foreign_energy_kernel = the foreign-lambda launch of gpu_launch_free_energy_kernel
 *                                          (kernel nbfe_foreign_cuda_kernel.cuh): the same pair evaluation at every
 *                                          foreign lambda, energies and dV/dlambda only, one launch per lambda.
 */
#include "nbfe_bodies.h"
#include "nbnxm_handle.cuh"

namespace nbb
{

__global__ void __launch_bounds__(128) nbfe_pair_kernel(const nbfe::Params p, const nbfe::Atoms a, const nbfe::List l)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j < l.numPairs)
    {
        nbfe::pair(p, a, l, j);
    }
}

} // namespace nbb

using nbb::fail;

extern "C" {

int nbnxm_b200_copy_fepparams(nbnxm_b200_t* nb, int have_fep, float alpha_coul, float alpha_vdw, int lambda_power,
                              float sigma6_with_invalid_sigma, float sigma6_minimum, float lambda_coul, float lambda_vdw)
{
    if (!nb) return fail("nbnxm_b200_copy_fepparams: null handle");
    if (lambda_power != 1 && lambda_power != 2) return fail("nbnxm_b200_copy_fepparams: soft-core lambda power %d (1 or 2)", lambda_power);
    nb->haveFep                   = have_fep != 0;
    nb->fepAlphaCoul              = alpha_coul;
    nb->fepAlphaVdw               = alpha_vdw;
    nb->fepLambdaPower            = lambda_power;
    nb->fepSigma6WithInvalidSigma = sigma6_with_invalid_sigma;
    nb->fepSigma6Minimum          = sigma6_minimum;
    nb->fepLambdaCoul             = lambda_coul;
    nb->fepLambdaVdw              = lambda_vdw;
    CU(cudaSetDevice(nb->device));
    CU(nb->fepDvdl.reserve(2));
    CU(cudaMemsetAsync(nb->fepDvdl.p, 0, sizeof(double) * 2, nb->stream[0]));
    return 0;
}

int nbnxm_b200_init_fep_atomdata(nbnxm_b200_t* nb, const float* q_a, const float* q_b, const int* type_a, const int* type_b,
                                 const float* lj_comb_a, const float* lj_comb_b)
{
    if (!nb || !q_a || !q_b) return fail("nbnxm_b200_init_fep_atomdata: null argument");
    const int  n    = nb->natoms;
    const bool comb = (nb->params.vdw_type == NBNXM_B200_VDW_CUT_COMB_GEOM || nb->params.vdw_type == NBNXM_B200_VDW_CUT_COMB_LB);
    if (comb && (!lj_comb_a || !lj_comb_b)) return fail("nbnxm_b200_init_fep_atomdata: this VdW flavor needs lj_comb of both end states");
    if (!comb && (!type_a || !type_b)) return fail("nbnxm_b200_init_fep_atomdata: this VdW flavor needs the atom types of both end states");
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[0];
    CU(cudaStreamSynchronize(nb->stream[0]));
    CU(cudaStreamSynchronize(nb->stream[1]));
    std::vector<float> q(size_t(n) * 2);
    for (int i = 0; i < n; i++)
    {
        q[2 * i]     = q_a[i];
        q[2 * i + 1] = q_b[i];
    }
    CU(nb->fepQ.reserve(size_t(n) * 2 + 2));
    CU(cudaMemcpyAsync(nb->fepQ.p, q.data(), sizeof(float) * 2 * n, cudaMemcpyHostToDevice, st));
    std::vector<int>   t;
    std::vector<float> lj;
    CU(nb->fepType.reserve(size_t(n) * 2 + 2));
    CU(nb->fepLjComb.reserve(size_t(n) * 4 + 4));
    if (!comb)
    {
        t.resize(size_t(n) * 2);
        for (int i = 0; i < n; i++)
        {
            t[2 * i]     = type_a[i];
            t[2 * i + 1] = type_b[i];
        }
        CU(cudaMemcpyAsync(nb->fepType.p, t.data(), sizeof(int) * 2 * n, cudaMemcpyHostToDevice, st));
    }
    else
    {
        lj.resize(size_t(n) * 4);
        for (int i = 0; i < n; i++)
        {
            lj[4 * i]     = lj_comb_a[2 * i];
            lj[4 * i + 1] = lj_comb_a[2 * i + 1];
            lj[4 * i + 2] = lj_comb_b[2 * i];
            lj[4 * i + 3] = lj_comb_b[2 * i + 1];
        }
        CU(cudaMemcpyAsync(nb->fepLjComb.p, lj.data(), sizeof(float) * 4 * n, cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st)); /* the staging vectors go out of scope */
    nb->haveFepAtomdata = true;
    return 0;
}

int nbnxm_b200_init_feppairlist(nbnxm_b200_t* nb, int iloc, int num_i, const int* iinr, const int* jindex, const int* jjnr,
                                const int* shift, const unsigned char* excl_fep)
{
    if (!nb || iloc < 0 || iloc > 1 || num_i < 0) return fail("nbnxm_b200_init_feppairlist: bad argument");
    if (num_i > 0 && (!iinr || !jindex || !jjnr || !shift)) return fail("nbnxm_b200_init_feppairlist: null list array");
    CU(cudaSetDevice(nb->device));
    nbnxm_b200::FepList& fl = nb->feplist[iloc];
    cudaStream_t         st = nb->stream[iloc];
    CU(cudaStreamSynchronize(st));
    /* jindex is an offset table starting at 0 (AtomPairlist::flatJList); anything else would leave j-entries unchecked */
    if (num_i > 0 && jindex[0] != 0) return fail("nbnxm_b200_init_feppairlist: jindex[0] is %d, expected 0", jindex[0]);
    const int numPairs = num_i > 0 ? jindex[num_i] : 0;
    for (int n = 0; n < num_i; n++)
    {
        if (iinr[n] < 0 || iinr[n] >= nb->natoms || shift[n] < 0 || shift[n] >= nbb::c_numShiftVectors || jindex[n + 1] < jindex[n])
        {
            return fail("nbnxm_b200_init_feppairlist: i-entry %d is outside the atom / shift range", n);
        }
    }
    std::vector<int>           pairEntry(numPairs);
    std::vector<unsigned char> excl(numPairs, 1);
    for (int n = 0; n < num_i; n++)
    {
        for (int j = jindex[n]; j < jindex[n + 1]; j++)
        {
            if (jjnr[j] < 0 || jjnr[j] >= nb->natoms) return fail("nbnxm_b200_init_feppairlist: j-atom %d outside the atom range", jjnr[j]);
            pairEntry[j] = n;
            if (excl_fep) excl[j] = excl_fep[j] ? 1 : 0;
        }
    }
    CU(fl.pairEntry.reserve(numPairs + 1));
    CU(fl.jjnr.reserve(numPairs + 1));
    CU(fl.exclFep.reserve(numPairs + 1));
    CU(fl.iinr.reserve(num_i + 1));
    CU(fl.shift.reserve(num_i + 1));
    if (numPairs > 0)
    {
        CU(cudaMemcpyAsync(fl.pairEntry.p, pairEntry.data(), sizeof(int) * numPairs, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(fl.jjnr.p, jjnr, sizeof(int) * numPairs, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(fl.exclFep.p, excl.data(), numPairs, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(fl.iinr.p, iinr, sizeof(int) * num_i, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(fl.shift.p, shift, sizeof(int) * num_i, cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st));
    fl.numI     = num_i;
    fl.numPairs = numPairs;
    return 0;
}

int nbnxm_b200_init_feppairlist_device(nbnxm_b200_t* nb, int iloc, int num_i, int num_pairs, const int* d_iinr, const int* d_shift,
                                       const int* d_pair_entry, const int* d_jjnr, const unsigned char* d_interacts)
{
    if (!nb || iloc < 0 || iloc > 1 || num_i < 0 || num_pairs < 0) return fail("nbnxm_b200_init_feppairlist_device: bad argument");
    if (num_pairs > 0 && (!d_iinr || !d_shift || !d_pair_entry || !d_jjnr || !d_interacts || num_i == 0))
    {
        return fail("nbnxm_b200_init_feppairlist_device: null list array");
    }
    CU(cudaSetDevice(nb->device));
    nbnxm_b200::FepList& fl = nb->feplist[iloc];
    cudaStream_t         st = nb->stream[iloc];
    CU(cudaStreamSynchronize(st)); /* kernels of earlier steps may still read the previous list */
    CU(fl.pairEntry.reserve(num_pairs + 1));
    CU(fl.jjnr.reserve(num_pairs + 1));
    CU(fl.exclFep.reserve(num_pairs + 1));
    CU(fl.iinr.reserve(num_i + 1));
    CU(fl.shift.reserve(num_i + 1));
    if (num_pairs > 0)
    {
        CU(cudaMemcpyAsync(fl.pairEntry.p, d_pair_entry, sizeof(int) * num_pairs, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fl.jjnr.p, d_jjnr, sizeof(int) * num_pairs, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fl.exclFep.p, d_interacts, num_pairs, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fl.iinr.p, d_iinr, sizeof(int) * num_i, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fl.shift.p, d_shift, sizeof(int) * num_i, cudaMemcpyDeviceToDevice, st));
    }
    fl.numI     = num_pairs > 0 ? num_i : 0;
    fl.numPairs = num_pairs;
    return 0;
}

/* kernel arguments of one launch at the coupling parameters (lambdaCoul, lambdaVdw) */
static int fepLaunchArgs(nbnxm_b200_t* nb, int iloc, float lambdaCoul, float lambdaVdw, nbfe::Params* pOut, nbfe::Atoms* aOut, nbfe::List* lOut)
{
    const nbnxm_b200::FepList& fl = nb->feplist[iloc];
    const nbnxm_b200_params_t& s  = nb->params;
    nbfe::Params               p{};
    switch (s.elec_type)
    {
        case NBNXM_B200_ELEC_CUT: p.elec = nbfe::ElecCut; break;
        case NBNXM_B200_ELEC_RF: p.elec = nbfe::ElecRF; break;
        case NBNXM_B200_ELEC_EWALD_TAB:
        case NBNXM_B200_ELEC_EWALD_ANA: p.elec = nbfe::ElecEwald; break;
        case NBNXM_B200_ELEC_EWALD_TAB_TWIN:
        case NBNXM_B200_ELEC_EWALD_ANA_TWIN:
            p.elec = nbfe::ElecEwald;
            p.twin = 1;
            break;
        default: return fail("perturbed kernels: electrostatics type %d has no perturbed kernel", s.elec_type);
    }
    switch (s.vdw_type)
    {
        case NBNXM_B200_VDW_CUT: p.vdw = nbfe::VdwCut; break;
        case NBNXM_B200_VDW_CUT_COMB_GEOM: p.vdw = nbfe::VdwCombGeom; break;
        case NBNXM_B200_VDW_CUT_COMB_LB: p.vdw = nbfe::VdwCombLB; break;
        case NBNXM_B200_VDW_FSWITCH: p.vdw = nbfe::VdwFSwitch; break;
        case NBNXM_B200_VDW_PSWITCH: p.vdw = nbfe::VdwPSwitch; break;
        default: return fail("perturbed kernels: VdW type %d (LJ-PME) has no perturbed kernel", s.vdw_type);
    }
    p.epsfac = s.epsfac; p.c_rf = s.c_rf; p.two_k_rf = s.two_k_rf; p.beta = s.ewald_beta; p.sh_ewald = s.sh_ewald;
    p.rcoulomb_sq = s.rcoulomb_sq; p.rvdw_sq = s.rvdw_sq; p.rvdw_switch = s.rvdw_switch;
    p.disp_c2 = s.disp_c2; p.disp_c3 = s.disp_c3; p.disp_cpot = s.disp_cpot;
    p.rep_c2 = s.rep_c2; p.rep_c3 = s.rep_c3; p.rep_cpot = s.rep_cpot;
    p.sw_c3 = s.sw_c3; p.sw_c4 = s.sw_c4; p.sw_c5 = s.sw_c5;
    p.alphaCoul = nb->fepAlphaCoul; p.alphaVdw = nb->fepAlphaVdw;
    p.sigma6WithInvalidSigma = nb->fepSigma6WithInvalidSigma; p.sigma6Minimum = nb->fepSigma6Minimum;
    p.lambdaCoul = lambdaCoul; p.lambdaVdw = lambdaVdw; p.lambdaPower = nb->fepLambdaPower;
    p.numTypes   = nb->numTypes;
    p.calcForces = 1;

    nbfe::Atoms a{};
    a.xq       = reinterpret_cast<const float*>(nb->xq.p);
    a.qAB      = nb->fepQ.p;
    a.typeAB   = nb->fepType.p;
    a.ljCombAB = nb->fepLjComb.p;
    a.nbfp     = reinterpret_cast<const float*>(nb->nbfp.p);
    a.shiftVec = nb->shiftVec.p;
    a.f4       = reinterpret_cast<float*>(nb->f4.p);
    a.fshift   = nb->fshift.p;
    a.energy   = nb->energy.p;
    a.dvdl     = nb->fepDvdl.p;
    nbfe::List l{};
    l.numPairs  = fl.numPairs;
    l.pairEntry = fl.pairEntry.p;
    l.iinr      = fl.iinr.p;
    l.shift     = fl.shift.p;
    l.jjnr      = fl.jjnr.p;
    l.exclFep   = fl.exclFep.p;
    *pOut = p;
    *aOut = a;
    *lOut = l;
    return 0;
}

int nbnxm_b200_launch_free_energy_kernel(nbnxm_b200_t* nb, int iloc, int compute_energy, int compute_virial)
{
    if (!nb || iloc < 0 || iloc > 1) return fail("nbnxm_b200_launch_free_energy_kernel: bad argument");
    if (!nb->haveFep) return fail("nbnxm_b200_launch_free_energy_kernel: call nbnxm_b200_copy_fepparams first");
    if (!nb->haveFepAtomdata) return fail("nbnxm_b200_launch_free_energy_kernel: call nbnxm_b200_init_fep_atomdata first");
    const nbnxm_b200::FepList& fl = nb->feplist[iloc];
    if (fl.numPairs == 0) return 0;
    nbfe::Params p;
    nbfe::Atoms  a;
    nbfe::List   l;
    if (fepLaunchArgs(nb, iloc, nb->fepLambdaCoul, nb->fepLambdaVdw, &p, &a, &l)) return 1;
    p.calcEnergy = compute_energy != 0;
    p.calcFshift = compute_virial != 0;
    CU(cudaSetDevice(nb->device));
    nbb::nbfe_pair_kernel<<<(fl.numPairs + 127) / 128, 128, 0, nb->stream[iloc]>>>(p, a, l);
    nb->launches++;
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_launch_foreign_energy_kernel(nbnxm_b200_t* nb, int iloc, int nlambda, const float* lambda_coul, const float* lambda_vdw)
{
    if (!nb || iloc < 0 || iloc > 1 || nlambda < 1 || !lambda_coul || !lambda_vdw) return fail("nbnxm_b200_launch_foreign_energy_kernel: bad argument");
    if (!nb->haveFep || !nb->haveFepAtomdata) return fail("nbnxm_b200_launch_foreign_energy_kernel: perturbed kernels are not set up");
    CU(cudaSetDevice(nb->device));
    cudaStream_t st = nb->stream[iloc];
    if (size_t(nlambda) * 4 > nb->fepForeign.alloc) CU(cudaStreamSynchronize(st));
    CU(nb->fepForeign.reserve(size_t(nlambda) * 4));
    CU(cudaMemsetAsync(nb->fepForeign.p, 0, sizeof(double) * 4 * nlambda, st));
    nb->fepNumForeign = nlambda;
    const nbnxm_b200::FepList& fl = nb->feplist[iloc];
    if (fl.numPairs == 0) return 0;
    for (int k = 0; k < nlambda; k++)
    {
        nbfe::Params p;
        nbfe::Atoms  a;
        nbfe::List   l;
        if (fepLaunchArgs(nb, iloc, lambda_coul[k], lambda_vdw[k], &p, &a, &l)) return 1;
        p.calcForces = 0;
        p.calcFshift = 0;
        p.calcEnergy = 1;
        a.energy     = nb->fepForeign.p + 4 * k;     /* E_lj, E_el */
        a.dvdl       = nb->fepForeign.p + 4 * k + 2; /* dV/dlambda: VdW, Coulomb */
        nbb::nbfe_pair_kernel<<<(fl.numPairs + 127) / 128, 128, 0, st>>>(p, a, l);
        nb->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

int nbnxm_b200_get_fep_foreign(nbnxm_b200_t* nb, int nlambda, double* out)
{
    if (!nb || !out || nlambda < 1 || nlambda > nb->fepNumForeign) return fail("nbnxm_b200_get_fep_foreign: bad argument");
    CU(cudaSetDevice(nb->device));
    CU(cudaStreamSynchronize(nb->stream[0]));
    if (nb->stream[1] != nb->stream[0]) CU(cudaStreamSynchronize(nb->stream[1]));
    CU(cudaMemcpy(out, nb->fepForeign.p, sizeof(double) * 4 * nlambda, cudaMemcpyDeviceToHost));
    return 0;
}

int nbnxm_b200_get_fep_dvdl(nbnxm_b200_t* nb, float* dvdl_lj, float* dvdl_el, int clear)
{
    if (!nb || !nb->fepDvdl.p) return fail("nbnxm_b200_get_fep_dvdl: no perturbed kernel was set up");
    CU(cudaSetDevice(nb->device));
    CU(cudaStreamSynchronize(nb->stream[0]));
    if (nb->stream[1] != nb->stream[0]) CU(cudaStreamSynchronize(nb->stream[1]));
    double v[2];
    CU(cudaMemcpy(v, nb->fepDvdl.p, sizeof(v), cudaMemcpyDeviceToHost));
    if (dvdl_lj) *dvdl_lj = static_cast<float>(v[0]);
    if (dvdl_el) *dvdl_el = static_cast<float>(v[1]);
    if (clear) CU(cudaMemset(nb->fepDvdl.p, 0, sizeof(v)));
    return 0;
}

} // extern "C"
