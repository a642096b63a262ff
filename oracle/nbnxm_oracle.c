/* TEST INFRASTRUCTURE — CPU oracle for the NBNXM cluster-pair nonbonded path.
 * See nbnxm_oracle.h for scope, parity status and the list of allowed callers.
 *
 * Restated (not copied) from:
 *   force loop      src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:326-717
 *                   (cross-checked with kernels_reference/kernel_gpu_ref.cpp:64-344)
 *   pair physics    src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-289
 *   prune           src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel_pruneonly.cuh:123-346
 *   list layout     src/gromacs/nbnxm/pairlist.h:189-287, pairlist.cpp:651-688
 */
#include "nbnxm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#    include <omp.h>
#endif

#define REAL double
#define REAL_IS_DOUBLE 1
#define FN(x) x##_d
#define SQRT sqrt
#define EXP exp
#define ERF erf
#include "nbnxm_oracle_body.h"
#undef REAL
#undef REAL_IS_DOUBLE
#undef FN
#undef SQRT
#undef EXP
#undef ERF

#define REAL float
#define REAL_IS_DOUBLE 0
#define FN(x) x##_f
#define SQRT sqrtf
#define EXP expf
#define ERF erff
#include "nbnxm_oracle_body.h"
#undef REAL
#undef REAL_IS_DOUBLE
#undef FN
#undef SQRT
#undef EXP
#undef ERF

int64_t orc_forces(const orc_params_t* p, int nsci, const orc_sci_t* sci, const orc_cjp_t* cjp,
                   const orc_excl_t* excl, const float* xq, const int* type, const float* lj_comb,
                   const float* nbfp, const float* nbfp_comb, const float* coulomb_tab,
                   const float* shift_vec, int calc_energy, int calc_fshift, double* f,
                   double* fshift, double* e)
{
    int64_t npairs = 0;
    for (int s = 0; s < nsci; s++)
    {
        double fs[3] = { 0, 0, 0 };
        npairs += sci_entry_d(p, &sci[s], cjp, excl, xq, type, lj_comb, nbfp, nbfp_comb, coulomb_tab,
                              shift_vec, calc_energy, f, fs, &e[0], &e[1]);
        if (calc_fshift)
        {
            for (int d = 0; d < 3; d++) fshift[3 * sci[s].shift + d] += fs[d];
        }
    }
    return npairs;
}

int64_t orc_forces_f32_omp(const orc_params_t* p, int nsci, const orc_sci_t* sci, const orc_cjp_t* cjp,
                           const orc_excl_t* excl, const float* xq, const int* type,
                           const float* lj_comb, const float* nbfp, const float* nbfp_comb,
                           const float* shift_vec, int natoms, int calc_energy, float* f,
                           double* e, int nthreads)
{
    int64_t npairs = 0;
    double  elj = 0, eel = 0;
    if (nthreads < 1) nthreads = 1;
    float* fbuf = (float*)calloc((size_t)nthreads * natoms * 3, sizeof(float));
#pragma omp parallel num_threads(nthreads) reduction(+ : npairs, elj, eel)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* ft = fbuf + (size_t)tid * natoms * 3;
        float  fs[3], el = 0, ee = 0;
#pragma omp for schedule(dynamic, 8)
        for (int s = 0; s < nsci; s++)
        {
            npairs += sci_entry_f(p, &sci[s], cjp, excl, xq, type, lj_comb, nbfp, nbfp_comb, NULL,
                                  shift_vec, calc_energy, ft, fs, &el, &ee);
        }
        elj += el;
        eel += ee;
#pragma omp barrier
#pragma omp for schedule(static)
        for (int a = 0; a < natoms * 3; a++)
        {
            float sum = 0;
            for (int t = 0; t < nthreads; t++) sum += fbuf[(size_t)t * natoms * 3 + a];
            f[a] += sum;
        }
    }
    free(fbuf);
    e[0] += elj;
    e[1] += eel;
    return npairs;
}

void orc_prune(const orc_params_t* p, int nsci, const orc_sci_t* sci_order, orc_cjp_t* cjp,
               uint32_t* imask_outer, const float* xq, const float* shift_vec, int fresh, int part,
               int nparts, int* sci_count)
{
    const float rlo2 = p->rlist_outer_sq, rli2 = p->rlist_inner_sq;
    for (int s = (fresh ? 0 : part); s < nsci; s += (fresh ? 1 : nparts))
    {
        const orc_sci_t* sc = &sci_order[s];
        const float*     sh = shift_vec + 3 * sc->shift;
        int              count = 0;
        for (int jp = sc->cj_begin; jp < sc->cj_end; jp++)
        {
            for (int w = 0; w < 2; w++)
            {
                uint32_t full, check, neu;
                if (fresh)
                {
                    full  = cjp[jp].imei[w].imask;
                    check = full;
                    neu   = 0;
                }
                else
                {
                    full  = imask_outer[jp * 2 + w];
                    neu   = cjp[jp].imei[w].imask;
                    check = neu ^ full;
                }
                if (!check) continue;
                for (int jm = 0; jm < 4; jm++)
                {
                    if (!(check & (0xffu << (jm * 8)))) continue;
                    const int cj = cjp[jp].cj[jm];
                    for (int i = 0; i < 8; i++)
                    {
                        const uint32_t bit = 1u << (jm * 8 + i);
                        if (!(check & bit)) continue;
                        int any_outer = 0, any_inner = 0;
                        for (int tj = 4 * w; tj < 4 * w + 4; tj++)
                        {
                            const float* xj = xq + 4 * (cj * 8 + tj);
                            for (int ti = 0; ti < 8; ti++)
                            {
                                const float* xa = xq + 4 * ((sc->sci * 8 + i) * 8 + ti);
                                /* xi = x + shift rounded to float, then rv = xi - xj, then
                                 * norm2 contracted by nvcc to fma(z,z,fma(y,y,x*x))
                                 * (pruneonly.cuh:208-210,285-286) */
                                const float xi = xa[0] + sh[0], yi = xa[1] + sh[1], zi = xa[2] + sh[2];
                                const float dx = xi - xj[0], dy = yi - xj[1], dz = zi - xj[2];
                                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                                any_outer |= (r2 < rlo2);
                                any_inner |= (r2 < rli2);
                            }
                        }
                        if (fresh && !any_outer) full &= ~bit;
                        if (any_inner) neu |= bit;
                    }
                }
                if (fresh)
                {
                    imask_outer[jp * 2 + w] = full;
                    count += __builtin_popcount(neu);
                }
                cjp[jp].imei[w].imask = neu;
            }
        }
        if (fresh && sci_count)
        {
            int idx      = 8192 - count - 1; /* c_sciHistogramSize, gpu_types_common.h:83 */
            sci_count[s] = idx > 0 ? idx : 0;
        }
    }
}

void orc_brute_force(const orc_params_t* p, int n, const float* x, const float* q, const int* type,
                     const float* nbfp, const float* nbfp_comb, const float* box,
                     const int* excl_index, const int* excl_atoms, int calc_energy, double* f,
                     double* e)
{
    const double rc2 = p->rcoulomb_sq;
    orc_params_t pp  = *p;
    /* brute force always reads the type table */
    if (pp.vdw_type == ORC_VDW_CUT_COMB_GEOM || pp.vdw_type == ORC_VDW_CUT_COMB_LB) pp.vdw_type = ORC_VDW_CUT;
    const int excl_forces =
            !(p->elec_type == ORC_ELEC_CUT && !calc_energy
              && !(p->vdw_type == ORC_VDW_EWALD_GEOM || p->vdw_type == ORC_VDW_EWALD_LB));
    for (int i = 0; i < n; i++)
    {
        if (calc_energy)
        {
            const double q2 = (double)p->epsfac * q[i] * q[i];
            if (p->elec_type == ORC_ELEC_CUT || p->elec_type == ORC_ELEC_RF)
                e[1] += -0.5 * p->c_rf * q2;
            else
                e[1] += -(double)p->ewald_beta * 0.564189583547756286948 * q2;
            if (p->vdw_type == ORC_VDW_EWALD_GEOM || p->vdw_type == ORC_VDW_EWALD_LB)
            {
                const double c2 = (double)p->ewaldcoeff_lj * p->ewaldcoeff_lj;
                e[0] += nbfp[2 * (type[i] * (p->ntypes + 1))] * 0.5 / 6.0 * (c2 * c2 * c2 / 6.0);
            }
        }
        for (int j = i + 1; j < n; j++)
        {
            double d[3];
            for (int k = 0; k < 3; k++)
            {
                d[k] = (double)x[3 * i + k] - (double)x[3 * j + k];
                d[k] -= box[k] * nearbyint(d[k] / box[k]);
            }
            const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (r2 >= rc2) continue;
            double int_bit = 1;
            for (int k = excl_index[i]; k < excl_index[i + 1]; k++)
                if (excl_atoms[k] == j) int_bit = 0;
            if (!excl_forces && int_bit == 0) continue;
            double c6, c12, c6grid;
            lj_params_d(&pp, i, j, type, NULL, nbfp, nbfp_comb, &c6, &c12, &c6grid);
            const double qq = (double)p->epsfac * q[i] * q[j];
            const double F  = pair_d(&pp, r2, qq, c6, c12, c6grid, int_bit, calc_energy, NULL, &e[0], &e[1]);
            for (int k = 0; k < 3; k++)
            {
                f[3 * i + k] += F * d[k];
                f[3 * j + k] -= F * d[k];
            }
        }
    }
}
