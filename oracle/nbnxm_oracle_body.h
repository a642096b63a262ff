/* TEST INFRASTRUCTURE — included twice by nbnxm_oracle.c (REAL = double, REAL = float).
 *
 * One atom-pair interaction, restated from the reference GPU kernel body
 * src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh:505-672 and the shared helpers
 * src/gromacs/nbnxm/nbnxm_kernel_utils.h:56-250 (force switch :69, LJ-PME :141,
 * potential switch :174, pmeCorrF :216, table interpolation :252).
 *
 * REAL == double: analytical Ewald uses libm erf/exp (the mathematical function the
 * reference's rational approximation pmeCorrF targets).
 * REAL == float : analytical Ewald uses the pmeCorrF rational approximation itself.
 */

static inline REAL FN(pme_corr_f)(REAL z2)
{
#if REAL_IS_DOUBLE
    /* pmeCorrF(z^2) = (2/sqrt(pi) * z * exp(-z^2) - erf(z)) / z^3, the function the reference's
     * rational polynomial approximates (nbnxm_kernel_utils.h:216). Series for small z. */
    if (z2 < 1e-8)
    {
        return -4.0 / (3.0 * 1.772453850905516027) + z2 * (4.0 / (5.0 * 1.772453850905516027));
    }
    const double z = sqrt(z2);
    return (2.0 / 1.772453850905516027 * z * exp(-z2) - erf(z)) / (z * z2);
#else
    const float FN6 = -1.7357322914161492954e-8f, FN5 = 1.4703624142580877519e-6f,
                FN4 = -0.000053401640219807709149f, FN3 = 0.0010054721316683106153f,
                FN2 = -0.019278317264888380590f, FN1 = 0.069670166153766424023f,
                FN0 = -0.75225204789749321333f;
    const float FD4 = 0.0011193462567257629232f, FD3 = 0.014866955030185295499f,
                FD2 = 0.11583842382862377919f, FD1 = 0.50736591960530292870f, FD0 = 1.0f;
    const float z4  = z2 * z2;
    float       pd0 = FD4 * z4 + FD2;
    const float pd1 = FD3 * z4 + FD1;
    pd0             = pd0 * z4 + FD0;
    pd0             = pd1 * z2 + pd0;
    pd0             = 1.0f / pd0;
    float pn0       = FN6 * z4 + FN4;
    float pn1       = FN5 * z4 + FN3;
    pn0             = pn0 * z4 + FN2;
    pn1             = pn1 * z4 + FN1;
    pn0             = pn0 * z4 + FN0;
    pn0             = pn1 * z2 + pn0;
    return pn0 * pd0;
#endif
}

/* c6/c12 are 6*C6 and 12*C12 (atomdata.cpp:579-588). qq = epsfac*qi*qj.
 * int_bit: 1 = normal pair, 0 = topology-excluded pair (still gets Ewald/RF correction).
 * Returns F/r; adds pair energies to *e_lj, *e_el when calc_e. */
static inline REAL FN(pair)(const orc_params_t* p, REAL r2, REAL qq, REAL c6, REAL c12, REAL c6grid,
                            REAL int_bit, int calc_e, const float* tab, REAL* e_lj, REAL* e_el)
{
    const REAL one_sixth = (REAL)(1.0 / 6.0), one_twelfth = (REAL)(1.0 / 12.0);
    const int  vdw = p->vdw_type, elec = p->elec_type;

    if (r2 < (REAL)3.82e-07f) /* c_nbnxmMinDistanceSquared, pairlist.h:154 */
    {
        r2 = (REAL)3.82e-07f;
    }
    const REAL inv_r  = (REAL)1.0 / SQRT(r2);
    const REAL inv_r2 = inv_r * inv_r;
    REAL       inv_r6 = inv_r2 * inv_r2 * inv_r2;
    inv_r6 *= int_bit;

    REAL F_invr = inv_r6 * (c12 * inv_r6 - c6) * inv_r2;
    REAL E_lj_p = int_bit
                  * (c12 * (inv_r6 * inv_r6 + (REAL)p->rep_cpot) * one_twelfth
                     - c6 * (inv_r6 + (REAL)p->disp_cpot) * one_sixth);

    if (vdw == ORC_VDW_FSWITCH)
    {
        const REAL r   = r2 * inv_r;
        REAL       rsw = r - (REAL)p->rvdw_switch;
        rsw            = rsw > 0 ? rsw : 0;
        F_invr += -c6 * ((REAL)p->disp_c2 + (REAL)p->disp_c3 * rsw) * rsw * rsw * inv_r
                  + c12 * ((REAL)p->rep_c2 + (REAL)p->rep_c3 * rsw) * rsw * rsw * inv_r;
        E_lj_p += c6 * ((REAL)p->disp_c2 / 3 + (REAL)p->disp_c3 / 4 * rsw) * rsw * rsw * rsw
                  - c12 * ((REAL)p->rep_c2 / 3 + (REAL)p->rep_c3 / 4 * rsw) * rsw * rsw * rsw;
    }
    if (vdw == ORC_VDW_EWALD_GEOM || vdw == ORC_VDW_EWALD_LB)
    {
        const REAL lje_coeff2   = (REAL)p->ewaldcoeff_lj * (REAL)p->ewaldcoeff_lj;
        const REAL lje_coeff6_6 = lje_coeff2 * lje_coeff2 * lje_coeff2 * one_sixth;
        const REAL inv_r6_nm    = inv_r2 * inv_r2 * inv_r2;
        const REAL cr2          = lje_coeff2 * r2;
        const REAL expmcr2      = EXP(-cr2);
        const REAL poly         = (REAL)1.0 + cr2 + (REAL)0.5 * cr2 * cr2;
        F_invr += c6grid * (inv_r6_nm - expmcr2 * (inv_r6_nm * poly + lje_coeff6_6)) * inv_r2;
        E_lj_p += one_sixth * c6grid
                  * (inv_r6_nm * ((REAL)1.0 - expmcr2 * poly) + (REAL)p->sh_lj_ewald * int_bit);
    }
    if (vdw == ORC_VDW_PSWITCH)
    {
        const REAL r   = r2 * inv_r;
        const REAL rsw = r - (REAL)p->rvdw_switch;
        if (rsw > 0)
        {
            const REAL c3 = p->sw_c3, c4 = p->sw_c4, c5 = p->sw_c5;
            const REAL sw  = 1 + (c3 + (c4 + c5 * rsw) * rsw) * rsw * rsw * rsw;
            const REAL dsw = (3 * c3 + (4 * c4 + 5 * c5 * rsw) * rsw) * rsw * rsw;
            F_invr         = F_invr * sw - inv_r * E_lj_p * dsw;
            E_lj_p *= sw;
        }
    }
    if (elec == ORC_ELEC_EWALD_TAB_TWIN || elec == ORC_ELEC_EWALD_ANA_TWIN)
    {
        /* VDW_CUTOFF_CHECK, nbnxm_cuda_kernel.cuh:615-624 */
        const REAL in = (r2 < (REAL)p->rvdw_sq) ? 1 : 0;
        F_invr *= in;
        E_lj_p *= in;
    }
    if (calc_e)
    {
        *e_lj += E_lj_p;
    }

    if (elec == ORC_ELEC_CUT)
    {
        F_invr += qq * int_bit * inv_r2 * inv_r;
        if (calc_e) *e_el += qq * (int_bit * inv_r - (REAL)p->c_rf);
    }
    else if (elec == ORC_ELEC_RF)
    {
        F_invr += qq * (int_bit * inv_r2 * inv_r - (REAL)p->two_k_rf);
        if (calc_e) *e_el += qq * (int_bit * inv_r + (REAL)0.5 * (REAL)p->two_k_rf * r2 - (REAL)p->c_rf);
    }
    else
    {
        const REAL beta = p->ewald_beta;
        if (elec == ORC_ELEC_EWALD_ANA || elec == ORC_ELEC_EWALD_ANA_TWIN)
        {
            F_invr += qq * (int_bit * inv_r2 * inv_r + FN(pme_corr_f)(beta * beta * r2) * beta * beta * beta);
        }
        else
        {
            /* interpolateCoulombForceR, nbnxm_kernel_utils.h:252-289 */
            const REAL norm = (REAL)p->coulomb_tab_scale * (r2 * inv_r);
            const int  idx  = (int)norm;
            const REAL frac = norm - idx;
            const REAL fr   = ((REAL)1.0 - frac) * tab[idx] + frac * tab[idx + 1];
            F_invr += qq * (int_bit * inv_r2 - fr) * inv_r;
        }
        if (calc_e)
        {
            *e_el += qq * (inv_r * (int_bit - ERF(r2 * inv_r * beta)) - int_bit * (REAL)p->sh_ewald);
        }
    }
    return F_invr;
}

/* LJ pair parameters for (ai, aj) following the flavor's parameter source
 * (nbnxm_cuda_kernel.cuh:519-536, nbnxm_kernel_utils.h:56-65,114-137). */
static inline void FN(lj_params)(const orc_params_t* p, int ai, int aj, const int* type, const float* lj_comb,
                                 const float* nbfp, const float* nbfp_comb, REAL* c6, REAL* c12, REAL* c6grid)
{
    *c6grid = 0;
    if (p->vdw_type == ORC_VDW_CUT_COMB_GEOM)
    {
        *c6  = (REAL)lj_comb[2 * ai] * (REAL)lj_comb[2 * aj];
        *c12 = (REAL)lj_comb[2 * ai + 1] * (REAL)lj_comb[2 * aj + 1];
    }
    else if (p->vdw_type == ORC_VDW_CUT_COMB_LB)
    {
        const REAL sigma   = (REAL)lj_comb[2 * ai] + (REAL)lj_comb[2 * aj];
        const REAL epsilon = (REAL)lj_comb[2 * ai + 1] * (REAL)lj_comb[2 * aj + 1];
        const REAL s2      = sigma * sigma;
        const REAL s6      = s2 * s2 * s2;
        *c6                = epsilon * s6;
        *c12               = *c6 * s6;
    }
    else
    {
        const int ti = type[ai], tj = type[aj];
        *c6  = nbfp[2 * (p->ntypes * ti + tj)];
        *c12 = nbfp[2 * (p->ntypes * ti + tj) + 1];
        if (p->vdw_type == ORC_VDW_EWALD_GEOM)
        {
            *c6grid = (REAL)nbfp_comb[2 * ti] * (REAL)nbfp_comb[2 * tj];
        }
        else if (p->vdw_type == ORC_VDW_EWALD_LB)
        {
            const REAL sigma   = (REAL)nbfp_comb[2 * ti] + (REAL)nbfp_comb[2 * tj];
            const REAL epsilon = (REAL)nbfp_comb[2 * ti + 1] * (REAL)nbfp_comb[2 * tj + 1];
            const REAL s2      = sigma * sigma;
            *c6grid            = epsilon * s2 * s2 * s2;
        }
    }
}

/* One sci entry of the GPU-layout list (all its cjPacked groups).
 * Loop structure and mask semantics: nbnxm_cuda_kernel.cuh:326-717. ftile accumulates with REAL. */
static inline int64_t FN(sci_entry)(const orc_params_t* p, const orc_sci_t* s, const orc_cjp_t* cjp,
                                    const orc_excl_t* excl, const float* xq, const int* type,
                                    const float* lj_comb, const float* nbfp, const float* nbfp_comb,
                                    const float* tab, const float* shift_vec, int calc_e, REAL* f,
                                    REAL* fshift_sum, REAL* e_lj, REAL* e_el)
{
    int64_t    npairs = 0;
    const int  central = 22; /* gmx::c_centralShiftIndex, pbcutil/ishift.h */
    const REAL rc2     = p->rcoulomb_sq;
    const float* sh    = shift_vec + 3 * s->shift;
    const int    excl_forces =
            !(p->elec_type == ORC_ELEC_CUT && !calc_e
              && !(p->vdw_type == ORC_VDW_EWALD_GEOM || p->vdw_type == ORC_VDW_EWALD_LB));

    if (calc_e && s->shift == central && cjp[s->cj_begin].cj[0] == s->sci * 8)
    {
        /* self terms on the diagonal sci entry, nbnxm_cuda_kernel.cuh:383-417 */
        REAL qsum = 0, c6sum = 0;
        for (int a = s->sci * 64; a < s->sci * 64 + 64; a++)
        {
            qsum += (REAL)xq[4 * a + 3] * (REAL)xq[4 * a + 3];
            if (p->vdw_type == ORC_VDW_EWALD_GEOM || p->vdw_type == ORC_VDW_EWALD_LB)
            {
                c6sum += nbfp[2 * (type[a] * (p->ntypes + 1))];
            }
        }
        if (p->vdw_type == ORC_VDW_EWALD_GEOM || p->vdw_type == ORC_VDW_EWALD_LB)
        {
            const REAL c2 = (REAL)p->ewaldcoeff_lj * (REAL)p->ewaldcoeff_lj;
            *e_lj += c6sum * (REAL)0.5 * (REAL)(1.0 / 6.0) * (c2 * c2 * c2 * (REAL)(1.0 / 6.0));
        }
        if (p->elec_type == ORC_ELEC_CUT || p->elec_type == ORC_ELEC_RF)
        {
            *e_el += (REAL)p->epsfac * qsum * (REAL)-0.5 * (REAL)p->c_rf;
        }
        else
        {
            *e_el += (REAL)p->epsfac * qsum * -(REAL)p->ewald_beta * (REAL)0.564189583547756286948;
        }
    }

    for (int jp = s->cj_begin; jp < s->cj_end; jp++)
    {
        for (int w = 0; w < 2; w++)
        {
            const uint32_t   imask = cjp[jp].imei[w].imask;
            const orc_excl_t* ex   = &excl[cjp[jp].imei[w].excl_ind];
            if (!imask) continue;
            for (int jm = 0; jm < 4; jm++)
            {
                if (!(imask & (0xffu << (jm * 8)))) continue;
                const int cj = cjp[jp].cj[jm];
                for (int i = 0; i < 8; i++)
                {
                    const uint32_t bit = 1u << (jm * 8 + i);
                    if (!(imask & bit)) continue;
                    const int ci = s->sci * 8 + i;
                    npairs += 32;
                    for (int tj = 4 * w; tj < 4 * w + 4; tj++)
                    {
                        const int aj = cj * 8 + tj;
                        for (int ti = 0; ti < 8; ti++)
                        {
                            const int  ai      = ci * 8 + ti;
                            const REAL int_bit = (ex->pair[(tj & 3) * 8 + ti] & bit) ? 1 : 0;
                            /* i coordinates are shifted and rounded to float first (:341) */
                            const float xi = xq[4 * ai] + sh[0], yi = xq[4 * ai + 1] + sh[1],
                                        zi = xq[4 * ai + 2] + sh[2];
                            const REAL dx = (REAL)xi - (REAL)xq[4 * aj], dy = (REAL)yi - (REAL)xq[4 * aj + 1],
                                       dz = (REAL)zi - (REAL)xq[4 * aj + 2];
                            const REAL r2 = dx * dx + dy * dy + dz * dz;
                            int        within;
                            if (excl_forces)
                            {
                                const int non_self = !(s->shift == central && tj <= ti);
                                within             = (r2 < rc2) && (non_self || ci != cj);
                            }
                            else
                            {
                                within = (r2 < rc2) && int_bit != 0;
                            }
                            if (!within) continue;
                            REAL c6, c12, c6grid;
                            FN(lj_params)(p, ai, aj, type, lj_comb, nbfp, nbfp_comb, &c6, &c12, &c6grid);
                            const REAL qq = (REAL)p->epsfac * (REAL)xq[4 * ai + 3] * (REAL)xq[4 * aj + 3];
                            const REAL F  = FN(pair)(p, r2, qq, c6, c12, c6grid, int_bit, calc_e, tab, e_lj, e_el);
                            f[3 * ai] += F * dx;
                            f[3 * ai + 1] += F * dy;
                            f[3 * ai + 2] += F * dz;
                            f[3 * aj] -= F * dx;
                            f[3 * aj + 1] -= F * dy;
                            f[3 * aj + 2] -= F * dz;
                            fshift_sum[0] += F * dx;
                            fshift_sum[1] += F * dy;
                            fshift_sum[2] += F * dz;
                        }
                    }
                }
            }
        }
    }
    return npairs;
}
