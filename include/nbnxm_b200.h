/* nbnxm_b200 — C ABI of the Blackwell-native NBNXM short-range nonbonded path.
 *
 * One opaque handle (nbnxm_b200_t) stands for the reference's `NbnxmGpu`
 * (src/gromacs/nbnxm/cuda/nbnxm_cuda_types.h:69).  Every entry point replaces one free function
 * of the reference's GPU-backend boundary, src/gromacs/nbnxm/nbnxm_gpu.h:68-313 and
 * src/gromacs/nbnxm/gpu_data_mgmt.h:66-181; the function it replaces is cited on each declaration.
 *
 * Conventions
 *  - plain pointers and sizes only; host pointers unless a name starts with d_.
 *  - every function returns 0 on success, non-zero on failure; nbnxm_b200_last_error() returns the
 *    message.  (The reference's functions do not return errors, they gmx_fatal(); the C++ shim in
 *    gromacs_b200/gmx_shim turns a non-zero status into gmx_fatal to keep that convention.)
 *  - there is no CPU fallback: without a usable sm_100 device nbnxm_b200_init fails.
 *  - iloc / aloc: 0 = Local, 1 = NonLocal (gmx::InteractionLocality / AtomLocality,
 *    src/gromacs/mdtypes/locality.h); aloc 2 = All where the reference allows it.
 *  - all work is stream-ordered on the handle's local / non-local streams like the reference
 *    (nbnxm_cuda_types.h:112); host buffers passed to async copies must stay alive until
 *    nbnxm_b200_wait_finish_task (pinned memory recommended, cf. HostAllocationPolicy).
 */
#ifndef NBNXM_B200_H
#define NBNXM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbnxm_b200 nbnxm_b200_t;

/* gmx::ElecType / gmx::VdwType, src/gromacs/nbnxm/nbnxm_enums.h:73-110 (same numbering) */
enum nbnxm_b200_elec_type
{
    NBNXM_B200_ELEC_CUT            = 0,
    NBNXM_B200_ELEC_RF             = 1,
    NBNXM_B200_ELEC_EWALD_TAB      = 2,
    NBNXM_B200_ELEC_EWALD_TAB_TWIN = 3,
    NBNXM_B200_ELEC_EWALD_ANA      = 4,
    NBNXM_B200_ELEC_EWALD_ANA_TWIN = 5,
    /* ElecType::None (no NBNxM electrostatics, e.g. FMM computes its own direct part): the LJ-only kernels.  Runs the
     * plain cut-off flavor with epsfac = 0: charges contribute neither forces nor energies, the pair cut-off stays
     * rcoulomb like in the reference's ElecNone kernels.  ElecType::Fmm (7) is not implemented, as in the reference
     * (nbnxm_gpu_data_mgmt.cpp:424-433). */
    NBNXM_B200_ELEC_NONE           = 6
};
enum nbnxm_b200_vdw_type
{
    NBNXM_B200_VDW_CUT           = 0,
    NBNXM_B200_VDW_CUT_COMB_GEOM = 1,
    NBNXM_B200_VDW_CUT_COMB_LB   = 2,
    NBNXM_B200_VDW_FSWITCH       = 3,
    NBNXM_B200_VDW_PSWITCH       = 4,
    NBNXM_B200_VDW_EWALD_GEOM    = 5,
    NBNXM_B200_VDW_EWALD_LB      = 6
};

/* Scalar part of gmx::NBParamGpu (src/gromacs/nbnxm/gpu_types_common.h:222-262), i.e. what
 * set_cutoff_parameters / initNbparam derive from interaction_const_t and PairlistParams
 * (src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:218-240, 462-520). */
typedef struct
{
    int   elec_type; /* nbnxm_b200_elec_type */
    int   vdw_type;  /* nbnxm_b200_vdw_type */
    float epsfac;
    float c_rf;
    float two_k_rf;
    float ewald_beta;
    float sh_ewald;
    float sh_lj_ewald;
    float ewaldcoeff_lj;
    float rcoulomb_sq;
    float rvdw_sq;
    float rvdw_switch;
    float rlist_outer_sq;
    float rlist_inner_sq;
    float disp_c2, disp_c3, disp_cpot; /* shift_consts_t dispersion_shift */
    float rep_c2, rep_c3, rep_cpot;    /* shift_consts_t repulsion_shift  */
    float sw_c3, sw_c4, sw_c5;         /* switch_consts_t vdw_switch      */
    float coulomb_tab_scale;
    int   use_dynamic_pruning;
} nbnxm_b200_params_t;

/* Byte-identical to nbnxm_sci_t / nbnxm_cj_packed_t / nbnxm_excl_t, src/gromacs/nbnxm/pairlist.h:189-287 */
typedef struct
{
    int sci, shift, cj_packed_begin, cj_packed_end;
} nbnxm_b200_sci_t;
typedef struct
{
    int cj[4];
    struct
    {
        unsigned int imask;
        int          excl_ind;
    } imei[2];
} nbnxm_b200_cj_packed_t;
typedef struct
{
    unsigned int pair[32];
} nbnxm_b200_excl_t;

/* gmx_wallclock_gpu_nbnxm_t, src/gromacs/timing/include/gromacs/timing/gpu_timing.h:80-92 (ms, counts) */
typedef struct
{
    double force_ms[2][2]; /* [prune?][energy?] */
    int    force_count[2][2];
    double prune_ms, rolling_prune_ms;
    int    prune_count, rolling_prune_count;
    double xq_h2d_ms, f_d2h_ms, pairlist_h2d_ms;
    /* the non-local share of force_ms / force_count (all flavors): the two localities run on different streams at the
     * same time, so their launch times must not be added up to a "kernel time per step" */
    double force_nonlocal_ms;
    int    force_nonlocal_count;
} nbnxm_b200_timings_t;

const char* nbnxm_b200_last_error(void);

/* gpu_init, nbnxm_gpu_data_mgmt.cpp:637.  nbfp: ntypes*ntypes (6*C6, 12*C12) pairs; nbfp_comb: ntypes
 * pairs or NULL; coulomb_tab: F*r table (EwaldCorrectionTables::tableF) or NULL.
 * local_stream / nonlocal_stream: cudaStream_t to run on, or NULL to let the library create them. */
int nbnxm_b200_init(nbnxm_b200_t** nb, int device, const nbnxm_b200_params_t* params, int ntypes,
                    const float* nbfp, const float* nbfp_comb, const float* coulomb_tab, int coulomb_tab_size,
                    int local_and_nonlocal, void* local_stream, void* nonlocal_stream);
/* gpu_free, nbnxm_gpu_data_mgmt.cpp:1724 */
int nbnxm_b200_free(nbnxm_b200_t* nb);
/* gpu_pme_loadbal_update_param, nbnxm_gpu_data_mgmt.cpp:701 */
int nbnxm_b200_update_params(nbnxm_b200_t* nb, const nbnxm_b200_params_t* params, const float* coulomb_tab,
                             int coulomb_tab_size);

/* gpu_init_pairlist, nbnxm_gpu_data_mgmt.cpp:739 */
int nbnxm_b200_init_pairlist(nbnxm_b200_t* nb, int iloc, const nbnxm_b200_sci_t* sci, int nsci,
                             const nbnxm_b200_cj_packed_t* cj_packed, int ncj_packed,
                             const nbnxm_b200_excl_t* excl, int nexcl, int na_ci);
/* the same for a list that is already in device memory (built by nbnxm_b200_gpu_search_build, include/nbnxm_b200_search.h):
 * device-to-device copies on the list's stream, no host staging */
int nbnxm_b200_init_pairlist_device(nbnxm_b200_t* nb, int iloc, const nbnxm_b200_sci_t* d_sci, int nsci,
                                    const nbnxm_b200_cj_packed_t* d_cj_packed, int ncj_packed,
                                    const nbnxm_b200_excl_t* d_excl, int nexcl, int na_ci);
/* gpu_init_atomdata, nbnxm_gpu_data_mgmt.cpp:1006.  atom_type and/or lj_comb (natoms float pairs)
 * as the flavor needs (types for Cut/FSwitch/PSwitch/Ewald*, lj_comb for CutComb*). */
int nbnxm_b200_init_atomdata(nbnxm_b200_t* nb, int natoms, int natoms_local, const int* atom_type,
                             const float* lj_comb);
/* the same buffers without the uploads: types / lj_comb / xq are written on the device
 * (nbnxm_b200_gpu_search_put_atoms_on_grid, include/nbnxm_b200_search.h) */
int nbnxm_b200_init_atomdata_device(nbnxm_b200_t* nb, int natoms, int natoms_local);
/* gpu_upload_shiftvec, nbnxm_gpu_data_mgmt.cpp:719.  shift_vec: 45 x 3 floats */
int nbnxm_b200_upload_shiftvec(nbnxm_b200_t* nb, const float* shift_vec, int dynamic_box);
/* gpu_copy_xq_to_gpu, nbnxm_gpu_data_mgmt.cpp:1489.  xq: natoms x 4 floats (nbat->x(), XYZQ) */
int nbnxm_b200_copy_xq_to_gpu(nbnxm_b200_t* nb, int aloc, const float* xq);
/* nbnxm_gpu_init_x_to_nbat_x, nbnxm_gpu_data_mgmt.cpp:1605.  One grid per call:
 * atom_index[natoms_nbat] (nbat slot -> rvec index, -1 for fillers), cells packed with
 * num_atoms_per_cell atoms; cxy_na / cxy_ind: per column atom count / first cell (ncolumns, ncolumns+1). */
int nbnxm_b200_init_x_to_nbat_x(nbnxm_b200_t* nb, int grid, int ngrids, const int* atom_index, int natoms_nbat,
                                const int* cxy_na, const int* cxy_ind, int ncolumns, int num_atoms_per_cell,
                                int atom_offset);
/* nbnxm_gpu_x_to_nbat_x, nbnxm_gpu_buffer_ops.cpp:59 / cuda/nbnxm_gpu_buffer_ops_internal.cu:73.
 * d_x: device rvec coordinates; x_ready_event: cudaEvent_t or NULL. */
int nbnxm_b200_x_to_nbat_x(nbnxm_b200_t* nb, const float* d_x, void* x_ready_event, int aloc);

/* gpu_launch_kernel, cuda/nbnxm_cuda.cu:516 */
int nbnxm_b200_launch_kernel(nbnxm_b200_t* nb, int iloc, int compute_energy, int compute_virial);
/* gpu_launch_kernel_pruneonly, cuda/nbnxm_cuda.cu:660 */
int nbnxm_b200_launch_kernel_pruneonly(nbnxm_b200_t* nb, int iloc, int num_parts);
/* gpu_launch_cpyback, nbnxm_gpu_data_mgmt.cpp:1268.  f: natoms x 3 floats (nbat order, XYZ) */
int nbnxm_b200_launch_cpyback(nbnxm_b200_t* nb, int aloc, float* f, int compute_energy, int compute_virial,
                              int use_gpu_f_buffer_ops);
/* gpu_try_finish_task / gpu_wait_finish_task, gpu_common.h:290 / :394.  Adds the staged energies and
 * shift forces (45 x 3) into e_lj, e_el, fshift for aloc == Local, like gpu_reduce_staged_outputs
 * (gpu_common.h:141).  *done = 1 when the task completed (always for wait). */
int nbnxm_b200_try_finish_task(nbnxm_b200_t* nb, int aloc, int compute_energy, int compute_virial,
                               float* e_lj, float* e_el, float* fshift, int* done);
int nbnxm_b200_wait_finish_task(nbnxm_b200_t* nb, int aloc, int compute_energy, int compute_virial,
                                float* e_lj, float* e_el, float* fshift);
/* gpu_clear_outputs, nbnxm_gpu_data_mgmt.cpp:1197 */
int nbnxm_b200_clear_outputs(nbnxm_b200_t* nb, int compute_virial);
/* nbnxmInsertNonlocalGpuDependency, nbnxm_gpu_data_mgmt.cpp:1464 */
int nbnxm_b200_insert_nonlocal_dependency(nbnxm_b200_t* nb, int iloc);
/* setupGpuShortRangeWorkLow / haveGpuShortRangeWork, nbnxm_gpu_data_mgmt.cpp:1244 / :1257 */
int nbnxm_b200_setup_short_range_work(nbnxm_b200_t* nb, int iloc, int have_bonded_work);
int nbnxm_b200_have_short_range_work(const nbnxm_b200_t* nb, int iloc);
/* gpu_min_ci_balanced, cuda/nbnxm_cuda_data_mgmt.cu:122 */
int nbnxm_b200_min_ci_balanced(const nbnxm_b200_t* nb);
/* gpu_is_kernel_ewald_analytical, nbnxm_gpu_data_mgmt.cpp:1238 */
int nbnxm_b200_is_kernel_ewald_analytical(const nbnxm_b200_t* nb);
/* gpu_get_timings / gpu_reset_timings, nbnxm_gpu_data_mgmt.cpp:1224 / :1230 */
int nbnxm_b200_get_timings(nbnxm_b200_t* nb, nbnxm_b200_timings_t* out);
int nbnxm_b200_reset_timings(nbnxm_b200_t* nb);
int nbnxm_b200_set_timing(nbnxm_b200_t* nb, int enable);

/* gpu_get_f / gpuGetNBAtomData, nbnxm_gpu_data_mgmt.cpp:1829 / :1823: device views.
 * d_f is float3-packed (natoms x 3), valid after nbnxm_b200_launch_cpyback's reduction stage
 * (use_gpu_f_buffer_ops = 1 keeps it on the device); d_xq is float4. */
int nbnxm_b200_get_device_buffers(nbnxm_b200_t* nb, float** d_xq, float** d_f, int* natoms);
/* gpuGetNBAtomData, nbnxm_gpu_data_mgmt.cpp:1823, for callers that ADD to the outputs on the device (GPU listed forces,
 * sim_util.cpp:1450): d_f (float3-packed) and d_fshift (float3[45]).  From this call on nbnxm_b200_clear_outputs zeroes
 * both, the copy-back stage adds the kernels' forces on top of what is in d_f, and the shift forces of d_fshift are
 * added to the ones nbnxm_b200_{try,wait}_finish_task return. */
int nbnxm_b200_get_shared_outputs(nbnxm_b200_t* nb, float** d_f, float** d_fshift);
/* ---- perturbed (free-energy) pair kernels (gromacs_b200/csrc/nbnxm_fep.cu) ----
 * copy_gpu_fepparams, gpu_data_mgmt.h:75 (soft-core and coupling parameters; lambda_power 1 or 2) */
int nbnxm_b200_copy_fepparams(nbnxm_b200_t* nb, int have_fep, float alpha_coul, float alpha_vdw, int lambda_power,
                              float sigma6_with_invalid_sigma, float sigma6_minimum, float lambda_coul,
                              float lambda_vdw);
/* the end-state part of gpu_init_atomdata (NBAtomDataGpu::q4 / atomTypes4 / ljComb4): charges of states A and B and,
 * as the VdW flavor needs, types or LJ combination parameters (2 per atom) of both states, nbat order, natoms of
 * nbnxm_b200_init_atomdata */
int nbnxm_b200_init_fep_atomdata(nbnxm_b200_t* nb, const float* q_a, const float* q_b, const int* type_a,
                                 const int* type_b, const float* lj_comb_a, const float* lj_comb_b);
/* gpu_init_feppairlist, nbnxm_gpu_data_mgmt.cpp:880: the perturbed atom-pair list in nbat indices (AtomPairlist:
 * iinr[num_i], jindex[num_i + 1], jjnr / excl_fep per j-entry, shift[num_i]); excl_fep 0 = excluded pair */
int nbnxm_b200_init_feppairlist(nbnxm_b200_t* nb, int iloc, int num_i, const int* iinr, const int* jindex,
                                const int* jjnr, const int* shift, const unsigned char* excl_fep);
/* the same list from device memory, in the layout the kernel reads (the device builder's perturbed-pair pass,
 * nbnxm_b200_gpu_search_set_perturbed): pair_entry[num_pairs] = i-entry of every pair, interacts 0 = excluded pair;
 * device-to-device copies on the locality's stream */
int nbnxm_b200_init_feppairlist_device(nbnxm_b200_t* nb, int iloc, int num_i, int num_pairs, const int* d_iinr,
                                       const int* d_shift, const int* d_pair_entry, const int* d_jjnr,
                                       const unsigned char* d_interacts);
/* gpu_launch_free_energy_kernel, nbnxm_gpu.h:115: adds into the same forces, shift forces and energies as
 * nbnxm_b200_launch_kernel, on the same stream */
int nbnxm_b200_launch_free_energy_kernel(nbnxm_b200_t* nb, int iloc, int compute_energy, int compute_virial);
/* the foreign-lambda launch of gpu_launch_free_energy_kernel (kernel nbfe_foreign_cuda_kernel.cuh): the pairs of the list
 * evaluated at nlambda other coupling parameters, energies only; results with nbnxm_b200_get_fep_foreign:
 * out[4 * k + 0..3] = E_lj, E_el, dV/dlambda VdW, dV/dlambda Coulomb at lambda k (synchronises) */
int nbnxm_b200_launch_foreign_energy_kernel(nbnxm_b200_t* nb, int iloc, int nlambda, const float* lambda_coul,
                                            const float* lambda_vdw);
int nbnxm_b200_get_fep_foreign(nbnxm_b200_t* nb, int nlambda, double* out);
/* dV/dlambda accumulated by the energy launches (NBAtomDataGpu::dvdlLJ / dvdlElec); synchronises */
int nbnxm_b200_get_fep_dvdl(nbnxm_b200_t* nb, float* dvdl_lj, float* dvdl_el, int clear);

/* GpuForceReduction::reinit / execute (src/gromacs/mdlib/gpuforcereduction_impl.cpp, reduceKernel in
 * gpuforcereduction_impl_internal.cu:61-118): f_total[a] (+)= f_nbat[cell[a]] (+ rvec_force_to_add[a]) for the atoms
 * [atom_start, atom_start + num_atoms), all arrays but `cell` in device memory, rvecs as packed float3.  cell[natoms]:
 * atom -> nbat slot (GridSet::cells()).  The nbat forces are read from the kernels' accumulator, so neither
 * nbnxm_b200_launch_cpyback nor a packing pass is needed first; the caller orders `stream` (NULL: the local stream)
 * after the force kernels, as the reference does with its dependency list. */
int nbnxm_b200_init_reduce_f(nbnxm_b200_t* nb, const int* cell, int natoms);
int nbnxm_b200_reduce_f(nbnxm_b200_t* nb, float* d_f_total, const float* d_rvec_force_to_add, int atom_start,
                        int num_atoms, int accumulate, void* stream);
/* streams the handle runs on (cudaStream_t), for callers that order their own work against it */
int nbnxm_b200_get_streams(nbnxm_b200_t* nb, void** local_stream, void** nonlocal_stream);

/* ---- inspection helpers (the reference reads the same data back in
 * GpuPairlistTest, src/gromacs/nbnxm/tests/pairlist.cpp:236) ---- */
/* copies the device cjPacked array (with pruned imasks), the outer-pruned imask array (2 per cjPacked),
 * sciSorted and the per-sci histogram index back to the host; any pointer may be NULL. */
int nbnxm_b200_download_pairlist(nbnxm_b200_t* nb, int iloc, nbnxm_b200_cj_packed_t* cj_packed,
                                 unsigned int* imask_outer, nbnxm_b200_sci_t* sci_sorted, int* sci_count,
                                 int* rolling_part);
/* number of atom pairs the last force launch on iloc evaluated (32 per set imask bit), counted on the
 * device when counting is enabled with nbnxm_b200_set_pair_counting */
int nbnxm_b200_set_pair_counting(nbnxm_b200_t* nb, int enable);
int nbnxm_b200_get_pair_count(nbnxm_b200_t* nb, int iloc, long long* npairs);
/* number of kernels this handle launched since init (for bench.py's gpu_launches) */
long long nbnxm_b200_launch_count(const nbnxm_b200_t* nb);

/* measured FP32 FMA peak of the device in TFLOP/s (pure-FFMA kernel, best of 5): the roofline denominator */
int nbnxm_b200_measure_fp32_peak(int device, double* tflops);

/* ---- x-slab halo exchange helpers (the reference's pack / unpack kernels,
 * src/gromacs/domdec/gpuhaloexchange_impl_gpu.cu:82-137); the transport (NCCL send/recv) is the caller's ---- */
/* d_send[i] = xq[index[i]] (+ shift on xyz); index on device */
int nbnxm_b200_pack_xq(nbnxm_b200_t* nb, const int* d_index, int n, const float* shift3, float* d_send, void* stream);
/* xq[first + i] = d_recv[i] */
int nbnxm_b200_unpack_xq(nbnxm_b200_t* nb, int first, int n, const float* d_recv, void* stream);
/* d_send[i] = f[first + i] (float4), then zeroes nothing */
int nbnxm_b200_pack_f(nbnxm_b200_t* nb, int first, int n, float* d_send, void* stream);
/* f[index[i]] += d_recv[i] */
int nbnxm_b200_unpack_add_f(nbnxm_b200_t* nb, const int* d_index, int n, const float* d_recv, void* stream);

/* ---- x-slab halo exchange over NCCL send/recv (NVLink), one process per GPU.  Replaces
 * gmx::GpuHaloExchange (src/gromacs/domdec/gpuhaloexchange.h:80-130) for a 1-D decomposition along x:
 * slabs are whole grid columns, so with atoms in nbat order the region sent to the -x neighbour is the
 * contiguous range [send_first, send_first + send_count) of the local atoms and the halo received from
 * the +x neighbour the contiguous range [recv_first, ...) of the non-local atoms.  All calls are
 * stream-ordered on the handle's non-local stream. ---- */
/* ncclGetUniqueId into a 128-byte buffer (rank 0 calls it, the caller broadcasts the bytes) */
int nbnxm_b200_halo_get_unique_id(char* id, int nbytes);
/* ncclCommInitRank: collective over the nranks handles (GpuHaloExchange constructor) */
int nbnxm_b200_halo_init(nbnxm_b200_t* nb, const char* id, int rank, int nranks);
int nbnxm_b200_halo_free(nbnxm_b200_t* nb);
/* GpuHaloExchange::reinitHalo, gpuhaloexchange_impl_gpu.cpp:152: called at search steps after gpu_init_atomdata */
int nbnxm_b200_halo_set_ranges(nbnxm_b200_t* nb, int send_first, int send_count, int recv_first, int recv_count);
/* communicateHaloCoordinates, gpuhaloexchange_impl_gpu.cpp:286: xq of our first columns -> rank-1, halo xq <- rank+1 */
int nbnxm_b200_halo_exchange_x(nbnxm_b200_t* nb);
/* communicateHaloForces, gpuhaloexchange_impl_gpu.cpp:340: halo f -> rank+1, f += what rank-1 computed on our atoms */
int nbnxm_b200_halo_exchange_f(nbnxm_b200_t* nb);
int nbnxm_b200_halo_set_timing(nbnxm_b200_t* nb, int enable);
int nbnxm_b200_halo_get_timings(nbnxm_b200_t* nb, double* x_ms, double* f_ms, int reset);

/* ---- the nonbonded part of one do_force step in one call (src/gromacs/mdlib/sim_util.cpp:1639-2442): what the
 * reference's do_force does around the two kernels, in its order: [copy xq H2D] -> clear outputs -> non-local
 * dependency + halo coordinates -> local kernel -> non-local kernel + halo forces -> rolling prune on its schedule
 * (local on even steps, non-local on odd ones with a halo; odd steps without, prunekerneldispatch.cpp:123-128) ->
 * force copy-back of both localities.  Nothing here is new functionality: it saves a caller that is not compiled
 * code (the Python bench and tests) a dozen foreign-function calls per step.
 * xq_host / f_host may be NULL (coordinates resident, forces left on the device). ---- */
/* ---- peer-memory halo: the B200 / NVSwitch way of the same exchange.  Instead of sending halo coordinates and
 * returning halo forces, every rank maps its +x neighbour's xq and force accumulator (CUDA IPC over NVLink); the
 * non-local kernel reads the neighbour's coordinates and reduces j-forces into the neighbour's accumulator directly
 * (red.global.add over NVLink), and ranks synchronise through step counters in device memory.  No transport call,
 * no pack / add kernels, no halo copy.  Replaces communicateHaloCoordinates + communicateHaloForces
 * (gpuhaloexchange_impl_gpu.cpp:286-470) when all ranks sit on one NVLink domain.
 * Set-up (after gpu_init_atomdata and halo_set_ranges at every search step): every rank exports a blob, the caller
 * hands each rank the blob of its +x neighbour (e.g. all_gather), then peer_import.  Use have_halo = 3 in
 * nbnxm_b200_do_force_step. ---- */
int nbnxm_b200_peer_blob_size(void);
int nbnxm_b200_peer_export(nbnxm_b200_t* nb, unsigned char* blob, int nbytes);
int nbnxm_b200_peer_import(nbnxm_b200_t* nb, const unsigned char* up_blob, int nbytes);
int nbnxm_b200_peer_close(nbnxm_b200_t* nb);
int nbnxm_b200_peer_error(nbnxm_b200_t* nb, int* error);

typedef struct nbnxm_b200_step_flags
{
    int compute_energy, compute_virial;
    int have_halo;              /* 1: the handle was created with local_and_nonlocal and halo ranges are set;
                                   3: peer-memory halo (nbnxm_b200_peer_import done);
                                   2: local and non-local lists but no transport (one slab of a decomposition run alone,
                                   halo coordinates resident: a profiling aid) */
    int dynamic_pruning;        /* 1: launch the rolling prune on its schedule */
    int rolling_prune_parts;
} nbnxm_b200_step_flags_t;
int nbnxm_b200_do_force_step(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* flags, const float* xq_host, float* f_host);
/* The same step for host-resident coordinates and forces (single rank), with the transfers pipelined against the
 * force kernel: atoms are cut into nchunks contiguous ranges (chunk_first_atom[nchunks + 1], whole grid columns), the
 * sci array - in the caller's order, grouped by the chunk of the i-atoms - into the matching ranges
 * (chunk_first_sci[nchunks + 1]); chunk_needs[k] is the bit mask of the atom chunks the entries of sci chunk k read
 * coordinates from / add forces to (from the outer list).  Coordinates go up chunk by chunk, the kernel of an sci chunk
 * starts when the chunks it needs have arrived, the forces of an atom chunk come down when every sci chunk touching it
 * has run.  The reference overlaps its non-local transfer with local work in the same spirit (sim_util.cpp:1772-1922).
 * Falls back to the plain sequence on steps with a fresh list. */
int nbnxm_b200_do_force_step_pipelined(nbnxm_b200_t* nb, int step, const nbnxm_b200_step_flags_t* flags, const float* xq_host,
                                       float* f_host, int nchunks, const int* chunk_first_atom, const int* chunk_first_sci,
                                       const unsigned int* chunk_needs);
/* Diagnostics: with the timeline enabled, a pipelined step records CUDA timing events per chunk; get_pipeline_timeline
 * synchronises the device and returns, for each of *nchunks chunks, ms[4 * c + 0..3] = end of its H2D copy, start and end of
 * its force kernel, end of its D2H copy, in ms after the start of the step. */
int nbnxm_b200_set_pipeline_timeline(nbnxm_b200_t* nb, int enable);
int nbnxm_b200_get_pipeline_timeline(nbnxm_b200_t* nb, int max_chunks, int* nchunks, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* NBNXM_B200_H */
