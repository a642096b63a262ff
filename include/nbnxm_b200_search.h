/* nbnxm_b200_search — grid binning (host) and GPU-layout pair-list construction (host and GPU builders) (C ABI).
 *
 * This is the caller side of the force path ("next" row of the scope table): it produces exactly the
 * inputs the reference hands to gpu_init_atomdata / gpu_init_pairlist, in the reference's formats:
 *   - atoms sorted into columns / 64-atom bins / 2x2x2 clusters of 8 (Grid::putOnGrid and
 *     sortCellsGpuGeometry, src/gromacs/nbnxm/grid.cpp:1612, :1169),
 *   - nbnxm_sci_t / nbnxm_cj_packed_t / nbnxm_excl_t arrays (src/gromacs/nbnxm/pairlist.h:189-287) with the
 *     mask conventions of src/gromacs/nbnxm/pairlist.cpp:651-688 (self/Newton exclusions), :1561-1660
 *     (topology exclusions), :1769-1879 (splitting of long i-entries for load balance).
 * It is an independent implementation (bounding-box search over a column grid); the list it builds is not
 * entry-for-entry the reference's list (any list that covers all pairs within rlist exactly once is valid), which
 * tests/test_pairsearch.py verifies against brute force.  Two builders produce the same grid and the same list:
 * the host one (nbnxm_b200_grid_* / nbnxm_b200_pairlist_*, OpenMP over i-bins) and the device one
 * (nbnxm_b200_gpu_search_*, data-parallel passes on the GPU that holds the coordinates).
 * Rectangular boxes only.
 */
#ifndef NBNXM_B200_SEARCH_H
#define NBNXM_B200_SEARCH_H

#include "nbnxm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbnxm_b200_grid nbnxm_b200_grid_t;

/* grid dimensions for natoms atoms in the box: columns of about 2 x 2 clusters of 8 atoms in cross-section
 * (Grid::setDimensions, src/gromacs/nbnxm/grid.cpp:161-311); ncx a multiple of nslabs when nslabs > 1 */
int nbnxm_b200_grid_dims(const float* box, int natoms, int nslabs, int* ncx, int* ncy);
/* nonbonded_verlet_t::putAtomsOnGrid (src/gromacs/nbnxm/nbnxm.cpp:78): bins natoms atoms (x: natoms x 3,
 * inside the rectangular box [0, box)) and sorts them into nbat order. */
int nbnxm_b200_grid_create(nbnxm_b200_grid_t** grid, const float* box, int natoms, const float* x, int nthreads);
/* the same with the number of x columns rounded down to a multiple of nslabs (x-slab decomposition over nslabs GPUs) */
int nbnxm_b200_grid_create_slabs(nbnxm_b200_grid_t** grid, const float* box, int natoms, const float* x, int nthreads, int nslabs);
int nbnxm_b200_grid_free(nbnxm_b200_grid_t* grid);
/* number of nbat slots (atoms padded to whole 64-atom bins) and bins, grid columns along x and y */
int nbnxm_b200_grid_info(const nbnxm_b200_grid_t* grid, int* natoms_nbat, int* nbins, int* ncx, int* ncy);
int nbnxm_b200_grid_box(const nbnxm_b200_grid_t* grid, float* box3);
/* atom_index[natoms_nbat]: nbat slot -> atom (-1 = filler) (GridSet::atomIndices);
 * first_bin_of_column[ncx*ncy + 1] (Grid::cellToBin_) */
int nbnxm_b200_grid_get_order(const nbnxm_b200_grid_t* grid, int* atom_index, int* first_bin_of_column);
/* nbnxm_atomdata_t::copy x / setAtomProperties (src/gromacs/nbnxm/atomdata.cpp:159-280, :1107): fills
 * xq[natoms_nbat*4] (fillers at -1e6 with q = 0), type[natoms_nbat] (fillers get ntypes-1) and, when
 * lj_comb_per_type != NULL (ntypes x 2), lj_comb[natoms_nbat*2]. Any output may be NULL. */
int nbnxm_b200_grid_fill_atomdata(const nbnxm_b200_grid_t* grid, const float* x, const float* q, const int* type,
                                  int ntypes, const float* lj_comb_per_type, float* xq, int* type_nbat,
                                  float* lj_comb);

/* nonbonded_verlet_t::constructPairlist (src/gromacs/nbnxm/pairlist.cpp:4056) for the GPU layout.
 * excl_index/excl_atoms: topology exclusions as CSR in atom order (may be NULL).
 * min_sci: split i-entries so that at least about this many are produced (gpu_min_ci_balanced); 0 = no split.
 * bin_begin/bin_end: only i-bins in [bin_begin, bin_end) get i-entries; j_bin_lo/j_bin_hi restrict the
 * j-clusters to bins in [j_bin_lo, j_bin_hi); pass 0, nbins, 0, nbins for the whole system.
 * inter_zone = 0: i and j ranges are the same zone: half-shell rule (backward PBC shifts only, j >= i on the
 *   central shift), like an intra-grid list of the reference (pairlist.cpp:3063-3067, 3226-3229).
 * inter_zone = 1: i bins are a home x-slab and j bins its +x neighbour slab (halo): every pair is listed
 *   once for all y/z shifts and only the x shift `required_tx` (0, or -1 across the periodic boundary),
 *   like the reference's inter-zone lists of the eighth-shell scheme (domdec/domdec_zones.cpp:55-83). */
int nbnxm_b200_pairlist_build(nbnxm_b200_grid_t* grid, float rlist, const int* excl_index, const int* excl_atoms,
                              int min_sci, int bin_begin, int bin_end, int j_bin_lo, int j_bin_hi, int inter_zone,
                              int required_tx, int nthreads);
int nbnxm_b200_pairlist_sizes(const nbnxm_b200_grid_t* grid, int* nsci, int* ncj_packed, int* nexcl,
                              long long* ncluster_pairs);
int nbnxm_b200_pairlist_copy(const nbnxm_b200_grid_t* grid, nbnxm_b200_sci_t* sci, nbnxm_b200_cj_packed_t* cj_packed,
                             nbnxm_b200_excl_t* excl);

/* make_fep_list for the GPU layout (pairlist.cpp:1414): splits every atom pair with a perturbed atom off the list built
 * last into an atom-pair list in nbat indices (the arguments of nbnxm_b200_init_feppairlist: iinr[num_i], jindex[num_i+1],
 * jjnr / interacts[num_j], shift[num_i]) and clears its bit in the cluster list (fetch the modified list with
 * nbnxm_b200_pairlist_sizes / _copy afterwards).  perturbed[natoms]: 1 for atoms whose charge or type differs between
 * the end states.  The cluster kernels' atom data must mask these atoms (charge 0, type ntypes - 1:
 * nbnxm_atomdata_mask_fep, atomdata.cpp:1039). */
int nbnxm_b200_pairlist_split_fep(nbnxm_b200_grid_t* grid, const unsigned char* perturbed);
int nbnxm_b200_pairlist_fep_sizes(const nbnxm_b200_grid_t* grid, int* num_i, int* num_j);
int nbnxm_b200_pairlist_fep_copy(const nbnxm_b200_grid_t* grid, int* iinr, int* jindex, int* jjnr, int* shift,
                                 unsigned char* interacts);

/* ---- host-side planning (gromacs_b200/csrc/hostplan.cpp) ----
 * x-slab decomposition over nslabs GPUs, one process per GPU: bins of slab r (home), of its one-sided halo (the first
 * columns of slab (r+1) % nslabs within rlist, plus one column of slack) and the x shift of the home x halo pairs
 * (-1 across the periodic boundary); the analogue of the reference's zone set-up, domdec/domdec_zones.cpp:55-83 */
int nbnxm_b200_slab_bin_ranges(const nbnxm_b200_grid_t* grid, int nslabs, int r, float rlist, int* home_begin, int* home_end,
                               int* halo_begin, int* halo_end, int* required_tx);
/* a list built on the global grid -> one rank's atom order (home bins first, then halo bins), in place: sci.sci relative
 * to the first home bin, j-clusters relative to the home range (halo = 0) or placed after it (halo = 1) */
int nbnxm_b200_pairlist_reindex(nbnxm_b200_sci_t* sci, int nsci, nbnxm_b200_cj_packed_t* cj_packed, int ncj_packed, int first_home_bin,
                                int first_halo_bin, int num_home_bins, int nclusters_total, int halo);
/* chunk plan of nbnxm_b200_do_force_step_pipelined (include/nbnxm_b200.h): at most 32 chunks of whole grid columns;
 * first_atom / first_sci: [nchunks + 1]; needs[k]: bit c set = entries of sci chunk k read or write atoms of chunk c;
 * the sci array is grouped by chunk in place (entries of a chunk keep their order) */
int nbnxm_b200_chunk_plan(const nbnxm_b200_grid_t* grid, nbnxm_b200_sci_t* sci, int nsci, const nbnxm_b200_cj_packed_t* cj_packed,
                          int ncj_packed, int nchunks_requested, int* nchunks, int* first_atom, int* first_sci, unsigned int* needs);

/* ---- the same list built on the GPU (gromacs_b200/csrc/nbnxm_gpusearch.cu) ----
 *
 * constructPairlist + gpu_init_pairlist (pairlist.cpp:4056, nbnxm_gpu_data_mgmt.cpp:739) without the host in
 * between: the list is built from the coordinates resident in the handle (`xq` after nbnxm_b200_copy_xq_to_gpu /
 * nbnxm_b200_x_to_nbat_x) and installed as the handle's list for `iloc`; it equals the list of
 * nbnxm_b200_pairlist_build (one thread) entry for entry, except for the numbering of the nbnxm_excl_t entries.
 * Order of calls at a search step: grid (nbnxm_b200_grid_create, host) -> nbnxm_b200_init_atomdata ->
 * nbnxm_b200_copy_xq_to_gpu -> nbnxm_b200_gpu_search_set_grid -> nbnxm_b200_gpu_search_build. */
typedef struct nbnxm_b200_gpu_search nbnxm_b200_gpu_search_t;

/* one search object per force handle; it runs on the handle's local stream and must be freed before the handle */
int nbnxm_b200_gpu_search_create(nbnxm_b200_gpu_search_t** search, nbnxm_b200_t* nb);
int nbnxm_b200_gpu_search_free(nbnxm_b200_gpu_search_t* search);
/* the grid of nbnxm_b200_grid_get_order / nbnxm_b200_grid_info (nbins * 64 must equal the handle's natoms) and the
 * topology exclusions (CSR in atom order, may be NULL); the arrays are copied before the call returns */
int nbnxm_b200_gpu_search_set_grid(nbnxm_b200_gpu_search_t* search, const float* box, int ncx, int ncy,
                                   const int* first_bin_of_column, const int* atom_index, int nbins, int natoms,
                                   const int* excl_index, const int* excl_atoms);
/* ---- the grid itself on the GPU: nonbonded_verlet_t::putAtomsOnGrid + setAtomProperties + gpu_init_atomdata
 * (nbnxm.cpp:78, atomdata.cpp:1107, nbnxm_gpu_data_mgmt.cpp:1006) from coordinates in device memory ----
 * set_atoms (once per topology): charges, types (atom order), per-type LJ combination parameters (ntypes x 2, for the
 * CutComb* flavors) and the topology exclusions (CSR in atom order); any array may be NULL.
 * put_atoms_on_grid (every search step): d_x = natoms rvecs in device memory, atom order, inside the box; bins the
 * atoms into the columns of nbnxm_b200_grid_dims, sorts every column (one block per column in its shared
 * memory, buckets along z + ranks; a column may hold at most 8192 atoms), sizes the handle's atom buffers for the new grid and writes
 * xq / types / lj_comb in nbat order, the slot -> atom map used by nbnxm_b200_x_to_nbat_x and the atom -> slot map
 * used by nbnxm_b200_reduce_f.  The order equals nbnxm_b200_grid_create's.  Follow with nbnxm_b200_gpu_search_build. */
int nbnxm_b200_gpu_search_set_atoms(nbnxm_b200_gpu_search_t* search, int natoms, const float* q, const int* type, int ntypes,
                                    const float* lj_comb_per_type, const int* excl_index, const int* excl_atoms);
int nbnxm_b200_gpu_search_put_atoms_on_grid(nbnxm_b200_gpu_search_t* search, const float* box, int nslabs, const float* d_x,
                                            void* x_ready_event, int* natoms_nbat, int* nbins, int* ncx, int* ncy);
/* host copies of the grid order (GridSet::atomIndices, Grid::cxy_ind) and the device time of the last gridding (ms) */
int nbnxm_b200_gpu_search_get_order(nbnxm_b200_gpu_search_t* search, int* atom_index, int* first_bin_of_column, float* grid_ms);
/* arguments as nbnxm_b200_pairlist_build (bin_end / j_bin_hi < 0: up to the last bin of the grid); the result becomes
 * the handle's list for iloc (haveFreshList set) */
int nbnxm_b200_gpu_search_build(nbnxm_b200_gpu_search_t* search, int iloc, float rlist, int min_sci, int bin_begin,
                                int bin_end, int j_bin_lo, int j_bin_hi, int inter_zone, int required_tx);
/* The search step of one x-slab of a multi-GPU run on the device (what gromacs_b200/multigpu.py's make_slab_plan does with the
 * host builder and nbnxm_b200_pairlist_reindex): `search` holds the grid of the whole system (put_atoms_on_grid with nslabs =
 * the number of ranks); `target` is the rank's handle on the same device.
 *   gather_slab: sizes target's atom buffers for (home bins + halo bins) * 64 atoms, home atoms local, and copies xq / types /
 *                lj_comb of the two bin ranges there, device to device.
 *   build_slab : iloc 0 = home x home (half shell), iloc 1 = home x halo (inter-zone mode, every y / z shift, required_tx = the
 *                x shift of the i-atoms, -1 across the periodic boundary); the list is re-indexed to the rank's order on the
 *                device and becomes target's list for iloc.  Bin ranges as nbnxm_b200_slab_bin_ranges gives them. */
int nbnxm_b200_gpu_search_gather_slab(nbnxm_b200_gpu_search_t* search, nbnxm_b200_t* target, int home_begin, int home_end,
                                      int halo_begin, int halo_end);
int nbnxm_b200_gpu_search_build_slab(nbnxm_b200_gpu_search_t* search, nbnxm_b200_t* target, int iloc, float rlist, int min_sci,
                                     int home_begin, int home_end, int halo_begin, int halo_end, int required_tx);
/* Perturbed (free-energy) atoms, make_fep_list (pairlist.cpp:1414-1560) on the device: perturbed[natoms] in atom order, host
 * memory (1 = charge or type differs between the end states; NULL: none).  From then on nbnxm_b200_gpu_search_build moves every
 * pair with a perturbed atom from the cluster list to an atom-pair list - the list nbnxm_b200_pairlist_split_fep makes on the
 * host, entry for entry - and installs it as the handle's perturbed list of that locality (nbnxm_b200_init_feppairlist without
 * the host), and nbnxm_b200_gpu_search_put_atoms_on_grid masks the perturbed atoms out of the cluster kernels' atom data
 * (charge 0, type ntypes - 1: nbnxm_atomdata_mask_fep, atomdata.cpp:1039).  Not for the slab builds. */
int nbnxm_b200_gpu_search_set_perturbed(nbnxm_b200_gpu_search_t* search, int natoms, const unsigned char* perturbed);
/* the perturbed list built last: sizes; copy in the host builder's form (iinr[num_i], jindex[num_i + 1], jjnr / interacts per
 * j-entry, shift[num_i]; tests) */
int nbnxm_b200_gpu_search_fep_sizes(const nbnxm_b200_gpu_search_t* search, int* num_i, int* num_j);
int nbnxm_b200_gpu_search_fep_download(nbnxm_b200_gpu_search_t* search, int* iinr, int* jindex, int* jjnr, int* shift,
                                       unsigned char* interacts);
/* sizes of the list built last, cluster pairs in it, device time of the build (ms, CUDA events) */
int nbnxm_b200_gpu_search_sizes(const nbnxm_b200_gpu_search_t* search, int* nsci, int* ncj_packed, int* nexcl,
                                long long* ncluster_pairs, float* build_ms);
/* copy of the list built last (tests, callers that keep a host copy) */
int nbnxm_b200_gpu_search_download(nbnxm_b200_gpu_search_t* search, nbnxm_b200_sci_t* sci,
                                   nbnxm_b200_cj_packed_t* cj_packed, nbnxm_b200_excl_t* excl);

#ifdef __cplusplus
}
#endif
#endif
