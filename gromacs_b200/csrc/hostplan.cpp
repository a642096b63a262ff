/* Host-side planning around the force path (include/nbnxm_b200_search.h), C++:
 *   - x-slab decomposition of one pair-search grid over G GPUs with a one-sided halo (the zone set-up the reference
 *     does in src/gromacs/domdec/domdec_zones.cpp:55-83 for its eighth-shell scheme, here for slabs along x),
 *   - re-indexing of a list built on the global grid to one rank's atom order (home bins, then halo bins),
 *   - the chunk plan of nbnxm_b200_do_force_step_pipelined (atoms cut into ranges of whole grid columns, the sci
 *     array grouped by the chunk of its i-atoms, per sci chunk the atom chunks its entries touch).
 * Integer bookkeeping only; no device code.
 */
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/nbnxm_b200_search.h"

namespace
{

int planFail(const char* msg)
{
    std::fprintf(stderr, "nbnxm_b200_search: %s\n", msg);
    return 1;
}

void slabColumns(int ncx, int nslabs, int r, int* c0, int* c1)
{
    *c0 = int((long long)ncx * r / nslabs);
    *c1 = int((long long)ncx * (r + 1) / nslabs);
}

} // namespace

extern "C" {

int nbnxm_b200_slab_bin_ranges(const nbnxm_b200_grid_t* grid, int nslabs, int r, float rlist, int* home_begin, int* home_end,
                               int* halo_begin, int* halo_end, int* required_tx)
{
    if (!grid || nslabs < 1 || r < 0 || r >= nslabs || !home_begin || !home_end || !halo_begin || !halo_end || !required_tx)
    {
        return planFail("slab_bin_ranges: bad argument");
    }
    int   nbins = 0, ncx = 0, ncy = 0;
    float box[3];
    nbnxm_b200_grid_info(grid, nullptr, &nbins, &ncx, &ncy);
    nbnxm_b200_grid_box(grid, box);
    if (nslabs < 2)
    {
        *home_begin  = 0;
        *home_end    = nbins;
        *halo_begin  = 0;
        *halo_end    = 0;
        *required_tx = 0;
        return 0;
    }
    std::vector<int> firstBin(size_t(ncx) * ncy + 1);
    nbnxm_b200_grid_get_order(grid, nullptr, firstBin.data());
    int cx0, cx1, nx0, nx1;
    slabColumns(ncx, nslabs, r, &cx0, &cx1);
    slabColumns(ncx, nslabs, (r + 1) % nslabs, &nx0, &nx1);
    /* the halo: the first columns of the +x neighbour within rlist of the slab boundary, plus one column of slack for
     * atoms binned by their cluster's lower corner */
    const double cell     = double(box[0]) / ncx;
    const int    ncolHalo = std::min(nx1 - nx0, int(std::ceil(double(rlist) / cell)) + 1);
    if (nslabs == 2 && (cx1 - cx0) < 2 * ncolHalo)
    {
        return planFail("slabs too thin for a one-sided halo");
    }
    *home_begin  = firstBin[size_t(cx0) * ncy];
    *home_end    = firstBin[size_t(cx1) * ncy];
    *halo_begin  = firstBin[size_t(nx0) * ncy];
    *halo_end    = firstBin[size_t(nx0 + ncolHalo) * ncy];
    *required_tx = (r == nslabs - 1) ? -1 : 0; /* home x halo pairs of the last slab cross the periodic boundary */
    return 0;
}

int nbnxm_b200_pairlist_reindex(nbnxm_b200_sci_t* sci, int nsci, nbnxm_b200_cj_packed_t* cj_packed, int ncj_packed, int first_home_bin,
                                int first_halo_bin, int num_home_bins, int nclusters_total, int halo)
{
    if ((nsci > 0 && !sci) || (ncj_packed > 0 && !cj_packed) || nclusters_total < 1) return planFail("pairlist_reindex: bad argument");
    for (int i = 0; i < nsci; i++)
    {
        sci[i].sci -= first_home_bin;
    }
    const long long offset = halo ? (long long)(num_home_bins - first_halo_bin) * 8 : -(long long)first_home_bin * 8;
    for (int g = 0; g < ncj_packed; g++)
    {
        for (int jm = 0; jm < 4; jm++)
        {
            /* unused slots of partially filled j-groups carry no mask bits and an unspecified index: keep them loadable */
            const long long cj   = (long long)cj_packed[g].cj[jm] + offset;
            cj_packed[g].cj[jm] = int(std::min<long long>(std::max<long long>(cj, 0), nclusters_total - 1));
        }
    }
    return 0;
}

int nbnxm_b200_chunk_plan(const nbnxm_b200_grid_t* grid, nbnxm_b200_sci_t* sci, int nsci, const nbnxm_b200_cj_packed_t* cj_packed,
                          int ncj_packed, int nchunks_requested, int* nchunks_out, int* first_atom, int* first_sci, unsigned int* needs)
{
    if (!grid || (nsci > 0 && !sci) || (ncj_packed > 0 && !cj_packed) || !nchunks_out || !first_atom || !first_sci || !needs)
    {
        return planFail("chunk_plan: bad argument");
    }
    int nbins = 0, ncx = 0, ncy = 0;
    nbnxm_b200_grid_info(grid, nullptr, &nbins, &ncx, &ncy);
    /* NBNXM_B200_CHUNK_WEIGHTS="w0,w1,...": relative chunk widths for A/B runs (their number replaces nchunks_requested) */
    std::vector<int> weightOverride;
    if (const char* ws = getenv("NBNXM_B200_CHUNK_WEIGHTS"))
    {
        for (const char* q = ws; *q != 0;)
        {
            char*      end = nullptr;
            const long v   = strtol(q, &end, 10);
            if (end == q) break;
            if (v > 0) weightOverride.push_back(int(v));
            q = (*end == ',') ? end + 1 : end;
        }
        if (weightOverride.size() > 32 || int(weightOverride.size()) > ncx) weightOverride.clear();
    }
    const int        nchunks = weightOverride.empty() ? std::max(1, std::min(std::min(nchunks_requested, 32), ncx)) : int(weightOverride.size());
    std::vector<int> firstBinOfColumn(size_t(ncx) * ncy + 1);
    nbnxm_b200_grid_get_order(grid, nullptr, firstBinOfColumn.data());
    /* Tapered chunk widths: the first kernel of a step cannot start before the chunks it reads have arrived, and the last
     * force copies cannot start before the last kernels have run, so both ends of the pipeline are exposed transfer time
     * (measured at 12.3 M atoms with 24 equal chunks: 0.5 ms before the first kernel, 0.6 ms of copies after the last,
     * profiles/r02i_timeline_12m_24.jsonl).  With eight chunks or more the two outermost chunks at each end of x (the
     * periodic wrap makes both ends neighbours, they are scheduled first and last) get a quarter, the next two half and three
     * quarters of the width of the chunks in the middle. */
    std::vector<int> weightSum(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++)
    {
        const int d  = std::min(c, nchunks - 1 - c);
        const int w  = nchunks < 8 ? 4 : (d < 2 ? 1 : (d == 2 ? 2 : (d == 3 ? 3 : 4)));
        weightSum[c + 1] = weightSum[c] + (weightOverride.empty() ? w : weightOverride[c]);
    }
    std::vector<int> firstBin(nchunks + 1);
    for (int c = 0; c <= nchunks; c++)
    {
        const int cx = int((long long)ncx * weightSum[c] / weightSum[nchunks]);
        firstBin[c]  = firstBinOfColumn[size_t(cx) * ncy];
        first_atom[c] = firstBin[c] * 64;
    }
    /* chunk of a bin: the last chunk starting at or below it (empty chunks share their start with the next one) */
    auto chunkOfBin = [&](int bin) { return int(std::upper_bound(firstBin.begin(), firstBin.end(), bin) - firstBin.begin()) - 1; };
    std::vector<int> chunkOfSci(nsci);
    for (int i = 0; i < nsci; i++)
    {
        chunkOfSci[i] = chunkOfBin(sci[i].sci);
    }
    for (int k = 0; k < nchunks; k++)
    {
        needs[k] = 1u << k;
    }
    /* atom chunks touched by the j-clusters of every entry (outer-list masks: a superset of what is evaluated) */
    for (int i = 0; i < nsci; i++)
    {
        const int k = chunkOfSci[i];
        if (k < 0 || k >= nchunks) return planFail("chunk_plan: sci entry outside the grid");
        for (int g = sci[i].cj_packed_begin; g < sci[i].cj_packed_end; g++)
        {
            const unsigned int any = cj_packed[g].imei[0].imask | cj_packed[g].imei[1].imask;
            for (int jm = 0; jm < 4; jm++)
            {
                if ((any >> (8 * jm)) & 0xffu)
                {
                    needs[k] |= 1u << chunkOfBin(cj_packed[g].cj[jm] / 8);
                }
            }
        }
    }
    /* the sci array grouped by chunk, entries of a chunk in list order */
    std::vector<int> order(nsci);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return chunkOfSci[a] < chunkOfSci[b]; });
    std::vector<nbnxm_b200_sci_t> sorted(nsci);
    std::vector<int>              count(nchunks + 1, 0);
    for (int i = 0; i < nsci; i++)
    {
        sorted[i] = sci[order[i]];
        count[chunkOfSci[order[i]] + 1]++;
    }
    if (nsci > 0) std::memcpy(sci, sorted.data(), sizeof(nbnxm_b200_sci_t) * nsci);
    first_sci[0] = 0;
    for (int k = 0; k < nchunks; k++)
    {
        first_sci[k + 1] = first_sci[k] + count[k + 1];
    }
    *nchunks_out = nchunks;
    return 0;
}

} // extern "C"
