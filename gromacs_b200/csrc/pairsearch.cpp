/* Host-side grid binning and GPU-layout pair-list construction (include/nbnxm_b200_search.h).
 *
 * Independent implementation of what the reference does in src/gromacs/nbnxm/grid.cpp (putOnGrid :1612,
 * sortCellsGpuGeometry :1169, target cell size :161-182) and src/gromacs/nbnxm/pairlist.cpp
 * (super-cluster search :2827-3310, make_cluster_list_supersub :813-966, self/Newton exclusion masks
 * :651-688, topology exclusions :1561-1660, splitting of i-entries :1769-1879).  Output formats are the
 * reference's (pairlist.h:189-287); the search itself is a plain bounding-box sweep over the column
 * grid, parallelised with OpenMP over i-bins.
 */
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#    include <omp.h>
#endif

#include "../../include/nbnxm_b200_search.h"
#include "gpusearch_bodies.h"

namespace
{

/* distance arithmetic shared with the GPU builder (explicit fmaf), so that both produce the same list */
using nbs::BB;
using nbs::bbDist2;

constexpr int   c_cl       = 8;  // atoms per cluster
constexpr int   c_binCl    = 8;  // clusters per bin
constexpr int   c_binAtoms = 64; // atoms per bin
constexpr int   c_central  = 22;
constexpr float c_farAway  = -1000000.0f; // filler coordinate, atomdata.cpp:171

struct ThreadList
{
    std::vector<nbnxm_b200_sci_t>       sci;
    std::vector<nbnxm_b200_cj_packed_t> cjp;
    std::vector<nbnxm_b200_excl_t>      excl;
    long long                           nClusterPairs = 0;
};

thread_local char g_err[256] = "";
int               fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "nbnxm_b200_search: %s\n", g_err);
    return 1;
}

} // namespace

struct nbnxm_b200_grid
{
    float box[3];
    int   natoms = 0;
    int   ncx = 1, ncy = 1;
    float cellSize[2];
    int   nbins = 0;
    std::vector<int>   colFirstBin; // ncx*ncy + 1
    std::vector<int>   atomIndex;   // slot -> atom
    std::vector<int>   slotOfAtom;  // atom -> slot
    std::vector<float> xs;          // slot-ordered coordinates (x,y,z), fillers far away
    std::vector<BB>    clBB;        // per cluster
    std::vector<int>   clCount;     // real atoms per cluster
    std::vector<BB>    binBB;

    // last list built
    std::vector<nbnxm_b200_sci_t>       sci;
    std::vector<nbnxm_b200_cj_packed_t> cjp;
    std::vector<nbnxm_b200_excl_t>      excl;
    long long                           nClusterPairs = 0;

    // perturbed atom-pair list split off the list above (nbnxm_b200_pairlist_split_fep)
    std::vector<int>           fepIinr, fepShift, fepJindex, fepJjnr;
    std::vector<unsigned char> fepInteracts;
    float                      lastRlist    = 0.0f;  // radius of the list built last
    bool                       fepSplitDone = false; // the split clears bits of the list: it must not run twice on one list
};

extern "C" {

int nbnxm_b200_grid_dims(const float* box, int natoms, int nslabs, int* ncx, int* ncy)
{
    if (!box || !ncx || !ncy || natoms <= 0 || nslabs < 1) return fail("grid_dims: bad argument");
    /* approximately cubic clusters of 8 atoms, 2x2 of them per column cross-section; round the column
     * count down (grid.cpp:161-182, 296-311) */
    const double density = natoms / (double(box[0]) * box[1] * box[2]);
    const double tlen    = std::cbrt(c_cl / density);
    *ncx                 = std::max(1, int(box[0] / (2 * tlen)));
    if (nslabs > 1)
    {
        /* x-slab decomposition over nslabs GPUs: a whole number of columns per slab, so that the slabs hold equal
         * numbers of atoms (the reference grids every domain separately) */
        *ncx = std::max(nslabs, (*ncx / nslabs) * nslabs);
    }
    *ncy = std::max(1, int(box[1] / (2 * tlen)));
    return 0;
}

int nbnxm_b200_grid_create(nbnxm_b200_grid_t** out, const float* box, int natoms, const float* x, int nthreads)
{
    return nbnxm_b200_grid_create_slabs(out, box, natoms, x, nthreads, 1);
}

int nbnxm_b200_grid_create_slabs(nbnxm_b200_grid_t** out, const float* box, int natoms, const float* x, int nthreads, int nslabs)
{
    if (!out || !box || !x || natoms <= 0 || nslabs < 1) return fail("grid_create: bad argument");
    if (nthreads < 1) nthreads = 1;
    nbnxm_b200_grid* g = new nbnxm_b200_grid();
    for (int d = 0; d < 3; d++) g->box[d] = box[d];
    g->natoms = natoms;
    nbnxm_b200_grid_dims(box, natoms, nslabs, &g->ncx, &g->ncy);
    g->cellSize[0]       = box[0] / g->ncx;
    g->cellSize[1]       = box[1] / g->ncy;
    const int ncol       = g->ncx * g->ncy;

    std::vector<int> colOfAtom(natoms), colCount(ncol, 0);
    for (int a = 0; a < natoms; a++)
    {
        int cx = int(x[3 * a] / g->cellSize[0]);
        int cy = int(x[3 * a + 1] / g->cellSize[1]);
        cx     = std::min(std::max(cx, 0), g->ncx - 1);
        cy     = std::min(std::max(cy, 0), g->ncy - 1);
        colOfAtom[a] = cx * g->ncy + cy; // x-major, grid.h:100
        colCount[colOfAtom[a]]++;
    }
    g->colFirstBin.assign(ncol + 1, 0);
    std::vector<int> colAtomStart(ncol + 1, 0);
    for (int c = 0; c < ncol; c++)
    {
        g->colFirstBin[c + 1] = g->colFirstBin[c] + (colCount[c] + c_binAtoms - 1) / c_binAtoms;
        colAtomStart[c + 1]   = colAtomStart[c] + colCount[c];
    }
    g->nbins = g->colFirstBin[ncol];
    std::vector<int> colAtoms(natoms), fill(ncol, 0);
    for (int a = 0; a < natoms; a++)
    {
        const int c                               = colOfAtom[a];
        colAtoms[colAtomStart[c] + fill[c]++] = a;
    }
    const int nslots = g->nbins * c_binAtoms;
    g->atomIndex.assign(nslots, -1);
    g->slotOfAtom.assign(natoms, -1);

#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int c = 0; c < ncol; c++)
    {
        int*      a0 = colAtoms.data() + colAtomStart[c];
        const int n  = colCount[c];
        auto      byDim = [&](int dim, bool descending) {
            return [=](int p, int q) {
                const float vp = x[3 * p + dim], vq = x[3 * q + dim];
                if (vp != vq) return descending ? vp > vq : vp < vq;
                return p < q;
            };
        };
        std::sort(a0, a0 + n, byDim(2, false));
        /* per bin: halves along z (already sorted), then y, then x, snaking the sort direction so that
         * consecutive clusters stay adjacent in space (cf. grid.cpp:1215-1290) */
        for (int s = 0; s < n; s += 32)
        {
            const int nz  = std::min(32, n - s);
            const int iz  = s / 32;
            std::sort(a0 + s, a0 + s + nz, byDim(1, (iz & 1) != 0));
            for (int t = 0; t < nz; t += 16)
            {
                const int ny = std::min(16, nz - t);
                const int iy = (s + t) / 16;
                std::sort(a0 + s + t, a0 + s + t + ny, byDim(0, (iy & 1) != 0));
            }
        }
        const int slot0 = g->colFirstBin[c] * c_binAtoms;
        for (int i = 0; i < n; i++)
        {
            g->atomIndex[slot0 + i] = a0[i];
            g->slotOfAtom[a0[i]]    = slot0 + i;
        }
    }
    /* slot-ordered coordinates and bounding boxes (real atoms only, atomdata.cpp:159-171) */
    g->xs.assign(size_t(nslots) * 3, c_farAway);
    g->clBB.resize(size_t(g->nbins) * c_binCl);
    g->clCount.assign(size_t(g->nbins) * c_binCl, 0);
    g->binBB.resize(g->nbins);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int b = 0; b < g->nbins; b++)
    {
        BB bb;
        for (int d = 0; d < 3; d++)
        {
            bb.lo[d] = 1e30f;
            bb.hi[d] = -1e30f;
        }
        for (int cl = 0; cl < c_binCl; cl++)
        {
            BB cb = bb;
            for (int d = 0; d < 3; d++)
            {
                cb.lo[d] = 1e30f;
                cb.hi[d] = -1e30f;
            }
            int cnt = 0;
            for (int i = 0; i < c_cl; i++)
            {
                const int slot = (b * c_binCl + cl) * c_cl + i;
                const int a    = g->atomIndex[slot];
                if (a >= 0)
                {
                    cnt++;
                    for (int d = 0; d < 3; d++)
                    {
                        const float v      = x[3 * a + d];
                        g->xs[3 * slot + d] = v;
                        cb.lo[d]           = std::min(cb.lo[d], v);
                        cb.hi[d]           = std::max(cb.hi[d], v);
                    }
                }
            }
            g->clBB[b * c_binCl + cl]    = cb;
            g->clCount[b * c_binCl + cl] = cnt;
            if (cnt)
            {
                for (int d = 0; d < 3; d++)
                {
                    bb.lo[d] = std::min(bb.lo[d], cb.lo[d]);
                    bb.hi[d] = std::max(bb.hi[d], cb.hi[d]);
                }
            }
        }
        g->binBB[b] = bb;
    }
    *out = g;
    return 0;
}

int nbnxm_b200_grid_free(nbnxm_b200_grid_t* g)
{
    delete g;
    return 0;
}

int nbnxm_b200_grid_info(const nbnxm_b200_grid_t* g, int* natoms_nbat, int* nbins, int* ncx, int* ncy)
{
    if (!g) return fail("null grid");
    if (natoms_nbat) *natoms_nbat = g->nbins * c_binAtoms;
    if (nbins) *nbins = g->nbins;
    if (ncx) *ncx = g->ncx;
    if (ncy) *ncy = g->ncy;
    return 0;
}

int nbnxm_b200_grid_box(const nbnxm_b200_grid_t* g, float* box)
{
    if (!g || !box) return fail("null grid");
    for (int d = 0; d < 3; d++) box[d] = g->box[d];
    return 0;
}

int nbnxm_b200_grid_get_order(const nbnxm_b200_grid_t* g, int* atom_index, int* first_bin_of_column)
{
    if (!g) return fail("null grid");
    if (atom_index) std::memcpy(atom_index, g->atomIndex.data(), sizeof(int) * g->atomIndex.size());
    if (first_bin_of_column) std::memcpy(first_bin_of_column, g->colFirstBin.data(), sizeof(int) * g->colFirstBin.size());
    return 0;
}

int nbnxm_b200_grid_fill_atomdata(const nbnxm_b200_grid_t* g, const float* x, const float* q, const int* type, int ntypes,
                                  const float* lj_comb_per_type, float* xq, int* type_nbat, float* lj_comb)
{
    if (!g) return fail("null grid");
    const int nslots = g->nbins * c_binAtoms;
#pragma omp parallel for schedule(static)
    for (int s = 0; s < nslots; s++)
    {
        const int a = g->atomIndex[s];
        if (xq)
        {
            xq[4 * s]     = a >= 0 ? x[3 * a] : c_farAway;
            xq[4 * s + 1] = a >= 0 ? x[3 * a + 1] : c_farAway;
            xq[4 * s + 2] = a >= 0 ? x[3 * a + 2] : c_farAway;
            xq[4 * s + 3] = (a >= 0 && q) ? q[a] : 0.0f;
        }
        const int t = (a >= 0 && type) ? type[a] : ntypes - 1; // filler type: all-zero LJ row, atomdata.cpp:505-520
        if (type_nbat) type_nbat[s] = t;
        if (lj_comb && lj_comb_per_type)
        {
            lj_comb[2 * s]     = lj_comb_per_type[2 * t];
            lj_comb[2 * s + 1] = lj_comb_per_type[2 * t + 1];
        }
    }
    return 0;
}

int nbnxm_b200_pairlist_build(nbnxm_b200_grid_t* g, float rlist, const int* excl_index, const int* excl_atoms, int min_sci,
                              int bin_begin, int bin_end, int j_bin_lo, int j_bin_hi, int inter_zone, int required_tx,
                              int nthreads)
{
    if (!g) return fail("null grid");
    if (bin_begin < 0 || bin_end > g->nbins || j_bin_lo < 0 || j_bin_hi > g->nbins) return fail("pairlist_build: bin range");
    for (int d = 0; d < 3; d++)
    {
        if (2 * rlist >= g->box[d]) return fail("pairlist_build: rlist %g must be shorter than half the box (%g)", rlist, g->box[d]);
    }
    if (nthreads < 1) nthreads = 1;
    g->lastRlist    = rlist;
    g->fepSplitDone = false;
    const float rl2 = rlist * rlist;
    /* bounding-box-only acceptance distance: rlist minus half the average x/y diagonal of a cluster
     * (pairlist.cpp: boundingbox_only_distance2) */
    const float rbb2 = nbs::bbOnlyDistance2(g->cellSize, rlist);
    const int   ncol = g->ncx * g->ncy;

    std::vector<int> colOfBin(g->nbins);
    for (int c = 0; c < ncol; c++)
    {
        for (int b = g->colFirstBin[c]; b < g->colFirstBin[c + 1]; b++) colOfBin[b] = c;
    }

    std::vector<ThreadList> tl(nthreads);
    const int               nIBins = bin_end - bin_begin;

#pragma omp parallel num_threads(nthreads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        ThreadList& L = tl[tid];
        L.excl.emplace_back();
        for (unsigned& p : L.excl[0].pair) p = 0xffffffffu;
        std::vector<int> posOfCluster(size_t(g->nbins) * c_binCl, -1);
        std::vector<int> jList;       // j-clusters of the current entry
        std::vector<unsigned> jMask;  // their i-cluster masks

        const int chunk0 = bin_begin + int((long long)nIBins * tid / nthreads);
        const int chunk1 = bin_begin + int((long long)nIBins * (tid + 1) / nthreads);
        for (int bi = chunk0; bi < chunk1; bi++)
        {
            const BB& ibb = g->binBB[bi];
            if (ibb.lo[0] > ibb.hi[0]) continue; // no real atoms
            for (int tz = -1; tz <= 1; tz++)
            {
                for (int ty = -1; ty <= 1; ty++)
                {
                    for (int tx = -1; tx <= 1; tx++)
                    {
                        const int shift = ((tz + 1) * 3 + (ty + 1)) * 5 + (tx + 2); // pbcutil/ishift.h
                        if (inter_zone)
                        {
                            if (tx != required_tx) continue;
                        }
                        else if (shift > c_central)
                        {
                            continue; // half shell: backward shifts only
                        }
                        const float sh[3] = { tx * g->box[0], ty * g->box[1], tz * g->box[2] };
                        /* quick reject against the unit cell */
                        bool out = false;
                        for (int d = 0; d < 3; d++)
                        {
                            if (ibb.lo[d] + sh[d] - rlist > g->box[d] || ibb.hi[d] + sh[d] + rlist < 0) out = true;
                        }
                        if (out) continue;

                        jList.clear();
                        jMask.clear();
                        const bool subDiag = (!inter_zone && shift == c_central);

                        auto scanBin = [&](int bj) {
                            if (bj < j_bin_lo || bj >= j_bin_hi) return;
                            const BB& jbb = g->binBB[bj];
                            if (jbb.lo[0] > jbb.hi[0]) return;
                            if (bbDist2(ibb, sh, jbb) >= rl2) return;
                            for (int cj = 0; cj < c_binCl; cj++)
                            {
                                const int gcj = bj * c_binCl + cj;
                                if (g->clCount[gcj] == 0) continue;
                                if (bbDist2(ibb, sh, g->clBB[gcj]) >= rl2) continue;
                                unsigned mask = 0;
                                for (int ci = 0; ci < c_binCl; ci++)
                                {
                                    const int gci = bi * c_binCl + ci;
                                    if (g->clCount[gci] == 0) continue;
                                    if (subDiag && bj == bi && ci > cj) continue;
                                    const float d2 = bbDist2(g->clBB[gci], sh, g->clBB[gcj]);
                                    if (d2 >= rl2) continue;
                                    bool in = d2 < rbb2;
                                    if (!in)
                                    {
                                        /* atom-pair check, like clusterpair_in_range */
                                        for (int i = 0; i < c_cl && !in; i++)
                                        {
                                            const float* xi = &g->xs[3 * (gci * c_cl + i)];
                                            if (xi[0] == c_farAway) continue;
                                            const float px = xi[0] + sh[0], py = xi[1] + sh[1], pz = xi[2] + sh[2];
                                            for (int j = 0; j < c_cl; j++)
                                            {
                                                const float* xj = &g->xs[3 * (gcj * c_cl + j)];
                                                const float  dx = px - xj[0], dy = py - xj[1], dz = pz - xj[2];
                                                if (nbs::dist2(dx, dy, dz) < rl2)
                                                {
                                                    in = true;
                                                    break;
                                                }
                                            }
                                        }
                                    }
                                    if (in) mask |= 1u << ci;
                                }
                                if (mask)
                                {
                                    jList.push_back(gcj);
                                    jMask.push_back(mask);
                                    L.nClusterPairs += __builtin_popcount(mask);
                                }
                            }
                        };

                        /* own bin first so that the diagonal cluster leads the central-shift entry
                         * (the force kernel keys the self-energy term on that, nbnxm_cuda_kernel.cuh:384) */
                        if (subDiag) scanBin(bi);
                        const float xlo = ibb.lo[0] + sh[0] - rlist, xhi = ibb.hi[0] + sh[0] + rlist;
                        const float ylo = ibb.lo[1] + sh[1] - rlist, yhi = ibb.hi[1] + sh[1] + rlist;
                        const int   cx0 = std::max(0, int(std::floor(xlo / g->cellSize[0])));
                        const int   cx1 = std::min(g->ncx - 1, int(std::floor(xhi / g->cellSize[0])));
                        const int   cy0 = std::max(0, int(std::floor(ylo / g->cellSize[1])));
                        const int   cy1 = std::min(g->ncy - 1, int(std::floor(yhi / g->cellSize[1])));
                        for (int cx = cx0; cx <= cx1; cx++)
                        {
                            for (int cy = cy0; cy <= cy1; cy++)
                            {
                                const int c = cx * g->ncy + cy;
                                for (int bj = g->colFirstBin[c]; bj < g->colFirstBin[c + 1]; bj++)
                                {
                                    if (subDiag && bj <= bi) continue; // own bin done, lower bins own the pair
                                    const float dzl = (ibb.lo[2] + sh[2]) - g->binBB[bj].hi[2];
                                    const float dzh = g->binBB[bj].lo[2] - (ibb.hi[2] + sh[2]);
                                    if (dzl >= rlist) continue;
                                    if (dzh >= rlist) break; // bins are sorted along z
                                    scanBin(bj);
                                }
                            }
                        }
                        if (jList.empty()) continue;

                        /* ---- close the entry: pack j-clusters in groups of 4 ---- */
                        const int nj      = int(jList.size());
                        const int nGroups = (nj + 3) / 4;
                        const int cjp0    = int(L.cjp.size());
                        L.cjp.resize(cjp0 + nGroups);
                        for (int gidx = 0; gidx < nGroups; gidx++)
                        {
                            nbnxm_b200_cj_packed_t& e = L.cjp[cjp0 + gidx];
                            std::memset(&e, 0, sizeof(e));
                            unsigned im = 0;
                            for (int jm = 0; jm < 4; jm++)
                            {
                                const int k = gidx * 4 + jm;
                                if (k < nj)
                                {
                                    e.cj[jm] = jList[k];
                                    im |= jMask[k] << (jm * 8);
                                    posOfCluster[jList[k]] = k;
                                }
                            }
                            e.imei[0].imask = im;
                            e.imei[1].imask = im; // both halves start identical, pairlist.cpp:951-954
                        }
                        auto exclMask = [&](int k, int half) -> nbnxm_b200_excl_t& {
                            nbnxm_b200_cj_packed_t& e = L.cjp[cjp0 + k / 4];
                            if (e.imei[half].excl_ind == 0)
                            {
                                e.imei[half].excl_ind = int(L.excl.size());
                                L.excl.emplace_back();
                                for (unsigned& p : L.excl.back().pair) p = 0xffffffffu;
                            }
                            return L.excl[e.imei[half].excl_ind];
                        };
                        /* self + Newton exclusions on diagonal cluster pairs: keep only j > i
                         * (setSelfAndNewtonExclusionsGpu, pairlist.cpp:651-688) */
                        if (subDiag)
                        {
                            for (int k = 0; k < nj; k++)
                            {
                                const int gcj = jList[k];
                                if (gcj / c_binCl != bi) continue;
                                const int ci = gcj % c_binCl;
                                if (!(jMask[k] & (1u << ci))) continue;
                                const unsigned bit = 1u << ((k & 3) * 8 + ci);
                                for (int half = 0; half < 2; half++)
                                {
                                    nbnxm_b200_excl_t& ex = exclMask(k, half);
                                    for (int jj = 0; jj < 4; jj++)
                                    {
                                        const int ja = half * 4 + jj;
                                        for (int ia = 0; ia < c_cl; ia++)
                                        {
                                            if (ja <= ia) ex.pair[jj * c_cl + ia] &= ~bit;
                                        }
                                    }
                                }
                            }
                        }
                        /* topology exclusions (setExclusionsForIEntry, pairlist.cpp:1561-1660) */
                        if (excl_index && excl_atoms)
                        {
                            for (int i = 0; i < c_binAtoms; i++)
                            {
                                const int islot = bi * c_binAtoms + i;
                                const int ia    = g->atomIndex[islot];
                                if (ia < 0) continue;
                                for (int e = excl_index[ia]; e < excl_index[ia + 1]; e++)
                                {
                                    const int ja = excl_atoms[e];
                                    if (ja == ia) continue;
                                    const int jslot = g->slotOfAtom[ja];
                                    if (subDiag && jslot <= islot) continue;
                                    const int gcj = jslot / c_cl;
                                    const int k   = posOfCluster[gcj];
                                    if (k < 0) continue;
                                    const int      ci  = i / c_cl;
                                    const unsigned bit = 1u << ((k & 3) * 8 + ci);
                                    if (!(L.cjp[cjp0 + k / 4].imei[0].imask & bit)) continue;
                                    const int jin  = jslot & (c_cl - 1);
                                    const int half = jin / 4;
                                    exclMask(k, half).pair[(jin & 3) * c_cl + (i & (c_cl - 1))] &= ~bit;
                                }
                            }
                        }
                        for (int k = 0; k < nj; k++) posOfCluster[jList[k]] = -1;
                        nbnxm_b200_sci_t s;
                        s.sci             = bi;
                        s.shift           = shift;
                        s.cj_packed_begin = cjp0;
                        s.cj_packed_end   = cjp0 + nGroups;
                        L.sci.push_back(s);
                    }
                }
            }
        }
    }

    /* ---- combine the per-thread lists (combine_nblists, pairlist.cpp:2314) ---- */
    size_t nsci = 0, ncjp = 0, nexcl = 1;
    for (const ThreadList& L : tl)
    {
        nsci += L.sci.size();
        ncjp += L.cjp.size();
        nexcl += L.excl.empty() ? 0 : L.excl.size() - 1;
    }
    /* split long i-entries so that about min_sci entries exist (split_sci_entry, pairlist.cpp:1769-1879) */
    int maxGroups = 1 << 30;
    if (min_sci > 0 && ncjp > 0)
    {
        maxGroups = std::max(1, int((ncjp + min_sci - 1) / min_sci));
    }
    g->sci.clear();
    g->cjp.resize(ncjp);
    g->excl.resize(nexcl);
    for (unsigned& p : g->excl[0].pair) p = 0xffffffffu;
    g->nClusterPairs = 0;
    size_t cjOff = 0, exOff = 1;
    for (const ThreadList& L : tl)
    {
        for (size_t i = 0; i < L.cjp.size(); i++)
        {
            nbnxm_b200_cj_packed_t e = L.cjp[i];
            for (int h = 0; h < 2; h++)
            {
                if (e.imei[h].excl_ind != 0) e.imei[h].excl_ind += int(exOff) - 1;
            }
            g->cjp[cjOff + i] = e;
        }
        for (size_t i = 1; i < L.excl.size(); i++) g->excl[exOff + i - 1] = L.excl[i];
        for (const nbnxm_b200_sci_t& s : L.sci)
        {
            for (int b = s.cj_packed_begin; b < s.cj_packed_end; b += maxGroups)
            {
                nbnxm_b200_sci_t t = s;
                t.cj_packed_begin  = int(cjOff) + b;
                t.cj_packed_end    = int(cjOff) + std::min(b + maxGroups, s.cj_packed_end);
                g->sci.push_back(t);
            }
        }
        cjOff += L.cjp.size();
        exOff += L.excl.empty() ? 0 : L.excl.size() - 1;
        g->nClusterPairs += L.nClusterPairs;
    }
    return 0;
}

int nbnxm_b200_pairlist_sizes(const nbnxm_b200_grid_t* g, int* nsci, int* ncj_packed, int* nexcl, long long* ncluster_pairs)
{
    if (!g) return fail("null grid");
    if (nsci) *nsci = int(g->sci.size());
    if (ncj_packed) *ncj_packed = int(g->cjp.size());
    if (nexcl) *nexcl = int(g->excl.size());
    if (ncluster_pairs) *ncluster_pairs = g->nClusterPairs;
    return 0;
}

int nbnxm_b200_pairlist_copy(const nbnxm_b200_grid_t* g, nbnxm_b200_sci_t* sci, nbnxm_b200_cj_packed_t* cj_packed,
                             nbnxm_b200_excl_t* excl)
{
    if (!g) return fail("null grid");
    if (sci && !g->sci.empty()) std::memcpy(sci, g->sci.data(), sizeof(*sci) * g->sci.size());
    if (cj_packed && !g->cjp.empty()) std::memcpy(cj_packed, g->cjp.data(), sizeof(*cj_packed) * g->cjp.size());
    if (excl && !g->excl.empty()) std::memcpy(excl, g->excl.data(), sizeof(*excl) * g->excl.size());
    return 0;
}

/* make_fep_list for the GPU layout (src/gromacs/nbnxm/pairlist.cpp:1414): every atom pair of the list built last in
 * which at least one atom is perturbed moves to an atom-pair list (i-atom, shift, j-atoms, "interacts" = its exclusion
 * bit) and its bit in the cluster list is cleared.  The caller masks the perturbed atoms in the cluster kernels' atom
 * data (charge 0, the non-interacting type: nbnxm_atomdata_mask_fep, atomdata.cpp:1039), so that a cleared bit
 * contributes nothing there, and evaluates the atom-pair list with the perturbed kernel
 * (nbnxm_b200_launch_free_energy_kernel).  Self pairs of perturbed atoms are listed as excluded pairs (the kernel
 * halves their reaction-field / Ewald correction). */
int nbnxm_b200_pairlist_split_fep(nbnxm_b200_grid_t* g, const unsigned char* perturbed)
{
    if (!g || !perturbed) return fail("pairlist_split_fep: null argument");
    if (g->fepSplitDone)
    {
        /* the bits of the moved pairs are gone from the cluster list: a second pass would list them as excluded pairs */
        return fail("pairlist_split_fep: the list built last has already been split; build a new list first");
    }
    g->fepSplitDone = true;
    /* like make_fep_list (pairlist.cpp:1519, rlist_fep2), interacting pairs beyond the list radius leave the cluster list
     * without entering the perturbed one: they are beyond the cut-off for the life of the list */
    const float rlist2 = g->lastRlist * g->lastRlist;
    g->fepIinr.clear();
    g->fepShift.clear();
    g->fepJjnr.clear();
    g->fepInteracts.clear();
    g->fepJindex.assign(1, 0);
    std::vector<int>           jOfI[c_binAtoms];
    std::vector<unsigned char> intOfI[c_binAtoms];
    /* exclusion entry of (group, half), private to the group so that bits can be cleared */
    auto ownExcl = [&](int group, int half) -> nbnxm_b200_excl_t& {
        nbnxm_b200_cj_packed_t& e = g->cjp[group];
        if (e.imei[half].excl_ind == 0)
        {
            e.imei[half].excl_ind = int(g->excl.size());
            g->excl.emplace_back();
            for (unsigned& w : g->excl.back().pair) w = 0xffffffffu;
        }
        return g->excl[e.imei[half].excl_ind];
    };
    for (const nbnxm_b200_sci_t& s : g->sci)
    {
        const int  bi      = s.sci;
        const bool central = (s.shift == c_central);
        bool       any     = false;
        const float sh[3]  = { float(s.shift % 5 - 2) * g->box[0], float((s.shift / 5) % 3 - 1) * g->box[1],
                               float(s.shift / 15 - 1) * g->box[2] }; // pbcutil/ishift.h
        for (int i = 0; i < c_binAtoms; i++)
        {
            jOfI[i].clear();
            intOfI[i].clear();
        }
        for (int group = s.cj_packed_begin; group < s.cj_packed_end; group++)
        {
            for (int jm = 0; jm < 4; jm++)
            {
                const unsigned imask = g->cjp[group].imei[0].imask;
                if (((imask >> (jm * 8)) & 0xffu) == 0) continue;
                const int gcj = g->cjp[group].cj[jm];
                for (int ci = 0; ci < c_binCl; ci++)
                {
                    const unsigned bit = 1u << (jm * 8 + ci);
                    if (!(imask & bit)) continue;
                    const int  gci      = bi * c_binCl + ci;
                    const bool diagonal = central && gci == gcj;
                    for (int ia = 0; ia < c_cl; ia++)
                    {
                        const int islot = gci * c_cl + ia;
                        const int ai    = g->atomIndex[islot];
                        if (ai < 0) continue;
                        for (int ja = 0; ja < c_cl; ja++)
                        {
                            const int jslot = gcj * c_cl + ja;
                            const int aj    = g->atomIndex[jslot];
                            if (aj < 0 || !(perturbed[ai] || perturbed[aj])) continue;
                            if (diagonal && ja < ia) continue; /* the mirrored pair owns it */
                            const int half = ja / 4;
                            const int word = (ja & 3) * c_cl + ia;
                            const int iloc = ci * c_cl + ia;
                            if (diagonal && ja == ia)
                            {
                                /* self pair: excluded by construction, listed for its exclusion correction */
                                jOfI[iloc].push_back(jslot);
                                intOfI[iloc].push_back(0);
                                any = true;
                                continue;
                            }
                            const int ex = g->cjp[group].imei[half].excl_ind;
                            const bool interacts = (g->excl[ex].pair[word] & bit) != 0;
                            if (interacts)
                            {
                                ownExcl(group, half).pair[word] &= ~bit;
                                const float dx = g->xs[3 * islot] + sh[0] - g->xs[3 * jslot];
                                const float dy = g->xs[3 * islot + 1] + sh[1] - g->xs[3 * jslot + 1];
                                const float dz = g->xs[3 * islot + 2] + sh[2] - g->xs[3 * jslot + 2];
                                /* the device form of the split (gpusearch_bodies.h, fepPairsOfIAtom) decides on the same bits */
                                if (!(nbs::dist2(dx, dy, dz) < rlist2)) continue;
                            }
                            jOfI[iloc].push_back(jslot);
                            intOfI[iloc].push_back(interacts ? 1 : 0);
                            any = true;
                        }
                    }
                }
            }
        }
        if (!any) continue;
        for (int i = 0; i < c_binAtoms; i++)
        {
            if (jOfI[i].empty()) continue;
            g->fepIinr.push_back(bi * c_binAtoms + i);
            g->fepShift.push_back(s.shift);
            g->fepJjnr.insert(g->fepJjnr.end(), jOfI[i].begin(), jOfI[i].end());
            g->fepInteracts.insert(g->fepInteracts.end(), intOfI[i].begin(), intOfI[i].end());
            g->fepJindex.push_back(int(g->fepJjnr.size()));
        }
    }
    return 0;
}

int nbnxm_b200_pairlist_fep_sizes(const nbnxm_b200_grid_t* g, int* num_i, int* num_j)
{
    if (!g) return fail("null grid");
    if (num_i) *num_i = int(g->fepIinr.size());
    if (num_j) *num_j = int(g->fepJjnr.size());
    return 0;
}

int nbnxm_b200_pairlist_fep_copy(const nbnxm_b200_grid_t* g, int* iinr, int* jindex, int* jjnr, int* shift, unsigned char* interacts)
{
    if (!g) return fail("null grid");
    auto cp = [](auto* dst, const auto& v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(v[0]) * v.size());
    };
    cp(iinr, g->fepIinr);
    cp(jindex, g->fepJindex);
    cp(jjnr, g->fepJjnr);
    cp(shift, g->fepShift);
    cp(interacts, g->fepInteracts);
    return 0;
}

} // extern "C"
