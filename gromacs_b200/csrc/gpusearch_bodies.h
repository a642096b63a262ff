/* Per-item bodies of the search step on the GPU (nbnxm_gpusearch.cu), written as host+device functions.
 *
 * Gridding (second half of this file) and pair-list construction (first half) are sequences of data-parallel passes
 * with a prefix sum between them.  Three kinds of pass:
 *   item passes   - one thread per work item, no cooperation: the body is a plain function of the item index;
 *   block passes  - one block per column (the column sort): stages of independent items separated by block barriers,
 *                   the body is a function of (block, stage, item) on the block's scratch memory;
 *   warp passes   - one warp per item (the cooperative form of the mask pass): per-lane functions joined by warp votes.
 * The CUDA kernels are thin wrappers (index arithmetic, barriers, votes); tests/kernel_emu runs the same bodies in
 * host loops so that the search logic is checked on a machine without a GPU (test infrastructure, not a product path:
 * the library itself only ever launches the kernels).
 *
 * What is built is what the reference's CPU search hands to gpu_init_atomdata / gpu_init_pairlist, in its formats
 * (src/gromacs/nbnxm/pairlist.h:189-287): atoms binned and sorted into columns, bins and clusters (Grid::putOnGrid,
 * grid.cpp:1612; sortCellsGpuGeometry :1169), super-cluster entries from a bounding-box sweep over the column grid
 * (pairlist.cpp:2827-3310), cluster-pair masks with the bounding-box / atom-pair distance test
 * (make_cluster_list_supersub, pairlist.cpp:813-966), self + Newton exclusions on the diagonal (:651-688), topology
 * exclusions (:1561-1660) and splitting of long i-entries (:1769-1879).
 *
 * Arithmetic is spelled with explicit fmaf so that the host builder (pairsearch.cpp), the emulation and the device
 * produce the same bits; grid order and lists are then equal entry for entry.
 *
 * Work items of the list passes:
 *   entry slot e = (bin - binBegin) * 27 + s, s = (tz+1)*9 + (ty+1)*3 + (tx+1): one candidate i-entry (bin, shift);
 *   bin pair p: (entry, j-bin) whose bounding boxes are within rlist, listed per entry in ascending j-bin order
 *   (own bin first on the central shift), so that the j-clusters of an entry are sorted by cluster index.
 */
#ifndef NBNXM_B200_GPUSEARCH_BODIES_H
#define NBNXM_B200_GPUSEARCH_BODIES_H

#include <math.h>
#include <string.h>

#include "../../include/nbnxm_b200.h"

#if defined(__CUDACC__)
#    define NBS_HD __host__ __device__ __forceinline__
#else
#    define NBS_HD inline
#endif

namespace nbs
{

constexpr int   c_cl        = 8;  /* atoms per cluster */
constexpr int   c_binCl     = 8;  /* clusters per bin (super-cluster) */
constexpr int   c_binAtoms  = 64; /* atoms per bin */
constexpr int   c_central   = 22; /* central shift index, pbcutil/ishift.h */
constexpr int   c_numSlots  = 27; /* shift candidates per bin */
constexpr float c_farAway   = -1000000.0f; /* filler coordinate, atomdata.cpp:171 */

struct BB
{
    float lo[3], hi[3];
};

struct alignas(16) XQ
{
    float x, y, z, q;
};

/* squared distance between bounding box a (shifted by sh) and b; 0 when they overlap */
NBS_HD float bbDist2(const BB& a, const float* sh, const BB& b)
{
    float d2 = 0.0f;
    for (int d = 0; d < 3; d++)
    {
        const float dl = (a.lo[d] + sh[d]) - b.hi[d];
        const float dh = b.lo[d] - (a.hi[d] + sh[d]);
        const float dm = fmaxf(fmaxf(dl, dh), 0.0f);
        d2             = fmaf(dm, dm, d2);
    }
    return d2;
}

/* square of the bounding-box distance below which a cluster pair is accepted without looking at the atoms: rlist minus
 * half the average x/y diagonal of a cluster (boundingbox_only_distance2, pairlist.cpp); host side */
inline float bbOnlyDistance2(const float* cellSize, float rlist)
{
    const float bbx = 0.5f * cellSize[0];
    const float bby = 0.5f * cellSize[1];
    const float rbb = fmaxf(0.0f, rlist - 0.5f * sqrtf(fmaf(bby, bby, bbx * bbx)));
    return rbb * rbb;
}

NBS_HD float dist2(float dx, float dy, float dz)
{
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* is any atom pair of i-cluster gci (shifted) and j-cluster gcj within rl2? (clusterpair_in_range, pairlist.cpp) */
NBS_HD bool clusterPairInRange(const XQ* xq, int gci, int gcj, const float* sh, float rl2)
{
    for (int i = 0; i < c_cl; i++)
    {
        const XQ xi = xq[gci * c_cl + i];
        if (xi.x == c_farAway)
        {
            continue;
        }
        const float px = xi.x + sh[0], py = xi.y + sh[1], pz = xi.z + sh[2];
        for (int j = 0; j < c_cl; j++)
        {
            const XQ xj = xq[gcj * c_cl + j];
            if (dist2(px - xj.x, py - xj.y, pz - xj.z) < rl2)
            {
                return true;
            }
        }
    }
    return false;
}

struct Grid
{
    float      box[3];
    float      cellSize[2];
    int        ncx, ncy, nbins, natoms;
    const int* colFirstBin; /* ncx*ncy + 1 */
    const int* atomIndex;   /* nbat slot -> atom, -1 = filler */
    const int* slotOfAtom;  /* atom -> nbat slot */
    const XQ*  xq;          /* nbat order */
    BB*        clBB;        /* per cluster */
    int*       clCount;     /* real atoms per cluster */
    BB*        binBB;       /* per bin, lo > hi when the bin holds no atom */
    const int* exclIndex;   /* topology exclusions, CSR in atom order; may be null */
    const int* exclAtoms;
};

struct Params
{
    float rlist, rl2, rbb2;
    int   binBegin, binEnd; /* i-bins */
    int   jBinLo, jBinHi;   /* j-bins */
    int   interZone, requiredTx;
    int   maxGroups;        /* cjPacked groups per sci entry after splitting */
};

struct Work
{
    int*           entryNumBinPairs; /* per entry slot (+1) */
    int*           entryBinPairOff;  /* exclusive scan of the above (+1: total) */
    int*           binPairJ;         /* per bin pair: j-bin */
    int*           binPairEntry;     /* per bin pair: entry slot */
    unsigned char* binPairMask;      /* per (bin pair, j-cluster): i-cluster mask */
    int*           entryNumJ;        /* j-clusters per entry */
    int*           entryGroups;      /* cjPacked groups per entry (+1) */
    int*           entryCjOff;       /* scan */
    int*           entryNumSci;      /* sci entries per entry after splitting (+1) */
    int*           entrySciOff;      /* scan */
    int*           entryNonEmpty;    /* 1 when the entry has j-clusters (+1) */
    int*           entryCompactOff;  /* scan */
    int*           compactEntry;     /* compact index -> entry slot */
    unsigned long long*     numClusterPairs;
    nbnxm_b200_sci_t*       sci;
    nbnxm_b200_cj_packed_t* cjp;
    nbnxm_b200_excl_t*      excl;
    int*                    exclFlag; /* per (cjPacked, half) (+1) */
    int*                    exclOff;  /* scan */
};

NBS_HD void atomicAddULL(unsigned long long* p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
#    if defined(_OPENMP)
#        pragma omp atomic
#    endif
    *p += v;
#endif
}

NBS_HD void atomicAndU32(unsigned int* p, unsigned int v)
{
#if defined(__CUDA_ARCH__)
    atomicAnd(p, v);
#else
#    if defined(_OPENMP)
#        pragma omp atomic
#    endif
    *p &= v;
#endif
}

/* ---- pass 0: bounding boxes from the coordinates (Grid::calcBoundingBoxes, grid.cpp) ---- */

NBS_HD void clusterBoundingBox(const Grid& g, int cluster)
{
    BB bb;
    for (int d = 0; d < 3; d++)
    {
        bb.lo[d] = 1e30f;
        bb.hi[d] = -1e30f;
    }
    int count = 0;
    for (int i = 0; i < c_cl; i++)
    {
        const int slot = cluster * c_cl + i;
        if (g.atomIndex[slot] >= 0)
        {
            const XQ v = g.xq[slot];
            count++;
            bb.lo[0] = fminf(bb.lo[0], v.x);
            bb.hi[0] = fmaxf(bb.hi[0], v.x);
            bb.lo[1] = fminf(bb.lo[1], v.y);
            bb.hi[1] = fmaxf(bb.hi[1], v.y);
            bb.lo[2] = fminf(bb.lo[2], v.z);
            bb.hi[2] = fmaxf(bb.hi[2], v.z);
        }
    }
    g.clBB[cluster]    = bb;
    g.clCount[cluster] = count;
}

NBS_HD void binBoundingBox(const Grid& g, int bin)
{
    BB bb;
    for (int d = 0; d < 3; d++)
    {
        bb.lo[d] = 1e30f;
        bb.hi[d] = -1e30f;
    }
    for (int cl = 0; cl < c_binCl; cl++)
    {
        if (g.clCount[bin * c_binCl + cl] > 0)
        {
            const BB cb = g.clBB[bin * c_binCl + cl];
            for (int d = 0; d < 3; d++)
            {
                bb.lo[d] = fminf(bb.lo[d], cb.lo[d]);
                bb.hi[d] = fmaxf(bb.hi[d], cb.hi[d]);
            }
        }
    }
    g.binBB[bin] = bb;
}

/* ---- entries ---- */

struct Entry
{
    int   bi = 0, shift = 0;
    float sh[3] = { 0.0f, 0.0f, 0.0f };
    bool  subDiag = false; /* central shift within one zone: half of the own bin, and only bins above it */
};

/* decodes entry slot e; false when the slot cannot hold an entry (shift not in the half shell / not the zone's x
 * shift, empty i-bin, shifted bin further than rlist from the unit cell) */
NBS_HD bool entryDecode(const Grid& g, const Params& p, int e, Entry& en)
{
    en.bi       = p.binBegin + e / c_numSlots;
    const int s = e % c_numSlots;
    const int tz = s / 9 - 1, ty = (s / 3) % 3 - 1, tx = s % 3 - 1;
    en.shift    = ((tz + 1) * 3 + (ty + 1)) * 5 + (tx + 2);
    if (p.interZone)
    {
        if (tx != p.requiredTx)
        {
            return false;
        }
    }
    else if (en.shift > c_central)
    {
        return false; /* half shell: backward shifts only */
    }
    en.sh[0]   = tx * g.box[0];
    en.sh[1]   = ty * g.box[1];
    en.sh[2]   = tz * g.box[2];
    en.subDiag = (!p.interZone && en.shift == c_central);
    const BB ibb = g.binBB[en.bi];
    if (ibb.lo[0] > ibb.hi[0])
    {
        return false;
    }
    for (int d = 0; d < 3; d++)
    {
        if (ibb.lo[d] + en.sh[d] - p.rlist > g.box[d] || ibb.hi[d] + en.sh[d] + p.rlist < 0)
        {
            return false;
        }
    }
    return true;
}

/* calls f(bj) for the j-bins of a (valid) entry whose bounding box is within rlist of the shifted i-bin, in list
 * order: own bin first on the central shift, then by column (x-major) and along z */
template<typename F>
NBS_HD void forEachJBin(const Grid& g, const Params& p, const Entry& en, F&& f)
{
    const BB ibb    = g.binBB[en.bi];
    auto     tryBin = [&](int bj) {
        if (bj < p.jBinLo || bj >= p.jBinHi)
        {
            return;
        }
        const BB jbb = g.binBB[bj];
        if (jbb.lo[0] > jbb.hi[0])
        {
            return;
        }
        if (bbDist2(ibb, en.sh, jbb) >= p.rl2)
        {
            return;
        }
        f(bj);
    };
    if (en.subDiag)
    {
        tryBin(en.bi);
    }
    const float xlo = ibb.lo[0] + en.sh[0] - p.rlist, xhi = ibb.hi[0] + en.sh[0] + p.rlist;
    const float ylo = ibb.lo[1] + en.sh[1] - p.rlist, yhi = ibb.hi[1] + en.sh[1] + p.rlist;
    int         cx0 = int(floorf(xlo / g.cellSize[0]));
    int         cx1 = int(floorf(xhi / g.cellSize[0]));
    int         cy0 = int(floorf(ylo / g.cellSize[1]));
    int         cy1 = int(floorf(yhi / g.cellSize[1]));
    cx0             = cx0 > 0 ? cx0 : 0;
    cy0             = cy0 > 0 ? cy0 : 0;
    cx1             = cx1 < g.ncx - 1 ? cx1 : g.ncx - 1;
    cy1             = cy1 < g.ncy - 1 ? cy1 : g.ncy - 1;
    for (int cx = cx0; cx <= cx1; cx++)
    {
        for (int cy = cy0; cy <= cy1; cy++)
        {
            const int c = cx * g.ncy + cy;
            for (int bj = g.colFirstBin[c]; bj < g.colFirstBin[c + 1]; bj++)
            {
                if (en.subDiag && bj <= en.bi)
                {
                    continue; /* own bin done, lower bins own the pair */
                }
                const float zhi = g.binBB[bj].hi[2], zlo = g.binBB[bj].lo[2];
                if ((ibb.lo[2] + en.sh[2]) - zhi >= p.rlist)
                {
                    continue;
                }
                if (zlo - (ibb.hi[2] + en.sh[2]) >= p.rlist)
                {
                    break; /* bins of a column are sorted along z */
                }
                tryBin(bj);
            }
        }
    }
}

/* ---- pass 1 / 2: count, then list, the bin pairs of entry slot e ---- */

template<bool WRITE>
NBS_HD void entryBinPairs(const Grid& g, const Params& p, const Work& w, int e)
{
    Entry en;
    int   n = 0;
    if (entryDecode(g, p, e, en))
    {
        const int off = WRITE ? w.entryBinPairOff[e] : 0;
        forEachJBin(g, p, en, [&](int bj) {
            if (WRITE)
            {
                w.binPairJ[off + n]     = bj;
                w.binPairEntry[off + n] = e;
            }
            n++;
        });
    }
    if (!WRITE)
    {
        w.entryNumBinPairs[e] = n;
    }
}

/* ---- pass 3: i-cluster mask of j-cluster (bin pair p, cj): item = p * 8 + cj ---- */

NBS_HD void binPairMask(const Grid& g, const Params& p, const Work& w, int item)
{
    const int pr = item >> 3, cj = item & 7;
    const int e  = w.binPairEntry[pr];
    const int bj = w.binPairJ[pr];
    Entry     en;
    entryDecode(g, p, e, en);
    const int    gcj  = bj * c_binCl + cj;
    unsigned int mask = 0;
    if (g.clCount[gcj] > 0)
    {
        const BB jbb = g.clBB[gcj];
        if (bbDist2(g.binBB[en.bi], en.sh, jbb) < p.rl2)
        {
            for (int ci = 0; ci < c_binCl; ci++)
            {
                const int gci = en.bi * c_binCl + ci;
                if (g.clCount[gci] == 0)
                {
                    continue;
                }
                if (en.subDiag && bj == en.bi && ci > cj)
                {
                    continue;
                }
                const float d2 = bbDist2(g.clBB[gci], en.sh, jbb);
                if (d2 >= p.rl2)
                {
                    continue;
                }
                /* bounding boxes closer than rbb are accepted without looking at the atoms */
                if (d2 < p.rbb2 || clusterPairInRange(g.xq, gci, gcj, en.sh, p.rl2))
                {
                    mask |= 1u << ci;
                }
            }
        }
    }
    w.binPairMask[item] = (unsigned char)mask;
}

/* ---- pass 3, warp-cooperative form: one warp per bin pair (64 cluster pairs), same result as binPairMask ----
 * phase A, lane l: the two cluster pairs (cj = l >> 2, ci = 2 (l & 3) and ci + 1) are classified from their bounding
 *   boxes: accepted (closer than rbb), to be checked atom by atom (between rbb and rlist), or out;
 * phase B, once per cluster pair to be checked: lane l tests i-atom l >> 2 against j-atoms 2 (l & 3) and + 1, a warp
 *   vote accepts the pair;
 * phase C, lanes 0..7: lane cj assembles the 8-bit i-cluster mask of j-cluster cj from the accept votes.
 * The per-lane functions below are shared by the CUDA kernel (ballot / any votes) and by the emulation (lane loops). */

struct BinPairWarp
{
    Entry en;
    int   bj = 0;
    bool  jbinInRange = false;
};

NBS_HD void maskWarpSetup(const Grid& g, const Params& p, const Work& w, int pr, BinPairWarp& bw)
{
    entryDecode(g, p, w.binPairEntry[pr], bw.en);
    bw.bj = w.binPairJ[pr];
}

/* classification of cluster pair (ci, cj): 0 = out, 1 = accepted from the bounding boxes, 2 = needs the atom check */
NBS_HD int maskWarpClassify(const Grid& g, const Params& p, const BinPairWarp& bw, int cj, int ci)
{
    const int gcj = bw.bj * c_binCl + cj;
    const int gci = bw.en.bi * c_binCl + ci;
    if (g.clCount[gcj] == 0 || g.clCount[gci] == 0)
    {
        return 0;
    }
    if (bw.en.subDiag && bw.bj == bw.en.bi && ci > cj)
    {
        return 0;
    }
    const BB jbb = g.clBB[gcj];
    if (bbDist2(g.binBB[bw.en.bi], bw.en.sh, jbb) >= p.rl2)
    {
        return 0;
    }
    const float d2 = bbDist2(g.clBB[gci], bw.en.sh, jbb);
    if (d2 >= p.rl2)
    {
        return 0;
    }
    return d2 < p.rbb2 ? 1 : 2;
}

/* lane's share of the atom check of cluster pair (ci, cj): i-atom lane >> 2 against j-atoms 2 (lane & 3), + 1 */
NBS_HD bool maskWarpAtomCheck(const Grid& g, const Params& p, const BinPairWarp& bw, int cj, int ci, int lane)
{
    const XQ xi = g.xq[(bw.en.bi * c_binCl + ci) * c_cl + (lane >> 2)];
    if (xi.x == c_farAway)
    {
        return false;
    }
    const float px = xi.x + bw.en.sh[0], py = xi.y + bw.en.sh[1], pz = xi.z + bw.en.sh[2];
    const int   j0 = (bw.bj * c_binCl + cj) * c_cl + (lane & 3) * 2;
    const XQ    xa = g.xq[j0], xb = g.xq[j0 + 1];
    return dist2(px - xa.x, py - xa.y, pz - xa.z) < p.rl2 || dist2(px - xb.x, py - xb.y, pz - xb.z) < p.rl2;
}

/* mask of j-cluster cj from the accept votes: vote0 bit l = pair (cj = l >> 2, ci = 2 (l & 3)), vote1: ci + 1 */
NBS_HD unsigned int maskWarpCompose(unsigned int vote0, unsigned int vote1, int cj)
{
    const unsigned int a = (vote0 >> (cj * 4)) & 0xfu, b = (vote1 >> (cj * 4)) & 0xfu;
    unsigned int       m = 0;
    for (int k = 0; k < 4; k++)
    {
        m |= ((a >> k) & 1u) << (2 * k);
        m |= ((b >> k) & 1u) << (2 * k + 1);
    }
    return m;
}

struct FBinPairMaskWarp
{
    Grid   g;
    Params p;
    Work   w;

#if defined(__CUDACC__)
    __device__ __forceinline__ void device(int pr, int lane) const
    {
        BinPairWarp bw;
        maskWarpSetup(g, p, w, pr, bw);
        const int    cj = lane >> 2, ci0 = (lane & 3) * 2;
        const int    c0 = maskWarpClassify(g, p, bw, cj, ci0);
        const int    c1 = maskWarpClassify(g, p, bw, cj, ci0 + 1);
        unsigned int vote0 = __ballot_sync(0xffffffffu, c0 == 1);
        unsigned int vote1 = __ballot_sync(0xffffffffu, c1 == 1);
        unsigned int need0 = __ballot_sync(0xffffffffu, c0 == 2);
        unsigned int need1 = __ballot_sync(0xffffffffu, c1 == 2);
        while (need0 != 0)
        {
            const int b = __ffs(need0) - 1;
            need0 &= need0 - 1;
            if (__any_sync(0xffffffffu, maskWarpAtomCheck(g, p, bw, b >> 2, (b & 3) * 2, lane)))
            {
                vote0 |= 1u << b;
            }
        }
        while (need1 != 0)
        {
            const int b = __ffs(need1) - 1;
            need1 &= need1 - 1;
            if (__any_sync(0xffffffffu, maskWarpAtomCheck(g, p, bw, b >> 2, (b & 3) * 2 + 1, lane)))
            {
                vote1 |= 1u << b;
            }
        }
        if (lane < c_binCl)
        {
            w.binPairMask[pr * c_binCl + lane] = (unsigned char)maskWarpCompose(vote0, vote1, lane);
        }
    }
#endif

#if defined(NBS_EMULATION)
    /* the same warp, lane by lane: compiled only into tests/kernel_emu (test infrastructure), never into the library */
    void host(int pr) const
    {
        BinPairWarp bw;
        maskWarpSetup(g, p, w, pr, bw);
        unsigned int vote0 = 0, vote1 = 0, need0 = 0, need1 = 0;
        for (int lane = 0; lane < 32; lane++)
        {
            const int cj = lane >> 2, ci0 = (lane & 3) * 2;
            const int c0 = maskWarpClassify(g, p, bw, cj, ci0);
            const int c1 = maskWarpClassify(g, p, bw, cj, ci0 + 1);
            vote0 |= (c0 == 1 ? 1u : 0u) << lane;
            vote1 |= (c1 == 1 ? 1u : 0u) << lane;
            need0 |= (c0 == 2 ? 1u : 0u) << lane;
            need1 |= (c1 == 2 ? 1u : 0u) << lane;
        }
        for (int half = 0; half < 2; half++)
        {
            unsigned int  need = half ? need1 : need0;
            unsigned int& vote = half ? vote1 : vote0;
            for (int b = 0; b < 32; b++)
            {
                if (need & (1u << b))
                {
                    bool any = false;
                    for (int lane = 0; lane < 32; lane++)
                    {
                        any = any || maskWarpAtomCheck(g, p, bw, b >> 2, (b & 3) * 2 + half, lane);
                    }
                    if (any)
                    {
                        vote |= 1u << b;
                    }
                }
            }
        }
        for (int lane = 0; lane < c_binCl; lane++)
        {
            w.binPairMask[pr * c_binCl + lane] = (unsigned char)maskWarpCompose(vote0, vote1, lane);
        }
    }
#endif
};

/* ---- pass 4: j-clusters and cjPacked groups per entry ---- */

NBS_HD void entryCountJ(const Work& w, int e)
{
    int                nj = 0;
    unsigned long long ncp = 0;
    for (int pr = w.entryBinPairOff[e]; pr < w.entryBinPairOff[e + 1]; pr++)
    {
        for (int cj = 0; cj < c_binCl; cj++)
        {
            const unsigned int m = w.binPairMask[pr * c_binCl + cj];
            if (m)
            {
                nj++;
#if defined(__CUDA_ARCH__)
                ncp += __popc(m);
#else
                ncp += __builtin_popcount(m);
#endif
            }
        }
    }
    w.entryNumJ[e]     = nj;
    w.entryGroups[e]   = (nj + 3) / 4;
    w.entryNonEmpty[e] = nj > 0 ? 1 : 0;
    if (ncp)
    {
        atomicAddULL(w.numClusterPairs, ncp);
    }
}

/* sci entries per entry slot once the total number of groups (hence maxGroups) is known */
NBS_HD void entryCountSci(const Params& p, const Work& w, int e)
{
    const int ng     = w.entryGroups[e];
    w.entryNumSci[e] = ng > 0 ? (ng + p.maxGroups - 1) / p.maxGroups : 0;
}

/* ---- pass 5: write the cjPacked groups and sci entries of entry slot e ---- */

NBS_HD void entryFill(const Grid& g, const Params& p, const Work& w, int e)
{
    const int nj = w.entryNumJ[e];
    if (nj == 0)
    {
        return;
    }
    Entry en;
    entryDecode(g, p, e, en);
    const int              cjp0 = w.entryCjOff[e];
    int                    k    = 0;
    nbnxm_b200_cj_packed_t grp;
    for (int pr = w.entryBinPairOff[e]; pr < w.entryBinPairOff[e + 1]; pr++)
    {
        const int bj = w.binPairJ[pr];
        for (int cj = 0; cj < c_binCl; cj++)
        {
            const unsigned int m = w.binPairMask[pr * c_binCl + cj];
            if (!m)
            {
                continue;
            }
            const int jm = k & 3;
            if (jm == 0)
            {
                for (int q = 0; q < 4; q++)
                {
                    grp.cj[q] = 0;
                }
                grp.imei[0].imask    = 0;
                grp.imei[0].excl_ind = 0;
                grp.imei[1].excl_ind = 0;
            }
            grp.cj[jm] = bj * c_binCl + cj;
            grp.imei[0].imask |= m << (jm * 8);
            k++;
            if (jm == 3 || k == nj)
            {
                grp.imei[1].imask     = grp.imei[0].imask; /* both halves start identical, pairlist.cpp:951-954 */
                w.cjp[cjp0 + (k - 1) / 4] = grp;
            }
        }
    }
    const int ng   = w.entryGroups[e];
    int       isci = w.entrySciOff[e];
    for (int b = 0; b < ng; b += p.maxGroups)
    {
        nbnxm_b200_sci_t s;
        s.sci             = en.bi;
        s.shift           = en.shift;
        s.cj_packed_begin = cjp0 + b;
        s.cj_packed_end   = cjp0 + (b + p.maxGroups < ng ? b + p.maxGroups : ng);
        w.sci[isci++]     = s;
    }
    w.compactEntry[w.entryCompactOff[e]] = e;
}

/* ---- pass 6: exclusions; item = compact entry * 64 + i-atom.  FILL = false marks the (cjPacked, half) mask
 * words that need their own exclusion entry, FILL = true clears the bits (after excl_ind has been assigned) ---- */

/* position of j-cluster gcj in the (sorted) j-list of an entry, -1 when absent */
NBS_HD int findJ(const Work& w, int cjp0, int nj, int gcj)
{
    int lo = 0, hi = nj - 1;
    while (lo <= hi)
    {
        const int mid = (lo + hi) >> 1;
        const int v   = w.cjp[cjp0 + (mid >> 2)].cj[mid & 3];
        if (v == gcj)
        {
            return mid;
        }
        if (v < gcj)
        {
            lo = mid + 1;
        }
        else
        {
            hi = mid - 1;
        }
    }
    return -1;
}

template<bool FILL>
NBS_HD void entryExclusions(const Grid& g, const Params& p, const Work& w, int item)
{
    const int e = w.compactEntry[item >> 6];
    const int i = item & 63;
    Entry     en;
    entryDecode(g, p, e, en);
    const int cjp0 = w.entryCjOff[e];
    const int nj   = w.entryNumJ[e];
    /* self + Newton exclusions on the diagonal cluster pair of i-cluster ci = i: only j > i interacts
     * (setSelfAndNewtonExclusionsGpu, pairlist.cpp:651-688) */
    if (en.subDiag && i < c_binCl)
    {
        const int ci = i;
        const int k  = findJ(w, cjp0, nj, en.bi * c_binCl + ci);
        if (k >= 0)
        {
            nbnxm_b200_cj_packed_t& grp = w.cjp[cjp0 + (k >> 2)];
            const unsigned int      bit = 1u << ((k & 3) * 8 + ci);
            if (grp.imei[0].imask & bit)
            {
                for (int half = 0; half < 2; half++)
                {
                    if (!FILL)
                    {
                        w.exclFlag[(cjp0 + (k >> 2)) * 2 + half] = 1;
                    }
                    else
                    {
                        nbnxm_b200_excl_t& ex = w.excl[grp.imei[half].excl_ind];
                        for (int jj = 0; jj < 4; jj++)
                        {
                            const int ja = half * 4 + jj;
                            for (int ia = 0; ia < c_cl; ia++)
                            {
                                if (ja <= ia)
                                {
                                    atomicAndU32(&ex.pair[jj * c_cl + ia], ~bit);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    /* topology exclusions of i-atom i (setExclusionsForIEntry, pairlist.cpp:1561-1660) */
    if (g.exclIndex != nullptr && g.exclAtoms != nullptr)
    {
        const int islot = en.bi * c_binAtoms + i;
        const int ia    = g.atomIndex[islot];
        if (ia >= 0)
        {
            for (int x = g.exclIndex[ia]; x < g.exclIndex[ia + 1]; x++)
            {
                const int ja = g.exclAtoms[x];
                if (ja == ia)
                {
                    continue;
                }
                const int jslot = g.slotOfAtom[ja];
                if (en.subDiag && jslot <= islot)
                {
                    continue;
                }
                const int k = findJ(w, cjp0, nj, jslot / c_cl);
                if (k < 0)
                {
                    continue;
                }
                nbnxm_b200_cj_packed_t& grp = w.cjp[cjp0 + (k >> 2)];
                const unsigned int      bit = 1u << ((k & 3) * 8 + i / c_cl);
                if (!(grp.imei[0].imask & bit))
                {
                    continue;
                }
                const int jin  = jslot & (c_cl - 1);
                const int half = jin / 4;
                if (!FILL)
                {
                    w.exclFlag[(cjp0 + (k >> 2)) * 2 + half] = 1;
                }
                else
                {
                    atomicAndU32(&w.excl[grp.imei[half].excl_ind].pair[(jin & 3) * c_cl + (i & (c_cl - 1))], ~bit);
                }
            }
        }
    }
}

/* ---- pass 7: exclusion index of (cjPacked, half) item; entry 0 stays the shared all-ones mask ---- */

NBS_HD void assignExclIndex(const Work& w, int item)
{
    if (w.exclFlag[item])
    {
        w.cjp[item >> 1].imei[item & 1].excl_ind = 1 + w.exclOff[item];
    }
}

/* ---- pass 8: perturbed (free-energy) pairs leave the cluster list for an atom-pair list (make_fep_list, pairlist.cpp:1414-1560,
 * for the GPU layout; the device form of nbnxm_b200_pairlist_split_fep, pairsearch.cpp).  Item = sci entry * 64 + i-atom: the
 * pairs of that i-atom with a perturbed partner, in the order of the entry's j-clusters.  MARK: flags the (cjPacked, half) mask
 * words that lose a bit (they need an exclusion entry of their own, like words with topology exclusions) and counts the pairs
 * that get listed; FILL: clears the bits and writes the pairs.  Whether a pair interacts comes from the topology exclusions
 * themselves, not from the mask bits, which other items of the fill pass are clearing at the same time.  Like make_fep_list
 * (rlist_fep2), interacting pairs beyond the list radius leave the cluster list without entering the perturbed one. ---- */

struct FepWork
{
    const unsigned char* slotPert;  /* per nbat slot: atom is perturbed (0 for fillers) */
    const unsigned char* clPert;    /* per cluster: any of its atoms is */
    int*                 count;     /* per item (+1): listed pairs */
    int*                 off;       /* scan */
    int*                 nonEmpty;  /* per item (+1) */
    int*                 iOff;      /* scan: index of the item's i-entry */
    int*                 pairEntry; /* the list: i-entry of every pair, j-atom slot, interacts (1) / excluded (0) */
    int*                 jjnr;
    unsigned char*       interacts;
    int*                 iinr; /* per i-entry: i-atom slot, shift index */
    int*                 shift;
};

/* per slot / cluster flags from the per-atom flags (atom order) */
struct FepFlags
{
    const unsigned char* perturbed; /* atom order */
    const int*           atomIndex;
    unsigned char*       slotPert;
    unsigned char*       clPert;
};

NBS_HD void clusterPerturbedFlags(const FepFlags& f, int cluster)
{
    unsigned char any = 0;
    for (int i = 0; i < c_cl; i++)
    {
        const int           slot = cluster * c_cl + i;
        const int           a    = f.atomIndex[slot];
        const unsigned char v    = (a >= 0 && f.perturbed[a]) ? 1 : 0;
        f.slotPert[slot]         = v;
        any |= v;
    }
    f.clPert[cluster] = any;
}

/* does the exclusion pass clear the bit of the pair (ai, aj)?  An atom's own entry in its exclusion list is skipped there
 * (entryExclusions), so an atom and its periodic image interact */
NBS_HD bool topologyExcluded(const Grid& g, int ai, int aj)
{
    if (g.exclIndex == nullptr || g.exclAtoms == nullptr || ai == aj)
    {
        return false;
    }
    for (int x = g.exclIndex[ai]; x < g.exclIndex[ai + 1]; x++)
    {
        if (g.exclAtoms[x] == aj)
        {
            return true;
        }
    }
    return false;
}

template<bool FILL>
NBS_HD void fepPairsOfIAtom(const Grid& g, const Params& p, const Work& w, const FepWork& f, int item)
{
    const int              isci  = item >> 6;
    const int              i     = item & 63;
    const nbnxm_b200_sci_t s     = w.sci[isci];
    const int              ci    = i / c_cl, ia = i & (c_cl - 1);
    const int              islot = s.sci * c_binAtoms + i;
    const int              ai    = g.atomIndex[islot];
    int                    n     = 0;
    const int              out0  = FILL ? f.off[item] : 0;
    const int              ient  = FILL ? f.iOff[item] : 0;
    if (ai >= 0)
    {
        const bool  central = (s.shift == c_central);
        const bool  iPert   = f.slotPert[islot] != 0;
        const float sh[3]   = { float(s.shift % 5 - 2) * g.box[0], float((s.shift / 5) % 3 - 1) * g.box[1],
                                float(s.shift / 15 - 1) * g.box[2] }; /* pbcutil/ishift.h */
        const XQ    xi      = g.xq[islot];
        for (int group = s.cj_packed_begin; group < s.cj_packed_end; group++)
        {
            const unsigned int imask = w.cjp[group].imei[0].imask;
            for (int jm = 0; jm < 4; jm++)
            {
                const unsigned int bit = 1u << (jm * 8 + ci);
                if (!(imask & bit))
                {
                    continue;
                }
                const int gcj = w.cjp[group].cj[jm];
                if (!iPert && !f.clPert[gcj])
                {
                    continue;
                }
                const bool diagonal = central && gcj == s.sci * c_binCl + ci;
                for (int ja = 0; ja < c_cl; ja++)
                {
                    const int jslot = gcj * c_cl + ja;
                    const int aj    = g.atomIndex[jslot];
                    if (aj < 0 || !(iPert || f.slotPert[jslot]))
                    {
                        continue;
                    }
                    if (diagonal && ja < ia)
                    {
                        continue; /* the mirrored pair owns it */
                    }
                    const int half      = ja / 4;
                    bool      interacts = false;
                    if (!(diagonal && ja == ia)) /* the self pair is excluded by construction, listed for its correction */
                    {
                        interacts = !topologyExcluded(g, ai, aj);
                        if (interacts)
                        {
                            if (!FILL)
                            {
                                w.exclFlag[group * 2 + half] = 1;
                            }
                            else
                            {
                                atomicAndU32(&w.excl[w.cjp[group].imei[half].excl_ind].pair[(ja & 3) * c_cl + ia], ~bit);
                            }
                            const XQ xj = g.xq[jslot];
                            if (!(dist2(xi.x + sh[0] - xj.x, xi.y + sh[1] - xj.y, xi.z + sh[2] - xj.z) < p.rl2))
                            {
                                continue;
                            }
                        }
                    }
                    if (FILL)
                    {
                        f.pairEntry[out0 + n] = ient;
                        f.jjnr[out0 + n]      = jslot;
                        f.interacts[out0 + n] = interacts ? 1 : 0;
                    }
                    n++;
                }
            }
        }
    }
    if (!FILL)
    {
        f.count[item]    = n;
        f.nonEmpty[item] = n > 0 ? 1 : 0;
    }
    else if (n > 0)
    {
        f.iinr[ient]  = islot;
        f.shift[ient] = s.shift;
    }
}

/* ======================================================================================================================
 * Gridding on the device: atoms (atom order, rvec) -> columns -> nbat order (Grid::putOnGrid, grid.cpp:1612;
 * sortCellsGpuGeometry :1169), the same order as the host gridder nbnxm_b200_grid_create (pairsearch.cpp): columns
 * x-major; within a column atoms sorted along z, then every 32 along +-y, then every 16 along +-x (the direction
 * alternates so that consecutive clusters stay adjacent); ties broken by atom index, so the order is unique.
 * ====================================================================================================================== */

struct GridBuild
{
    float        cellSize[2];
    int          ncx, ncy, natoms;
    const float* x;            /* atom order, 3 floats per atom */
    int*         colOfAtom;    /* natoms */
    int*         colCount;     /* atoms per column (+1) */
    int*         colAtomStart; /* scan (+1) */
    int*         colBins;      /* 64-atom bins per column (+1) */
    int*         colFirstBin;  /* scan (+1) */
    int*         colFill;      /* scatter counters */
    int*         colAtoms;     /* atoms grouped by column */
    int*         rankOfAtom;   /* block form of the scatter: place of the atom among the atoms of its block in its column */
    int*         maxColCount;  /* scalar */
    int*         atomIndex;    /* nbat slot -> atom, prefilled with -1 */
    int*         slotOfAtom;   /* atom -> nbat slot */
};

NBS_HD int atomicAddInt(int* p, int v)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const int old = *p;
    *p += v;
    return old;
#endif
}

NBS_HD void atomicMaxInt(int* p, int v)
{
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p)
    {
        *p = v;
    }
#endif
}

NBS_HD int gridColumn(const GridBuild& g, int a)
{
    int cx = int(g.x[3 * a] / g.cellSize[0]);
    int cy = int(g.x[3 * a + 1] / g.cellSize[1]);
    cx     = cx < 0 ? 0 : (cx > g.ncx - 1 ? g.ncx - 1 : cx);
    cy     = cy < 0 ? 0 : (cy > g.ncy - 1 ? g.ncy - 1 : cy);
    return cx * g.ncy + cy; /* x-major, grid.h:100 */
}

/* pass G1: column of atom a */
NBS_HD void gridColumnOfAtom(const GridBuild& g, int a)
{
    const int col  = gridColumn(g, a);
    g.colOfAtom[a] = col;
    atomicAddInt(&g.colCount[col], 1);
}

/* pass G2: bins of column c */
NBS_HD void gridColumnBins(const GridBuild& g, int c)
{
    const int n  = g.colCount[c];
    g.colBins[c] = (n + c_binAtoms - 1) / c_binAtoms;
    atomicMaxInt(g.maxColCount, n);
}

/* pass G3: scatter into columns (any order: the sort below is a total order) */
NBS_HD void gridScatterAtom(const GridBuild& g, int a)
{
    const int col = g.colOfAtom[a];
    const int pos = atomicAddInt(&g.colFill[col], 1);
    g.colAtoms[g.colAtomStart[col] + pos] = a;
}

/* Passes G1 and G3 as block passes: one block per c_gridBlockAtoms consecutive atoms with the column counters of the block in
 * its scratch memory, so that the atomics of 12 M atoms onto a few thousand global counters (2.2 + 2.1 of the 4.8 ms of a
 * gridding at 12.3 M atoms) become shared-memory atomics plus one global atomic per (block, column it touches).  Same results:
 * the column counts are sums, the order of the atoms inside a column is arbitrary at this point (the sort makes it unique). */
constexpr int c_gridBlockAtoms = 4096;

struct ColumnCountBlock
{
    GridBuild g;
    int       ncol;

    NBS_HD int numStages(int) const { return 3; }
    NBS_HD int numAtoms(int b) const
    {
        const int n = g.natoms - b * c_gridBlockAtoms;
        return n < c_gridBlockAtoms ? n : c_gridBlockAtoms;
    }
    NBS_HD int numItems(int b, int s) const { return s == 1 ? numAtoms(b) : ncol; }
    NBS_HD void operator()(int b, int s, int t, void* scratch) const
    {
        int* hist = static_cast<int*>(scratch);
        if (s == 0)
        {
            hist[t] = 0;
        }
        else if (s == 1)
        {
            const int a    = b * c_gridBlockAtoms + t;
            const int col  = gridColumn(g, a);
            g.colOfAtom[a] = col;
            atomicAddInt(&hist[col], 1);
        }
        else if (hist[t] != 0)
        {
            atomicAddInt(&g.colCount[t], hist[t]);
        }
    }
};

struct ColumnScatterBlock
{
    GridBuild g;
    int       ncol;

    NBS_HD int numStages(int) const { return 4; }
    NBS_HD int numAtoms(int b) const
    {
        const int n = g.natoms - b * c_gridBlockAtoms;
        return n < c_gridBlockAtoms ? n : c_gridBlockAtoms;
    }
    NBS_HD int numItems(int b, int s) const { return (s == 1 || s == 3) ? numAtoms(b) : ncol; }
    NBS_HD void operator()(int b, int s, int t, void* scratch) const
    {
        int* hist = static_cast<int*>(scratch);
        if (s == 0)
        {
            hist[t] = 0;
        }
        else if (s == 1)
        {
            const int a     = b * c_gridBlockAtoms + t;
            g.rankOfAtom[a] = atomicAddInt(&hist[g.colOfAtom[a]], 1);
        }
        else if (s == 2)
        {
            /* the block's atoms of column t get hist[t] consecutive places in the column: hist[t] becomes the first of them */
            const int v = hist[t];
            hist[t]     = v != 0 ? atomicAddInt(&g.colFill[t], v) : 0;
        }
        else
        {
            const int a   = b * c_gridBlockAtoms + t;
            const int col = g.colOfAtom[a];
            g.colAtoms[g.colAtomStart[col] + hist[col] + g.rankOfAtom[a]] = a;
        }
    }
};

/* pass G4: one block per column, bitonic networks in the block's scratch memory (key[nPad], idx[nPad]); the stages
 * are separated by block barriers, every item of a stage touches its own elements only */
NBS_HD int nextPow2AtLeast32(int n)
{
    int p = 32;
    while (p < n)
    {
        p <<= 1;
    }
    return p;
}

NBS_HD int log2Int(int p)
{
    int l = 0;
    while ((1 << l) < p)
    {
        l++;
    }
    return l;
}

struct ColumnSort
{
    GridBuild g;
    int       scratchPad; /* elements of scratch per block (max nPad) */

    NBS_HD int numStages(int c) const
    {
        const int n = g.colCount[c];
        if (n == 0)
        {
            return 0;
        }
        const int l = log2Int(nextPow2AtLeast32(n));
        return 1 + l * (l + 1) / 2 + 1 + 15 + 1 + 10 + 1;
    }

    /* decodes stage s: kind 0 = load key of dimension dim (segment length seg for the direction), 1 = bitonic (k, j,
     * kmax), 2 = store */
    NBS_HD void decode(int c, int s, int& kind, int& dim, int& k, int& j, int& kmax) const
    {
        const int nPad = nextPow2AtLeast32(g.colCount[c]);
        const int l    = log2Int(nPad);
        const int s1   = l * (l + 1) / 2;
        int       q;
        if (s == 0)
        {
            kind = 0;
            dim  = 2;
            kmax = nPad;
            return;
        }
        else if (s <= s1)
        {
            q    = s - 1;
            kmax = nPad;
            dim  = 2;
        }
        else if (s == s1 + 1)
        {
            kind = 0;
            dim  = 1;
            kmax = 32;
            return;
        }
        else if (s <= s1 + 1 + 15)
        {
            q    = s - (s1 + 2);
            kmax = 32;
            dim  = 1;
        }
        else if (s == s1 + 17)
        {
            kind = 0;
            dim  = 0;
            kmax = 16;
            return;
        }
        else if (s <= s1 + 17 + 10)
        {
            q    = s - (s1 + 18);
            kmax = 16;
            dim  = 0;
        }
        else
        {
            kind = 2;
            dim  = 0;
            kmax = 0;
            return;
        }
        kind = 1;
        for (k = 2;; k <<= 1)
        {
            const int nj = log2Int(k);
            if (q < nj)
            {
                j = k >> (q + 1);
                break;
            }
            q -= nj;
        }
    }

    NBS_HD int numItems(int c, int s) const
    {
        const int nPad = nextPow2AtLeast32(g.colCount[c]);
        int       kind = 0, dim = 0, k = 0, j = 0, kmax = 0;
        decode(c, s, kind, dim, k, j, kmax);
        return kind == 1 ? nPad / 2 : nPad;
    }

    NBS_HD void operator()(int c, int s, int t, void* scratch) const
    {
        const int n    = g.colCount[c];
        float*    key  = static_cast<float*>(scratch);
        int*      idx  = reinterpret_cast<int*>(key + scratchPad);
        int       kind = 0, dim = 0, k = 0, j = 0, kmax = 0;
        decode(c, s, kind, dim, k, j, kmax);
        if (kind == 0)
        {
            /* (re)load the sort key; z: ascending over the column; y / x: direction by segment parity */
            if (dim == 2)
            {
                idx[t] = t < n ? g.colAtoms[g.colAtomStart[c] + t] : 0x7fffffff;
            }
            float v = 3.0e38f;
            if (t < n)
            {
                v = g.x[3 * idx[t] + dim];
                if (dim != 2 && ((t / kmax) & 1))
                {
                    v = -v;
                }
            }
            key[t] = v;
        }
        else if (kind == 1)
        {
            const int  i   = 2 * j * (t / j) + (t % j);
            const int  p   = i + j;
            const bool asc = (k == kmax) || ((i & k) == 0);
            const float ka = key[i], kb = key[p];
            const int   ia = idx[i], ib = idx[p];
            const bool  bLessA = (kb < ka) || (kb == ka && ib < ia);
            const bool  aLessB = (ka < kb) || (ka == kb && ia < ib);
            if (asc ? bLessA : aLessB)
            {
                key[i] = kb;
                key[p] = ka;
                idx[i] = ib;
                idx[p] = ia;
            }
        }
        else if (t < n)
        {
            const int slot       = g.colFirstBin[c] * c_binAtoms + t;
            g.atomIndex[slot]    = idx[t];
            g.slotOfAtom[idx[t]] = slot;
        }
    }
};

/* pass G4, bucket form (the default): the same total order as the bitonic networks above - z ascending over the column, then
 * every 32 along +-y, then every 16 along +-x, ties by atom index - from counting instead of compare-exchange networks:
 *   z: the atoms of the column go into numBuckets(n) ~ n / 8 buckets by a monotone function of z (spread over the column's own
 *      z range), bucket sizes are scanned, atoms scattered to their bucket in any order, and every atom's rank inside its
 *      bucket (the number of atoms of the bucket that come before it) gives its final place;
 *   y, x: the rank of an atom among the 32 (16) atoms of its segment, direction by segment parity.
 * 13 stages with a block barrier each instead of ~120, no padding to a power of two, ~60 comparisons per atom of a liquid
 * instead of ~110 compare-exchanges per padded slot with the stage decoded per item.  A column whose atoms crowd into few
 * buckets (all z equal) costs bucket size comparisons per atom: slower, same order.
 * Scratch per block: keyA, idxA, keyB, idxB [scratchPad each], count[c_maxBuckets + 1], start[c_maxBuckets + 1], range[2]. */
constexpr int c_maxBuckets = 1024;
constexpr int c_scanChunk  = 32; /* bucket counts scanned in chunks of 32, then the 32 chunk totals, then added back */

NBS_HD int floatBits(float v)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(v);
#else
    int i;
    memcpy(&i, &v, sizeof(i));
    return i;
#endif
}
NBS_HD float bitsFloat(int i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float v;
    memcpy(&v, &i, sizeof(v));
    return v;
#endif
}
/* int whose signed order is the order of the floats (finite values) */
NBS_HD int   floatOrderedInt(float v) { const int i = floatBits(v); return i >= 0 ? i : (i ^ 0x7fffffff); }
NBS_HD float orderedIntFloat(int i) { return bitsFloat(i >= 0 ? i : (i ^ 0x7fffffff)); }

NBS_HD void atomicMinInt(int* p, int v)
{
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p)
    {
        *p = v;
    }
#endif
}

struct ColumnBucketSort
{
    GridBuild g;
    int       scratchPad; /* elements per key / index array of the scratch (>= the tallest column) */

    static size_t scratchBytes(int pad) { return sizeof(int) * (size_t(4) * pad + 2 * (c_maxBuckets + 1) + 2); }

    NBS_HD static int numBuckets(int n)
    {
        const int nb = (n + 7) / 8;
        return nb > c_maxBuckets ? c_maxBuckets : nb;
    }

    NBS_HD int numStages(int c) const { return g.colCount[c] == 0 ? 0 : 13; }

    NBS_HD int numItems(int c, int s) const
    {
        const int n  = g.colCount[c];
        const int nb = numBuckets(n);
        switch (s)
        {
            case 0: return n > nb + 1 ? n : nb + 1;
            case 3: return (nb + c_scanChunk - 1) / c_scanChunk;
            case 4: return 1;
            case 5: return nb + 1;
            default: return n;
        }
    }

    /* bucket of z: monotone in z (the subtraction, the multiplication by a positive factor and the conversion are) */
    NBS_HD static int bucketOf(float z, float zmin, float scale, int nb)
    {
        const int b = int((z - zmin) * scale);
        return b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    }

    /* (ka, ia) before (kb, ib) */
    NBS_HD static bool before(float ka, int ia, float kb, int ib) { return (ka < kb) || (ka == kb && ia < ib); }

    NBS_HD void operator()(int c, int s, int t, void* scratch) const
    {
        const int n     = g.colCount[c];
        const int nb    = numBuckets(n);
        float*    keyA  = static_cast<float*>(scratch);
        int*      idxA  = reinterpret_cast<int*>(keyA + scratchPad);
        float*    keyB  = reinterpret_cast<float*>(idxA + scratchPad);
        int*      idxB  = reinterpret_cast<int*>(keyB + scratchPad);
        int*      count = idxB + scratchPad;          /* [nb + 1] */
        int*      start = count + c_maxBuckets + 1;   /* [nb + 1] */
        int*      range = start + c_maxBuckets + 1;   /* ordered ints of the smallest and the largest z */
        switch (s)
        {
            case 0: /* load z keys, clear the bucket counters, initialise the z range */
                if (t < n)
                {
                    const int a = g.colAtoms[g.colAtomStart[c] + t];
                    idxA[t]     = a;
                    keyA[t]     = g.x[3 * a + 2];
                }
                if (t <= nb)
                {
                    count[t] = 0;
                }
                if (t == 0)
                {
                    range[0] = 0x7fffffff;
                    range[1] = -0x7fffffff - 1;
                }
                break;
            case 1: /* z range of the column; the plain reads only filter, the atomics decide */
            {
                const int o = floatOrderedInt(keyA[t]);
                if (o < range[0]) atomicMinInt(&range[0], o);
                if (o > range[1]) atomicMaxInt(&range[1], o);
                break;
            }
            case 2: /* bucket sizes */
            {
                const float zmin = orderedIntFloat(range[0]), zmax = orderedIntFloat(range[1]);
                const float scale = zmax > zmin ? float(nb) / (zmax - zmin) : 0.0f;
                atomicAddInt(&count[bucketOf(keyA[t], zmin, scale, nb)], 1);
                break;
            }
            case 3: /* exclusive scan of the sizes, chunk by chunk; the chunk total goes to start[chunk end] for stage 4 */
            {
                const int b0 = t * c_scanChunk, b1 = b0 + c_scanChunk < nb ? b0 + c_scanChunk : nb;
                int       sum = 0;
                for (int b = b0; b < b1; b++)
                {
                    const int v = count[b];
                    start[b]    = sum;
                    sum += v;
                }
                keyB[t] = bitsFloat(sum); /* chunk totals parked in keyB (not in use yet) */
                break;
            }
            case 4: /* scan of the chunk totals */
            {
                const int nchunks = (nb + c_scanChunk - 1) / c_scanChunk;
                int       sum     = 0;
                for (int k = 0; k < nchunks; k++)
                {
                    const int v = floatBits(keyB[k]);
                    keyB[k]     = bitsFloat(sum);
                    sum += v;
                }
                break;
            }
            case 5: /* add the chunk offsets; start[nb] = n */
                if (t < nb)
                {
                    start[t] += floatBits(keyB[t / c_scanChunk]);
                }
                else
                {
                    start[nb] = n;
                }
                break;
            case 6: /* scatter into the buckets, any order inside a bucket (count[] runs back down to 0) */
            {
                const float zmin = orderedIntFloat(range[0]), zmax = orderedIntFloat(range[1]);
                const float scale = zmax > zmin ? float(nb) / (zmax - zmin) : 0.0f;
                const int   b   = bucketOf(keyA[t], zmin, scale, nb);
                const int   pos = start[b] + atomicAddInt(&count[b], -1) - 1;
                keyB[pos]       = keyA[t];
                idxB[pos]       = idxA[t];
                break;
            }
            case 7: /* rank inside the bucket -> place in the column */
            {
                const float zmin = orderedIntFloat(range[0]), zmax = orderedIntFloat(range[1]);
                const float scale = zmax > zmin ? float(nb) / (zmax - zmin) : 0.0f;
                const float k = keyB[t];
                const int   a = idxB[t];
                const int   b = bucketOf(k, zmin, scale, nb);
                int         r = 0;
                for (int q = start[b]; q < start[b + 1]; q++)
                {
                    r += before(keyB[q], idxB[q], k, a) ? 1 : 0;
                }
                idxA[start[b] + r] = a;
                break;
            }
            case 8: /* y keys, direction by the parity of the 32-atom segment */
            {
                const float v = g.x[3 * idxA[t] + 1];
                keyA[t]       = ((t / 32) & 1) ? -v : v;
                break;
            }
            case 9: /* rank inside the segment of 32 */
            {
                const int   q0 = t & ~31, q1 = q0 + 32 < n ? q0 + 32 : n;
                const float k  = keyA[t];
                const int   a  = idxA[t];
                int         r  = 0;
                for (int q = q0; q < q1; q++)
                {
                    r += before(keyA[q], idxA[q], k, a) ? 1 : 0;
                }
                idxB[q0 + r] = a;
                break;
            }
            case 10: /* x keys, direction by the parity of the 16-atom segment */
            {
                const float v = g.x[3 * idxB[t]];
                keyB[t]       = ((t / 16) & 1) ? -v : v;
                break;
            }
            case 11: /* rank inside the segment of 16 */
            {
                const int   q0 = t & ~15, q1 = q0 + 16 < n ? q0 + 16 : n;
                const float k  = keyB[t];
                const int   a  = idxB[t];
                int         r  = 0;
                for (int q = q0; q < q1; q++)
                {
                    r += before(keyB[q], idxB[q], k, a) ? 1 : 0;
                }
                idxA[q0 + r] = a;
                break;
            }
            default: /* 12: store */
            {
                const int slot      = g.colFirstBin[c] * c_binAtoms + t;
                g.atomIndex[slot]   = idxA[t];
                g.slotOfAtom[idxA[t]] = slot;
                break;
            }
        }
    }
};

/* pass G5: atom data in nbat order (nbnxm_atomdata_t::copy x / setAtomProperties, atomdata.cpp:159-280, :1107);
 * fillers: x = -1e6, q = 0, type = ntypes - 1 */
struct AtomFill
{
    const int*   atomIndex;
    const float* x;
    const float* q;             /* atom order, may be null */
    const int*   type;          /* atom order, may be null */
    const float* ljCombPerType; /* ntypes x 2, may be null */
    const unsigned char* perturbed; /* atom order, may be null: these atoms are masked out of the cluster kernels' atom data
                                       (charge 0, type ntypes - 1: nbnxm_atomdata_mask_fep, atomdata.cpp:1039) */
    int          ntypes;
    XQ*          xq;
    int*         typeNbat;   /* may be null */
    float*       ljCombNbat; /* 2 per slot, may be null */
};

NBS_HD void fillAtomSlot(const AtomFill& f, int slot)
{
    const int a = f.atomIndex[slot];
    XQ        v;
    v.x = a >= 0 ? f.x[3 * a] : c_farAway;
    v.y = a >= 0 ? f.x[3 * a + 1] : c_farAway;
    v.z = a >= 0 ? f.x[3 * a + 2] : c_farAway;
    const bool masked = (a >= 0 && f.perturbed != nullptr && f.perturbed[a] != 0);
    v.q = (a >= 0 && f.q != nullptr && !masked) ? f.q[a] : 0.0f;
    f.xq[slot]  = v;
    const int t = (a >= 0 && f.type != nullptr && !masked) ? f.type[a] : f.ntypes - 1;
    if (f.typeNbat != nullptr)
    {
        f.typeNbat[slot] = t;
    }
    if (f.ljCombNbat != nullptr && f.ljCombPerType != nullptr)
    {
        f.ljCombNbat[2 * slot]     = f.ljCombPerType[2 * t];
        f.ljCombNbat[2 * slot + 1] = f.ljCombPerType[2 * t + 1];
    }
}

} // namespace nbs

#endif
