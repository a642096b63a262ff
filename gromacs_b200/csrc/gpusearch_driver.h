/* Pass sequence of the GPU pair-list builder, written once over a backend:
 *   nbnxm_gpusearch.cu instantiates it with the CUDA backend (kernels on the handle's stream, device buffers),
 *   tests/kernel_emu/search_emu.cpp with a host-loop backend that runs the same bodies on the CPU (test only).
 *
 * Backend concept:
 *   template<typename T> struct Buf { T* p; ... };          grow-only buffer
 *   int  reserve(Buf<T>&, size_t count)                      0 = ok; contents are not preserved
 *   int  zero(void* p, size_t bytes), int ones(void* p, size_t bytes)
 *   int  upload(T* dst, const T* hostSrc, size_t count)
 *   int  scan(const int* in, int* out, int n)                exclusive prefix sum
 *   int  readInt(const int* p, int* hostValue), readULL(...) synchronising read-back of one value
 *   int  forEach(int n, F f)                                 f(i) for i in [0, n), f copied by value
 *   int  forEachWarp(int warps, F f)                         f.device(w, lane) on the 32 lanes of a warp (CUDA) or
 *                                                            f.host(w), the same warp lane by lane (emulation)
 *   int  forEachBlock(int blocks, int threads, size_t scratchBytes, F f)
 *                                                            per block b: for s < f.numStages(b): f(b, s, t, scratch) for
 *                                                            t < f.numItems(b, s), a block barrier between stages
 */
#ifndef NBNXM_B200_GPUSEARCH_DRIVER_H
#define NBNXM_B200_GPUSEARCH_DRIVER_H

#include <cmath>
#include <cstddef>

#include "gpusearch_bodies.h"

namespace nbs
{

/* pass functors (arguments by value: they become kernel parameters) */
struct FSlotOfAtom
{
    const int* atomIndex;
    int*       slotOfAtom;
    NBS_HD void operator()(int slot) const
    {
        const int a = atomIndex[slot];
        if (a >= 0)
        {
            slotOfAtom[a] = slot;
        }
    }
};
struct FClusterBB
{
    Grid g;
    NBS_HD void operator()(int i) const { clusterBoundingBox(g, i); }
};
struct FBinBB
{
    Grid g;
    NBS_HD void operator()(int i) const { binBoundingBox(g, i); }
};
/* entry passes run slot-major (neighbouring threads: neighbouring bins, same shift) so that warps stay coherent;
 * the arrays are indexed bin-major, the order of the list */
template<bool WRITE>
struct FEntryBinPairs
{
    Grid   g;
    Params p;
    Work   w;
    int    numIBins;
    NBS_HD void operator()(int i) const { entryBinPairs<WRITE>(g, p, w, (i % numIBins) * c_numSlots + i / numIBins); }
};
struct FBinPairMask
{
    Grid   g;
    Params p;
    Work   w;
    NBS_HD void operator()(int i) const { binPairMask(g, p, w, i); }
};
struct FEntryCountJ
{
    Work w;
    NBS_HD void operator()(int i) const { entryCountJ(w, i); }
};
struct FEntryCountSci
{
    Params p;
    Work   w;
    NBS_HD void operator()(int i) const { entryCountSci(p, w, i); }
};
struct FEntryFill
{
    Grid   g;
    Params p;
    Work   w;
    NBS_HD void operator()(int i) const { entryFill(g, p, w, i); }
};
template<bool FILL>
struct FEntryExclusions
{
    Grid   g;
    Params p;
    Work   w;
    NBS_HD void operator()(int i) const { entryExclusions<FILL>(g, p, w, i); }
};
struct FAssignExclIndex
{
    Work w;
    NBS_HD void operator()(int i) const { assignExclIndex(w, i); }
};

struct FClusterPerturbed
{
    FepFlags f;
    NBS_HD void operator()(int i) const { clusterPerturbedFlags(f, i); }
};
template<bool FILL>
struct FFepPairs
{
    Grid    g;
    Params  p;
    Work    w;
    FepWork f;
    NBS_HD void operator()(int i) const { fepPairsOfIAtom<FILL>(g, p, w, f, i); }
};

struct FGridColumnOfAtom
{
    GridBuild g;
    NBS_HD void operator()(int i) const { gridColumnOfAtom(g, i); }
};
struct FGridColumnBins
{
    GridBuild g;
    NBS_HD void operator()(int i) const { gridColumnBins(g, i); }
};
struct FGridScatterAtom
{
    GridBuild g;
    NBS_HD void operator()(int i) const { gridScatterAtom(g, i); }
};
struct FAtomFill
{
    AtomFill f;
    NBS_HD void operator()(int i) const { fillAtomSlot(f, i); }
};

constexpr int c_maxColumnAtoms = 8192; /* one block sorts a column in its shared memory (136 KB of scratch at this height) */

template<typename BE>
struct SearchState
{
    template<typename T>
    using Buf = typename BE::template Buf<T>;

    Grid g{};
    bool haveGrid = false;

    Buf<int> colFirstBin, atomIndex, slotOfAtom, clCount, exclIndex, exclAtoms;
    Buf<BB>  clBB, binBB;

    /* gridding on the device */
    Buf<int>   colOfAtom, colCount, colAtomStart, colBins, colFill, colAtoms, maxColCount, rankOfAtom;
    Buf<float> qAtom, ljCombPerType; /* static per-atom / per-type properties, atom order */
    Buf<int>   typeAtom;
    int        ntypes = 0, maxColumnAtoms = 0;
    bool       haveQ = false, haveType = false, haveLjComb = false;

    Buf<int>           entryNumBinPairs, entryBinPairOff, binPairJ, binPairEntry;
    Buf<unsigned char> binPairMask;
    Buf<int>           entryNumJ, entryGroups, entryCjOff, entryNumSci, entrySciOff, entryNonEmpty, entryCompactOff;
    Buf<int>           compactEntry, exclFlag, exclOff;
    Buf<unsigned long long> numClusterPairs;

    /* the list built last */
    Buf<nbnxm_b200_sci_t>       sci;
    Buf<nbnxm_b200_cj_packed_t> cjp;
    Buf<nbnxm_b200_excl_t>      excl;
    /* pass 3 as one warp per bin pair (FBinPairMaskWarp) instead of one thread per (bin pair, j-cluster) */
    bool      cooperativeMasks = false;
    bool      bitonicColumnSort = false; /* the column sort as bitonic networks (ColumnSort) instead of buckets + ranks */
    bool      globalColumnAtomics = false; /* passes G1 / G3 one thread per atom with global atomics instead of block passes */

    /* perturbed (free-energy) atoms: when set, buildPairlist moves their pairs from the cluster list to an atom-pair list
     * (pass 8) and fillAtomData masks them out of the cluster kernels' atom data */
    bool               havePerturbed = false;
    Buf<unsigned char> perturbedAtom, slotPert, clPert, fepInteracts;
    Buf<int>           fepCount, fepOff, fepNonEmpty, fepIOff, fepPairEntry, fepJjnr, fepIinr, fepShift;
    int                numFepI = 0, numFepPairs = 0;
    int       nsci = 0, ncjp = 0, nexcl = 0, numBinPairs = 0, numEntries = 0;
    long long numClusterPairsHost = 0;
};

#define NBS_TRY(call)        \
    do                       \
    {                        \
        if ((call) != 0)     \
        {                    \
            return 1;        \
        }                    \
    } while (0)

template<typename BE>
int setGridPointers(BE& be, SearchState<BE>& st, const float* box, int ncx, int ncy, int nbins, int natoms)
{
    for (int d = 0; d < 3; d++)
    {
        st.g.box[d] = box[d];
    }
    st.g.cellSize[0] = box[0] / ncx;
    st.g.cellSize[1] = box[1] / ncy;
    st.g.ncx         = ncx;
    st.g.ncy         = ncy;
    st.g.nbins       = nbins;
    st.g.natoms      = natoms;
    NBS_TRY(be.reserve(st.clCount, size_t(nbins) * c_binCl));
    NBS_TRY(be.reserve(st.clBB, size_t(nbins) * c_binCl));
    NBS_TRY(be.reserve(st.binBB, nbins));
    st.g.colFirstBin = st.colFirstBin.p;
    st.g.atomIndex   = st.atomIndex.p;
    st.g.slotOfAtom  = st.slotOfAtom.p;
    st.g.clCount     = st.clCount.p;
    st.g.clBB        = st.clBB.p;
    st.g.binBB       = st.binBB.p;
    st.haveGrid      = true;
    return 0;
}

/* topology exclusions, CSR in atom order (host arrays; null: none) */
template<typename BE>
int setExclusions(BE& be, SearchState<BE>& st, int natoms, const int* exclIndex, const int* exclAtoms)
{
    st.g.exclIndex = nullptr;
    st.g.exclAtoms = nullptr;
    if (exclIndex != nullptr && exclAtoms != nullptr)
    {
        const int nex = exclIndex[natoms];
        NBS_TRY(be.reserve(st.exclIndex, natoms + 1));
        NBS_TRY(be.reserve(st.exclAtoms, nex > 0 ? nex : 1));
        NBS_TRY(be.upload(st.exclIndex.p, exclIndex, natoms + 1));
        if (nex > 0)
        {
            NBS_TRY(be.upload(st.exclAtoms.p, exclAtoms, nex));
        }
        st.g.exclIndex = st.exclIndex.p;
        st.g.exclAtoms = st.exclAtoms.p;
    }
    return 0;
}

/* grid description from the host gridder (putAtomsOnGrid): columns, atom order, topology exclusions */
template<typename BE>
int setGrid(BE& be, SearchState<BE>& st, const float* box, int ncx, int ncy, const int* firstBinOfColumn,
            const int* atomIndex, int nbins, int natoms, const int* exclIndex, const int* exclAtoms)
{
    const int ncol   = ncx * ncy;
    const int nslots = nbins * c_binAtoms;
    NBS_TRY(be.reserve(st.colFirstBin, ncol + 1));
    NBS_TRY(be.reserve(st.atomIndex, nslots));
    NBS_TRY(be.reserve(st.slotOfAtom, natoms));
    NBS_TRY(be.upload(st.colFirstBin.p, firstBinOfColumn, ncol + 1));
    NBS_TRY(be.upload(st.atomIndex.p, atomIndex, nslots));
    NBS_TRY(be.ones(st.slotOfAtom.p, sizeof(int) * natoms)); /* -1 */
    NBS_TRY(setGridPointers(be, st, box, ncx, ncy, nbins, natoms));
    NBS_TRY(setExclusions(be, st, natoms, exclIndex, exclAtoms));
    NBS_TRY(be.forEach(nslots, FSlotOfAtom{ st.atomIndex.p, st.slotOfAtom.p }));
    return 0;
}

/* static atom properties in atom order (host arrays, any may be null), uploaded once per topology */
template<typename BE>
int setAtomProperties(BE& be, SearchState<BE>& st, int natoms, const float* q, const int* type, int ntypes,
                      const float* ljCombPerType)
{
    st.ntypes     = ntypes;
    st.haveQ      = q != nullptr;
    st.haveType   = type != nullptr;
    st.haveLjComb = ljCombPerType != nullptr;
    if (q != nullptr)
    {
        NBS_TRY(be.reserve(st.qAtom, natoms));
        NBS_TRY(be.upload(st.qAtom.p, q, natoms));
    }
    if (type != nullptr)
    {
        NBS_TRY(be.reserve(st.typeAtom, natoms));
        NBS_TRY(be.upload(st.typeAtom.p, type, natoms));
    }
    if (ljCombPerType != nullptr)
    {
        NBS_TRY(be.reserve(st.ljCombPerType, size_t(ntypes) * 2));
        NBS_TRY(be.upload(st.ljCombPerType.p, ljCombPerType, size_t(ntypes) * 2));
    }
    return 0;
}

/* nonbonded_verlet_t::putAtomsOnGrid (nbnxm.cpp:78) on the backend: x is natoms rvecs in backend memory, atom order,
 * inside the rectangular box; ncx / ncy from nbnxm_b200_grid_dims.  Leaves the grid (columns, atom order) in the
 * search state, ready for buildPairlist; *nbinsOut bins of 64 atoms. */
template<typename BE>
int putAtomsOnGrid(BE& be, SearchState<BE>& st, const float* box, int ncx, int ncy, int natoms, const float* x, int* nbinsOut)
{
    const int ncol = ncx * ncy;
    NBS_TRY(be.reserve(st.colOfAtom, natoms));
    NBS_TRY(be.reserve(st.colAtoms, natoms));
    NBS_TRY(be.reserve(st.colCount, ncol + 1));
    NBS_TRY(be.reserve(st.colAtomStart, ncol + 1));
    NBS_TRY(be.reserve(st.colBins, ncol + 1));
    NBS_TRY(be.reserve(st.colFirstBin, ncol + 1));
    NBS_TRY(be.reserve(st.colFill, ncol + 1));
    NBS_TRY(be.reserve(st.maxColCount, 1));
    NBS_TRY(be.zero(st.colCount.p, sizeof(int) * (ncol + 1)));
    NBS_TRY(be.zero(st.colBins.p, sizeof(int) * (ncol + 1)));
    NBS_TRY(be.zero(st.colFill.p, sizeof(int) * (ncol + 1)));
    NBS_TRY(be.zero(st.maxColCount.p, sizeof(int)));
    GridBuild gb{};
    gb.cellSize[0]  = box[0] / ncx;
    gb.cellSize[1]  = box[1] / ncy;
    gb.ncx          = ncx;
    gb.ncy          = ncy;
    gb.natoms       = natoms;
    gb.x            = x;
    gb.colOfAtom    = st.colOfAtom.p;
    gb.colCount     = st.colCount.p;
    gb.colAtomStart = st.colAtomStart.p;
    gb.colBins      = st.colBins.p;
    gb.colFirstBin  = st.colFirstBin.p;
    gb.colFill      = st.colFill.p;
    gb.colAtoms     = st.colAtoms.p;
    gb.maxColCount  = st.maxColCount.p;
    /* column counters of a block of atoms in shared memory where they fit (200 KB: 51 200 columns) */
    const bool blockCounters = !st.globalColumnAtomics && size_t(ncol) * sizeof(int) <= size_t(200) * 1024 && natoms > 0;
    const int  gridBlocks    = (natoms + c_gridBlockAtoms - 1) / c_gridBlockAtoms;
    if (blockCounters)
    {
        NBS_TRY(be.reserve(st.rankOfAtom, natoms));
        gb.rankOfAtom = st.rankOfAtom.p;
        NBS_TRY(be.forEachBlock(gridBlocks, 1024, size_t(ncol) * sizeof(int), ColumnCountBlock{ gb, ncol }));
    }
    else
    {
        NBS_TRY(be.forEach(natoms, FGridColumnOfAtom{ gb }));
    }
    NBS_TRY(be.forEach(ncol, FGridColumnBins{ gb }));
    NBS_TRY(be.scan(gb.colCount, gb.colAtomStart, ncol + 1));
    NBS_TRY(be.scan(gb.colBins, gb.colFirstBin, ncol + 1));
    int nbins = 0;
    {
        unsigned long long unused = 0; /* both sizes with one synchronisation */
        NBS_TRY(be.readInts2ULL(gb.colFirstBin + ncol, &nbins, gb.maxColCount, &st.maxColumnAtoms, nullptr, &unused));
    }
    if (st.maxColumnAtoms > c_maxColumnAtoms)
    {
        return be.fail("gridding on the device: a grid column holds more atoms than one block sorts (8192); use the host gridder");
    }
    const int nslots = nbins * c_binAtoms;
    NBS_TRY(be.reserve(st.atomIndex, nslots));
    NBS_TRY(be.reserve(st.slotOfAtom, natoms));
    NBS_TRY(be.ones(st.atomIndex.p, sizeof(int) * nslots)); /* -1: filler */
    gb.atomIndex  = st.atomIndex.p;
    gb.slotOfAtom = st.slotOfAtom.p;
    if (blockCounters)
    {
        NBS_TRY(be.forEachBlock(gridBlocks, 1024, size_t(ncol) * sizeof(int), ColumnScatterBlock{ gb, ncol }));
    }
    else
    {
        NBS_TRY(be.forEach(natoms, FGridScatterAtom{ gb }));
    }
    if (st.bitonicColumnSort)
    {
        const int nPad    = nextPow2AtLeast32(st.maxColumnAtoms);
        const int threads = nPad / 2 < 1024 ? nPad / 2 : 1024;
        NBS_TRY(be.forEachBlock(ncol, threads, size_t(nPad) * (sizeof(float) + sizeof(int)), ColumnSort{ gb, nPad }));
    }
    else
    {
        const int pad     = (st.maxColumnAtoms + 31) / 32 * 32;
        const int threads = pad < 1024 ? (pad < 32 ? 32 : pad) : 1024;
        NBS_TRY(be.forEachBlock(ncol, threads, ColumnBucketSort::scratchBytes(pad), ColumnBucketSort{ gb, pad }));
    }
    NBS_TRY(setGridPointers(be, st, box, ncx, ncy, nbins, natoms));
    *nbinsOut = nbins;
    return 0;
}

/* atom data in nbat order from the grid in the state and the properties of setAtomProperties */
template<typename BE>
int fillAtomData(BE& be, SearchState<BE>& st, const float* x, XQ* xq, int* typeNbat, float* ljCombNbat)
{
    AtomFill f{};
    f.atomIndex     = st.atomIndex.p;
    f.x             = x;
    f.q             = st.haveQ ? st.qAtom.p : nullptr;
    f.type          = st.haveType ? st.typeAtom.p : nullptr;
    f.ljCombPerType = st.haveLjComb ? st.ljCombPerType.p : nullptr;
    f.perturbed     = st.havePerturbed ? st.perturbedAtom.p : nullptr;
    f.ntypes        = st.ntypes;
    f.xq            = xq;
    f.typeNbat      = typeNbat;
    f.ljCombNbat    = ljCombNbat;
    return be.forEach(st.g.nbins * c_binAtoms, FAtomFill{ f });
}

/* constructPairlist (pairlist.cpp:4056) for the GPU layout from coordinates xq in nbat order; arguments as
 * nbnxm_b200_pairlist_build (include/nbnxm_b200_search.h) */
template<typename BE>
int buildPairlist(BE& be, SearchState<BE>& st, const XQ* xq, float rlist, int minSci, int binBegin, int binEnd, int jBinLo,
                  int jBinHi, int interZone, int requiredTx)
{
    Grid& g = st.g;
    g.xq    = xq;
    Params p;
    p.rlist = rlist;
    p.rl2   = rlist * rlist;
    p.rbb2  = bbOnlyDistance2(g.cellSize, rlist);
    p.binBegin      = binBegin;
    p.binEnd        = binEnd;
    p.jBinLo        = jBinLo;
    p.jBinHi        = jBinHi;
    p.interZone     = interZone;
    p.requiredTx    = requiredTx;
    p.maxGroups     = 1 << 30;

    const int numIBins = binEnd - binBegin;
    const int nE       = numIBins * c_numSlots;
    st.nsci = st.ncjp = st.numBinPairs = st.numEntries = 0;
    st.numFepI = st.numFepPairs = 0;
    st.nexcl               = 1;
    st.numClusterPairsHost = 0;

    /* pass 0: bounding boxes of the current coordinates */
    NBS_TRY(be.forEach(g.nbins * c_binCl, FClusterBB{ g }));
    NBS_TRY(be.forEach(g.nbins, FBinBB{ g }));

    NBS_TRY(be.reserve(st.entryNumBinPairs, nE + 1));
    NBS_TRY(be.reserve(st.entryBinPairOff, nE + 1));
    NBS_TRY(be.reserve(st.entryNumJ, nE + 1));
    NBS_TRY(be.reserve(st.entryGroups, nE + 1));
    NBS_TRY(be.reserve(st.entryCjOff, nE + 1));
    NBS_TRY(be.reserve(st.entryNumSci, nE + 1));
    NBS_TRY(be.reserve(st.entrySciOff, nE + 1));
    NBS_TRY(be.reserve(st.entryNonEmpty, nE + 1));
    NBS_TRY(be.reserve(st.entryCompactOff, nE + 1));
    NBS_TRY(be.reserve(st.numClusterPairs, 1));
    NBS_TRY(be.zero(st.entryNumBinPairs.p, sizeof(int) * (nE + 1)));
    NBS_TRY(be.zero(st.entryGroups.p, sizeof(int) * (nE + 1)));
    NBS_TRY(be.zero(st.entryNumSci.p, sizeof(int) * (nE + 1)));
    NBS_TRY(be.zero(st.entryNonEmpty.p, sizeof(int) * (nE + 1)));
    NBS_TRY(be.zero(st.numClusterPairs.p, sizeof(unsigned long long)));

    Work w{};
    w.entryNumBinPairs = st.entryNumBinPairs.p;
    w.entryBinPairOff  = st.entryBinPairOff.p;
    w.entryNumJ        = st.entryNumJ.p;
    w.entryGroups      = st.entryGroups.p;
    w.entryCjOff       = st.entryCjOff.p;
    w.entryNumSci      = st.entryNumSci.p;
    w.entrySciOff      = st.entrySciOff.p;
    w.entryNonEmpty    = st.entryNonEmpty.p;
    w.entryCompactOff  = st.entryCompactOff.p;
    w.numClusterPairs  = st.numClusterPairs.p;

    /* excl entry 0: the shared all-ones mask */
    auto finishEmpty = [&]() -> int {
        NBS_TRY(be.reserve(st.excl, 1));
        NBS_TRY(be.ones(st.excl.p, sizeof(nbnxm_b200_excl_t)));
        return 0;
    };
    if (nE == 0)
    {
        return finishEmpty();
    }

    /* pass 1: bin pairs per entry slot, scan, pass 2: list them */
    NBS_TRY(be.forEach(nE, FEntryBinPairs<false>{ g, p, w, numIBins }));
    NBS_TRY(be.scan(w.entryNumBinPairs, w.entryBinPairOff, nE + 1));
    NBS_TRY(be.readInt(w.entryBinPairOff + nE, &st.numBinPairs));
    if (st.numBinPairs == 0)
    {
        return finishEmpty();
    }
    if (st.numBinPairs >= (1 << 28))
    {
        return be.fail("pair search: too many bin pairs for 32-bit item indices");
    }
    NBS_TRY(be.reserve(st.binPairJ, st.numBinPairs));
    NBS_TRY(be.reserve(st.binPairEntry, st.numBinPairs));
    NBS_TRY(be.reserve(st.binPairMask, size_t(st.numBinPairs) * c_binCl));
    w.binPairJ     = st.binPairJ.p;
    w.binPairEntry = st.binPairEntry.p;
    w.binPairMask  = st.binPairMask.p;
    NBS_TRY(be.forEach(nE, FEntryBinPairs<true>{ g, p, w, numIBins }));

    /* pass 3: cluster-pair masks */
    if (st.cooperativeMasks)
    {
        NBS_TRY(be.forEachWarp(st.numBinPairs, FBinPairMaskWarp{ g, p, w }));
    }
    else
    {
        NBS_TRY(be.forEach(st.numBinPairs * c_binCl, FBinPairMask{ g, p, w }));
    }

    /* pass 4: sizes */
    NBS_TRY(be.forEach(nE, FEntryCountJ{ w }));
    NBS_TRY(be.scan(w.entryGroups, w.entryCjOff, nE + 1));
    NBS_TRY(be.scan(w.entryNonEmpty, w.entryCompactOff, nE + 1));
    {
        /* the three sizes of this stage with one synchronisation */
        unsigned long long ncp = 0;
        NBS_TRY(be.readInts2ULL(w.entryCjOff + nE, &st.ncjp, w.entryCompactOff + nE, &st.numEntries, w.numClusterPairs, &ncp));
        st.numClusterPairsHost = (long long)ncp;
    }
    if (st.ncjp == 0)
    {
        return finishEmpty();
    }
    if (st.ncjp >= (1 << 29))
    {
        return be.fail("pair search: too many cjPacked groups for 32-bit item indices");
    }
    /* split long i-entries so that about minSci entries exist (split_sci_entry, pairlist.cpp:1769-1879) */
    if (minSci > 0)
    {
        const int mg = (st.ncjp + minSci - 1) / minSci;
        p.maxGroups  = mg > 1 ? mg : 1;
    }
    NBS_TRY(be.forEach(nE, FEntryCountSci{ p, w }));
    NBS_TRY(be.scan(w.entryNumSci, w.entrySciOff, nE + 1));
    NBS_TRY(be.readInt(w.entrySciOff + nE, &st.nsci));

    /* pass 5: the list */
    NBS_TRY(be.reserve(st.sci, st.nsci));
    NBS_TRY(be.reserve(st.cjp, st.ncjp));
    NBS_TRY(be.reserve(st.compactEntry, st.numEntries));
    NBS_TRY(be.reserve(st.exclFlag, size_t(st.ncjp) * 2 + 1));
    NBS_TRY(be.reserve(st.exclOff, size_t(st.ncjp) * 2 + 1));
    NBS_TRY(be.zero(st.exclFlag.p, sizeof(int) * (size_t(st.ncjp) * 2 + 1)));
    w.sci          = st.sci.p;
    w.cjp          = st.cjp.p;
    w.compactEntry = st.compactEntry.p;
    w.exclFlag     = st.exclFlag.p;
    w.exclOff      = st.exclOff.p;
    NBS_TRY(be.forEach(nE, FEntryFill{ g, p, w }));

    /* passes 6 / 7: exclusion masks */
    int numExclExtra = 0;
    NBS_TRY(be.forEach(st.numEntries * c_binAtoms, FEntryExclusions<false>{ g, p, w }));
    /* pass 8, mark: mask words that lose the bits of perturbed pairs get an exclusion entry of their own as well */
    FepWork   fw{};
    const int numFepItems = st.havePerturbed ? st.nsci * c_binAtoms : 0;
    if (st.havePerturbed)
    {
        if (st.nsci >= (1 << 25))
        {
            return be.fail("pair search: too many sci entries for the perturbed-pair pass");
        }
        NBS_TRY(be.reserve(st.slotPert, size_t(g.nbins) * c_binAtoms));
        NBS_TRY(be.reserve(st.clPert, size_t(g.nbins) * c_binCl));
        NBS_TRY(be.reserve(st.fepCount, numFepItems + 1));
        NBS_TRY(be.reserve(st.fepOff, numFepItems + 1));
        NBS_TRY(be.reserve(st.fepNonEmpty, numFepItems + 1));
        NBS_TRY(be.reserve(st.fepIOff, numFepItems + 1));
        NBS_TRY(be.zero(st.fepCount.p + numFepItems, sizeof(int)));
        NBS_TRY(be.zero(st.fepNonEmpty.p + numFepItems, sizeof(int)));
        NBS_TRY(be.forEach(g.nbins * c_binCl, FClusterPerturbed{ FepFlags{ st.perturbedAtom.p, g.atomIndex, st.slotPert.p, st.clPert.p } }));
        fw.slotPert = st.slotPert.p;
        fw.clPert   = st.clPert.p;
        fw.count    = st.fepCount.p;
        fw.off      = st.fepOff.p;
        fw.nonEmpty = st.fepNonEmpty.p;
        fw.iOff     = st.fepIOff.p;
        NBS_TRY(be.forEach(numFepItems, FFepPairs<false>{ g, p, w, fw }));
    }
    NBS_TRY(be.scan(w.exclFlag, w.exclOff, st.ncjp * 2 + 1));
    NBS_TRY(be.readInt(w.exclOff + st.ncjp * 2, &numExclExtra));
    st.nexcl = 1 + numExclExtra;
    NBS_TRY(be.reserve(st.excl, st.nexcl));
    NBS_TRY(be.ones(st.excl.p, sizeof(nbnxm_b200_excl_t) * st.nexcl));
    w.excl = st.excl.p;
    if (numExclExtra > 0)
    {
        NBS_TRY(be.forEach(st.ncjp * 2, FAssignExclIndex{ w }));
        NBS_TRY(be.forEach(st.numEntries * c_binAtoms, FEntryExclusions<true>{ g, p, w }));
    }
    if (st.havePerturbed)
    {
        /* pass 8, fill: sizes of the perturbed list, then its pairs; their bits leave the cluster list */
        NBS_TRY(be.scan(fw.count, fw.off, numFepItems + 1));
        NBS_TRY(be.scan(fw.nonEmpty, fw.iOff, numFepItems + 1));
        unsigned long long unused = 0;
        NBS_TRY(be.readInts2ULL(fw.off + numFepItems, &st.numFepPairs, fw.iOff + numFepItems, &st.numFepI, nullptr, &unused));
        NBS_TRY(be.reserve(st.fepPairEntry, st.numFepPairs + 1));
        NBS_TRY(be.reserve(st.fepJjnr, st.numFepPairs + 1));
        NBS_TRY(be.reserve(st.fepInteracts, st.numFepPairs + 1));
        NBS_TRY(be.reserve(st.fepIinr, st.numFepI + 1));
        NBS_TRY(be.reserve(st.fepShift, st.numFepI + 1));
        fw.pairEntry = st.fepPairEntry.p;
        fw.jjnr      = st.fepJjnr.p;
        fw.interacts = st.fepInteracts.p;
        fw.iinr      = st.fepIinr.p;
        fw.shift     = st.fepShift.p;
        if (st.numFepPairs > 0)
        {
            NBS_TRY(be.forEach(numFepItems, FFepPairs<true>{ g, p, w, fw }));
        }
    }
    return 0;
}

/* per-atom perturbed flags (atom order, host memory; null: no perturbed atoms) for the builds and griddings that follow */
template<typename BE>
int setPerturbed(BE& be, SearchState<BE>& st, int natoms, const unsigned char* perturbed)
{
    st.havePerturbed = false;
    if (perturbed == nullptr || natoms <= 0)
    {
        return 0;
    }
    NBS_TRY(be.reserve(st.perturbedAtom, natoms));
    NBS_TRY(be.upload(st.perturbedAtom.p, perturbed, natoms));
    st.havePerturbed = true;
    return 0;
}

} // namespace nbs

#endif
