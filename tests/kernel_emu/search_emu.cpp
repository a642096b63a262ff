/* TEST INFRASTRUCTURE — host-loop backend for the pass sequence of the GPU pair-list builder.
 *
 * Runs gromacs_b200/csrc/gpusearch_driver.h + gpusearch_bodies.h (the exact bodies the CUDA kernels wrap) on the CPU,
 * one loop iteration per CUDA thread, so that tests/test_gpusearch_emu.py can check the search logic against the
 * independent host builder (pairsearch.cpp) without a GPU.  Nothing in the product links or loads this file.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define NBS_EMULATION 1
#include "../../gromacs_b200/csrc/gpusearch_driver.h"

namespace
{

struct HostBackend
{
    template<typename T>
    struct Buf
    {
        T*     p     = nullptr;
        size_t n     = 0;
        size_t alloc = 0;
    };
    template<typename T>
    int reserve(Buf<T>& b, size_t count)
    {
        b.n = count;
        if (count > b.alloc)
        {
            std::free(b.p);
            b.alloc = count + count / 5 + 64;
            b.p     = static_cast<T*>(std::malloc(b.alloc * sizeof(T)));
            /* poison, so that a pass reading what no pass wrote shows up */
            std::memset(static_cast<void*>(b.p), 0x5a, b.alloc * sizeof(T));
        }
        return b.p == nullptr;
    }
    int zero(void* p, size_t bytes)
    {
        std::memset(p, 0, bytes);
        return 0;
    }
    int ones(void* p, size_t bytes)
    {
        std::memset(p, 0xff, bytes);
        return 0;
    }
    template<typename T>
    int upload(T* dst, const T* src, size_t count)
    {
        std::memcpy(dst, src, count * sizeof(T));
        return 0;
    }
    int scan(const int* in, int* out, int n)
    {
        int sum = 0;
        for (int i = 0; i < n; i++)
        {
            const int v = in[i];
            out[i]      = sum;
            sum += v;
        }
        return 0;
    }
    int readInt(const int* p, int* v)
    {
        *v = *p;
        return 0;
    }
    int readInts2ULL(const int* p0, int* v0, const int* p1, int* v1, const unsigned long long* p2, unsigned long long* v2)
    {
        *v0 = *p0;
        *v1 = *p1;
        if (p2) *v2 = *p2;
        return 0;
    }
    int readULL(const unsigned long long* p, unsigned long long* v)
    {
        *v = *p;
        return 0;
    }
    template<typename F>
    int forEach(int n, F f)
    {
        /* reversed order: no pass may depend on the order its items run in */
        for (int i = n - 1; i >= 0; i--)
        {
            f(i);
        }
        return 0;
    }
    template<typename F>
    int forEachWarp(int n, F f)
    {
        for (int i = n - 1; i >= 0; i--)
        {
            f.host(i);
        }
        return 0;
    }
    template<typename F>
    int forEachBlock(int blocks, int /*threads*/, size_t scratchBytes, F f)
    {
        void* scratch = std::malloc(scratchBytes);
        for (int b = blocks - 1; b >= 0; b--)
        {
            std::memset(scratch, 0x5a, scratchBytes);
            const int ns = f.numStages(b);
            for (int s = 0; s < ns; s++)
            {
                for (int t = f.numItems(b, s) - 1; t >= 0; t--)
                {
                    f(b, s, t, scratch);
                }
            }
        }
        std::free(scratch);
        return 0;
    }
    int fail(const char* msg)
    {
        std::fprintf(stderr, "search_emu: %s\n", msg);
        return 1;
    }
};

HostBackend                   g_be;
nbs::SearchState<HostBackend> g_st;

} // namespace

extern "C" {

int search_emu_set_grid(const float* box, int ncx, int ncy, const int* first_bin_of_column, const int* atom_index, int nbins,
                        int natoms, const int* excl_index, const int* excl_atoms)
{
    return nbs::setGrid(g_be, g_st, box, ncx, ncy, first_bin_of_column, atom_index, nbins, natoms, excl_index, excl_atoms);
}

int search_emu_set_bitonic_column_sort(int on)
{
    g_st.bitonicColumnSort = on != 0;
    /* the older forms travel together: bitonic networks with the per-atom global atomics of passes G1 / G3 */
    g_st.globalColumnAtomics = on != 0;
    return 0;
}

int search_emu_set_cooperative_masks(int on)
{
    g_st.cooperativeMasks = on != 0;
    return 0;
}

int search_emu_put_atoms_on_grid(const float* box, int ncx, int ncy, int natoms, const float* x, const int* excl_index,
                                 const int* excl_atoms, int* nbins)
{
    if (nbs::putAtomsOnGrid(g_be, g_st, box, ncx, ncy, natoms, x, nbins)) return 1;
    return nbs::setExclusions(g_be, g_st, natoms, excl_index, excl_atoms);
}

int search_emu_get_order(int* atom_index, int* first_bin_of_column)
{
    std::memcpy(atom_index, g_st.atomIndex.p, sizeof(int) * g_st.g.nbins * nbs::c_binAtoms);
    std::memcpy(first_bin_of_column, g_st.colFirstBin.p, sizeof(int) * (g_st.g.ncx * g_st.g.ncy + 1));
    return 0;
}

int search_emu_fill_atomdata(int natoms, const float* x, const float* q, const int* type, int ntypes, const float* lj_comb_per_type,
                             float* xq, int* type_nbat, float* lj_comb)
{
    if (nbs::setAtomProperties(g_be, g_st, natoms, q, type, ntypes, lj_comb_per_type)) return 1;
    return nbs::fillAtomData(g_be, g_st, x, reinterpret_cast<nbs::XQ*>(xq), type_nbat, lj_comb);
}

int search_emu_build(const float* xq, float rlist, int min_sci, int bin_begin, int bin_end, int j_bin_lo, int j_bin_hi,
                     int inter_zone, int required_tx, int* sizes, long long* ncluster_pairs)
{
    if (nbs::buildPairlist(g_be, g_st, reinterpret_cast<const nbs::XQ*>(xq), rlist, min_sci, bin_begin, bin_end, j_bin_lo,
                           j_bin_hi, inter_zone, required_tx))
    {
        return 1;
    }
    sizes[0]        = g_st.nsci;
    sizes[1]        = g_st.ncjp;
    sizes[2]        = g_st.nexcl;
    sizes[3]        = g_st.numBinPairs;
    *ncluster_pairs = g_st.numClusterPairsHost;
    return 0;
}

/* perturbed atoms (atom order; null: none) for the builds that follow: pass 8 of buildPairlist */
int search_emu_set_perturbed(int natoms, const unsigned char* perturbed)
{
    return nbs::setPerturbed(g_be, g_st, natoms, perturbed);
}

int search_emu_fep_sizes(int* num_i, int* num_pairs)
{
    *num_i     = g_st.numFepI;
    *num_pairs = g_st.numFepPairs;
    return 0;
}

int search_emu_fep_copy(int* iinr, int* shift, int* pair_entry, int* jjnr, unsigned char* interacts)
{
    if (g_st.numFepI)
    {
        std::memcpy(iinr, g_st.fepIinr.p, sizeof(int) * g_st.numFepI);
        std::memcpy(shift, g_st.fepShift.p, sizeof(int) * g_st.numFepI);
    }
    if (g_st.numFepPairs)
    {
        std::memcpy(pair_entry, g_st.fepPairEntry.p, sizeof(int) * g_st.numFepPairs);
        std::memcpy(jjnr, g_st.fepJjnr.p, sizeof(int) * g_st.numFepPairs);
        std::memcpy(interacts, g_st.fepInteracts.p, g_st.numFepPairs);
    }
    return 0;
}

int search_emu_copy(nbnxm_b200_sci_t* sci, nbnxm_b200_cj_packed_t* cjp, nbnxm_b200_excl_t* excl)
{
    if (g_st.nsci) std::memcpy(sci, g_st.sci.p, sizeof(*sci) * g_st.nsci);
    if (g_st.ncjp) std::memcpy(cjp, g_st.cjp.p, sizeof(*cjp) * g_st.ncjp);
    std::memcpy(excl, g_st.excl.p, sizeof(*excl) * g_st.nexcl);
    return 0;
}

} // extern "C"
