/* TEST INFRASTRUCTURE — runs the per-pair body of the perturbed nonbonded kernel (gromacs_b200/csrc/nbfe_bodies.h, the
 * exact code the CUDA kernel wraps) in a host loop, one iteration per CUDA thread, so that tests/test_fep_emu.py can
 * check it against the pinned oracle (oracle/nbfe_oracle.py) without a GPU.  Nothing in the product links this file. */
#include <vector>

#include "../../gromacs_b200/csrc/nbfe_bodies.h"

extern "C" int fep_emu_run(const nbfe::Params* p, const float* xq, const float* qAB, const int* typeAB, const float* ljCombAB,
                           const float* nbfp, const float* shiftVec, int numI, const int* iinr, const int* jindex, const int* jjnr,
                           const int* shift, const unsigned char* exclFep, float* f4, double* fshift, double* energy, double* dvdl)
{
    const int        numPairs = jindex[numI];
    std::vector<int> pairEntry(numPairs);
    for (int n = 0; n < numI; n++)
    {
        for (int j = jindex[n]; j < jindex[n + 1]; j++)
        {
            pairEntry[j] = n;
        }
    }
    nbfe::Atoms a{ xq, qAB, typeAB, ljCombAB, nbfp, shiftVec, f4, fshift, energy, dvdl };
    nbfe::List  l{ numPairs, pairEntry.data(), iinr, shift, jjnr, exclFep };
    /* reversed order: the result may not depend on the order the pairs run in (beyond float summation order) */
    for (int j = numPairs - 1; j >= 0; j--)
    {
        nbfe::pair(*p, a, l, j);
    }
    return 0;
}
