// One translation unit per variant of the resident search kernel (resident_k22.cu ...): the
// variants compile in parallel and flat_search.cu only sees these three entry points.
#pragma once
#include "search_resident.cuh"

namespace bp {

template <int MAXW, int MAXEW>
struct ResidentVariant {
    static int blocks_per_sm();
    static void launch(int grid, cudaStream_t st, const SearchParams& p, const ResidentParams& x);
};

#ifdef BP_RESIDENT_MAXW
template <>
int ResidentVariant<BP_RESIDENT_MAXW, BP_RESIDENT_MAXEW>::blocks_per_sm()
{
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, fs_search_resident_kernel<BP_RESIDENT_MAXW, BP_RESIDENT_MAXEW>, 128, 0);
    return per_sm < 1 ? 1 : per_sm;
}
template <>
void ResidentVariant<BP_RESIDENT_MAXW, BP_RESIDENT_MAXEW>::launch(int grid, cudaStream_t st,
                                                                 const SearchParams& p,
                                                                 const ResidentParams& x)
{
    fs_search_resident_kernel<BP_RESIDENT_MAXW, BP_RESIDENT_MAXEW><<<grid, 128, 0, st>>>(p, x);
}
#endif

}  // namespace bp
