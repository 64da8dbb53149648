// bitslice_part.cu — instantiates the bit-sliced scan kernels for the k with k mod 4 == APC_BS_PART
// (compiled four times by the Makefile so that the objects build in parallel).
#include "bitslice_core.cuh"

#ifndef APC_BS_PART
#error "compile with -DAPC_BS_PART=0..3"
#endif

#define APC_BS_CAT2(a, b) a##b
#define APC_BS_CAT(a, b) APC_BS_CAT2(a, b)

namespace apc {

cudaError_t APC_BS_CAT(launch_bs_part, APC_BS_PART)(BsLaunchCtx &l) {
    switch (l.c->k) {
#define APC_BS_CASE(K_) case K_: return launch_bs_k<K_>(l);
#ifdef APC_BS_ONLY_K // experiment builds: one k only (context creation time against the number of kernels)
#if (APC_BS_ONLY_K % 4) == APC_BS_PART || (APC_BS_ONLY_K % 4 == 0 && APC_BS_PART == 0)
        APC_BS_CASE(APC_BS_ONLY_K)
#endif
#elif APC_BS_PART == 0
        APC_BS_CASE(4) APC_BS_CASE(8) APC_BS_CASE(12) APC_BS_CASE(16) APC_BS_CASE(20) APC_BS_CASE(24) APC_BS_CASE(28)
        APC_BS_CASE(32)
#elif APC_BS_PART == 1
        APC_BS_CASE(5) APC_BS_CASE(9) APC_BS_CASE(13) APC_BS_CASE(17) APC_BS_CASE(21) APC_BS_CASE(25) APC_BS_CASE(29)
#elif APC_BS_PART == 2
        APC_BS_CASE(2) APC_BS_CASE(6) APC_BS_CASE(10) APC_BS_CASE(14) APC_BS_CASE(18) APC_BS_CASE(22) APC_BS_CASE(26)
        APC_BS_CASE(30)
#else
        APC_BS_CASE(3) APC_BS_CASE(7) APC_BS_CASE(11) APC_BS_CASE(15) APC_BS_CASE(19) APC_BS_CASE(23) APC_BS_CASE(27)
        APC_BS_CASE(31)
#endif
#undef APC_BS_CASE
    default: return cudaErrorInvalidValue;
    }
}

#ifdef APC_BS_STATS
// deep rows computed / tests, of the kernels of this object since the last call; resets the counters
cudaError_t APC_BS_CAT(bs_stats_part, APC_BS_PART)(double *ratio) {
    unsigned long long h[2] = {0, 0}, z[2] = {0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(h, g_bs_stats, sizeof(h));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_bs_stats, z, sizeof(z));
    *ratio = h[1] ? (double)h[0] / (double)h[1] : -1.0;
    return e;
}
#endif

} // namespace apc
