// peak_kernels.cu — integer-pipe peak microbenchmark (SURVEY.md §8d, step 4 of
// the build plan): the driver's MEASURED_PEAKS.json only holds HBM and bf16
// tensor peaks, while K1 is bound by the integer pipes.  Three kernels issue
// independent chains of (a) LOP3 only — ALU pipe, (b) IMAD only — FMA pipe,
// (c) the 6.5 : 5 LOP3 : IMAD blend of the scan kernel's column step; the
// returned figures are lane-operations per second over the whole GPU.
#include "apc_internal.h"

namespace apc {

constexpr int kPeakChains = 8;
constexpr int kPeakIters = 4096;
constexpr int kPeakThreads = 256;

template <int MODE>
__global__ void __launch_bounds__(kPeakThreads) int_peak_kernel(uint32_t a, uint32_t b, uint32_t mul,
                                                                uint32_t *out) {
    uint32_t x[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; i++) x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int i = 0; i < kPeakChains; i++) {
            // every op takes a neighbouring chain as an operand so that neither
            // nvcc nor ptxas can collapse consecutive ops into one
            const int j = (i + 1) & (kPeakChains - 1), l = (i + 3) & (kPeakChains - 1);
            if (MODE == 0) { // 4 LOP3
                x[i] = (x[i] & x[j]) ^ a;
                x[i] = (x[i] | x[l]) ^ b;
                x[i] = (x[i] ^ x[j]) | a;
                x[i] = (x[i] & b) ^ x[l];
            } else if (MODE == 1) { // 4 IMAD
                x[i] = x[i] * mul + x[j];
                x[i] = x[i] * mul + x[l];
                x[i] = x[i] * mul + x[j];
                x[i] = x[i] * mul + x[l];
            } else { // 2 LOP3 + 2 IMAD interleaved
                x[i] = (x[i] & x[j]) ^ a;
                x[i] = x[i] * mul + x[l];
                x[i] = (x[i] | x[l]) ^ b;
                x[i] = x[i] * mul + x[j];
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kPeakChains; i++) r ^= x[i];
    if (r == 0x12345678u) out[0] = r; // keep the chains alive
}

template <int MODE>
static cudaError_t time_mode(const Ctx &c, uint32_t *d_out, double *ops_per_s) {
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&e1);
    if (e != cudaSuccess) return e;
    const int grid = c.sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, c.stream);
        int_peak_kernel<MODE><<<grid, kPeakThreads, 0, c.stream>>>(0x5a5a1234u + rep, 0x0ff0c3a5u, 5u + 2 * rep, d_out);
        cudaEventRecord(e1, c.stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    const double ops = (double)grid * kPeakThreads * (double)kPeakIters * kPeakChains * 4.0;
    *ops_per_s = ops / (best * 1e-3);
    return cudaGetLastError();
}

cudaError_t measure_int_peak(const Ctx &c, double *lop3, double *imad, double *mixed) {
    uint32_t *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_out, 16);
    if (e != cudaSuccess) return e;
    double v[3] = {0, 0, 0};
    e = time_mode<0>(c, d_out, &v[0]);
    if (e == cudaSuccess) e = time_mode<1>(c, d_out, &v[1]);
    if (e == cudaSuccess) e = time_mode<2>(c, d_out, &v[2]);
    cudaFree(d_out);
    if (lop3) *lop3 = v[0];
    if (imad) *imad = v[1];
    if (mixed) *mixed = v[2];
    return e;
}

} // namespace apc
