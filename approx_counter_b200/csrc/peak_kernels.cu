// peak_kernels.cu — integer-pipe peak microbenchmark (SURVEY.md §8d, step 4 of
// the build plan): the driver's MEASURED_PEAKS.json only holds HBM and bf16
// tensor peaks, while K1 is bound by the integer pipes.  Three kernels issue
// independent chains of (a) LOP3 only — ALU pipe, (b) IMAD only — FMA pipe,
// (c) a 1 : 1 LOP3 : IMAD blend; the
// returned figures are lane-operations per second over the whole GPU.
#include <cstring>

#include "apc_internal.h"
#include "scan_core.cuh"

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
    const uint32_t mulr = mul + (threadIdx.x >> 31); // == mul, but opaque: lives in a vector register
    (void)mulr;
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
            } else if (MODE == 5) { // 4 LOP3 with three register operands each
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
                x[i] = lop3<0xEA>(x[i], x[l], x[j]);
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
                x[i] = lop3<0xFE>(x[i], x[l], x[j]);
            } else if (MODE == 6) { // K1's blend: 13 LOP3 (3 regs) : 6 IMAD (reg*mul+reg)
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
                x[i] = x[i] * mul + x[l];
                x[i] = lop3<0xEA>(x[i], x[l], x[j]);
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
            } else if (MODE == 8) { // same blend, multiplier in a vector register (3 register reads per IMAD)
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
                x[i] = x[i] * mulr + x[l];
                x[i] = lop3<0xEA>(x[i], x[l], x[j]);
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
            } else if (MODE == 7) { // same blend, IMAD without a register addend
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
                x[i] = x[i] * mul;
                x[i] = lop3<0xEA>(x[i], x[l], x[j]);
                x[i] = lop3<0x96>(x[i], x[j], x[l]);
            } else if (MODE == 3) { // 4 IMAD.HI
                x[i] = __umulhi(x[i], mul) + x[j];
                x[i] = __umulhi(x[i], a) + x[l];
                x[i] = __umulhi(x[i], mul) + x[j];
                x[i] = __umulhi(x[i], b) + x[l];
            } else if (MODE == 4) { // 4 IMAD.WIDE (64-bit result, both halves consumed)
                unsigned long long t = (unsigned long long)x[i] * mul + (((unsigned long long)x[l] << 32) | x[j]);
                x[i] = (uint32_t)t; x[j] = (uint32_t)(t >> 32);
                t = (unsigned long long)x[i] * mul + (((unsigned long long)x[l] << 32) | x[j]);
                x[i] = (uint32_t)t; x[j] = (uint32_t)(t >> 32);
                t = (unsigned long long)x[i] * mul + (((unsigned long long)x[l] << 32) | x[j]);
                x[i] = (uint32_t)t; x[j] = (uint32_t)(t >> 32);
                t = (unsigned long long)x[i] * mul + (((unsigned long long)x[l] << 32) | x[j]);
                x[i] = (uint32_t)t; x[j] = (uint32_t)(t >> 32);
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

// The scan kernel's column update fed from registers (no text loads, no table
// look-ups): what the two integer pipes can sustain on exactly K1's instruction
// mix (6.5 LOP3 : 3 IMAD per unit and column), for SETS independent 4-word state
// sets per thread and MB resident CTAs of THREADS threads per SM.
template <int NW, int SETS, int THREADS, int MB>
__global__ void __launch_bounds__(THREADS, MB) scan_core_kernel(uint32_t seed, uint32_t mul, uint32_t iters,
                                                                uint32_t *out) {
    ScanState st[SETS];
    uint4 eqa[SETS], eqb[SETS];
    const uint32_t t = threadIdx.x * 2654435761u + blockIdx.x * 40503u + seed;
    const uint32_t m = mul - 1;
#pragma unroll
    for (int s = 0; s < SETS; s++) {
        Column<NW>::init(st[s], mul, m);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            st[s].r0[w] ^= t + w + s * 11u; st[s].r1[w] ^= t * 3u + w + s; st[s].s2[w] ^= t * 7u + w + s;
        }
        eqa[s] = make_uint4(t ^ (0x1111u + s), t ^ 0x2222u, t ^ 0x4444u, t ^ 0x8888u);
        eqb[s] = make_uint4(~t + s, t * 5u, t * 9u, t * 17u);
    }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int s = 0; s < SETS; s++) step2<NW>(st[s], eqa[s], eqb[s], mul, m);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int s = 0; s < SETS; s++)
#pragma unroll
        for (int w = 0; w < 4; w++)
            r ^= st[s].a0[w] ^ st[s].a1[w] ^ st[s].a2[w] ^ st[s].r0[w] ^ st[s].r1[w] ^ st[s].s2[w];
    if (r == 0x12345678u) out[0] = r;
}

template <int NW, int SETS, int THREADS, int MB>
static cudaError_t time_scan_core(const Ctx &c, uint32_t *d_out, double *unit_cols_per_s) {
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&e1);
    if (e != cudaSuccess) return e;
    const int grid = c.sm_count * MB;
    const uint32_t iters = 2048;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, c.stream);
        scan_core_kernel<NW, SETS, THREADS, MB><<<grid, THREADS, 0, c.stream>>>(17u + rep, 4u, iters, d_out);
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
    // per thread and iteration: 8 x step2 = 16 columns x SETS x (4 / NW) units
    *unit_cols_per_s = (double)grid * THREADS * iters * 16.0 * SETS * (4.0 / NW) / (best * 1e-3);
    return cudaGetLastError();
}

#ifdef APC_BS_STATS
cudaError_t bs_stats_part0(double *);
cudaError_t bs_stats_part1(double *);
cudaError_t bs_stats_part2(double *);
cudaError_t bs_stats_part3(double *);
#endif

cudaError_t microbench(const Ctx &c, const char *name, double *value) {
#ifdef APC_BS_STATS
    if (!std::strcmp(name, "bs_stats")) {
        cudaError_t e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) return e;
        switch (c.k & 3) {
        case 0: return bs_stats_part0(value);
        case 1: return bs_stats_part1(value);
        case 2: return bs_stats_part2(value);
        default: return bs_stats_part3(value);
        }
    }
#endif
    uint32_t *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_out, 16);
    if (e != cudaSuccess) return e;
    if (!std::strcmp(name, "lop3")) e = time_mode<0>(c, d_out, value);
    else if (!std::strcmp(name, "imad")) e = time_mode<1>(c, d_out, value);
    else if (!std::strcmp(name, "mixed")) e = time_mode<2>(c, d_out, value);
    else if (!std::strcmp(name, "imad_hi")) e = time_mode<3>(c, d_out, value);
    else if (!std::strcmp(name, "lop3_3reg")) e = time_mode<5>(c, d_out, value);
    else if (!std::strcmp(name, "blend_3reg")) e = time_mode<6>(c, d_out, value);
    else if (!std::strcmp(name, "blend_2reg")) e = time_mode<7>(c, d_out, value);
    else if (!std::strcmp(name, "blend_rrr")) e = time_mode<8>(c, d_out, value);
    else if (!std::strcmp(name, "imad_wide")) e = time_mode<4>(c, d_out, value);
    // core_<sets>_<threads>_<ctas per SM>: warps per SMSP = threads/128 * ctas
    else if (!std::strcmp(name, "core_1_256_2")) e = time_scan_core<1, 1, 256, 2>(c, d_out, value);
    else if (!std::strcmp(name, "core_1_256_3")) e = time_scan_core<1, 1, 256, 3>(c, d_out, value);
    else if (!std::strcmp(name, "core_1_256_4")) e = time_scan_core<1, 1, 256, 4>(c, d_out, value);
    else if (!std::strcmp(name, "core_1_128_9")) e = time_scan_core<1, 1, 128, 9>(c, d_out, value);
    else if (!std::strcmp(name, "core_2_256_1")) e = time_scan_core<1, 2, 256, 1>(c, d_out, value);
    else if (!std::strcmp(name, "core_2_256_2")) e = time_scan_core<1, 2, 256, 2>(c, d_out, value);
    else if (!std::strcmp(name, "core_2_128_5")) e = time_scan_core<1, 2, 128, 5>(c, d_out, value);
    else if (!std::strcmp(name, "core_3_128_3")) e = time_scan_core<1, 3, 128, 3>(c, d_out, value);
    else if (!std::strcmp(name, "core64_1_256_3")) e = time_scan_core<2, 1, 256, 3>(c, d_out, value);
    else e = cudaErrorInvalidValue;
    cudaFree(d_out);
    return e;
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
