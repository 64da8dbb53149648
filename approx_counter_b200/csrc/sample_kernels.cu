// sample_kernels.cu — K2: turn the sampled read ends (ASCII) into scan tiles.
//
// Replaces the reference's StringSet<Dna5String> sample (:38, filled by
// sampleSequences :415-476) and the SeqAn index built over it (:537-541).
// Layout: see apc_internal.h.  One thread produces one uint4 (16 bases of one
// read); thread order is lane-fastest so the stores of a warp are one
// contiguous 512-byte line group.
#include "apc_internal.h"

namespace apc {

__device__ __forceinline__ uint32_t ascii_to_code(uint8_t ch) {
    const uint32_t u = ch & 0xDFu; // fold case
    uint32_t code = kCodeN;        // everything that is not ACGT is N (Dna5)
    code = (u == 'A') ? 0x00u : code;
    code = (u == 'C') ? 0x10u : code;
    code = (u == 'G') ? 0x20u : code;
    code = (u == 'T') ? 0x30u : code;
    return code;
}

__device__ __forceinline__ uint4 encode_chunk(const uint8_t *__restrict__ src, uint32_t avail) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t pos = i * 4 + j;
            const uint32_t code = pos < avail ? ascii_to_code(src[pos]) : kCodeN;
            v |= code << (8 * j);
        }
        w[i] = v;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void build_tiles_uniform_kernel(const uint8_t *__restrict__ ascii, uint64_t n_reads,
                                           uint32_t read_len, uint32_t chunks, uint64_t n_items,
                                           uint4 *__restrict__ tiles, uint32_t *__restrict__ lens) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items) return;
    const uint32_t lane = idx & 31;
    const uint64_t tc = idx >> 5; // tile * chunks + chunk
    const uint32_t chunk = tc % chunks;
    const uint64_t tile = tc / chunks;
    const uint64_t read = tile * kTileReads + lane;
    uint4 out = make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u);
    if (read < n_reads) {
        const uint32_t from = chunk * kChunkBases;
        const uint32_t avail = read_len > from ? read_len - from : 0;
        out = encode_chunk(ascii + read * read_len + from, avail);
        if (chunk == 0) lens[read] = read_len;
    }
    tiles[idx] = out;
}

__global__ void build_tiles_ragged_kernel(const uint8_t *__restrict__ ascii,
                                          const uint64_t *__restrict__ offs, uint64_t n_reads,
                                          uint32_t chunks, uint64_t n_items,
                                          uint4 *__restrict__ tiles, uint32_t *__restrict__ lens) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items) return;
    const uint32_t lane = idx & 31;
    const uint64_t tc = idx >> 5;
    const uint32_t chunk = tc % chunks;
    const uint64_t tile = tc / chunks;
    const uint64_t read = tile * kTileReads + lane;
    uint4 out = make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u);
    if (read < n_reads) {
        const uint64_t b = offs[read];
        const uint32_t len = (uint32_t)(offs[read + 1] - b);
        const uint32_t from = chunk * kChunkBases;
        const uint32_t avail = len > from ? len - from : 0;
        out = encode_chunk(ascii + b + from, avail);
        if (chunk == 0) lens[read] = len;
    }
    tiles[idx] = out;
}

cudaError_t launch_build_tiles_uniform(const uint8_t *d_ascii, uint64_t n_reads, uint32_t read_len,
                                       uint32_t chunks, uint32_t n_tiles, uint4 *d_tiles,
                                       uint32_t *d_lens, cudaStream_t s) {
    const uint64_t n_items = (uint64_t)n_tiles * chunks * kTileReads;
    if (n_items == 0) return cudaSuccess;
    const int threads = 256;
    const uint64_t blocks = (n_items + threads - 1) / threads;
    build_tiles_uniform_kernel<<<(unsigned)blocks, threads, 0, s>>>(d_ascii, n_reads, read_len, chunks,
                                                                    n_items, d_tiles, d_lens);
    return cudaGetLastError();
}

cudaError_t launch_build_tiles_ragged(const uint8_t *d_ascii, const uint64_t *d_offs, uint64_t n_reads,
                                      uint32_t chunks, uint32_t n_tiles, uint4 *d_tiles,
                                      uint32_t *d_lens, cudaStream_t s) {
    const uint64_t n_items = (uint64_t)n_tiles * chunks * kTileReads;
    if (n_items == 0) return cudaSuccess;
    const int threads = 256;
    const uint64_t blocks = (n_items + threads - 1) / threads;
    build_tiles_ragged_kernel<<<(unsigned)blocks, threads, 0, s>>>(d_ascii, d_offs, n_reads, chunks,
                                                                   n_items, d_tiles, d_lens);
    return cudaGetLastError();
}

} // namespace apc
