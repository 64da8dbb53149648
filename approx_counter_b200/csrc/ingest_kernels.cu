// ingest_kernels.cu — K6: FASTA / FASTQ ingest and read sampling on the device.
//
// Replaces, for FASTA and 4-line FASTQ (what basecallers write), the reference's readRecords (:819-825) and the walk
// and copies of sampleSequences (:447-471): the file's bytes are copied to HBM once, the records are indexed there, and every sample (start
// and end of a run, every run of -mr) is gathered from the resident bytes straight into the staging buffer the
// layout kernels read — no host parse, no host-side copies of the read ends, no second upload.
//
//   count_newlines_kernel   one pass over the bytes: newlines per 4 KB tile (one warp)   (HBM-bound: reads 1 B / B)
//   [prefix sum over the tiles: cub::DeviceScan, tiles = bytes / 4096 elements]
//   write_newlines_kernel   second pass: the byte offset of every newline, ascending    (reads 1 B / B, L2-warm for
//                                                                                         files below ~100 MB)
//   index_records_kernel    one warp per 32 records: grammar check (header / '+' markers, quality length, no blanks
//                           inside the sequence) and the record's (sequence offset, length)
//   pick_*_kernel + prefix  the first nb_sample ids of the caller's shuffled order with length >= 2*cut (:447-461)
//   gather_ends_kernel      prefix(cut) (:466) or the last cut+1 bases (:463) of the chosen reads -> ASCII rows
//
// FASTA that is not one line per sequence (wrapped sequences, blank lines, headers without a sequence line) is first
// re-laid as single-line records in a second buffer (measure_lines_kernel + prefix sum + unwrap_lines_kernel) and then
// indexed like the rest.  Wrapped FASTQ and blanks inside a sequence line are reported as APC_ERR_FORMAT by
// apc_ingest_fastx and are the host parser's job (csrc/host/host_util.cpp).
#include <cub/device/device_scan.cuh>

#include "apc_internal.h"

namespace apc {

constexpr int kIngestThreads = 256;                                     // 8 warps, one tile each
constexpr int kIngestIters = (int)(kIngestTileBytes / (32 * 16));       // uint4 loads per lane and tile
static_assert(kIngestIters * 32 * 16 == (int)kIngestTileBytes, "tile = 32 lanes x iterations x 16 bytes");

// bit 8 j + 7 of word i of the result = byte 4 i + j of the 16 is '\n'
__device__ __forceinline__ uint4 newline_bits16(const uint4 v) {
    return make_uint4(__vcmpeq4(v.x, 0x0A0A0A0Au) & 0x80808080u, __vcmpeq4(v.y, 0x0A0A0A0Au) & 0x80808080u,
                      __vcmpeq4(v.z, 0x0A0A0A0Au) & 0x80808080u, __vcmpeq4(v.w, 0x0A0A0A0Au) & 0x80808080u);
}
__device__ __forceinline__ uint32_t popc4(const uint4 b) { return __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w); }
// the same as one 16-bit mask: bit b = byte b of the 16 is '\n'
__device__ __forceinline__ uint32_t newline_mask16(const uint4 v) {
    const uint4 b = newline_bits16(v);
    const uint32_t w[4] = {b.x, b.y, b.z, b.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) m |= ((((w[i] >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * i); // bits 0, 8, 16, 24 -> 0..3
    return m;
}

// One warp per tile of kIngestTileBytes: no shared memory, no barriers — a warp's loads of one step are one contiguous
// 512-byte request, and the order of the newlines inside a tile is (step, lane, byte).
__global__ void __launch_bounds__(kIngestThreads)
count_newlines_kernel(const uint4 *__restrict__ file, const uint64_t n_tiles, uint64_t *__restrict__ tile_nl) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (kIngestThreads / 32);
    for (uint64_t tile = (uint64_t)blockIdx.x * (kIngestThreads / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps) {
        const uint4 *p = file + tile * (kIngestTileBytes / 16) + lane;
        uint4 v[kIngestIters];
#pragma unroll
        for (int it = 0; it < kIngestIters; it++) v[it] = __ldg(p + it * 32);
        uint32_t c = 0;
#pragma unroll
        for (int it = 0; it < kIngestIters; it++) c += popc4(newline_bits16(v[it]));
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) tile_nl[tile] = c;
    }
}

__global__ void __launch_bounds__(kIngestThreads)
write_newlines_kernel(const uint4 *__restrict__ file, const uint64_t n_tiles, const uint64_t *__restrict__ tile_base,
                      uint64_t *__restrict__ nl) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (kIngestThreads / 32);
    for (uint64_t tile = (uint64_t)blockIdx.x * (kIngestThreads / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += warps) {
        const uint4 *p = file + tile * (kIngestTileBytes / 16) + lane;
        uint4 v[kIngestIters];
#pragma unroll
        for (int it = 0; it < kIngestIters; it++) v[it] = __ldg(p + it * 32);
        uint64_t base = tile_base[tile];
#pragma unroll
        for (int it = 0; it < kIngestIters; it++) {
            uint32_t m = newline_mask16(v[it]);
            const uint32_t has = __ballot_sync(0xFFFFFFFFu, m != 0);
            if (!has) continue; // 512-byte steps without a newline cost nothing more
            const uint32_t c = __popc(m);
            uint32_t incl = c, total; // inclusive prefix over the warp
            if (__popc(has) <= 4) { // a few lanes hold newlines: their counts are handed round one by one
                total = 0;
                for (uint32_t rest = has; rest; rest &= rest - 1) {
                    const uint32_t j = (uint32_t)__ffs((int)rest) - 1;
                    const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
                    incl += j < lane ? cj : 0;
                    total += cj;
                }
            } else {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += o;
                }
                total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            uint64_t at = base + incl - c;
            const uint64_t byte0 = tile * kIngestTileBytes + ((uint64_t)it * 32 + lane) * 16;
            while (m) {
                nl[at++] = byte0 + (uint32_t)(__ffs((int)m) - 1);
                m &= m - 1;
            }
            base += total;
        }
    }
}

struct LineTable {
    const uint8_t *file;
    const uint64_t *nl;
    uint64_t n_nl, n_bytes; // line i = [i ? nl[i-1] + 1 : 0, i < n_nl ? nl[i] : n_bytes); n_nl + 1 lines
    __device__ __forceinline__ uint64_t begin(uint64_t i) const { return i ? nl[i - 1] + 1 : 0; }
    __device__ __forceinline__ uint64_t end(uint64_t i) const { return i < n_nl ? nl[i] : n_bytes; }
};

__device__ __forceinline__ bool is_blank(uint8_t ch) { return ch == ' ' || ch == '\t' || ch == '\r'; }

// bit b of the result = byte b of the 16 is a blank (' ', TAB or CR)
__device__ __forceinline__ uint32_t blank_mask16(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t eq = (__vcmpeq4(w[i], 0x20202020u) | __vcmpeq4(w[i], 0x09090909u) | __vcmpeq4(w[i], 0x0D0D0D0Du)) & 0x08040201u;
        eq = (eq | (eq >> 8) | (eq >> 16) | (eq >> 24)) & 0xFu;
        m |= eq << (4 * i);
    }
    return m;
}

constexpr int kIndexWarps = 8;

// blanks among the 16 bytes at `at` (16-byte aligned) that lie inside [sb, se); 0 when the piece is outside
__device__ __forceinline__ uint32_t blanks_inside(const uint8_t *file, uint64_t at, uint64_t sb, uint64_t se) {
    if (at >= se) return 0;
    uint32_t m = blank_mask16(__ldg(reinterpret_cast<const uint4 *>(file + at)));
    if (at < sb) m &= 0xFFFFu << (uint32_t)(sb - at);
    if (at + 16 > se) m &= 0xFFFFu >> (uint32_t)(at + 16 - se);
    return m;
}

// One warp per 32 consecutive records.  FASTA: lines 2r (header, '>') and 2r+1 (sequence).  FASTQ: lines 4r ('@'
// header), 4r+1 (sequence), 4r+2 ('+'), 4r+3 (quality, as long as the sequence).  Trailing CR / blanks of the sequence
// line are dropped like the host parser does; blanks inside it, or any other arrangement of lines, raise the flag.
// First every lane settles the line boundaries and marker bytes of its own record (the loads of 32 records in flight
// together), then the warp reads the 32 sequence lines one after the other as aligned 16-byte pieces per lane, two
// 512-byte steps requested at a time.
template <bool FASTQ>
__global__ void __launch_bounds__(kIndexWarps * 32)
index_records_kernel(const LineTable t, const uint64_t n_records, uint64_t *__restrict__ rec_start,
                     uint32_t *__restrict__ rec_len, uint32_t *__restrict__ flag) {
    constexpr int PER = FASTQ ? 4 : 2;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t base = ((uint64_t)blockIdx.x * kIndexWarps + (threadIdx.x >> 5)) * 32;
    if (base >= n_records) return;
    const uint64_t r = base + lane;
    const bool valid = r < n_records;
    uint64_t edge[PER + 1]; // edge[j] = end of line PER * r - 1 + j ("-1": the newline in front of the record)
#pragma unroll
    for (int j = 0; j <= PER; j++) {
        const uint64_t line = PER * r + j;
        edge[j] = !valid ? 0 : line == 0 ? ~(uint64_t)0 : (line - 1 < t.n_nl ? t.nl[line - 1] : t.n_bytes);
    }
    uint64_t sb = edge[1] + 1, se = edge[2];
    bool bad = false;
    if (valid) {
        const uint8_t head = t.file[edge[0] + 1]; // an empty header line reads its own '\n' here
        uint8_t plus = '+';
        uint64_t qb = 0, qe = 0;
        if (FASTQ) {
            plus = t.file[edge[2] + 1];
            qb = edge[PER - 1] + 1, qe = edge[PER];
            while (qe > qb && t.file[qe - 1] == '\r') qe--;
        }
        while (se > sb && is_blank(t.file[se - 1])) se--;
        bad = head != (FASTQ ? '@' : '>') || plus != '+' || (se - sb) > 0xFFFFFFFFull;
        if (se > sb) bad |= t.file[sb] == (FASTQ ? '+' : '>');
        if (FASTQ) bad |= (qe - qb) != (se - sb);
        rec_start[r] = sb;
        rec_len[r] = (uint32_t)(se - sb);
    } else {
        sb = se = 0;
    }
    // blanks inside the sequences (the file buffer is 256-byte aligned and padded to whole tiles)
    const uint32_t count = (uint32_t)min((uint64_t)32, n_records - base);
    for (uint32_t i = 0; i < count; i++) {
        const uint64_t b = __shfl_sync(0xFFFFFFFFu, sb, i), e = __shfl_sync(0xFFFFFFFFu, se, i);
        for (uint64_t at = (b & ~(uint64_t)15) + 16 * lane; at < e; at += 2 * 16 * 32) {
            const uint32_t m0 = blanks_inside(t.file, at, b, e), m1 = blanks_inside(t.file, at + 16 * 32, b, e);
            bad |= (m0 | m1) != 0;
        }
    }
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0 && bad) atomicOr(flag, 1u);
}

// ---- wrapped FASTA -> single-line records --------------------------------------------------------------------------
// A header line (first byte '>') opens a record; every other line adds its bytes, without trailing blanks, to the
// record's sequence (blank lines add nothing — the host parser reads the same grammar).  The re-laid file is
// ">\n" + sequence + "\n" per record, so the single-line kernels above index it; blanks INSIDE a line travel with it and
// are caught there.  Line 0 is a header (apc_ingest_fastx checked the file's first byte).
// bytes each line contributes: 2 for the first header (">\n"), 3 for the others ("\n>\n": the newline that closes the
// previous record's sequence), its stripped length otherwise; one thread per line
__global__ void measure_lines_kernel(const LineTable t, uint64_t *__restrict__ line_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > t.n_nl) {
        if (i == t.n_nl + 1) line_len[i] = 0; // the slot that receives the total
        return;
    }
    const uint64_t b = t.begin(i);
    uint64_t e = t.end(i);
    while (e > b && is_blank(t.file[e - 1])) e--;
    line_len[i] = (e > b && t.file[b] == '>') ? (i == 0 ? 2 : 3) : e - b;
}

// one warp per line: the separator bytes of a header, or the line's bytes to their place in the re-laid file
__global__ void __launch_bounds__(kIndexWarps * 32)
unwrap_lines_kernel(const LineTable t, const uint64_t *__restrict__ line_off, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t i = (uint64_t)blockIdx.x * kIndexWarps + (threadIdx.x >> 5);
    if (i > t.n_nl) return;
    const uint64_t off = line_off[i], len = line_off[i + 1] - off, b = t.begin(i);
    if (len && t.file[b] == '>') {
        if (lane < len) out[off + lane] = (i == 0 ? ">\n" : "\n>\n")[lane];
        return;
    }
    for (uint64_t j = lane; j < len; j += 32) out[off + j] = t.file[b + j];
}

// :447-461 — position i of the (shuffled) order is taken when its read has at least 2 * cut bases
__global__ void pick_flag_kernel(const uint32_t *__restrict__ order, const uint64_t n, const uint32_t *__restrict__ rec_len,
                                 const uint64_t min_len, uint32_t *__restrict__ flags, uint32_t *__restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t id = order ? order[i] : i;
    if (id >= n) { // not a permutation of 0..n-1
        atomicOr(flag, 2u);
        flags[i] = 0;
        return;
    }
    flags[i] = rec_len[id] >= min_len ? 1u : 0u;
}

// chosen reads in sampling order, and where each one's sampled end begins in the file: prefix(seq, cut) (:466) starts
// at the sequence, suffix(seq, len - 1 - cut) (:463: cut + 1 bases) cut + 1 bases before its end
__global__ void pick_choose_kernel(const uint32_t *__restrict__ order, const uint64_t n, const uint32_t *__restrict__ flags,
                                   const uint32_t *__restrict__ pos, const uint64_t nb_sample,
                                   const uint64_t *__restrict__ rec_start, const uint32_t *__restrict__ rec_len,
                                   const uint32_t cut, const bool bot, uint64_t *__restrict__ src_off) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i] && pos[i] < nb_sample) {
        const uint32_t id = order ? order[i] : (uint32_t)i;
        src_off[pos[i]] = rec_start[id] + (bot ? (uint64_t)rec_len[id] - 1 - cut : 0);
    }
}

// row j of the sample = row_len bytes of the file from src_off[j].  One thread fills 16 consecutive bytes of the
// row-major sample (one 128-bit store; they may straddle two rows).
__global__ void gather_ends_kernel(const uint8_t *__restrict__ file, const uint64_t *__restrict__ src_off,
                                   const uint64_t n_bytes_out, const uint32_t row_len, uint8_t *__restrict__ stage) {
    const uint64_t idx = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (idx >= n_bytes_out) return;
    uint64_t j = idx / row_len;
    uint32_t o = (uint32_t)(idx - j * row_len);
    const uint32_t n = (uint32_t)min((uint64_t)16, n_bytes_out - idx);
    const uint64_t rows = n_bytes_out / row_len;
    const uint8_t *src = file + src_off[j];
    uint64_t next_off = j + 1 < rows ? src_off[j + 1] : 0; // requested together with this row's
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (uint32_t b = 0; b < 16; b++) {
        if (b < n) {
            w[b >> 2] |= (uint32_t)src[o] << (8 * (b & 3));
            if (++o == row_len) { // on to the next row (rows shorter than 16 bytes: more than once)
                o = 0;
                j++;
                src = file + next_off;
                if (row_len < 16 && j + 1 < rows) next_off = src_off[j + 1];
            }
        }
    }
    if (n == 16) {
        *reinterpret_cast<uint4 *>(stage + idx) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
        for (uint32_t b = 0; b < 16; b++)
            if (b < n) stage[idx + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
    }
}

static unsigned grid_for(uint64_t items, int per_block) {
    return (unsigned)std::min<uint64_t>((items + per_block - 1) / per_block, 0x7FFFFFFFull);
}

cudaError_t launch_count_newlines(const uint8_t *d_file, uint64_t n_tiles, uint64_t *d_tile_nl, cudaStream_t s) {
    if (!n_tiles) return cudaSuccess;
    count_newlines_kernel<<<grid_for(n_tiles, kIngestThreads / 32), kIngestThreads, 0, s>>>(reinterpret_cast<const uint4 *>(d_file), n_tiles,
                                                                         d_tile_nl);
    return cudaGetLastError();
}

cudaError_t launch_write_newlines(const uint8_t *d_file, uint64_t n_tiles, const uint64_t *d_tile_base, uint64_t *d_nl,
                                  cudaStream_t s) {
    if (!n_tiles) return cudaSuccess;
    write_newlines_kernel<<<grid_for(n_tiles, kIngestThreads / 32), kIngestThreads, 0, s>>>(reinterpret_cast<const uint4 *>(d_file), n_tiles,
                                                                         d_tile_base, d_nl);
    return cudaGetLastError();
}

// exclusive prefix sums (in place for the u64 form); d_temp == nullptr: only the temp size is returned
cudaError_t ingest_prefix_u64(void *d_temp, size_t &temp_bytes, uint64_t *d_inout, uint64_t n, cudaStream_t s) {
    return cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_inout, d_inout, (int64_t)n, s);
}

cudaError_t ingest_prefix_u32(void *d_temp, size_t &temp_bytes, const uint32_t *d_in, uint32_t *d_out, uint64_t n,
                              cudaStream_t s) {
    return cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_in, d_out, (int64_t)n, s);
}

cudaError_t launch_index_records(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes, bool fastq,
                                 uint64_t n_records, uint64_t *d_rec_start, uint32_t *d_rec_len, uint32_t *d_flag,
                                 cudaStream_t s) {
    if (!n_records) return cudaSuccess;
    const LineTable t{d_file, d_nl, n_nl, n_bytes};
    const unsigned grid = grid_for(n_records, kIndexWarps * 32);
    if (fastq) index_records_kernel<true><<<grid, kIndexWarps * 32, 0, s>>>(t, n_records, d_rec_start, d_rec_len, d_flag);
    else index_records_kernel<false><<<grid, kIndexWarps * 32, 0, s>>>(t, n_records, d_rec_start, d_rec_len, d_flag);
    return cudaGetLastError();
}

cudaError_t launch_measure_lines(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes,
                                 uint64_t *d_line_len, cudaStream_t s) {
    const LineTable t{d_file, d_nl, n_nl, n_bytes};
    measure_lines_kernel<<<grid_for(n_nl + 2, 256), 256, 0, s>>>(t, d_line_len);
    return cudaGetLastError();
}

cudaError_t launch_unwrap_lines(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes,
                                const uint64_t *d_line_off, uint8_t *d_out, cudaStream_t s) {
    const LineTable t{d_file, d_nl, n_nl, n_bytes};
    unwrap_lines_kernel<<<grid_for(n_nl + 1, kIndexWarps), kIndexWarps * 32, 0, s>>>(t, d_line_off, d_out);
    return cudaGetLastError();
}

cudaError_t launch_pick_reads(const uint32_t *d_order, uint64_t n, const uint64_t *d_rec_start, const uint32_t *d_rec_len,
                              uint32_t cut, bool bot, uint64_t nb_sample, uint32_t *d_flags, uint32_t *d_pos,
                              uint64_t *d_src_off, void *d_temp, size_t temp_bytes, uint32_t *d_flag, cudaStream_t s) {
    if (!n) return cudaSuccess;
    pick_flag_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_order, n, d_rec_len, 2ull * cut, d_flags, d_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if ((e = ingest_prefix_u32(d_temp, temp_bytes, d_flags, d_pos, n, s)) != cudaSuccess) return e;
    pick_choose_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_order, n, d_flags, d_pos, nb_sample, d_rec_start, d_rec_len, cut, bot,
                                                      d_src_off);
    return cudaGetLastError();
}

cudaError_t launch_gather_ends(const uint8_t *d_file, const uint64_t *d_src_off, uint64_t n_sampled, uint32_t row_len,
                               uint8_t *d_stage, cudaStream_t s) {
    const uint64_t n_out = n_sampled * row_len;
    if (!n_out) return cudaSuccess;
    gather_ends_kernel<<<grid_for((n_out + 15) / 16, 256), 256, 0, s>>>(d_file, d_src_off, n_out, row_len, d_stage);
    return cudaGetLastError();
}

} // namespace apc
