// apc_internal.h — context layout and launch declarations shared by libapc's
// translation units.  Not part of the ABI (see include/apc.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "apc.h"

namespace apc {

// ---- scan-tile geometry -------------------------------------------------------
// The sampled text lives in HBM as tiles of 32 reads (one read per lane).  A
// tile is stored column-chunk-major: uint4 tile[chunks][32]; each uint4 holds
// 16 consecutive bases of one read, one byte per base, already scaled to the
// byte offset of that base's row in the 16-byte-per-row match table:
//   A=0x00 C=0x10 G=0x20 T=0x30 N/pad=0x40.
// A warp's 128-bit loads of one chunk are therefore one coalesced 512-byte
// request, and a base costs one byte-extract before it becomes an LDS address.
constexpr int kTileReads = 32;
constexpr int kChunkBases = 16;
constexpr uint32_t kCodeN = 0x40;
constexpr int kPeqRows = 5;        // A C G T N
constexpr int kWordsPerThread = 4; // u32 state words per thread = one 16-byte table row
constexpr int kScanWarps = 8;      // warps per CTA of the scan kernel

// ---- bit-sliced kernel: unit shapes -----------------------------------------------------------
// A unit is G k-mers that share their first K - T bases (rows) and is scanned by one warp
// (bs_group_kernel<K, K - T, G>).  bs_shape(k, s) is shape s of the table for this k, g == 0 where
// the shape does not exist (fewer than two shared rows, or more rows than the register file holds
// without spills).  Heaviest units first: that is also the launch order.
struct BsShape {
    int t, g;
};
constexpr int kBsShapes = 22;
constexpr int kBsMaxRows = 60;
constexpr int kBsAlivePct = 30;
constexpr uint32_t kBsMediumShapes = 0x3FFC00u; // shapes 10..21 of bs_shape: units of at most 48 rows at k = 16
constexpr uint32_t kBsSmallShapes = 0x3F0000u; // shapes 16..21 of bs_shape: units of at most 31 rows at k = 16 (small samples)
constexpr BsShape bs_shape(int k, int s) {
    // round 2: units of up to 60 rows still compile to 255 registers with at most a few dozen bytes of spills, and
    // ten larger shapes — greedy selection over the C2 / C3 / C4 query sets with the planner's cost model, each kept
    // only where the GPU agreed (tools/ab_shapes.sh; (8, 6) and (7, 7) were chosen by the model and lost to their
    // spills) — lead the table: C3 676 -> 838, C4 652 -> 857, C2 385 -> 408 kGCUPS
    const BsShape table[kBsShapes] = {{4, 10}, {9, 5}, {3, 13}, {6, 7}, {5, 9}, {3, 14}, {10, 4}, {7, 5}, {8, 5}, {6, 5}, {2, 16},
                                      {6, 6},  {5, 6}, {8, 4},  {2, 12}, {3, 8}, {6, 3}, {2, 8},  {3, 6},  {3, 4}, {1, 4}, {k - k / 2, 2}};
    const BsShape sh = table[s];
    const int p = k - sh.t;
    if (p < 2 || sh.t < 1 || p + sh.g * sh.t > kBsMaxRows) return BsShape{0, 0};
    for (int j = 0; j < s; j++)
        if (table[j].t == sh.t && table[j].g == sh.g) return BsShape{0, 0};
    return sh;
}
// Rows >= bs_check_row_host(k) of a unit are computed only in the columns where something can reach them
// (dead-row skipping, bitslice_core.cuh); k = no such split.
constexpr int bs_check_row_host(int k) {
#ifdef APC_BS_CHECK_ROW
    return APC_BS_CHECK_ROW < k ? APC_BS_CHECK_ROW : k; // A/B builds (tools/build_ab.sh)
#else
    // A/B on the BASELINE workloads (tools/ab_skip.sh): k = 16: row 12 / 13 / 14 -> 346 / 365 / 353 kGCUPS;
    // k = 20: 549 / 618 / 625; k = 32: 13 / 14 -> 556 / 585
    // round 2, with the larger units (tools/ab_bench.sh): k = 16: 12 / 13 / 14 -> 371 / 403 / 382; k = 20: 13 / 14 / 15 ->
    // 788 / 831 / 795; k = 32: 13 / 14 / 15 -> 784 / 859 / 866
    return k >= 28 ? 15 : k >= 18 ? 14 : k >= 16 ? 13 : k;
#endif
}
// The scan kernels prefetch one column pair past either end of a read: the plane buffer is padded by that much.
constexpr int kBsPadCols = 2;
struct BsRange { // super-groups (1024 reads) and reads [lo, hi) of one scan
    uint32_t sg_first, n_sg;
    uint64_t lo, hi;
};

struct ScanVariant {
    int nw; // u32 words per unit (1 or 2); 0 = bit-sliced kernel (bitslice_kernel.cu)
    int f;  // k-mers interleaved per unit
    bool bitslice() const { return nw == 0; }
    bool pairing() const { return nw == 0 && f == 1; } // f == 0: bit-sliced without k-mer pairing (scan_variant 7)
    int queries_per_group() const { return nw ? f * kWordsPerThread / nw : 1; }
    int rows_per_unit() const { return nw ? 32 * nw / f : 32; } // complete k-mer rows a unit can hold
};

// device scratch of the exact stage (exact_kernels.cu); grown on demand, kept
// across calls so the per-end loop does not re-allocate
struct ExactScratch {
    void *d_keys_a = nullptr;  size_t keys_a_cap = 0;  // key stream, later the unique keys
    void *d_keys_b = nullptr;  size_t keys_b_cap = 0;  // sorted keys
    void *d_temp = nullptr;    size_t temp_cap = 0;    // cub temp storage
    uint32_t *d_start = nullptr; size_t start_cap = 0; // run starts, d+1 entries
    uint16_t *d_dsum = nullptr;  size_t dsum_cap = 0;  // dimer sums of the unique keys
    uint32_t *d_block_a = nullptr; size_t block_a_cap = 0;
    uint32_t *d_block_b = nullptr; size_t block_b_cap = 0;
    unsigned long long *d_hist = nullptr; size_t hist_cap = 0;
    unsigned long long *d_counters = nullptr; size_t counters_cap = 0;
    uint64_t *d_forb = nullptr; size_t forb_cap = 0;
    uint64_t *d_sel_k = nullptr; size_t sel_k_cap = 0;
    uint64_t *d_sel_c = nullptr; size_t sel_c_cap = 0;
    uint64_t launches = 0;
};

struct Ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {}; // upload 0-1, scan 2-3, approx 4-5, exact stage 6-7
    cudaEvent_t ev_table = nullptr; // completion of the last match-table copy out of h_pinned
    bool table_copy_pending = false;
    std::string err;

    // sample
    uint4 *d_tiles = nullptr;
    size_t tiles_bytes = 0;
    uint4 *d_planes = nullptr;  // bit planes of the same text (bitslice_kernel.cu), [super-group][column][32 groups],
                                // kBsPadCols columns of padding in front and behind (the kernels prefetch past both ends)
    size_t planes_cap = 0, planes_bytes = 0;
    uint4 *planes() const { return d_planes + kBsPadCols * 32; }
    uint32_t *d_lens = nullptr; // per-read length — exact stage
    size_t lens_cap = 0;
    bool has_sample = false;
    uint64_t n_reads = 0;
    uint32_t max_len = 0;
    uint64_t total_bases = 0;
    uint32_t n_tiles = 0;
    uint32_t chunks = 0; // uint4 chunks per read
    bool uniform_len = true;

    // queries
    uint8_t k = 0;
    uint32_t n_kmers = 0;
    ScanVariant variant{1, 1};
    uint32_t n_groups = 0;
    uint64_t *d_kmers = nullptr; // bit-sliced kernel: the query k-mers in scan order (units of shape 0, 1, ..., then
                                 // the ungrouped ones; members of suffix-sharing units reversed), then u32 perm[n]
    uint32_t bs_units[kBsShapes] = {}; // units per shape
    size_t kmers_cap = 0;
    uint32_t *d_peq = nullptr; // [n_groups][5][4]
    size_t peq_cap = 0;
    unsigned long long *d_counts = nullptr; // [n_groups * queries_per_group]
    size_t counts_cap = 0;
    unsigned int *d_job_counter = nullptr;  // job queue heads of the persistent scan kernels (self re-arming), one per
                                            // concurrent launch: [kBsShapes + 1]
    unsigned long long *d_deep_lop3 = nullptr; // LOP3 warp instructions the scan kernels spent on deep rows (dead-row
                                               // skipping makes that data dependent) since the last apc_scan_stats_read
    // host side of the same statistics, accumulated per scan from the plan
    mutable double stat_lop3_top = 0., stat_lop3_all = 0., stat_lop3_single = 0.;
    mutable uint64_t stat_scans = 0;
    cudaStream_t bs_streams[kBsShapes] = {}; // side streams of the bit-sliced scan (one launch per shape, concurrent)
    cudaEvent_t bs_join[kBsShapes] = {};
    cudaEvent_t bs_fork = nullptr;

    // CUDA graph of the launches of one scan (fork, one kernel per shape in use, join): captured when the same
    // scan is issued a second time, replayed afterwards — a C1-sized scan is launch-bound otherwise
    cudaGraphExec_t scan_graph = nullptr;
    uint64_t plan_gen = 0;          // bumped by every upload / set_queries / option change that affects a scan
    uint64_t graph_gen = ~0ull;     // plan_gen the graph was captured for (~0 = none)
    uint64_t last_scan_gen = ~0ull; // plan_gen of the previous scan
    unsigned long long *graph_dst = nullptr, *last_scan_dst = nullptr;
    cudaStream_t graph_stream = nullptr, last_scan_stream = nullptr;
    uint64_t graph_launches = 0;
    double graph_lop3[3] = {0., 0., 0.}; // the plan statistics of one replay (top, all, one k-mer per warp)
    int opt_graph = 0; // 1 = a repeated scan is captured and replayed (apc_set_option "scan_graph")

    // multi-GPU: an NCCL communicator (one rank per context), libnccl loaded on first use (apc_comm.cpp)
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;

    // staging
    uint8_t *d_stage = nullptr;
    size_t stage_cap = 0;
    uint64_t *d_stage_offs = nullptr;
    size_t stage_offs_cap = 0;
    void *h_pinned = nullptr;
    size_t pinned_cap = 0;

    // resident input file and its record index (device ingest, ingest_kernels.cu)
    uint8_t *d_file = nullptr;       // the file's bytes, padded with zeros to a whole number of tiles
    size_t file_cap = 0;
    uint8_t *d_file2 = nullptr;      // wrapped FASTA re-laid as single-line records is built here, then swapped with d_file
    size_t file2_cap = 0;
    uint64_t *d_line_off = nullptr;  // per line of a wrapped file: bytes it contributes to the re-laid file, then their prefix sum
    size_t line_off_cap = 0;
    bool file_unwrapped = false;     // the resident bytes are the re-laid form, not the caller's
    uint64_t *d_tile_nl = nullptr;   // newlines per tile, then (in place) their exclusive prefix sum; [tiles + 1]
    size_t tile_nl_cap = 0;
    uint64_t *d_nl = nullptr;        // byte offsets of the newlines, ascending
    size_t nl_cap = 0;
    uint64_t *d_rec_start = nullptr; // per record: byte offset of its sequence line
    size_t rec_start_cap = 0;
    uint32_t *d_rec_len = nullptr;   // per record: bases on that line
    size_t rec_len_cap = 0;
    uint32_t *d_pick = nullptr;      // sampling scratch: order | eligible flags | their prefix sum (n u32 each) | file offsets of the chosen ends (n u64)
    size_t pick_cap = 0;
    uint8_t *h_file_stage[2] = {nullptr, nullptr}; // page-locked staging of the file copy (two pieces in flight)
    cudaEvent_t ev_file_stage[2] = {};
    int opt_ingest_staging = 1;      // 0: hand the caller's (pageable) buffer to cudaMemcpyAsync directly
    void *d_ingest_temp = nullptr;   // cub temp storage of the two prefix sums
    size_t ingest_temp_cap = 0;
    uint32_t *d_ingest_flag = nullptr; // [0] malformed input seen by a kernel
    cudaEvent_t ev_ingest[4] = {};   // copy 0-1, index 1-2, sample 2-3 (of the most recent calls)
    bool has_file = false;
    bool stage_is_sample = false;    // d_stage holds the ASCII rows of the resident sample (apc_download_sample)
    int file_fastq = 0;
    uint64_t file_bytes = 0;         // without trailing blank lines
    uint64_t n_records = 0;
    float ingest_ms[3] = {0.f, 0.f, 0.f}; // copy, index, sample

    // options
    int opt_variant = 0;
    int opt_tiles_per_job = 0;
    uint32_t opt_shape_mask = 0xFFFFFFFFu; // unit shapes the bit-sliced scan may use (bit s = shape s of bs_shape)
    int opt_alive_pct = kBsAlivePct;       // planner: expected share of columns in which deep rows are computed
    uint64_t opt_first_read = 0;  // scan only reads [first, first+n) of the resident sample
    int64_t opt_n_reads = -1;     // -1 = to the end

    apc_timing timing{};
    ExactScratch exact;

    // upper bound of the number of k-windows of the resident sample (:496)
    uint64_t total_windows(uint32_t k) const {
        if (uniform_len) return max_len >= k ? n_reads * (uint64_t)(max_len - k + 1) : 0;
        return total_bases;
    }
};

int fail(Ctx *c, int status, const char *what, cudaError_t e = cudaSuccess);

#define APC_CUDA(ctx, call)                                                   \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ::apc::fail((ctx), APC_ERR_CUDA, #call, e__); \
    } while (0)

// sample_kernels.cu
cudaError_t launch_build_tiles_uniform(const uint8_t *d_ascii, uint64_t n_reads, uint32_t read_len,
                                       uint32_t chunks, uint32_t n_tiles, uint4 *d_tiles,
                                       uint32_t *d_lens, cudaStream_t s);
cudaError_t launch_build_tiles_ragged(const uint8_t *d_ascii, const uint64_t *d_offs, uint64_t n_reads,
                                      uint32_t chunks, uint32_t n_tiles, uint4 *d_tiles,
                                      uint32_t *d_lens, cudaStream_t s);

// ingest_kernels.cu
constexpr uint64_t kIngestTileBytes = 4096;  // one warp: 8 steps of 32 x 16 bytes
constexpr uint64_t kIngestStageBytes = (uint64_t)16 << 20; // one staging piece of the file copy
cudaError_t launch_count_newlines(const uint8_t *d_file, uint64_t n_tiles, uint64_t *d_tile_nl, cudaStream_t s);
cudaError_t launch_write_newlines(const uint8_t *d_file, uint64_t n_tiles, const uint64_t *d_tile_base, uint64_t *d_nl,
                                  cudaStream_t s);
cudaError_t ingest_prefix_u64(void *d_temp, size_t &temp_bytes, uint64_t *d_inout, uint64_t n, cudaStream_t s);
cudaError_t ingest_prefix_u32(void *d_temp, size_t &temp_bytes, const uint32_t *d_in, uint32_t *d_out, uint64_t n,
                              cudaStream_t s);
cudaError_t launch_index_records(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes, bool fastq,
                                 uint64_t n_records, uint64_t *d_rec_start, uint32_t *d_rec_len, uint32_t *d_flag,
                                 cudaStream_t s);
cudaError_t launch_measure_lines(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes,
                                 uint64_t *d_line_len, cudaStream_t s);
cudaError_t launch_unwrap_lines(const uint8_t *d_file, const uint64_t *d_nl, uint64_t n_nl, uint64_t n_bytes,
                                const uint64_t *d_line_off, uint8_t *d_out, cudaStream_t s);
cudaError_t launch_pick_reads(const uint32_t *d_order, uint64_t n, const uint64_t *d_rec_start, const uint32_t *d_rec_len,
                              uint32_t cut, bool bot, uint64_t nb_sample, uint32_t *d_flags, uint32_t *d_pos,
                              uint64_t *d_src_off, void *d_temp, size_t temp_bytes, uint32_t *d_flag, cudaStream_t s);
cudaError_t launch_gather_ends(const uint8_t *d_file, const uint64_t *d_src_off, uint64_t n_sampled, uint32_t row_len,
                               uint8_t *d_stage, cudaStream_t s);

// scan_kernel.cu
ScanVariant pick_variant(int k, int forced);
void build_peq_tables(const uint64_t *kmers, uint32_t n_kmers, int k, ScanVariant v, uint32_t *table,
                      uint32_t &n_groups);
cudaError_t launch_scan(const Ctx &c, unsigned long long *d_counts, uint64_t *launches);

// bitslice_kernel.cu
cudaError_t launch_build_planes(const Ctx &c);
cudaError_t launch_bs_scan(const Ctx &c, uint64_t read_lo, uint64_t read_hi, unsigned long long *d_counts,
                           uint32_t sg_per_job, uint64_t *launches);
void bs_group_queries(const uint64_t *kmers, uint32_t n, int k, uint32_t shape_mask, float alive,
                      std::vector<uint32_t> &order, std::vector<uint8_t> &reversed, uint32_t (&units)[kBsShapes]);
uint64_t bs_reverse_kmer(uint64_t kmer, int k);
cudaError_t warm_bs_kernels(const Ctx &c, int k);

// exact_kernels.cu
int exact_count_select(Ctx *c, uint8_t k, float lc_adjusted, uint64_t lim, uint64_t solid_km,
                       const uint64_t *forbidden, uint64_t n_forbidden, uint64_t capacity,
                       std::vector<uint64_t> &kmers, std::vector<uint64_t> &counts, uint64_t *n_needed,
                       uint64_t *n_distinct, uint64_t *n_had_n);
void free_exact_scratch(Ctx *c);
int exact_reserve(Ctx *c, uint8_t k, uint64_t max_windows);

// peak_kernels.cu
cudaError_t measure_int_peak(const Ctx &c, double *lop3, double *imad, double *mixed);
cudaError_t microbench(const Ctx &c, const char *name, double *value);

} // namespace apc

// host/host_util.cpp: memcpy split over the host threads (OpenMP)
namespace apch {
void parallel_copy(void *dst, const void *src, size_t n);
}

struct apc_ctx : public apc::Ctx {};
