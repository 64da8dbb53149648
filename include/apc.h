/*
 * apc.h — C ABI of libapc, the B200 (sm_100a) implementation of
 * approx_counter's counting path.
 *
 * The reference (qbonenfant/approx_counter, one C++ translation unit) has no
 * plugin / FFI layer: its hot path is two static functions called from main().
 * This header is the boundary a maintainer would bind instead of those calls;
 * each entry point cites the reference code it replaces
 * (/root/reference/approx_counter.cpp:line).  INTEGRATION.md shows the patch.
 *
 * Conventions: plain C types, caller-owned HOST buffers unless a parameter is
 * documented as a device pointer, `int` status return (APC_OK or a negative
 * APC_ERR_*), no exceptions cross the ABI.  One apc_ctx drives one GPU; a
 * context is not re-entrant (one call at a time per context).  There is no
 * CPU fallback: without a usable CUDA device apc_create fails.
 *
 * k-mers are 2-bit packed exactly like the reference's dna2int (:55-62):
 * A=0 C=1 G=2 T=3, first base in the most significant used bits, k <= 32.
 */
#ifndef APC_H
#define APC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APC_VERSION 2

#define APC_OK 0
#define APC_ERR_INVALID (-1)    /* bad argument (NULL, k outside [2,32], ...) */
#define APC_ERR_CUDA (-2)       /* a CUDA runtime call failed; see apc_last_error */
#define APC_ERR_NO_DEVICE (-3)  /* no CUDA device / device index out of range */
#define APC_ERR_NO_SAMPLE (-4)  /* a stage needs apc_upload_sample first */
#define APC_ERR_NO_QUERIES (-5) /* apc_scan before apc_set_queries */
#define APC_ERR_NOMEM (-6)      /* host or device allocation failed */
#define APC_ERR_CAPACITY (-7)   /* caller-provided output capacity too small */
#define APC_ERR_COMM (-8)       /* NCCL unavailable or a collective failed; see apc_last_error */
#define APC_ERR_FORMAT (-9)     /* apc_ingest_fastx: input outside the device parser's grammar */
#define APC_MAXERR 2            /* compile-time edit bound, reference :25 */

typedef struct apc_ctx apc_ctx;

/* Per-stage device times of the most recent calls, CUDA events on the
 * context's stream, milliseconds. */
typedef struct apc_timing {
    float upload_ms;  /* H2D + layout kernel of apc_upload_sample        */
    float exact_ms;   /* apc_exact_topn device work                      */
    float scan_ms;    /* approximate-count kernel(s) of the last scan    */
    float total_ms;   /* whole last apc_approx_count incl. H2D/D2H       */
    uint64_t scan_launches;  /* kernels launched by the last scan        */
    uint64_t exact_launches; /* kernels launched by the last exact stage */
} apc_timing;

int apc_version(void);
const char *apc_strerror(int status);
/* number of visible CUDA devices, or a negative APC_ERR_* */
int apc_device_count(void);

/* Bind a context to CUDA device `device` and create its stream.  Replaces
 * nothing in the reference (which is CPU/OpenMP, :547); this is the one-time
 * set-up a host adds before the per-end loop (:858). */
int apc_create(int device, apc_ctx **out);
void apc_destroy(apc_ctx *ctx);
/* Text of the last failure on this context ("" if none). */
const char *apc_last_error(const apc_ctx *ctx);
/* Run subsequent work on an existing cudaStream_t (e.g. a torch stream) so
 * the caller can bracket it with its own events.  NULL = the context's own
 * stream; for the legacy default stream pass cudaStreamLegacy (0x1). */
int apc_set_stream(apc_ctx *ctx, void *cuda_stream);
int apc_sync(apc_ctx *ctx);

/* Optional one-time set-up for hosts that run once and exit (the drop-in binary
 * calls it beside the parsing of its input, :819-825): allocates the device
 * buffers of a sample of n_reads x read_len, the exact stage's scratch for its
 * k-windows and the tables of n_kmers queries, and runs the whole path once on
 * a 64-read dummy sample so that every kernel the real calls launch is loaded
 * (CUDA loads kernels lazily at their first launch).  Leaves the context
 * without sample and queries; every other entry point works without it. */
int apc_reserve(apc_ctx *ctx, uint64_t n_reads, uint32_t read_len, uint8_t k,
                uint32_t n_kmers);

/* ---- the sampled text ------------------------------------------------------
 * Replaces the `sequence_set_type sample` that sampleSequences returns (:415,
 * :867) and errorCount indexes (:537-541): instead of a SeqAn FM index the
 * sampled read ends live in HBM as scan tiles (32 reads, column-major, one
 * pre-scaled code byte per base; N and padding never match).
 * `bases`: ASCII, row-major, n_reads x read_len (sl for starts, sl+1 for
 * ends, :463/:466).  Letters other than ACGTacgt are N (Dna5, :38). */
int apc_upload_sample(apc_ctx *ctx, const uint8_t *bases, uint64_t n_reads,
                      uint32_t read_len);
/* Same, without the final synchronisation: `bases` must stay valid (and, to
 * overlap with other work, be page-locked) until apc_sync returns. */
int apc_upload_sample_async(apc_ctx *ctx, const uint8_t *bases, uint64_t n_reads,
                            uint32_t read_len);
/* Ragged form: read r is bases[offsets[r] .. offsets[r+1]). */
int apc_upload_sample_ragged(apc_ctx *ctx, const uint8_t *bases,
                             const uint64_t *offsets, uint64_t n_reads);
int apc_sample_info(const apc_ctx *ctx, uint64_t *n_reads, uint32_t *max_len,
                    uint64_t *total_bases);

/* ---- ingest on the device: FASTA / FASTQ bytes -> sampled read ends -----------
 * Replaces readRecords (:819-825) and the walk and copies of sampleSequences
 * (:447-471): the file's bytes are copied to HBM once, the records are indexed
 * there (newline index + a grammar check per record), and every sample is
 * gathered from the resident bytes straight into the staging buffer of the
 * layout kernels.  Taken: FASTA ('>' header lines; a record's other lines,
 * trailing blanks dropped, are its sequence — wrapped sequences, blank lines
 * and headers without a sequence line included: such files are re-laid as one
 * line per sequence on the device first) and 4-line FASTQ ('@' header,
 * sequence, '+' line, quality of the sequence's length); LF or CRLF.  Refused
 * with APC_ERR_FORMAT, and left to the host parser (apch_reads_load): wrapped
 * FASTQ, and blanks INSIDE a sequence line.  `file_bytes` is a HOST pointer (a
 * read-only mapping of the file will do); it is not needed after the call
 * returns. */
int apc_ingest_fastx(apc_ctx *ctx, const uint8_t *file_bytes, uint64_t n_bytes,
                     uint64_t *n_records_out, int *is_fastq_out);
/* seqs[first .. first+n) lengths (`length(sequence_set[id])`, :461) of the
 * resident file. */
int apc_ingest_lengths(apc_ctx *ctx, uint64_t first, uint64_t n,
                       uint32_t *lens_out);
/* sampleSequences (:415-476) over the resident file: walks `order` (HOST, the
 * shuffled read ids of :423-429 — apch_shuffle_order gives the reference's
 * mt19937 + std::shuffle; NULL = 0, 1, 2, ...), takes the first nb_sample reads
 * of length >= 2*cut (:447-461) and makes their first `cut` bases (bot == 0,
 * :466) or last cut+1 bases (bot != 0, :463) the context's sample, exactly as
 * apc_upload_sample would with the rows sampleSequences returns.  n_order must
 * equal the record count.  Synchronises. */
int apc_sample_resident(apc_ctx *ctx, const uint32_t *order, uint64_t n_order,
                        uint64_t nb_sample, uint32_t cut, int bot,
                        uint64_t *n_sampled_out);
/* Multi-GPU hosts: rows [first_read, first_read + n_reads) of the ASCII sample
 * resident in `src` (another context, normally another GPU, after
 * apc_sample_resident or apc_upload_sample there) become the sample of `ctx`,
 * copied GPU to GPU (cudaMemcpyPeerAsync: over NVLink where the devices are
 * peers) — a rank's shard of the reads without a trip through the host.  src's
 * sample must be complete (apc_sample_resident and apc_upload_sample
 * synchronise) and must not be replaced before apc_sync(ctx) returns;
 * asynchronous like apc_upload_sample_async. */
int apc_upload_sample_peer(apc_ctx *ctx, const apc_ctx *src, uint64_t first_read,
                           uint64_t n_reads);
/* ASCII rows (n_reads x read_len, see apc_sample_info) of the resident sample
 * as uploaded or gathered — what sampleSequences would have returned.  Not
 * available after apc_upload_sample_ragged. */
int apc_download_sample(apc_ctx *ctx, uint8_t *bases_out, uint64_t capacity);
/* Device times of the last apc_ingest_fastx (H2D copy of the file; newline
 * index + record index kernels) and apc_sample_resident (pick + gather +
 * layout kernels), CUDA events, milliseconds.  Any pointer may be NULL. */
int apc_ingest_timing(const apc_ctx *ctx, float *copy_ms, float *index_ms,
                      float *sample_ms);

/* ---- exact count + filter + top-N -------------------------------------------
 * Replaces count_kmers (:487-519, called :874) followed by get_most_frequent
 * (:396-405, called :898): counts every N-free window of the uploaded sample
 * that passes the low-complexity filter (:214-234, float32 `>=`, threshold
 * already adjusted by :183-186) and is not in `forbidden` (:330-332), then
 * returns the first `lim` k-mers in CompareCount order (:275-305: count desc,
 * complexity asc, k-mer value desc).  kmers_out/counts_out hold >= lim
 * entries; *n_out = number written.  n_distinct/n_had_n (optional) receive
 * count.size() (:883) and the skipped-window tally (:506). */
int apc_exact_topn(apc_ctx *ctx, uint8_t k, float lc_adjusted, uint64_t lim,
                   const uint64_t *forbidden, uint64_t n_forbidden,
                   uint64_t *kmers_out, uint64_t *counts_out, uint64_t *n_out,
                   uint64_t *n_distinct, uint64_t *n_had_n);
/* Replaces get_solid_kmers (:372-388, called :892): every k-mer with count >=
 * solid_km, in CompareCount order (the reference leaves ties unspecified).
 * Returns APC_ERR_CAPACITY (and the needed size in *n_out) if capacity is
 * too small. */
int apc_exact_solid(apc_ctx *ctx, uint8_t k, float lc_adjusted, uint64_t solid_km,
                    const uint64_t *forbidden, uint64_t n_forbidden,
                    uint64_t *kmers_out, uint64_t *counts_out, uint64_t capacity,
                    uint64_t *n_out, uint64_t *n_distinct, uint64_t *n_had_n);

/* ---- approximate count: the hot path ------------------------------------------
 * Replaces errorCount (:531-601, called :922) minus its index build: for each
 * k-mer, counts_out[i] = sum over e=0..2 of the number of sampled reads
 * flagged at error level e (:553-565, :589-596), i.e. sum over reads of
 * max(0, 3 - d) with d the minimum edit distance between the k-mer and any
 * substring of the read (text N matches nothing).  Only the k-mer values are
 * read, as in :584.  Host in, host out; H2D/D2H inside the call. */
int apc_approx_count(apc_ctx *ctx, uint8_t k, const uint64_t *kmers,
                     uint32_t n_kmers, uint64_t *counts_out);
/* Same, enqueued on the context's stream without waiting: counts_out is
 * filled once apc_sync returns.  `kmers` is consumed before the call returns.
 * Two contexts on one device (e.g. read starts and read ends) overlap the
 * upload of one sample with the scan of the other this way. */
int apc_approx_count_async(apc_ctx *ctx, uint8_t k, const uint64_t *kmers,
                           uint32_t n_kmers, uint64_t *counts_out);

/* Split-phase form of the same call for resident data and multi-GPU use:
 * set_queries uploads the match tables, scan launches the kernel on the
 * context's stream (asynchronous; d_counts is a DEVICE pointer to n_kmers
 * uint64 that the kernel overwrites, or NULL to use the context's own
 * buffer), get_counts synchronises and copies the context's buffer back.
 * A multi-GPU host scans one read shard per GPU and sums the count vectors
 * (one all-reduce of n_kmers uint64 — see INTEGRATION.md). */
int apc_set_queries(apc_ctx *ctx, uint8_t k, const uint64_t *kmers,
                    uint32_t n_kmers);
int apc_scan(apc_ctx *ctx, uint64_t *d_counts);
int apc_get_counts(apc_ctx *ctx, uint64_t *counts_out);
/* Device address of the context's own count buffer (n_kmers uint64). */
uint64_t *apc_counts_device_ptr(apc_ctx *ctx);

/* ---- multi-GPU: reads sharded, counts summed ------------------------------------
 * Replaces the reference's only parallelism, the OpenMP team over k-mers that
 * shares one index (:547-599) and its `omp critical` result merge (:595-596):
 * the sampled reads are independent and a count is a sum over reads (:589-596),
 * so every GPU scans one shard of the reads for ALL k-mers and the per-k-mer
 * count vectors are summed with one small all-reduce over NVLink.  One context
 * (= one GPU) is one rank; the ranks may live in one process (a host thread
 * per GPU) or in one process per GPU.  libnccl.so.2 is loaded on first use.
 *
 *   rank 0:     apc_comm_unique_id(id)  -> hand `id` to every rank (any channel)
 *   every rank: apc_comm_init_rank(ctx, n_ranks, rank, id)     (collective)
 *   every rank: apc_scan_allreduce(ctx, NULL) [or apc_scan x2 ends, then one
 *               apc_allreduce_counts over both vectors], apc_get_counts
 */
#define APC_COMM_ID_BYTES 128
int apc_comm_unique_id(uint8_t id_out[APC_COMM_ID_BYTES]);
int apc_comm_init_rank(apc_ctx *ctx, int n_ranks, int rank,
                       const uint8_t id[APC_COMM_ID_BYTES]);
int apc_comm_destroy(apc_ctx *ctx);
/* rank / size of the context's communicator (0 / 1 without one) */
int apc_comm_info(const apc_ctx *ctx, int *rank, int *n_ranks);
/* In-place sum over the ranks of n uint64 at the DEVICE pointer d_counts (NULL:
 * the context's own count buffer, n ignored), enqueued on the context's
 * stream.  A context without a communicator returns APC_OK without doing
 * anything, so single- and multi-GPU hosts share one code path. */
int apc_allreduce_counts(apc_ctx *ctx, uint64_t *d_counts, uint64_t n);
/* apc_scan followed by apc_allreduce_counts on the same stream: the whole
 * errorCount of a sharded sample (:531-601), asynchronous. */
int apc_scan_allreduce(apc_ctx *ctx, uint64_t *d_counts);

int apc_last_timing(const apc_ctx *ctx, apc_timing *out);

/* What the scans since the previous call cost on the integer ALU pipe, in LOP3
 * warp instructions (x 32 = lane operations): the bit-sliced kernel spends 5
 * per automaton row, text column and 1024 reads (minus 7 per unit and column
 * for the constant cells of rows 0-1).  `executed` counts the rows the kernels
 * really computed — the deep rows of a unit are skipped in the columns where
 * nothing can reach them, which the kernels tally per job at run time —,
 * `planned` every row of the scan plan in every column, `one_kmer_per_warp`
 * the same without any row sharing between k-mers.  Synchronises the stream
 * and resets the tally.  Zeros for the row-packed scan variants. */
typedef struct apc_scan_stats {
    uint64_t scans;
    double lop3_executed;
    double lop3_top;     /* part of `executed` spent in rows computed in every column */
    double lop3_planned;
    double lop3_one_kmer_per_warp;
} apc_scan_stats;
int apc_scan_stats_read(apc_ctx *ctx, apc_scan_stats *out);
/* Kernels launched by the most recent apc_scan (no synchronisation). */
uint64_t apc_last_scan_launches(const apc_ctx *ctx);

/* Tuning / test knobs.  "scan_variant": 0 or 7 = bit-sliced kernel (reads
 * packed into words, the default), 8 = best row-packed kernel for k, 1 = one
 * k-mer per 32-bit word, 2 = two per word (k<=16), 3 = three per word
 * (k<=10), 6 = three per 64-bit pair (k<=21).  "tiles_per_job": work per job
 * of the persistent warps (0 auto): 32-read tiles for the row-packed
 * kernels, 1024-read super-groups for the bit-sliced one.  "shape_mask": the
 * unit shapes the default kernel may group k-mers into (bit s = shape s of
 * apc_plan_queries; default all, 0 = one k-mer per warp) and "plan_alive_pct":
 * the share of text columns (percent) in which the planner expects a unit's
 * deep rows to be computed; both applied at the next apc_set_queries.
 * "scan_graph": 1 = a scan that is issued again unchanged is captured in a CUDA
 * graph and replayed from then on (one launch instead of up to 13 + fork/join
 * events), 0 (default) = always launch directly — back-to-back scans hide the
 * launches behind the previous scan anyway, and the graph's node dependencies
 * cost about what it saves (DESIGN.md tuning log).  "ingest_staging": 1
 * (default) = apc_ingest_fastx copies files of 64 MB and more through two
 * page-locked 16 MB pieces filled by the host threads, 0 = hands the caller's
 * buffer to cudaMemcpyAsync as it is. */
int apc_set_option(apc_ctx *ctx, const char *name, int64_t value);

/* The scan plan apc_set_queries would build for these k-mers (needs no GPU):
 * the default kernel scans k-mers that share a prefix — or, walking the text
 * backwards, a suffix — in units whose common rows are computed once.
 * order_out[n_kmers] = the k-mer indices in scan order (units of shape 0, 1,
 * ..., then the ungrouped k-mers), reversed_out[n_kmers] = 1 where the k-mer
 * at that position is scanned reversed, units_out[APC_PLAN_SHAPES] = units per
 * shape, shape_t_out / shape_g_out[APC_PLAN_SHAPES] = private bases per member
 * and members per unit of each shape for this k (0 where the shape does not
 * exist).  Any output pointer may be NULL.  Introspection for tests and
 * tuning; the counts do not depend on the plan. */
#define APC_PLAN_SHAPES 22
int apc_plan_queries(uint8_t k, const uint64_t *kmers, uint32_t n_kmers,
                     uint32_t *order_out, uint8_t *reversed_out,
                     uint32_t *units_out, int32_t *shape_t_out,
                     int32_t *shape_g_out);

/* Integer-pipe peak microbenchmark (roofline denominator):
 * runs dependent-free LOP3 / IMAD / mixed chains on every SM and returns
 * warp-level lane-ops per second for each. */
int apc_measure_int_peak(apc_ctx *ctx, double *lop3_ops_per_s,
                         double *imad_ops_per_s, double *mixed_ops_per_s);

/* Single microbenchmarks behind the figures above and the tuning notes in
 * "lop3", "imad", "mixed", "imad_hi", "imad_wide" return lane-ops
 * per second; "core_<sets>_<threads>_<ctas>" (see peak_kernels.cu) run
 * the scan kernel's column update from registers (no loads) and return
 * unit-columns per second.  "bs_stats" exists only in counting builds
 * (-DAPC_BS_STATS): share of the dead-row tests after which the deep rows
 * were computed since the last call. */
int apc_microbench(apc_ctx *ctx, const char *name, double *value);

#ifdef __cplusplus
}
#endif
#endif /* APC_H */
