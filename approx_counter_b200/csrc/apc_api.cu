// apc_api.cu — the C ABI of libapc (include/apc.h): context, sample upload,
// query upload, scan, result read-back, timing.  Plain pointers and sizes
// only; every CUDA failure becomes a status code plus apc_last_error text.
#include <algorithm>
#include <cstring>
#include <exception>
#include <new>

#include "apc_internal.h"

namespace apc {

int fail(Ctx *c, int status, const char *what, cudaError_t e) {
    if (c) {
        c->err = what ? what : "";
        if (e != cudaSuccess) {
            c->err += ": ";
            c->err += cudaGetErrorString(e);
        }
    }
    return status;
}

static int bind(Ctx *c) {
    if (!c) return APC_ERR_INVALID;
    APC_CUDA(c, cudaSetDevice(c->device));
    return APC_OK;
}

template <typename T>
static int grow(Ctx *c, T *&ptr, size_t &cap, size_t need_bytes) {
    if (need_bytes <= cap && ptr) return APC_OK;
    if (ptr) {
        APC_CUDA(c, cudaStreamSynchronize(c->stream));
        APC_CUDA(c, cudaFree(ptr));
        ptr = nullptr;
        cap = 0;
    }
    if (need_bytes == 0) need_bytes = 16;
    // 1/8 of slack: the `end` sample of a run is one base per read longer than the `start`
    // sample (:463) and must not force a re-allocation of every buffer
    need_bytes += need_bytes / 8;
    cudaError_t e = cudaMalloc((void **)&ptr, need_bytes);
    if (e != cudaSuccess) {
        ptr = nullptr;
        return fail(c, e == cudaErrorMemoryAllocation ? APC_ERR_NOMEM : APC_ERR_CUDA, "cudaMalloc", e);
    }
    cap = need_bytes;
    return APC_OK;
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventSynchronize(b) != cudaSuccess) return 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) return 0.f;
    return ms;
}

static int prepare_sample(Ctx *c, uint64_t n_reads, uint32_t max_len, uint64_t total_bases) {
    if (n_reads > 0xFFFFFFFFull * kTileReads) return fail(c, APC_ERR_INVALID, "too many reads");
    // the context holds no sample until every buffer of the new one exists (a failed call must not leave
    // new sizes next to old buffers)
    c->has_sample = false;
    c->plan_gen++;
    const uint32_t n_tiles = (uint32_t)((n_reads + kTileReads - 1) / kTileReads);
    const uint32_t chunks = (max_len + kChunkBases - 1) / kChunkBases;
    // + one chunk of padding: the scan kernel always prefetches the next 512 bytes
    const size_t tiles_bytes = ((size_t)n_tiles * chunks + 1) * kTileReads * sizeof(uint4);
    int st = grow(c, c->d_tiles, c->tiles_bytes, tiles_bytes);
    if (st) return st;
    // bit planes: 32 tiles per super-group, one uint4 per (column, tile), + two columns of padding at
    // either end for the kernels' prefetch (forward and backward walks)
    const size_t n_sg = ((size_t)n_tiles + 31) / 32;
    const size_t planes_bytes = n_sg * chunks * kChunkBases * 32 * sizeof(uint4);
    if ((st = grow(c, c->d_planes, c->planes_cap, planes_bytes + 2 * apc::kBsPadCols * 32 * sizeof(uint4)))) return st;
    const size_t lens_bytes = ((size_t)n_tiles * kTileReads + 1) * sizeof(uint32_t);
    if ((st = grow(c, c->d_lens, c->lens_cap, lens_bytes))) return st;
    APC_CUDA(c, cudaMemsetAsync(c->d_lens, 0, lens_bytes, c->stream));
    c->n_reads = n_reads;
    c->max_len = max_len;
    c->total_bases = total_bases;
    c->n_tiles = n_tiles;
    c->chunks = chunks;
    c->planes_bytes = planes_bytes;
    return APC_OK; // has_sample is set by the caller once the layout kernels are enqueued
}

// Newline index and record index of the n_eff bytes resident in d_file (tail padded with zeros up to a whole tile).
// ok = false: the lines do not form one-sequence-line records (n_nl is valid either way, d_nl holds the newlines).
static int ingest_index(Ctx *c, uint64_t n_eff, bool fastq, uint64_t &n_nl, uint64_t &n_rec, bool &ok) {
    int st;
    ok = false;
    n_rec = 0;
    const uint64_t n_tiles = (n_eff + kIngestTileBytes - 1) / kIngestTileBytes;
    if ((st = grow(c, c->d_tile_nl, c->tile_nl_cap, (size_t)(n_tiles + 1) * sizeof(uint64_t)))) return st;
    size_t temp_bytes = 0;
    APC_CUDA(c, ingest_prefix_u64(nullptr, temp_bytes, c->d_tile_nl, n_tiles + 1, c->stream));
    if ((st = grow(c, c->d_ingest_temp, c->ingest_temp_cap, temp_bytes))) return st;
    APC_CUDA(c, cudaMemsetAsync(c->d_tile_nl + n_tiles, 0, sizeof(uint64_t), c->stream));
    APC_CUDA(c, cudaMemsetAsync(c->d_ingest_flag, 0, 4 * sizeof(uint32_t), c->stream));
    APC_CUDA(c, launch_count_newlines(c->d_file, n_tiles, c->d_tile_nl, c->stream));
    temp_bytes = c->ingest_temp_cap;
    APC_CUDA(c, ingest_prefix_u64(c->d_ingest_temp, temp_bytes, c->d_tile_nl, n_tiles + 1, c->stream));
    APC_CUDA(c, cudaMemcpyAsync(&n_nl, c->d_tile_nl + n_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    if ((st = grow(c, c->d_nl, c->nl_cap, (size_t)std::max<uint64_t>(1, n_nl) * sizeof(uint64_t)))) return st;
    APC_CUDA(c, launch_write_newlines(c->d_file, n_tiles, c->d_tile_nl, c->d_nl, c->stream));
    const uint64_t n_lines = n_nl + 1, per = fastq ? 4 : 2;
    if (n_lines % per) return APC_OK; // not 4-line FASTQ / not single-line FASTA
    n_rec = n_lines / per;
    if ((st = grow(c, c->d_rec_start, c->rec_start_cap, (size_t)n_rec * sizeof(uint64_t)))) return st;
    if ((st = grow(c, c->d_rec_len, c->rec_len_cap, (size_t)n_rec * sizeof(uint32_t)))) return st;
    APC_CUDA(c, launch_index_records(c->d_file, c->d_nl, n_nl, n_eff, fastq, n_rec, c->d_rec_start, c->d_rec_len,
                                     c->d_ingest_flag, c->stream));
    uint32_t flag = 0;
    APC_CUDA(c, cudaMemcpyAsync(&flag, c->d_ingest_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    ok = flag == 0;
    return APC_OK;
}

// FASTA with any line structure -> single-line records in a second buffer (d_nl must hold the newlines of d_file).
static int ingest_unwrap(Ctx *c, uint64_t n_eff, uint64_t n_nl, uint64_t &n_eff2) {
    int st;
    const uint64_t n_lines = n_nl + 1;
    if ((st = grow(c, c->d_line_off, c->line_off_cap, (size_t)(n_lines + 1) * sizeof(uint64_t)))) return st;
    size_t temp_bytes = 0;
    APC_CUDA(c, ingest_prefix_u64(nullptr, temp_bytes, c->d_line_off, n_lines + 1, c->stream));
    if ((st = grow(c, c->d_ingest_temp, c->ingest_temp_cap, temp_bytes))) return st;
    APC_CUDA(c, launch_measure_lines(c->d_file, c->d_nl, n_nl, n_eff, c->d_line_off, c->stream));
    temp_bytes = c->ingest_temp_cap;
    APC_CUDA(c, ingest_prefix_u64(c->d_ingest_temp, temp_bytes, c->d_line_off, n_lines + 1, c->stream));
    APC_CUDA(c, cudaMemcpyAsync(&n_eff2, c->d_line_off + n_lines, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    const uint64_t padded = (n_eff2 + kIngestTileBytes - 1) / kIngestTileBytes * kIngestTileBytes;
    if ((st = grow(c, c->d_file2, c->file2_cap, (size_t)std::max<uint64_t>(padded, kIngestTileBytes)))) return st;
    APC_CUDA(c, launch_unwrap_lines(c->d_file, c->d_nl, n_nl, n_eff, c->d_line_off, c->d_file2, c->stream));
    if (padded > n_eff2) APC_CUDA(c, cudaMemsetAsync(c->d_file2 + n_eff2, 0, padded - n_eff2, c->stream));
    std::swap(c->d_file, c->d_file2);
    std::swap(c->file_cap, c->file2_cap);
    return APC_OK;
}

} // namespace apc

using apc::Ctx;

extern "C" {

int apc_version(void) { return APC_VERSION; }

const char *apc_strerror(int status) {
    switch (status) {
    case APC_OK: return "ok";
    case APC_ERR_INVALID: return "invalid argument";
    case APC_ERR_CUDA: return "CUDA runtime error";
    case APC_ERR_NO_DEVICE: return "no usable CUDA device";
    case APC_ERR_NO_SAMPLE: return "no sample uploaded";
    case APC_ERR_NO_QUERIES: return "no queries set";
    case APC_ERR_NOMEM: return "out of memory";
    case APC_ERR_CAPACITY: return "output capacity too small";
    case APC_ERR_COMM: return "communicator (NCCL) error";
    case APC_ERR_FORMAT: return "input outside the device parser's grammar (use the host parser)";
    default: return "unknown status";
    }
}

int apc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return APC_ERR_NO_DEVICE;
    }
    return n;
}

int apc_create(int device, apc_ctx **out) {
    if (!out) return APC_ERR_INVALID;
    *out = nullptr;
    int n = apc_device_count();
    if (n <= 0 || device < 0 || device >= n) return APC_ERR_NO_DEVICE;
    apc_ctx *c = new (std::nothrow) apc_ctx();
    if (!c) return APC_ERR_NOMEM;
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_table, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_job_counter, (apc::kBsShapes + 1) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->d_job_counter, 0, (apc::kBsShapes + 1) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_deep_lop3, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->d_deep_lop3, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->bs_fork, cudaEventDisableTiming);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev_ingest[i]);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_ingest_flag, 4 * sizeof(uint32_t));
    for (int i = 0; i < apc::kBsShapes && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->bs_streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->bs_join[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        apc_destroy(c);
        return APC_ERR_CUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return APC_OK;
}

void apc_destroy(apc_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &s : c->bs_streams) // a scan that failed half way may not have joined its side streams
        if (s) cudaStreamSynchronize(s);
    cudaFree(c->d_tiles);
    cudaFree(c->d_lens);
    cudaFree(c->d_planes);
    cudaFree(c->d_kmers);
    cudaFree(c->d_peq);
    cudaFree(c->d_counts);
    cudaFree(c->d_job_counter);
    cudaFree(c->d_deep_lop3);
    if (c->scan_graph) cudaGraphExecDestroy(c->scan_graph);
    apc_comm_destroy(c);
    cudaFree(c->d_stage);
    cudaFree(c->d_stage_offs);
    apc::free_exact_scratch(c);
    cudaFree(c->d_file);
    cudaFree(c->d_file2);
    cudaFree(c->d_line_off);
    cudaFree(c->d_tile_nl);
    cudaFree(c->d_nl);
    cudaFree(c->d_rec_start);
    cudaFree(c->d_rec_len);
    cudaFree(c->d_pick);
    cudaFree(c->d_ingest_temp);
    cudaFree(c->d_ingest_flag);
    for (int b = 0; b < 2; b++) {
        if (c->h_file_stage[b]) cudaFreeHost(c->h_file_stage[b]);
        if (c->ev_file_stage[b]) cudaEventDestroy(c->ev_file_stage[b]);
    }
    for (auto &e : c->ev_ingest)
        if (e) cudaEventDestroy(e);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->ev_table) cudaEventDestroy(c->ev_table);
    if (c->bs_fork) cudaEventDestroy(c->bs_fork);
    for (int i = 0; i < apc::kBsShapes; i++) {
        if (c->bs_join[i]) cudaEventDestroy(c->bs_join[i]);
        if (c->bs_streams[i]) cudaStreamDestroy(c->bs_streams[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char *apc_last_error(const apc_ctx *c) { return c ? c->err.c_str() : "null context"; }

int apc_set_stream(apc_ctx *c, void *cuda_stream) {
    if (!c) return APC_ERR_INVALID;
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return APC_OK;
}

int apc_sync(apc_ctx *c) {
    int st = apc::bind(c);
    if (st) return st;
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    return APC_OK;
}

int apc_upload_sample_async(apc_ctx *c, const uint8_t *bases, uint64_t n_reads, uint32_t read_len) {
    int st = apc::bind(c);
    if (st) return st;
    if (!bases && n_reads * read_len) return apc::fail(c, APC_ERR_INVALID, "bases is NULL");
    APC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    const size_t bytes = (size_t)n_reads * read_len;
    if ((st = apc::prepare_sample(c, n_reads, read_len, bytes))) return st;
    c->uniform_len = true;
    c->stage_is_sample = false;
    if ((st = apc::grow(c, c->d_stage, c->stage_cap, bytes))) return st;
    if (bytes) APC_CUDA(c, cudaMemcpyAsync(c->d_stage, bases, bytes, cudaMemcpyHostToDevice, c->stream));
    APC_CUDA(c, apc::launch_build_tiles_uniform(c->d_stage, n_reads, read_len, c->chunks, c->n_tiles,
                                                c->d_tiles, c->d_lens, c->stream));
    APC_CUDA(c, apc::launch_build_planes(*c));
    APC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    c->has_sample = true;
    c->stage_is_sample = true;
    c->timing.upload_ms = -1.f; // resolved lazily by apc_last_timing
    return APC_OK;
}

int apc_upload_sample(apc_ctx *c, const uint8_t *bases, uint64_t n_reads, uint32_t read_len) {
    int st = apc_upload_sample_async(c, bases, n_reads, read_len);
    if (st) return st;
    APC_CUDA(c, cudaStreamSynchronize(c->stream)); // caller may free `bases` on return
    c->timing.upload_ms = apc::elapsed(c->ev[0], c->ev[1]);
    return APC_OK;
}

// no C++ exception may cross the ABI: host containers that fail to allocate become APC_ERR_NOMEM
#define APC_TRY try {
#define APC_CATCH(ctx)                                                                   \
    }                                                                                    \
    catch (const std::bad_alloc &) { return apc::fail((ctx), APC_ERR_NOMEM, "host allocation failed"); } \
    catch (const std::exception &e) { return apc::fail((ctx), APC_ERR_INVALID, e.what()); }              \
    catch (...) { return apc::fail((ctx), APC_ERR_INVALID, "unknown C++ exception"); }

int apc_upload_sample_ragged(apc_ctx *c, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads) {
    APC_TRY
    int st = apc::bind(c);
    if (st) return st;
    if (!offsets && n_reads) return apc::fail(c, APC_ERR_INVALID, "offsets is NULL");
    uint64_t total = n_reads ? offsets[n_reads] - offsets[0] : 0;
    if (!bases && total) return apc::fail(c, APC_ERR_INVALID, "bases is NULL");
    uint32_t max_len = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        if (offsets[r + 1] < offsets[r]) return apc::fail(c, APC_ERR_INVALID, "offsets not monotone");
        const uint64_t len = offsets[r + 1] - offsets[r];
        if (len > 0xFFFFFFFFull) return apc::fail(c, APC_ERR_INVALID, "read too long");
        if (len > max_len) max_len = (uint32_t)len;
    }
    APC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if ((st = apc::prepare_sample(c, n_reads, max_len, total))) return st;
    c->uniform_len = false;
    c->stage_is_sample = false;
    if ((st = apc::grow(c, c->d_stage, c->stage_cap, (size_t)total))) return st;
    if ((st = apc::grow(c, c->d_stage_offs, c->stage_offs_cap, (size_t)(n_reads + 1) * sizeof(uint64_t)))) return st;
    if (n_reads) {
        const uint64_t base0 = offsets[0];
        if (total) APC_CUDA(c, cudaMemcpyAsync(c->d_stage, bases + base0, total, cudaMemcpyHostToDevice, c->stream));
        std::vector<uint64_t> rel(n_reads + 1);
        for (uint64_t r = 0; r <= n_reads; r++) rel[r] = offsets[r] - base0;
        APC_CUDA(c, cudaMemcpyAsync(c->d_stage_offs, rel.data(), (n_reads + 1) * sizeof(uint64_t),
                                    cudaMemcpyHostToDevice, c->stream));
        APC_CUDA(c, cudaStreamSynchronize(c->stream)); // rel goes out of scope
    }
    APC_CUDA(c, apc::launch_build_tiles_ragged(c->d_stage, c->d_stage_offs, n_reads, c->chunks, c->n_tiles,
                                               c->d_tiles, c->d_lens, c->stream));
    APC_CUDA(c, apc::launch_build_planes(*c));
    APC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->has_sample = true;
    c->timing.upload_ms = apc::elapsed(c->ev[0], c->ev[1]);
    return APC_OK;
    APC_CATCH(c)
}

// ---- ingest on the device (ingest_kernels.cu) ---------------------------------------------------------------
int apc_ingest_fastx(apc_ctx *c, const uint8_t *file_bytes, uint64_t n_bytes, uint64_t *n_records_out, int *is_fastq_out) {
    int st = apc::bind(c);
    if (st) return st;
    if (!file_bytes && n_bytes) return apc::fail(c, APC_ERR_INVALID, "file_bytes is NULL");
    c->has_file = false;
    c->n_records = 0;
    if (n_records_out) *n_records_out = 0;
    if (is_fastq_out) *is_fastq_out = 0;
    // blank lines at the end of the file are dropped here: the last line of what is copied is never empty
    uint64_t n_eff = n_bytes;
    while (n_eff > 0) {
        const uint8_t ch = file_bytes[n_eff - 1];
        if (ch != '\n' && ch != '\r' && ch != ' ' && ch != '\t') break;
        n_eff--;
    }
    // so are blank lines in front of the first record (the host parser skips them too)
    while (n_eff > 0 && (*file_bytes == '\n' || *file_bytes == '\r' || *file_bytes == ' ' || *file_bytes == '\t')) {
        file_bytes++;
        n_eff--;
    }
    c->file_bytes = n_eff;
    c->file_unwrapped = false;
    c->ingest_ms[0] = c->ingest_ms[1] = 0.f;
    if (n_eff == 0) { // no record at all
        c->has_file = true;
        c->file_fastq = 0;
        return APC_OK;
    }
    if (file_bytes[0] != '>' && file_bytes[0] != '@')
        return apc::fail(c, APC_ERR_FORMAT, "input does not start with '>' or '@'");
    const bool fastq = file_bytes[0] == '@';
    const uint64_t n_tiles = (n_eff + apc::kIngestTileBytes - 1) / apc::kIngestTileBytes;
    const uint64_t padded = n_tiles * apc::kIngestTileBytes;
    if ((st = apc::grow(c, c->d_file, c->file_cap, (size_t)padded))) return st;
    APC_CUDA(c, cudaEventRecord(c->ev_ingest[0], c->stream));
    if (c->opt_ingest_staging && n_eff >= 4 * apc::kIngestStageBytes) {
        // two page-locked pieces in flight: the host threads fill one (apch::parallel_copy) while the copy engine
        // drains the other — the driver's own staging of pageable memory is one thread deep (about 11 GB/s)
        for (int b = 0; b < 2; b++) {
            if (!c->h_file_stage[b]) {
                cudaError_t e = cudaHostAlloc((void **)&c->h_file_stage[b], apc::kIngestStageBytes, cudaHostAllocDefault);
                if (e != cudaSuccess) {
                    c->h_file_stage[b] = nullptr;
                    return apc::fail(c, APC_ERR_NOMEM, "cudaHostAlloc (file staging)", e);
                }
            }
            if (!c->ev_file_stage[b]) APC_CUDA(c, cudaEventCreateWithFlags(&c->ev_file_stage[b], cudaEventDisableTiming));
        }
        uint64_t piece = 0;
        for (uint64_t off = 0; off < n_eff; off += apc::kIngestStageBytes, piece++) {
            const uint64_t len = std::min<uint64_t>(apc::kIngestStageBytes, n_eff - off);
            const int b = (int)(piece & 1);
            if (piece >= 2) APC_CUDA(c, cudaEventSynchronize(c->ev_file_stage[b]));
            apch::parallel_copy(c->h_file_stage[b], file_bytes + off, (size_t)len);
            APC_CUDA(c, cudaMemcpyAsync(c->d_file + off, c->h_file_stage[b], len, cudaMemcpyHostToDevice, c->stream));
            APC_CUDA(c, cudaEventRecord(c->ev_file_stage[b], c->stream));
        }
    } else {
        // small files: pageable memory, staged by the driver
        for (uint64_t off = 0; off < n_eff; off += (uint64_t)64 << 20) {
            const uint64_t len = std::min<uint64_t>((uint64_t)64 << 20, n_eff - off);
            APC_CUDA(c, cudaMemcpyAsync(c->d_file + off, file_bytes + off, len, cudaMemcpyHostToDevice, c->stream));
        }
    }
    if (padded > n_eff) APC_CUDA(c, cudaMemsetAsync(c->d_file + n_eff, 0, padded - n_eff, c->stream));
    APC_CUDA(c, cudaEventRecord(c->ev_ingest[1], c->stream));
    uint64_t n_nl = 0, n_rec = 0;
    bool ok = false;
    if ((st = apc::ingest_index(c, n_eff, fastq, n_nl, n_rec, ok))) return st;
    if (!ok && !fastq) {
        // not one sequence line per record: wrapped FASTA (or blank lines, or a header without a sequence line).  The
        // records are re-laid as ">\n<sequence on one line>\n" in a second buffer, which then takes the file's place.
        uint64_t n_eff2 = 0;
        if ((st = apc::ingest_unwrap(c, n_eff, n_nl, n_eff2))) return st;
        n_eff = n_eff2;
        c->file_bytes = n_eff;
        if ((st = apc::ingest_index(c, n_eff, false, n_nl, n_rec, ok))) return st;
        c->file_unwrapped = true;
    }
    APC_CUDA(c, cudaEventRecord(c->ev_ingest[2], c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->ingest_ms[0] = apc::elapsed(c->ev_ingest[0], c->ev_ingest[1]);
    c->ingest_ms[1] = apc::elapsed(c->ev_ingest[1], c->ev_ingest[2]);
    if (!ok)
        return apc::fail(c, APC_ERR_FORMAT, fastq ? "not 4-line FASTQ (wrapped record, blank line, or quality length)"
                                                 : "not FASTA the device parser takes (blanks inside a sequence line)");
    if (n_rec > 0x7FFFFFFFull) return apc::fail(c, APC_ERR_INVALID, "too many records (the reference's read ids are int)");
    c->has_file = true;
    c->file_fastq = fastq ? 1 : 0;
    c->n_records = n_rec;
    if (n_records_out) *n_records_out = n_rec;
    if (is_fastq_out) *is_fastq_out = c->file_fastq;
    return APC_OK;
}

int apc_ingest_lengths(apc_ctx *c, uint64_t first, uint64_t n, uint32_t *lens_out) {
    int st = apc::bind(c);
    if (st) return st;
    if (!c->has_file) return apc::fail(c, APC_ERR_NO_SAMPLE, "apc_ingest_fastx first");
    if (first > c->n_records || n > c->n_records - first || (!lens_out && n)) return apc::fail(c, APC_ERR_INVALID, "range");
    if (n) APC_CUDA(c, cudaMemcpyAsync(lens_out, c->d_rec_len + first, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    return APC_OK;
}

int apc_sample_resident(apc_ctx *c, const uint32_t *order, uint64_t n_order, uint64_t nb_sample, uint32_t cut, int bot,
                        uint64_t *n_sampled_out) {
    int st = apc::bind(c);
    if (st) return st;
    if (n_sampled_out) *n_sampled_out = 0;
    if (!c->has_file) return apc::fail(c, APC_ERR_NO_SAMPLE, "apc_ingest_fastx first");
    const uint64_t n = c->n_records;
    if (order && n_order != n) return apc::fail(c, APC_ERR_INVALID, "order must hold every record id once");
    if (cut > 0x7FFFFFF0u) return apc::fail(c, APC_ERR_INVALID, "cut too large");
    const uint32_t row_len = cut + (bot ? 1u : 0u);
    APC_CUDA(c, cudaEventRecord(c->ev_ingest[2], c->stream));
    uint64_t n_sampled = 0;
    uint32_t *d_ord = nullptr, *d_flags = nullptr, *d_pos = nullptr;
    uint64_t *d_src_off = nullptr;
    if (n && cut && nb_sample) { // :461 takes nothing when cut == 0 (current_cut_size > 0 is part of the test)
        const uint64_t n4 = (n + 3) & ~(uint64_t)3; // keeps the u64 part aligned
        if ((st = apc::grow(c, c->d_pick, c->pick_cap, (size_t)n4 * 3 * sizeof(uint32_t) + (size_t)n * sizeof(uint64_t)))) return st;
        d_ord = c->d_pick, d_flags = d_ord + n4, d_pos = d_flags + n4;
        d_src_off = reinterpret_cast<uint64_t *>(d_pos + n4);
        size_t temp_bytes = 0;
        APC_CUDA(c, apc::ingest_prefix_u32(nullptr, temp_bytes, d_flags, d_pos, n, c->stream));
        if ((st = apc::grow(c, c->d_ingest_temp, c->ingest_temp_cap, temp_bytes))) return st;
        if (order) APC_CUDA(c, cudaMemcpyAsync(d_ord, order, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        APC_CUDA(c, cudaMemsetAsync(c->d_ingest_flag, 0, 4 * sizeof(uint32_t), c->stream));
        APC_CUDA(c, apc::launch_pick_reads(order ? d_ord : nullptr, n, c->d_rec_start, c->d_rec_len, cut, bot != 0, nb_sample,
                                           d_flags, d_pos, d_src_off, c->d_ingest_temp, c->ingest_temp_cap, c->d_ingest_flag,
                                           c->stream));
        uint32_t last[2] = {0, 0}, flag = 0;
        APC_CUDA(c, cudaMemcpyAsync(&last[0], d_pos + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        APC_CUDA(c, cudaMemcpyAsync(&last[1], d_flags + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        APC_CUDA(c, cudaMemcpyAsync(&flag, c->d_ingest_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        APC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (flag) return apc::fail(c, APC_ERR_INVALID, "order holds an id outside [0, n_records)");
        n_sampled = std::min<uint64_t>(nb_sample, (uint64_t)last[0] + last[1]);
    }
    const size_t bytes = (size_t)n_sampled * row_len;
    if ((st = apc::prepare_sample(c, n_sampled, row_len, bytes))) return st;
    c->uniform_len = true;
    c->stage_is_sample = false;
    if ((st = apc::grow(c, c->d_stage, c->stage_cap, bytes))) return st;
    APC_CUDA(c, apc::launch_gather_ends(c->d_file, d_src_off, n_sampled, row_len, c->d_stage, c->stream));
    APC_CUDA(c, apc::launch_build_tiles_uniform(c->d_stage, n_sampled, row_len, c->chunks, c->n_tiles, c->d_tiles, c->d_lens,
                                                c->stream));
    APC_CUDA(c, apc::launch_build_planes(*c));
    APC_CUDA(c, cudaEventRecord(c->ev_ingest[3], c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->has_sample = true;
    c->stage_is_sample = true;
    c->ingest_ms[2] = apc::elapsed(c->ev_ingest[2], c->ev_ingest[3]);
    c->timing.upload_ms = c->ingest_ms[2];
    if (n_sampled_out) *n_sampled_out = n_sampled;
    return APC_OK;
}

int apc_upload_sample_peer(apc_ctx *c, const apc_ctx *src, uint64_t first_read, uint64_t n_reads) {
    int st = apc::bind(c);
    if (st) return st;
    if (!src || src == c) return apc::fail(c, APC_ERR_INVALID, "src must be another context");
    if (!src->has_sample || !src->uniform_len || !src->stage_is_sample)
        return apc::fail(c, APC_ERR_NO_SAMPLE, "src holds no sample with ASCII rows");
    if (first_read > src->n_reads || n_reads > src->n_reads - first_read) return apc::fail(c, APC_ERR_INVALID, "rows outside src's sample");
    const uint32_t read_len = src->max_len;
    if (src->device != c->device) { // direct GPU-to-GPU copies (NVLink) where the devices are peers; staged by the driver otherwise
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c->device, src->device) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return apc::fail(c, APC_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
            (void)cudaGetLastError();
        }
    }
    APC_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    const size_t bytes = (size_t)n_reads * read_len;
    if ((st = apc::prepare_sample(c, n_reads, read_len, bytes))) return st;
    c->uniform_len = true;
    c->stage_is_sample = false;
    if ((st = apc::grow(c, c->d_stage, c->stage_cap, bytes))) return st;
    if (bytes)
        APC_CUDA(c, cudaMemcpyPeerAsync(c->d_stage, c->device, src->d_stage + (size_t)first_read * read_len, src->device, bytes,
                                        c->stream));
    APC_CUDA(c, apc::launch_build_tiles_uniform(c->d_stage, n_reads, read_len, c->chunks, c->n_tiles, c->d_tiles, c->d_lens,
                                                c->stream));
    APC_CUDA(c, apc::launch_build_planes(*c));
    APC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    c->has_sample = true;
    c->stage_is_sample = true;
    c->timing.upload_ms = -1.f; // resolved lazily by apc_last_timing
    return APC_OK;
}

int apc_download_sample(apc_ctx *c, uint8_t *bases_out, uint64_t capacity) {
    int st = apc::bind(c);
    if (st) return st;
    if (!c->has_sample) return apc::fail(c, APC_ERR_NO_SAMPLE, "no sample");
    if (!c->uniform_len || !c->stage_is_sample) return apc::fail(c, APC_ERR_INVALID, "the resident sample has no ASCII rows");
    const uint64_t bytes = c->n_reads * c->max_len;
    if (bytes > capacity) return apc::fail(c, APC_ERR_CAPACITY, "capacity");
    if (bytes && !bases_out) return apc::fail(c, APC_ERR_INVALID, "bases_out is NULL");
    if (bytes) APC_CUDA(c, cudaMemcpyAsync(bases_out, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    return APC_OK;
}

int apc_ingest_timing(const apc_ctx *c, float *copy_ms, float *index_ms, float *sample_ms) {
    if (!c) return APC_ERR_INVALID;
    if (copy_ms) *copy_ms = c->ingest_ms[0];
    if (index_ms) *index_ms = c->ingest_ms[1];
    if (sample_ms) *sample_ms = c->ingest_ms[2];
    return APC_OK;
}

int apc_reserve(apc_ctx *c, uint64_t n_reads, uint32_t read_len, uint8_t k, uint32_t n_kmers) {
    APC_TRY
    int st = apc::bind(c);
    if (st) return st;
    if (k < 2 || k > 32) return apc::fail(c, APC_ERR_INVALID, "k must be in [2,32]");
    // the buffers of the expected sample and of the exact stage's k-windows
    if ((st = apc::prepare_sample(c, n_reads, read_len, n_reads * read_len))) return st;
    if ((st = apc::grow(c, c->d_stage, c->stage_cap, (size_t)n_reads * read_len))) return st;
    const uint64_t windows = read_len >= k ? n_reads * (uint64_t)(read_len - k + 1) : 0;
    if ((st = apc::exact_reserve(c, k, windows))) return st;
    // the whole path once on a 64-read dummy sample: loads the layout kernels, the exact stage (the library sort
    // included) and the scan kernels the dummy's plan uses; then every other scan kernel of this k
    {
        const uint32_t len = std::max<uint32_t>(read_len, (uint32_t)k + 8u);
        std::vector<uint8_t> dummy((size_t)64 * len);
        uint32_t x = 12345u;
        for (auto &b : dummy) {
            x = x * 1664525u + 1013904223u;
            b = (uint8_t)"ACGT"[x >> 30];
        }
        for (uint32_t r = 0; r < 64; r += 2)
            std::memcpy(&dummy[(size_t)r * len], &dummy[0], k + 4u); // a shared prefix: something to group
        if ((st = apc_upload_sample(c, dummy.data(), 64, len))) return st;
        const uint32_t lim = std::max<uint32_t>(16u, std::min<uint32_t>(n_kmers, 64u));
        std::vector<uint64_t> km(lim), ct(lim);
        uint64_t n_top = 0;
        if ((st = apc_exact_topn(c, k, 1e30f, lim, nullptr, 0, km.data(), ct.data(), &n_top, nullptr, nullptr))) return st;
        if (n_top && (st = apc_approx_count(c, k, km.data(), (uint32_t)n_top, ct.data()))) return st;
    }
    APC_CUDA(c, apc::warm_bs_kernels(*c, k));
    if (n_kmers) { // the query tables and the count vector
        const size_t words = (size_t)n_kmers * 3;
        if ((st = apc::grow(c, c->d_kmers, c->kmers_cap, words * sizeof(uint32_t)))) return st;
        if ((st = apc::grow(c, c->d_counts, c->counts_cap, (size_t)n_kmers * sizeof(unsigned long long)))) return st;
    }
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->has_sample = false; // the dummy is not a sample anyone asked for
    c->k = 0;
    c->plan_gen++;
    c->timing = apc_timing{};
    c->stat_scans = 0;
    c->stat_lop3_top = c->stat_lop3_all = c->stat_lop3_single = 0.;
    APC_CUDA(c, cudaMemsetAsync(c->d_deep_lop3, 0, sizeof(unsigned long long), c->stream));
    return APC_OK;
    APC_CATCH(c)
}

int apc_sample_info(const apc_ctx *c, uint64_t *n_reads, uint32_t *max_len, uint64_t *total_bases) {
    if (!c) return APC_ERR_INVALID;
    if (n_reads) *n_reads = c->n_reads;
    if (max_len) *max_len = c->max_len;
    if (total_bases) *total_bases = c->total_bases;
    return APC_OK;
}

static int exact_common(apc_ctx *c, uint8_t k, float lc, uint64_t lim, uint64_t solid_km,
                        const uint64_t *forbidden, uint64_t n_forbidden, uint64_t *kmers_out,
                        uint64_t *counts_out, uint64_t capacity, uint64_t *n_out, uint64_t *n_distinct,
                        uint64_t *n_had_n) {
    APC_TRY
    int st = apc::bind(c);
    if (st) return st;
    if (k < 2 || k > 32) return apc::fail(c, APC_ERR_INVALID, "k must be in [2,32]");
    if (!c->has_sample) return apc::fail(c, APC_ERR_NO_SAMPLE, "exact stage: no sample uploaded");
    if (!n_out || (!forbidden && n_forbidden)) return apc::fail(c, APC_ERR_INVALID, "NULL argument");
    std::vector<uint64_t> km, ct;
    uint64_t needed = 0;
    st = apc::exact_count_select(c, k, lc, lim, solid_km, forbidden, n_forbidden, capacity, km, ct, &needed,
                                 n_distinct, n_had_n);
    if (st == APC_ERR_CAPACITY) {
        *n_out = needed;
        return apc::fail(c, APC_ERR_CAPACITY, "output capacity too small");
    }
    if (st) return st;
    *n_out = km.size();
    if (km.size() > capacity) return apc::fail(c, APC_ERR_CAPACITY, "output capacity too small");
    if (!km.empty() && (!kmers_out || !counts_out)) return apc::fail(c, APC_ERR_INVALID, "NULL output");
    for (size_t i = 0; i < km.size(); i++) { kmers_out[i] = km[i]; counts_out[i] = ct[i]; }
    return APC_OK;
    APC_CATCH(c)
}

int apc_exact_topn(apc_ctx *c, uint8_t k, float lc_adjusted, uint64_t lim, const uint64_t *forbidden,
                   uint64_t n_forbidden, uint64_t *kmers_out, uint64_t *counts_out, uint64_t *n_out,
                   uint64_t *n_distinct, uint64_t *n_had_n) {
    return exact_common(c, k, lc_adjusted, lim, 0, forbidden, n_forbidden, kmers_out, counts_out, lim, n_out,
                        n_distinct, n_had_n);
}

int apc_exact_solid(apc_ctx *c, uint8_t k, float lc_adjusted, uint64_t solid_km, const uint64_t *forbidden,
                    uint64_t n_forbidden, uint64_t *kmers_out, uint64_t *counts_out, uint64_t capacity,
                    uint64_t *n_out, uint64_t *n_distinct, uint64_t *n_had_n) {
    if (solid_km == 0) return apc::fail(c, APC_ERR_INVALID, "solid_km must be > 0");
    return exact_common(c, k, lc_adjusted, ~0ull, solid_km, forbidden, n_forbidden, kmers_out, counts_out,
                        capacity, n_out, n_distinct, n_had_n);
}

int apc_set_queries(apc_ctx *c, uint8_t k, const uint64_t *kmers, uint32_t n_kmers) {
    APC_TRY
    int st = apc::bind(c);
    if (st) return st;
    if (k < 2 || k > 32) return apc::fail(c, APC_ERR_INVALID, "k must be in [2,32]");
    if (!kmers && n_kmers) return apc::fail(c, APC_ERR_INVALID, "kmers is NULL");
    if (n_kmers > 0x7FFFFFFFu) return apc::fail(c, APC_ERR_INVALID, "too many k-mers");
    if (k < 32)
        for (uint32_t i = 0; i < n_kmers; i++)
            if (kmers[i] >> (2 * k)) return apc::fail(c, APC_ERR_INVALID, "k-mer value wider than 2k bits");
    // Everything is staged in locals and published to the context only after every allocation has
    // succeeded; until then the context has no queries (k = 0), so a failed call cannot leave new sizes
    // next to old buffers.
    c->k = 0;
    c->plan_gen++;
    const apc::ScanVariant variant = apc::pick_variant(k, c->opt_variant);
    // the match tables are built in a context-owned pinned buffer so that the H2D copy is
    // truly asynchronous; the buffer is only rewritten once its previous copy has completed
    const uint32_t qg = variant.queries_per_group();
    const bool bs = variant.bitslice();
    // row-packed kernels take match tables, the bit-sliced kernel the k-mers themselves
    const size_t table_words = bs ? (size_t)n_kmers * 3 : (size_t)((n_kmers + qg - 1) / qg) * apc::kPeqRows * apc::kWordsPerThread;
    if (c->table_copy_pending) {
        APC_CUDA(c, cudaEventSynchronize(c->ev_table));
        c->table_copy_pending = false;
    }
    if (table_words * sizeof(uint32_t) > c->pinned_cap) {
        if (c->h_pinned) APC_CUDA(c, cudaFreeHost(c->h_pinned));
        c->h_pinned = nullptr;
        c->pinned_cap = 0;
        const size_t cap = std::max<size_t>(table_words * sizeof(uint32_t) * 2, 1 << 16);
        cudaError_t e = cudaMallocHost(&c->h_pinned, cap);
        if (e != cudaSuccess) return apc::fail(c, APC_ERR_NOMEM, "cudaMallocHost", e);
        c->pinned_cap = cap;
    }
    void *d_dst = nullptr;
    uint32_t n_groups = 0;
    uint32_t units[apc::kBsShapes] = {};
    if (bs) {
        // k-mers in scan order (units of prefix- or suffix-sharing k-mers first, shape by shape), then the
        // index of each in the caller's order; bit 31 marks the members of a unit that is scanned backwards
        std::vector<uint32_t> order;
        std::vector<uint8_t> reversed;
        uint32_t shape_mask = variant.pairing() ? c->opt_shape_mask : 0u;
        // A small scan is bound by parallelism, not by throughput: a job is one unit x 1024 reads walked column by
        // column, and with few jobs per resident warp the scan takes as long as its longest jobs.  So the largest
        // shapes are left out while there are fewer than 8 jobs per resident warp (units of at most 48 rows: 10 000
        // reads x 5 000 k-mers 0.54 -> 0.49 ms per step), and below 2 jobs per warp only the shapes of at most 31 rows
        // are used (tools/c1_probe.py on C1, 10 000 reads x 500 k-mers: 0.156 -> 0.105 ms per step).
        if (c->has_sample && shape_mask == 0xFFFFFFFFu) {
            const uint64_t n_sg = ((uint64_t)c->n_tiles + 31) / 32, jobs = n_sg * n_kmers / 6;
            const uint64_t warps = (uint64_t)c->sm_count * 8;
            if (jobs < 2 * warps) shape_mask = apc::kBsSmallShapes;
            else if (jobs < 8 * warps) shape_mask = apc::kBsMediumShapes;
        }
        apc::bs_group_queries(kmers, n_kmers, k, shape_mask, (float)c->opt_alive_pct / 100.f, order, reversed, units);
        uint64_t *hk = (uint64_t *)c->h_pinned;
        uint32_t *hp = (uint32_t *)(hk + n_kmers);
        for (uint32_t i = 0; i < n_kmers; i++) {
            hk[i] = reversed[i] ? apc::bs_reverse_kmer(kmers[order[i]], k) : kmers[order[i]];
            hp[i] = order[i] | (reversed[i] ? 0x80000000u : 0u);
        }
        n_groups = n_kmers;
        if ((st = apc::grow(c, c->d_kmers, c->kmers_cap, table_words * sizeof(uint32_t)))) return st;
        d_dst = c->d_kmers;
    } else {
        apc::build_peq_tables(kmers, n_kmers, k, variant, (uint32_t *)c->h_pinned, n_groups);
        if ((st = apc::grow(c, c->d_peq, c->peq_cap, table_words * sizeof(uint32_t)))) return st;
        d_dst = c->d_peq;
    }
    const size_t slots = (size_t)n_groups * qg;
    if ((st = apc::grow(c, c->d_counts, c->counts_cap, slots * sizeof(unsigned long long)))) return st;
    if (table_words) {
        APC_CUDA(c, cudaMemcpyAsync(d_dst, c->h_pinned, table_words * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                    c->stream));
        APC_CUDA(c, cudaEventRecord(c->ev_table, c->stream));
        c->table_copy_pending = true;
    }
    c->n_kmers = n_kmers;
    c->variant = variant;
    c->n_groups = n_groups;
    for (int i = 0; i < apc::kBsShapes; i++) c->bs_units[i] = units[i];
    c->k = k;
    return APC_OK;
    APC_CATCH(c)
}

int apc_plan_queries(uint8_t k, const uint64_t *kmers, uint32_t n_kmers, uint32_t *order_out, uint8_t *reversed_out,
                     uint32_t *units_out, int32_t *shape_t_out, int32_t *shape_g_out) {
    static_assert(APC_PLAN_SHAPES == apc::kBsShapes, "apc.h and apc_internal.h disagree on the number of shapes");
    if (k < 2 || k > 32 || (!kmers && n_kmers) || n_kmers > 0x7FFFFFFFu) return APC_ERR_INVALID;
    try {
        std::vector<uint32_t> order;
        std::vector<uint8_t> reversed;
        uint32_t units[apc::kBsShapes];
        apc::bs_group_queries(kmers, n_kmers, k, 0xFFFFFFFFu, (float)apc::kBsAlivePct / 100.f, order, reversed, units);
        for (uint32_t i = 0; i < n_kmers; i++) {
            if (order_out) order_out[i] = order[i];
            if (reversed_out) reversed_out[i] = reversed[i];
        }
        for (int s = 0; s < apc::kBsShapes; s++) {
            const apc::BsShape sh = apc::bs_shape(k, s);
            if (units_out) units_out[s] = units[s];
            if (shape_t_out) shape_t_out[s] = sh.g ? sh.t : 0;
            if (shape_g_out) shape_g_out[s] = sh.g;
        }
    } catch (...) {
        return APC_ERR_NOMEM;
    }
    return APC_OK;
}

// The launches of one scan, directly or as a replay of their CUDA graph.  A scan that is issued a second time
// unchanged (same sample, queries, options, destination and stream) is captured while it is launched —
// stream capture follows the fork/join events onto the side streams — and replayed from then on: one
// cudaGraphLaunch instead of up to 13 kernel launches and 24 event calls, which is most of a C1-sized scan.
static int scan_launches(apc_ctx *c, unsigned long long *dst) {
    const bool same = c->opt_graph && c->variant.bitslice() && c->last_scan_gen == c->plan_gen &&
                      c->last_scan_dst == dst && c->last_scan_stream == c->stream;
    c->last_scan_gen = c->plan_gen;
    c->last_scan_dst = dst;
    c->last_scan_stream = c->stream;
    if (same && c->scan_graph && c->graph_gen == c->plan_gen && c->graph_dst == dst && c->graph_stream == c->stream) {
        APC_CUDA(c, cudaGraphLaunch(c->scan_graph, c->stream));
        c->timing.scan_launches = c->graph_launches;
        // the plan's share of the statistics (the kernels tally the data-dependent part themselves)
        c->stat_scans++;
        c->stat_lop3_top += c->graph_lop3[0];
        c->stat_lop3_all += c->graph_lop3[1];
        c->stat_lop3_single += c->graph_lop3[2];
        return APC_OK;
    }
    if (!same) {
        APC_CUDA(c, apc::launch_scan(*c, dst, &c->timing.scan_launches));
        return APC_OK;
    }
    // second identical scan: capture it (nothing runs during capture), then launch the graph
    if (c->scan_graph) {
        cudaGraphExecDestroy(c->scan_graph);
        c->scan_graph = nullptr;
        c->graph_gen = ~0ull;
    }
    const double before[3] = {c->stat_lop3_top, c->stat_lop3_all, c->stat_lop3_single};
    const uint64_t scans_before = c->stat_scans;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { // the stream cannot be captured (e.g. the legacy default stream): launch directly
        cudaGetLastError();
        c->opt_graph = 0;
        APC_CUDA(c, apc::launch_scan(*c, dst, &c->timing.scan_launches));
        return APC_OK;
    }
    const cudaError_t le = apc::launch_scan(*c, dst, &c->timing.scan_launches);
    e = cudaStreamEndCapture(c->stream, &graph);
    if (le != cudaSuccess || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->opt_graph = 0; // do not try again on this context
        c->stat_scans = scans_before;
        c->stat_lop3_top = before[0]; c->stat_lop3_all = before[1]; c->stat_lop3_single = before[2];
        APC_CUDA(c, apc::launch_scan(*c, dst, &c->timing.scan_launches));
        return APC_OK;
    }
    e = cudaGraphInstantiate(&c->scan_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        c->scan_graph = nullptr;
        cudaGetLastError();
        c->opt_graph = 0;
        c->stat_scans = scans_before;
        c->stat_lop3_top = before[0]; c->stat_lop3_all = before[1]; c->stat_lop3_single = before[2];
        APC_CUDA(c, apc::launch_scan(*c, dst, &c->timing.scan_launches));
        return APC_OK;
    }
    c->graph_gen = c->plan_gen;
    c->graph_dst = dst;
    c->graph_stream = c->stream;
    c->graph_launches = c->timing.scan_launches;
    c->graph_lop3[0] = c->stat_lop3_top - before[0];
    c->graph_lop3[1] = c->stat_lop3_all - before[1];
    c->graph_lop3[2] = c->stat_lop3_single - before[2];
    APC_CUDA(c, cudaGraphLaunch(c->scan_graph, c->stream));
    return APC_OK;
}

int apc_scan(apc_ctx *c, uint64_t *d_counts) {
    int st = apc::bind(c);
    if (st) return st;
    if (!c->has_sample) return apc::fail(c, APC_ERR_NO_SAMPLE, "apc_scan: no sample uploaded");
    if (c->k == 0) return apc::fail(c, APC_ERR_NO_QUERIES, "apc_scan: no queries");
    unsigned long long *dst = d_counts ? (unsigned long long *)d_counts : c->d_counts;
    APC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((st = scan_launches(c, dst))) return st;
    APC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    c->timing.scan_ms = -1.f; // resolved lazily by apc_last_timing / apc_get_counts
    return APC_OK;
}

int apc_scan_allreduce(apc_ctx *c, uint64_t *d_counts) {
    int st = apc_scan(c, d_counts);
    if (st) return st;
    return apc_allreduce_counts(c, d_counts, c->n_kmers);
}

int apc_scan_stats_read(apc_ctx *c, apc_scan_stats *out) {
    int st = apc::bind(c);
    if (st) return st;
    if (!out) return apc::fail(c, APC_ERR_INVALID, "NULL argument");
    unsigned long long deep = 0;
    APC_CUDA(c, cudaMemcpyAsync(&deep, c->d_deep_lop3, sizeof deep, cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaMemsetAsync(c->d_deep_lop3, 0, sizeof deep, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    out->scans = c->stat_scans;
    out->lop3_top = c->stat_lop3_top;
    out->lop3_executed = c->stat_lop3_top + (double)deep;
    out->lop3_planned = c->stat_lop3_all;
    out->lop3_one_kmer_per_warp = c->stat_lop3_single;
    c->stat_scans = 0;
    c->stat_lop3_top = c->stat_lop3_all = c->stat_lop3_single = 0.;
    return APC_OK;
}

int apc_get_counts(apc_ctx *c, uint64_t *counts_out) {
    int st = apc::bind(c);
    if (st) return st;
    if (c->k == 0) return apc::fail(c, APC_ERR_NO_QUERIES, "apc_get_counts: no queries");
    if (!counts_out && c->n_kmers) return apc::fail(c, APC_ERR_INVALID, "counts_out is NULL");
    if (c->n_kmers)
        APC_CUDA(c, cudaMemcpyAsync(counts_out, c->d_counts, (size_t)c->n_kmers * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    return APC_OK;
}

uint64_t *apc_counts_device_ptr(apc_ctx *c) { return c ? (uint64_t *)c->d_counts : nullptr; }

int apc_approx_count_async(apc_ctx *c, uint8_t k, const uint64_t *kmers, uint32_t n_kmers, uint64_t *counts_out) {
    int st = apc::bind(c);
    if (st) return st;
    if (!c->has_sample) return apc::fail(c, APC_ERR_NO_SAMPLE, "apc_approx_count: no sample uploaded");
    if (!counts_out && n_kmers) return apc::fail(c, APC_ERR_INVALID, "counts_out is NULL");
    APC_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    if ((st = apc_set_queries(c, k, kmers, n_kmers))) return st;
    if ((st = apc_scan(c, nullptr))) return st;
    if (n_kmers)
        APC_CUDA(c, cudaMemcpyAsync(counts_out, c->d_counts, (size_t)n_kmers * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, c->stream));
    APC_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    c->timing.total_ms = -1.f;
    return APC_OK;
}

int apc_approx_count(apc_ctx *c, uint8_t k, const uint64_t *kmers, uint32_t n_kmers, uint64_t *counts_out) {
    int st = apc_approx_count_async(c, k, kmers, n_kmers, counts_out);
    if (st) return st;
    APC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->timing.total_ms = apc::elapsed(c->ev[4], c->ev[5]);
    return APC_OK;
}

uint64_t apc_last_scan_launches(const apc_ctx *c) { return c ? c->timing.scan_launches : 0; }

int apc_last_timing(const apc_ctx *cc, apc_timing *out) {
    apc_ctx *c = const_cast<apc_ctx *>(cc);
    if (!c || !out) return APC_ERR_INVALID;
    if (c->timing.scan_ms < 0.f) c->timing.scan_ms = apc::elapsed(c->ev[2], c->ev[3]);
    if (c->timing.upload_ms < 0.f) c->timing.upload_ms = apc::elapsed(c->ev[0], c->ev[1]);
    if (c->timing.total_ms < 0.f) c->timing.total_ms = apc::elapsed(c->ev[4], c->ev[5]);
    *out = c->timing;
    return APC_OK;
}

int apc_set_option(apc_ctx *c, const char *name, int64_t value) {
    if (!c || !name) return APC_ERR_INVALID;
    c->plan_gen++; // any option may change what a scan launches
    if (!std::strcmp(name, "scan_graph")) {
        c->opt_graph = value != 0;
        return APC_OK;
    }
    if (!std::strcmp(name, "ingest_staging")) {
        c->opt_ingest_staging = value != 0;
        return APC_OK;
    }
    if (!std::strcmp(name, "scan_variant")) {
        if (value != 0 && value != 1 && value != 2 && value != 3 && value != 6 && value != 7 && value != 8)
            return apc::fail(c, APC_ERR_INVALID, "scan_variant must be 0,1,2,3,6,7 or 8");
        c->opt_variant = (int)value;
        return APC_OK;
    }
    if (!std::strcmp(name, "tiles_per_job")) {
        if (value < 0 || value > (1 << 20)) return apc::fail(c, APC_ERR_INVALID, "tiles_per_job out of range");
        c->opt_tiles_per_job = (int)value;
        return APC_OK;
    }
    if (!std::strcmp(name, "shape_mask")) { // takes effect at the next apc_set_queries
        if (value < 0 || value > 0xFFFFFFFFll) return apc::fail(c, APC_ERR_INVALID, "shape_mask out of range");
        c->opt_shape_mask = (uint32_t)value;
        return APC_OK;
    }
    if (!std::strcmp(name, "plan_alive_pct")) { // takes effect at the next apc_set_queries
        if (value < 0 || value > 100) return apc::fail(c, APC_ERR_INVALID, "plan_alive_pct must be in [0,100]");
        c->opt_alive_pct = (int)value;
        return APC_OK;
    }
    if (!std::strcmp(name, "scan_first_read")) {
        if (value < 0 || value % apc::kTileReads) return apc::fail(c, APC_ERR_INVALID, "scan_first_read must be a multiple of 32");
        c->opt_first_read = (uint64_t)value;
        return APC_OK;
    }
    if (!std::strcmp(name, "scan_n_reads")) {
        c->opt_n_reads = value < 0 ? -1 : value;
        return APC_OK;
    }
    return apc::fail(c, APC_ERR_INVALID, "unknown option");
}

int apc_measure_int_peak(apc_ctx *c, double *lop3, double *imad, double *mixed) {
    int st = apc::bind(c);
    if (st) return st;
    APC_CUDA(c, apc::measure_int_peak(*c, lop3, imad, mixed));
    return APC_OK;
}

int apc_microbench(apc_ctx *c, const char *name, double *value) {
    int st = apc::bind(c);
    if (st) return st;
    if (!name || !value) return apc::fail(c, APC_ERR_INVALID, "NULL argument");
    APC_CUDA(c, apc::microbench(*c, name, value));
    return APC_OK;
}

} // extern "C"
