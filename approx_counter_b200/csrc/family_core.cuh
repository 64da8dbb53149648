// family_core.cuh — the FAMILY form of the bit-sliced scan (K1): one CTA scans a whole family of query k-mers
// — an adapter window and its one- and two-error variants — for 1024 reads, and the rows they share are
// computed once per family instead of once per unit.
//
// Why: in the one-warp-per-unit kernels (bitslice_core.cuh) a unit holds at most 48 automaton rows (the
// register file), so a trunk that 100 k-mers share is recomputed by each of the dozen units that hold them;
// with dead-row skipping four fifths of the executed LOP3 are such top rows (bench.py:
// lop3_top_share_of_executed).  The query set of the reference's pipeline (errorCount's input,
// /root/reference/approx_counter.cpp:531-601, :584) is a handful of prefix families (in either direction),
// so the remedy is to share at the family level:
//
//  * the family's TRUNK — the member on the heavy path of the family's prefix tree — is scanned by warp 0
//    (the producer), which also stages the text masks of the column block in shared memory for everyone and
//    PUBLISHES, per text column, the state of the trunk rows other k-mers branch off from;
//  * the other members form units as before — a stretch of ST rows their k-mers share among themselves,
//    then G private tails of T rows — but a unit no longer starts at row 0: it ATTACHES to trunk row
//    a - 1 (a = K - T - ST, the bases it shares with the trunk) and takes its carry from the published
//    state (three LDS per column) instead of computing a rows;
//  * seven consumer warps walk the same column block after the producer (two-stage ring, named barriers:
//    the producer/consumer pattern of the PTX manual's bar.arrive / bar.sync example).  A consumer owns a
//    LIST of units whose rows sum to about one register file's worth; between column blocks the state of a
//    unit is parked in shared memory (6 LDS/STS per row and 8 columns against 40 LOP3), so a unit's shape
//    is compile-time code while a warp's work is data — balanced by the planner, no idle warps.
//
// Dead-row skipping carries over: a unit's rows below absolute row M are computed in every column, the rest
// only while something can reach them; a unit that attaches at or below row M has NO top rows and costs one
// carry load + one vote in a quiet column.  Exactness argument as in bitslice_core.cuh; the only new fact is
// that for a unit whose first row is deep, "level 2 of row M-1" is replaced by "level 2 of the trunk row it
// attaches to" (its carry-in), which is what reaches it.
#pragma once

#include "bitslice_core.cuh"

namespace apc {

constexpr int kFamWarps = 8;       // 1 producer + 7 consumers, one CTA per SM (255 registers per thread)
constexpr int kFamConsumers = kFamWarps - 1;
constexpr int kFamStageCols = 8;   // text columns per ring stage
constexpr int fam_stage_words(int k) { return kFamStageCols * 4 * 32 + (kFamStageCols + 1) * fam_pub(k) * 3 * 32; } // u32 per stage
// per pass, staged once per job: the units' k-mers (at most one per parked row), their perm entries, the unit records
constexpr int fam_table_words(int k) { return fam_state_rows(k) * 2 + fam_state_rows(k) + (fam_state_rows(k) / 4) * 2; }
constexpr size_t fam_smem_bytes(int k) {
    return (size_t)(2 * fam_stage_words(k) + fam_state_rows(k) * 3 * 32 + fam_table_words(k) + 32) * sizeof(uint32_t);
}
static_assert(fam_smem_bytes(16) <= 227 * 1024 && fam_smem_bytes(20) <= 227 * 1024 && fam_smem_bytes(32) <= 227 * 1024, "shared memory of one SM");

__device__ __forceinline__ void fam_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kFamWarps * 32) : "memory"); }
__device__ __forceinline__ void fam_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kFamWarps * 32) : "memory"); }

// A value every lane holds alike, said so that ptxas keeps what is derived from it in uniform registers (the row
// offsets then cost no per-row address arithmetic: LDS R, [R + UR]).
__device__ __forceinline__ uint64_t fam_uniform64(uint64_t v) {
    const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, 0), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), 0);
    return ((uint64_t)hi << 32) | lo;
}

struct FamStage {
    uint32_t *text; // [col][base][lane]
    uint32_t *pub;  // [col + 1][published row][level][lane]; slot 0 = the column before the stage
};
template <int K>
__device__ __forceinline__ FamStage fam_stage(uint32_t *smem, int buf) {
    uint32_t *s = smem + buf * fam_stage_words(K);
    return FamStage{s, s + kFamStageCols * 4 * 32};
}

// One column block (kFamStageCols columns) of one consumer unit: stretch of ST rows + G tails of T rows attached to
// published trunk row `slot` (= a - 1 - first published row).  TOPS / TOPT: rows of the stretch / of every tail that lie
// above the checkpoint row M and are computed in every column.
template <int PUB, int ST, int T, int G, int TOPS, int TOPT>
__device__ __forceinline__ void fam_unit_block(const FamStage &stg, uint32_t *__restrict__ state, const uint32_t slot,
                                               const uint64_t *kmers, const uint32_t *perm,
                                               const bool first, const bool last, const uint32_t vm,
                                               unsigned long long *__restrict__ counts, unsigned long long &deep_lop3,
                                               const uint32_t lane) {
    static_assert(TOPS <= ST && TOPT <= T && (TOPT == 0 || TOPS == ST), "top rows are a prefix of the unit's rows");
    constexpr int STN = ST > 0 ? ST : 1; // zero-length arrays are not C++
    constexpr bool kAllTop = TOPS == ST && TOPT == T;
    constexpr int kDeepRows = (ST - TOPS) + G * (T - TOPT);
    uint32_t off_s[STN], off_t[G][T];
    {
        const uint64_t k0 = fam_uniform64(kmers[0]);
#pragma unroll
        for (int i = 0; i < ST; i++) off_s[i] = (uint32_t)((k0 >> (2 * (T + ST - 1 - i))) & 3u) * kPlaneRow;
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        const uint64_t kg = fam_uniform64(kmers[g]);
#pragma unroll
        for (int i = 0; i < T; i++) off_t[g][i] = (uint32_t)((kg >> (2 * (T - 1 - i))) & 3u) * kPlaneRow;
    }
    uint32_t s0[STN], s1[STN], s2[STN], x0[G][T], x1[G][T], x2[G][T];
    if (first) { // rows >= 2 start empty
#pragma unroll
        for (int j = 0; j < ST; j++) s0[j] = s1[j] = s2[j] = 0;
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int j = 0; j < T; j++) x0[g][j] = x1[g][j] = x2[g][j] = 0;
    } else {
        const uint32_t *p = state + lane;
#pragma unroll
        for (int j = 0; j < ST; j++) { s0[j] = p[0]; s1[j] = p[32]; s2[j] = p[64]; p += 96; }
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int j = 0; j < T; j++) { x0[g][j] = p[0]; x1[g][j] = p[32]; x2[g][j] = p[64]; p += 96; }
    }
    // the trunk row this unit branches off from: its state before the block's first column
    const uint32_t *pub = stg.pub + slot * 96 + lane;
    uint32_t t0 = pub[0], t1 = pub[32], t2 = pub[64];
    bool deep_zero = false; // unknown at the start of a block: the first quiet column re-checks
    uint32_t n_run = 0;
    // two columns per iteration: the second writes the state back into the registers the first read it from
    // (a one-column loop body costs a register move per state word)
#pragma unroll 1
    for (int cp = 0; cp < kFamStageCols / 2; cp++) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int col = 2 * cp + half;
        pub += PUB * 96;
        const uint32_t u0 = pub[0], u1 = pub[32], u2 = pub[64];
        BsCarry c{t0, t1, t2, u0, u1};
        const uint32_t ck_in = t2 | u2; // can anything reach the unit's first row in this column
        t0 = u0; t1 = u1; t2 = u2;
        const char *slot_lane = reinterpret_cast<const char *>(stg.text + col * 128) + lane * 4;
        if constexpr (kAllTop) {
            if constexpr (ST > 0) bs_rows<STN, 2, false, 0, ST>(s0, s1, s2, c, slot_lane, off_s);
#pragma unroll
            for (int g = 0; g < G; g++) {
                BsCarry cg = c;
                bs_rows<T, 2, true>(x0[g], x1[g], x2[g], cg, slot_lane, off_t[g]);
            }
        } else if constexpr (TOPT > 0) { // split inside the tails
            if constexpr (ST > 0) bs_rows<STN, 2, false, 0, ST>(s0, s1, s2, c, slot_lane, off_s);
            BsCarry cg[G];
            uint32_t ck = 0;
#pragma unroll
            for (int g = 0; g < G; g++) {
                cg[g] = c;
                ck |= x2[g][TOPT - 1];
                bs_rows<T, 2, true, 0, TOPT>(x0[g], x1[g], x2[g], cg[g], slot_lane, off_t[g]);
                ck |= x2[g][TOPT - 1];
            }
            bool run = __any_sync(0xFFFFFFFFu, ck != 0);
            if (!run && !deep_zero) {
                uint32_t z = 0;
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int j = TOPT; j < T - 1; j++) z |= x2[g][j];
                run = __any_sync(0xFFFFFFFFu, z != 0);
            }
            deep_zero = !run;
            n_run += run;
            if (run) {
#pragma unroll
                for (int g = 0; g < G; g++) bs_rows<T, 2, true, TOPT, T>(x0[g], x1[g], x2[g], cg[g], slot_lane, off_t[g]);
            }
        } else { // split inside the stretch (TOPS > 0) or above the unit (TOPS == 0: every row is deep)
            uint32_t ck = ck_in;
            if constexpr (TOPS > 0) {
                ck = s2[TOPS - 1];
                bs_rows<STN, 2, false, 0, TOPS>(s0, s1, s2, c, slot_lane, off_s);
                ck |= s2[TOPS - 1];
            }
            bool run = __any_sync(0xFFFFFFFFu, ck != 0);
            if (!run && !deep_zero) {
                uint32_t z = 0;
#pragma unroll
                for (int j = TOPS; j < ST; j++) z |= s2[j];
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int j = 0; j < T - 1; j++) z |= x2[g][j];
                run = __any_sync(0xFFFFFFFFu, z != 0);
            }
            deep_zero = !run;
            n_run += run;
            if (run) {
                if constexpr (ST > TOPS) bs_rows<STN, 2, false, TOPS, ST>(s0, s1, s2, c, slot_lane, off_s);
#pragma unroll
                for (int g = 0; g < G; g++) {
                    BsCarry cg = c;
                    bs_rows<T, 2, true>(x0[g], x1[g], x2[g], cg, slot_lane, off_t[g]);
                }
            }
        }
    }
    }
    deep_lop3 += n_run * (5 * kDeepRows);
    if (!last) {
        uint32_t *p = state + lane;
#pragma unroll
        for (int j = 0; j < ST; j++) { p[0] = s0[j]; p[32] = s1[j]; p[64] = s2[j]; p += 96; }
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int j = 0; j < T; j++) { p[0] = x0[g][j]; p[32] = x1[g][j]; p[64] = x2[g][j]; p += 96; }
    } else {
#pragma unroll
        for (int g = 0; g < G; g++) { // the sticky last row of each tail holds the hits (:589-593)
            const uint32_t cnt = __popc(x0[g][T - 1] & vm) + __popc(x1[g][T - 1] & vm) + __popc(x2[g][T - 1] & vm);
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, cnt);
            const uint32_t pi = perm[g];
            if (lane == 0 && t && pi != 0xFFFFFFFFu) atomicAdd(&counts[pi & 0x7FFFFFFFu], (unsigned long long)t);
        }
    }
}

// dispatch on the unit's shape (compile-time code per shape, the work list of a warp is data)
template <int K, int S>
__device__ __forceinline__ void fam_unit_dispatch(const uint32_t shape, const FamStage &stg, uint32_t *state,
                                                  const uint64_t *kmers, const uint32_t *perm, const bool first,
                                                  const bool last, const uint32_t vm, unsigned long long *counts,
                                                  unsigned long long &deep_lop3, const uint32_t lane) {
    if constexpr (S < kFamShapes) {
        constexpr FamShape sh = fam_shape(S);
        if constexpr (fam_shape_valid(K, S)) {
            if (shape == (uint32_t)S) {
                constexpr int a = K - sh.t - sh.st, M = bs_check_row(K);
                constexpr int tops = M - a < 0 ? 0 : M - a > sh.st ? sh.st : M - a;
                constexpr int topt = M - a - sh.st < 0 ? 0 : M - a - sh.st > sh.t ? sh.t : M - a - sh.st;
                fam_unit_block<fam_pub(K), sh.st, sh.t, sh.g, tops, topt>(stg, state, (uint32_t)(a - 1 - (K - 1 - fam_pub(K))), kmers,
                                                                          perm, first, last, vm, counts, deep_lop3, lane);
                return;
            }
        }
        fam_unit_dispatch<K, S + 1>(shape, stg, state, kmers, perm, first, last, vm, counts, deep_lop3, lane);
    }
}

template <int K>
__global__ void __launch_bounds__(kFamWarps * 32, 1)
bs_family_kernel(const uint4 *__restrict__ planes, const uint32_t sg_first, const uint32_t n_sg, const uint32_t cols,
                 const uint32_t read_len, const uint64_t range_lo, const uint64_t range_hi,
                 const FamPass *__restrict__ passes, const FamUnit *__restrict__ units, const uint64_t *__restrict__ kmers,
                 const uint32_t *__restrict__ perm, const uint32_t n_passes, const uint32_t n_jobs,
                 unsigned long long *__restrict__ counts, unsigned int *__restrict__ job_counter,
                 unsigned long long *__restrict__ deep_lop3_out) {
    constexpr int kFamPub = fam_pub(K);
    static_assert(K - 1 - kFamPub >= 2, "the published trunk rows lie below the constant rows 0-1");
    extern __shared__ __align__(16) uint32_t fam_smem[];
    uint32_t *state_area = fam_smem + 2 * fam_stage_words(K);
    uint64_t *s_kmers = reinterpret_cast<uint64_t *>(state_area + fam_state_rows(K) * 96); // 8-byte aligned: all terms are even
    uint32_t *s_perm = reinterpret_cast<uint32_t *>(s_kmers + fam_state_rows(K));
    FamUnit *s_units = reinterpret_cast<FamUnit *>(s_perm + fam_state_rows(K));
    volatile uint32_t *s_job = reinterpret_cast<uint32_t *>(s_units + fam_state_rows(K) / 4);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_stages = (read_len + kFamStageCols - 1) / kFamStageCols;
    constexpr int kPubFirst = K - 1 - kFamPub; // first published trunk row
    unsigned long long deep_lop3 = 0;

    for (;;) {
        if (threadIdx.x == 0) {
            uint32_t job = atomicAdd(job_counter, 1u);
            if (job == n_jobs + gridDim.x - 1u) atomicExch(job_counter, 0u); // last fetch of the launch re-arms the queue
            *s_job = job;
        }
        fam_bar_sync(0);
        // (the broadcast from lane 0 tells ptxas that the value is warp-uniform: everything derived from it —
        // the unit records, the k-mers, the row offsets — then lives in uniform registers)
        const uint32_t job = __shfl_sync(0xFFFFFFFFu, *s_job, 0);
        if (job >= n_jobs) break;
        const uint32_t pi = job % n_passes, sg = job / n_passes;
        const FamPass *pass = passes + pi;
        const bool reverse = (__ldg(&pass->trunk_perm) >> 31) != 0;
        {   // stage the pass's tables in shared memory: no global load is left inside the column blocks
            const uint32_t u0 = __ldg(&pass->unit_first), nu = __ldg(&pass->warp_end[kFamConsumers - 1]);
            const uint32_t k0 = nu ? __ldg(&units[u0].first_kmer) : 0u;
            uint32_t nk = 0;
            if (nu) {
                const FamUnit lastu = units[u0 + nu - 1];
                nk = lastu.first_kmer + (uint32_t)fam_shape_g(lastu.shape) - k0;
            }
            for (uint32_t i = threadIdx.x; i < nu; i += blockDim.x) {
                FamUnit un = units[u0 + i];
                un.first_kmer -= k0;
                s_units[i] = un;
            }
            for (uint32_t i = threadIdx.x; i < nk; i += blockDim.x) {
                s_kmers[i] = __ldg(kmers + k0 + i);
                s_perm[i] = __ldg(perm + k0 + i);
            }
        }
        fam_bar_sync(0); // tables staged; everyone has read the job before thread 0 may overwrite it
        const uint32_t vm = bs_valid_mask(((uint64_t)(sg_first + sg) * kGroupsPerSuper + lane) * 32, range_lo, range_hi);

        if (warp == 0) {
            // ---- producer: stage the text masks, scan the trunk, publish the rows units branch off from
            const uint64_t trunk = __ldg(&pass->trunk);
            uint32_t off[K];
#pragma unroll
            for (int i = 0; i < K; i++) off[i] = (uint32_t)((trunk >> (2 * (K - 1 - i))) & 3u) * kPlaneRow;
            uint32_t r0[K], r1[K], r2[K];
            bs_rows_init<K, 0>(r0, r1, r2);
            const int64_t cstep = reverse ? -(int64_t)kGroupsPerSuper : (int64_t)kGroupsPerSuper;
            const size_t col0 = reverse ? (size_t)n_stages * kFamStageCols - 1 : 0;
            const uint4 *p = planes + ((size_t)(sg_first + sg) * cols + col0) * kGroupsPerSuper + lane;
            uint4 ma = __ldg(p), mb = __ldg(p + cstep);
            for (uint32_t s = 0; s < n_stages; s++) {
                const int buf = (int)(s & 1);
                if (s >= 2) fam_bar_sync(3 + buf); // the consumers are done with what this buffer held
                const FamStage stg = fam_stage<K>(fam_smem, buf);
                {   // slot 0: the trunk's state before the stage's first column
                    uint32_t *q = stg.pub + lane;
#pragma unroll
                    for (int r = 0; r < kFamPub; r++) { q[0] = r0[kPubFirst + r]; q[32] = r1[kPubFirst + r]; q[64] = r2[kPubFirst + r]; q += 96; }
                }
#pragma unroll 1
                for (int pr = 0; pr < kFamStageCols / 2; pr++) {
                    p += 2 * cstep;
                    const uint4 na = __ldg(p), nb = __ldg(p + cstep); // padded by kBsPadCols columns at both ends
                    uint32_t *ta = stg.text + (2 * pr) * 128 + lane, *tb = ta + 128;
                    ta[0] = ma.x; ta[32] = ma.y; ta[64] = ma.z; ta[96] = ma.w;
                    tb[0] = mb.x; tb[32] = mb.y; tb[64] = mb.z; tb[96] = mb.w;
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        BsCarry c = bs_carry_init();
                        bs_rows<K, 0, true>(r0, r1, r2, c, reinterpret_cast<const char *>(half ? tb : ta), off);
                        uint32_t *q = stg.pub + (2 * pr + half + 1) * (kFamPub * 96) + lane;
#pragma unroll
                        for (int r = 0; r < kFamPub; r++) { q[0] = r0[kPubFirst + r]; q[32] = r1[kPubFirst + r]; q[64] = r2[kPubFirst + r]; q += 96; }
                    }
                    ma = na; mb = nb;
                }
                fam_bar_arrive(1 + buf); // stage full
            }
            // the last stages' "empty" arrivals have no refill waiting for them: consume them so that the
            // barriers are balanced when the next job starts
            if (n_stages >= 2) fam_bar_sync(3 + (int)(n_stages & 1));
            fam_bar_sync(3 + (int)((n_stages - 1) & 1));
            const uint32_t cnt = __popc(r0[K - 1] & vm) + __popc(r1[K - 1] & vm) + __popc(r2[K - 1] & vm);
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, cnt);
            const uint32_t tp = __ldg(&pass->trunk_perm) & 0x7FFFFFFFu;
            if (lane == 0 && t && tp != 0x7FFFFFFFu) atomicAdd(&counts[tp], (unsigned long long)t);
        } else {
            // ---- consumers: every unit of this warp's list, one column block after the other
            const uint32_t w = warp - 1;
            const uint32_t u_first = __shfl_sync(0xFFFFFFFFu, w ? __ldg(&pass->warp_end[w - 1]) : 0u, 0);
            const uint32_t u_end = __shfl_sync(0xFFFFFFFFu, __ldg(&pass->warp_end[w]), 0);
            for (uint32_t s = 0; s < n_stages; s++) {
                const int buf = (int)(s & 1);
                fam_bar_sync(1 + buf);
                const FamStage stg = fam_stage<K>(fam_smem, buf);
                for (uint32_t u = u_first; u < u_end; u++) {
                    const uint2 raw = *reinterpret_cast<const uint2 *>(&s_units[u]);
                    const uint32_t first_kmer = __shfl_sync(0xFFFFFFFFu, raw.x, 0), sr = __shfl_sync(0xFFFFFFFFu, raw.y, 0);
                    fam_unit_dispatch<K, 0>(sr & 0xFFFFu, stg, state_area + (size_t)(sr >> 16) * 96, s_kmers + first_kmer,
                                            s_perm + first_kmer, s == 0, s + 1 == n_stages, vm, counts, deep_lop3, lane);
                }
                fam_bar_arrive(3 + buf); // stage consumed
            }
        }
    }
    if (lane == 0 && deep_lop3) atomicAdd(deep_lop3_out, deep_lop3);
}

// ---- launch: one more concurrent launch of the scan (its own stream and job queue)
template <int K>
static cudaError_t launch_family_k(BsLaunchCtx &l) {
    const Ctx &c = *l.c;
    if constexpr (K < kFamMinK) {
        return c.fam_passes ? cudaErrorInvalidValue : cudaSuccess; // the planner builds no family for this k
    } else {
        constexpr size_t kFamSmemBytes = fam_smem_bytes(K);
        if (c.fam_passes == 0) return cudaSuccess;
        // per device and cheap: set on every launch (a process may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(bs_family_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFamSmemBytes);
        if (e != cudaSuccess) return e;
        const uint64_t jobs = (uint64_t)l.r.n_sg * c.fam_passes;
        if (jobs > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.sm_count, jobs);
        const int slot = l.slot++;
        cudaStream_t s = c.stream;
        if (slot > 0) {
            s = c.bs_streams[slot - 1];
            if ((e = cudaStreamWaitEvent(s, c.bs_fork, 0)) != cudaSuccess) return e;
        }
        bs_family_kernel<K><<<grid, kFamWarps * 32, kFamSmemBytes, s>>>(
            c.planes(), l.r.sg_first, l.r.n_sg, c.chunks * kChunkBases, c.max_len, l.r.lo, l.r.hi, c.d_fam_passes, c.d_fam_units,
            c.d_fam_kmers, c.d_fam_perm, c.fam_passes, (uint32_t)jobs, l.d_counts, c.d_job_counter + slot, c.d_deep_lop3);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        (*l.launches)++;
        if (slot > 0) {
            if ((e = cudaEventRecord(c.bs_join[slot - 1], s)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(c.stream, c.bs_join[slot - 1], 0)) != cudaSuccess) return e;
        }
        // statistics: the rows the plan computes in every column / in all
        const double cols_sg = (double)l.r.n_sg * (double)(((c.max_len + kFamStageCols - 1) / kFamStageCols) * kFamStageCols);
        l.lop3_top += cols_sg * c.fam_lop3_top_per_col;
        l.lop3_all += cols_sg * c.fam_lop3_all_per_col;
        return cudaSuccess;
    }
}

} // namespace apc
