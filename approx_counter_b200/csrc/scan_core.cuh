// scan_core.cuh — the column update of the approximate-count automaton (device
// inlines shared by scan_kernel.cu and the pipe-mix microbenchmark in
// peak_kernels.cu).  See scan_kernel.cu for the algorithm.
#pragma once
#include <cstdint>

namespace apc {

// Boolean steps are pinned to one LOP3 each (inline PTX is opaque to nvcc's
// re-association, which otherwise regroups the OR chains into more LOP3s).
// LUT = f(0xF0, 0xCC, 0xAA).
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
__device__ __forceinline__ uint32_t and2(uint32_t a, uint32_t b) { return lop3<0xC0>(a, b, 0u); }       // a & b
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xEA>(a, b, c); } // (a & b) | c
__device__ __forceinline__ uint32_t or3(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xFE>(a, b, c); }    // a | b | c

// ---- automaton state of one thread: four 32-bit words ----------------------------
// Wu–Manber levels R0 ⊆ R1 ⊆ R2 (bit i*F+f = "prefix i+1 of k-mer f matches a text
// suffix with <= e edits").  Per column
//     R0' = S0 & Eq                          S0 = (R0 << F) | 1s
//     R1' = (S1 & Eq) | R0 | S0 | S0'        S1 = (R1 << F) | 1s,  S0' = (R0' << F) | 1s
//     R2' = (S2 & Eq) | R1 | S1 | S1'        S2 =  R2 << F
// The shifted copies S_e are CARRIED from column to column: S0' and S1' are needed by
// this column's deletion terms anyway and are exactly next column's S0 and S1, so a
// column costs three shifts (one IMAD each: R*2^F + 1s) instead of five, and
// 5 LOP3 + 1.5 LOP3 of hit accumulation.  R2 itself is only needed for the hit
// accumulator, so it is not kept.
struct ScanState {
    uint32_t r0[4], r1[4];        // levels 0 and 1 (insertion terms of the next column)
    uint32_t s0[4], s1[4], s2[4]; // shifted levels
    uint32_t a0[4], a1[4], a2[4]; // OR of every column's levels: bit k-1 = "read flagged"
};

template <int NW>
struct Column;

// one unit per word
template <>
struct Column<1> {
    static __device__ __forceinline__ void init(ScanState &st, const uint32_t mul, const uint32_t m) {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            st.r0[w] = 0;
            st.r1[w] = m;           // prefix 1 by one deletion
            const uint32_t r2 = m * mul + m; // prefixes 1..2 by deletions
            st.s0[w] = m;
            st.s1[w] = m * mul + m;
            st.s2[w] = r2 * mul;
            st.a0[w] = 0; st.a1[w] = m; st.a2[w] = r2;
        }
    }
    // one column; the new levels come back in n0/n1/n2 for the accumulator
    static __device__ __forceinline__ void core(ScanState &st, const uint4 eq4, const uint32_t mul, const uint32_t m,
                                                uint32_t (&n0)[4], uint32_t (&n1)[4], uint32_t (&n2)[4]) {
        const uint32_t eq[4] = {eq4.x, eq4.y, eq4.z, eq4.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            n0[u] = and2(st.s0[u], eq[u]);
            const uint32_t t1 = and_or(st.s1[u], eq[u], st.r0[u]);
            const uint32_t s0n = n0[u] * mul + m;
            n1[u] = or3(t1, st.s0[u], s0n);
            const uint32_t t2 = and_or(st.s2[u], eq[u], st.r1[u]);
            const uint32_t s1n = n1[u] * mul + m;
            n2[u] = or3(t2, st.s1[u], s1n);
            st.s2[u] = n2[u] * mul;
            st.r0[u] = n0[u]; st.r1[u] = n1[u];
            st.s0[u] = s0n; st.s1[u] = s1n;
        }
    }
};

// two units of 64 bits: words (0,1) and (2,3) as (lo,hi); always three k-mers per unit.
// The 64-bit shift is IMAD (low word, FMA pipe) + SHF.L.W funnel shift (high word, ALU
// pipe).  IMAD.WIDE + IMAD would keep the ALU pipe free, but IMAD.WIDE measures 2.7x an
// IMAD on B200 and the all-funnel form is 14 % faster end to end (3.52 vs 3.08 Tcol/s, k=20).
template <>
struct Column<2> {
    static constexpr int kF = 3;
    static __device__ __forceinline__ void shl(uint32_t lo, uint32_t hi, uint32_t mul, uint32_t add,
                                               uint32_t &olo, uint32_t &ohi) {
        olo = lo * mul + add;              // IMAD
        ohi = __funnelshift_l(lo, hi, kF); // SHF.L.W.U32.HI
    }
    static __device__ __forceinline__ void init(ScanState &st, const uint32_t mul, const uint32_t m) {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const bool low = (w & 1) == 0; // rows 0..2 of every k-mer live in the low word
            const uint32_t r2 = m * mul + m;
            st.r0[w] = 0;
            st.r1[w] = low ? m : 0;
            st.s0[w] = low ? m : 0;
            st.s1[w] = low ? m * mul + m : 0;
            st.s2[w] = low ? r2 * mul : 0;
            st.a0[w] = 0; st.a1[w] = low ? m : 0; st.a2[w] = low ? r2 : 0;
        }
    }
    static __device__ __forceinline__ void core(ScanState &st, const uint4 eq4, const uint32_t mul, const uint32_t m,
                                                uint32_t (&n0)[4], uint32_t (&n1)[4], uint32_t (&n2)[4]) {
        const uint32_t eq[4] = {eq4.x, eq4.y, eq4.z, eq4.w};
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
            n0[u] = and2(st.s0[u], eq[u]);
            n0[u + 1] = and2(st.s0[u + 1], eq[u + 1]);
            const uint32_t t1l = and_or(st.s1[u], eq[u], st.r0[u]);
            const uint32_t t1h = and_or(st.s1[u + 1], eq[u + 1], st.r0[u + 1]);
            uint32_t s0l, s0h, s1l, s1h;
            shl(n0[u], n0[u + 1], mul, m, s0l, s0h);
            n1[u] = or3(t1l, st.s0[u], s0l);
            n1[u + 1] = or3(t1h, st.s0[u + 1], s0h);
            const uint32_t t2l = and_or(st.s2[u], eq[u], st.r1[u]);
            const uint32_t t2h = and_or(st.s2[u + 1], eq[u + 1], st.r1[u + 1]);
            shl(n1[u], n1[u + 1], mul, m, s1l, s1h);
            n2[u] = or3(t2l, st.s1[u], s1l);
            n2[u + 1] = or3(t2h, st.s1[u + 1], s1h);
            shl(n2[u], n2[u + 1], mul, 0u, st.s2[u], st.s2[u + 1]);
            st.r0[u] = n0[u]; st.r0[u + 1] = n0[u + 1];
            st.r1[u] = n1[u]; st.r1[u + 1] = n1[u + 1];
            st.s0[u] = s0l; st.s0[u + 1] = s0h;
            st.s1[u] = s1l; st.s1[u + 1] = s1h;
        }
    }
};

// One column, accumulate immediately (tail columns).
template <int NW>
__device__ __forceinline__ void step1(ScanState &st, const uint4 eq, const uint32_t mul, const uint32_t m) {
    uint32_t n0[4], n1[4], n2[4];
    Column<NW>::core(st, eq, mul, m, n0, n1, n2);
#pragma unroll
    for (int w = NW - 1; w < 4; w += NW) { // row k-1 lives in the last word of a unit
        st.a0[w] |= n0[w]; st.a1[w] |= n1[w]; st.a2[w] |= n2[w];
    }
}

__device__ __forceinline__ uint4 lds_row(const uint32_t *table, const uint32_t off) {
    return *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(table) + off);
}

// 2*T columns, one 3-input OR per level and word for all of them.  When the unit has
// spare rows above row k-1 (k smaller than the rows the unit can hold) the match tables
// mark those rows "always equal", so a hit in row k-1 keeps moving up one row per column
// instead of being shifted out: the accumulator then only has to look at every T-th
// column (T - 1 <= spare rows) and the per-column cost of hit accumulation drops from
// 1.5 LOP3 to 1.5 / T.  Bits that reach the spare rows at levels 1 and 2 through the
// other transitions are hits of the same or an earlier column at the same or a lower
// level, so any set bit in rows >= k-1 of an accumulator is a genuine hit.
template <int NW, int T>
__device__ __forceinline__ void stepT(ScanState &st, const uint32_t *table, const uint32_t (&off)[2 * T],
                                      const uint32_t mul, const uint32_t m) {
    uint32_t p0[4], p1[4], p2[4], n0[4], n1[4], n2[4];
#pragma unroll
    for (int c = 0; c < T; c++) Column<NW>::core(st, lds_row(table, off[c]), mul, m, p0, p1, p2);
#pragma unroll
    for (int c = T; c < 2 * T; c++) Column<NW>::core(st, lds_row(table, off[c]), mul, m, n0, n1, n2);
#pragma unroll
    for (int w = NW - 1; w < 4; w += NW) {
        st.a0[w] = or3(st.a0[w], p0[w], n0[w]);
        st.a1[w] = or3(st.a1[w], p1[w], n1[w]);
        st.a2[w] = or3(st.a2[w], p2[w], n2[w]);
    }
}

// Two columns, one 3-input OR per level and word.
template <int NW>
__device__ __forceinline__ void step2(ScanState &st, const uint4 eqa, const uint4 eqb, const uint32_t mul,
                                      const uint32_t m) {
    uint32_t p0[4], p1[4], p2[4], n0[4], n1[4], n2[4];
    Column<NW>::core(st, eqa, mul, m, p0, p1, p2);
    Column<NW>::core(st, eqb, mul, m, n0, n1, n2);
#pragma unroll
    for (int w = NW - 1; w < 4; w += NW) {
        st.a0[w] = or3(st.a0[w], p0[w], n0[w]);
        st.a1[w] = or3(st.a1[w], p1[w], n1[w]);
        st.a2[w] = or3(st.a2[w], p2[w], n2[w]);
    }
}

} // namespace apc
