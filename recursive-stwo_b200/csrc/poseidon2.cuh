// Poseidon2 over M31, width 16, R_F = 4+4, R_P = 14, x^5 — one state per thread, all 16 words
// in registers.
//
// Replaces the value side of the reference's
//   primitives/poseidon31/src/implementation.rs:108-149 (poseidon2_permute),
//   :7-18 / :20-58 (4x4 and 16x16 external MDS), :103-106 (pow5),
//   parameters.rs:6-190 (constants; generated header include/stwo_b200_poseidon2_constants.h).
//
// Arithmetic plan (range classes from m31.cuh):
//  * S-box: x in C0 -> x^5 in C1 with 3 IMAD.WIDE + 3 LEA.HI + 2 folds.  The product a*(2b) puts
//    floor(ab/2^31) in the high word and (ab mod 2^31)<<1 in the low word, so one LEA.HI reduces.
//  * Linear layers never fold inside: every output is one 64-bit dot product (coefficients
//    pre-doubled, next round constant pre-doubled and added as the seed), then a single LEA.HI.
//    Output class "C0+": <= p + 2^19.
//  * Internal (partial-round) matrix 1 + diag: y_i = S + d_i*s_i computed as
//    IMAD.WIDE(s_i, 2*d_i, 2*S) with S the *unreduced* 64-bit sum, then LEA.HI.
//
// Two code shapes of the same arithmetic:
//   permute<true>  : every round unrolled, constants as immediates (≈4.2 k SASS instructions, 67 KB)
//   permute<false> : round loops rolled, constants from __constant__ memory (≈1 k instructions) —
//                    fits the instruction cache when the permutation sits inside a path/sponge loop.
#pragma once
#include "m31.cuh"
#include "../../include/stwo_b200_poseidon2_constants.h"

namespace poseidon2 {

struct consts {
    u32 diag[16];
    u32 rc_first[64];
    u32 rc_part[14];
    u32 rc_last[64];
};
#if defined(__CUDACC__)
__device__
#endif
static constexpr consts K = {STWO_P2_DIAG16, STWO_P2_RC_FIRST, STWO_P2_RC_PARTIAL, STWO_P2_RC_LAST};

// Tables for the rolled shape.  seeds[k][0..15] = 2*rc[i] of the round that FOLLOWS the k-th
// external MDS of a half (0 when none follows); seeds[k][16..19] = -(sum of the column's four
// seeds) mod 2^64, which removes the seeds from the column sums.
struct rolled_tables {
    u64 first[5][20];    // MDS #0 (initial) .. #4 of the first half
    u64 last[4][20];     // MDS after each of the last four rounds
    u32 rc_part[14];
    u32 rc_last0[16];
    u32 diag2[16];       // 2*diag
};
constexpr rolled_tables make_tables() {
    rolled_tables t = {};
    for (int k = 0; k < 5; k++)
        for (int i = 0; i < 16; i++) t.first[k][i] = k < 4 ? 2ull * K.rc_first[16 * k + i] : 0ull;
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 16; i++) t.last[k][i] = k < 3 ? 2ull * K.rc_last[16 * (k + 1) + i] : 0ull;
    for (int k = 0; k < 5; k++)
        for (int j = 0; j < 4; j++)
            t.first[k][16 + j] = 0ull - (t.first[k][j] + t.first[k][j + 4] + t.first[k][j + 8] + t.first[k][j + 12]);
    for (int k = 0; k < 4; k++)
        for (int j = 0; j < 4; j++)
            t.last[k][16 + j] = 0ull - (t.last[k][j] + t.last[k][j + 4] + t.last[k][j + 8] + t.last[k][j + 12]);
    for (int i = 0; i < 14; i++) t.rc_part[i] = K.rc_part[i];
    for (int i = 0; i < 16; i++) t.rc_last0[i] = K.rc_last[i];
    for (int i = 0; i < 16; i++) t.diag2[i] = 2u * K.diag[i];
    return t;
}
#if defined(__CUDACC__)
static __constant__ rolled_tables c_tab = make_tables();
__device__
#endif
static constexpr rolled_tables h_tab = make_tables();
#if defined(__CUDA_ARCH__)
#define P2_TAB c_tab
#else
#define P2_TAB h_tab
#endif

// hi + (lo >> 1) of a 64-bit value whose true value is half of it: H*2^31 + L -> H + L
HD u32 half_reduce(u64 x2) { return (u32)(x2 >> 32) + ((u32)x2 >> 1); }

// C0+ (<= p + 2^19) or C1 -> x^5 in C1
HD u32 sbox(u32 x) {
    x = m31::fold(x);                                    // C0
    u32 x2d = x << 1;                                    // 2x < 2^32
    u32 x2 = m31::fold(m31::mul_lazy_pre2(x, x2d));      // x^2 in C0
    u32 x4 = m31::fold(m31::mul_lazy_pre2(x2, x2 << 1)); // x^4 in C0
    return m31::mul_lazy_pre2(x4, x2d);                  // x^5 in C1
}

// External MDS circ(2*M4, M4, M4, M4), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
// (reference implementation.rs:7-58).  Inputs: any u32.  Outputs: (M x + rc) in C0+ (<= p + 161).
// seed[i] = 2*rc[i], seed[16+j] = -(column seed sums).
HD void ext_mds(u32 s[16], const u64 *seed) {
    u64 t[16];
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const u32 x0 = s[4 * b], x1 = s[4 * b + 1], x2 = s[4 * b + 2], x3 = s[4 * b + 3];
        t[4 * b + 0] = seed[4 * b + 0] + (u64)x0 * 10u + (u64)x1 * 14u + (u64)x2 * 2u + (u64)x3 * 6u;
        t[4 * b + 1] = seed[4 * b + 1] + (u64)x0 * 8u + (u64)x1 * 12u + (u64)x2 * 2u + (u64)x3 * 2u;
        t[4 * b + 2] = seed[4 * b + 2] + (u64)x0 * 2u + (u64)x1 * 6u + (u64)x2 * 10u + (u64)x3 * 14u;
        t[4 * b + 3] = seed[4 * b + 3] + (u64)x0 * 2u + (u64)x1 * 2u + (u64)x2 * 8u + (u64)x3 * 12u;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u64 col = t[j] + t[j + 4] + t[j + 8] + t[j + 12] + seed[16 + j];
#pragma unroll
        for (int b = 0; b < 4; b++) s[4 * b + j] = half_reduce(t[4 * b + j] + col);
    }
}

// one internal round on state in C0+ ; rc = constant of THIS round (added to s0 before the S-box)
HD void internal_round(u32 s[16], u32 rc) {
    u32 x = m31::fold(s[0]) + rc;          // C1
    s[0] = sbox(x);                        // sbox folds first; result C1
    u64 sum2 = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sum2 += (u64)s[i] * 2u;   // 2*S < 2^37
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = half_reduce(sum2 + (u64)s[i] * (u64)(2u * K.diag[i]));
}

// canonical in, canonical out
template <bool UNROLLED>
HD void permute(u32 s[16]) {
    if (UNROLLED) {
        ext_mds(s, h_tab.first[0]);
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox(s[i]);
            ext_mds(s, h_tab.first[r + 1]);
        }
#pragma unroll
        for (int r = 0; r < 14; r++) internal_round(s, K.rc_part[r]);
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = m31::fold(s[i]) + K.rc_last[i];   // C0+ -> C1
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox(s[i]);
            ext_mds(s, h_tab.last[r]);
        }
    } else {
        ext_mds(s, P2_TAB.first[0]);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox(s[i]);
            ext_mds(s, P2_TAB.first[r + 1]);
        }
#pragma unroll 1
        for (int r = 0; r < 14; r++) internal_round(s, P2_TAB.rc_part[r]);
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = m31::fold(s[i]) + P2_TAB.rc_last0[i];
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox(s[i]);
            ext_mds(s, P2_TAB.last[r]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = m31::canon(m31::fold(s[i]));
}

// two independent states through the rolled rounds together: every loop body holds two independent dependency chains
HD void permute2(u32 a[16], u32 b[16]) {
    ext_mds(a, P2_TAB.first[0]); ext_mds(b, P2_TAB.first[0]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) { a[i] = sbox(a[i]); b[i] = sbox(b[i]); }
        ext_mds(a, P2_TAB.first[r + 1]); ext_mds(b, P2_TAB.first[r + 1]);
    }
#pragma unroll 1
    for (int r = 0; r < 14; r++) { internal_round(a, P2_TAB.rc_part[r]); internal_round(b, P2_TAB.rc_part[r]); }
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = m31::fold(a[i]) + P2_TAB.rc_last0[i]; b[i] = m31::fold(b[i]) + P2_TAB.rc_last0[i]; }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) { a[i] = sbox(a[i]); b[i] = sbox(b[i]); }
        ext_mds(a, P2_TAB.last[r]); ext_mds(b, P2_TAB.last[r]);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = m31::canon(m31::fold(a[i])); b[i] = m31::canon(m31::fold(b[i])); }
}

}  // namespace poseidon2
