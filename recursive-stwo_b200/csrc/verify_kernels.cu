// K3-K5 and the batch driver: kernels that dispatch the per-thread verifier stages of verify.cuh, and the C ABI of
// stwo_b200_verify_proofs_batch*.  Grid layouts keep the lanes of a warp on the same tree / log-size group of
// different proofs, so same-shape batches execute almost divergence-free.
#include "common.cuh"
#include "verify.cuh"
#include "synth.cuh"
#include <vector>
#include <map>
#include <algorithm>
#include <array>
#include <string.h>
#include <stdlib.h>

using namespace stwo_b200;
using verify::Workspace;

namespace {
constexpr int kT = 64;

// every kernel covers the proofs [p0, p0 + pn) so that a batch can be cut into slices that run on separate streams
__global__ void __launch_bounds__(kT) k_fiat_shamir(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn) verify::stage_fiat_shamir(ws, p0 + idx);
}
__global__ void __launch_bounds__(kT) k_single_tree(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn * 4) verify::stage_single_tree(ws, p0 + idx % pn, idx / pn);
}
__global__ void __launch_bounds__(kT) k_group(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn * fri::MAX_LOGS) verify::stage_group(ws, p0 + idx % pn, idx / pn);
}
// 16 lanes per (proof, log-size group): the coefficient chain of build_group is strided over the lanes
__global__ void __launch_bounds__(kT) k_group_coop(const Workspace ws, u32 p0, u32 pn) {
    const u32 unit = (blockIdx.x * kT + threadIdx.x) / 16;
    if (unit >= pn * fri::MAX_LOGS) return;              // whole groups leave together
    struct Co16 {
        u32 l; unsigned mask;
        __device__ __forceinline__ u32 lane() const { return l; }
        __device__ __forceinline__ u32 size() const { return 16; }
        __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    } co;
    co.l = threadIdx.x % 16;
    co.mask = 0xffffu << (16 * ((threadIdx.x % 32) / 16));
    verify::stage_group_coop(co, ws, p0 + unit % pn, unit / pn);
}
__global__ void __launch_bounds__(kT) k_answer(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    const u32 per_g = pn * ws.shape.n_queries;
    if (idx < per_g * fri::MAX_LOGS) {
        u32 g = idx / per_g, r = idx % per_g;
        verify::stage_answer(ws, p0 + r / ws.shape.n_queries, g, r % ws.shape.n_queries);
    }
}
__global__ void __launch_bounds__(kT) k_folds(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn) verify::stage_folds(ws, p0 + idx);
}
__global__ void __launch_bounds__(kT) k_pair_tree(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn * ws.shape.n_fri_trees()) verify::stage_pair_tree(ws, p0 + idx % pn, idx / pn);
}
// Cooperative tree rebuilds (decommit_coop.cuh): G lanes of a warp per (proof, tree).  G trades latency (long chains per
// lane when small) against lane utilisation (idle lanes near the root when large): 16 for small batches, 4 for large ones.
template <int G>
struct CoopGroup {
    u32 l; unsigned mask;
    __device__ __forceinline__ u32 lane() const { return l; }
    __device__ __forceinline__ u32 size() const { return G; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
constexpr int kCoopThreads = 64;          // upper bound; the launch shrinks the block when the group tables would not fit
constexpr int kTreeBlocksDefault = 8;     // measured on B200 at 4096 proofs: 1.46 + 2.38 ms (1) -> 1.40 + 2.27 ms (8) -> 1.37 + 2.31 ms (10)
// MB: resident blocks per SM the register allocation aims at (1: unconstrained = 152 registers, 6 blocks; 8: 128 registers, no spills;
// 10: 96 registers, a few spilled words): STWO_B200_TREE_BLOCKS selects, see tree_min_blocks()
template <int G, int MB>
__global__ void __launch_bounds__(kCoopThreads, MB) k_single_tree_coop(const Workspace ws, u32 p0, u32 pn, u32 tab_words) {
    extern __shared__ u32 smem[];
    const u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (grp >= pn * 4) return;                                       // whole groups leave together
    CoopGroup<G> co;
    co.l = threadIdx.x % G;
    co.mask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (G * ((threadIdx.x % 32) / G));
    verify::stage_single_tree_coop(co, ws, p0 + grp % pn, grp / pn, smem + (threadIdx.x / G) * tab_words);
}
template <int G, int MB>
__global__ void __launch_bounds__(kCoopThreads, MB) k_pair_tree_coop(const Workspace ws, u32 p0, u32 pn, u32 tab_words) {
    extern __shared__ u32 smem[];
    const u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (grp >= pn * ws.shape.n_fri_trees()) return;
    CoopGroup<G> co;
    co.l = threadIdx.x % G;
    co.mask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (G * ((threadIdx.x % 32) / G));
    verify::stage_pair_tree_coop(co, ws, p0 + grp % pn, grp / pn, smem + (threadIdx.x / G) * tab_words);
}

template <int G>
__global__ void __launch_bounds__(kCoopThreads) k_parse_coop(const Workspace ws, u32 p0, u32 pn, u32 tab_words) {
    extern __shared__ u32 smem[];
    const u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (grp >= pn) return;
    CoopGroup<G> co;
    co.l = threadIdx.x % G;
    co.mask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (G * ((threadIdx.x % 32) / G));
    verify::stage_parse_coop(co, ws, p0 + grp, smem + (threadIdx.x / G) * tab_words);
}
// ---- transcript with the Poseidon2 state spread over 16 lanes (one word per lane, warp-shuffle MDS mixing) ---------------
// The Fiat-Shamir chain is strictly sequential (105-255 permutations) and gates every later stage, so what matters is the
// LATENCY of one permutation: ~6.3 us with the state in one thread's registers (tools/perm_latency_microbench.cu), ~2 us
// with one state word per lane: the S-boxes of a full round run in parallel, the external matrix is 6 shuffles, the internal
// matrix a 4-step butterfly sum.  Same transcript order as fs::transcript (components/recursive/fiat_shamir/src/lib.rs:39-131).
// The warp holds two proofs (one per half) and stays CONVERGENT from the first instruction to the last: the shuffles name all 32 lanes
// (width 16 keeps them inside a half), so they compile to bare SHFL -- with a half-warp mask every one of the ~150 shuffles of a
// permutation came wrapped in WARPSYNC.COLLECTIVE / BSSY / BSYNC (2 500 of them in the kernel), on the critical path of a latency chain.
// A half without a proof of its own (ragged tail, a proof that did not parse) repeats its neighbour's and stores nothing.
struct Lanes16 {
    u32 l;                                                  // l = 0..15: which state word this lane holds
    bool active;                                            // this half owns a proof: its stores count
    // this lane's constants, loaded once per kernel: the chain is latency bound, a constant fetched inside a round is on the critical path
    u32 rcf[4], rcl[4], dg, c0, c1, c2, c3;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int r = 0; r < 4; r++) { rcf[r] = poseidon2::K.rc_first[16 * r + l]; rcl[r] = poseidon2::K.rc_last[16 * r + l]; }
        dg = poseidon2::K.diag[l];
        const u32 row = l & 3u;
        c0 = row == 0 ? 5u : row == 1 ? 4u : 1u; c1 = row == 0 ? 7u : row == 1 ? 6u : row == 2 ? 3u : 1u;
        c2 = row < 2 ? 1u : row == 2 ? 5u : 4u; c3 = row == 0 ? 3u : row == 1 ? 1u : row == 2 ? 7u : 6u;
    }
    // any value below 2^61 whose high word is small -> canonical, in one step: hi * 2^32 + lo = 2 hi + (lo mod 2^31) + (lo >> 31) (mod p).
    // The general m31::red64 folds twice; on this chain every dependent instruction is ~5 cycles of latency nobody else fills.
    __device__ __forceinline__ static u32 red_small(u64 x) {
        const u32 lo = (u32)x, hi = (u32)(x >> 32);          // hi < 2^29: the sum below is < 2p
        return m31::canon((lo & M31_P) + (lo >> 31) + (hi << 1));
    }
    __device__ __forceinline__ u32 get(u32 x, u32 src) const { return __shfl_sync(0xffffffffu, x, src, 16); }
    __device__ __forceinline__ u32 get_xor(u32 x, u32 m) const { return __shfl_xor_sync(0xffffffffu, x, m, 16); }
    // circ(2 M4, M4, M4, M4) x, M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] (primitives/poseidon31/src/implementation.rs:7-58)
    __device__ __forceinline__ u32 ext_mds(u32 x) const {
        const u32 base = l & ~3u;
        const u32 x0 = get(x, base), x1 = get(x, base + 1), x2 = get(x, base + 2), x3 = get(x, base + 3);
        const u32 t = red_small((u64)c0 * x0 + (u64)c1 * x1 + (u64)c2 * x2 + (u64)c3 * x3);
        // column sum over the four blocks as one lazy 34-bit sum: two shuffles side by side instead of two dependent add-reduce steps
        const u32 t4 = get_xor(t, 4), t8 = get_xor(t, 8), t12 = get_xor(t, 12);
        return red_small((u64)t + t + t4 + t8 + t12);
    }
    __device__ __forceinline__ static u32 pow5(u32 x) { const u32 x2 = m31::mulc(x, x), x4 = m31::mulc(x2, x2); return m31::mulc(x4, x); }
    __device__ u32 permute(u32 x) const {                   // canonical in, canonical out; implementation.rs:108-149
        x = ext_mds(x);
#pragma unroll
        for (int r = 0; r < 4; r++) x = ext_mds(pow5(m31::addc(x, rcf[r])));
#pragma unroll
        for (int r = 0; r < 14; r++) {
            const u32 sb = pow5(m31::addc(x, poseidon2::K.rc_part[r]));     // every lane computes it, lane 0 keeps it: no branch
            x = l == 0 ? sb : x;
            // sum of the 16 words: two butterfly steps on lazy sums (values < 2^31, four of them < 2^33), then one reduction
            const u64 s1 = (u64)x + get_xor(x, 1) + get_xor(x, 2) + get_xor(x, 3);
            const u32 q = red_small(s1);
            const u64 s2 = (u64)q + get_xor(q, 4) + get_xor(q, 8) + get_xor(q, 12);
            x = red_small(s2 + (u64)x * dg);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) x = ext_mds(pow5(m31::addc(x, rcl[r])));
        return x;
    }
};
struct Channel16 {                                          // primitives/channel/src/lib.rs:23-58 on a lane-spread state
    Lanes16 g; u32 s, n_sent, n_perms;
    u32 *sink;                                              // optional: output state of permutation k at sink[16 k ..], a word per lane
    size_t in_delta;                                        // != 0: its input state in_delta words further
    __device__ void absorb(u32 rate_word) {
        s = g.l < 8 ? rate_word : s;
        if (sink && in_delta && g.active) sink[in_delta + 16 * (size_t)n_perms + g.l] = s;
        s = g.permute(s);
        if (sink && g.active) sink[16 * (size_t)n_perms + g.l] = s;
        n_sent = 0; n_perms++;
    }
    __device__ void mix8(const u32 *w8) { absorb(g.l < 8 ? w8[g.l] : 0u); }
    __device__ void mix4(const u32 *w4) { absorb(g.l < 4 ? w4[g.l] : 0u); }
    __device__ void mix4v(u32 v) { absorb(g.l < 4 ? v : 0u); }              // lanes 0..3 already hold the four words
    __device__ void mix44(const u32 *a, const u32 *b) { absorb(g.l < 4 ? a[g.l] : g.l < 8 ? b[g.l - 4] : 0u); }
    __device__ u32 draw() {                                 // lanes 0..7 return the eight drawn words
        const u32 t0 = g.l == 0 ? n_sent : g.l < 8 ? 0u : s;
        if (sink && in_delta && g.active) sink[in_delta + 16 * (size_t)n_perms + g.l] = t0;
        const u32 t = g.permute(t0);
        if (sink && g.active) sink[16 * (size_t)n_perms + g.l] = t;
        n_sent++; n_perms++;
        return t;
    }
};
// Trip counts come from the batch's shape (a kernel parameter: provably warp-uniform, so the loops stay convergent for the compiler too),
// not from the proof's descriptor (equal for every proof that parsed, but loaded from memory).
__device__ __forceinline__ void transcript16(const Lanes16 &g, const u32 *w, const proof::Desc &d, fs::Out &o, u32 *sink, const size_t in_delta,
                                             const u32 n_inner, const u32 n_last_coeffs, const u32 n_queries, const u32 pow_bits) {
    Channel16 ch{g, 0u, 0u, 0u, sink, in_delta};
    const u32 l = g.l;
    const bool act = g.active;
    auto store_q = [&](qm31_t *dst, u32 t, u32 first_lane) { if (act && l >= first_lane && l < first_lane + 4) dst->v[l - first_lane] = t; };
    ch.mix8(w + d.commitments[0]);
    ch.mix4v(l == 0 ? d.log_size_plonk : 0u);
    ch.mix4v(l == 0 ? d.log_size_poseidon : 0u);
    ch.mix8(w + d.commitments[1]);
    u32 t = ch.draw();
    store_q(&o.z, t, 0); store_q(&o.alpha, t, 4);
    ch.mix8(w + d.stmt1);
    ch.mix8(w + d.commitments[2]);
    t = ch.draw(); store_q(&o.random_coeff, t, 0);
    ch.mix8(w + d.commitments[3]);
    t = ch.draw(); store_q(&o.oods_t, t, 0);
    {
        // the OODS point from t (every lane computes it from the broadcast words; lane 0 stores)
        const qm31_t ot = qm31::mk(g.get(t, 0), g.get(t, 1), g.get(t, 2), g.get(t, 3));
        if (act && l == 0) {
            const qm31_t t2 = fs::qmul(ot, ot);
            const qm31_t inv = fs::qinv(qm31::add_m31(t2, 1));
            o.oods_x = fs::qmul(fs::qsub(qm31::one(), t2), inv);
            o.oods_y = fs::qmul(fs::qadd(ot, ot), inv);
        }
    }
    // The chain is latency bound, so no permutation may wait for its input: the rate words of absorb k + 1 are loaded before the
    // permutation of absorb k starts (the loads do not depend on the state).
    // sampled values, flattened tree -> column -> mask, two per permutation: word of this lane (lanes 0..7) of pair j
    auto pair_word = [&](u32 j) -> u32 {
        const u32 sidx = 2 * j + (l >> 2);
        if (l >= 8 || sidx >= proof::TOTAL_SAMPLES) return 0u;
        u32 tr, c, m = 0;
        if (sidx < 50) { tr = 0; c = sidx; }
        else if (sidx < 110) { tr = 1; c = sidx - 50; }
        else if (sidx < 134) {
            const u32 r = sidx - 110;
            tr = 2;
            if (r < 4) c = r;
            else if (r < 12) { c = 4 + (r - 4) / 2; m = (r - 4) & 1u; }
            else if (r < 16) c = 8 + (r - 12);
            else { c = 12 + (r - 16) / 2; m = (r - 16) & 1u; }
        } else { tr = 3; c = sidx - 134; }
        return w[proof::sample_off(d, tr, c, m) + (l & 3u)];
    };
    {
        constexpr u32 n_pairs = (proof::TOTAL_SAMPLES + 1) / 2;
        u32 cur = pair_word(0);
        for (u32 j = 0; j < n_pairs; j++) {
            const u32 nxt = j + 1 < n_pairs ? pair_word(j + 1) : 0u;
            ch.absorb(cur);
            cur = nxt;
        }
    }
    t = ch.draw(); store_q(&o.after_coeff, t, 0);
    {
        u32 cur = l < 8 ? w[d.fl_commitment + l] : 0u;
        for (u32 i = 0; i <= n_inner; i++) {
            const u32 nxt = (i < n_inner && l < 8) ? w[d.in_commitment[i] + l] : 0u;
            ch.absorb(cur);
            t = ch.draw(); store_q(&o.fri_alphas[i], t, 0);
            cur = nxt;
        }
    }
    {
        const u32 n_mix = (n_last_coeffs + 1) / 2;
        auto coeff_word = [&](u32 i) -> u32 { return (l < 8 && 8 * i + l < 4 * n_last_coeffs) ? w[d.last_coeffs + 8 * i + l] : 0u; };
        u32 cur = coeff_word(0);
        for (u32 i = 0; i < n_mix; i++) {
            const u32 nxt = i + 1 < n_mix ? coeff_word(i + 1) : 0u;
            ch.absorb(cur);
            cur = nxt;
        }
    }
    const u64 nonce = (u64)w[d.pow_nonce] | ((u64)w[d.pow_nonce + 1] << 32);
    const u32 limb = l == 0 ? (u32)(nonce & ((1u << 22) - 1)) : l == 1 ? (u32)((nonce >> 22) & ((1u << 21) - 1)) : l == 2 ? (u32)((nonce >> 43) & ((1u << 21) - 1)) : 0u;
    ch.mix4v(limb);
    if (act && l >= 8) o.digest_after_nonce[l - 8] = ch.s;
    if (act && l == 8) o.pow_ok = (ch.s & ((1u << pow_bits) - 1)) == 0;
    u32 got = 0;
    for (u32 k = 0; k < (n_queries + 3) / 4; k++) {
        t = ch.draw();
        if (act && l < 8 && got + l < n_queries) o.raw_queries[got + l] = t;
        got += 8;
    }
    if (act && l == 0) o.n_transcript_perms = ch.n_perms;
}
__global__ void __launch_bounds__(kT) k_transcript16(const Workspace ws, u32 p0, u32 pn) {
    const u32 grp = (blockIdx.x * kT + threadIdx.x) / 16;
    const bool own = grp < pn && ws.desc[p0 + grp].ok;
    const bool other = __shfl_xor_sync(0xffffffffu, own ? 1u : 0u, 16) != 0;
    if (!own && !other) return;                              // whole warps leave together
    const u32 p = p0 + (own ? grp : (grp ^ 1u));             // a half without a proof shadows its neighbour (same shape, same trip counts)
    const proof::Desc &d = ws.desc[p];
    Lanes16 g;
    g.l = threadIdx.x % 16;
    g.active = own;
    g.init();
    verify::Detail &dt = ws.detail[p];
    transcript16(g, ws.blob(p), d, dt.fs, ws.perm_out_of(p, 0), ws.in_delta(), ws.shape.n_inner, 1u << ws.shape.log_last, ws.shape.n_queries,
                 ws.shape.pow_bits);
    __syncwarp();
    if (own && g.l == 0) verify::stage_after_transcript(ws, p);
}
__global__ void __launch_bounds__(kT) k_transcript(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn) verify::stage_transcript(ws, p0 + idx);
}
__global__ void __launch_bounds__(kT) k_oods(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn) verify::stage_oods(ws, p0 + idx);
}
template <int G>
__global__ void __launch_bounds__(kCoopThreads) k_folds_coop(const Workspace ws, u32 p0, u32 pn, u32 tab_words) {
    extern __shared__ u32 smem[];
    const u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (grp >= pn) return;
    CoopGroup<G> co;
    co.l = threadIdx.x % G;
    co.mask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (G * ((threadIdx.x % 32) / G));
    verify::stage_folds_coop(co, ws, p0 + grp, smem + (threadIdx.x / G) * tab_words);
}

__global__ void __launch_bounds__(kT) k_single_path(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    const u32 per_t = pn * ws.shape.n_queries;
    if (idx < per_t * 4) {
        u32 t = idx / per_t, r = idx % per_t;
        verify::stage_single_path(ws, p0 + r / ws.shape.n_queries, t, r % ws.shape.n_queries);
    }
}
__global__ void __launch_bounds__(kT) k_pair_path(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    const u32 per_f = pn * ws.shape.n_queries;
    if (idx < per_f * ws.shape.n_fri_trees()) {
        u32 f = idx / per_f, r = idx % per_f;
        verify::stage_pair_path(ws, p0 + r / ws.shape.n_queries, f, r % ws.shape.n_queries);
    }
}
__global__ void __launch_bounds__(kT) k_verdict(const Workspace ws, u32 p0, u32 pn, uint8_t *verdict, uint8_t *stage) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx >= pn) return;
    const u32 p = p0 + idx;
    verify::stage_verdict(ws, p);
    if (verdict) verdict[p] = (uint8_t)ws.detail[p].verdict;
    if (stage) stage[p] = (uint8_t)ws.detail[p].stage;
}

__global__ void __launch_bounds__(kT) k_synth_open(const Workspace ws, u32 p0, u32 pn) {
    u32 idx = blockIdx.x * kT + threadIdx.x;
    if (idx < pn) synth::stage_open(ws, p0 + idx);
}
// stream pools for sliced batches: per slice a main chain and a side stream for what does not gate the FRI chain.  One pool per CALLER
// stream (up to kPools; later callers share the last one): passes a caller issues on different streams -- the shape groups of a mixed
// batch, the lanes of a pipeline -- then really run beside each other instead of queueing on one set of pooled streams.
constexpr int kSlices = 4, kPools = 6;
struct Pool {
    cudaStream_t owner = nullptr, s[2 * kSlices] = {nullptr};
    cudaEvent_t fork = nullptr, fs_done[kSlices] = {nullptr}, tree_done[kSlices] = {nullptr}, side_done[kSlices] = {nullptr}, join[kSlices] = {nullptr};
    bool ready = false;
    bool init(cudaStream_t st) {
        for (auto &x : s) if (cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return false;
        for (int i = 0; i < kSlices; i++)
            for (cudaEvent_t *e : {&fs_done[i], &tree_done[i], &side_done[i], &join[i]})
                if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return false;
        owner = st; ready = true;
        return true;
    }
};
Pool g_pools[kPools];
void pools_destroy() {
    for (Pool &p : g_pools) {
        if (!p.ready) continue;
        for (auto &x : p.s) if (x) cudaStreamDestroy(x);
        if (p.fork) cudaEventDestroy(p.fork);
        for (int i = 0; i < kSlices; i++)
            for (cudaEvent_t e : {p.fs_done[i], p.tree_done[i], p.side_done[i], p.join[i]}) if (e) cudaEventDestroy(e);
        p = Pool();
    }
}
Pool *pool_for(cudaStream_t st) {
    static int n_pools = 0;                                // STWO_B200_VERIFY_POOLS=k (1..6): profiling
    if (!n_pools) { const char *e = getenv("STWO_B200_VERIFY_POOLS"); n_pools = e ? atoi(e) : kPools; if (n_pools < 1 || n_pools > kPools) n_pools = kPools; }
    for (int i = 0; i < n_pools; i++) {
        if (g_pools[i].ready && g_pools[i].owner == st) return &g_pools[i];
        if (!g_pools[i].ready) return g_pools[i].init(st) ? &g_pools[i] : nullptr;
    }
    return &g_pools[n_pools - 1];
}

// group width of the tree-rebuild kernels: 0 = one thread per tree (decommit.cuh); STWO_B200_TREE_G overrides the choice
int tree_group_width(u32 n_proofs) {
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("STWO_B200_TREE_G");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) return forced;
    return n_proofs <= 1024 ? 16 : 8;
}
// ... and for the stages whose unit of work is one QUERY (a node of a tree layer per query, a fold chain per query): a whole warp per
// unit when the shape has more queries than 32 lanes would leave idle and the batch is small -- the 80-query shapes of a mixed batch are
// a latency chain (34 proofs: 2.2 + 2.8 + 1.0 ms in the two tree stages and the folds with 16 lanes)
int query_group_width(u32 n_proofs, u32 n_queries) {
    const int g = tree_group_width(n_proofs);
    return (g == 16 && n_queries > 32 && n_proofs <= 256) ? 32 : g;
}
// shared memory of a block = (threads / G) group tables; large query counts need the opt-in limit
constexpr size_t kCoopSmemMax = 160 * 1024;
template <class K>
bool coop_launch(K kernel, int G, size_t groups, u32 tab_words, const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    int threads = kCoopThreads;
    while (threads > 32 && (size_t)(threads / G) * tab_words * 4 > kCoopSmemMax) threads /= 2;
    const size_t smem = (size_t)(threads / G) * tab_words * 4;
    if (smem > kCoopSmemMax) return false;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmemMax) != cudaSuccess) return false;
    kernel<<<(unsigned)((groups * G + threads - 1) / threads), threads, smem, st>>>(ws, p0, n, tab_words);
    return true;
}
int tree_min_blocks() {
    static int mb = -1;
    if (mb < 0) { const char *e = getenv("STWO_B200_TREE_BLOCKS"); const int v = e ? atoi(e) : kTreeBlocksDefault; mb = v >= 10 ? 10 : v >= 8 ? 8 : 1; }
    return mb;
}
template <int G>
bool launch_single_tree_g(size_t groups, u32 tab, const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    switch (tree_min_blocks()) {
    case 10: return coop_launch(k_single_tree_coop<G, 10>, G, groups, tab, ws, p0, n, st);
    case 8: return coop_launch(k_single_tree_coop<G, 8>, G, groups, tab, ws, p0, n, st);
    default: return coop_launch(k_single_tree_coop<G, 1>, G, groups, tab, ws, p0, n, st);
    }
}
template <int G>
bool launch_pair_tree_g(size_t groups, u32 tab, const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    switch (tree_min_blocks()) {
    case 10: return coop_launch(k_pair_tree_coop<G, 10>, G, groups, tab, ws, p0, n, st);
    case 8: return coop_launch(k_pair_tree_coop<G, 8>, G, groups, tab, ws, p0, n, st);
    default: return coop_launch(k_pair_tree_coop<G, 1>, G, groups, tab, ws, p0, n, st);
    }
}
void launch_single_tree(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    const int G = query_group_width(ws.n_proofs, ws.shape.n_queries);
    const u32 nq = ws.shape.n_queries, tab = decommit::single_tab_words(nq) + nq;
    const size_t groups = (size_t)n * 4;
    bool done = false;
    if (G == 32) done = launch_single_tree_g<32>(groups, tab, ws, p0, n, st);
    else if (G == 16) done = launch_single_tree_g<16>(groups, tab, ws, p0, n, st);
    else if (G == 8) done = launch_single_tree_g<8>(groups, tab, ws, p0, n, st);
    else if (G == 4) done = launch_single_tree_g<4>(groups, tab, ws, p0, n, st);
    if (!done) k_single_tree<<<(unsigned)((groups + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}
void launch_pair_tree(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    const int G = query_group_width(ws.n_proofs, ws.shape.n_queries);
    const u32 nq = ws.shape.n_queries, tab = decommit::pair_tab_words(nq) + nq;
    const size_t groups = (size_t)n * ws.shape.n_fri_trees();
    bool done = false;
    if (G == 32) done = launch_pair_tree_g<32>(groups, tab, ws, p0, n, st);
    else if (G == 16) done = launch_pair_tree_g<16>(groups, tab, ws, p0, n, st);
    else if (G == 8) done = launch_pair_tree_g<8>(groups, tab, ws, p0, n, st);
    else if (G == 4) done = launch_pair_tree_g<4>(groups, tab, ws, p0, n, st);
    if (!done) k_pair_tree<<<(unsigned)((groups + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}

// parse (a warp per proof shares the canonicity walk) + transcript: what gates every later stage
void launch_parse_transcript(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    if (tree_group_width(ws.n_proofs) == 0) { k_fiat_shamir<<<(unsigned)((n + kT - 1) / kT), kT, 0, st>>>(ws, p0, n); return; }
    coop_launch(k_parse_coop<32>, 32, n, verify::parse_tab_words(), ws, p0, n, st);
    static int lanes16 = -1;
    if (lanes16 < 0) { const char *e = getenv("STWO_B200_TRANSCRIPT_LANES"); lanes16 = (e && atoi(e) == 1) ? 0 : 1; }   // "1": thread-local state
    if (lanes16) k_transcript16<<<(unsigned)(((size_t)n * 16 + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
    else k_transcript<<<(unsigned)((n + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}
void launch_oods(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    if (tree_group_width(ws.n_proofs) == 0) return;                  // k_fiat_shamir already did it
    k_oods<<<(unsigned)((n + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}
void launch_folds(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    // one query per lane: 16 lanes per proof (8 when the batch alone fills the GPU)
    const int G = query_group_width(ws.n_proofs, ws.shape.n_queries);
    const u32 tab = verify::folds_tab_words(ws.shape.n_queries);
    bool done = false;
    if (G == 32) done = coop_launch(k_folds_coop<32>, 32, n, tab, ws, p0, n, st);
    else if (G == 16) done = coop_launch(k_folds_coop<16>, 16, n, tab, ws, p0, n, st);
    else if (G == 8 || G == 4) done = coop_launch(k_folds_coop<8>, 8, n, tab, ws, p0, n, st);
    if (!done) k_folds<<<(unsigned)((n + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}

void launch_group(const Workspace &ws, u32 p0, u32 n, cudaStream_t st) {
    if (tree_group_width(ws.n_proofs) == 0) k_group<<<(unsigned)(((size_t)n * fri::MAX_LOGS + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
    else k_group_coop<<<(unsigned)(((size_t)n * fri::MAX_LOGS * 16 + kT - 1) / kT), kT, 0, st>>>(ws, p0, n);
}
cudaEvent_t g_ev[STWO_B200_N_STAGE_KERNELS + 1] = {nullptr};
bool g_timed_valid = false;
inline unsigned nblk(size_t n) { return (unsigned)((n + kT - 1) / kT); }

bool shape_ok(const stwo_b200_proof_shape *s) {
    return s && proof::shape_consistent(s->log_size_plonk, s->log_size_poseidon, s->pow_bits, s->log_blowup, s->log_last, s->n_queries, s->n_inner);
}
}  // namespace

void stwo_b200::verify_pools_destroy() { pools_destroy(); }
static_assert(sizeof(stwo_b200_proof_shape) == sizeof(verify::Shape), "shape mirrors");
static_assert(sizeof(stwo_b200_verify_detail) == sizeof(verify::Detail), "detail mirrors");

extern "C" int32_t stwo_b200_proof_shape_of(const uint8_t *blob, size_t len, stwo_b200_proof_shape *out) {
    if (!blob || !out || (len & 3) || ((uintptr_t)blob & 3)) return STWO_B200_E_BAD_ARG;
    static thread_local proof::Desc d;
    if (!proof::parse(reinterpret_cast<const u32 *>(blob), len / 4, d)) return STWO_B200_E_SHAPE;
    out->log_size_plonk = d.log_size_plonk; out->log_size_poseidon = d.log_size_poseidon; out->pow_bits = d.pow_bits;
    out->log_blowup = d.log_blowup; out->log_last = d.log_last; out->n_queries = d.n_queries; out->n_inner = d.n_inner;
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_shape_from_config(const stwo_b200_pcs_config *config, uint32_t log_size_plonk, uint32_t log_size_poseidon,
                                               stwo_b200_proof_shape *out) {
    if (!config || !out) return STWO_B200_E_BAD_ARG;
    if (log_size_plonk == 0 || log_size_plonk > proof::MAX_LOG_SIZE || log_size_poseidon == 0 || log_size_poseidon > proof::MAX_LOG_SIZE ||
        config->log_blowup == 0 || config->log_blowup > proof::MAX_LOG_BLOWUP || config->log_last > proof::MAX_LOG_LAST)
        return STWO_B200_E_SHAPE;
    const uint32_t max_first = proof::expected_max_first(log_size_plonk, log_size_poseidon, config->log_blowup);
    if (max_first < config->log_last + config->log_blowup + 1) return STWO_B200_E_SHAPE;
    *out = {log_size_plonk, log_size_poseidon, config->pow_bits, config->log_blowup, config->log_last, config->n_queries,
            max_first - (config->log_last + config->log_blowup + 1)};
    return shape_ok(out) ? STWO_B200_OK : STWO_B200_E_SHAPE;
}

extern "C" size_t stwo_b200_verify_workspace_bytes(const stwo_b200_proof_shape *shape, uint32_t n_proofs) {
    if (!shape_ok(shape)) return 0;
    Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape, sizeof ws.shape);
    ws.n_proofs = n_proofs;
    return verify::carve(ws, nullptr);
}

extern "C" uint32_t stwo_b200_proof_record_slots(const stwo_b200_proof_shape *shape) {
    if (!shape_ok(shape)) return 0;
    verify::Shape v;
    memcpy(&v, shape, sizeof v);
    return verify::hint_layout(v).total;
}

extern "C" uint64_t stwo_b200_proof_perms(const stwo_b200_proof_shape *shape) {
    // permutations of the per-query (circuit) path of one proof, excluding the transcript: trees 0-3 + FRI layers
    if (!shape_ok(shape)) return 0;
    verify::Shape v;
    memcpy(&v, shape, sizeof v);
    uint64_t n = 0;
    for (u32 t = 0; t < 4; t++) {
        stwo_b200_path_shape ps;
        ps.depth = v.tree_depth(t);
        for (u32 h = 0; h <= ps.depth; h++) ps.n_cols[h] = 0;
        if (t < 3) { ps.n_cols[v.log_plonk()] += proof::plonk_cols(t); ps.n_cols[v.log_pos()] += proof::n_cols(t) - proof::plonk_cols(t); }
        else ps.n_cols[v.max_first()] = 8;
        n += (uint64_t)merkle::path_perms(ps) * v.n_queries;
    }
    for (u32 f = 0; f < v.n_fri_trees(); f++) {
        u32 depth = v.fri_depth(f), mask = v.fri_data_mask(f), per = 4;
        for (u32 h = 0; h < depth; h++) per += ((mask >> h) & 1u) ? (h >= 1 ? 5 : 3) : 1;
        n += (uint64_t)per * v.n_queries;
    }
    return n;
}

static int32_t verify_batch_impl(const uint32_t *host_blobs, const uint64_t *host_blob_off, const uint32_t *blobs, const uint64_t *blob_off,
                                 uint32_t n_proofs, const stwo_b200_proof_shape *shape, const uint32_t *input_idx,
                                 const uint32_t *input_vals, uint32_t n_inputs, uint32_t flags, void *workspace, size_t workspace_bytes,
                                 uint8_t *verdict, uint8_t *stage, void *stream);
extern "C" int32_t stwo_b200_verify_proofs_batch_dev(const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                                     const stwo_b200_proof_shape *shape, const uint32_t *input_idx,
                                                     const uint32_t *input_vals, uint32_t n_inputs, uint32_t flags,
                                                     void *workspace, size_t workspace_bytes, uint8_t *verdict, uint8_t *stage,
                                                     void *stream) {
    return verify_batch_impl(nullptr, nullptr, blobs, blob_off, n_proofs, shape, input_idx, input_vals, n_inputs, flags, workspace,
                             workspace_bytes, verdict, stage, stream);
}
extern "C" int32_t stwo_b200_verify_proofs_batch_pinned_dev(const uint32_t *host_blobs, const uint64_t *host_blob_off, uint32_t *blobs,
                                                            uint64_t *blob_off, uint32_t n_proofs, const stwo_b200_proof_shape *shape,
                                                            const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                                            uint32_t flags, void *workspace, size_t workspace_bytes, uint8_t *verdict,
                                                            uint8_t *stage, void *stream) {
    if (!host_blobs || !host_blob_off) return STWO_B200_E_BAD_ARG;
    return verify_batch_impl(host_blobs, host_blob_off, blobs, blob_off, n_proofs, shape, input_idx, input_vals, n_inputs, flags, workspace,
                             workspace_bytes, verdict, stage, stream);
}
static int32_t verify_batch_impl(const uint32_t *host_blobs, const uint64_t *host_blob_off, const uint32_t *blobs, const uint64_t *blob_off,
                                 uint32_t n_proofs, const stwo_b200_proof_shape *shape, const uint32_t *input_idx,
                                 const uint32_t *input_vals, uint32_t n_inputs, uint32_t flags, void *workspace, size_t workspace_bytes,
                                 uint8_t *verdict, uint8_t *stage, void *stream) {
    STWO_CHECK_DEVICE();
    if (n_proofs == 0) return STWO_B200_OK;
    if (!shape_ok(shape)) return STWO_B200_E_SHAPE;
    if (!blobs || !blob_off || !workspace || (n_inputs && (!input_idx || !input_vals))) return STWO_B200_E_BAD_ARG;
    Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape, sizeof ws.shape);
    ws.n_proofs = n_proofs; ws.blobs = blobs; ws.blob_off = blob_off;
    ws.input_idx = input_idx; ws.input_vals = input_vals; ws.n_inputs = n_inputs;
    if (verify::carve(ws, (uint8_t *)workspace) > workspace_bytes) return STWO_B200_E_BAD_ARG;
    if (!stwo_b200::record_inputs()) ws.perm_in = nullptr;         // STWO_B200_RECORD_INPUTS=0 (profiling): outputs only, the check re-executes
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nq = shape->n_queries, nf = ws.shape.n_fri_trees();
    const bool timed = flags & STWO_B200_VERIFY_TIMED, full = flags & STWO_B200_VERIFY_FULL;
    // Full mode: the cooperative tree rebuilds hash every node once and hand its permutation states to the queries whose path
    // runs through it.  The thread-per-path kernels (every path hashed again from its hints) remain as the checker of that
    // record (STWO_B200_VERIFY_PATH_KERNELS) and for the thread-per-tree rebuilds (STWO_B200_TREE_G=0), which keep no record.
    const bool path_kernels = full && ((flags & STWO_B200_VERIFY_PATH_KERNELS) || tree_group_width(n_proofs) == 0);
    ws.mode = (full ? verify::MODE_FULL : 0u) | (path_kernels ? verify::MODE_PATH_KERNELS : 0u);
    g_timed_valid = false;
    if (host_blobs)      // the offsets are needed by every slice: they go first, on the caller's stream
        STWO_CUDA(cudaMemcpyAsync(const_cast<uint64_t *>(blob_off), host_blob_off, ((size_t)n_proofs + 1) * 8, cudaMemcpyHostToDevice, st));
    const u32 upto = flags & (STWO_B200_VERIFY_UPTO_TRANSCRIPT | STWO_B200_VERIFY_UPTO_ANSWERS | STWO_B200_VERIFY_UPTO_FOLDS);
    if (timed || n_proofs < 256 || upto || (flags & STWO_B200_VERIFY_ONE_STREAM)) {
        if (host_blobs)
            STWO_CUDA(cudaMemcpyAsync(const_cast<uint32_t *>(blobs), host_blobs, host_blob_off[n_proofs] * 4, cudaMemcpyHostToDevice, st));
        // one stream, stage after stage (clean per-stage timings; small batches)
        const u32 n = n_proofs;
        if (timed && !g_ev[0]) for (int i = 0; i <= STWO_B200_N_STAGE_KERNELS; i++) STWO_CUDA(cudaEventCreate(&g_ev[i]));
        int e = 0;
#define MARK() do { if (timed) cudaEventRecord(g_ev[e], st); e++; } while (0)
        // stage-level callers (stwo_b200_channel_replay_batch / _fri_answers_batch / _fri_fold_batch) stop early; the verdict kernel
        // then reports what failed so far
        auto stop = [&](int launched) {
            k_verdict<<<nblk(n), kT, 0, st>>>(ws, 0, n, verdict, stage);
            note_launch(launched + 1);
            return cuda_status(cudaGetLastError());
        };
        MARK(); launch_parse_transcript(ws, 0, n, st); launch_oods(ws, 0, n, st);
        if (upto & STWO_B200_VERIFY_UPTO_TRANSCRIPT) return stop(3);
        MARK(); launch_single_tree(ws, 0, n, st);
        MARK(); launch_group(ws, 0, n, st);
        MARK(); k_answer<<<nblk((size_t)n * fri::MAX_LOGS * nq), kT, 0, st>>>(ws, 0, n);
        if (upto & STWO_B200_VERIFY_UPTO_ANSWERS) return stop(6);
        MARK(); launch_folds(ws, 0, n, st);
        if (upto & STWO_B200_VERIFY_UPTO_FOLDS) return stop(7);
        MARK(); launch_pair_tree(ws, 0, n, st);
        MARK();
        if (path_kernels) k_single_path<<<nblk((size_t)n * 4 * nq), kT, 0, st>>>(ws, 0, n);
        MARK();
        if (path_kernels) k_pair_path<<<nblk((size_t)n * nf * nq), kT, 0, st>>>(ws, 0, n);
        MARK(); k_verdict<<<nblk(n), kT, 0, st>>>(ws, 0, n, verdict, stage);
        MARK();
#undef MARK
        g_timed_valid = timed;
        note_launch((path_kernels ? 9 : 7) + (tree_group_width(ws.n_proofs) ? 2 : 0));      // parse + transcript + oods instead of one kernel
        return cuda_status(cudaGetLastError());
    }
    // sliced: the per-proof and per-tree stages have far fewer threads than the GPU holds, so independent slices of the
    // batch run concurrently on pooled streams, and the commitment-tree path recomputation runs beside the FRI chain
    Pool *pool = pool_for(st);
    if (!pool) return cuda_status(cudaGetLastError());
    STWO_CUDA(cudaEventRecord(pool->fork, st));
    for (int sl = 0; sl < kSlices; sl++) {
        const u32 p0 = (u32)((uint64_t)n_proofs * sl / kSlices), p1 = (u32)((uint64_t)n_proofs * (sl + 1) / kSlices), n = p1 - p0;
        cudaStream_t a = pool->s[2 * sl], b = pool->s[2 * sl + 1];
        STWO_CUDA(cudaStreamWaitEvent(a, pool->fork, 0));
        if (host_blobs && n)     // this slice's blobs, on the stream that consumes them: overlaps the other slices' kernels
            STWO_CUDA(cudaMemcpyAsync(const_cast<uint32_t *>(blobs) + host_blob_off[p0], host_blobs + host_blob_off[p0],
                                      (host_blob_off[p1] - host_blob_off[p0]) * 4, cudaMemcpyHostToDevice, a));
        launch_parse_transcript(ws, p0, n, a);
        // the side stream takes what does not gate the FRI chain: the OODS / logup check, then the commitment-tree paths
        STWO_CUDA(cudaEventRecord(pool->fs_done[sl], a));
        STWO_CUDA(cudaStreamWaitEvent(b, pool->fs_done[sl], 0));
        launch_oods(ws, p0, n, b);
        launch_single_tree(ws, p0, n, a);
        if (path_kernels) {
            STWO_CUDA(cudaEventRecord(pool->tree_done[sl], a));
            STWO_CUDA(cudaStreamWaitEvent(b, pool->tree_done[sl], 0));
            k_single_path<<<nblk((size_t)n * 4 * nq), kT, 0, b>>>(ws, p0, n);
        }
        STWO_CUDA(cudaEventRecord(pool->side_done[sl], b));
        launch_group(ws, p0, n, a);
        k_answer<<<nblk((size_t)n * fri::MAX_LOGS * nq), kT, 0, a>>>(ws, p0, n);
        launch_folds(ws, p0, n, a);
        launch_pair_tree(ws, p0, n, a);
        if (path_kernels) k_pair_path<<<nblk((size_t)n * nf * nq), kT, 0, a>>>(ws, p0, n);
        STWO_CUDA(cudaStreamWaitEvent(a, pool->side_done[sl], 0));
        k_verdict<<<nblk(n), kT, 0, a>>>(ws, p0, n, verdict, stage);
        STWO_CUDA(cudaEventRecord(pool->join[sl], a));
        STWO_CUDA(cudaStreamWaitEvent(st, pool->join[sl], 0));
        note_launch((path_kernels ? 9 : 7) + (tree_group_width(ws.n_proofs) ? 2 : 0));      // parse + transcript + oods instead of one kernel
    }
    return cuda_status(cudaGetLastError());
}

// Verification of synthetic FRI + Merkle instances (synth.cuh; BASELINE configs[4] part i): Fiat-Shamir over the FRI commitments and the
// opened first-layer values (k_synth_open), then the same fold and FRI tree-rebuild kernels a real proof goes through (K3 channel, K5
// folds, K2 Merkle).  FULL: the per-query roots and the permutation record of the FRI trees are produced like for a real proof.
extern "C" int32_t stwo_b200_synth_verify_batch_dev(const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                                    const stwo_b200_proof_shape *shape, uint32_t flags, void *workspace, size_t workspace_bytes,
                                                    uint8_t *verdict, uint8_t *stage, void *stream) {
    STWO_CHECK_DEVICE();
    if (n_proofs == 0) return STWO_B200_OK;
    if (!shape_ok(shape)) return STWO_B200_E_SHAPE;
    if (!blobs || !blob_off || !workspace) return STWO_B200_E_BAD_ARG;
    Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape, sizeof ws.shape);
    ws.n_proofs = n_proofs; ws.blobs = blobs; ws.blob_off = blob_off;
    if (verify::carve(ws, (uint8_t *)workspace) > workspace_bytes) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool full = flags & STWO_B200_VERIFY_FULL;
    const bool path_kernels = full && ((flags & STWO_B200_VERIFY_PATH_KERNELS) || tree_group_width(n_proofs) == 0);
    ws.mode = (full ? verify::MODE_FULL : 0u) | (path_kernels ? verify::MODE_PATH_KERNELS : 0u);
    const u32 n = n_proofs;
    const size_t nq = shape->n_queries, nf = ws.shape.n_fri_trees();
    k_synth_open<<<nblk(n), kT, 0, st>>>(ws, 0, n);
    launch_folds(ws, 0, n, st);
    launch_pair_tree(ws, 0, n, st);
    if (path_kernels) k_pair_path<<<nblk((size_t)n * nf * nq), kT, 0, st>>>(ws, 0, n);
    k_verdict<<<nblk(n), kT, 0, st>>>(ws, 0, n, verdict, stage);
    note_launch(path_kernels ? 5 : 4);
    return cuda_status(cudaGetLastError());
}

extern "C" int32_t stwo_b200_verify_stage_ms(float *ms) {
    STWO_CHECK_DEVICE();
    if (!ms || !g_timed_valid) return STWO_B200_E_BAD_ARG;
    STWO_CUDA(cudaEventSynchronize(g_ev[STWO_B200_N_STAGE_KERNELS]));
    for (int i = 0; i < STWO_B200_N_STAGE_KERNELS; i++) STWO_CUDA(cudaEventElapsedTime(&ms[i], g_ev[i], g_ev[i + 1]));
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_verify_fetch(const void *workspace, const stwo_b200_proof_shape *shape, uint32_t n_proofs, uint32_t p,
                                          uint32_t what, void *out, size_t out_bytes, void *stream) {
    STWO_CHECK_DEVICE();
    if (!shape_ok(shape) || p >= n_proofs || !workspace || !out) return STWO_B200_E_BAD_ARG;
    Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape, sizeof ws.shape);
    ws.n_proofs = n_proofs;
    verify::carve(ws, (uint8_t *)const_cast<void *>(workspace));
    const size_t nq = shape->n_queries, nf = ws.shape.n_fri_trees(), nt = ws.shape.n_trees();
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case STWO_B200_FETCH_DETAIL: src = ws.detail + p; bytes = sizeof(verify::Detail); break;
        case STWO_B200_FETCH_DOMAIN_POINTS: src = ws.domain_points + (size_t)p * fri::MAX_LOGS * nq * 2; bytes = fri::MAX_LOGS * nq * 8; break;
        case STWO_B200_FETCH_ANSWERS: src = ws.answers + (size_t)p * fri::MAX_LOGS * nq * 4; bytes = fri::MAX_LOGS * nq * 16; break;
        case STWO_B200_FETCH_CIRCLE_FOLDS: src = ws.circle_folds + (size_t)p * fri::MAX_LOGS * nq * 4; bytes = fri::MAX_LOGS * nq * 16; break;
        case STWO_B200_FETCH_LINE_FOLDS: src = ws.line_folds + (size_t)p * proof::MAX_INNER * nq * 4; bytes = proof::MAX_INNER * nq * 16; break;
        case STWO_B200_FETCH_LAST_EVALS: src = ws.last_evals + (size_t)p * nq * 4; bytes = nq * 16; break;
        case STWO_B200_FETCH_PATH_ROOTS: src = ws.path_roots + (size_t)p * nt * nq * 8; bytes = nt * nq * 32; break;
        case STWO_B200_FETCH_PATH_COLS: src = ws.path_cols + (size_t)p * 4 * nq * verify::PATH_COLS_STRIDE; bytes = 4 * nq * verify::PATH_COLS_STRIDE * 4; break;
        case STWO_B200_FETCH_PATH_SIBLINGS: src = ws.path_sib + (size_t)p * 4 * nq * verify::MAX_DEPTH * 8; bytes = 4 * nq * verify::MAX_DEPTH * 32; break;
        case STWO_B200_FETCH_PAIR_HINTS: src = ws.pair_hints + (size_t)p * nf * nq * verify::PAIR_HINT_WORDS; bytes = nf * nq * verify::PAIR_HINT_WORDS * 4; break;
        case STWO_B200_FETCH_PERM_RECORD: src = ws.perm_out_of(p, 0); bytes = (size_t)ws.hint_total * 64; break;
        case STWO_B200_FETCH_PERM_RECORD_INPUTS: src = ws.perm_out_of(p, 0) + ws.in_delta(); bytes = (size_t)ws.hint_total * 64; break;
        case STWO_B200_FETCH_RECORD_TREES: src = ws.hint_trees + p; bytes = 4; break;
        default: return STWO_B200_E_BAD_ARG;
    }
    if (out_bytes < bytes) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    STWO_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

// All proofs of a batch at once: the per-proof arrays of the workspace are contiguous, so a stage's results come back with one copy
extern "C" int32_t stwo_b200_verify_fetch_batch(const void *workspace, const stwo_b200_proof_shape *shape, uint32_t n_proofs, uint32_t what,
                                                void *out, size_t out_bytes, void *stream) {
    STWO_CHECK_DEVICE();
    if (!shape_ok(shape) || !n_proofs || !workspace || !out) return STWO_B200_E_BAD_ARG;
    Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape, sizeof ws.shape);
    ws.n_proofs = n_proofs;
    verify::carve(ws, (uint8_t *)const_cast<void *>(workspace));
    const size_t nq = shape->n_queries, n = n_proofs;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case STWO_B200_FETCH_DETAIL: src = ws.detail; bytes = n * sizeof(verify::Detail); break;
        case STWO_B200_FETCH_DOMAIN_POINTS: src = ws.domain_points; bytes = n * fri::MAX_LOGS * nq * 8; break;
        case STWO_B200_FETCH_ANSWERS: src = ws.answers; bytes = n * fri::MAX_LOGS * nq * 16; break;
        case STWO_B200_FETCH_CIRCLE_FOLDS: src = ws.circle_folds; bytes = n * fri::MAX_LOGS * nq * 16; break;
        case STWO_B200_FETCH_LINE_FOLDS: src = ws.line_folds; bytes = n * proof::MAX_INNER * nq * 16; break;
        case STWO_B200_FETCH_LAST_EVALS: src = ws.last_evals; bytes = n * nq * 16; break;
        default: return STWO_B200_E_BAD_ARG;
    }
    if (out_bytes < bytes) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    STWO_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

// Host-pointer batch entry: blobs may have different shapes; they are grouped by shape and verified one group at a time.
extern "C" int32_t stwo_b200_verify_proofs_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs,
                                                 const stwo_b200_pcs_config *configs, uint32_t n_configs,
                                                 const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                                 uint32_t flags, uint8_t *verdict, uint8_t *stage) {
    STWO_CHECK_DEVICE();
    if (n_proofs == 0) return STWO_B200_OK;
    // the PcsConfig is the CALLER's (FiatShamirHints::new(&proof, config, ..), components/hints/src/fiat_shamir.rs:69-73): a blob
    // is never verified under parameters it chose itself
    if (!blobs || !lens || !verdict || !configs || n_configs == 0) return STWO_B200_E_BAD_ARG;
    std::map<std::array<uint32_t, 7>, std::vector<uint32_t>> groups;
    for (uint32_t i = 0; i < n_proofs; i++) {
        stwo_b200_proof_shape s;
        // header-only shape read; malformed blobs are rejected here exactly as the device parse would
        bool allowed = false;
        if (blobs[i] && stwo_b200_proof_shape_of(blobs[i], lens[i], &s) == STWO_B200_OK)
            for (uint32_t c = 0; c < n_configs && !allowed; c++)
                allowed = configs[c].pow_bits == s.pow_bits && configs[c].log_blowup == s.log_blowup && configs[c].log_last == s.log_last &&
                          configs[c].n_queries == s.n_queries;
        if (!allowed) {                 // malformed, or a PcsConfig the caller did not ask for
            verdict[i] = proof::REJECT;
            if (stage) stage[i] = proof::ST_PARSE;
            continue;
        }
        std::array<uint32_t, 7> k;
        memcpy(k.data(), &s, sizeof s);
        groups[k].push_back(i);
    }
    // every shape group gets its own region of the staging area and one of a few streams: the groups are small and latency
    // bound (a transcript is a sequential chain), so they run beside each other; one synchronisation at the end
    struct Plan { stwo_b200_proof_shape s; const std::vector<uint32_t> *ids; std::vector<uint64_t> off; size_t base, b_blobs, b_off, b_in, b_ws; size_t out_at; uint8_t *d_verdict; cudaStream_t st; };
    std::vector<Plan> plans;
    size_t total = 0, out_total = 0;
    for (auto &kv : groups) {
        Plan pl;
        memcpy(&pl.s, kv.first.data(), sizeof pl.s);
        pl.ids = &kv.second;
        const uint32_t n = (uint32_t)kv.second.size();
        pl.off.assign(n + 1, 0);
        for (uint32_t k = 0; k < n; k++) pl.off[k + 1] = pl.off[k] + lens[kv.second[k]] / 4;
        pl.b_blobs = align_up(pl.off[n] * 4, 256); pl.b_off = align_up((n + 1) * 8, 256); pl.b_in = align_up((size_t)n_inputs * 20 + 4, 256);
        pl.b_ws = align_up(stwo_b200_verify_workspace_bytes(&pl.s, n), 256);
        pl.base = total; pl.out_at = out_total;
        total += pl.b_blobs + pl.b_off + pl.b_in + pl.b_ws + align_up(2 * (size_t)n, 256);
        out_total += 2 * (size_t)n;
        plans.push_back(std::move(pl));
    }
    int32_t rc = stage_reserve(total);
    if (rc) return rc;
    constexpr int kGroupStreams = 4;
    static cudaStream_t gs[kGroupStreams] = {nullptr};
    if (!gs[0]) for (int i = 0; i < kGroupStreams; i++) STWO_CUDA(cudaStreamCreateWithFlags(&gs[i], cudaStreamNonBlocking));
    // nothing of an earlier call on the staging stream may still be using the area
    STWO_CUDA(cudaStreamSynchronize(stage_stream()));
    std::vector<uint8_t> out(out_total);
    // largest groups first: the long chains start early and the short ones fill in beside them
    std::vector<size_t> order(plans.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return plans[x].b_ws > plans[y].b_ws; });
    int next = 0;
    for (size_t oi : order) {
        Plan &pl = plans[oi];
        cudaStream_t st = gs[next++ % kGroupStreams];
        const std::vector<uint32_t> &ids = *pl.ids;
        const uint32_t n = (uint32_t)ids.size();
        uint8_t *d = stage_dev() + pl.base;
        u32 *d_blobs = (u32 *)d; d += pl.b_blobs;
        u64 *d_off = (u64 *)d; d += pl.b_off;
        u32 *d_idx = (u32 *)d, *d_vals = d_idx + n_inputs; d += pl.b_in;
        uint8_t *d_ws = d; d += pl.b_ws;
        uint8_t *d_verdict = d, *d_stage = d + n;
        for (uint32_t k = 0; k < n; k++)
            STWO_CUDA(cudaMemcpyAsync(d_blobs + pl.off[k], blobs[ids[k]], lens[ids[k]], cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d_off, pl.off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        if (n_inputs) {
            STWO_CUDA(cudaMemcpyAsync(d_idx, input_idx, n_inputs * 4, cudaMemcpyHostToDevice, st));
            STWO_CUDA(cudaMemcpyAsync(d_vals, input_vals, n_inputs * 16, cudaMemcpyHostToDevice, st));
        }
        rc = stwo_b200_verify_proofs_batch_dev(d_blobs, d_off, n, &pl.s, d_idx, d_vals, n_inputs, flags, d_ws, pl.b_ws, d_verdict, d_stage, st);
        if (rc) return rc;
        pl.d_verdict = d_verdict; pl.st = st;
    }
    // the read-backs go last: a copy into pageable host memory blocks the host until the group has finished
    for (size_t oi : order) {
        Plan &pl = plans[oi];
        STWO_CUDA(cudaMemcpyAsync(out.data() + pl.out_at, pl.d_verdict, 2 * pl.ids->size(), cudaMemcpyDeviceToHost, pl.st));
    }
    for (int i = 0; i < kGroupStreams; i++) STWO_CUDA(cudaStreamSynchronize(gs[i]));
    for (const Plan &pl : plans) {
        const uint32_t n = (uint32_t)pl.ids->size();
        for (uint32_t k = 0; k < n; k++) {
            verdict[(*pl.ids)[k]] = out[pl.out_at + k];
            if (stage) stage[(*pl.ids)[k]] = out[pl.out_at + n + k];
        }
    }
    return STWO_B200_OK;
}
