// Proof wire format on device: bincode-1.3 PlonkWithPoseidonProof<Poseidon31MerkleHasher> blobs are
// consumed in place (little-endian u32 words); one pass per proof records where every section starts.
//
// Struct usage in the reference: components/recursive/data_structures/src/lib.rs:98-223 (proof vars),
// components/hints/src/fiat_shamir.rs:69-216 (field order of the stwo types), SURVEY.md App. A (layout).
// Everything here is HD so the CPU test tier runs the same code (tests/hostsim).
#pragma once
#include "m31.cuh"

namespace proof {

constexpr u32 MAX_QUERIES = 128;
constexpr u32 MAX_INNER = 32;
constexpr u32 MAX_LOG_SIZE = 28;        // component log sizes (stmt0) are untrusted header words: bounded BEFORE any sum is formed
constexpr u32 MAX_LOG_BLOWUP = 16;
constexpr u32 MAX_LOG_LAST = 12;
constexpr u32 MAX_FIRST = 29;           // largest committed column log size

// The FRI layer count is not free: the composition columns are committed at log size composition_log_degree_bound - 1 +
// blowup, where the bound is the largest constraint degree bound of the two components (Plonk: degree 3 -> log_size + 2,
// Poseidon: is_full * (x + rc)^5 -> log_size + 3; components/hints/src/fiat_shamir.rs:130-135), and stwo's
// FriVerifier::commit (called at fiat_shamir.rs:177-183) rejects a proof whose inner-layer count does not bring that
// size down to log_last + blowup (FriVerificationError::InvalidNumFriLayers).  All 16 fixtures satisfy the relation.
HD u32 expected_max_first(u32 log_size_plonk, u32 log_size_poseidon, u32 log_blowup) {
    const u32 a = log_size_plonk + 1, b = log_size_poseidon + 2;
    return (a > b ? a : b) + log_blowup;
}
// every bound and relation a (header, PcsConfig) pair must satisfy; no sum can wrap once the three bounds hold
HD bool shape_consistent(u32 log_size_plonk, u32 log_size_poseidon, u32 pow_bits, u32 log_blowup, u32 log_last, u64 n_queries, u64 n_inner) {
    if (log_size_plonk == 0 || log_size_plonk > MAX_LOG_SIZE || log_size_poseidon == 0 || log_size_poseidon > MAX_LOG_SIZE) return false;
    if (log_blowup == 0 || log_blowup > MAX_LOG_BLOWUP || log_last > MAX_LOG_LAST || pow_bits >= 32) return false;
    if (n_queries == 0 || n_queries > MAX_QUERIES || n_inner >= MAX_INNER) return false;
    const u32 max_first = log_last + log_blowup + 1 + (u32)n_inner;
    return max_first <= MAX_FIRST && max_first == expected_max_first(log_size_plonk, log_size_poseidon, log_blowup);
}
// columns per commitment tree (preprocessed, trace, interaction, composition) and how many of the leading ones belong
// to the Plonk component (the rest to the Poseidon component)
HD u32 n_cols(u32 t) { return t == 0 ? 50u : t == 1 ? 60u : t == 2 ? 16u : 8u; }
HD u32 plonk_cols(u32 t) { return t == 0 ? 10u : t == 1 ? 12u : t == 2 ? 8u : 0u; }
constexpr u32 TOTAL_SAMPLES = 50 + 60 + 8 + 2 * 8 + 8;

// verdict / stage codes shared with the C ABI (include/stwo_b200.h)
enum { ACCEPT = 0, REJECT = 1, UNSUPPORTED = 2 };
enum { ST_OK = 0, ST_PARSE = 1, ST_POW = 2, ST_LOGUP = 3, ST_OODS = 4, ST_MERKLE = 5, ST_FRI_FIRST = 6, ST_FRI_INNER = 7,
       ST_FRI_LAST = 8, ST_UNSUPPORTED = 9 };

struct Desc {                    // offsets are in u32 words from the start of this proof's blob
    u32 ok;
    u32 log_size_plonk, log_size_poseidon, pow_bits, log_blowup, log_last, n_queries, n_inner;
    u32 max_first, log_plonk, log_pos;          // derived: column log sizes incl. blow-up
    u32 stmt1;
    u32 commitments[4];
    u32 sampled[4];                             // first column record of each tree
    u32 hash_witness[4], n_hash_witness[4];
    u32 queried[4], n_queried[4];
    u32 pow_nonce;
    u32 fl_fri_witness, fl_n_fri_witness, fl_hash_witness, fl_n_hash_witness, fl_commitment;
    u32 in_fri_witness[MAX_INNER], in_n_fri_witness[MAX_INNER], in_hash_witness[MAX_INNER],
        in_n_hash_witness[MAX_INNER], in_commitment[MAX_INNER];
    u32 last_coeffs, n_last_coeffs;
};

// number of mask values of column c of tree t: the last logup batch of each component is sampled at [-1, 0]
// (components/recursive/composition/src/data_structures.rs:189-207)
HD u32 n_masks(u32 t, u32 c) { return (t == 2 && (c & 4)) ? 2u : 1u; }
// word offset of sample m of column c of tree t (column record = u64 length + n_masks QM31)
HD u32 sample_off(const Desc &d, u32 t, u32 c, u32 m) {
    u32 off = d.sampled[t];
    if (t == 2) off += (c < 4 ? c * 6 : c < 8 ? 24 + (c - 4) * 10 : c < 12 ? 64 + (c - 8) * 6 : 88 + (c - 12) * 10);
    else off += c * 6;
    return off + 2 + 4 * m;
}

struct Reader {
    const u32 *w; size_t n, at; bool bad;
    // deferred canonicity check: when `ranges` is set, words() records (offset, count) pairs there instead of walking the
    // range, and a group of lanes checks them afterwards (verify::stage_parse_coop); a full list falls back to the walk
    u32 *ranges = nullptr; u32 n_ranges = 0, max_ranges = 0;
    HDM u32 r32() { if (bad || at + 1 > n) { bad = true; return 0; } return w[at++]; }
    HDM u64 r64() { if (bad || at + 2 > n) { bad = true; return 0; } u64 v = (u64)w[at] | ((u64)w[at + 1] << 32); at += 2; return v; }
    // n_words canonical M31 words; returns their offset
    HDM u32 words(u64 n_words) {
        if (bad || n_words > n - at) { bad = true; return 0; }
        u32 off = (u32)at;
        if (ranges && n_ranges < max_ranges) { ranges[2 * n_ranges] = off; ranges[2 * n_ranges + 1] = (u32)n_words; n_ranges++; }
        else for (u64 i = 0; i < n_words; i++) if (w[at + i] >= M31_P) bad = true;
        at += (size_t)n_words;
        return off;
    }
    HDM void decommitment(u32 &hw, u32 &n_hw) {
        u64 n = r64();
        if (n > (1u << 24)) { bad = true; return; }
        n_hw = (u32)n;
        hw = words(n * 8);
        if (r64() != 0) bad = true;          // column_witness must be empty (components/hints/src/decommit.rs:71)
    }
};

// Returns true when the blob is a well-formed proof of a supported shape.
HD bool parse(const u32 *w, size_t n_words, Desc &d, u32 *ranges = nullptr, u32 max_ranges = 0, u32 *n_ranges_out = nullptr) {
    Reader r{w, n_words, 0, false};
    r.ranges = ranges; r.max_ranges = max_ranges;
    struct Done { Reader &r; u32 *out; HDM ~Done() { if (out) *out = r.n_ranges; } } done{r, n_ranges_out};
    d.ok = 0;
    d.log_size_plonk = r.r32();
    d.log_size_poseidon = r.r32();
    d.stmt1 = r.words(8);
    d.pow_bits = r.r32();
    d.log_blowup = r.r32();
    d.log_last = r.r32();
    u64 nq = r.r64();
    // header words are attacker-controlled: bound each one before it is shifted by, looped over or added to anything
    if (r.bad || nq == 0 || nq > MAX_QUERIES || d.pow_bits >= 32 || d.log_last > MAX_LOG_LAST) return false;
    if (d.log_size_plonk == 0 || d.log_size_plonk > MAX_LOG_SIZE || d.log_size_poseidon == 0 || d.log_size_poseidon > MAX_LOG_SIZE ||
        d.log_blowup == 0 || d.log_blowup > MAX_LOG_BLOWUP)
        return false;
    d.n_queries = (u32)nq;
    if (r.r64() != 4) return false;
    for (int t = 0; t < 4; t++) d.commitments[t] = r.words(8);
    if (r.r64() != 4) return false;
    for (u32 t = 0; t < 4; t++) {
        if (r.r64() != n_cols(t)) return false;
        d.sampled[t] = (u32)r.at;
        for (u32 c = 0; c < n_cols(t); c++) {
            if (r.r64() != n_masks(t, c)) return false;
            r.words(4 * n_masks(t, c));
        }
    }
    if (r.r64() != 4) return false;
    for (int t = 0; t < 4; t++) r.decommitment(d.hash_witness[t], d.n_hash_witness[t]);
    if (r.r64() != 4) return false;
    for (int t = 0; t < 4; t++) {
        u64 n = r.r64();
        if (n > (1u << 24)) return false;
        d.n_queried[t] = (u32)n;
        d.queried[t] = r.words(n);
    }
    d.pow_nonce = (u32)r.at;
    r.r64();
    {
        u64 n = r.r64();
        if (n > (1u << 24)) return false;
        d.fl_n_fri_witness = (u32)n;
        d.fl_fri_witness = r.words(n * 4);
        r.decommitment(d.fl_hash_witness, d.fl_n_hash_witness);
        d.fl_commitment = r.words(8);
    }
    u64 ni = r.r64();
    if (r.bad || ni >= MAX_INNER) return false;
    d.n_inner = (u32)ni;
    for (u32 i = 0; i < d.n_inner; i++) {
        u64 n = r.r64();
        if (n > (1u << 24)) return false;
        d.in_n_fri_witness[i] = (u32)n;
        d.in_fri_witness[i] = r.words(n * 4);
        r.decommitment(d.in_hash_witness[i], d.in_n_hash_witness[i]);
        d.in_commitment[i] = r.words(8);
    }
    u64 nl = r.r64();
    if (r.bad || nl != (1ull << d.log_last)) return false;
    d.n_last_coeffs = (u32)nl;
    d.last_coeffs = r.words(nl * 4);
    r.r32();                                   // last_layer_poly.log_size
    if (r.bad || r.at != n_words) return false;
    d.max_first = d.log_last + d.log_blowup + 1 + d.n_inner;
    d.log_plonk = d.log_size_plonk + d.log_blowup;
    d.log_pos = d.log_size_poseidon + d.log_blowup;
    if (!shape_consistent(d.log_size_plonk, d.log_size_poseidon, d.pow_bits, d.log_blowup, d.log_last, d.n_queries, d.n_inner)) return false;
    d.ok = 1;
    return true;
}

}  // namespace proof
