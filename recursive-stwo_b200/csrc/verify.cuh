// Batched native verification of PlonkWithPoseidon stwo proofs: workspace layout and the per-thread
// stage functions the kernels in verify_kernels.cu dispatch.  Proofs of one batch share a shape
// (PcsConfig + component log sizes), so every per-proof array has a fixed stride.
//
// Stage order and failure codes follow the reference driver examples/single-proof/src/main.rs:33-83:
//   FiatShamirHints/Results -> CompositionCheck -> DecommitHints + AnswerResults -> FirstLayerHints,
//   InnerLayersHints + FoldingResults.
#pragma once
#include "fri.cuh"
#include "merkle.cuh"
#include "decommit_coop.cuh"

namespace verify {

using proof::Desc;

struct Shape {            // mirror of stwo_b200_proof_shape (include/stwo_b200.h)
    u32 log_size_plonk, log_size_poseidon, pow_bits, log_blowup, log_last, n_queries, n_inner;
    HDM u32 max_first() const { return log_last + log_blowup + 1 + n_inner; }
    HDM u32 log_plonk() const { return log_size_plonk + log_blowup; }
    HDM u32 log_pos() const { return log_size_poseidon + log_blowup; }
    HDM u32 tree_depth(u32 t) const { return t == 3 ? max_first() : (log_plonk() > log_pos() ? log_plonk() : log_pos()); }
    HDM u32 tree_cols(u32 t) const { return proof::n_cols(t); }
    HDM u32 n_fri_trees() const { return 1 + n_inner; }
    HDM u32 fri_depth(u32 f) const { return f == 0 ? max_first() : max_first() - f; }
    HDM u32 fri_data_mask(u32 f) const {
        if (f) return 1u << fri_depth(f);
        return (1u << max_first()) | (1u << log_plonk()) | (1u << log_pos());
    }
    HDM u32 n_trees() const { return 4 + n_fri_trees(); }
    HDM bool matches(const Desc &d) const {
        return d.log_size_plonk == log_size_plonk && d.log_size_poseidon == log_size_poseidon && d.pow_bits == pow_bits &&
               d.log_blowup == log_blowup && d.log_last == log_last && d.n_queries == n_queries && d.n_inner == n_inner;
    }
};

// Where the native verifier leaves the output state of the permutations that the verifier CIRCUIT executes as well (the
// transcript and every per-query authentication path: together exactly the circuit's Poseidon flow).  K6 may take them as hints
// instead of permuting again (tape::Perm::hint); check_poseidon_invocations still re-executes every flow entry.
// Slot order inside a path = execution order of merkle::path_root / decommit::pair_path_root.
constexpr u32 HINT_TRANSCRIPT_SLOTS = 512;
struct HintLayout {
    u32 single_base[4], single_n[4];
    u32 pair_base[proof::MAX_INNER + 1], pair_n[proof::MAX_INNER + 1];
    u32 total;
    HDM u32 single_slot(u32 t, u32 i) const { return single_base[t] + i * single_n[t]; }
    HDM u32 pair_slot(u32 f, u32 i) const { return pair_base[f] + i * pair_n[f]; }
};
HD void single_path_shape(const Shape &s, u32 t, stwo_b200_path_shape &shp) {
    const u32 depth = s.tree_depth(t);
    shp.depth = depth;
    for (u32 h = 0; h <= depth; h++) shp.n_cols[h] = 0;
    if (t < 3) { shp.n_cols[s.log_plonk()] += proof::plonk_cols(t); shp.n_cols[s.log_pos()] += proof::n_cols(t) - proof::plonk_cols(t); }
    else shp.n_cols[s.max_first()] = 8;
}
HD HintLayout hint_layout(const Shape &s) {
    HintLayout h;
    u32 at = HINT_TRANSCRIPT_SLOTS;
    for (u32 t = 0; t < 4; t++) {
        stwo_b200_path_shape shp;
        single_path_shape(s, t, shp);
        h.single_n[t] = merkle::path_perms(shp);
        h.single_base[t] = at;
        at += h.single_n[t] * s.n_queries;
    }
    for (u32 f = 0; f <= proof::MAX_INNER; f++) { h.pair_base[f] = at; h.pair_n[f] = 0; }
    for (u32 f = 0; f < s.n_fri_trees(); f++) {
        h.pair_n[f] = decommit::pair_path_perms(s.fri_depth(f), s.fri_data_mask(f));
        h.pair_base[f] = at;
        at += h.pair_n[f] * s.n_queries;
    }
    h.total = at;
    return h;
}

constexpr u32 MAX_DEPTH = 30;
constexpr u32 PATH_COLS_STRIDE = 64;                           // >= widest tree (60 columns)
constexpr u32 PAIR_HINT_WORDS = 2 * decommit::MAX_DATA_LAYERS * 4 + (MAX_DEPTH - 1) * 8;   // self, sib, sibling hashes

// Per-proof results kept for the caller (every Fiat-Shamir draw, the OODS values, the counters).
struct Detail {
    fs::Out fs;
    qm31_t oods_computed, oods_expected;
    u32 n_logs, log_sizes[fri::MAX_LOGS];
    u32 fail_mask;                       // bit s set <=> stage s failed
    u32 verdict, stage;
    u32 n_perms_hints, n_perms_paths;
};

struct Workspace {
    Shape shape;
    u32 n_proofs;
    const u32 *blobs;                    // all proofs, word-addressed
    const u64 *blob_off;                 // n_proofs + 1 word offsets
    const u32 *input_idx; const u32 *input_vals; u32 n_inputs;
    Desc *desc;
    Detail *detail;
    // commitment trees: [p][t][i]
    u32 *path_cols;                      // PATH_COLS_STRIDE words
    u32 *path_sib;                       // MAX_DEPTH * 8 words
    u32 *path_roots;                     // [p][tree 0..n_trees)[i][8]   recomputed per-query roots (full mode)
    u32 *single_scratch;                 // [p][t][22 * nq]
    // answers / folds
    fri::Group *groups;                  // [p][g]
    u32 *domain_points;                  // [p][g][i][2]
    u32 *answers;                        // [p][g][i][4]
    u32 *circle_folds;                   // [p][g][i][4]
    u32 *line_folds;                     // [p][li][i][4]
    u32 *last_evals;                     // [p][i][4]
    u32 *fri_vals;                       // [p][f][2 * nq * MAX_LOGS... see fri_vals_stride][4]
    u32 *n_fri_vals;                     // [p][f]
    u32 *pair_hints;                     // [p][f][i][PAIR_HINT_WORDS]
    u32 *pair_scratch;                   // [p][f][88 * nq]
    u32 *fold_buf;                       // [p][max(1, 2^(log_last-1))][4]  last-layer polynomial fold buffer
    u32 *perm_out;                       // [p][HintLayout::total][16]  output state of every transcript / per-query path permutation
    u32 *perm_in;                        // [p][HintLayout::total][16]  its input state (same slots): what check_poseidon_invocations compares
                                         //      a flow entry with, instead of executing the permutation a second time
    u32 *hint_trees;                     // [p]  trees (of n_trees()) whose part of perm_out is complete for THIS run; the circuit's tape
                                         //      evaluation takes a proof's permutations from perm_out only when all of them are
    u32 mode;                            // MODE_* (set by the host per run)

    HDM const u32 *blob(u32 p) const { return blobs + blob_off[p]; }
    HDM size_t blob_words(u32 p) const { return (size_t)(blob_off[p + 1] - blob_off[p]); }
    HDM u32 nq() const { return shape.n_queries; }
    HDM u32 *cols_of(u32 p, u32 t, u32 i) const { return path_cols + (((size_t)p * 4 + t) * nq() + i) * PATH_COLS_STRIDE; }
    HDM u32 *sib_of(u32 p, u32 t, u32 i) const { return path_sib + (((size_t)p * 4 + t) * nq() + i) * (MAX_DEPTH * 8); }
    HDM u32 *root_of(u32 p, u32 tree, u32 i) const { return path_roots + (((size_t)p * shape.n_trees() + tree) * nq() + i) * 8; }
    HDM u32 fri_vals_stride() const { return 2 * nq() * fri::MAX_LOGS * 4; }
    HDM u32 *vals_of(u32 p, u32 f) const { return fri_vals + ((size_t)p * shape.n_fri_trees() + f) * fri_vals_stride(); }
    HDM u32 *nvals_of(u32 p, u32 f) const { return n_fri_vals + (size_t)p * shape.n_fri_trees() + f; }
    HDM u32 *hint_of(u32 p, u32 f, u32 i) const { return pair_hints + (((size_t)p * shape.n_fri_trees() + f) * nq() + i) * PAIR_HINT_WORDS; }
    // tree-walk scratch: thread (p, tree) owns word i at base[i * n_threads + thread] on the device (coalesced across
    // the lanes of a warp, which hold the same tree of different proofs); contiguous per thread on the host
#if defined(__CUDA_ARCH__)
    HDM decommit::Strided scratch_single(u32 p, u32 t) const { decommit::Strided s; s.p = single_scratch + ((size_t)t * n_proofs + p); s.stride = 4 * n_proofs; return s; }
    HDM decommit::Strided scratch_pair(u32 p, u32 f) const { decommit::Strided s; s.p = pair_scratch + ((size_t)f * n_proofs + p); s.stride = shape.n_fri_trees() * n_proofs; return s; }
#else
    HDM decommit::Strided scratch_single(u32 p, u32 t) const { decommit::Strided s; s.p = single_scratch + ((size_t)p * 4 + t) * decommit::SINGLE_SCRATCH_WORDS_PER_QUERY * nq(); s.stride = 1; return s; }
    HDM decommit::Strided scratch_pair(u32 p, u32 f) const { decommit::Strided s; s.p = pair_scratch + ((size_t)p * shape.n_fri_trees() + f) * decommit::PAIR_SCRATCH_WORDS_PER_QUERY * nq(); s.stride = 1; return s; }
#endif
    HDM u32 *q4(u32 *base, u32 p, u32 a, u32 na, u32 i) const { return base + (((size_t)p * na + a) * nq() + i) * 4; }
    u32 hint_total, pair_hint_base;      // HintLayout::total / pair_base[0] of `shape` (set by carve)
    HDM u32 *perm_out_of(u32 p, u32 slot) const { return perm_out ? perm_out + ((size_t)p * hint_total + slot) * 16 : nullptr; }
    HDM size_t in_delta() const { return perm_out && perm_in ? (size_t)(perm_in - perm_out) : 0; }   // input slot = output slot + in_delta words
};

// MODE_FULL: record the circuit's permutations (perm_out) and the per-query roots.  MODE_PATH_KERNELS: the record is produced by
// the thread-per-path kernels (stage_single_path / stage_pair_path: every path hashed again from its hints, as
// SinglePathMerkleProof::verify does) instead of by the cooperative tree rebuilds, which hash every node once.
enum : u32 { MODE_FULL = 1u, MODE_PATH_KERNELS = 2u };

HD void fail(Detail &dt, u32 stage) { dt.fail_mask |= 1u << stage; }
#if defined(__CUDA_ARCH__)
#define VERIFY_ATOMIC_OR(ptr, v) atomicOr((ptr), (v))
#define VERIFY_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define VERIFY_ATOMIC_OR(ptr, v) (*(ptr) |= (v))
#define VERIFY_ATOMIC_ADD(ptr, v) (*(ptr) += (v))
#endif
HD void fail_shared(Detail *dt, u32 stage) { VERIFY_ATOMIC_OR(&dt->fail_mask, 1u << stage); }

// ---- stage 1: parse + transcript + logup + OODS -------------------------------------------------------------------------
// Three pieces so that only what gates the rest (parse, transcript) sits on the critical path: the canonicity walk over the
// blob is shared by a group of lanes, and the OODS / logup check runs beside the tree rebuilds.
HD void reset_detail(Detail &dt) {
    dt.fail_mask = 0; dt.verdict = proof::REJECT; dt.stage = 0; dt.n_perms_hints = 0; dt.n_perms_paths = 0; dt.n_logs = 0;
}
// what follows the transcript proper: PoW verdict, log sizes, the query-collision check
HD void stage_after_transcript(const Workspace &ws, u32 p);
// transcript + PoW + the query-collision check (thread per proof)
HD void stage_transcript(const Workspace &ws, u32 p) {
    Desc &d = ws.desc[p];
    if (!d.ok) return;
    fs::transcript(ws.blob(p), d, ws.detail[p].fs, ws.perm_out_of(p, 0), ws.in_delta());
    stage_after_transcript(ws, p);
}
HD void stage_after_transcript(const Workspace &ws, u32 p) {
    Desc &d = ws.desc[p];
    Detail &dt = ws.detail[p];
    VERIFY_ATOMIC_ADD(&dt.n_perms_paths, dt.fs.n_transcript_perms);
    if (!dt.fs.pow_ok) fail_shared(&dt, proof::ST_POW);
    dt.n_logs = fri::log_sizes(d, dt.log_sizes);
    // duplicated queries at the largest size are not supported by the reference (answer/src/lib.rs:190-195)
    for (u32 i = 0; i < d.n_queries; i++)
        for (u32 j = 0; j < i; j++)
            if (fri::position(d, dt.fs.raw_queries[i], d.max_first) == fri::position(d, dt.fs.raw_queries[j], d.max_first)) {
                fail_shared(&dt, proof::ST_UNSUPPORTED);
                i = d.n_queries; break;
            }
}
// logup sum + OODS composition check (thread per proof; needs only the transcript)
HD void stage_oods(const Workspace &ws, u32 p) {
    Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    if (!fs::logup_sum_ok(w, d, dt.fs, ws.input_idx, ws.input_vals, ws.n_inputs)) fail_shared(&dt, proof::ST_LOGUP);
    if (!fs::oods_ok(w, d, dt.fs, &dt.oods_computed, &dt.oods_expected)) fail_shared(&dt, proof::ST_OODS);
}
HD void stage_fiat_shamir(const Workspace &ws, u32 p) {
    Desc &d = ws.desc[p];
    Detail &dt = ws.detail[p];
    reset_detail(dt);
    if (ws.hint_trees) ws.hint_trees[p] = 0;
    if (!proof::parse(ws.blob(p), ws.blob_words(p), d) || !ws.shape.matches(d)) { d.ok = 0; fail(dt, proof::ST_PARSE); return; }
    stage_transcript(ws, p);
    stage_oods(ws, p);
}
// cooperative parse: lane 0 walks the structure, every lane checks a share of the field-element ranges for canonicity.
// tab: group-shared, parse_tab_words() words.
constexpr u32 PARSE_MAX_RANGES = 320;
HD u32 parse_tab_words() { return 2 * PARSE_MAX_RANGES + 8; }
template <class Co>
HD void stage_parse_coop(const Co &co, const Workspace &ws, u32 p, u32 *tab) {
    Desc &d = ws.desc[p];
    Detail &dt = ws.detail[p];
    u32 *ranges = tab, *ctl = tab + 2 * PARSE_MAX_RANGES;      // [0] n_ranges  [1] structure ok  [2] non-canonical word seen
    const u32 *w = ws.blob(p);
    if (co.lane() == 0) {
        reset_detail(dt);
        if (ws.hint_trees) ws.hint_trees[p] = 0;
        u32 n_ranges = 0;
        const bool ok = proof::parse(w, ws.blob_words(p), d, ranges, PARSE_MAX_RANGES, &n_ranges) && ws.shape.matches(d);
        ctl[0] = n_ranges; ctl[1] = ok ? 1u : 0u; ctl[2] = 0;
    }
    co.sync();
    if (ctl[1]) {
        u32 bad = 0;
        for (u32 r = 0; r < ctl[0]; r++) {
            const u32 off = ranges[2 * r], len = ranges[2 * r + 1];
            for (u32 i = co.lane(); i < len; i += co.size()) bad |= (w[off + i] >= M31_P) ? 1u : 0u;
        }
        if (bad) ctl[2] = 1;
    }
    co.sync();
    if (co.lane() == 0 && (!ctl[1] || ctl[2])) { d.ok = 0; fail(dt, proof::ST_PARSE); }
    co.sync();
}

// ---- stage 2: commitment-tree decommitment -> per-query paths (one thread per (proof, tree)) -----------------------
HD void stage_single_tree(const Workspace &ws, u32 p, u32 t) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries;
    decommit::SingleShape sh;
    if (t < 3) {
        sh.hA = d.log_plonk; sh.nA = proof::plonk_cols(t); sh.hB = d.log_pos; sh.nB = proof::n_cols(t) - proof::plonk_cols(t);
        sh.depth = sh.hA > sh.hB ? sh.hA : sh.hB;
    } else { sh.hA = d.max_first; sh.nA = 8; sh.hB = 0; sh.nB = 0; sh.depth = d.max_first; }
    u32 q[proof::MAX_QUERIES], idx[decommit::SINGLE_IDX_WORDS_PER_QUERY * proof::MAX_QUERIES];
    for (u32 i = 0; i < nq; i++) q[i] = fri::position(d, dt.fs.raw_queries[i], sh.depth);
    u32 perms = 0;
    bool ok = decommit::single_tree(sh, q, nq, w + d.queried[t], d.n_queried[t], w + d.hash_witness[t], d.n_hash_witness[t],
                                    w + d.commitments[t], ws.cols_of(p, t, 0), PATH_COLS_STRIDE, ws.sib_of(p, t, 0), MAX_DEPTH * 8,
                                    ws.scratch_single(p, t), idx, &perms);
    VERIFY_ATOMIC_ADD(&dt.n_perms_hints, perms);
    if (!ok) fail_shared(&dt, proof::ST_MERKLE);
}

// cooperative form: a group of lanes per (proof, tree); tab = group-shared scratch of decommit::single_tab_words(nq) words.
// The node tables use this tree's slice of single_scratch contiguously (16 * nq words).
template <class Co>
HD void stage_single_tree_coop(const Co &co, const Workspace &ws, u32 p, u32 t, u32 *tab) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries;
    decommit::SingleShape sh;
    if (t < 3) {
        sh.hA = d.log_plonk; sh.nA = proof::plonk_cols(t); sh.hB = d.log_pos; sh.nB = proof::n_cols(t) - proof::plonk_cols(t);
        sh.depth = sh.hA > sh.hB ? sh.hA : sh.hB;
    } else { sh.hA = d.max_first; sh.nA = 8; sh.hB = 0; sh.nB = 0; sh.depth = d.max_first; }
    u32 *q = tab + decommit::single_tab_words(nq);                     // query positions, group-shared
    for (u32 i = co.lane(); i < nq; i += co.size()) q[i] = fri::position(d, dt.fs.raw_queries[i], sh.depth);
    co.sync();
    u32 perms = 0;
    u32 *nodes = ws.single_scratch + ((size_t)p * 4 + t) * decommit::SINGLE_SCRATCH_WORDS_PER_QUERY * nq;
    // full mode: the rebuild hands every node's permutation states to the queries whose path runs through the node
    decommit::PermRec rec{nullptr, 0, nullptr};
    const bool record = (ws.mode & MODE_FULL) && !(ws.mode & MODE_PATH_KERNELS) && ws.perm_out;
    if (record) {
        stwo_b200_path_shape shp;
        u32 at = HINT_TRANSCRIPT_SLOTS;
        for (u32 tt = 0; tt <= t; tt++) { single_path_shape(ws.shape, tt, shp); if (tt < t) at += merkle::path_perms(shp) * nq; }
        rec.base = ws.perm_out_of(p, at); rec.per_query = merkle::path_perms(shp); rec.roots = ws.root_of(p, t, 0); rec.in_delta = ws.in_delta();
    }
    const bool ok = decommit::single_tree_coop(co, sh, q, nq, w + d.queried[t], d.n_queried[t], w + d.hash_witness[t], d.n_hash_witness[t],
                                               w + d.commitments[t], ws.cols_of(p, t, 0), PATH_COLS_STRIDE, ws.sib_of(p, t, 0), MAX_DEPTH * 8,
                                               nodes, tab, &perms, rec);
    if (co.lane() == 0) {
        VERIFY_ATOMIC_ADD(&dt.n_perms_hints, perms);
        if (!ok) fail_shared(&dt, proof::ST_MERKLE);
        if (record && decommit::single_rec_complete(tab, nq)) {
            VERIFY_ATOMIC_ADD(&dt.n_perms_paths, rec.per_query * nq);  // the circuit's path permutations this record covers
            VERIFY_ATOMIC_ADD(&ws.hint_trees[p], 1u);
        }
    }
}
template <class Co>
HD void stage_pair_tree_coop(const Co &co, const Workspace &ws, u32 p, u32 f, u32 *tab) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries, depth = ws.shape.fri_depth(f);
    u32 *q = tab + decommit::pair_tab_words(nq);
    for (u32 i = co.lane(); i < nq; i += co.size()) q[i] = fri::position(d, dt.fs.raw_queries[i], depth);
    co.sync();
    const u32 *hw = w + (f ? d.in_hash_witness[f - 1] : d.fl_hash_witness);
    const u32 n_hw = f ? d.in_n_hash_witness[f - 1] : d.fl_n_hash_witness;
    const u32 *root = w + (f ? d.in_commitment[f - 1] : d.fl_commitment);
    u32 *hint = ws.hint_of(p, f, 0);
    u32 *self_vals = hint, *sib_vals = hint + nq * decommit::MAX_DATA_LAYERS * 4, *sib_hashes = sib_vals + nq * decommit::MAX_DATA_LAYERS * 4;
    u32 *nodes = ws.pair_scratch + ((size_t)p * ws.shape.n_fri_trees() + f) * decommit::PAIR_SCRATCH_WORDS_PER_QUERY * nq;
    u32 perms = 0;
    decommit::PermRec rec{nullptr, 0, nullptr};
    const bool record = (ws.mode & MODE_FULL) && !(ws.mode & MODE_PATH_KERNELS) && ws.perm_out;
    if (record) {
        u32 at = ws.pair_hint_base;
        for (u32 ff = 0; ff < f; ff++) at += decommit::pair_path_perms(ws.shape.fri_depth(ff), ws.shape.fri_data_mask(ff)) * nq;
        rec.base = ws.perm_out_of(p, at); rec.per_query = decommit::pair_path_perms(depth, ws.shape.fri_data_mask(f)); rec.roots = ws.root_of(p, 4 + f, 0);
        rec.in_delta = ws.in_delta();
    }
    const bool ok = decommit::pair_tree_coop(co, depth, ws.shape.fri_data_mask(f), q, nq, ws.vals_of(p, f), *ws.nvals_of(p, f), hw, n_hw, root,
                                             self_vals, sib_vals, sib_hashes, nodes, tab, &perms, rec);
    if (co.lane() == 0) {
        VERIFY_ATOMIC_ADD(&dt.n_perms_hints, perms);
        if (!ok) fail_shared(&dt, f ? proof::ST_FRI_INNER : proof::ST_FRI_FIRST);
        if (record && decommit::pair_rec_complete(tab, nq)) {
            VERIFY_ATOMIC_ADD(&dt.n_perms_paths, rec.per_query * nq);
            VERIFY_ATOMIC_ADD(&ws.hint_trees[p], 1u);
        }
    }
}

// full mode: SinglePathMerkleProofVar::verify for one (proof, tree, query)
HD void stage_single_path(const Workspace &ws, u32 p, u32 t, u32 i) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    stwo_b200_path_shape shp;
    const u32 depth = ws.shape.tree_depth(t);
    single_path_shape(ws.shape, t, shp);
    u32 root[8];
    u32 *sink = nullptr;
    if (ws.perm_out) {
        // slot base of (tree t, query i): the same prefix sums as hint_layout(), without building the whole table per thread
        u32 at = HINT_TRANSCRIPT_SLOTS;
        for (u32 tt = 0; tt < t; tt++) { stwo_b200_path_shape o; single_path_shape(ws.shape, tt, o); at += merkle::path_perms(o) * ws.nq(); }
        sink = ws.perm_out_of(p, at + i * merkle::path_perms(shp));
    }
    merkle::path_root(shp, fri::position(d, dt.fs.raw_queries[i], depth), ws.cols_of(p, t, i), ws.sib_of(p, t, i), root, sink, ws.in_delta());
    decommit::cp8(ws.root_of(p, t, i), root);
    VERIFY_ATOMIC_ADD(&dt.n_perms_paths, merkle::path_perms(shp));
    if (!decommit::eq8(root, ws.blob(p) + d.commitments[t])) fail_shared(&dt, proof::ST_MERKLE);
}

// ---- stage 3: answers (thread per (proof, group) then per (proof, group, query)) -------------------------------------
HD void stage_group(const Workspace &ws, u32 p, u32 g) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    if (g >= dt.n_logs) return;
    if (!fri::build_group(ws.blob(p), d, dt.fs, dt.log_sizes[g], ws.groups[(size_t)p * fri::MAX_LOGS + g])) fail_shared(&dt, proof::ST_PARSE);
}
template <class Co>
HD void stage_group_coop(const Co &co, const Workspace &ws, u32 p, u32 g) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    if (g >= dt.n_logs) return;
    if (!fri::build_group_coop(co, ws.blob(p), d, dt.fs, dt.log_sizes[g], ws.groups[(size_t)p * fri::MAX_LOGS + g]) && co.lane() == 0)
        fail_shared(&dt, proof::ST_PARSE);
}
HD void stage_answer(const Workspace &ws, u32 p, u32 g, u32 i) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    if (g >= dt.n_logs) return;
    const u32 L = dt.log_sizes[g];
    const u32 q = fri::position(d, dt.fs.raw_queries[i], L);
    const cpoint_t dp = fri::domain_point(L, q);
    u32 *dpo = ws.domain_points + (((size_t)p * fri::MAX_LOGS + g) * ws.nq() + i) * 2;
    dpo[0] = dp.x; dpo[1] = dp.y;
    u32 row[fri::MAX_GROUP_SAMPLES];
    const u32 *paths[4] = {ws.cols_of(p, 0, i), ws.cols_of(p, 1, i), ws.cols_of(p, 2, i), ws.cols_of(p, 3, i)};
    fri::gather_row(d, L, paths, row);
    qm31_t ans = qm31::zero();
    if (!fri::row_quotient(ws.groups[(size_t)p * fri::MAX_LOGS + g], row, dp, ans)) fail_shared(&dt, proof::ST_FRI_FIRST);
    fs::qstore(ws.q4(ws.answers, p, g, fri::MAX_LOGS, i), ans);
}

// ---- stage 4: folds (one thread per proof): first-layer evaluations, circle folds, line folds, last layer ---------------
HD void stage_folds(const Workspace &ws, u32 p) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries, max_first = d.max_first;
    u32 pos[proof::MAX_QUERIES], sp[proof::MAX_QUERIES];
    // first layer: per log size descending, per sorted pair subset, both evaluations (queried: answer, else fri_witness)
    {
        u32 *vals = ws.vals_of(p, 0), nv = 0, wi = 0;
        bool ok = true;
        for (u32 g = 0; g < dt.n_logs && ok; g++) {
            const u32 L = dt.log_sizes[g];
            for (u32 i = 0; i < nq; i++) sp[i] = pos[i] = fri::position(d, dt.fs.raw_queries[i], L);
            const u32 ns = decommit::sort_unique(sp, nq);
            for (u32 k = 0; k < ns && ok;) {
                const u32 start = sp[k] & ~1u;
                for (u32 e = start; e < start + 2; e++) {
                    const u32 *src;
                    if (k < ns && sp[k] == e) {
                        u32 i = 0;
                        while (pos[i] != e) i++;
                        src = ws.q4(ws.answers, p, g, fri::MAX_LOGS, i);
                        k++;
                    } else {
                        if (wi >= d.fl_n_fri_witness) { ok = false; break; }
                        src = w + d.fl_fri_witness + 4 * wi++;
                    }
                    for (int c = 0; c < 4; c++) vals[nv++] = src[c];
                }
            }
        }
        if (wi != d.fl_n_fri_witness) ok = false;
        *ws.nvals_of(p, 0) = nv;
        if (!ok) fail_shared(&dt, proof::ST_FRI_FIRST);
    }
    // circle folds: sibling evaluation of query i at log L = the other member of its pair in the first-layer values
    for (u32 g = 0; g < dt.n_logs; g++) {
        const u32 L = dt.log_sizes[g];
        for (u32 i = 0; i < nq; i++) pos[i] = fri::position(d, dt.fs.raw_queries[i], L);
        // locate each pair inside vals: pairs are stored in sorted order per group; recompute the offsets
        u32 base = 0;
        for (u32 g2 = 0; g2 < g; g2++) {
            for (u32 i = 0; i < nq; i++) sp[i] = fri::position(d, dt.fs.raw_queries[i], dt.log_sizes[g2]) >> 1;
            base += decommit::sort_unique(sp, nq) * 8;
        }
        for (u32 i = 0; i < nq; i++) sp[i] = pos[i] >> 1;
        const u32 npairs = decommit::sort_unique(sp, nq);
        const u32 *vals = ws.vals_of(p, 0);
        for (u32 i = 0; i < nq; i++) {
            const int pi = decommit::find(sp, npairs, pos[i] >> 1);
            const u32 *pr = vals + base + 8 * (u32)pi;
            const qm31_t l = fs::qload(pr), r = fs::qload(pr + 4);
            const qm31_t self = (pos[i] & 1u) ? r : l, sib = (pos[i] & 1u) ? l : r;
            const cpoint_t pt = circle::dbl(fri::absolute_point(L, pos[i]));
            const qm31_t f = fri::fold_pair(self, sib, pos[i], fs::minv(pt.y), dt.fs.fri_alphas[max_first - L]);
            fs::qstore(ws.q4(ws.circle_folds, p, g, fri::MAX_LOGS, i), f);
        }
    }
    // inner layers
    qm31_t folded[proof::MAX_QUERIES];
    for (u32 i = 0; i < nq; i++) folded[i] = qm31::zero();
    u32 log_size = max_first;
    bool inner_ok = true;
    for (u32 li = 0; li < d.n_inner; li++) {
        for (u32 g = 0; g < dt.n_logs; g++)
            if (dt.log_sizes[g] == log_size) {
                const qm31_t a2 = fs::qmul(dt.fs.fri_alphas[li], dt.fs.fri_alphas[li]);
                for (u32 i = 0; i < nq; i++) folded[i] = fs::qadd(fs::qmul(a2, folded[i]), fs::qload(ws.q4(ws.circle_folds, p, g, fri::MAX_LOGS, i)));
            }
        log_size -= 1;
        for (u32 i = 0; i < nq; i++) sp[i] = pos[i] = fri::position(d, dt.fs.raw_queries[i], log_size);
        const u32 ns = decommit::sort_unique(sp, nq);
        u32 *vals = ws.vals_of(p, 1 + li), nv = 0, wi = 0;
        // decommitted evaluations: one (left, right) pair per sorted pair index; a missing sibling comes from fri_witness
        qm31_t sib_of_unique[proof::MAX_QUERIES];
        for (u32 k = 0; k < ns; k++) {
            const u32 e = sp[k];
            u32 i = 0;
            while (pos[i] != e) i++;
            qm31_t sv;
            const int sk = decommit::find(sp, ns, e ^ 1u);
            if (sk >= 0) { u32 j = 0; while (pos[j] != (e ^ 1u)) j++; sv = folded[j]; }
            else if (wi < d.in_n_fri_witness[li]) sv = fs::qload(w + d.in_fri_witness[li] + 4 * wi++);
            else { inner_ok = false; sv = qm31::zero(); }
            sib_of_unique[k] = sv;
            if (k == 0 || (sp[k - 1] >> 1) != (e >> 1)) {
                const qm31_t l = (e & 1u) ? sv : folded[i], r = (e & 1u) ? folded[i] : sv;
                fs::qstore(vals + nv, l); fs::qstore(vals + nv + 4, r); nv += 8;
            }
        }
        if (wi != d.in_n_fri_witness[li]) inner_ok = false;
        *ws.nvals_of(p, 1 + li) = nv;
        for (u32 i = 0; i < nq; i++) {
            const int k = decommit::find(sp, ns, pos[i]);
            const u32 x_inv = fs::minv(fri::absolute_point(log_size, pos[i]).x);
            folded[i] = fri::fold_pair(folded[i], sib_of_unique[k], pos[i], x_inv, dt.fs.fri_alphas[li + 1]);
        }
        for (u32 i = 0; i < nq; i++) fs::qstore(ws.q4(ws.line_folds, p, li, proof::MAX_INNER, i), folded[i]);
    }
    if (!inner_ok) fail_shared(&dt, proof::ST_FRI_INNER);
    // last layer
    bool last_ok = true;
    const u32 fold_cap = d.log_last ? (1u << (d.log_last - 1)) : 1u;
    qm31_t *buf = (qm31_t *)(ws.fold_buf + (size_t)p * fold_cap * 4);
    for (u32 i = 0; i < nq; i++) {
        const cpoint_t ab = fri::absolute_point(log_size, fri::position(d, dt.fs.raw_queries[i], log_size));
        const u32 x = m31::subc(m31::mulc(ab.x, ab.x), m31::mulc(ab.y, ab.y));
        qm31_t ev;
        if (d.log_last == 0) ev = fs::qload(w + d.last_coeffs);
        else {
            u32 dbl[16];
            dbl[0] = x;
            for (u32 k = 1; k < d.log_last; k++) { u32 sq = m31::mulc(dbl[k - 1], dbl[k - 1]); dbl[k] = m31::subc(m31::addc(sq, sq), 1); }
            u32 n = 1u << d.log_last;
            {
                n >>= 1;
                for (u32 k = 0; k < n; k++)
                    buf[k] = fs::qadd(fs::qload(w + d.last_coeffs + 8 * k), qm31::mul_m31(fs::qload(w + d.last_coeffs + 8 * k + 4), dbl[d.log_last - 1]));
                for (u32 lev = d.log_last - 1; lev-- > 0;) {
                    n >>= 1;
                    for (u32 k = 0; k < n; k++) buf[k] = fs::qadd(buf[2 * k], qm31::mul_m31(buf[2 * k + 1], dbl[lev]));
                }
                ev = buf[0];
            }
        }
        fs::qstore(ws.last_evals + ((size_t)p * nq + i) * 4, ev);
        if (!qm31::eq(ev, folded[i])) last_ok = false;
    }
    if (!last_ok) fail_shared(&dt, proof::ST_FRI_LAST);
}

// Cooperative form of stage_folds: a group of lanes per proof, one query per lane for the arithmetic (domain points, M31
// inversions, folds, last-layer polynomial), lane 0 for the short stream-order bookkeeping (which evaluation of a pair comes
// from fri_witness).  tab: group-shared, folds_tab_words(nq) words.
HD u32 folds_tab_words(u32 nq) { return 3 * nq + 8 * nq + 8 + 4 * nq; }
// Sorted distinct values of pos[i] >> shift, by all lanes of the group: sp[0..ns) ascending, rep[k] = the first query that
// has value sp[k].  flag: nq scratch words.  Rank sort: every lane ranks its own queries against all the others.
template <class Co>
HD u32 coop_sort_unique(const Co &co, const u32 *pos, u32 nq, u32 shift, u32 *sp, u32 *rep, u32 *flag, u32 *count) {
    const u32 L = co.lane(), G = co.size();
    for (u32 i = L; i < nq; i += G) {
        const u32 v = pos[i] >> shift;
        u32 first = 1;
        for (u32 j = 0; j < i; j++) if ((pos[j] >> shift) == v) { first = 0; break; }
        flag[i] = first;
    }
    co.sync();
    for (u32 i = L; i < nq; i += G) {
        if (!flag[i]) continue;
        const u32 v = pos[i] >> shift;
        u32 rank = 0;
        for (u32 j = 0; j < nq; j++) rank += (flag[j] && (pos[j] >> shift) < v) ? 1u : 0u;
        sp[rank] = v; rep[rank] = i;
    }
    if (L == 0) {
        u32 c = 0;
        for (u32 j = 0; j < nq; j++) c += flag[j];
        *count = c;
    }
    co.sync();
    return *count;
}
template <class Co>
HD void stage_folds_coop(const Co &co, const Workspace &ws, u32 p, u32 *tab) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries, max_first = d.max_first, L = co.lane(), G = co.size();
    u32 *pos = tab, *sp = pos + nq;
    u32 *folded = sp + 2 * nq, *sibv = folded + 4 * nq;               // qm31 per query / per unique position
    u32 *ctl = sibv + 4 * nq;                                          // [0] ns  [1] ok flags
    u32 *rep = ctl + 8, *flag = rep + nq, *widx = flag + nq, *pidx = widx + nq;
    if (L == 0) ctl[1] = 0;
    // ---- first layer: per log size descending, per sorted pair subset, both evaluations
    {
        u32 *vals = ws.vals_of(p, 0);
        if (L == 0) { ctl[2] = 0; ctl[3] = 0; }                        // nv, wi
        co.sync();
        for (u32 g = 0; g < dt.n_logs; g++) {
            const u32 Lg = dt.log_sizes[g];
            for (u32 i = L; i < nq; i += G) pos[i] = fri::position(d, dt.fs.raw_queries[i], Lg);
            co.sync();
            const u32 ns = coop_sort_unique(co, pos, nq, 0, sp, rep, flag, ctl);
            if (L == 0) {
                u32 nv = ctl[2], wi = ctl[3];
                bool ok = true;
                for (u32 k = 0; k < ns && ok;) {
                    const u32 start = sp[k] & ~1u;
                    for (u32 e = start; e < start + 2; e++) {
                        const u32 *src;
                        if (k < ns && sp[k] == e) {
                            src = ws.q4(ws.answers, p, g, fri::MAX_LOGS, rep[k]);
                            k++;
                        } else {
                            if (wi >= d.fl_n_fri_witness) { ok = false; break; }
                            src = w + d.fl_fri_witness + 4 * wi++;
                        }
                        for (int c = 0; c < 4; c++) vals[nv++] = src[c];
                    }
                }
                if (!ok) ctl[1] |= 1;
                ctl[2] = nv; ctl[3] = wi;
            }
            co.sync();
        }
        if (L == 0) {
            if (ctl[3] != d.fl_n_fri_witness) ctl[1] |= 1;
            *ws.nvals_of(p, 0) = ctl[2];
            if (ctl[1] & 1) fail_shared(&dt, proof::ST_FRI_FIRST);
        }
        co.sync();
    }
    // ---- circle folds, one query per lane
    {
        u32 base = 0;
        const u32 *vals = ws.vals_of(p, 0);
        for (u32 g = 0; g < dt.n_logs; g++) {
            const u32 Lg = dt.log_sizes[g];
            for (u32 i = L; i < nq; i += G) pos[i] = fri::position(d, dt.fs.raw_queries[i], Lg);
            co.sync();
            const u32 npairs = coop_sort_unique(co, pos, nq, 1, sp, rep, flag, ctl);
            for (u32 i = L; i < nq; i += G) {
                const int pi = decommit::find(sp, npairs, pos[i] >> 1);
                const u32 *pr = vals + base + 8 * (u32)pi;
                const qm31_t l = fs::qload(pr), r = fs::qload(pr + 4);
                const qm31_t self = (pos[i] & 1u) ? r : l, sib = (pos[i] & 1u) ? l : r;
                const cpoint_t pt = circle::dbl(fri::absolute_point(Lg, pos[i]));
                const qm31_t f = fri::fold_pair(self, sib, pos[i], fs::minv(pt.y), dt.fs.fri_alphas[max_first - Lg]);
                fs::qstore(ws.q4(ws.circle_folds, p, g, fri::MAX_LOGS, i), f);
            }
            co.sync();
            base += npairs * 8;
        }
    }
    // ---- inner layers
    for (u32 i = L; i < nq; i += G) fs::qstore(folded + 4 * i, qm31::zero());
    co.sync();
    u32 log_size = max_first;
    for (u32 li = 0; li < d.n_inner; li++) {
        for (u32 g = 0; g < dt.n_logs; g++)
            if (dt.log_sizes[g] == log_size) {
                const qm31_t a2 = fs::qmul(dt.fs.fri_alphas[li], dt.fs.fri_alphas[li]);
                for (u32 i = L; i < nq; i += G)
                    fs::qstore(folded + 4 * i, fs::qadd(fs::qmul(a2, fs::qload(folded + 4 * i)), fs::qload(ws.q4(ws.circle_folds, p, g, fri::MAX_LOGS, i))));
            }
        log_size -= 1;
        for (u32 i = L; i < nq; i += G) pos[i] = fri::position(d, dt.fs.raw_queries[i], log_size);
        co.sync();
        const u32 ns = coop_sort_unique(co, pos, nq, 0, sp, rep, flag, ctl);
        // per distinct position k: does its sibling come from the witness (widx), does it open a new pair (pidx)?
        for (u32 k = L; k < ns; k += G) {
            widx[k] = decommit::find(sp, ns, sp[k] ^ 1u) < 0 ? 1u : 0u;
            pidx[k] = (k == 0 || (sp[k - 1] >> 1) != (sp[k] >> 1)) ? 1u : 0u;
        }
        co.sync();
        if (L == 0) {                                  // exclusive prefix sums; bit 31 keeps the flag
            u32 wi = 0, np = 0;
            for (u32 k = 0; k < ns; k++) {
                const u32 fw = widx[k], fp = pidx[k];
                widx[k] = wi | (fw << 31); pidx[k] = np | (fp << 31);
                wi += fw; np += fp;
            }
            if (wi != d.in_n_fri_witness[li]) ctl[1] |= 2;
            *ws.nvals_of(p, 1 + li) = 8 * np;
        }
        co.sync();
        {
            u32 *vals = ws.vals_of(p, 1 + li);
            for (u32 k = L; k < ns; k += G) {
                const u32 e = sp[k];
                qm31_t sv;
                if (widx[k] >> 31) {
                    const u32 wi = widx[k] & 0x7fffffffu;
                    sv = wi < d.in_n_fri_witness[li] ? fs::qload(w + d.in_fri_witness[li] + 4 * wi) : qm31::zero();
                } else sv = fs::qload(folded + 4 * rep[(u32)decommit::find(sp, ns, e ^ 1u)]);
                fs::qstore(sibv + 4 * k, sv);
                if (pidx[k] >> 31) {
                    const qm31_t fv = fs::qload(folded + 4 * rep[k]);
                    const qm31_t l = (e & 1u) ? sv : fv, r = (e & 1u) ? fv : sv;
                    u32 *out = vals + 8 * (pidx[k] & 0x7fffffffu);
                    fs::qstore(out, l); fs::qstore(out + 4, r);
                }
            }
        }
        co.sync();
        for (u32 i = L; i < nq; i += G) {
            const int k = decommit::find(sp, ns, pos[i]);
            const u32 x_inv = fs::minv(fri::absolute_point(log_size, pos[i]).x);
            const qm31_t f = fri::fold_pair(fs::qload(folded + 4 * i), fs::qload(sibv + 4 * (u32)k), pos[i], x_inv, dt.fs.fri_alphas[li + 1]);
            fs::qstore(ws.q4(ws.line_folds, p, li, proof::MAX_INNER, i), f);
            // lanes read folded[] of other queries only through sibv (filled before the barrier): writing now is safe
            fs::qstore(folded + 4 * i, f);
        }
        co.sync();
    }
    if (L == 0 && (ctl[1] & 2)) fail_shared(&dt, proof::ST_FRI_INNER);
    // ---- last layer, one query per lane (each lane folds the polynomial in its own slice of fold_buf when it is large)
    bool last_ok = true;
    for (u32 i = L; i < nq; i += G) {
        const cpoint_t ab = fri::absolute_point(log_size, fri::position(d, dt.fs.raw_queries[i], log_size));
        const u32 x = m31::subc(m31::mulc(ab.x, ab.x), m31::mulc(ab.y, ab.y));
        qm31_t ev;
        if (d.log_last == 0) ev = fs::qload(w + d.last_coeffs);
        else {
            // Horner-free evaluation of the fold: value = sum_k coeff[k] * prod_{bits b of k} dbl[log_last - 1 - b'] -- done
            // iteratively over the coefficients with a per-lane stack of partial sums (depth log_last), no shared buffer
            u32 dbl[16];
            dbl[0] = x;
            for (u32 k = 1; k < d.log_last; k++) { u32 sq = m31::mulc(dbl[k - 1], dbl[k - 1]); dbl[k] = m31::subc(m31::addc(sq, sq), 1); }
            qm31_t stack[16];
            u32 depth_used = 0;
            const u32 n = 1u << d.log_last;
            for (u32 k = 0; k < n; k++) {
                qm31_t cur = fs::qload(w + d.last_coeffs + 4 * k);
                // merge completed pairs: after coefficient k, every trailing 1-bit of k closes a level
                u32 kk = k, lev = d.log_last;
                while (kk & 1u) {
                    lev--;
                    cur = fs::qadd(stack[--depth_used], qm31::mul_m31(cur, dbl[lev]));
                    kk >>= 1;
                }
                stack[depth_used++] = cur;
            }
            ev = stack[0];
        }
        fs::qstore(ws.last_evals + ((size_t)p * nq + i) * 4, ev);
        if (!qm31::eq(ev, fs::qload(folded + 4 * i))) last_ok = false;
    }
    if (!last_ok) fail_shared(&dt, proof::ST_FRI_LAST);
}

// ---- stage 5: FRI layer trees (one thread per (proof, layer)) ---------------------------------------------------------
HD void stage_pair_tree(const Workspace &ws, u32 p, u32 f) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 *w = ws.blob(p);
    const u32 nq = d.n_queries, depth = ws.shape.fri_depth(f);
    u32 q[proof::MAX_QUERIES], idx[decommit::PAIR_IDX_WORDS_PER_QUERY * proof::MAX_QUERIES];
    for (u32 i = 0; i < nq; i++) q[i] = fri::position(d, dt.fs.raw_queries[i], depth);
    const u32 *hw = w + (f ? d.in_hash_witness[f - 1] : d.fl_hash_witness);
    const u32 n_hw = f ? d.in_n_hash_witness[f - 1] : d.fl_n_hash_witness;
    const u32 *root = w + (f ? d.in_commitment[f - 1] : d.fl_commitment);
    u32 *hint = ws.hint_of(p, f, 0);
    // pair_tree writes packed per-query arrays; give it strided views by running it on a packed scratch tail and spreading
    const decommit::Strided scratch = ws.scratch_pair(p, f);
    u32 perms = 0;
    // packed outputs live at the start of the hint block of this tree (nq * PAIR_HINT_WORDS words available)
    u32 *self_vals = hint, *sib_vals = hint + nq * decommit::MAX_DATA_LAYERS * 4, *sib_hashes = sib_vals + nq * decommit::MAX_DATA_LAYERS * 4;
    bool ok = decommit::pair_tree(depth, ws.shape.fri_data_mask(f), q, nq, ws.vals_of(p, f), *ws.nvals_of(p, f), hw, n_hw, root,
                                  self_vals, sib_vals, sib_hashes, scratch, idx, &perms);
    VERIFY_ATOMIC_ADD(&dt.n_perms_hints, perms);
    if (!ok) fail_shared(&dt, f ? proof::ST_FRI_INNER : proof::ST_FRI_FIRST);
}
// packed hint accessors (see stage_pair_tree)
HD const u32 *hint_self(const Workspace &ws, u32 p, u32 f, u32 i) { return ws.hint_of(p, f, 0) + i * decommit::MAX_DATA_LAYERS * 4; }
HD const u32 *hint_sib(const Workspace &ws, u32 p, u32 f, u32 i) {
    return ws.hint_of(p, f, 0) + ws.nq() * decommit::MAX_DATA_LAYERS * 4 + i * decommit::MAX_DATA_LAYERS * 4;
}
HD const u32 *hint_hashes(const Workspace &ws, u32 p, u32 f, u32 i) {
    return ws.hint_of(p, f, 0) + 2 * ws.nq() * decommit::MAX_DATA_LAYERS * 4 + (size_t)i * (ws.shape.fri_depth(f) - 1) * 8;
}

// full mode: SinglePairMerkleProofVar::verify for one (proof, layer, query) + consistency of the opened values with the folds
HD void stage_pair_path(const Workspace &ws, u32 p, u32 f, u32 i) {
    const Desc &d = ws.desc[p];
    if (!d.ok) return;
    Detail &dt = ws.detail[p];
    const u32 depth = ws.shape.fri_depth(f);
    const u32 q = fri::position(d, dt.fs.raw_queries[i], depth);
    u32 root[8], perms = 0;
    u32 *sink = nullptr;
    if (ws.perm_out) {
        u32 at = ws.pair_hint_base;
        for (u32 ff = 0; ff < f; ff++) at += decommit::pair_path_perms(ws.shape.fri_depth(ff), ws.shape.fri_data_mask(ff)) * ws.nq();
        sink = ws.perm_out_of(p, at + i * decommit::pair_path_perms(depth, ws.shape.fri_data_mask(f)));
    }
    decommit::pair_path_root(depth, ws.shape.fri_data_mask(f), q, hint_self(ws, p, f, i), hint_sib(ws, p, f, i), hint_hashes(ws, p, f, i), root, &perms, sink,
                             ws.in_delta());
    decommit::cp8(ws.root_of(p, 4 + f, i), root);
    VERIFY_ATOMIC_ADD(&dt.n_perms_paths, perms);
    const u32 *want = ws.blob(p) + (f ? d.in_commitment[f - 1] : d.fl_commitment);
    if (!decommit::eq8(root, want)) fail_shared(&dt, f ? proof::ST_FRI_INNER : proof::ST_FRI_FIRST);
}

// ---- stage 6: verdict (one thread per proof) ------------------------------------------------------------------------
HD void stage_verdict(const Workspace &ws, u32 p) {
    Detail &dt = ws.detail[p];
    // the thread-per-path kernels write every slot of a parsed proof's record (they ran before this stage)
    if ((ws.mode & MODE_FULL) && (ws.mode & MODE_PATH_KERNELS) && ws.hint_trees && ws.desc[p].ok) ws.hint_trees[p] = ws.shape.n_trees();
    const u32 order[9] = {proof::ST_PARSE, proof::ST_POW, proof::ST_LOGUP, proof::ST_OODS, proof::ST_UNSUPPORTED, proof::ST_MERKLE,
                          proof::ST_FRI_FIRST, proof::ST_FRI_INNER, proof::ST_FRI_LAST};
    dt.verdict = proof::ACCEPT; dt.stage = proof::ST_OK;
    for (int k = 0; k < 9; k++)
        if (dt.fail_mask & (1u << order[k])) {
            dt.stage = order[k];
            dt.verdict = order[k] == proof::ST_UNSUPPORTED ? proof::UNSUPPORTED : proof::REJECT;
            break;
        }
}


// ---- workspace sizing (host) ------------------------------------------------------------------------------------------
struct Carver {
    uint8_t *base; size_t at;
    template <class T> T *take(size_t n) {
        at = (at + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + at) : nullptr;
        at += n * sizeof(T);
        return p;
    }
};
// lays the per-proof arrays out behind `base` (nullptr: only measures); returns the bytes needed
inline size_t carve(Workspace &ws, uint8_t *base) {
    Carver c{base, 0};
    const size_t n = ws.n_proofs, nq = ws.shape.n_queries, nf = ws.shape.n_fri_trees(), nt = ws.shape.n_trees();
    ws.desc = c.take<Desc>(n);
    ws.detail = c.take<Detail>(n);
    ws.path_cols = c.take<u32>(n * 4 * nq * PATH_COLS_STRIDE);
    ws.path_sib = c.take<u32>(n * 4 * nq * MAX_DEPTH * 8);
    ws.path_roots = c.take<u32>(n * nt * nq * 8);
    ws.single_scratch = c.take<u32>(n * 4 * decommit::SINGLE_SCRATCH_WORDS_PER_QUERY * nq);
    ws.groups = c.take<fri::Group>(n * fri::MAX_LOGS);
    ws.domain_points = c.take<u32>(n * fri::MAX_LOGS * nq * 2);
    ws.answers = c.take<u32>(n * fri::MAX_LOGS * nq * 4);
    ws.circle_folds = c.take<u32>(n * fri::MAX_LOGS * nq * 4);
    ws.line_folds = c.take<u32>(n * proof::MAX_INNER * nq * 4);
    ws.last_evals = c.take<u32>(n * nq * 4);
    ws.fri_vals = c.take<u32>(n * nf * ws.fri_vals_stride());
    ws.n_fri_vals = c.take<u32>(n * nf);
    ws.pair_hints = c.take<u32>(n * nf * nq * PAIR_HINT_WORDS);
    ws.pair_scratch = c.take<u32>(n * nf * decommit::PAIR_SCRATCH_WORDS_PER_QUERY * nq);
    ws.fold_buf = c.take<u32>(n * (ws.shape.log_last ? ((size_t)1 << (ws.shape.log_last - 1)) : 1) * 4);
    const HintLayout hl = hint_layout(ws.shape);
    ws.hint_total = hl.total; ws.pair_hint_base = hl.pair_base[0];
    ws.perm_out = c.take<u32>(n * (size_t)hl.total * 16);
    ws.perm_in = c.take<u32>(n * (size_t)hl.total * 16);
    ws.hint_trees = c.take<u32>(n);
    return (c.at + 255) & ~(size_t)255;
}

}  // namespace verify
