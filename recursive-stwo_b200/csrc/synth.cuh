// Synthetic FRI + Merkle instances (BASELINE configs[4] part i): per instance, seeded low-degree columns of
// the shape's three log sizes -> committed in one mixed-degree first-layer tree -> alpha from a Poseidon channel over the roots ->
// folded layer by layer (every inner layer committed) -> last-layer polynomial interpolated from the final layer -> queries drawn
// from the channel -> batched decommitments in stwo's layout.  Nothing here is OODS-consistent (there is
// no STARK behind the columns); it is exactly the FRI query phase + its Merkle decommitments that the verifier then checks, on
// instances that all differ: different query positions, node sharing and witness consumption order in every lane.
//
// This header is the element-level arithmetic, HD so that tests/hostsim builds whole instances on the CPU box; synth_kernels.cu runs
// the same functions one thread per element.  The prover-side rules mirror what the verifier consumes:
//   first-layer values / fri_witness order   components/hints/src/folding.rs:414-451
//   inner-layer fri_witness order             components/hints/src/folding.rs:497-502
//   hash_witness order (left before right, layer by layer)   stwo MerkleVerifier::verify as replayed in decommit.cuh (pair_tree)
//   folds                                     components/recursive/folding/src/lib.rs:56-204, primitives/line/src/lib.rs:39-67
#pragma once
#include "verify.cuh"

namespace synth {

constexpr u32 HDR = 256, MAGIC = 0x53594E54u;          // 'SYNT'
// header words of an instance blob; every section offset is in words from the start of the blob
enum : u32 { H_MAGIC = 0, H_SHAPE = 1 /* 7 words */, H_FL_NFW = 8, H_FL_NHW = 9, H_IN_NFW = 10 /* 32 */, H_IN_NHW = 42 /* 32 */, H_NONCE = 74 /* 2 */,
             H_TOTAL = 76, H_OFF_FLC = 80, H_OFF_LAST = 81, H_OFF_ANS = 82, H_OFF_FLFW = 83, H_OFF_FLHW = 84, H_NLAST = 85,
             H_OFF_INC = 96 /* 32 */, H_OFF_INFW = 128 /* 32 */, H_OFF_INHW = 160 /* 32 */ };

struct Layout {
    u32 off_flc, off_last, off_ans, off_flfw, off_flhw, off_inc[proof::MAX_INNER], off_infw[proof::MAX_INNER], off_inhw[proof::MAX_INNER], total;
};
// capacities are worst cases over the query positions: the counts actually used are in the header
HD Layout layout(const verify::Shape &s) {
    Layout l;
    const u32 nq = s.n_queries;
    u32 at = HDR;
    l.off_flc = at; at += 8;
    for (u32 i = 0; i < s.n_inner; i++) { l.off_inc[i] = at; at += 8; }
    l.off_last = at; at += 4u << s.log_last;
    l.off_ans = at; at += fri::MAX_LOGS * nq * 4;
    l.off_flfw = at; at += fri::MAX_LOGS * nq * 4;
    l.off_flhw = at; at += (s.max_first() + 1) * 2 * nq * 8;
    for (u32 i = 0; i < s.n_inner; i++) {
        l.off_infw[i] = at; at += nq * 4;
        l.off_inhw[i] = at; at += (s.fri_depth(1 + i) + 1) * 2 * nq * 8;
    }
    l.total = at;
    return l;
}
HD void write_header(u32 *blob, const verify::Shape &s, const Layout &l) {
    for (u32 k = 0; k < HDR; k++) blob[k] = 0;
    blob[H_MAGIC] = MAGIC;
    blob[H_SHAPE + 0] = s.log_size_plonk; blob[H_SHAPE + 1] = s.log_size_poseidon; blob[H_SHAPE + 2] = s.pow_bits; blob[H_SHAPE + 3] = s.log_blowup;
    blob[H_SHAPE + 4] = s.log_last; blob[H_SHAPE + 5] = s.n_queries; blob[H_SHAPE + 6] = s.n_inner;
    blob[H_TOTAL] = l.total; blob[H_OFF_FLC] = l.off_flc; blob[H_OFF_LAST] = l.off_last; blob[H_OFF_ANS] = l.off_ans;
    blob[H_OFF_FLFW] = l.off_flfw; blob[H_OFF_FLHW] = l.off_flhw; blob[H_NLAST] = 1u << s.log_last;
    for (u32 i = 0; i < s.n_inner; i++) { blob[H_OFF_INC + i] = l.off_inc[i]; blob[H_OFF_INFW + i] = l.off_infw[i]; blob[H_OFF_INHW + i] = l.off_inhw[i]; }
}
// section offsets of `shape` into a Desc (no validation; the generator uses it while the blob is being built)
HD void set_offsets(const verify::Shape &shape, proof::Desc &d) {
    const Layout l = layout(shape);
    d.log_size_plonk = shape.log_size_plonk; d.log_size_poseidon = shape.log_size_poseidon; d.pow_bits = shape.pow_bits; d.log_blowup = shape.log_blowup;
    d.log_last = shape.log_last; d.n_queries = shape.n_queries; d.n_inner = shape.n_inner;
    d.max_first = shape.max_first(); d.log_plonk = shape.log_plonk(); d.log_pos = shape.log_pos();
    d.pow_nonce = H_NONCE;
    d.fl_commitment = l.off_flc; d.fl_fri_witness = l.off_flfw; d.fl_hash_witness = l.off_flhw;
    for (u32 i = 0; i < shape.n_inner; i++) { d.in_commitment[i] = l.off_inc[i]; d.in_fri_witness[i] = l.off_infw[i]; d.in_hash_witness[i] = l.off_inhw[i]; }
    d.last_coeffs = l.off_last; d.n_last_coeffs = 1u << shape.log_last;
}
// The verifier's view of an instance: a Desc whose FRI sections point into the blob (the commitment-tree / sampled-value sections of a
// real proof do not exist).  Returns false for anything that is not an instance of `shape`.
HD bool fill_desc(const u32 *blob, size_t n_words, const verify::Shape &shape, proof::Desc &d) {
    d.ok = 0;
    if (n_words < HDR || blob[H_MAGIC] != MAGIC) return false;
    verify::Shape s;
    s.log_size_plonk = blob[H_SHAPE + 0]; s.log_size_poseidon = blob[H_SHAPE + 1]; s.pow_bits = blob[H_SHAPE + 2]; s.log_blowup = blob[H_SHAPE + 3];
    s.log_last = blob[H_SHAPE + 4]; s.n_queries = blob[H_SHAPE + 5]; s.n_inner = blob[H_SHAPE + 6];
    if (s.log_size_plonk != shape.log_size_plonk || s.log_size_poseidon != shape.log_size_poseidon || s.pow_bits != shape.pow_bits ||
        s.log_blowup != shape.log_blowup || s.log_last != shape.log_last || s.n_queries != shape.n_queries || s.n_inner != shape.n_inner)
        return false;
    const Layout l = layout(shape);
    if (blob[H_TOTAL] != l.total || n_words < l.total) return false;
    d.log_size_plonk = s.log_size_plonk; d.log_size_poseidon = s.log_size_poseidon; d.pow_bits = s.pow_bits; d.log_blowup = s.log_blowup;
    d.log_last = s.log_last; d.n_queries = s.n_queries; d.n_inner = s.n_inner;
    d.max_first = shape.max_first(); d.log_plonk = shape.log_plonk(); d.log_pos = shape.log_pos();
    d.pow_nonce = H_NONCE;
    d.fl_commitment = l.off_flc; d.fl_fri_witness = l.off_flfw; d.fl_hash_witness = l.off_flhw;
    d.fl_n_fri_witness = blob[H_FL_NFW]; d.fl_n_hash_witness = blob[H_FL_NHW];
    if (d.fl_n_fri_witness > fri::MAX_LOGS * s.n_queries || d.fl_n_hash_witness > (shape.max_first() + 1) * 2 * s.n_queries) return false;
    for (u32 i = 0; i < s.n_inner; i++) {
        d.in_commitment[i] = l.off_inc[i]; d.in_fri_witness[i] = l.off_infw[i]; d.in_hash_witness[i] = l.off_inhw[i];
        d.in_n_fri_witness[i] = blob[H_IN_NFW + i]; d.in_n_hash_witness[i] = blob[H_IN_NHW + i];
        if (d.in_n_fri_witness[i] > s.n_queries || d.in_n_hash_witness[i] > (shape.fri_depth(1 + i) + 1) * 2 * s.n_queries) return false;
    }
    d.last_coeffs = l.off_last; d.n_last_coeffs = 1u << s.log_last;
    // every field element of the used sections must be canonical
    for (u32 k = l.off_flc; k < l.off_flfw + 4 * d.fl_n_fri_witness; k++) if (blob[k] >= M31_P) return false;      // commitments .. answers, used witnesses
    for (u32 k = 0; k < 8 * d.fl_n_hash_witness; k++) if (blob[l.off_flhw + k] >= M31_P) return false;
    for (u32 i = 0; i < s.n_inner; i++) {
        for (u32 k = 0; k < 4 * d.in_n_fri_witness[i]; k++) if (blob[l.off_infw[i] + k] >= M31_P) return false;
        for (u32 k = 0; k < 8 * d.in_n_hash_witness[i]; k++) if (blob[l.off_inhw[i] + k] >= M31_P) return false;
    }
    d.ok = 1;
    return true;
}

// ---- the Fiat-Shamir part of an instance: a fresh channel over the FRI commitments (the tail of fs::transcript) ----------------------
// `upto`: number of commitments to absorb (1 = first layer only ... 1 + n_inner = all); with all of them the last-layer polynomial,
// the nonce and the query draws follow.  The generator calls it after every commit (it needs alpha_k before it can fold layer k).
HD void transcript(const u32 *w, const proof::Desc &d, fs::Out &o, u32 upto) {
    fs::Channel ch;
    ch.init(nullptr);
    u32 dr[8];
    ch.mix8(w + d.fl_commitment);
    ch.draw(dr); o.fri_alphas[0] = fs::qload(dr);
    for (u32 i = 0; i < d.n_inner && i + 1 < upto; i++) {
        ch.mix8(w + d.in_commitment[i]);
        ch.draw(dr); o.fri_alphas[i + 1] = fs::qload(dr);
    }
    if (upto < 1 + d.n_inner) return;
    for (u32 i = 0; i < d.n_last_coeffs; i += 2) {
        if (i + 1 < d.n_last_coeffs) ch.mix8(w + d.last_coeffs + 4 * i);
        else ch.mix4(w + d.last_coeffs + 4 * i);
    }
    const u64 nonce = (u64)w[d.pow_nonce] | ((u64)w[d.pow_nonce + 1] << 32);
    u32 nf[4] = {(u32)(nonce & ((1u << 22) - 1)), (u32)((nonce >> 22) & ((1u << 21) - 1)), (u32)((nonce >> 43) & ((1u << 21) - 1)), 0};
    ch.mix4(nf);
    for (int i = 0; i < 8; i++) o.digest_after_nonce[i] = ch.st[8 + i];
    o.pow_ok = (ch.st[8] & ((1u << d.pow_bits) - 1)) == 0;
    u32 got = 0;
    for (u32 k = 0; k < (d.n_queries + 3) / 4; k++) {
        ch.draw(dr);
        for (int j = 0; j < 8 && got < d.n_queries; j++) o.raw_queries[got++] = dr[j];
    }
    o.n_transcript_perms = ch.n_perms;
}

// The reference does not support two queries on the same position of the largest domain (components/recursive/answer/src/lib.rs:190-195):
// the generator grinds the nonce past such draws, the way a prover grinds it for the proof of work.
HD bool queries_distinct(const proof::Desc &d, const fs::Out &o) {
    for (u32 i = 0; i < d.n_queries; i++)
        for (u32 j = 0; j < i; j++)
            if (fri::position(d, o.raw_queries[i], d.max_first) == fri::position(d, o.raw_queries[j], d.max_first)) return false;
    return true;
}

// ---- seeded low-degree columns ----------------------------------------------------------------------------------------------------
HD u64 splitmix(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
constexpr u32 N_TERMS = 8;
// term t of column g of instance `seed`: exponent e < 2^(L - blowup) in the circle FFT basis and a QM31 coefficient; term 0 has the top
// exponent bit set, so the column has the full degree its log size allows
HD void term(u64 seed, u32 g, u32 t, u32 log_degree, u32 &e, qm31_t &c) {
    u64 r = splitmix(seed * 0x100000001B3ull + g * 1315423911ull + t * 2654435761ull + 12345);
    e = log_degree ? (u32)(r & ((1u << log_degree) - 1u)) : 0u;
    if (t == 0 && log_degree) e |= 1u << (log_degree - 1);
    u32 v[4];
    for (int k = 0; k < 4; k++) { r = splitmix(r); v[k] = (u32)(r % M31_P); }
    c = qm31::mk(v[0], v[1], v[2], v[3]);
}
// value of column g (log size L) at committed position i: sum_t c_t * y^{e_t bit 0} * prod_{j >= 1} pi^{j-1}(x)^{e_t bit j}, (x, y) the
// circle-domain point of the position (fri::domain_point)
HD qm31_t column_value(u64 seed, u32 g, u32 L, u32 log_blowup, u32 i) {
    const cpoint_t p = fri::domain_point(L, i);
    const u32 log_degree = L - log_blowup;
    u32 chain[32];
    chain[0] = p.y;
    u32 x = p.x;
    for (u32 j = 1; j < log_degree; j++) { chain[j] = x; const u32 sq = m31::mulc(x, x); x = m31::subc(m31::addc(sq, sq), 1); }
    qm31_t acc = qm31::zero();
    for (u32 t = 0; t < N_TERMS; t++) {
        u32 e; qm31_t c;
        term(seed, g, t, log_degree, e, c);
        u32 b = 1;
        for (u32 j = 0; j < log_degree; j++) if ((e >> j) & 1u) b = m31::mulc(b, chain[j]);
        acc = qm31::add(acc, qm31::mul_m31(c, b));
    }
    return acc;
}

// ---- folds of whole layers (the prover side of fri::fold_pair) ---------------------------------------------------------------------------
// element j of the line layer of log size L - 1 obtained from the circle column of log size L
HD qm31_t circle_fold_at(u32 L, u32 j, qm31_t even, qm31_t odd, qm31_t alpha) {
    const cpoint_t pt = circle::dbl(fri::absolute_point(L, 2 * j));
    return fri::fold_pair(even, odd, 2 * j, fs::minv(pt.y), alpha);
}
// element j of the line layer of log size L - 1 obtained from the line layer of log size L
HD qm31_t line_fold_at(u32 L, u32 j, qm31_t even, qm31_t odd, qm31_t alpha) {
    return fri::fold_pair(even, odd, 2 * j, fs::minv(fri::absolute_point(L, 2 * j).x), alpha);
}
// x-coordinate of position r of a line layer of log size L: the pair (2m, 2m + 1) sits at (+x, -x), x = absolute_point(L, 2m).x
HD u32 line_x(u32 L, u32 r) {
    const u32 x = fri::absolute_point(L, r).x;
    return (r & 1u) ? m31::negc(x) : x;
}
// Coefficients (LinePolyVar order, primitives/line/src/lib.rs:39-67) of the polynomial of degree < 2^k through the first 2^k evaluations of
// a line layer of log size L: f(x) = f_e(pi(x)) + x f_o(pi(x)), coefficients = those of f_e then those of f_o.  vals: 2^k QM31, destroyed;
// out: 2^k QM31.  tmp: 2^k QM31 scratch.
HD void interpolate_line(u32 L, u32 k, qm31_t *vals, qm31_t *out, qm31_t *tmp) {
    const u32 inv2 = (M31_P + 1) / 2;
    // level t works on blocks of 2^(k - t) consecutive values; within a block the values sit at positions 0 .. of a layer of log size L - t
    for (u32 t = 0; t < k; t++) {
        const u32 blk = 1u << (k - t), n_blk = 1u << t;
        for (u32 b = 0; b < n_blk; b++) {
            qm31_t *v = vals + (size_t)b * blk;
            for (u32 m = 0; m < blk / 2; m++) {
                const qm31_t a = v[2 * m], c = v[2 * m + 1];
                const u32 x = line_x(L - t, 2 * m);
                tmp[m] = qm31::mul_m31(qm31::add(a, c), inv2);
                tmp[blk / 2 + m] = qm31::mul_m31(qm31::sub(a, c), m31::mulc(inv2, fs::minv(x)));
            }
            for (u32 m = 0; m < blk; m++) v[m] = tmp[m];
        }
    }
    for (u32 m = 0; m < (1u << k); m++) out[m] = vals[m];
}

// ---- Merkle nodes of the generator's full trees -------------------------------------------------------------------------------------------
HD void node_hash(const u32 *left, const u32 *right, const u32 *val4, u32 out[8]) {       // hash_node(children?, one QM31 column value?)
    decommit::hash_node2(left, right, val4, val4 ? 4 : 0, out);
}

// ---- decommitment of one FRI tree, prover side ---------------------------------------------------------------------------------------------
// Walks the layers exactly like decommit::pair_tree (the verifier's replay) and writes, where the verifier takes a hash from hash_witness,
// the hash of that node of the FULL tree.  node(h, pos) -> const u32* (8 words) of the full tree's node at layer h.  Returns the count.
template <class NodeAt>
HD u32 emit_hash_witness(u32 depth, u32 data_mask, const u32 *q, u32 nq, NodeAt node, u32 *hw, u32 *scratch /* 6 * nq words */) {
    const u32 cap = 2 * nq;
    u32 *qs = scratch, *cpos = qs + cap, *npos = cpos + cap;
    for (u32 i = 0; i < nq; i++) qs[i] = q[i];
    u32 n = decommit::sort_unique(qs, nq), cm = 0, wi = 0;
    for (u32 h = depth + 1; h-- > 0;) {
        if (h < depth) {
            for (u32 k = 0; k < n; k++) qs[k] >>= 1;
            n = decommit::sort_unique(qs, n);
        }
        const bool data = (data_mask >> h) & 1u;
        u32 m = 0;
        if (data) {
            for (u32 k = 0; k < n; k++) { npos[m++] = qs[k]; npos[m++] = qs[k] ^ 1u; }
            m = decommit::sort_unique(npos, m);
        } else for (u32 k = 0; k < n; k++) npos[m++] = qs[k];
        if (h < depth)
            for (u32 a = 0; a < m; a++) {
                const u32 p = npos[a];
                if (decommit::find(cpos, cm, p << 1) < 0) { const u32 *s = node(h + 1, p << 1); for (int c = 0; c < 8; c++) hw[8 * wi + c] = s[c]; wi++; }
                if (decommit::find(cpos, cm, (p << 1) + 1) < 0) { const u32 *s = node(h + 1, (p << 1) + 1); for (int c = 0; c < 8; c++) hw[8 * wi + c] = s[c]; wi++; }
            }
        for (u32 a = 0; a < m; a++) cpos[a] = npos[a];
        cm = m;
    }
    return wi;
}

// ---- verifier, stage 1 of an instance: what parse + transcript + answers are for a real proof ------------------------------------------
// Desc from the blob header, Fiat-Shamir over the FRI commitments (alphas, PoW, queries), and the opened first-layer values -- for a real
// proof the DEEP quotient answers, here part of the instance ("fri_answers supplied as witnesses") and authenticated by the first-layer
// tree like them.  The folds and the FRI tree rebuilds then run unchanged (verify::stage_folds*, stage_pair_tree*).
HD void stage_open(const verify::Workspace &ws, u32 p) {
    proof::Desc &d = ws.desc[p];
    verify::Detail &dt = ws.detail[p];
    verify::reset_detail(dt);
    // the four commitment trees have no part in an instance, so nothing of theirs can be missing from the permutation record: the
    // count starts at 4 and reaches n_trees() when every FRI tree's part is complete (what the folding-stage circuit waits for)
    if (ws.hint_trees) ws.hint_trees[p] = 4;
    const u32 *w = ws.blob(p);
    if (!fill_desc(w, ws.blob_words(p), ws.shape, d)) { d.ok = 0; verify::fail(dt, proof::ST_PARSE); return; }
    transcript(w, d, dt.fs, 1 + d.n_inner);
    verify::stage_after_transcript(ws, p);
    const u32 nq = d.n_queries, off = layout(ws.shape).off_ans;
    for (u32 k = 0; k < fri::MAX_LOGS * nq * 4; k++) ws.answers[(size_t)p * fri::MAX_LOGS * nq * 4 + k] = w[off + k];
}

}  // namespace synth
