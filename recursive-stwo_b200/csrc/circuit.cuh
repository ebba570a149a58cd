// Witness stream of the recursive verifier circuit: resolves a tape::Src word against one proof's blob and the batched
// verifier's workspace (verify.cuh), i.e. it plays the role of the host structs the reference hands to the circuit:
// the proof itself, FiatShamirHints.oods_point (examples/single-proof/src/main.rs:57), DecommitHints
// (components/hints/src/decommit.rs:186-241) and First/InnerLayersHints (components/hints/src/folding.rs:290-601).
#pragma once
#include "tape.cuh"
#include "verify.cuh"

namespace circuit {

HD u32 gather_word(const verify::Workspace &ws, u32 p, u32 src, const u32 *extra = nullptr) {
    const proof::Desc &d = ws.desc[p];
    const u32 *w = ws.blob(p);
    const u32 a = tape::src_a(src), i = tape::src_i(src), k = tape::src_k(src);
    switch (tape::src_section(src)) {
    case tape::S_STMT0: return w[k];
    case tape::S_STMT1: return w[d.stmt1 + k];
    case tape::S_COMMITMENT: return w[d.commitments[a] + k];
    case tape::S_SAMPLED: return w[proof::sample_off(d, a, i, k >> 2) + (k & 3u)];
    case tape::S_FRI_COMMITMENT: return w[(a ? d.in_commitment[a - 1] : d.fl_commitment) + k];
    case tape::S_LAST_COEFFS: return w[d.last_coeffs + k];
    case tape::S_POW_LIMB: {
        // data_structures/src/lib.rs:189-205: 22 / 21 / 21-bit limbs of the nonce
        const u64 n = (u64)w[d.pow_nonce] | ((u64)w[d.pow_nonce + 1] << 32);
        return k == 0 ? (u32)(n & ((1u << 22) - 1)) : k == 1 ? (u32)((n >> 22) & ((1u << 21) - 1)) : (u32)((n >> 43) & ((1u << 21) - 1));
    }
    case tape::S_OODS: return k < 4 ? ws.detail[p].fs.oods_x.v[k] : ws.detail[p].fs.oods_y.v[k - 4];
    case tape::S_PATH_COL: return ws.cols_of(p, a, i)[k];
    case tape::S_PATH_SIB: return ws.sib_of(p, a, i)[k];
    case tape::S_PAIR_SELF: return verify::hint_self(ws, p, a, i)[k];
    case tape::S_PAIR_SIB: return verify::hint_sib(ws, p, a, i)[k];
    case tape::S_PAIR_HASH: return verify::hint_hashes(ws, p, a, i)[k];
    case tape::S_FS: {
        const fs::Out &f = ws.detail[p].fs;
        if (k == tape::FS_ZERO) return 0;
        if (k >= tape::FS_QUERY_BASE) return fri::position(d, f.raw_queries[k - tape::FS_QUERY_BASE], d.max_first);
        const qm31_t *q = k < 4 ? &f.oods_t : k < 8 ? &f.z : k < 12 ? &f.alpha : k < 16 ? &f.random_coeff : k < 20 ? &f.after_coeff : &f.fri_alphas[(k - 20) / 4];
        return q->v[k & 3u];
    }
    case tape::S_EXTRA: return extra ? extra[k] : 0;
    case tape::S_ANSWER: return ws.q4(ws.answers, p, a, fri::MAX_LOGS, i)[k];
    default: return 0;
    }
}

// Public-input hashes of the last-layer circuit: Poseidon31MerkleHasher::hash_node(None, values) of (a) every sampled value
// (components/last/fiat_shamir/src/lib.rs:42-54) and (b) every > 8-column opening of a commitment-tree query
// (components/last/answer/src/data_structures/merkle_proofs.rs:186-204).  One job = 8 output words.
struct ExtraJob { u32 kind, tree, query, col_off, n_cols, slot; };     // kind 0: sampled values, 1: path columns
HD void extra_job(const verify::Workspace &ws, u32 p, const ExtraJob &j, u32 *extra) {
    const proof::Desc &d = ws.desc[p];
    const u32 *w = ws.blob(p);
    u32 out[8];
    if (j.kind == 0) {
        // flattened sampled values: tree -> column -> mask -> 4 words
        merkle::hash_node(nullptr, [&](u32 c) {
            u32 q = c >> 2, t = 0, col = 0;
            for (t = 0; t < 4; t++) {
                u32 n = 0;
                for (col = 0; col < proof::n_cols(t); col++) n += proof::n_masks(t, col);
                if (q < n) break;
                q -= n;
            }
            for (col = 0; col < proof::n_cols(t); col++) { const u32 m = proof::n_masks(t, col); if (q < m) break; q -= m; }
            return w[proof::sample_off(d, t, col, q) + (c & 3u)];
        }, proof::TOTAL_SAMPLES * 4, out);
    } else {
        const u32 *cols = ws.cols_of(p, j.tree, j.query) + j.col_off;
        merkle::hash_node(nullptr, [&](u32 c) { return cols[c]; }, j.n_cols, out);
    }
    for (int k = 0; k < 8; k++) extra[j.slot + k] = out[k];
}

}  // namespace circuit
