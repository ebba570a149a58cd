// Witness stream of the recursive verifier circuit: resolves a tape::Src word against one proof's blob and the batched
// verifier's workspace (verify.cuh), i.e. it plays the role of the host structs the reference hands to the circuit:
// the proof itself, FiatShamirHints.oods_point (examples/single-proof/src/main.rs:57), DecommitHints
// (components/hints/src/decommit.rs:186-241) and First/InnerLayersHints (components/hints/src/folding.rs:290-601).
#pragma once
#include "tape.cuh"
#include "verify.cuh"

namespace circuit {

HD u32 gather_word(const verify::Workspace &ws, u32 p, u32 src) {
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
    default: return 0;
    }
}

}  // namespace circuit
