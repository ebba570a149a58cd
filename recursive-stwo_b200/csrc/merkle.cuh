// Poseidon31 Merkle hashing on device: stwo's Poseidon31MerkleHasher::hash_node and the
// per-query authentication path walk, written so that each thread owns one sponge / one path and
// the (rolled) permutation has a single call site per kernel loop.
//
// Follows the reference's restatement of hash_node:
//   primitives/merkle/src/lib.rs:50-91   (hash_m31_columns_get_rate: leaf = rate(perm(0^8 || cap)))
//   primitives/merkle/src/lib.rs:141-181 (hash_m31_columns_get_capacity: 8 words / chunk, zero pad,
//                                          capacity chained)
//   primitives/merkle/src/lib.rs:9-48    (hash_tree = rate(perm(l || r)); with column:
//                                          rate(perm(hash_tree || cap(cols))))
// and the path walk of components/recursive/data_structures/src/lib.rs:315-354 /
// components/hints/src/decommit.rs:22-42.
#pragma once
#include "poseidon2.cuh"
#include "../../include/stwo_b200.h"

namespace merkle {

// hash_node for one node.  load_col(c) returns column value c of this node.
// children == nullptr => leaf layer.  out: 8 words.
template <class LoadCol>
HD void hash_node(const u32 *children, LoadCol load_col, u32 n_cols, u32 out[8]) {
    u32 st[16];
    const u32 n_chunks = (!children && n_cols == 0) ? 1u : (n_cols + 7) / 8;
    // step -1 (only with children): tree hash; steps 0..n_chunks-1: sponge; last: combine / finalise
    u32 tree[8];
    int step = children ? -1 : 0;
#pragma unroll
    for (int i = 8; i < 16; i++) st[i] = 0;
    const int last = (int)n_chunks;            // index of the finalising permutation
    const bool need_final = !children || n_cols > 0;
    for (;;) {
        if (step < 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) st[i] = children[i];
        } else if (step < last) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 c = 8 * (u32)step + i;
                st[i] = c < n_cols ? load_col(c) : 0u;
            }
        } else {
            // finalise: leaf -> 0^8 || cap ; inner with columns -> tree || cap
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = children ? tree[i] : 0u;
        }
        poseidon2::permute<false>(st);
        if (step < 0) {
            if (!need_final) break;
#pragma unroll
            for (int i = 0; i < 8; i++) { tree[i] = st[i]; st[8 + i] = 0; }
        } else if (step == last) {
            break;
        }
        step++;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = st[i];
}

// Walk one authentication path; returns the recomputed root in out[8].
//   cols: this path's column values (leaf layer first, then injected layers, descending log size)
//   sib : depth x 8 sibling words, leaf level first
// sink (optional): receives the 16-word output state of every permutation, in execution order
// in_delta != 0: the input state of each permutation goes in_delta words beyond its output state (the parallel input record)
HD void path_root(const stwo_b200_path_shape &shape, u32 index, const u32 *cols, const u32 *sib, u32 out[8], u32 *sink = nullptr, size_t in_delta = 0) {
    enum { PH_SPONGE = 0, PH_FINAL_LEAF = 1, PH_NODE = 2, PH_COMBINE = 3 };
    u32 st[16];
    u32 saved[8];
#pragma unroll
    for (int i = 0; i < 8; i++) saved[i] = 0;
#pragma unroll
    for (int i = 8; i < 16; i++) st[i] = 0;
    const u32 depth = shape.depth;
    u32 h = depth;                       // layer currently being completed
    u32 rem = shape.n_cols[depth];       // column words still to absorb at this layer
    int ph = PH_SPONGE;
    bool leaf = true;
    for (;;) {
        if (ph == PH_SPONGE) {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = (u32)i < rem ? cols[i] : 0u;
            u32 take = rem < 8 ? rem : 8;
            cols += take;
            rem -= take;
        } else if (ph == PH_FINAL_LEAF) {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = 0;
        } else if (ph == PH_NODE) {
            const u32 lvl = depth - h;                   // sibling level, leaf level = 0
            const u32 *s8 = sib + 8 * lvl;
            const bool right = (index >> lvl) & 1u;      // we are the right child
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 mine = st[i], other = s8[i];
                st[i] = right ? other : mine;
                st[8 + i] = right ? mine : other;
            }
        } else {   // PH_COMBINE: tree hash || capacity of this layer's columns (already in st[8..16])
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = saved[i];
        }
        if (sink && in_delta) {
#pragma unroll
            for (int i = 0; i < 16; i++) sink[in_delta + i] = st[i];
        }
        poseidon2::permute<false>(st);
        if (sink) {
#pragma unroll
            for (int i = 0; i < 16; i++) sink[i] = st[i];
            sink += 16;
        }
        // transitions
        if (ph == PH_SPONGE) {
            if (rem == 0) ph = leaf ? PH_FINAL_LEAF : PH_COMBINE;
            continue;
        }
        if (ph == PH_NODE) {
            h--;
            rem = shape.n_cols[h];
            if (rem > 0) {           // absorb this layer's columns next, then combine
#pragma unroll
                for (int i = 0; i < 8; i++) { saved[i] = st[i]; st[8 + i] = 0; }
                ph = PH_SPONGE;
                leaf = false;
                continue;
            }
        }
        // st[0..8] is the complete hash of layer h
        if (h == 0) break;
        ph = PH_NODE;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = st[i];
}

HD u32 path_perms(const stwo_b200_path_shape &shape) {
    u32 n = (shape.n_cols[shape.depth] + 7) / 8 + 1;
    for (u32 h = 0; h < shape.depth; h++) n += 1 + (shape.n_cols[h] ? (shape.n_cols[h] + 7) / 8 + 1 : 0);
    return n;
}

}  // namespace merkle
