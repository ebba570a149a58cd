// Hint pre-pass on device: stwo's batched Merkle decommitments are replayed layer by layer
// (recomputing every node the batch touches, consuming `hash_witness` and the queried values in
// stream order) and re-shaped into independent per-query authentication paths, which is what the
// verifier circuit consumes.  One thread owns one tree of one proof; node tables are small sorted
// arrays in a caller-provided scratch area.
//
// Follows
//   components/hints/src/decommit.rs:44-183   SinglePathMerkleProof::from_stwo_proof (commitment trees)
//   components/hints/src/folding.rs:93-287    SinglePairMerkleProof::from_stwo_proof (FRI layer trees)
//   components/hints/src/folding.rs:33-91     SinglePairMerkleProof::verify
#pragma once
#include "fiat_shamir.cuh"

namespace decommit {

using fs::permute_mem;

// Scratch arrays are addressed through a stride so that, on the device, word i of every thread of a warp is adjacent
// (coalesced) instead of each thread owning a contiguous block; on the host the stride is 1.
struct Strided {
    u32 *p; u32 stride;
    HDM u32 &operator[](u32 i) const { return p[(size_t)i * stride]; }
    HDM Strided at(u32 off) const { Strided r; r.p = p + (size_t)off * stride; r.stride = stride; return r; }
};
HD void ld8(const Strided &s, u32 k, u32 out[8]) { for (u32 i = 0; i < 8; i++) out[i] = s[8 * k + i]; }
HD void st8(const Strided &s, u32 k, const u32 in[8]) { for (u32 i = 0; i < 8; i++) s[8 * k + i] = in[i]; }

// hash_node on register-resident children (L == nullptr: leaf); cols is contiguous memory (the proof blob).
// primitives/merkle/src/lib.rs:9-181
// sink (optional, advanced): receives the 16-word output state of every permutation, in execution order; with in_delta != 0 the
// 16-word INPUT state of the same permutation goes in_delta words further (the input record runs parallel to the output record)
HD void tap(u32 **sink, const u32 *st) {
    if (sink && *sink) { for (int i = 0; i < 16; i++) (*sink)[i] = st[i]; *sink += 16; }
}
HD void tap_in(u32 **sink, const u32 *st, size_t in_delta) {
    if (sink && *sink && in_delta) for (int i = 0; i < 16; i++) (*sink)[in_delta + i] = st[i];
}
HD void hash_node2(const u32 *L, const u32 *R, const u32 *cols, u32 nc, u32 *out, u32 *tree_out = nullptr, u32 **sink = nullptr, size_t in_delta = 0) {
    u32 st[16];
    u32 tree[8];
    if (L) {
        for (int i = 0; i < 8; i++) { st[i] = L[i]; st[8 + i] = R[i]; }
        tap_in(sink, st, in_delta);
        permute_mem(st);
        tap(sink, st);
        if (tree_out) for (int i = 0; i < 8; i++) tree_out[i] = st[i];
        if (nc == 0) { for (int i = 0; i < 8; i++) out[i] = st[i]; return; }
        for (int i = 0; i < 8; i++) tree[i] = st[i];
    }
    for (int i = 8; i < 16; i++) st[i] = 0;
    u32 n_chunks = nc ? (nc + 7) / 8 : 1;
    for (u32 c = 0; c < n_chunks; c++) {
        for (u32 i = 0; i < 8; i++) st[i] = 8 * c + i < nc ? cols[8 * c + i] : 0u;
        tap_in(sink, st, in_delta);
        permute_mem(st);
        tap(sink, st);
    }
    for (int i = 0; i < 8; i++) st[i] = L ? tree[i] : 0u;
    tap_in(sink, st, in_delta);
    permute_mem(st);
    tap(sink, st);
    for (int i = 0; i < 8; i++) out[i] = st[i];
}
// The same node with a callback around every permutation: tap(j, st, false) sees the 16-word input state of the node's j-th permutation,
// tap(j, st, true) its output state (execution order: [hash of the children,] sponge chunks, finalisation).  The cooperative tree rebuilds use it to hand each
// state to the queries whose authentication path runs through this node (decommit_coop.cuh).
template <class Tap>
HD void hash_node2_tap(const u32 *L, const u32 *R, const u32 *cols, u32 nc, u32 *out, u32 *tree_out, Tap tap) {
    u32 st[16];
    u32 tree[8];
    u32 j = 0;
    if (L) {
        for (int i = 0; i < 8; i++) { st[i] = L[i]; st[8 + i] = R[i]; }
        tap(j, st, false);
        permute_mem(st);
        tap(j++, st, true);
        if (tree_out) for (int i = 0; i < 8; i++) tree_out[i] = st[i];
        if (nc == 0) { for (int i = 0; i < 8; i++) out[i] = st[i]; return; }
        for (int i = 0; i < 8; i++) tree[i] = st[i];
    }
    for (int i = 8; i < 16; i++) st[i] = 0;
    u32 n_chunks = nc ? (nc + 7) / 8 : 1;
    for (u32 c = 0; c < n_chunks; c++) {
        for (u32 i = 0; i < 8; i++) st[i] = 8 * c + i < nc ? cols[8 * c + i] : 0u;
        tap(j, st, false);
        permute_mem(st);
        tap(j++, st, true);
    }
    for (int i = 0; i < 8; i++) st[i] = L ? tree[i] : 0u;
    tap(j, st, false);
    permute_mem(st);
    tap(j++, st, true);
    for (int i = 0; i < 8; i++) out[i] = st[i];
}
HD u32 node_perms(bool leaf, u32 nc) { return leaf ? (nc ? (nc + 7) / 8 : 1) + 1 : 1 + (nc ? (nc + 7) / 8 + 1 : 0); }

HD bool eq8(const u32 *a, const u32 *b) {
    u32 d = 0;
    for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
    return d == 0;
}
HD void cp8(u32 *dst, const u32 *src) { for (int i = 0; i < 8; i++) dst[i] = src[i]; }

// in-place insertion sort + dedup of n <= 256 values; returns the unique count
template <class A>
HD u32 sort_unique(A v, u32 n) {
    for (u32 i = 1; i < n; i++) {
        u32 x = v[i], j = i;
        while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; }
        v[j] = x;
    }
    u32 m = 0;
    for (u32 i = 0; i < n; i++) if (m == 0 || v[m - 1] != v[i]) v[m++] = v[i];
    return m;
}
template <class A>
HD int find(A v, u32 n, u32 x) {
    u32 lo = 0, hi = n;
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (v[mid] < x) lo = mid + 1; else hi = mid; }
    return (lo < n && v[lo] == x) ? (int)lo : -1;
}

// ---- commitment trees ---------------------------------------------------------------------------------
constexpr u32 SINGLE_SCRATCH_WORDS_PER_QUERY = 16;   // two layers of node hashes
constexpr u32 SINGLE_IDX_WORDS_PER_QUERY = 6;        // thread-local position / index tables

// Columns live at two layers at most: nA columns at log size hA and nB at hB (the Plonk and the Poseidon
// component; composition tree: nB = 0).  depth = max(hA, hB) = log size of the leaf layer.
struct SingleShape {
    u32 depth, hA, nA, hB, nB;
    HDM u32 ncols(u32 h) const { return (h == hA ? nA : 0u) + (h == hB ? nB : 0u); }
};

// q[nq]: query positions at the leaf layer.  Writes, for query i, its column values (leaf layer first, then the
// injected layer) to path_cols + i*cpp and its sibling hashes (leaf level first) to path_sib + i*sib_stride.
// Returns true iff both streams are consumed exactly and the recomputed root equals `root`.
HD bool single_tree(const SingleShape &sh, const u32 *q, u32 nq, const u32 *values, u32 n_values, const u32 *hw, u32 n_hw,
                    const u32 *root, u32 *path_cols, u32 cpp, u32 *path_sib, u32 sib_stride, Strided scratch, u32 *idx, u32 *perms) {
    // node hashes live in the (global, strided) scratch; the small position / index tables in thread-local memory
    Strided chash = scratch, phash = chash.at(8 * nq);
    u32 *cpos = idx, *ppos = cpos + nq, *sibsrc = ppos + nq, *par = sibsrc + nq, *colsrc = par + nq, *qnode = colsrc + nq;
    const u32 depth = sh.depth;
    for (u32 i = 0; i < nq; i++) cpos[i] = q[i];
    u32 m = sort_unique(cpos, nq);
    for (u32 i = 0; i < nq; i++) qnode[i] = (u32)find(cpos, m, q[i]);
    u32 vi = 0, wi = 0, np = 0;
    u32 nc = sh.ncols(depth);
    u32 h8[8], l8[8], r8[8];
    for (u32 k = 0; k < m; k++) {
        if (vi + nc > n_values) return false;
        hash_node2(nullptr, nullptr, values + vi, nc, h8);
        st8(chash, k, h8);
        np += node_perms(true, nc);
        colsrc[k] = vi;
        vi += nc;
    }
    for (u32 i = 0; i < nq; i++) {
        const u32 src = colsrc[qnode[i]];
        for (u32 c = 0; c < nc; c++) path_cols[i * cpp + c] = values[src + c];
    }
    u32 colpos = nc;
    for (u32 h = depth; h-- > 0;) {
        nc = sh.ncols(h);
        u32 j = 0, k = 0;
        while (k < m) {
            const u32 ps = cpos[k];
            const bool pair = k + 1 < m && cpos[k + 1] == (ps ^ 1u);
            if (pair) {
                ld8(chash, k, l8); ld8(chash, k + 1, r8);
                sibsrc[k] = k + 1; sibsrc[k + 1] = k; par[k] = j; par[k + 1] = j;
            } else {
                if (wi >= n_hw) return false;
                const u32 *W = hw + 8 * wi;
                sibsrc[k] = 0x80000000u | wi;
                par[k] = j;
                wi++;
                if (ps & 1u) { cp8(l8, W); ld8(chash, k, r8); } else { ld8(chash, k, l8); cp8(r8, W); }
            }
            if (vi + nc > n_values) return false;
            hash_node2(l8, r8, values + vi, nc, h8);
            st8(phash, j, h8);
            np += node_perms(false, nc);
            ppos[j] = ps >> 1;
            colsrc[j] = vi;
            vi += nc;
            j++;
            k += pair ? 2 : 1;
        }
        for (u32 i = 0; i < nq; i++) {
            const u32 kq = qnode[i];
            const u32 s = sibsrc[kq];
            u32 *dst = path_sib + (size_t)i * sib_stride + (depth - 1 - h) * 8;
            if (s & 0x80000000u) cp8(dst, hw + 8 * (s & 0x7fffffffu)); else ld8(chash, s, dst);
            const u32 pj = par[kq];
            qnode[i] = pj;
            const u32 src = colsrc[pj];
            for (u32 c = 0; c < nc; c++) path_cols[i * cpp + colpos + c] = values[src + c];
        }
        colpos += nc;
        u32 *tp = cpos; cpos = ppos; ppos = tp;
        Strided t = chash; chash = phash; phash = t;
        m = j;
    }
    if (perms) *perms += np;
    ld8(chash, 0, h8);
    return vi == n_values && wi == n_hw && m == 1 && eq8(h8, root);
}

// ---- FRI layer trees (query and sibling both opened; QM31 leaf = 4 words) -------------------------------------
constexpr u32 PAIR_SCRATCH_WORDS_PER_QUERY = 2 * 40;   // five tables of node hashes, 2*nq nodes each
constexpr u32 PAIR_IDX_WORDS_PER_QUERY = 2 * 4;
constexpr u32 MAX_DATA_LAYERS = 3;

// data_mask bit h set <=> evaluations are committed at the layer of log size h (bit `depth` always set).
// vals: the decommitted evaluations in stream order (layer by layer descending, positions ascending, 4 words each).
// Per query i:  self_vals/sib_vals [(i*MAX_DATA_LAYERS + d)*4] for the d-th data layer (descending),
//               sib_hashes [(i*(depth-1) + j)*8], j = depth-1-h for layers h = depth-1 .. 1:
//               the sibling's hash (plain layers) or the sibling's hash *without* its own evaluation (data layers).
HD bool pair_tree(u32 depth, u32 data_mask, const u32 *q, u32 nq, const u32 *vals, u32 n_vals, const u32 *hw, u32 n_hw,
                  const u32 *root, u32 *self_vals, u32 *sib_vals, u32 *sib_hashes, Strided scratch, u32 *idx, u32 *perms) {
    const u32 cap = 2 * nq;
    Strided chash = scratch, nhash = chash.at(8 * cap), ntree = nhash.at(8 * cap), nL = ntree.at(8 * cap), nR = nL.at(8 * cap);
    u32 *qs = idx, *cpos = qs + cap, *npos = cpos + cap, *nval = npos + cap;
    for (u32 i = 0; i < nq; i++) qs[i] = q[i];
    u32 n = sort_unique(qs, nq);
    u32 cm = 0;                    // child table size
    u32 vi = 0, wi = 0, np = 0, d_idx = 0;
    u32 h8[8], t8[8], l8[8], r8[8];
    for (u32 h = depth + 1; h-- > 0;) {
        if (h < depth) {
            for (u32 k = 0; k < n; k++) qs[k] >>= 1;
            n = sort_unique(qs, n);
        }
        const bool data = (data_mask >> h) & 1u;
        u32 m = 0;
        if (data) {
            for (u32 k = 0; k < n; k++) { npos[m++] = qs[k]; npos[m++] = qs[k] ^ 1u; }
            m = sort_unique(npos, m);
        } else {
            if (h == depth) return false;
            for (u32 k = 0; k < n; k++) npos[m++] = qs[k];
        }
        for (u32 a = 0; a < m; a++) {
            const u32 *val = nullptr;
            if (data) {
                if (vi + 4 > n_vals) return false;
                val = vals + vi; nval[a] = vi; vi += 4;
            }
            if (h == depth) {
                hash_node2(nullptr, nullptr, val, 4, h8);
                st8(nhash, a, h8);
                np += 2;
            } else {
                const u32 p = npos[a];
                int li = find(cpos, cm, p << 1), ri = find(cpos, cm, (p << 1) + 1);
                if (li >= 0) ld8(chash, (u32)li, l8); else { if (wi >= n_hw) return false; cp8(l8, hw + 8 * wi++); }
                if (ri >= 0) ld8(chash, (u32)ri, r8); else { if (wi >= n_hw) return false; cp8(r8, hw + 8 * wi++); }
                st8(nL, a, l8); st8(nR, a, r8);
                hash_node2(l8, r8, val, data ? 4 : 0, h8, t8);
                st8(nhash, a, h8); st8(ntree, a, t8);
                np += data ? 3 : 1;
            }
        }
        // per-query extraction
        for (u32 i = 0; i < nq; i++) {
            const u32 qh = q[i] >> (depth - h);
            if (data) {
                int a_self = find(npos, m, qh), a_sib = find(npos, m, qh ^ 1u);
                if (a_self < 0 || (a_sib < 0 && h > 0)) return false;
                const u32 vs = nval[a_self];
                for (int c = 0; c < 4; c++) self_vals[(i * MAX_DATA_LAYERS + d_idx) * 4 + c] = vals[vs + c];
                if (a_sib >= 0) {
                    const u32 vb = nval[a_sib];
                    for (int c = 0; c < 4; c++) sib_vals[(i * MAX_DATA_LAYERS + d_idx) * 4 + c] = vals[vb + c];
                }
                if (h != depth && h >= 1) ld8(ntree, (u32)a_sib, sib_hashes + ((size_t)i * (depth - 1) + (depth - 1 - h)) * 8);
            }
            // the child layer h+1, when it carries no data, takes its sibling from this parent's two children
            const u32 hc = h + 1;
            if (h < depth && hc < depth && !((data_mask >> hc) & 1u)) {
                int pa = find(npos, m, qh);
                if (pa < 0) return false;
                const u32 qc = q[i] >> (depth - hc);
                ld8((qc & 1u) ? nL : nR, (u32)pa, sib_hashes + ((size_t)i * (depth - 1) + (depth - 1 - hc)) * 8);
            }
        }
        if (data) d_idx++;
        // this layer becomes the child table
        for (u32 a = 0; a < m; a++) cpos[a] = npos[a];
        for (u32 a = 0; a < 8 * m; a++) chash[a] = nhash[a];
        cm = m;
    }
    if (perms) *perms += np;
    ld8(chash, 0, h8);
    return vi == n_vals && wi == n_hw && cm == 1 && eq8(h8, root);
}

// SinglePairMerkleProof::verify for one query (components/hints/src/folding.rs:33-91): recompute the root from the
// per-query hints.  Returns the root in out[8].
HD void pair_path_root(u32 depth, u32 data_mask, u32 qpos, const u32 *self_vals, const u32 *sib_vals, const u32 *sib_hashes,
                       u32 *out, u32 *perms, u32 *sink = nullptr, size_t in_delta = 0) {
    u32 self_h[8], sib_h[8];
    u32 d_idx = 0, np = 4;
    u32 **sk = sink ? &sink : nullptr;
    hash_node2(nullptr, nullptr, self_vals, 4, self_h, nullptr, sk, in_delta);
    hash_node2(nullptr, nullptr, sib_vals, 4, sib_h, nullptr, sk, in_delta);
    d_idx = 1;
    for (u32 i = 0; i < depth; i++) {
        const u32 h = depth - i - 1;
        const bool data = (data_mask >> h) & 1u;
        const u32 *L = ((qpos >> i) & 1u) ? sib_h : self_h, *R = ((qpos >> i) & 1u) ? self_h : sib_h;
        u32 nxt[8];
        if (!data) {
            hash_node2(L, R, nullptr, 0, nxt, nullptr, sk, in_delta);
            np += 1;
            if (i != depth - 1) cp8(sib_h, sib_hashes + 8 * i);
        } else {
            hash_node2(L, R, self_vals + 4 * d_idx, 4, nxt, nullptr, sk, in_delta);
            np += 3;
            if (h >= 1) {
                // sibling = rate(perm(sibling tree hash || capacity(sibling evaluation)))
                u32 st[16];
                for (int k = 0; k < 4; k++) st[k] = sib_vals[4 * d_idx + k];
                for (int k = 4; k < 16; k++) st[k] = 0;
                tap_in(sk, st, in_delta);
                permute_mem(st);
                tap(sk, st);
                for (int k = 0; k < 8; k++) st[k] = sib_hashes[8 * i + k];
                tap_in(sk, st, in_delta);
                permute_mem(st);
                tap(sk, st);
                cp8(sib_h, st);
                np += 2;
            }
            d_idx++;
        }
        cp8(self_h, nxt);
    }
    cp8(out, self_h);
    if (perms) *perms += np;
}
// permutations pair_path_root executes (a function of the tree's shape only)
HD u32 pair_path_perms(u32 depth, u32 data_mask) {
    u32 np = 4;
    for (u32 i = 0; i < depth; i++) {
        const u32 h = depth - i - 1;
        np += ((data_mask >> h) & 1u) ? (h >= 1 ? 5u : 3u) : 1u;
    }
    return np;
}

}  // namespace decommit
