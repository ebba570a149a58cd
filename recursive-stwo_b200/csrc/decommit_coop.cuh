// Cooperative form of the hint pre-pass (decommit.cuh): a GROUP of lanes rebuilds one tree of one proof.
//
// The batched decommitment is replayed layer by layer as before, but every layer is split into
//   plan    -- which nodes exist, where their children / sibling / column values come from (stream positions are
//              prefix sums over the sorted node list; no hashing), computed lane-parallel + one short scan,
//   hash    -- one node per lane (Poseidon2 sponge / compression), the only expensive part,
//   extract -- one query per lane copies its sibling hash and column values into its per-query path,
// with a group barrier between the phases.  The critical path is then ~depth x (perms per node) permutations instead of
// (nodes per tree) x (perms per node): 45 instead of ~350 for an FRI layer tree of shape S.
//
// Follows the same reference code as decommit.cuh:
//   components/hints/src/decommit.rs:44-183   SinglePathMerkleProof::from_stwo_proof (commitment trees)
//   components/hints/src/folding.rs:93-287    SinglePairMerkleProof::from_stwo_proof (FRI layer trees)
//
// `Co` abstracts the group: lane(), size(), sync().  On the device it is a 16-lane half-warp (CoopHalfWarp in
// verify_kernels.cu); on the host (tests/hostsim) a group of one, so the very same code runs in the CPU test tier.
// `tab` is group-shared scratch (shared memory on the device), `nodes` group-private global scratch for node hashes.
#pragma once
#include "decommit.cuh"

namespace decommit {

struct CoopOne {                                   // host / single-thread group
    HDM u32 lane() const { return 0; }
    HDM u32 size() const { return 1; }
    HDM void sync() const {}
};

constexpr u32 W_FLAG = 0x80000000u;                // source index flag: "take hash_witness[idx & ~W_FLAG]"

HD void load_hash(const u32 *nodes_tab, const u32 *hw, u32 src, u32 out[8]) {
    const u32 *p = (src & W_FLAG) ? hw + 8 * (size_t)(src & ~W_FLAG) : nodes_tab + 8 * (size_t)src;
    for (int i = 0; i < 8; i++) out[i] = p[i];
}

// Where the rebuild leaves, per query, the output state of every permutation of that query's authentication path -- the very
// permutations SinglePathMerkleProofVar::verify / SinglePairMerkleProofVar::verify execute in the verifier circuit
// (components/recursive/data_structures/src/lib.rs:315-354,400-464).  A node of the partial tree is hashed ONCE here
// (like from_stwo_proof, components/hints/src/decommit.rs:44-183, folding.rs:93-287) and its states are handed to every query
// whose path runs through it, in the slot order of merkle::path_root / pair_path_root (verify.cuh HintLayout).
struct PermRec {
    u32 *base;            // slot 0 of query 0 of this tree; nullptr: no record
    u32 per_query;        // slots per query
    u32 *roots;           // nq x 8: the recomputed root per query (= the rebuilt root); may be null
    size_t in_delta = 0;  // != 0: the input state of a permutation is recorded in_delta words beyond its output state
    HDM u32 *slot(u32 query, u32 k) const { return base + ((size_t)query * per_query + k) * 16; }
    // where the state seen by a tap goes: the output record, or (before the permutation) the parallel input record; nullptr = nowhere
    HDM u32 *slot(u32 query, u32 k, bool after) const { return after ? slot(query, k) : in_delta ? slot(query, k) + in_delta : nullptr; }
};
// The queries whose path runs through one node, as a bit set (n_queries <= 128): built once per node, walked once per permutation
// state -- scanning all queries for every state costs 2 x n_queries steps per permutation, a fifth of the permutation itself at 80 queries.
struct QSet {
    u32 w[4];
    HDM void clear() { w[0] = w[1] = w[2] = w[3] = 0; }
    template <class F> HDM void build(u32 nq, F member) {
        for (u32 k = 0; k < 4; k++) {
            u32 m = 0;
            for (u32 b = 0; b < 32 && 32 * k + b < nq; b++) if (member(32 * k + b)) m |= 1u << b;
            w[k] = m;
        }
    }
    template <class F> HDM void each(F f) const {
        for (u32 k = 0; k < 4; k++)
            for (u32 m = w[k]; m; m &= m - 1) {
#if defined(__CUDA_ARCH__)
                f(32 * k + (u32)__ffs((int)m) - 1);
#else
                f(32 * k + (u32)__builtin_ctz(m));
#endif
            }
    }
};
HD void st16(u32 *dst, const u32 *st) {
#if defined(__CUDA_ARCH__)
    uint4 *d = reinterpret_cast<uint4 *>(dst);          // slots are 64-byte aligned
    d[0] = make_uint4(st[0], st[1], st[2], st[3]); d[1] = make_uint4(st[4], st[5], st[6], st[7]);
    d[2] = make_uint4(st[8], st[9], st[10], st[11]); d[3] = make_uint4(st[12], st[13], st[14], st[15]);
#else
    for (int i = 0; i < 16; i++) dst[i] = st[i];
#endif
}

// number of tab words single_tree_coop needs
HD u32 single_tab_words(u32 nq) { return 7 * nq + 8; }
// after single_tree_coop returned: did it walk every level (is the PermRec complete)?
HD bool single_rec_complete(const u32 *tab, u32 nq) { return tab[7 * nq + 5] != 0; }

template <class Co>
HD bool single_tree_coop(const Co &co, const SingleShape &sh, const u32 *q, u32 nq, const u32 *values, u32 n_values, const u32 *hw, u32 n_hw,
                         const u32 *root, u32 *path_cols, u32 cpp, u32 *path_sib, u32 sib_stride, u32 *nodes, u32 *tab, u32 *perms,
                         const PermRec rec = PermRec{nullptr, 0, nullptr}) {
    u32 *cpos = tab, *ppos = cpos + nq, *lsrc = ppos + nq, *rsrc = lsrc + nq, *sibsrc = rsrc + nq, *par = sibsrc + nq, *qnode = par + nq;
    u32 *ctl = qnode + nq;                         // [0] m  [1] wi  [2] vi  [3] fail  [4] perms
    u32 *chash = nodes, *phash = nodes + 8 * (size_t)nq;
    const u32 depth = sh.depth, L = co.lane(), G = co.size();
    // ---- sorted unique leaf positions by rank (lane-parallel), query -> node map
    if (L == 0) { ctl[1] = 0; ctl[2] = 0; ctl[3] = 0; ctl[4] = 0; }
    for (u32 i = L; i < nq; i += G) {                                  // first occurrence of its position?
        u32 first = 1;
        for (u32 j = 0; j < i; j++) if (q[j] == q[i]) { first = 0; break; }
        sibsrc[i] = first;
    }
    co.sync();
    for (u32 i = L; i < nq; i += G) {                                  // rank among the distinct positions
        u32 rank = 0;
        for (u32 j = 0; j < nq; j++) rank += (sibsrc[j] && q[j] < q[i]) ? 1u : 0u;
        qnode[i] = rank;
        if (sibsrc[i]) cpos[rank] = q[i];
    }
    co.sync();
    if (L == 0) { u32 m = 0; for (u32 i = 0; i < nq; i++) m += sibsrc[i]; ctl[0] = m; }
    co.sync();
    u32 m = ctl[0];
    u32 nc = sh.ncols(depth);
    // ---- leaf layer: node k takes values[k * nc ..]
    if ((size_t)m * nc > n_values) { if (L == 0) ctl[3] = 1; }
    co.sync();
    if (!ctl[3]) {
        for (u32 k = L; k < m; k += G) {
            u32 h8[8];
            QSet mine; mine.clear();
            if (rec.base) mine.build(nq, [&](u32 i) { return qnode[i] == k; });
            hash_node2_tap(nullptr, nullptr, values + (size_t)k * nc, nc, h8, nullptr, [&](u32 j, const u32 *st, bool after) {
                if (rec.base && (after || rec.in_delta)) mine.each([&](u32 i) { st16(rec.slot(i, j, after), st); });
            });
            for (int i = 0; i < 8; i++) chash[8 * k + i] = h8[i];
        }
        for (u32 i = L; i < nq; i += G) {
            const u32 src = qnode[i] * nc;
            for (u32 c = 0; c < nc; c++) path_cols[i * cpp + c] = values[src + c];
        }
        if (L == 0) { ctl[2] = m * nc; ctl[4] = m * node_perms(true, nc); }
    }
    co.sync();
    u32 colpos = nc;
    u32 slot_off = node_perms(true, nc);               // path slots used so far (merkle::path_root order)
    for (u32 h = depth; h-- > 0;) {
        nc = sh.ncols(h);
        // ---- plan: a child starts a parent unless it is the odd partner of the previous child
        if (!ctl[3]) {
            for (u32 k = L; k < m; k += G) {
                const u32 ps = cpos[k];
                const bool has_prev = k > 0 && cpos[k - 1] == (ps ^ 1u), has_next = k + 1 < m && cpos[k + 1] == (ps ^ 1u);
                const bool starts = !((ps & 1u) && has_prev);
                lsrc[k] = starts ? 1u : 0u;                           // flags for the scan
                rsrc[k] = (starts && !has_next && !has_prev) ? 1u : 0u;   // needs a witness sibling
            }
        }
        co.sync();
        if (L == 0 && !ctl[3]) {
            // exclusive scans (short: m <= n_queries): parent index and witness index per starting child
            u32 j = 0, w = ctl[1];
            for (u32 k = 0; k < m; k++) {
                const u32 starts = lsrc[k], needw = rsrc[k];
                par[k] = starts ? j : j - 1;
                sibsrc[k] = needw ? (W_FLAG | w) : (starts ? k + 1 : k - 1);
                j += starts; w += needw;
            }
            if (w > n_hw || (size_t)ctl[2] + (size_t)j * nc > n_values) ctl[3] = 1;
            ctl[0] = j; ctl[1] = w;
        }
        co.sync();
        const u32 mp = ctl[0], vi0 = ctl[2];
        if (!ctl[3]) {
            // parent j <- its starting child: recompute sources from the child tables (one child per lane)
            for (u32 k = L; k < m; k += G) {
                const u32 ps = cpos[k];
                const bool starts = !((ps & 1u) && k > 0 && cpos[k - 1] == (ps ^ 1u));
                if (!starts) continue;
                const u32 j = par[k], s = sibsrc[k];
                u32 l8[8], r8[8], h8[8];
                if (ps & 1u) { load_hash(chash, hw, s, l8); load_hash(chash, hw, k, r8); }
                else { load_hash(chash, hw, k, l8); load_hash(chash, hw, s, r8); }
                QSet mine; mine.clear();
                if (rec.base) mine.build(nq, [&](u32 i) { return par[qnode[i]] == j; });
                hash_node2_tap(l8, r8, values + vi0 + (size_t)j * nc, nc, h8, nullptr, [&](u32 jj, const u32 *st, bool after) {
                    if (rec.base && (after || rec.in_delta)) mine.each([&](u32 i) { st16(rec.slot(i, slot_off + jj, after), st); });
                });
                for (int i = 0; i < 8; i++) phash[8 * j + i] = h8[i];
                ppos[j] = ps >> 1;
            }
            // ---- extract: sibling hash of every query's node at this level, then move the query to its parent
            for (u32 i = L; i < nq; i += G) {
                const u32 kq = qnode[i];
                u32 s8[8];
                load_hash(chash, hw, sibsrc[kq], s8);
                u32 *dst = path_sib + (size_t)i * sib_stride + (depth - 1 - h) * 8;
                for (int c = 0; c < 8; c++) dst[c] = s8[c];
                const u32 pj = par[kq];
                const u32 src = vi0 + pj * nc;
                for (u32 c = 0; c < nc; c++) path_cols[i * cpp + colpos + c] = values[src + c];
                lsrc[i] = pj;                                          // staged: qnode is read by other lanes in this phase
            }
        }
        co.sync();
        if (!ctl[3]) {
            for (u32 i = L; i < nq; i += G) qnode[i] = lsrc[i];
            if (L == 0) { ctl[2] = vi0 + mp * nc; ctl[4] += mp * node_perms(false, nc); }
        }
        co.sync();
        colpos += nc;
        slot_off += node_perms(false, nc);
        u32 *tp = cpos; cpos = ppos; ppos = tp;
        u32 *th = chash; chash = phash; phash = th;
        m = mp;
    }
    bool ok = !ctl[3] && ctl[2] == n_values && ctl[1] == n_hw && m == 1 && eq8(chash, root);
    if (perms && L == 0) *perms += ctl[4];
    if (rec.roots && !ctl[3] && m == 1) for (u32 i = L; i < nq; i += G) cp8(rec.roots + 8 * (size_t)i, chash);
    if (L == 0) ctl[5] = ctl[3] ? 0u : 1u;             // the record is complete (every level was walked)
    co.sync();
    return ok;
}

// ---- FRI layer trees ---------------------------------------------------------------------------------------------------
HD u32 pair_tab_words(u32 nq) { return 2 * nq + 5 * 2 * nq + 8; }
HD bool pair_rec_complete(const u32 *tab, u32 nq) { return tab[2 * nq + 5 * 2 * nq + 7] != 0; }

// Same contract as pair_tree (decommit.cuh).  nodes: 5 tables of 8 * 2nq words (child hash, node hash, node tree hash,
// node left child, node right child).
template <class Co>
HD bool pair_tree_coop(const Co &co, u32 depth, u32 data_mask, const u32 *q, u32 nq, const u32 *vals, u32 n_vals, const u32 *hw, u32 n_hw,
                       const u32 *root, u32 *self_vals, u32 *sib_vals, u32 *sib_hashes, u32 *nodes, u32 *tab, u32 *perms,
                       const PermRec rec = PermRec{nullptr, 0, nullptr}) {
    const u32 cap = 2 * nq, L = co.lane(), G = co.size();
    u32 *qs = tab, *qtmp = qs + nq;                                    // sorted unique query positions of the layer
    u32 *cpos = qtmp + nq, *npos = cpos + cap, *lsrc = npos + cap, *rsrc = lsrc + cap, *flag = rsrc + cap;
    u32 *ctl = flag + cap;                                             // [0] n  [1] wi  [2] vi  [3] fail  [4] perms  [5] m  [6] cm
    u32 *chash = nodes, *nhash = chash + 8 * (size_t)cap, *ntree = nhash + 8 * (size_t)cap, *nL = ntree + 8 * (size_t)cap, *nR = nL + 8 * (size_t)cap;
    if (L == 0) {
        for (u32 i = 0; i < nq; i++) qs[i] = q[i];
        ctl[0] = sort_unique(qs, nq);
        ctl[1] = ctl[2] = ctl[3] = ctl[4] = 0; ctl[5] = 0; ctl[6] = 0;
    }
    co.sync();
    u32 d_idx = 0;
    u32 slot_off = 4;                                  // path slots before the level being hashed (pair_path_root order: 2 + 2 leaf perms first)
    for (u32 h = depth + 1; h-- > 0;) {
        const bool data = (data_mask >> h) & 1u;
        // ---- plan (lane 0: the node list is short and sorted; this is integer work only)
        if (L == 0 && !ctl[3]) {
            u32 n = ctl[0];
            if (h < depth) {
                u32 nn = 0;
                for (u32 k = 0; k < n; k++) { const u32 p = qs[k] >> 1; if (nn == 0 || qs[nn - 1] != p) qs[nn++] = p; }
                n = nn; ctl[0] = n;
            }
            u32 m = 0;
            if (data) {
                for (u32 k = 0; k < n; k++) {
                    const u32 e = qs[k] & ~1u;
                    if (m == 0 || npos[m - 1] != (e | 1u)) { npos[m++] = e; npos[m++] = e | 1u; }
                }
            } else {
                if (h == depth) ctl[3] = 1;
                for (u32 k = 0; k < n; k++) npos[m++] = qs[k];
            }
            ctl[5] = m;
            if (data && (size_t)ctl[2] + 4 * (size_t)m > n_vals) ctl[3] = 1;
        }
        co.sync();
        const u32 m = ctl[5], cm = ctl[6], vi0 = ctl[2];
        // children of every node: index in the child table or "missing" (lane-parallel binary searches)
        if (!ctl[3] && h < depth) {
            for (u32 a = L; a < m; a += G) {
                const u32 p = npos[a];
                const int li = find(cpos, cm, p << 1), ri = find(cpos, cm, (p << 1) + 1);
                lsrc[a] = li >= 0 ? (u32)li : W_FLAG;
                rsrc[a] = ri >= 0 ? (u32)ri : W_FLAG;
            }
        }
        co.sync();
        if (L == 0 && !ctl[3] && h < depth) {                          // witness indices: left before right, nodes in order
            u32 w = ctl[1];
            for (u32 a = 0; a < m; a++) {
                if (lsrc[a] == W_FLAG) lsrc[a] = W_FLAG | w++;
                if (rsrc[a] == W_FLAG) rsrc[a] = W_FLAG | w++;
            }
            if (w > n_hw) ctl[3] = 1;
            ctl[1] = w;
        }
        co.sync();
        // ---- hash: one node per lane
        if (!ctl[3]) {
            for (u32 a = L; a < m; a += G) {
                const u32 *val = data ? vals + vi0 + 4 * (size_t)a : nullptr;
                u32 h8[8], t8[8], l8[8], r8[8];
                const u32 pos = npos[a], sh_q = depth - h;
                if (h == depth) {
                    // a leaf is the self opening of the queries at its position (slots 0, 1) and the sibling opening of the
                    // queries at the other member of its pair (slots 2, 3)
                    QSet self, sibl; self.clear(); sibl.clear();
                    if (rec.base) { self.build(nq, [&](u32 i) { return q[i] == pos; }); sibl.build(nq, [&](u32 i) { return (q[i] ^ 1u) == pos; }); }
                    hash_node2_tap(nullptr, nullptr, val, 4, h8, nullptr, [&](u32 j, const u32 *st, bool after) {
                        if (rec.base && (after || rec.in_delta)) {
                            self.each([&](u32 i) { st16(rec.slot(i, j, after), st); });
                            sibl.each([&](u32 i) { st16(rec.slot(i, 2 + j, after), st); });
                        }
                    });
                    for (int i = 0; i < 8; i++) nhash[8 * a + i] = h8[i];
                } else {
                    load_hash(chash, hw, lsrc[a], l8);
                    load_hash(chash, hw, rsrc[a], r8);
                    // self node of a query: all its permutations; sibling node at a data layer: the two that fold the sibling's
                    // own evaluation into its tree hash (the tree hash itself is a hint of the path)
                    QSet self, sibl; self.clear(); sibl.clear();
                    if (rec.base) {
                        self.build(nq, [&](u32 i) { return (q[i] >> sh_q) == pos; });
                        if (data && h >= 1) sibl.build(nq, [&](u32 i) { return ((q[i] >> sh_q) ^ 1u) == pos; });
                    }
                    hash_node2_tap(l8, r8, val, data ? 4 : 0, h8, t8, [&](u32 j, const u32 *st, bool after) {
                        if (rec.base && (after || rec.in_delta)) {
                            self.each([&](u32 i) { st16(rec.slot(i, slot_off + j, after), st); });
                            if (j >= 1) sibl.each([&](u32 i) { st16(rec.slot(i, slot_off + 2 + j, after), st); });
                        }
                    });
                    for (int i = 0; i < 8; i++) { nhash[8 * a + i] = h8[i]; ntree[8 * a + i] = t8[i]; nL[8 * a + i] = l8[i]; nR[8 * a + i] = r8[i]; }
                }
            }
        }
        co.sync();
        // ---- extract: one query per lane
        if (!ctl[3]) {
            for (u32 i = L; i < nq; i += G) {
                const u32 qh = q[i] >> (depth - h);
                if (data) {
                    const int a_self = find(npos, m, qh), a_sib = find(npos, m, qh ^ 1u);
                    if (a_self < 0 || (a_sib < 0 && h > 0)) { ctl[3] = 1; continue; }
                    for (int c = 0; c < 4; c++) self_vals[(i * MAX_DATA_LAYERS + d_idx) * 4 + c] = vals[vi0 + 4 * (u32)a_self + c];
                    if (a_sib >= 0) for (int c = 0; c < 4; c++) sib_vals[(i * MAX_DATA_LAYERS + d_idx) * 4 + c] = vals[vi0 + 4 * (u32)a_sib + c];
                    if (h != depth && h >= 1) {
                        u32 *dst = sib_hashes + ((size_t)i * (depth - 1) + (depth - 1 - h)) * 8;
                        for (int c = 0; c < 8; c++) dst[c] = ntree[8 * (u32)a_sib + c];
                    }
                }
                const u32 hc = h + 1;
                if (h < depth && hc < depth && !((data_mask >> hc) & 1u)) {
                    const int pa = find(npos, m, qh);
                    if (pa < 0) { ctl[3] = 1; continue; }
                    const u32 qc = q[i] >> (depth - hc);
                    const u32 *src = ((qc & 1u) ? nL : nR) + 8 * (u32)pa;
                    u32 *dst = sib_hashes + ((size_t)i * (depth - 1) + (depth - 1 - hc)) * 8;
                    for (int c = 0; c < 8; c++) dst[c] = src[c];
                }
            }
        }
        co.sync();
        if (!ctl[3]) {
            for (u32 a = L; a < m; a += G) cpos[a] = npos[a];
            if (L == 0) {
                ctl[6] = m;
                if (data) ctl[2] = vi0 + 4 * m;
                ctl[4] += h == depth ? 2 * m : (data ? 3 * m : m);
            }
        }
        co.sync();
        if (data) d_idx++;
        if (h < depth) slot_off += data ? (h >= 1 ? 5u : 3u) : 1u;
        u32 *t = chash; chash = nhash; nhash = t;                      // this layer becomes the child table
    }
    const bool ok = !ctl[3] && ctl[2] == n_vals && ctl[1] == n_hw && ctl[6] == 1 && eq8(chash, root);
    if (perms && L == 0) *perms += ctl[4];
    if (rec.roots && !ctl[3] && ctl[6] == 1) for (u32 i = L; i < nq; i += G) cp8(rec.roots + 8 * (size_t)i, chash);
    if (L == 0) ctl[7] = ctl[3] ? 0u : 1u;             // the record is complete (every level was walked)
    co.sync();
    return ok;
}

}  // namespace decommit
