/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc.h).
 * Poseidon2-M31 permutation, stwo hash_node, Poseidon31 channel. */
#include <string.h>
#include "orc.h"
#include "../include/stwo_b200_poseidon2_constants.h"

static const uint32_t DIAG[16] = STWO_P2_DIAG16;
static const uint32_t RC_FIRST[64] = STWO_P2_RC_FIRST;
static const uint32_t RC_PART[14] = STWO_P2_RC_PARTIAL;
static const uint32_t RC_LAST[64] = STWO_P2_RC_LAST;

/* reference primitives/poseidon31/src/implementation.rs:7-18 */
static void mds4(m31 *x) {
    m31 t0 = m31_add(x[0], x[1]);
    m31 t1 = m31_add(x[2], x[3]);
    m31 t2 = m31_add(m31_dbl(x[1]), t1);
    m31 t3 = m31_add(m31_dbl(x[3]), t0);
    m31 t4 = m31_add(m31_dbl(m31_dbl(t1)), t3);
    m31 t5 = m31_add(m31_dbl(m31_dbl(t0)), t2);
    x[0] = m31_add(t3, t5); x[1] = t5; x[2] = m31_add(t2, t4); x[3] = t4;
}
/* reference implementation.rs:20-58: circ(2*M4, M4, M4, M4) */
static void mds16(m31 *s) {
    m31 t[16];
    memcpy(t, s, sizeof t);
    for (int b = 0; b < 4; b++) mds4(t + 4 * b);
    for (int j = 0; j < 4; j++) {
        m31 col = m31_add(m31_add(t[j], t[j + 4]), m31_add(t[j + 8], t[j + 12]));
        for (int b = 0; b < 4; b++) s[4 * b + j] = m31_add(t[4 * b + j], col);
    }
}
static inline m31 pow5(m31 a) { m31 b = m31_mul(a, a); return m31_mul(m31_mul(b, b), a); }

/* reference implementation.rs:108-149 */
void orc_poseidon2_permute(uint32_t s[16]) {
    mds16(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = pow5(m31_add(s[i], RC_FIRST[16 * r + i]));
        mds16(s);
    }
    for (int r = 0; r < 14; r++) {
        s[0] = pow5(m31_add(s[0], RC_PART[r]));
        m31 sum = 0;
        for (int i = 0; i < 16; i++) sum = m31_add(sum, s[i]);
        for (int i = 0; i < 16; i++) s[i] = m31_add(sum, m31_mul(s[i], DIAG[i]));
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = pow5(m31_add(s[i], RC_LAST[16 * r + i]));
        mds16(s);
    }
}

void orc_poseidon2_permute_batch(uint32_t *states, size_t n) {
    for (size_t i = 0; i < n; i++) orc_poseidon2_permute(states + 16 * i);
}

/* column sponge: 8 M31 per chunk, zero padded, capacity chained
 * (reference primitives/merkle/src/lib.rs:141-181) */
void orc_hash_column_get_capacity(const uint32_t *cols, size_t n, uint32_t out[8]) {
    uint32_t st[16];
    uint32_t cap[8] = {0};
    size_t n_chunks = (n + 7) / 8;
    if (n_chunks == 0) n_chunks = 1;
    for (size_t c = 0; c < n_chunks; c++) {
        for (int i = 0; i < 8; i++) st[i] = (8 * c + i < n) ? cols[8 * c + i] : 0;
        memcpy(st + 8, cap, 32);
        orc_poseidon2_permute(st);
        memcpy(cap, st + 8, 32);
    }
    memcpy(out, cap, 32);
}

/* stwo Poseidon31MerkleHasher::hash_node as restated by
 * reference primitives/merkle/src/lib.rs:9-91 and used at
 * components/hints/src/decommit.rs:22-42, folding.rs:33-91 */
void orc_hash_node(const uint32_t *left, const uint32_t *right,
                   const uint32_t *cols, size_t n_cols, uint32_t out[8]) {
    uint32_t st[16];
    if (!left) {                      /* leaf: rate of permute(0^8 || cap(cols)) */
        memset(st, 0, 32);
        orc_hash_column_get_capacity(cols, n_cols, st + 8);
        orc_poseidon2_permute(st);
        memcpy(out, st, 32);
        return;
    }
    memcpy(st, left, 32);
    memcpy(st + 8, right, 32);
    orc_poseidon2_permute(st);
    if (n_cols) {                     /* lib.rs:12-20: permute(hash_tree || cap(cols)) */
        orc_hash_column_get_capacity(cols, n_cols, st + 8);
        orc_poseidon2_permute(st);
    }
    memcpy(out, st, 32);
}

uint64_t orc_merkle_build(const uint32_t *leaves, uint32_t log_n, uint32_t n_cols,
                          uint32_t *nodes) {
    uint64_t perms = 0;
    size_t n = (size_t)1 << log_n;
    uint32_t *layer = nodes + (n - 1) * 8;
    for (size_t i = 0; i < n; i++) orc_hash_node(NULL, NULL, leaves + i * n_cols, n_cols, layer + 8 * i);
    perms += n * ((n_cols + 7) / 8 + 1);
    for (uint32_t k = log_n; k-- > 0;) {
        size_t m = (size_t)1 << k;
        uint32_t *child = nodes + (2 * m - 1) * 8;
        uint32_t *cur = nodes + (m - 1) * 8;
        for (size_t i = 0; i < m; i++)
            orc_hash_node(child + 16 * i, child + 16 * i + 8, NULL, 0, cur + 8 * i);
        perms += m;
    }
    return perms;
}

int orc_merkle_path_verify(const uint32_t *leaf, uint32_t n_cols, uint32_t index,
                           const uint32_t *siblings, uint32_t depth,
                           const uint32_t root[8], uint32_t out_root[8]) {
    uint32_t cur[8];
    orc_hash_node(NULL, NULL, leaf, n_cols, cur);
    for (uint32_t i = 0; i < depth; i++) {
        const uint32_t *sib = siblings + 8 * i;
        if ((index >> i) & 1) orc_hash_node(sib, cur, NULL, 0, cur);
        else orc_hash_node(cur, sib, NULL, 0, cur);
    }
    if (out_root) memcpy(out_root, cur, 32);
    return memcmp(cur, root, 32) == 0;
}

/* Poseidon31 channel, reference primitives/channel/src/lib.rs:23-58 */
void orc_channel_init(orc_channel *c) { memset(c, 0, sizeof *c); }
void orc_channel_mix_root(orc_channel *c, const uint32_t root[8]) {
    uint32_t st[16];
    memcpy(st, root, 32); memcpy(st + 8, c->digest, 32);
    orc_poseidon2_permute(st);
    memcpy(c->digest, st + 8, 32); c->n_sent = 0; c->n_perms++;
}
void orc_channel_mix_felts2(orc_channel *c, const uint32_t a[4], const uint32_t b[4]) {
    uint32_t h[8];
    memcpy(h, a, 16);
    if (b) memcpy(h + 4, b, 16); else memset(h + 4, 0, 16);
    orc_channel_mix_root(c, h);
}
void orc_channel_draw(orc_channel *c, uint32_t out[8]) {
    uint32_t st[16] = {0};
    st[0] = c->n_sent++;
    memcpy(st + 8, c->digest, 32);
    orc_poseidon2_permute(st);
    memcpy(out, st, 32); c->n_perms++;
}

/* Mixed-degree authentication path: columns injected at inner layers
 * (reference components/recursive/data_structures/src/lib.rs:315-354;
 * components/hints/src/decommit.rs:22-42).  n_cols[h] = columns at the layer of
 * log size h (h = depth is the leaf layer); cols = leaf layer values first,
 * then each injected layer in descending h; siblings leaf level first. */
void orc_merkle_path_root_mixed(uint32_t depth, const uint32_t *n_cols, uint32_t index,
                                const uint32_t *cols, const uint32_t *siblings, uint32_t out[8]) {
    uint32_t cur[8];
    orc_hash_node(NULL, NULL, cols, n_cols[depth], cur);
    cols += n_cols[depth];
    for (uint32_t i = 0; i < depth; i++) {
        uint32_t h = depth - 1 - i;
        const uint32_t *sib = siblings + 8 * i;
        if ((index >> i) & 1) orc_hash_node(sib, cur, cols, n_cols[h], cur);
        else orc_hash_node(cur, sib, cols, n_cols[h], cur);
        cols += n_cols[h];
    }
    memcpy(out, cur, 32);
}

/* ---- multi-threaded drivers for the CPU baseline (pthreads; the reference itself is
 * single-threaded -- constraint_system/src/lib.rs:33 -- so threads only ever split
 * independent states / nodes / trees) ------------------------------------------------ */
#include <pthread.h>
typedef struct { void (*fn)(size_t, size_t, void *); size_t lo, hi; void *ctx; } orc_job;
static void *orc_job_run(void *p) { orc_job *j = (orc_job *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void orc_parallel_for(size_t n, unsigned n_threads, void (*fn)(size_t, size_t, void *), void *ctx) {
    if (n_threads <= 1 || n < 2 * (size_t)n_threads) { fn(0, n, ctx); return; }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; orc_job jobs[256];
    for (unsigned t = 0; t < n_threads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / n_threads; jobs[t].hi = n * (t + 1) / n_threads;
        pthread_create(&th[t], NULL, orc_job_run, &jobs[t]);
    }
    for (unsigned t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}
static void permute_range(size_t lo, size_t hi, void *ctx) {
    uint32_t *s = (uint32_t *)ctx;
    for (size_t i = lo; i < hi; i++) orc_poseidon2_permute(s + 16 * i);
}
void orc_poseidon2_permute_batch_mt(uint32_t *states, size_t n, unsigned n_threads) {
    orc_parallel_for(n, n_threads, permute_range, states);
}
typedef struct { const uint32_t *leaves; uint32_t n_cols; uint32_t *layer; const uint32_t *child; } build_ctx;
static void leaf_range(size_t lo, size_t hi, void *p) {
    build_ctx *c = (build_ctx *)p;
    for (size_t i = lo; i < hi; i++) orc_hash_node(NULL, NULL, c->leaves + i * c->n_cols, c->n_cols, c->layer + 8 * i);
}
static void node_range(size_t lo, size_t hi, void *p) {
    build_ctx *c = (build_ctx *)p;
    for (size_t i = lo; i < hi; i++) orc_hash_node(c->child + 16 * i, c->child + 16 * i + 8, NULL, 0, c->layer + 8 * i);
}
uint64_t orc_merkle_build_mt(const uint32_t *leaves, uint32_t log_n, uint32_t n_cols, uint32_t *nodes, unsigned n_threads) {
    size_t n = (size_t)1 << log_n;
    build_ctx c = { leaves, n_cols, nodes + (n - 1) * 8, NULL };
    orc_parallel_for(n, n_threads, leaf_range, &c);
    uint64_t perms = n * ((n_cols + 7) / 8 + 1);
    for (uint32_t k = log_n; k-- > 0;) {
        size_t m = (size_t)1 << k;
        c.child = nodes + (2 * m - 1) * 8; c.layer = nodes + (m - 1) * 8;
        orc_parallel_for(m, n_threads, node_range, &c);
        perms += m;
    }
    return perms;
}
typedef struct { uint32_t depth; const uint32_t *n_cols; const uint32_t *index, *cols, *sib, *roots; uint32_t cpp; uint8_t *verdict; } path_ctx;
static void path_range(size_t lo, size_t hi, void *p) {
    path_ctx *c = (path_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        uint32_t r[8];
        orc_merkle_path_root_mixed(c->depth, c->n_cols, c->index[i], c->cols + i * c->cpp, c->sib + i * c->depth * 8, r);
        c->verdict[i] = memcmp(r, c->roots, 32) == 0;
    }
}
/* all paths against roots[0..8] */
void orc_merkle_paths_verify_mt(uint32_t depth, const uint32_t *n_cols, size_t n_paths, const uint32_t *index,
                                const uint32_t *cols, const uint32_t *sib, const uint32_t *root, uint8_t *verdict,
                                unsigned n_threads) {
    uint32_t cpp = 0;
    for (uint32_t h = 0; h <= depth; h++) cpp += n_cols[h];
    path_ctx c = { depth, n_cols, index, cols, sib, root, cpp, verdict };
    orc_parallel_for(n_paths, n_threads, path_range, &c);
}
