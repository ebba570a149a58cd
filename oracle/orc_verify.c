/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc.h).
 * Native CPU verifier of a PlonkWithPoseidon stwo proof with Poseidon31 Merkle/channel:
 * the value side of examples/single-proof (reference examples/single-proof/src/main.rs:33-83),
 * i.e. FiatShamirResults -> CompositionCheck -> AnswerResults -> FoldingResults, together
 * with the hint pre-pass (DecommitHints / FirstLayerHints / InnerLayersHints) that re-shapes
 * stwo's batched decommitments into per-query paths.  Each block cites what it follows. */
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* ---- small helpers ---------------------------------------------------------- */
static qm31 qm31_load(const uint32_t *w) { return qm31_mk(w[0], w[1], w[2], w[3]); }
/* combine_ef: v0 + i v1 + u v2 + iu v3  (composition/src/data_structures.rs:143-146) */
static qm31 combine_ef(qm31 a, qm31 b, qm31 c, qm31 d) {
    return qm31_add(qm31_add(a, qm31_shift_i(b)), qm31_add(qm31_shift_j(c), qm31_shift_ij(d)));
}
static qm31 qpow5(qm31 x) { qm31 x2 = qm31_mul(x, x); return qm31_mul(qm31_mul(x2, x2), x); }

typedef struct { qm31 x, y; } qpoint;
/* circle group law over QM31 with an M31 point (primitives/circle/src/lib.rs:236-250) */
static qpoint qpoint_add_m31(qpoint p, cpoint q) {
    qpoint r;
    r.x = qm31_sub(qm31_mul_m31(p.x, q.x), qm31_mul_m31(p.y, q.y));
    r.y = qm31_add(qm31_mul_m31(p.x, q.y), qm31_mul_m31(p.y, q.x));
    return r;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}
static uint32_t sort_dedup(uint32_t *v, uint32_t n) {
    qsort(v, n, 4, cmp_u32);
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; i++) if (m == 0 || v[m - 1] != v[i]) v[m++] = v[i];
    return m;
}
static int find_pos(const uint32_t *v, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (v[mid] < x) lo = mid + 1; else hi = mid; }
    return (lo < n && v[lo] == x) ? (int)lo : -1;
}

/* ---- transcript (components/recursive/fiat_shamir/src/lib.rs:39-131) ----------- */
static void channel_mix_qm31_pair(orc_channel *c, const uint32_t *a, const uint32_t *b) {
    orc_channel_mix_felts2(c, a, b);
}
static qm31 channel_draw_first(orc_channel *c) { uint32_t o[8]; orc_channel_draw(c, o); return qm31_load(o); }

static int transcript(const orc_proof *p, orc_verify_out *o) {
    orc_channel ch;
    orc_channel_init(&ch);
    orc_channel_mix_root(&ch, p->commitments[0]);
    uint32_t f[4] = { p->log_size_plonk, 0, 0, 0 };
    orc_channel_mix_felts2(&ch, f, NULL);
    f[0] = p->log_size_poseidon;
    orc_channel_mix_felts2(&ch, f, NULL);
    orc_channel_mix_root(&ch, p->commitments[1]);
    uint32_t d[8];
    orc_channel_draw(&ch, d);
    o->z = qm31_load(d); o->alpha = qm31_load(d + 4);
    channel_mix_qm31_pair(&ch, p->plonk_total_sum.v, p->poseidon_total_sum.v);
    orc_channel_mix_root(&ch, p->commitments[2]);
    o->random_coeff = channel_draw_first(&ch);
    orc_channel_mix_root(&ch, p->commitments[3]);
    o->oods_t = channel_draw_first(&ch);
    /* CirclePointQM31Var::from_t (primitives/circle/src/lib.rs:204-219) */
    qm31 t2 = qm31_mul(o->oods_t, o->oods_t);
    qm31 inv = qm31_inv(qm31_add(t2, qm31_from_m31(1)));
    o->oods_x = qm31_mul(qm31_sub(qm31_from_m31(1), t2), inv);
    o->oods_y = qm31_mul(qm31_add(o->oods_t, o->oods_t), inv);
    /* sampled values, flattened tree -> column -> mask, two per permutation */
    const uint32_t *pend = NULL;
    for (int t = 0; t < 4; t++)
        for (uint32_t c = 0; c < p->n_cols[t]; c++)
            for (uint32_t m = 0; m < p->n_masks[t][c]; m++) {
                const uint32_t *v = p->sampled[t][c] + 4 * m;
                if (pend) { channel_mix_qm31_pair(&ch, pend, v); pend = NULL; }
                else pend = v;
            }
    if (pend) orc_channel_mix_felts2(&ch, pend, NULL);
    o->after_coeff = channel_draw_first(&ch);
    orc_channel_mix_root(&ch, p->first_layer.commitment);
    o->fri_alphas[0] = channel_draw_first(&ch);
    for (uint32_t i = 0; i < p->n_inner; i++) {
        orc_channel_mix_root(&ch, p->inner[i].commitment);
        o->fri_alphas[i + 1] = channel_draw_first(&ch);
    }
    for (uint64_t i = 0; i < p->n_last_coeffs; i += 2)
        orc_channel_mix_felts2(&ch, p->last_coeffs + 4 * i, i + 1 < p->n_last_coeffs ? p->last_coeffs + 4 * (i + 1) : NULL);
    /* nonce limbs 22/21/21 bits (components/recursive/data_structures/src/lib.rs:197-213) */
    uint32_t nf[4] = { (uint32_t)(p->pow_nonce & ((1u << 22) - 1)), (uint32_t)((p->pow_nonce >> 22) & ((1u << 21) - 1)),
                       (uint32_t)((p->pow_nonce >> 43) & ((1u << 21) - 1)), 0 };
    orc_channel_mix_felts2(&ch, nf, NULL);
    memcpy(o->digest_after_nonce, ch.digest, 32);
    int pow_ok = p->pow_bits >= 32 ? 0 : (ch.digest[0] & ((1u << p->pow_bits) - 1)) == 0;
    uint32_t nq = p->n_queries, got = 0;
    for (uint32_t k = 0; k < (nq + 3) / 4; k++) {
        orc_channel_draw(&ch, d);
        for (int j = 0; j < 8 && got < nq; j++) o->raw_queries[got++] = d[j];
    }
    o->n_transcript_perms = (uint32_t)ch.n_perms;
    return pow_ok;
}

/* ---- OODS composition (components/recursive/composition/src) ------------------ */
typedef struct {
    const orc_proof *p;
    qm31 acc, random_coeff, denom_inv;
    qm31 z, alpha_pow[3];
    /* mask cursors: [interaction] -> next column inside this component's sub-span */
    uint32_t col[3], base[3];
    qm31 frac_num[8], frac_den[8];
    uint32_t n_fracs;
    qm31 cumsum_shift;
} eval_row;

static qm31 mask1(eval_row *e, int tree) {            /* next_interaction_mask(tree, [0]) */
    uint32_t c = e->base[tree] + e->col[tree]++;
    return qm31_load(e->p->sampled[tree][c]);
}
static void add_constraint(eval_row *e, qm31 v) {     /* data_structures.rs:167-170 + :25-27 */
    e->acc = qm31_add(qm31_mul(e->acc, e->random_coeff), qm31_mul(v, e->denom_inv));
}
static void add_to_relation(eval_row *e, qm31 mult, const qm31 *vals, int n) {   /* :148-165 */
    qm31 den = vals[0];                               /* alpha^0 = 1 */
    for (int i = 1; i < n; i++) den = qm31_add(den, qm31_mul(e->alpha_pow[i], vals[i]));
    den = qm31_sub(den, e->z);
    e->frac_num[e->n_fracs] = mult; e->frac_den[e->n_fracs] = den; e->n_fracs++;
}
static int finalize_logup(eval_row *e, uint32_t batch) {   /* :172-210 */
    uint32_t nb = (e->n_fracs + batch - 1) / batch;
    qm31 prev_col = qm31_from_m31(0);
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t lo = b * batch, hi = lo + batch < e->n_fracs ? lo + batch : e->n_fracs;
        qm31 num = e->frac_num[lo], den = e->frac_den[lo];
        for (uint32_t k = lo + 1; k < hi; k++) {
            num = qm31_add(qm31_mul(num, e->frac_den[k]), qm31_mul(e->frac_num[k], den));
            den = qm31_mul(den, e->frac_den[k]);
        }
        uint32_t c0 = e->base[2] + e->col[2];
        e->col[2] += 4;
        const orc_proof *p = e->p;
        if (b + 1 < nb) {
            for (int k = 0; k < 4; k++) if (p->n_masks[2][c0 + k] != 1) return -1;
            qm31 cur = combine_ef(qm31_load(p->sampled[2][c0]), qm31_load(p->sampled[2][c0 + 1]),
                                  qm31_load(p->sampled[2][c0 + 2]), qm31_load(p->sampled[2][c0 + 3]));
            qm31 diff = qm31_sub(cur, prev_col);
            prev_col = cur;
            add_constraint(e, qm31_sub(qm31_mul(diff, den), num));
        } else {
            for (int k = 0; k < 4; k++) if (p->n_masks[2][c0 + k] != 2) return -1;
            qm31 prev_row = combine_ef(qm31_load(p->sampled[2][c0]), qm31_load(p->sampled[2][c0 + 1]),
                                       qm31_load(p->sampled[2][c0 + 2]), qm31_load(p->sampled[2][c0 + 3]));
            qm31 cur = combine_ef(qm31_load(p->sampled[2][c0] + 4), qm31_load(p->sampled[2][c0 + 1] + 4),
                                  qm31_load(p->sampled[2][c0 + 2] + 4), qm31_load(p->sampled[2][c0 + 3] + 4));
            qm31 diff = qm31_sub(qm31_sub(cur, prev_row), prev_col);
            qm31 fixed = qm31_add(diff, e->cumsum_shift);
            add_constraint(e, qm31_sub(qm31_mul(fixed, den), num));
        }
    }
    return 0;
}
/* coset_vanishing for a canonic coset: x doubled log_size-1 times (composition/src/lib.rs:18-29) */
static qm31 coset_vanishing(qm31 x, uint32_t log_size) {
    for (uint32_t i = 1; i < log_size; i++) { qm31 sq = qm31_mul(x, x); x = qm31_sub(qm31_add(sq, sq), qm31_from_m31(1)); }
    return x;
}
static void q_apply_m4(qm31 *x) {     /* poseidon.rs:10-24 */
    qm31 t0 = qm31_add(x[0], x[1]), t02 = qm31_add(t0, t0), t1 = qm31_add(x[2], x[3]), t12 = qm31_add(t1, t1);
    qm31 t2 = qm31_add(qm31_add(x[1], x[1]), t1), t3 = qm31_add(qm31_add(x[3], x[3]), t0);
    qm31 t4 = qm31_add(qm31_add(t12, t12), t3), t5 = qm31_add(qm31_add(t02, t02), t2);
    x[0] = qm31_add(t3, t5); x[1] = t5; x[2] = qm31_add(t2, t4); x[3] = t4;
}
static void q_external(qm31 *s) {     /* poseidon.rs:28-50 */
    for (int i = 0; i < 4; i++) q_apply_m4(s + 4 * i);
    for (int j = 0; j < 4; j++) {
        qm31 t = qm31_add(qm31_add(s[j], s[j + 4]), qm31_add(s[j + 8], s[j + 12]));
        for (int i = 0; i < 4; i++) s[4 * i + j] = qm31_add(s[4 * i + j], t);
    }
}
static void q_internal(qm31 *s) {     /* poseidon.rs:55-66 */
    qm31 sum = s[0];
    for (int i = 1; i < 16; i++) sum = qm31_add(sum, s[i]);
    s[0] = qm31_add(s[0], qm31_add(qm31_add(s[0], s[0]), sum));
    for (int i = 1; i < 16; i++) s[i] = qm31_add(qm31_mul_m31(s[i], 1u << (i + 1)), sum);
}

static int eval_plonk(eval_row *e) {   /* plonk.rs:8-82 */
    qm31 one = qm31_from_m31(1);
    qm31 a_wire = mask1(e, 0), b_wire = mask1(e, 0), c_wire = mask1(e, 0), op = mask1(e, 0);
    qm31 mult_a = mask1(e, 0), mult_b = mask1(e, 0), mult_c = mask1(e, 0);
    qm31 poseidon_wire = mask1(e, 0), mult_poseidon = mask1(e, 0), enforce_c_m31 = mask1(e, 0);
    qm31 v[12];
    for (int i = 0; i < 12; i++) v[i] = mask1(e, 1);
    add_constraint(e, qm31_mul(enforce_c_m31, v[9]));
    add_constraint(e, qm31_mul(enforce_c_m31, v[10]));
    add_constraint(e, qm31_mul(enforce_c_m31, v[11]));
    qm31 a = combine_ef(v[0], v[1], v[2], v[3]), b = combine_ef(v[4], v[5], v[6], v[7]), c = combine_ef(v[8], v[9], v[10], v[11]);
    add_constraint(e, qm31_sub(qm31_sub(c, qm31_mul(op, qm31_add(a, b))), qm31_mul(qm31_mul(qm31_sub(one, op), a), b)));
    qm31 r2[2], r3[3];
    r2[0] = a; r2[1] = a_wire; add_to_relation(e, mult_a, r2, 2);
    r2[0] = b; r2[1] = b_wire; add_to_relation(e, mult_b, r2, 2);
    r2[0] = c; r2[1] = c_wire; add_to_relation(e, mult_c, r2, 2);
    r3[0] = poseidon_wire; r3[1] = a; r3[2] = b; add_to_relation(e, qm31_neg(mult_poseidon), r3, 3);
    return finalize_logup(e, 2);
}

static int eval_poseidon(eval_row *e) {   // poseidon.rs:73-241
    qm31 one = qm31_from_m31(1);
    qm31 is_first = mask1(e, 0), is_last = mask1(e, 0), is_full = mask1(e, 0);
    qm31 not_first = qm31_sub(one, is_first), not_last = qm31_sub(one, is_last), is_partial = qm31_sub(not_first, is_full);
    qm31 round_id = mask1(e, 0);
    qm31 rc0[16], rc1[16];
    for (int i = 0; i < 16; i++) rc0[i] = mask1(e, 0);
    for (int i = 0; i < 16; i++) rc1[i] = mask1(e, 0);
    qm31 ext1 = mask1(e, 0), ext2 = mask1(e, 0), ext1_nz = mask1(e, 0), ext2_nz = mask1(e, 0);
    qm31 swap_addr = rc0[0];
    qm31 in[16], mid[16], out[16], s[16];
    for (int i = 0; i < 16; i++) in[i] = mask1(e, 1);
    for (int i = 0; i < 16; i++) mid[i] = mask1(e, 1);
    for (int i = 0; i < 16; i++) out[i] = mask1(e, 1);
    qm31 swap = mid[0], nswap = qm31_sub(one, swap);
    for (int i = 0; i < 16; i++)
        s[i] = i < 8 ? qm31_add(qm31_mul(in[i], nswap), qm31_mul(in[i + 8], swap))
                     : qm31_add(qm31_mul(in[i - 8], swap), qm31_mul(in[i], nswap));
    q_external(s);
    for (int i = 0; i < 16; i++) add_constraint(e, qm31_mul(is_first, qm31_sub(s[i], out[i])));
    for (int i = 0; i < 16; i++) s[i] = qpow5(qm31_add(in[i], rc0[i]));
    for (int i = 0; i < 16; i++) { add_constraint(e, qm31_mul(is_full, qm31_sub(mid[i], s[i]))); s[i] = mid[i]; }
    q_external(s);
    for (int i = 0; i < 16; i++) s[i] = qpow5(qm31_add(s[i], rc1[i]));
    q_external(s);
    for (int i = 0; i < 16; i++) add_constraint(e, qm31_mul(is_full, qm31_sub(out[i], s[i])));
    for (int i = 0; i < 16; i++) s[i] = in[i];
    for (int r = 0; r < 14; r++) {
        s[0] = qpow5(qm31_add(s[0], rc0[r]));
        add_constraint(e, qm31_mul(is_partial, qm31_sub(mid[r], s[0])));
        s[0] = mid[r];
        q_internal(s);
    }
    for (int i = 0; i < 16; i++) add_constraint(e, qm31_mul(is_partial, qm31_sub(out[i], s[i])));
    qm31 in_left = qm31_add(round_id, round_id), in_right = qm31_add(in_left, one);
    qm31 out_left = qm31_add(in_right, one), out_right = qm31_add(out_left, one);
    qm31 r3[3], r2[2];
    r3[0] = qm31_add(qm31_mul(is_first, ext1), qm31_mul(not_first, in_left));
    r3[1] = combine_ef(in[0], in[1], in[2], in[3]); r3[2] = combine_ef(in[4], in[5], in[6], in[7]);
    add_to_relation(e, qm31_sub(qm31_mul(ext1_nz, is_first), not_first), r3, 3);
    r3[0] = qm31_add(qm31_mul(is_first, ext2), qm31_mul(not_first, in_right));
    r3[1] = combine_ef(in[8], in[9], in[10], in[11]); r3[2] = combine_ef(in[12], in[13], in[14], in[15]);
    add_to_relation(e, qm31_sub(qm31_mul(ext2_nz, is_first), not_first), r3, 3);
    r3[0] = qm31_add(qm31_mul(is_last, ext1), qm31_mul(not_last, out_left));
    r3[1] = combine_ef(out[0], out[1], out[2], out[3]); r3[2] = combine_ef(out[4], out[5], out[6], out[7]);
    add_to_relation(e, qm31_add(qm31_mul(ext1_nz, is_last), not_last), r3, 3);
    r3[0] = qm31_add(qm31_mul(is_last, ext2), qm31_mul(not_last, out_right));
    r3[1] = combine_ef(out[8], out[9], out[10], out[11]); r3[2] = combine_ef(out[12], out[13], out[14], out[15]);
    add_to_relation(e, qm31_add(qm31_mul(ext2_nz, is_last), not_last), r3, 3);
    r2[0] = swap; r2[1] = swap_addr;
    add_to_relation(e, qm31_mul(is_first, not_last), r2, 2);
    return finalize_logup(e, 3);
}

static int oods_check(const orc_proof *p, orc_verify_out *o) {
    if (p->n_cols[0] != 50 || p->n_cols[1] != 60 || p->n_cols[2] != 16 || p->n_cols[3] != 8) return -1;
    for (int t = 0; t < 4; t++)
        for (uint32_t c = 0; c < p->n_cols[t]; c++)
            if (p->n_masks[t][c] != ((t == 2 && (c & 4)) ? 2u : 1u)) return -1;
    eval_row e;
    memset(&e, 0, sizeof e);
    e.p = p; e.random_coeff = o->random_coeff; e.z = o->z;
    e.alpha_pow[0] = qm31_from_m31(1); e.alpha_pow[1] = o->alpha; e.alpha_pow[2] = qm31_mul(o->alpha, o->alpha);
    e.acc = qm31_from_m31(0);
    /* Plonk component: tree0 cols 0-9, tree1 0-11, tree2 0-7 */
    e.denom_inv = qm31_inv(coset_vanishing(o->oods_x, p->log_size_plonk));
    e.cumsum_shift = qm31_mul_m31(p->plonk_total_sum, m31_inv(m31_pow(2, p->log_size_plonk)));
    if (eval_plonk(&e)) return -1;
    /* Poseidon component: tree0 10-49, tree1 12-59, tree2 8-15 */
    e.base[0] = 10; e.base[1] = 12; e.base[2] = 8; e.col[0] = e.col[1] = e.col[2] = 0; e.n_fracs = 0;
    e.denom_inv = qm31_inv(coset_vanishing(o->oods_x, p->log_size_poseidon));
    e.cumsum_shift = qm31_mul_m31(p->poseidon_total_sum, m31_inv(m31_pow(2, p->log_size_poseidon)));
    if (eval_poseidon(&e)) return -1;
    o->oods_computed = e.acc;
    const uint32_t *const *sv = p->sampled[3];
    qm31 left = combine_ef(qm31_load(sv[0]), qm31_load(sv[1]), qm31_load(sv[2]), qm31_load(sv[3]));
    qm31 right = combine_ef(qm31_load(sv[4]), qm31_load(sv[5]), qm31_load(sv[6]), qm31_load(sv[7]));
    uint32_t bound = o->max_first_log - p->log_blowup + 1;      /* composition_log_degree_bound */
    qm31 x = o->oods_x;
    for (uint32_t i = 0; i + 2 < bound; i++) { qm31 sq = qm31_mul(x, x); x = qm31_sub(qm31_add(sq, sq), qm31_from_m31(1)); }
    o->oods_expected = qm31_add(left, qm31_mul(right, x));
    return qm31_eq(o->oods_computed, o->oods_expected) ? 0 : 1;
}

/* ---- batched -> per-query decommitment, single-value trees ----------------------------
 * SinglePathMerkleProof::from_stwo_proof + verify (components/hints/src/decommit.rs:22-184). */
typedef struct { uint32_t pos; uint32_t hash[8]; const uint32_t *cols; } dnode;
typedef struct {
    uint32_t depth;
    uint32_t n_cols[33];                 /* by layer log size */
    uint32_t n_layer[33];                /* nodes known per layer (queried-path nodes + witnesses) */
    dnode *layer[33];                    /* sorted by pos */
} partial_tree;

static dnode *layer_find(const partial_tree *t, uint32_t h, uint32_t pos) {
    uint32_t lo = 0, hi = t->n_layer[h];
    while (lo < hi) { uint32_t m = (lo + hi) / 2; if (t->layer[h][m].pos < pos) lo = m + 1; else hi = m; }
    return (lo < t->n_layer[h] && t->layer[h][lo].pos == pos) ? &t->layer[h][lo] : NULL;
}
static int cmp_dnode(const void *a, const void *b) { return cmp_u32(&((const dnode *)a)->pos, &((const dnode *)b)->pos); }

/* returns 0 ok / 1 reject (streams not exactly consumed, root mismatch) */
static int single_tree_rebuild(uint32_t depth, const uint32_t *n_cols_by_log, const uint32_t *raw_q, uint32_t nq,
                               const uint32_t *values, uint64_t n_values, const orc_decommitment *dec,
                               partial_tree *t, uint64_t *perms) {
    memset(t, 0, sizeof *t);
    t->depth = depth;
    memcpy(t->n_cols, n_cols_by_log, 33 * 4);
    if (dec->n_column_witness) return 1;
    uint32_t pos[ORC_MAX_QUERIES];
    memcpy(pos, raw_q, nq * 4);
    uint32_t np = sort_dedup(pos, nq);
    uint64_t vi = 0, hi = 0;
    /* leaf layer */
    t->layer[depth] = calloc(2 * np + 2, sizeof(dnode));
    for (uint32_t k = 0; k < np; k++) {
        uint32_t nc = t->n_cols[depth];
        if (vi + nc > n_values) return 1;
        dnode *n = &t->layer[depth][k];
        n->pos = pos[k]; n->cols = values + vi;
        orc_hash_node(NULL, NULL, values + vi, nc, n->hash);
        *perms += (nc + 7) / 8 + 1;
        vi += nc;
    }
    t->n_layer[depth] = np;
    for (uint32_t h = depth; h-- > 0;) {
        uint32_t nchild = t->n_layer[h + 1];
        dnode *child = t->layer[h + 1];
        uint32_t n_known = nchild;      /* witnesses are appended past the sorted prefix, sorted at the end */
        t->layer[h] = calloc(2 * np + 2, sizeof(dnode));
        uint32_t m = 0, nc = t->n_cols[h];
        for (uint32_t k = 0; k < nchild; k++) {
            uint32_t ps = child[k].pos;
            if (m && t->layer[h][m - 1].pos == (ps >> 1)) continue;       /* parent already built */
            if (vi + nc > n_values) return 1;
            const uint32_t *cv = values + vi;
            vi += nc;
            const uint32_t *sib;
            if (k + 1 < nchild && child[k + 1].pos == (ps ^ 1)) sib = child[k + 1].hash;
            else {
                if (hi >= dec->n_hash_witness) return 1;
                dnode *w = &child[n_known++];
                w->pos = ps ^ 1; memcpy(w->hash, dec->hash_witness + 8 * hi, 32); w->cols = NULL;
                sib = w->hash; hi++;
            }
            dnode *n = &t->layer[h][m++];
            n->pos = ps >> 1; n->cols = cv;
            if (ps & 1) orc_hash_node(sib, child[k].hash, cv, nc, n->hash);
            else orc_hash_node(child[k].hash, sib, cv, nc, n->hash);
            *perms += 1 + (nc ? (nc + 7) / 8 + 1 : 0);
        }
        t->n_layer[h] = m;
        t->n_layer[h + 1] = n_known;
        qsort(child, n_known, sizeof(dnode), cmp_dnode);
    }
    if (vi != n_values || hi != dec->n_hash_witness) return 1;
    return 0;
}
static void partial_tree_free(partial_tree *t) { for (int h = 0; h < 33; h++) free(t->layer[h]); }

/* per-query path out of the partial tree + SinglePathMerkleProof::verify */
static int single_path_root(const partial_tree *t, uint32_t q, uint32_t *cols_out, uint32_t *sib_out, uint32_t root_out[8], uint64_t *perms) {
    uint32_t depth = t->depth, nc_total = 0, cur = q;
    for (uint32_t h = depth + 1; h-- > 0;) {
        dnode *n = layer_find(t, h, cur);
        if (!n || (t->n_cols[h] && !n->cols)) return 1;
        memcpy(cols_out + nc_total, n->cols, t->n_cols[h] * 4);
        nc_total += t->n_cols[h];
        if (h > 0) {
            dnode *s = layer_find(t, h, cur ^ 1);
            if (!s) return 1;
            memcpy(sib_out + 8 * (depth - h), s->hash, 32);
        }
        cur >>= 1;
    }
    orc_merkle_path_root_mixed(depth, t->n_cols, q, cols_out, sib_out, root_out);
    for (uint32_t h = 0; h <= depth; h++) *perms += (h < depth) + (t->n_cols[h] ? (t->n_cols[h] + 7) / 8 + 1 : 0);
    return 0;
}

/* ---- FRI pair trees (components/hints/src/folding.rs:21-288) ---------------------------- */
typedef struct { uint32_t pos; uint32_t hash[8]; uint32_t tree_hash[8]; uint32_t val[4]; int has_val; } pnode;
typedef struct { uint32_t depth; uint8_t has_data[33]; uint32_t n_layer[33]; pnode *layer[33]; } pair_tree;
static int cmp_pnode(const void *a, const void *b) { return cmp_u32(&((const pnode *)a)->pos, &((const pnode *)b)->pos); }
static pnode *pfind(const pair_tree *t, uint32_t h, uint32_t pos) {
    uint32_t lo = 0, hi = t->n_layer[h];
    while (lo < hi) { uint32_t m = (lo + hi) / 2; if (t->layer[h][m].pos < pos) lo = m + 1; else hi = m; }
    return (lo < t->n_layer[h] && t->layer[h][lo].pos == pos) ? &t->layer[h][lo] : NULL;
}
/* child hash from the previous layer or the next witness (left before right) */
static const uint32_t *child_hash(pair_tree *t, uint32_t h_child, uint32_t sorted_n, uint32_t *n_known, uint32_t pos,
                                  const orc_decommitment *dec, uint64_t *hi) {
    pnode *L = t->layer[h_child];
    uint32_t lo = 0, up = sorted_n;
    while (lo < up) { uint32_t m = (lo + up) / 2; if (L[m].pos < pos) lo = m + 1; else up = m; }
    if (lo < sorted_n && L[lo].pos == pos) return L[lo].hash;
    for (uint32_t k = sorted_n; k < *n_known; k++) if (L[k].pos == pos) return L[k].hash;
    if (*hi >= dec->n_hash_witness) return NULL;
    pnode *w = &L[(*n_known)++];
    memset(w, 0, sizeof *w);
    w->pos = pos; memcpy(w->hash, dec->hash_witness + 8 * (*hi), 32); (*hi)++;
    return w->hash;
}
static int pair_tree_rebuild(uint32_t depth, const uint8_t *has_data, const uint32_t *leaf_q, uint32_t nq,
                             const uint32_t *values, uint64_t n_values, const orc_decommitment *dec,
                             const uint32_t *root, pair_tree *t, uint64_t *perms) {
    memset(t, 0, sizeof *t);
    t->depth = depth;
    memcpy(t->has_data, has_data, 33);
    if (dec->n_column_witness) return 1;
    uint32_t q[ORC_MAX_QUERIES];
    memcpy(q, leaf_q, nq * 4);
    uint32_t n = nq;
    uint64_t vi = 0, hi = 0;
    for (uint32_t h = depth + 1; h-- > 0;) {
        n = sort_dedup(q, n);
        t->layer[h] = calloc(4 * nq + 4, sizeof(pnode));
        uint32_t m = 0;
        uint32_t sorted_child = h < depth ? t->n_layer[h + 1] : 0, known_child = sorted_child;
        if (has_data[h]) {
            uint32_t ss[2 * ORC_MAX_QUERIES], ns = 0;
            for (uint32_t k = 0; k < n; k++) { ss[ns++] = q[k]; ss[ns++] = q[k] ^ 1; }
            ns = sort_dedup(ss, ns);
            for (uint32_t k = 0; k < ns; k++) {
                if (vi + 4 > n_values) return 1;
                pnode *nd = &t->layer[h][m++];
                nd->pos = ss[k]; memcpy(nd->val, values + vi, 16); nd->has_val = 1; vi += 4;
            }
            for (uint32_t k = 0; k < m; k++) {
                pnode *nd = &t->layer[h][k];
                if (h == depth) { orc_hash_node(NULL, NULL, nd->val, 4, nd->hash); *perms += 2; }
                else {
                    const uint32_t *l = child_hash(t, h + 1, sorted_child, &known_child, nd->pos << 1, dec, &hi);
                    if (!l) return 1;
                    const uint32_t *r = child_hash(t, h + 1, sorted_child, &known_child, (nd->pos << 1) + 1, dec, &hi);
                    if (!r) return 1;
                    orc_hash_node(l, r, NULL, 0, nd->tree_hash);
                    orc_hash_node(l, r, nd->val, 4, nd->hash);
                    *perms += 3;
                }
            }
        } else {
            if (h == depth) return 1;
            for (uint32_t k = 0; k < n; k++) {
                pnode *nd = &t->layer[h][m++];
                nd->pos = q[k];
                const uint32_t *l = child_hash(t, h + 1, sorted_child, &known_child, nd->pos << 1, dec, &hi);
                if (!l) return 1;
                const uint32_t *r = child_hash(t, h + 1, sorted_child, &known_child, (nd->pos << 1) + 1, dec, &hi);
                if (!r) return 1;
                orc_hash_node(l, r, NULL, 0, nd->hash);
                *perms += 1;
            }
        }
        t->n_layer[h] = m;
        if (h < depth) { t->n_layer[h + 1] = known_child; qsort(t->layer[h + 1], known_child, sizeof(pnode), cmp_pnode); }
        for (uint32_t k = 0; k < n; k++) q[k] >>= 1;
    }
    if (vi != n_values || hi != dec->n_hash_witness) return 1;
    if (t->n_layer[0] != 1 || memcmp(t->layer[0][0].hash, root, 32)) return 1;
    return 0;
}
static void pair_tree_free(pair_tree *t) { for (int h = 0; h < 33; h++) free(t->layer[h]); }

/* SinglePairMerkleProof::verify on the path extracted for leaf query q; self/sibling values returned per data layer */
static int pair_path_root(const pair_tree *t, uint32_t q, qm31 *self_vals, qm31 *sib_vals, uint32_t root_out[8], uint64_t *perms,
                          uint32_t *sib_hash_out /* (depth-1) x 8, SinglePairMerkleProof.sibling_hashes, or NULL */) {
    uint32_t depth = t->depth, cur = q;
    uint32_t self_h[8], sib_h[8];
    pnode *s = pfind(t, depth, cur), *b = pfind(t, depth, cur ^ 1);
    if (!s || !b || !s->has_val || !b->has_val) return 1;
    orc_hash_node(NULL, NULL, s->val, 4, self_h);
    orc_hash_node(NULL, NULL, b->val, 4, sib_h);
    *perms += 4;
    self_vals[depth] = qm31_load(s->val); sib_vals[depth] = qm31_load(b->val);
    for (uint32_t i = 0; i < depth; i++) {
        uint32_t h = depth - i - 1;
        uint32_t parent = cur >> 1;
        const uint32_t *cv = NULL;
        pnode *ps = NULL, *pb = NULL;
        if (t->has_data[h]) {
            ps = pfind(t, h, parent); pb = pfind(t, h, parent ^ 1);
            if (!ps || !ps->has_val) return 1;
            if (h > 0 && (!pb || !pb->has_val)) return 1;
            cv = ps->val;
            self_vals[h] = qm31_load(ps->val);
            if (pb) sib_vals[h] = qm31_load(pb->val);
        }
        if (cur & 1) orc_hash_node(sib_h, self_h, cv, cv ? 4 : 0, self_h);
        else orc_hash_node(self_h, sib_h, cv, cv ? 4 : 0, self_h);
        *perms += cv ? 3 : 1;
        if (h > 0) {
            if (!t->has_data[h]) {
                pnode *sb = pfind(t, h, parent ^ 1);
                if (!sb) return 1;
                memcpy(sib_h, sb->hash, 32);
                if (sib_hash_out) memcpy(sib_hash_out + 8 * i, sb->hash, 32);
            } else {
                /* sibling: its tree hash (hash witness of the per-query proof) combined with its own column hash */
                uint32_t st[16];
                memcpy(st, pb->tree_hash, 32);
                if (sib_hash_out) memcpy(sib_hash_out + 8 * i, pb->tree_hash, 32);
                orc_hash_column_get_capacity(pb->val, 4, st + 8);
                orc_poseidon2_permute(st);
                memcpy(sib_h, st, 32);
                *perms += 2;
            }
        }
        cur = parent;
    }
    memcpy(root_out, self_h, 32);
    return 0;
}

/* ---- domain points (primitives/query/src/lib.rs:56-168) ------------------- */
/* absolute point of position q at log size L: half_odds(L).at(bitrev(q >> 1, L - 1)) */
static cpoint absolute_point(uint32_t L, uint32_t q) { return cp_half_odds_at(L, orc_bitrev(q >> 1, L - 1)); }

/* ---- answers (components/recursive/answer/src) ------------------------------------------ */
typedef struct { int shift; uint32_t comp_log; qpoint point; uint32_t n; uint32_t col[160]; qm31 val[160]; } sample_batch;

static _Thread_local orc_hints *g_hints;      /* optional sink for the per-query hints (orc_verify_proof_hints) */

/* Steps 7-9 of the verifier -- the FRI query phase: first-layer pair evaluations + decommitment + circle folds, inner layers, last layer
 * (components/hints/src/folding.rs:296-601; components/recursive/folding/src/lib.rs:12-205; primitives/line/src/lib.rs:39-67).  Needs
 * o->fri_alphas, o->fri_answers, o->n_logs / log_sizes and the query positions per log size.  Returns 1 after setting stage / verdict
 * when a check fails.  Also the whole of orc_fri_verify_synth (the FRI-only instances of BASELINE configs[4] part i). */
static int fri_stage(const orc_proof *p, orc_verify_out *o, uint32_t (*pos_at)[ORC_MAX_QUERIES], uint32_t max_first, uint32_t nq) {
#define FAIL(stage_, verdict_) do { o->stage = (stage_); o->verdict = (verdict_); return 1; } while (0)
    /* 7. FRI first layer: rebuild pair evaluations, decommit, circle folds
     *    (hints/folding.rs:296-452; recursive/folding/src/lib.rs:22-90) */
    static _Thread_local qm31 self_v[ORC_MAX_QUERIES][33], sib_v[ORC_MAX_QUERIES][33];
    {
        uint64_t wi = 0;
        static _Thread_local uint32_t vals[ORC_MAX_LOGS * 2 * ORC_MAX_QUERIES * 4];
        uint64_t nv = 0;
        uint8_t has_data[33] = {0};
        for (uint32_t g = 0; g < o->n_logs; g++) {
            uint32_t L = o->log_sizes[g];
            has_data[L] = 1;
            /* sorted unique positions and the answer of each (first occurrence) */
            uint32_t sp[ORC_MAX_QUERIES];
            memcpy(sp, pos_at[L], nq * 4);
            uint32_t ns = sort_dedup(sp, nq);
            for (uint32_t k = 0; k < ns;) {
                uint32_t start = (sp[k] >> 1) << 1;
                for (uint32_t e = start; e < start + 2; e++) {
                    qm31 v;
                    if (k < ns && sp[k] == e) {
                        uint32_t i = 0;
                        while (pos_at[L][i] != e) i++;
                        v = o->fri_answers[g][i];
                        k++;
                    } else {
                        if (wi >= p->first_layer.n_fri_witness) FAIL(ORC_STAGE_FRI_FIRST, 1);
                        v = qm31_load(p->first_layer.fri_witness + 4 * wi++);
                    }
                    memcpy(vals + nv, v.v, 16); nv += 4;
                }
            }
        }
        if (wi != p->first_layer.n_fri_witness) FAIL(ORC_STAGE_FRI_FIRST, 1);
        pair_tree pt;
        int bad = pair_tree_rebuild(max_first, has_data, pos_at[max_first], nq, vals, nv, &p->first_layer.decommitment,
                                    p->first_layer.commitment, &pt, &o->n_perms_hints);
        for (uint32_t i = 0; i < nq && !bad; i++) {
            bad = pair_path_root(&pt, pos_at[max_first][i], self_v[i], sib_v[i], o->path_roots[4][i], &o->n_perms_paths,
                                 g_hints ? g_hints->pair_sib_hash[0][i][0] : NULL);
            if (g_hints && !bad) {
                g_hints->pair_depth[0] = max_first;
                memcpy(g_hints->pair_has_data[0], has_data, 33);
                for (uint32_t h = 0; h <= max_first; h++) if (has_data[h]) {
                    memcpy(g_hints->pair_self[0][i][h], self_v[i][h].v, 16); memcpy(g_hints->pair_sib[0][i][h], sib_v[i][h].v, 16);
                }
            }
            if (!bad && memcmp(o->path_roots[4][i], p->first_layer.commitment, 32)) bad = 1;
        }
        pair_tree_free(&pt);
        if (bad) FAIL(ORC_STAGE_FRI_FIRST, 1);
        for (uint32_t g = 0; g < o->n_logs; g++) {
            uint32_t L = o->log_sizes[g];
            for (uint32_t i = 0; i < nq; i++) {
                uint32_t q = pos_at[L][i];
                /* self column must equal the computed answer (recursive/folding/src/lib.rs:36-54) */
                if (!qm31_eq(self_v[i][L], o->fri_answers[g][i])) FAIL(ORC_STAGE_FRI_FIRST, 1);
                cpoint pt2 = cp_dbl(absolute_point(L, q));
                m31 y_inv = m31_inv(pt2.y);
                qm31 l = (q & 1) ? sib_v[i][L] : self_v[i][L], r = (q & 1) ? self_v[i][L] : sib_v[i][L];
                qm31 nl = qm31_add(l, r), nr = qm31_mul_m31(qm31_sub(l, r), y_inv);
                o->circle_folds[g][i] = qm31_add(nl, qm31_mul(nr, o->fri_alphas[max_first - L]));
            }
        }
    }

    /* 8. inner layers (hints/folding.rs:459-601; recursive/folding/src/lib.rs:122-192) */
    {
        qm31 folded[ORC_MAX_QUERIES];
        for (uint32_t i = 0; i < nq; i++) folded[i] = qm31_from_m31(0);
        uint32_t log_size = max_first;
        for (uint32_t li = 0; li < p->n_inner; li++) {
            for (uint32_t g = 0; g < o->n_logs; g++)
                if (o->log_sizes[g] == log_size) {
                    qm31 a2 = qm31_mul(o->fri_alphas[li], o->fri_alphas[li]);
                    for (uint32_t i = 0; i < nq; i++) folded[i] = qm31_add(qm31_mul(a2, folded[i]), o->circle_folds[g][i]);
                }
            log_size -= 1;
            const orc_fri_layer *layer = &p->inner[li];
            /* decommitted values: for each sorted unique position, (left, right) with missing siblings from fri_witness */
            uint32_t sp[ORC_MAX_QUERIES];
            memcpy(sp, pos_at[log_size], nq * 4);
            uint32_t ns = sort_dedup(sp, nq);
            static _Thread_local uint32_t vals[2 * ORC_MAX_QUERIES * 4];
            uint64_t nv = 0, wi = 0;
            uint32_t last_pair = 0xffffffffu;
            for (uint32_t k = 0; k < ns; k++) {
                uint32_t e = sp[k];
                uint32_t i = 0;
                while (pos_at[log_size][i] != e) i++;
                qm31 v = folded[i], sv;
                int sib_known = find_pos(sp, ns, e ^ 1);
                if (sib_known >= 0) { uint32_t j = 0; while (pos_at[log_size][j] != (e ^ 1)) j++; sv = folded[j]; }
                else {
                    if (wi >= layer->n_fri_witness) FAIL(ORC_STAGE_FRI_INNER, 1);
                    sv = qm31_load(layer->fri_witness + 4 * wi++);
                }
                if ((e >> 1) != last_pair) {
                    qm31 l = (e & 1) ? sv : v, r = (e & 1) ? v : sv;
                    memcpy(vals + nv, l.v, 16); memcpy(vals + nv + 4, r.v, 16); nv += 8;
                    last_pair = e >> 1;
                }
            }
            if (wi != layer->n_fri_witness) FAIL(ORC_STAGE_FRI_INNER, 1);
            uint8_t has_data[33] = {0};
            has_data[log_size] = 1;
            pair_tree pt;
            int bad = pair_tree_rebuild(log_size, has_data, pos_at[log_size], nq, vals, nv, &layer->decommitment,
                                        layer->commitment, &pt, &o->n_perms_hints);
            for (uint32_t i = 0; i < nq && !bad; i++) {
                bad = pair_path_root(&pt, pos_at[log_size][i], self_v[i], sib_v[i], o->path_roots[5 + li][i], &o->n_perms_paths,
                                     g_hints ? g_hints->pair_sib_hash[1 + li][i][0] : NULL);
                if (g_hints && !bad) {
                    g_hints->pair_depth[1 + li] = log_size;
                    memcpy(g_hints->pair_has_data[1 + li], has_data, 33);
                    memcpy(g_hints->pair_self[1 + li][i][log_size], self_v[i][log_size].v, 16);
                    memcpy(g_hints->pair_sib[1 + li][i][log_size], sib_v[i][log_size].v, 16);
                }
                if (!bad && memcmp(o->path_roots[5 + li][i], layer->commitment, 32)) bad = 1;
            }
            pair_tree_free(&pt);
            if (bad) FAIL(ORC_STAGE_FRI_INNER, 1);
            for (uint32_t i = 0; i < nq; i++) {
                uint32_t q = pos_at[log_size][i];
                if (!qm31_eq(folded[i], self_v[i][log_size])) FAIL(ORC_STAGE_FRI_INNER, 1);
                m31 x_inv = m31_inv(absolute_point(log_size, q).x);
                qm31 l = (q & 1) ? sib_v[i][log_size] : self_v[i][log_size], r = (q & 1) ? self_v[i][log_size] : sib_v[i][log_size];
                qm31 nl = qm31_add(l, r), nr = qm31_mul_m31(qm31_sub(l, r), x_inv);
                folded[i] = qm31_add(nl, qm31_mul(nr, o->fri_alphas[li + 1]));
                o->line_folds[li][i] = folded[i];
            }
        }
        /* 9. last layer (recursive/folding/src/lib.rs:194-204; primitives/line/src/lib.rs:39-67) */
        for (uint32_t i = 0; i < nq; i++) {
            uint32_t q = pos_at[log_size][i];
            cpoint ab = absolute_point(log_size, q);
            m31 x = m31_sub(m31_mul(ab.x, ab.x), m31_mul(ab.y, ab.y));
            uint32_t lg = p->log_last;
            qm31 eval;
            if (p->n_last_coeffs == 1) eval = qm31_load(p->last_coeffs);
            else {
                m31 dbl[32];
                dbl[0] = x;
                for (uint32_t k = 1; k < lg; k++) { m31 sq = m31_mul(dbl[k - 1], dbl[k - 1]); dbl[k] = m31_sub(m31_add(sq, sq), 1); }
                /* fold(values, factors): lhs + rhs * factors[0], recursively -> iterative from the innermost factor */
                static _Thread_local qm31 buf[1 << 12];
                if (lg > 12) FAIL(ORC_STAGE_PARSE, 1);
                uint32_t n = 1u << lg;
                for (uint32_t k = 0; k < n; k++) buf[k] = qm31_load(p->last_coeffs + 4 * k);
                for (uint32_t lev = lg; lev-- > 0;) {
                    /* adjacent pairs at the deepest level use the LAST factor */
                    n >>= 1;
                    for (uint32_t k = 0; k < n; k++) buf[k] = qm31_add(buf[2 * k], qm31_mul_m31(buf[2 * k + 1], dbl[lev]));
                }
                eval = buf[0];
            }
            o->last_layer_evals[i] = eval;
            if (!qm31_eq(folded[i], eval)) FAIL(ORC_STAGE_FRI_LAST, 1);
        }
    }
    return 0;
#undef FAIL
}

int orc_verify_proof(const uint8_t *blob, size_t len, const uint32_t *input_idx, const uint32_t *input_vals,
                     uint32_t n_inputs, orc_verify_out *o) {
    static _Thread_local orc_proof P;
    orc_proof *p = &P;
    memset(o, 0, sizeof *o);
    o->verdict = 1;
#define FAIL(stage_, verdict_) do { o->stage = (stage_); o->verdict = (verdict_); return 0; } while (0)
    if (orc_proof_parse(blob, len, p) != 0) FAIL(ORC_STAGE_PARSE, 1);
    const uint32_t nq = p->n_queries, blow = p->log_blowup;
    if (p->n_last_coeffs != (1ull << p->log_last) || p->n_inner + 1 > ORC_MAX_INNER) FAIL(ORC_STAGE_PARSE, 1);
    const uint32_t max_first = p->log_last + blow + 1 + p->n_inner;
    /* untrusted header words: bounded before they are added, shifted by or looped over */
    if (p->log_size_plonk == 0 || p->log_size_plonk > 28 || p->log_size_poseidon == 0 || p->log_size_poseidon > 28 || blow == 0 || blow > 16 ||
        p->log_last > 12 || p->pow_bits >= 32)
        FAIL(ORC_STAGE_PARSE, 1);
    /* FRI layer count vs column bounds: the composition columns are committed at composition_log_degree_bound - 1 + blowup
     * (components/hints/src/fiat_shamir.rs:130-135), bound = max(log_size_plonk + 2, log_size_poseidon + 3) (constraint degrees
     * 3 and 6 of the two components), and FriVerifier::commit (fiat_shamir.rs:177-183; stwo core/fri.rs, dependency absent from
     * the tree) fails with InvalidNumFriLayers unless the inner layers fold that down to log_last + blowup. */
    {
        const uint32_t a = p->log_size_plonk + 1, b = p->log_size_poseidon + 2;
        if (max_first > 29 || max_first != (a > b ? a : b) + blow) FAIL(ORC_STAGE_PARSE, 1);
    }
    o->max_first_log = max_first; o->n_inner = p->n_inner; o->n_queries = nq;

    /* 1. transcript + PoW */
    if (!transcript(p, o)) FAIL(ORC_STAGE_POW, 1);
    o->n_perms_paths += o->n_transcript_perms;

    /* 2. logup total sum (fiat_shamir/src/lib.rs:133-141) */
    {
        qm31 sum = qm31_from_m31(0);
        for (uint32_t i = 0; i < n_inputs; i++) {
            qm31 v = qm31_load(input_vals + 4 * i);
            qm31 t = qm31_sub(qm31_add(v, qm31_mul(qm31_from_m31(input_idx[i]), o->alpha)), o->z);
            if (qm31_is_zero(t)) FAIL(ORC_STAGE_LOGUP, 1);
            sum = qm31_add(sum, qm31_inv(t));
        }
        sum = qm31_add(qm31_add(sum, p->poseidon_total_sum), p->plonk_total_sum);
        if (!qm31_is_zero(sum)) FAIL(ORC_STAGE_LOGUP, 1);
    }

    /* 3. OODS */
    {
        int r = oods_check(p, o);
        if (r < 0) FAIL(ORC_STAGE_PARSE, 1);
        if (r > 0) FAIL(ORC_STAGE_OODS, 1);
    }

    /* 4. query positions per log size (hints/fiat_shamir.rs:239-252) */
    const uint32_t log_plonk = p->log_size_plonk + blow, log_pos = p->log_size_poseidon + blow;
    {
        uint32_t ls[3] = { max_first, log_plonk, log_pos };
        qsort(ls, 3, 4, cmp_u32);
        o->n_logs = 0;
        for (int i = 2; i >= 0; i--) if (o->n_logs == 0 || o->log_sizes[o->n_logs - 1] != ls[i]) o->log_sizes[o->n_logs++] = ls[i];
    }
    uint32_t pos_at[31][ORC_MAX_QUERIES];
    for (uint32_t L = 1; L <= max_first; L++)
        for (uint32_t i = 0; i < nq; i++) pos_at[L][i] = (o->raw_queries[i] & ((1u << max_first) - 1)) >> (max_first - L);
    for (uint32_t g = 0; g < o->n_logs; g++) memcpy(o->query_pos[g], pos_at[o->log_sizes[g]], nq * 4);
    {   /* the reference panics on duplicated queries at the largest size (answer/src/lib.rs:190-195) */
        uint32_t tmp[ORC_MAX_QUERIES];
        memcpy(tmp, pos_at[max_first], nq * 4);
        if (sort_dedup(tmp, nq) != nq) FAIL(ORC_STAGE_UNSUPPORTED, 2);
    }

    /* 5. commitment-tree decommitments -> per-query paths (hints/decommit.rs; data_structures/src/lib.rs:315-354) */
    static _Thread_local uint32_t path_cols[4][ORC_MAX_QUERIES][64];
    {
        const uint32_t split[4][2] = { {10, 40}, {12, 48}, {8, 8}, {0, 0} };
        for (int t = 0; t < 4; t++) {
            uint32_t ncl[33] = {0};
            uint32_t depth;
            if (t < 3) { ncl[log_plonk] += split[t][0]; ncl[log_pos] += split[t][1]; depth = log_plonk > log_pos ? log_plonk : log_pos; }
            else { ncl[max_first] = 8; depth = max_first; }
            partial_tree pt;
            int bad = single_tree_rebuild(depth, ncl, pos_at[depth], nq, p->queried_values[t], p->n_queried_values[t],
                                          &p->decommitments[t], &pt, &o->n_perms_hints);
            if (!bad && (pt.n_layer[0] < 1 || memcmp(pt.layer[0][0].hash, p->commitments[t], 32))) bad = 1;
            for (uint32_t i = 0; i < nq && !bad; i++) {
                uint32_t sib[32 * 8];
                bad = single_path_root(&pt, pos_at[depth][i], path_cols[t][i], sib, o->path_roots[t][i], &o->n_perms_paths);
                if (g_hints && !bad) {
                    g_hints->single_depth[t] = depth;
                    memcpy(g_hints->single_ncols[t], ncl, sizeof ncl);
                    memcpy(g_hints->single_cols[t][i], path_cols[t][i], sizeof path_cols[t][i]);
                    memcpy(g_hints->single_sib[t][i], sib, depth * 32);
                }
                if (!bad && memcmp(o->path_roots[t][i], p->commitments[t], 32)) bad = 1;
            }
            partial_tree_free(&pt);
            if (bad) FAIL(ORC_STAGE_MERKLE, 1);
        }
    }

    /* 6. FRI answers per log size (answer/src/lib.rs:294-382, data_structures.rs) */
    {
        qpoint oods = { o->oods_x, o->oods_y };
        /* column descriptors in flatten order: tree-major; component log size, mask shifts */
        for (uint32_t g = 0; g < o->n_logs; g++) {
            uint32_t L = o->log_sizes[g];
            static _Thread_local sample_batch batches[4];
            uint32_t nb = 0, col_index = 0;
            for (int t = 0; t < 4; t++)
                for (uint32_t c = 0; c < p->n_cols[t]; c++) {
                    static const uint32_t split0[4] = { 10, 12, 8, 0 };
                    uint32_t comp_log, clog;
                    if (t == 3) { comp_log = 0; clog = max_first; }
                    else if (c < split0[t]) { comp_log = p->log_size_plonk; clog = log_plonk; }
                    else { comp_log = p->log_size_poseidon; clog = log_pos; }
                    if (clog != L) continue;
                    for (uint32_t m = 0; m < p->n_masks[t][c]; m++) {
                        int shift = (p->n_masks[t][c] == 2 && m == 0) ? -1 : 0;
                        uint32_t key_log = shift ? comp_log : 0;
                        uint32_t b;
                        for (b = 0; b < nb; b++) if (batches[b].shift == shift && batches[b].comp_log == key_log) break;
                        if (b == nb) {
                            if (nb == 4) FAIL(ORC_STAGE_PARSE, 1);
                            batches[b].shift = shift; batches[b].comp_log = key_log; batches[b].n = 0;
                            /* mask point: oods + shift * step(comp_log), step = gen(comp_log) (answer/src/lib.rs:62-72) */
                            batches[b].point = shift ? qpoint_add_m31(oods, cp_neg(cp_subgroup_gen(comp_log))) : oods;
                            nb++;
                        }
                        if (batches[b].n >= 160) FAIL(ORC_STAGE_PARSE, 1);
                        batches[b].col[batches[b].n] = col_index;
                        batches[b].val[batches[b].n] = qm31_load(p->sampled[t][c] + 4 * m);
                        batches[b].n++;
                    }
                    col_index++;
                }
            /* line coefficients with the running alpha (data_structures.rs:137-189) */
            static _Thread_local qm31 ca[4][160], cb[4][160], cc[4][160];
            qm31 alpha = qm31_mk(0, 0, m31_neg(2), 0);
            for (uint32_t b = 0; b < nb; b++) {
                cm31 y0 = qm31_lo(batches[b].point.y), y1 = qm31_hi(batches[b].point.y);
                for (uint32_t k = 0; k < batches[b].n; k++) {
                    cm31 v0 = qm31_lo(batches[b].val[k]), v1 = qm31_hi(batches[b].val[k]);
                    cm31 bb = cm31_sub(cm31_mul(v0, y1), cm31_mul(v1, y0));
                    ca[b][k] = qm31_mul_cm31(alpha, v1);
                    cb[b][k] = qm31_mul_cm31(alpha, bb);
                    cc[b][k] = qm31_mul_cm31(alpha, y1);
                    alpha = qm31_mul(alpha, o->after_coeff);
                }
            }
            for (uint32_t i = 0; i < nq; i++) {
                uint32_t q = pos_at[L][i];
                cpoint ab = absolute_point(L, q);
                cpoint dp = cp_dbl(ab);
                if (q & 1) dp = cp_neg(dp);
                o->domain_points[g][i] = dp;
                /* queried values of this row at this log size: trees 0..3 in order */
                uint32_t row[160], nrow = 0;
                for (int t = 0; t < 4; t++) {
                    static const uint32_t split0[4] = { 10, 12, 8, 0 }, split1[4] = { 40, 48, 8, 0 };
                    if (t == 3) { if (L == max_first) { memcpy(row + nrow, path_cols[3][i], 32); nrow += 8; } continue; }
                    /* path_cols layout: descending log size: larger component first */
                    uint32_t first_is_plonk = log_plonk >= log_pos;
                    uint32_t off_plonk, off_pos;
                    if (log_plonk == log_pos) { off_plonk = 0; off_pos = split0[t]; }
                    else if (first_is_plonk) { off_plonk = 0; off_pos = split0[t]; }
                    else { off_pos = 0; off_plonk = split1[t]; }
                    if (log_plonk == L) { memcpy(row + nrow, path_cols[t][i] + off_plonk, split0[t] * 4); nrow += split0[t]; }
                    if (log_pos == L) { memcpy(row + nrow, path_cols[t][i] + off_pos, split1[t] * 4); nrow += split1[t]; }
                }
                qm31 acc = qm31_from_m31(0);
                for (uint32_t b = 0; b < nb; b++) {
                    cm31 prx = qm31_lo(batches[b].point.x), pix = qm31_hi(batches[b].point.x);
                    cm31 pry = qm31_lo(batches[b].point.y), piy = qm31_hi(batches[b].point.y);
                    cm31 a = cm31_mul(cm31_sub(prx, cm31_mk(dp.x, 0)), piy);
                    cm31 bq = cm31_mul(cm31_sub(pry, cm31_mk(dp.y, 0)), pix);
                    cm31 den = cm31_sub(a, bq);
                    if (den.a == 0 && den.b == 0) FAIL(ORC_STAGE_FRI_FIRST, 1);
                    cm31 dinv = cm31_inv(den);
                    qm31 num = qm31_from_m31(0);
                    for (uint32_t k = 0; k < batches[b].n; k++) {
                        qm31 value = qm31_mul_m31(cc[b][k], row[batches[b].col[k]]);
                        qm31 lin = qm31_add(qm31_mul_m31(ca[b][k], dp.y), cb[b][k]);
                        num = qm31_add(num, qm31_sub(value, lin));
                    }
                    acc = qm31_add(acc, qm31_mul_cm31(num, dinv));
                }
                o->fri_answers[g][i] = acc;
            }
        }
    }

    if (fri_stage(p, o, pos_at, max_first, nq)) return 0;
    o->verdict = 0; o->stage = ORC_OK;
    return 0;
#undef FAIL
}

/* same run, additionally exporting the per-query decommitment hints the verifier circuit takes as witnesses
 * (components/hints/src/decommit.rs:10-16 SinglePathMerkleProof, folding.rs:21-28 SinglePairMerkleProof) */
/* The verifier as the reference calls it: under the CALLER's PcsConfig (FiatShamirHints::new(&proof, config, ..),
 * components/hints/src/fiat_shamir.rs:69-73).  cfg = {pow_bits, log_blowup, log_last, n_queries}.  stwo reads the FRI / PoW
 * parameters from `config` and the proof carries its own copy; a proof whose copy differs cannot verify (wrong query count,
 * wrong last-layer size, ...) -- restated here as a rejection at the parse stage. */
int orc_verify_proof_cfg(const uint8_t *blob, size_t len, const uint32_t *cfg, const uint32_t *input_idx, const uint32_t *input_vals,
                         uint32_t n_inputs, orc_verify_out *o) {
    static _Thread_local orc_proof P;
    if (orc_proof_parse(blob, len, &P) == 0 &&
        (P.pow_bits != cfg[0] || P.log_blowup != cfg[1] || P.log_last != cfg[2] || P.n_queries != cfg[3])) {
        memset(o, 0, sizeof *o);
        o->verdict = 1; o->stage = ORC_STAGE_PARSE;
        return 0;
    }
    return orc_verify_proof(blob, len, input_idx, input_vals, n_inputs, o);
}

/* FRI-only verifier for the synthetic FRI + Merkle instances of BASELINE configs[4] part i: a fresh channel over
 * the FRI commitments (the tail of `transcript` above: components/recursive/fiat_shamir/src/lib.rs:84-130), the opened first-layer values
 * taken from the instance in place of the DEEP quotient answers, then fri_stage.  Blob layout: the 256-word header of
 * recursive-stwo_b200/csrc/synth.cuh (shape at words 1..7, witness counts at 8 / 9 / 10+i / 42+i, nonce at 74, section offsets at 80..191).
 * TEST INFRASTRUCTURE like the rest of oracle/. */
int orc_fri_verify_synth(const uint32_t *w, size_t n_words, orc_verify_out *o) {
    static _Thread_local orc_proof P;
    orc_proof *p = &P;
    memset(p, 0, sizeof *p);
    memset(o, 0, sizeof *o);
    o->verdict = 1;
#define FAIL(stage_, verdict_) do { o->stage = (stage_); o->verdict = (verdict_); return 0; } while (0)
    if (n_words < 256 || w[0] != 0x53594E54u || w[76] > n_words) FAIL(ORC_STAGE_PARSE, 1);
    p->log_size_plonk = w[1]; p->log_size_poseidon = w[2]; p->pow_bits = w[3]; p->log_blowup = w[4]; p->log_last = w[5]; p->n_queries = w[6];
    p->n_inner = w[7];
    const uint32_t nq = p->n_queries, blow = p->log_blowup;
    if (nq == 0 || nq > ORC_MAX_QUERIES || p->n_inner + 1 > ORC_MAX_INNER || p->log_last > 12 || p->pow_bits >= 32 || blow == 0 || blow > 16 ||
        p->log_size_plonk == 0 || p->log_size_plonk > 28 || p->log_size_poseidon == 0 || p->log_size_poseidon > 28)
        FAIL(ORC_STAGE_PARSE, 1);
    const uint32_t max_first = p->log_last + blow + 1 + p->n_inner;
    const uint32_t log_plonk = p->log_size_plonk + blow, log_pos = p->log_size_poseidon + blow;
    if (max_first > 29 || log_plonk > max_first || log_pos > max_first) FAIL(ORC_STAGE_PARSE, 1);
    for (size_t k = 256; k < w[76]; k++) if (w[k] >= 0x7fffffffu) FAIL(ORC_STAGE_PARSE, 1);      /* unused capacity is zero */
    p->first_layer.commitment = w + w[80];
    p->first_layer.fri_witness = w + w[83]; p->first_layer.n_fri_witness = w[8];
    p->first_layer.decommitment.hash_witness = w + w[84]; p->first_layer.decommitment.n_hash_witness = w[9];
    for (uint32_t i = 0; i < p->n_inner; i++) {
        p->inner[i].commitment = w + w[96 + i];
        p->inner[i].fri_witness = w + w[128 + i]; p->inner[i].n_fri_witness = w[10 + i];
        p->inner[i].decommitment.hash_witness = w + w[160 + i]; p->inner[i].decommitment.n_hash_witness = w[42 + i];
    }
    p->last_coeffs = w + w[81]; p->n_last_coeffs = w[85]; p->last_log_size = p->log_last;
    if (p->n_last_coeffs != (1ull << p->log_last)) FAIL(ORC_STAGE_PARSE, 1);
    p->pow_nonce = (uint64_t)w[74] | ((uint64_t)w[75] << 32);
    o->max_first_log = max_first; o->n_inner = p->n_inner; o->n_queries = nq;
    /* channel over the FRI commitments */
    {
        orc_channel ch;
        orc_channel_init(&ch);
        uint32_t d[8];
        orc_channel_mix_root(&ch, p->first_layer.commitment);
        o->fri_alphas[0] = channel_draw_first(&ch);
        for (uint32_t i = 0; i < p->n_inner; i++) {
            orc_channel_mix_root(&ch, p->inner[i].commitment);
            o->fri_alphas[i + 1] = channel_draw_first(&ch);
        }
        for (uint64_t i = 0; i < p->n_last_coeffs; i += 2)
            orc_channel_mix_felts2(&ch, p->last_coeffs + 4 * i, i + 1 < p->n_last_coeffs ? p->last_coeffs + 4 * (i + 1) : NULL);
        uint32_t nf[4] = { (uint32_t)(p->pow_nonce & ((1u << 22) - 1)), (uint32_t)((p->pow_nonce >> 22) & ((1u << 21) - 1)),
                           (uint32_t)((p->pow_nonce >> 43) & ((1u << 21) - 1)), 0 };
        orc_channel_mix_felts2(&ch, nf, NULL);
        memcpy(o->digest_after_nonce, ch.digest, 32);
        const int pow_ok = (ch.digest[0] & ((1u << p->pow_bits) - 1)) == 0;
        uint32_t got = 0;
        for (uint32_t k = 0; k < (nq + 3) / 4; k++) {
            orc_channel_draw(&ch, d);
            for (int j = 0; j < 8 && got < nq; j++) o->raw_queries[got++] = d[j];
        }
        o->n_transcript_perms = (uint32_t)ch.n_perms;
        o->n_perms_paths += o->n_transcript_perms;
        if (!pow_ok) FAIL(ORC_STAGE_POW, 1);
    }
    {
        uint32_t ls[3] = { max_first, log_plonk, log_pos };
        qsort(ls, 3, 4, cmp_u32);
        o->n_logs = 0;
        for (int i = 2; i >= 0; i--) if (o->n_logs == 0 || o->log_sizes[o->n_logs - 1] != ls[i]) o->log_sizes[o->n_logs++] = ls[i];
    }
    static _Thread_local uint32_t pos_at[31][ORC_MAX_QUERIES];
    for (uint32_t L = 1; L <= max_first; L++)
        for (uint32_t i = 0; i < nq; i++) pos_at[L][i] = (o->raw_queries[i] & ((1u << max_first) - 1)) >> (max_first - L);
    for (uint32_t g = 0; g < o->n_logs; g++) memcpy(o->query_pos[g], pos_at[o->log_sizes[g]], nq * 4);
    {
        uint32_t tmp[ORC_MAX_QUERIES];
        memcpy(tmp, pos_at[max_first], nq * 4);
        if (sort_dedup(tmp, nq) != nq) FAIL(ORC_STAGE_UNSUPPORTED, 2);
    }
    for (uint32_t g = 0; g < o->n_logs; g++)
        for (uint32_t i = 0; i < nq; i++) o->fri_answers[g][i] = qm31_load(w + w[82] + (g * nq + i) * 4);
    if (fri_stage(p, o, pos_at, max_first, nq)) return 0;
    o->verdict = 0; o->stage = ORC_OK;
    return 0;
#undef FAIL
}

int orc_verify_proof_hints(const uint8_t *blob, size_t len, const uint32_t *input_idx, const uint32_t *input_vals,
                           uint32_t n_inputs, orc_verify_out *o, orc_hints *h) {
    memset(h, 0, sizeof *h);
    g_hints = h;
    int r = orc_verify_proof(blob, len, input_idx, input_vals, n_inputs, o);
    g_hints = NULL;
    return r;
}

/* the per-query FRI hints of a synthetic instance (the pair_* part of orc_hints; the single_* part stays zero) */
int orc_fri_verify_synth_hints(const uint32_t *w, size_t n_words, orc_verify_out *o, orc_hints *h) {
    memset(h, 0, sizeof *h);
    g_hints = h;
    int r = orc_fri_verify_synth(w, n_words, o);
    g_hints = NULL;
    return r;
}

/* ---- pthread batch driver for the CPU baseline: proofs are independent ----------------------------------- */
#include <pthread.h>
typedef struct {
    const uint8_t *blobs; const uint64_t *off; uint32_t lo, hi;
    const uint32_t *idx, *vals; uint32_t n_inputs; uint8_t *verdict, *stage; uint64_t perms;
} vjob;
static void *vjob_run(void *p) {
    vjob *j = (vjob *)p;
    orc_verify_out *o = malloc(sizeof *o);
    for (uint32_t i = j->lo; i < j->hi; i++) {
        orc_verify_proof(j->blobs + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->idx, j->vals, j->n_inputs, o);
        j->verdict[i] = (uint8_t)o->verdict; j->stage[i] = (uint8_t)o->stage;
        j->perms += o->n_perms_hints + o->n_perms_paths;
    }
    free(o);
    return NULL;
}
/* blobs: back to back, off: n+1 BYTE offsets (each 4-byte aligned); returns the permutations executed */
uint64_t orc_verify_batch_mt(const uint8_t *blobs, const uint64_t *off, uint32_t n, const uint32_t *idx, const uint32_t *vals,
                             uint32_t n_inputs, uint8_t *verdict, uint8_t *stage, unsigned n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > n) n_threads = n ? n : 1;
    pthread_t th[256]; vjob jobs[256];
    for (unsigned t = 0; t < n_threads; t++) {
        vjob j = { blobs, off, (uint32_t)((uint64_t)n * t / n_threads), (uint32_t)((uint64_t)n * (t + 1) / n_threads), idx, vals, n_inputs, verdict, stage, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, vjob_run, &jobs[t]);
    }
    uint64_t perms = 0;
    for (unsigned t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); perms += jobs[t].perms; }
    return perms;
}
