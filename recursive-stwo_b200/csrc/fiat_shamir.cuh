// Per-proof sequential stages of the verifier: Fiat-Shamir transcript, logup total-sum check and
// the OODS composition check.  One thread owns one proof.
//
// Value side of
//   primitives/channel/src/lib.rs:23-58                       (ChannelVar)
//   components/recursive/fiat_shamir/src/lib.rs:31-141        (FiatShamirResults::compute)
//   primitives/circle/src/lib.rs:204-233                      (OODS point from t, x-only doubling)
//   components/recursive/composition/src/{lib,plonk,poseidon,data_structures}.rs (CompositionCheck)
#pragma once
#include "poseidon2.cuh"
#include "proof.cuh"

namespace fs {

#if defined(__CUDACC__)
#define NOINLINE_D static __host__ __device__ __noinline__
#else
#define NOINLINE_D static
#endif

// out-of-line copies keep the straight-line stages small: the permutation is ~1.2k instructions
NOINLINE_D void permute_mem(u32 *st) {
    u32 s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = st[i];
    poseidon2::permute<false>(s);
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = s[i];
}
NOINLINE_D qm31_t qmul(qm31_t a, qm31_t b) { return qm31::mul(a, b); }
NOINLINE_D qm31_t qinv(qm31_t a) { return qm31::inv(a); }
NOINLINE_D u32 minv(u32 a) { return m31::inv(a); }

HD qm31_t qload(const u32 *w) { return qm31::mk(w[0], w[1], w[2], w[3]); }
HD void qstore(u32 *w, qm31_t v) { w[0] = v.v[0]; w[1] = v.v[1]; w[2] = v.v[2]; w[3] = v.v[3]; }
HD qm31_t qadd(qm31_t a, qm31_t b) { return qm31::add(a, b); }
HD qm31_t qsub(qm31_t a, qm31_t b) { return qm31::sub(a, b); }
// multiplication by i, u, iu (primitives/fields/src/qm31.rs:402-418)
HD qm31_t shift_i(qm31_t x) { return qm31::mk(m31::negc(x.v[1]), x.v[0], m31::negc(x.v[3]), x.v[2]); }
HD qm31_t shift_u(qm31_t x) {   // (a + bu) u = (2+i) b + a u
    u32 b0 = x.v[2], b1 = x.v[3];
    return qm31::mk(m31::subc(m31::addc(b0, b0), b1), m31::addc(m31::addc(b1, b1), b0), x.v[0], x.v[1]);
}
HD qm31_t combine_ef(qm31_t a, qm31_t b, qm31_t c, qm31_t d) {
    return qadd(qadd(a, shift_i(b)), qadd(shift_u(c), shift_u(shift_i(d))));
}

// ---- channel ---------------------------------------------------------------------------------------
struct Channel {
    u32 st[16];        // st[8..16] = digest between operations
    u32 n_sent;
    u32 n_perms;
    u32 *sink;         // optional: the 16-word output state of permutation k goes to sink[16 k ..]
    size_t in_delta;   // != 0: its input state goes in_delta words further (the parallel input record)
    HDM void init(u32 *sink_ = nullptr, size_t in_delta_ = 0) { for (int i = 0; i < 16; i++) st[i] = 0; n_sent = 0; n_perms = 0; sink = sink_; in_delta = in_delta_; }
    HDM void emit(const u32 *t) { if (sink) for (int i = 0; i < 16; i++) sink[16 * (size_t)n_perms + i] = t[i]; }
    HDM void emit_in(const u32 *t) { if (sink && in_delta) for (int i = 0; i < 16; i++) sink[in_delta + 16 * (size_t)n_perms + i] = t[i]; }
    HDM void mix8(const u32 *w8) {            // mix_root / mix_two_felts: digest <- capacity(perm(w8 || digest))
        for (int i = 0; i < 8; i++) st[i] = w8[i];
        emit_in(st);
        permute_mem(st);
        emit(st);
        n_sent = 0; n_perms++;
    }
    HDM void mix4(const u32 *w4) {            // mix_one_felt
        u32 t[8] = {w4[0], w4[1], w4[2], w4[3], 0, 0, 0, 0};
        mix8(t);
    }
    HDM void mix44(const u32 *a, const u32 *b) {
        u32 t[8] = {a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]};
        mix8(t);
    }
    HDM void draw(u32 out[8]) {               // rate(perm([n_sent,0..] || digest)); digest unchanged
        u32 t[16];
        t[0] = n_sent++;
        for (int i = 1; i < 8; i++) t[i] = 0;
        for (int i = 8; i < 16; i++) t[i] = st[i];
        emit_in(t);
        permute_mem(t);
        emit(t);
        for (int i = 0; i < 8; i++) out[i] = t[i];
        n_perms++;
    }
};

struct Out {                                   // Fiat-Shamir results of one proof
    qm31_t z, alpha, random_coeff, oods_t, oods_x, oods_y, after_coeff;
    qm31_t fri_alphas[proof::MAX_INNER + 1];
    u32 digest_after_nonce[8];
    u32 raw_queries[proof::MAX_QUERIES];
    u32 n_transcript_perms;
    u32 pow_ok;
};

HD void transcript(const u32 *w, const proof::Desc &d, Out &o, u32 *sink = nullptr, size_t in_delta = 0) {
    Channel ch;
    ch.init(sink, in_delta);
    ch.mix8(w + d.commitments[0]);
    u32 f[4] = {d.log_size_plonk, 0, 0, 0};
    ch.mix4(f);
    f[0] = d.log_size_poseidon;
    ch.mix4(f);
    ch.mix8(w + d.commitments[1]);
    u32 dr[8];
    ch.draw(dr);
    o.z = qload(dr); o.alpha = qload(dr + 4);
    ch.mix8(w + d.stmt1);                       // plonk_total_sum || poseidon_total_sum are adjacent in the blob
    ch.mix8(w + d.commitments[2]);
    ch.draw(dr); o.random_coeff = qload(dr);
    ch.mix8(w + d.commitments[3]);
    ch.draw(dr); o.oods_t = qload(dr);
    {
        qm31_t t2 = qmul(o.oods_t, o.oods_t);
        qm31_t inv = qinv(qm31::add_m31(t2, 1));
        o.oods_x = qmul(qsub(qm31::one(), t2), inv);
        o.oods_y = qmul(qadd(o.oods_t, o.oods_t), inv);
    }
    // sampled values in tree -> column -> mask order, two per permutation
    const u32 *pend = nullptr;
    for (u32 t = 0; t < 4; t++)
        for (u32 c = 0; c < proof::n_cols(t); c++)
            for (u32 m = 0; m < proof::n_masks(t, c); m++) {
                const u32 *v = w + proof::sample_off(d, t, c, m);
                if (pend) { ch.mix44(pend, v); pend = nullptr; } else pend = v;
            }
    if (pend) ch.mix4(pend);
    ch.draw(dr); o.after_coeff = qload(dr);
    ch.mix8(w + d.fl_commitment);
    ch.draw(dr); o.fri_alphas[0] = qload(dr);
    for (u32 i = 0; i < d.n_inner; i++) {
        ch.mix8(w + d.in_commitment[i]);
        ch.draw(dr); o.fri_alphas[i + 1] = qload(dr);
    }
    for (u32 i = 0; i < d.n_last_coeffs; i += 2) {
        if (i + 1 < d.n_last_coeffs) ch.mix8(w + d.last_coeffs + 4 * i);     // two adjacent QM31
        else ch.mix4(w + d.last_coeffs + 4 * i);
    }
    // 64-bit nonce as 22/21/21-bit limbs (components/recursive/data_structures/src/lib.rs:197-213)
    u64 nonce = (u64)w[d.pow_nonce] | ((u64)w[d.pow_nonce + 1] << 32);
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

// sum_inputs 1/(v + idx*alpha - z) + plonk_total_sum + poseidon_total_sum == 0
HD bool logup_sum_ok(const u32 *w, const proof::Desc &d, const Out &o, const u32 *input_idx, const u32 *input_vals, u32 n_inputs) {
    qm31_t sum = qm31::zero();
    for (u32 i = 0; i < n_inputs; i++) {
        qm31_t t = qsub(qadd(qload(input_vals + 4 * i), qm31::mul_m31(o.alpha, input_idx[i])), o.z);
        if (qm31::is_zero(t)) return false;
        sum = qadd(sum, qinv(t));
    }
    sum = qadd(qadd(sum, qload(w + d.stmt1 + 4)), qload(w + d.stmt1));
    return qm31::is_zero(sum);
}

// ---- OODS ------------------------------------------------------------------------------------------
struct EvalRow {
    const u32 *w; const proof::Desc *d;
    qm31_t acc, random_coeff, denom_inv, z, alpha, alpha2, cumsum_shift;
    u32 col[3], base[3];
    qm31_t num[5], den[5];
    u32 n_fracs;
    HDM qm31_t mask(u32 tree) { u32 c = base[tree] + col[tree]++; return qload(w + proof::sample_off(*d, tree, c, 0)); }
    HDM void constraint(qm31_t v) { acc = qadd(qmul(acc, random_coeff), qmul(v, denom_inv)); }
    HDM void relation2(qm31_t mult, qm31_t v0, qm31_t v1) { num[n_fracs] = mult; den[n_fracs] = qsub(qadd(v0, qmul(alpha, v1)), z); n_fracs++; }
    HDM void relation3(qm31_t mult, qm31_t v0, qm31_t v1, qm31_t v2) {
        num[n_fracs] = mult; den[n_fracs] = qsub(qadd(qadd(v0, qmul(alpha, v1)), qmul(alpha2, v2)), z); n_fracs++;
    }
    HDM qm31_t ext_mask(u32 c0, u32 m) {
        return combine_ef(qload(w + proof::sample_off(*d, 2, c0, m)), qload(w + proof::sample_off(*d, 2, c0 + 1, m)),
                          qload(w + proof::sample_off(*d, 2, c0 + 2, m)), qload(w + proof::sample_off(*d, 2, c0 + 3, m)));
    }
    HDM void finalize_logup(u32 batch) {
        u32 nb = (n_fracs + batch - 1) / batch;
        qm31_t prev_col = qm31::zero();
        for (u32 b = 0; b < nb; b++) {
            u32 lo = b * batch, hi = lo + batch < n_fracs ? lo + batch : n_fracs;
            qm31_t p = num[lo], q = den[lo];
            for (u32 k = lo + 1; k < hi; k++) { p = qadd(qmul(p, den[k]), qmul(num[k], q)); q = qmul(q, den[k]); }
            u32 c0 = base[2] + col[2];
            col[2] += 4;
            if (b + 1 < nb) {
                qm31_t cur = ext_mask(c0, 0);
                qm31_t diff = qsub(cur, prev_col);
                prev_col = cur;
                constraint(qsub(qmul(diff, q), p));
            } else {
                qm31_t prev_row = ext_mask(c0, 0), cur = ext_mask(c0, 1);
                qm31_t fixed = qadd(qsub(qsub(cur, prev_row), prev_col), cumsum_shift);
                constraint(qsub(qmul(fixed, q), p));
            }
        }
    }
};

HD qm31_t double_x(qm31_t x) { qm31_t sq = qmul(x, x); return qm31::sub_m31(qadd(sq, sq), 1); }
HD qm31_t pow5(qm31_t x) { qm31_t x2 = qmul(x, x); return qmul(qmul(x2, x2), x); }
HD void ext_matrix(qm31_t *s) {
    for (int b = 0; b < 4; b++) {
        qm31_t *x = s + 4 * b;
        qm31_t t0 = qadd(x[0], x[1]), t02 = qadd(t0, t0), t1 = qadd(x[2], x[3]), t12 = qadd(t1, t1);
        qm31_t t2 = qadd(qadd(x[1], x[1]), t1), t3 = qadd(qadd(x[3], x[3]), t0);
        qm31_t t4 = qadd(qadd(t12, t12), t3), t5 = qadd(qadd(t02, t02), t2);
        x[0] = qadd(t3, t5); x[1] = t5; x[2] = qadd(t2, t4); x[3] = t4;
    }
    for (int j = 0; j < 4; j++) {
        qm31_t t = qadd(qadd(s[j], s[j + 4]), qadd(s[j + 8], s[j + 12]));
        for (int b = 0; b < 4; b++) s[4 * b + j] = qadd(s[4 * b + j], t);
    }
}
HD void int_matrix(qm31_t *s) {
    qm31_t sum = s[0];
    for (int i = 1; i < 16; i++) sum = qadd(sum, s[i]);
    s[0] = qadd(s[0], qadd(qadd(s[0], s[0]), sum));
    for (int i = 1; i < 16; i++) s[i] = qadd(qm31::mul_m31(s[i], 1u << (i + 1)), sum);
}

HD void eval_plonk(EvalRow &e) {
    qm31_t a_wire = e.mask(0), b_wire = e.mask(0), c_wire = e.mask(0), op = e.mask(0);
    qm31_t mult_a = e.mask(0), mult_b = e.mask(0), mult_c = e.mask(0), poseidon_wire = e.mask(0), mult_poseidon = e.mask(0);
    qm31_t enforce_c_m31 = e.mask(0);
    qm31_t v[12];
    for (int i = 0; i < 12; i++) v[i] = e.mask(1);
    e.constraint(qmul(enforce_c_m31, v[9]));
    e.constraint(qmul(enforce_c_m31, v[10]));
    e.constraint(qmul(enforce_c_m31, v[11]));
    qm31_t a = combine_ef(v[0], v[1], v[2], v[3]), b = combine_ef(v[4], v[5], v[6], v[7]), c = combine_ef(v[8], v[9], v[10], v[11]);
    e.constraint(qsub(qsub(c, qmul(op, qadd(a, b))), qmul(qmul(qsub(qm31::one(), op), a), b)));
    e.relation2(mult_a, a, a_wire);
    e.relation2(mult_b, b, b_wire);
    e.relation2(mult_c, c, c_wire);
    e.relation3(qm31::neg(mult_poseidon), poseidon_wire, a, b);
    e.finalize_logup(2);
}

// The 52 preprocessed and 48 trace masks of the Poseidon component are read where they are used (they sit in the proof blob, a few
// cache lines): held in registers they are 96 QM31 = 384 words, which spilled; live now are the 16-element state and the accumulators.
HD void eval_poseidon(EvalRow &e) {
    const qm31_t one = qm31::one();
    const u32 *w = e.w;
    const proof::Desc &d = *e.d;
    const u32 b0 = e.base[0] + e.col[0], b1 = e.base[1] + e.col[1];
    auto pre = [&](u32 c) { return qload(w + proof::sample_off(d, 0, b0 + c, 0)); };      // is_first, is_last, is_full, round_id, rc0[16], rc1[16], ext..
    auto trc = [&](u32 c) { return qload(w + proof::sample_off(d, 1, b1 + c, 0)); };      // in[16], mid[16], out[16]
    auto rc0 = [&](u32 i) { return pre(4 + i); };
    auto rc1 = [&](u32 i) { return pre(20 + i); };
    auto in = [&](u32 i) { return trc(i); };
    auto mid = [&](u32 i) { return trc(16 + i); };
    auto out = [&](u32 i) { return trc(32 + i); };
    e.col[0] += 40; e.col[1] += 48;
    const qm31_t is_first = pre(0), is_last = pre(1), is_full = pre(2);
    const qm31_t not_first = qsub(one, is_first), not_last = qsub(one, is_last), is_partial = qsub(not_first, is_full);
    qm31_t s[16];
    const qm31_t swap = mid(0), nswap = qsub(one, swap);
    for (int i = 0; i < 16; i++)
        s[i] = i < 8 ? qadd(qmul(in(i), nswap), qmul(in(i + 8), swap)) : qadd(qmul(in(i - 8), swap), qmul(in(i), nswap));
    ext_matrix(s);
    for (int i = 0; i < 16; i++) e.constraint(qmul(is_first, qsub(s[i], out(i))));
    for (int i = 0; i < 16; i++) s[i] = pow5(qadd(in(i), rc0(i)));
    for (int i = 0; i < 16; i++) { const qm31_t m = mid(i); e.constraint(qmul(is_full, qsub(m, s[i]))); s[i] = m; }
    ext_matrix(s);
    for (int i = 0; i < 16; i++) s[i] = pow5(qadd(s[i], rc1(i)));
    ext_matrix(s);
    for (int i = 0; i < 16; i++) e.constraint(qmul(is_full, qsub(out(i), s[i])));
    for (int i = 0; i < 16; i++) s[i] = in(i);
    for (int r = 0; r < 14; r++) {
        s[0] = pow5(qadd(s[0], rc0(r)));
        const qm31_t m = mid(r);
        e.constraint(qmul(is_partial, qsub(m, s[0])));
        s[0] = m;
        int_matrix(s);
    }
    for (int i = 0; i < 16; i++) e.constraint(qmul(is_partial, qsub(out(i), s[i])));
    const qm31_t round_id = pre(3), ext1 = pre(36), ext2 = pre(37), ext1_nz = pre(38), ext2_nz = pre(39);
    const qm31_t in_left = qadd(round_id, round_id), in_right = qadd(in_left, one), out_left = qadd(in_right, one), out_right = qadd(out_left, one);
    e.relation3(qsub(qmul(ext1_nz, is_first), not_first), qadd(qmul(is_first, ext1), qmul(not_first, in_left)),
                combine_ef(in(0), in(1), in(2), in(3)), combine_ef(in(4), in(5), in(6), in(7)));
    e.relation3(qsub(qmul(ext2_nz, is_first), not_first), qadd(qmul(is_first, ext2), qmul(not_first, in_right)),
                combine_ef(in(8), in(9), in(10), in(11)), combine_ef(in(12), in(13), in(14), in(15)));
    e.relation3(qadd(qmul(ext1_nz, is_last), not_last), qadd(qmul(is_last, ext1), qmul(not_last, out_left)),
                combine_ef(out(0), out(1), out(2), out(3)), combine_ef(out(4), out(5), out(6), out(7)));
    e.relation3(qadd(qmul(ext2_nz, is_last), not_last), qadd(qmul(is_last, ext2), qmul(not_last, out_right)),
                combine_ef(out(8), out(9), out(10), out(11)), combine_ef(out(12), out(13), out(14), out(15)));
    e.relation2(qmul(is_first, not_last), swap, rc0(0));
    e.finalize_logup(3);
}

// returns true when the composition evaluated from the sampled values matches the committed one
HD bool oods_ok(const u32 *w, const proof::Desc &d, const Out &o, qm31_t *computed, qm31_t *expected) {
    EvalRow e;
    e.w = w; e.d = &d; e.random_coeff = o.random_coeff; e.z = o.z; e.alpha = o.alpha; e.alpha2 = qmul(o.alpha, o.alpha);
    e.acc = qm31::zero();
    for (int i = 0; i < 3; i++) { e.col[i] = 0; e.base[i] = 0; }
    e.n_fracs = 0;
    qm31_t x = o.oods_x;
    for (u32 i = 1; i < d.log_size_plonk; i++) x = double_x(x);        // coset_vanishing of a canonic coset
    e.denom_inv = qinv(x);
    e.cumsum_shift = qm31::mul_m31(qload(w + d.stmt1), minv(1u << d.log_size_plonk));
    eval_plonk(e);
    e.base[0] = 10; e.base[1] = 12; e.base[2] = 8; e.col[0] = e.col[1] = e.col[2] = 0; e.n_fracs = 0;
    x = o.oods_x;
    for (u32 i = 1; i < d.log_size_poseidon; i++) x = double_x(x);
    e.denom_inv = qinv(x);
    e.cumsum_shift = qm31::mul_m31(qload(w + d.stmt1 + 4), minv(1u << d.log_size_poseidon));
    eval_poseidon(e);
    qm31_t left = combine_ef(qload(w + proof::sample_off(d, 3, 0, 0)), qload(w + proof::sample_off(d, 3, 1, 0)),
                             qload(w + proof::sample_off(d, 3, 2, 0)), qload(w + proof::sample_off(d, 3, 3, 0)));
    qm31_t right = combine_ef(qload(w + proof::sample_off(d, 3, 4, 0)), qload(w + proof::sample_off(d, 3, 5, 0)),
                              qload(w + proof::sample_off(d, 3, 6, 0)), qload(w + proof::sample_off(d, 3, 7, 0)));
    u32 bound = d.max_first - d.log_blowup + 1;                          // composition_log_degree_bound
    x = o.oods_x;
    for (u32 i = 0; i + 2 < bound; i++) x = double_x(x);
    *computed = e.acc;
    *expected = qadd(left, qmul(right, x));
    return qm31::eq(*computed, *expected);
}

}  // namespace fs
