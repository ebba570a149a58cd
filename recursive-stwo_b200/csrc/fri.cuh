// Query positions, circle-domain points, DEEP quotient answers and FRI folds.
//
// Value side of
//   primitives/query/src/lib.rs:19-168                       (positions per log size, point carried with the query)
//   components/recursive/answer/src/lib.rs:34-382            (AnswerResults::compute, fri_answers_for_log_size)
//   components/recursive/answer/src/data_structures.rs:42-189 (sample batches, line coefficients, row quotients)
//   components/hints/src/folding.rs:296-452,459-601          (first-layer evaluation rebuild, inner-layer folds)
//   components/recursive/folding/src/lib.rs:56-204           (circle fold, line folds, last-layer check)
//   primitives/line/src/lib.rs:39-67                         (LinePolyVar::eval_at_point)
#pragma once
#include "decommit.cuh"

namespace fri {

using fs::qadd; using fs::qsub; using fs::qmul; using fs::qload; using fs::qstore; using fs::minv;

constexpr u32 MAX_LOGS = 3;

struct QPoint { qm31_t x, y; };

// log sizes with committed columns, descending, deduplicated: {max_first, log_plonk, log_pos}
HD u32 log_sizes(const proof::Desc &d, u32 out[MAX_LOGS]) {
    u32 v[3] = {d.max_first, d.log_plonk, d.log_pos};
    for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++) if (v[j] > v[i]) { u32 t = v[i]; v[i] = v[j]; v[j] = t; }
    u32 n = 0;
    for (int i = 0; i < 3; i++) if (n == 0 || out[n - 1] != v[i]) out[n++] = v[i];
    return n;
}
HD u32 position(const proof::Desc &d, u32 raw, u32 L) { return (raw & ((1u << d.max_first) - 1u)) >> (d.max_first - L); }

// k * G for the circle generator G of order 2^31, k in [0, 2^31): table-free double-and-add (31 steps).
// Coset::half_odds(L).at(j) = (2^(29-L) + j * 2^(31-L)) * G.
HD cpoint_t half_odds_at(u32 L, u32 j) { return circle::mul_gen((1u << (29 - L)) + (j << (31 - L))); }
// absolute point of position q at log size L (PointCarryingQueryVar::get_absolute_point)
HD cpoint_t absolute_point(u32 L, u32 q) { return half_odds_at(L, circle::bitrev(q >> 1, L - 1)); }
// circle-domain point of position q at log size L (get_next_point): double, negate when the low bit is set
HD cpoint_t domain_point(u32 L, u32 q) {
    cpoint_t p = circle::dbl(absolute_point(L, q));
    return (q & 1u) ? circle::conj(p) : p;
}

// ---- answers ---------------------------------------------------------------------------------------------
// Column c of tree t: component log size (without blow-up) and committed log size.
HD void column_info(const proof::Desc &d, u32 t, u32 c, u32 &comp_log, u32 &col_log) {
    if (t == 3) { comp_log = 0; col_log = d.max_first; }
    else if (c < proof::plonk_cols(t)) { comp_log = d.log_size_plonk; col_log = d.log_plonk; }
    else { comp_log = d.log_size_poseidon; col_log = d.log_pos; }
}

constexpr u32 MAX_BATCHES = 4;
constexpr u32 MAX_GROUP_SAMPLES = 160;
// Sample batches of one log-size group with the line coefficients (alpha*a, alpha*b, alpha*c) of every sample;
// samples are stored batch-major: batch b owns [start[b], start[b+1]).
struct Group {
    u32 n_batches;
    u32 n_cols;                                       // columns (row width) at this log size
    u32 start[MAX_BATCHES + 1];
    QPoint point[MAX_BATCHES];
    u32 col[MAX_GROUP_SAMPLES];                       // index into the row of queried values of this log size
    qm31_t ca[MAX_GROUP_SAMPLES], cb[MAX_GROUP_SAMPLES], cc[MAX_GROUP_SAMPLES];
};

HD bool build_group(const u32 *w, const proof::Desc &d, const fs::Out &o, u32 L, Group &g) {
    int shift_of[MAX_BATCHES]; u32 key_log[MAX_BATCHES]; u32 cnt[MAX_BATCHES], fill[MAX_BATCHES];
    g.n_batches = 0;
    // pass 0 discovers the batches (keyed by mask shift, first-seen order over tree -> column -> mask) and counts
    // their samples; pass 1 places every sample batch-major
    for (int pass = 0; pass < 2; pass++) {
        u32 col_index = 0;
        for (u32 t = 0; t < 4; t++)
            for (u32 c = 0; c < proof::n_cols(t); c++) {
                u32 comp_log, col_log;
                column_info(d, t, c, comp_log, col_log);
                if (col_log != L) continue;
                const u32 nm = proof::n_masks(t, c);
                for (u32 m = 0; m < nm; m++) {
                    const int shift = (nm == 2 && m == 0) ? -1 : 0;
                    const u32 kl = shift ? comp_log : 0;
                    u32 b = 0;
                    while (b < g.n_batches && !(shift_of[b] == shift && key_log[b] == kl)) b++;
                    if (pass == 0) {
                        if (b == g.n_batches) {
                            if (b == MAX_BATCHES) return false;
                            shift_of[b] = shift; key_log[b] = kl; cnt[b] = 0;
                            g.point[b].x = o.oods_x; g.point[b].y = o.oods_y;
                            if (shift) {   // oods + (-1) * step, step = generator of the subgroup of order 2^comp_log
                                cpoint_t s = circle::conj(circle::mul_gen(1u << (31 - comp_log)));
                                g.point[b].x = qsub(qm31::mul_m31(o.oods_x, s.x), qm31::mul_m31(o.oods_y, s.y));
                                g.point[b].y = qadd(qm31::mul_m31(o.oods_x, s.y), qm31::mul_m31(o.oods_y, s.x));
                            }
                            g.n_batches++;
                        }
                        cnt[b]++;
                    } else {
                        const u32 at = g.start[b] + fill[b]++;
                        g.col[at] = col_index;
                        g.ca[at] = qload(w + proof::sample_off(d, t, c, m));    // sampled value; coefficients replace it below
                    }
                }
                col_index++;
            }
        g.n_cols = col_index;
        if (pass == 0) {
            g.start[0] = 0;
            for (u32 b = 0; b < g.n_batches; b++) { g.start[b + 1] = g.start[b] + cnt[b]; fill[b] = 0; }
            if (g.start[g.n_batches] > MAX_GROUP_SAMPLES) return false;
        }
    }
    // complex_conjugate_line_coeffs with the running alpha = -2u * random_coeff^k
    qm31_t alpha = qm31::mk(0, 0, M31_P - 2, 0);
    for (u32 b = 0; b < g.n_batches; b++) {
        const cm31_t y0 = qm31::lo(g.point[b].y), y1 = qm31::hi(g.point[b].y);
        for (u32 k = g.start[b]; k < g.start[b + 1]; k++) {
            const qm31_t v = g.ca[k];
            const cm31_t v0 = qm31::lo(v), v1 = qm31::hi(v);
            const cm31_t bb = cm31::sub(cm31::mul(v0, y1), cm31::mul(v1, y0));
            g.ca[k] = qm31::mul_cm31(alpha, v1);
            g.cb[k] = qm31::mul_cm31(alpha, bb);
            g.cc[k] = qm31::mul_cm31(alpha, y1);
            alpha = qmul(alpha, o.after_coeff);
        }
    }
    return true;
}

// The same with a GROUP of lanes per (proof, log size): the structure passes are cheap integer walks every lane repeats (stores are dealt
// out), the coefficient loop -- a chain of ~140 dependent QM31 products in the sequential form, alpha_{k+1} = alpha_k * after_coeff --
// is strided: lane j owns the samples k = j, j + G, .. and walks alpha in steps of after_coeff^G.  Co: lane(), size(), sync().
template <class Co>
HD bool build_group_coop(const Co &co, const u32 *w, const proof::Desc &d, const fs::Out &o, u32 L, Group &g) {
    const u32 lane = co.lane(), G = co.size();
    int shift_of[MAX_BATCHES]; u32 key_log[MAX_BATCHES]; u32 cnt[MAX_BATCHES], fill[MAX_BATCHES], start[MAX_BATCHES + 1];
    u32 n_batches = 0;
    bool ok = true;
    for (int pass = 0; pass < 2 && ok; pass++) {
        u32 col_index = 0;
        for (u32 t = 0; t < 4 && ok; t++)
            for (u32 c = 0; c < proof::n_cols(t) && ok; c++) {
                u32 comp_log, col_log;
                column_info(d, t, c, comp_log, col_log);
                if (col_log != L) continue;
                const u32 nm = proof::n_masks(t, c);
                for (u32 m = 0; m < nm; m++) {
                    const int shift = (nm == 2 && m == 0) ? -1 : 0;
                    const u32 kl = shift ? comp_log : 0;
                    u32 b = 0;
                    while (b < n_batches && !(shift_of[b] == shift && key_log[b] == kl)) b++;
                    if (pass == 0) {
                        if (b == n_batches) {
                            if (b == MAX_BATCHES) { ok = false; break; }
                            shift_of[b] = shift; key_log[b] = kl; cnt[b] = 0;
                            if (lane == b % G) {
                                g.point[b].x = o.oods_x; g.point[b].y = o.oods_y;
                                if (shift) {
                                    cpoint_t s = circle::conj(circle::mul_gen(1u << (31 - comp_log)));
                                    g.point[b].x = qsub(qm31::mul_m31(o.oods_x, s.x), qm31::mul_m31(o.oods_y, s.y));
                                    g.point[b].y = qadd(qm31::mul_m31(o.oods_x, s.y), qm31::mul_m31(o.oods_y, s.x));
                                }
                            }
                            n_batches++;
                        }
                        cnt[b]++;
                    } else {
                        const u32 at = start[b] + fill[b]++;
                        if (at % G == lane) {
                            g.col[at] = col_index;
                            g.ca[at] = qload(w + proof::sample_off(d, t, c, m));
                        }
                    }
                }
                col_index++;
            }
        if (lane == 0) g.n_cols = col_index;
        if (pass == 0 && ok) {
            start[0] = 0;
            for (u32 b = 0; b < n_batches; b++) { start[b + 1] = start[b] + cnt[b]; fill[b] = 0; }
            if (start[n_batches] > MAX_GROUP_SAMPLES) ok = false;
        }
    }
    if (!ok) return false;                              // every lane took the same decisions
    if (lane == 0) { g.n_batches = n_batches; for (u32 b = 0; b <= n_batches; b++) g.start[b] = start[b]; }
    co.sync();                                          // points and sampled values are in place
    // alpha_k = -2u * after_coeff^k over the group's samples in batch-major order; lane j: k = j, j + G, ..
    qm31_t step = o.after_coeff, alpha = qm31::mk(0, 0, M31_P - 2, 0);
    {
        qm31_t sq = o.after_coeff;                       // after_coeff^(2^i)
        for (u32 bit = 1; bit < G; bit <<= 1) {
            if (lane & bit) alpha = qmul(alpha, sq);
            sq = qmul(sq, sq);
        }
        step = G == 1 ? o.after_coeff : sq;              // after_coeff^G (G a power of two)
    }
    u32 b = 0;
    for (u32 k = lane; k < start[n_batches]; k += G) {
        while (k >= start[b + 1]) b++;
        const cm31_t y0 = qm31::lo(g.point[b].y), y1 = qm31::hi(g.point[b].y);
        const qm31_t v = g.ca[k];
        const cm31_t v0 = qm31::lo(v), v1 = qm31::hi(v);
        const cm31_t bb = cm31::sub(cm31::mul(v0, y1), cm31::mul(v1, y0));
        g.ca[k] = qm31::mul_cm31(alpha, v1);
        g.cb[k] = qm31::mul_cm31(alpha, bb);
        g.cc[k] = qm31::mul_cm31(alpha, y1);
        alpha = qmul(alpha, step);
    }
    co.sync();
    return true;
}

// accumulate_row_quotients for one query: row = queried values of all columns of this log size
HD bool row_quotient(const Group &g, const u32 *row, cpoint_t dp, qm31_t &out) {
    qm31_t acc = qm31::zero();
    for (u32 b = 0; b < g.n_batches; b++) {
        const cm31_t prx = qm31::lo(g.point[b].x), pix = qm31::hi(g.point[b].x);
        const cm31_t pry = qm31::lo(g.point[b].y), piy = qm31::hi(g.point[b].y);
        const cm31_t den = cm31::sub(cm31::mul(cm31::sub(prx, cm31::mk(dp.x, 0)), piy), cm31::mul(cm31::sub(pry, cm31::mk(dp.y, 0)), pix));
        if ((den.a | den.b) == 0) return false;
        const cm31_t dinv = cm31::inv(den);
        qm31_t num = qm31::zero();
        for (u32 k = g.start[b]; k < g.start[b + 1]; k++) {
            const qm31_t value = qm31::mul_m31(g.cc[k], row[g.col[k]]);
            const qm31_t lin = qadd(qm31::mul_m31(g.ca[k], dp.y), g.cb[k]);
            num = qadd(num, qsub(value, lin));
        }
        acc = qadd(acc, qm31::mul_cm31(num, dinv));
    }
    out = acc;
    return true;
}

// Gather the row of queried values at log size L for one query from the four per-query commitment-tree paths.
// paths[t] points at the query's column values of tree t in path order (leaf layer first, then the injected layer).
HD u32 gather_row(const proof::Desc &d, u32 L, const u32 *const paths[4], u32 *row) {
    u32 n = 0;
    for (u32 t = 0; t < 3; t++) {
        const u32 np = proof::plonk_cols(t), ns = proof::n_cols(t) - np;
        // the larger component is the leaf layer; equal sizes share the leaf layer, Plonk columns first
        const u32 off_plonk = d.log_plonk >= d.log_pos ? 0 : ns, off_pos = d.log_plonk >= d.log_pos ? np : 0;
        if (d.log_plonk == L) for (u32 c = 0; c < np; c++) row[n++] = paths[t][off_plonk + c];
        if (d.log_pos == L) for (u32 c = 0; c < ns; c++) row[n++] = paths[t][off_pos + c];
    }
    if (d.max_first == L) for (u32 c = 0; c < 8; c++) row[n++] = paths[3][c];
    return n;
}

// ---- folds ------------------------------------------------------------------------------------------------------
// LinePolyVar::eval_at_point: coefficients folded with the x-doublings, innermost pairs take the last doubling
HD qm31_t eval_last_poly(const u32 *coeffs, u32 log_n, u32 x, qm31_t *buf) {
    if (log_n == 0) return qload(coeffs);
    u32 dbl[16];
    dbl[0] = x;
    for (u32 k = 1; k < log_n; k++) { u32 sq = m31::mulc(dbl[k - 1], dbl[k - 1]); dbl[k] = m31::subc(m31::addc(sq, sq), 1); }
    u32 n = 1u << log_n;
    for (u32 k = 0; k < n; k++) buf[k] = qload(coeffs + 4 * k);
    for (u32 lev = log_n; lev-- > 0;) {
        n >>= 1;
        for (u32 k = 0; k < n; k++) buf[k] = qadd(buf[2 * k], qm31::mul_m31(buf[2 * k + 1], dbl[lev]));
    }
    return buf[0];
}

// fold of one pair: (l + r) + alpha * (l - r) * inv  with (l, r) ordered by the low bit of the position
HD qm31_t fold_pair(qm31_t self, qm31_t sib, u32 pos, u32 inv, qm31_t alpha) {
    const qm31_t l = (pos & 1u) ? sib : self, r = (pos & 1u) ? self : sib;
    return qadd(qadd(l, r), qmul(qm31::mul_m31(qsub(l, r), inv), alpha));
}

}  // namespace fri
