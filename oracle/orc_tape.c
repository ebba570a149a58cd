/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc.h).
 * CPU restatement of the VALUE side of the circuit DSL and of the constraint system's finalisation loops, replaying
 * the value log oracle/orc_dsl.py records while it builds the circuit (one record per variable, in creation order):
 *   primitives/fields/src/{m31,cm31,qm31}.rs          value = ... of add / mul / neg / inv / decompose
 *   primitives/bits/src/lib.rs:48-82                  bit decomposition
 *   primitives/poseidon31/src/lib.rs:282-407          poseidon2_permute on the (swapped) halves, PoseidonEntry hashes
 *   constraint_system/src/plonk_with_poseidon.rs:337-380 check_arithmetics, :468-519 check_poseidon_invocations,
 *                                                :521-628 generate_plonk_with_poseidon_circuit (the 12 value columns + op)
 * It is what bench.py times as the CPU baseline of trace generation (the reference does the same arithmetic while it
 * appends rows); the append bookkeeping itself (Vec pushes, Rc/RefCell) is not replayed, so the baseline is favourable
 * to the CPU. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

enum { LOG_ADD = 1, LOG_MUL, LOG_MULC, LOG_IN, LOG_INV_M31, LOG_INV_QM31, LOG_CINV_RE, LOG_CINV_IM, LOG_COORD, LOG_BIT, LOG_PERM,
       LOG_M4, LOG_POW5M4, LOG_HADAMARD, LOG_GRANDSUM, LOG_POW4 };   /* the last five: constraint_system/src/plonk_without_poseidon.rs:108-245 */
#define NO_VAR 0xffffffffu

/* the M4 MDS block on the four coordinates (plonk_without_poseidon.rs:115-125) */
static qm31 q_m4(qm31 x) {
    m31 t0 = m31_add(x.v[0], x.v[1]), t1 = m31_add(x.v[2], x.v[3]);
    m31 t2 = m31_add(m31_dbl(x.v[1]), t1), t3 = m31_add(m31_dbl(x.v[3]), t0);
    m31 t4 = m31_add(m31_dbl(m31_dbl(t1)), t3), t5 = m31_add(m31_dbl(m31_dbl(t0)), t2);
    return qm31_mk(m31_add(t3, t5), t5, m31_add(t2, t4), t4);
}
static qm31 q_had(qm31 a, qm31 b) { return qm31_mk(m31_mul(a.v[0], b.v[0]), m31_mul(a.v[1], b.v[1]), m31_mul(a.v[2], b.v[2]), m31_mul(a.v[3], b.v[3])); }
static qm31 q_pow4(qm31 a) { qm31 s = q_had(a, a); return q_had(s, s); }
static qm31 q_grandsum(qm31 a, qm31 b) {
    m31 s = m31_add(m31_add(m31_add(a.v[0], a.v[1]), m31_add(a.v[2], a.v[3])), m31_add(m31_add(b.v[0], b.v[1]), m31_add(b.v[2], b.v[3])));
    return qm31_mk(s, s, s, s);
}
static qm31 ldq(const uint32_t *vars, uint32_t i) { return qm31_mk(vars[4 * i], vars[4 * i + 1], vars[4 * i + 2], vars[4 * i + 3]); }
static void stq(uint32_t *vars, uint32_t i, qm31 q) { memcpy(vars + 4 * i, q.v, 16); }

/* ops: n_ops x 8 words {kind, dst, a, b, v0..v3}; perms: n_perms x 32 words {l_kind, l_a, l_b, r_kind, r_a, r_b, swap_var,
 * out[4], pad[5], left literal[8], right literal[8]}.  vars: n_vars x 4 (written); flow_hash: n_perms x 32; flow_swap: n_perms. */
void orc_circuit_replay(const uint32_t *ops, uint32_t n_ops, const uint32_t *perms, uint32_t *vars, uint32_t *flow_hash, uint8_t *flow_swap) {
    stq(vars, 0, qm31_mk(0, 0, 0, 0)); stq(vars, 1, qm31_mk(1, 0, 0, 0)); stq(vars, 2, qm31_mk(0, 1, 0, 0)); stq(vars, 3, qm31_mk(0, 0, 1, 0));
    for (uint32_t k = 0; k < n_ops; k++) {
        const uint32_t *o = ops + 8 * k;
        const uint32_t dst = o[1], a = o[2], b = o[3];
        switch (o[0]) {
        case LOG_ADD: stq(vars, dst, qm31_add(ldq(vars, a), ldq(vars, b))); break;
        case LOG_MUL: stq(vars, dst, qm31_mul(ldq(vars, a), ldq(vars, b))); break;
        case LOG_MULC: stq(vars, dst, qm31_mul_m31(ldq(vars, a), b)); break;
        case LOG_IN: memcpy(vars + 4 * dst, o + 4, 16); break;
        case LOG_INV_M31: stq(vars, dst, qm31_from_m31(m31_inv(vars[4 * a]))); break;
        case LOG_INV_QM31: stq(vars, dst, qm31_inv(ldq(vars, a))); break;
        case LOG_CINV_RE: stq(vars, dst, qm31_from_m31(cm31_inv(qm31_lo(ldq(vars, a))).a)); break;
        case LOG_CINV_IM: stq(vars, dst, qm31_from_m31(cm31_inv(qm31_lo(ldq(vars, a))).b)); break;
        case LOG_COORD: stq(vars, dst, qm31_from_m31(vars[4 * a + b])); break;
        case LOG_BIT: stq(vars, dst, qm31_from_m31((vars[4 * a] >> b) & 1u)); break;
        case LOG_M4: stq(vars, dst, q_m4(ldq(vars, a))); break;
        case LOG_POW5M4: stq(vars, dst, q_m4(q_had(ldq(vars, a), ldq(vars, b)))); break;
        case LOG_HADAMARD: stq(vars, dst, q_had(ldq(vars, a), ldq(vars, b))); break;
        case LOG_GRANDSUM: stq(vars, dst, q_grandsum(ldq(vars, a), ldq(vars, b))); break;
        case LOG_POW4: stq(vars, dst, q_pow4(ldq(vars, a))); break;
        case LOG_PERM: {
            const uint32_t *p = perms + 32 * dst;
            uint32_t in[16], st[16];
            for (int h = 0; h < 2; h++) {
                const uint32_t *d = p + 3 * h;
                if (d[0]) memcpy(in + 8 * h, p + 16 + 8 * h, 32);
                else { memcpy(in + 8 * h, vars + 4 * d[1], 16); memcpy(in + 8 * h + 4, vars + 4 * d[2], 16); }
            }
            const int swap = p[6] != NO_VAR && vars[4 * p[6]] != 0;
            memcpy(st, in + (swap ? 8 : 0), 32); memcpy(st + 8, in + (swap ? 0 : 8), 32);
            orc_poseidon2_permute(st);
            memcpy(flow_hash + 32 * dst, in, 64); memcpy(flow_hash + 32 * dst + 16, st, 64);
            flow_swap[dst] = (uint8_t)swap;
            for (int q = 0; q < 4; q++) if (p[7 + q] != NO_VAR) memcpy(vars + 4 * p[7 + q], st + 4 * q, 16);
            break;
        }
        default: break;
        }
    }
}

/* wiring: 6 x n_rows words (a_wire, b_wire, c_wire, poseidon_wire, enforce_c_m31, op).  Returns the first bad row or -1. */
int64_t orc_circuit_check_arithmetics(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars) {
    const uint32_t *aw = wiring, *bw = wiring + n_rows, *cw = wiring + 2 * (size_t)n_rows, *enf = wiring + 4 * (size_t)n_rows, *op = wiring + 5 * (size_t)n_rows;
    for (uint32_t i = 0; i < n_rows; i++) {
        const qm31 a = ldq(vars, aw[i]), b = ldq(vars, bw[i]), c = ldq(vars, cw[i]);
        const qm31 want = qm31_add(qm31_mul_m31(qm31_add(a, b), op[i]), qm31_mul_m31(qm31_mul(a, b), m31_sub(1, op[i])));
        if (!qm31_eq(want, c)) return i;
        if (enf[i] && (c.v[1] | c.v[2] | c.v[3])) return i;
    }
    return -1;
}

/* Plonk-without-Poseidon system (plonk_without_poseidon.rs:410-599): wiring = 7 x n_rows words (a_wire, b_wire, c_wire, op1..op4) */
int64_t orc_circuit_check_arithmetics_without(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars) {
    const uint32_t *aw = wiring, *bw = wiring + n_rows, *cw = wiring + 2 * (size_t)n_rows, *op1 = wiring + 3 * (size_t)n_rows,
                   *op2 = wiring + 4 * (size_t)n_rows, *op3 = wiring + 5 * (size_t)n_rows, *op4 = wiring + 6 * (size_t)n_rows;
    for (uint32_t i = 0; i < n_rows; i++) {
        const qm31 a = ldq(vars, aw[i]), b = ldq(vars, bw[i]), c = ldq(vars, cw[i]);
        const uint32_t sel = op2[i] * 4 + op3[i] * 2 + op4[i];
        qm31 want;
        if (op2[i] > 1 || op3[i] > 1 || op4[i] > 1) return i;
        if (sel && op1[i] != 1) return i;
        switch (sel) {
        case 0: want = qm31_add(qm31_mul_m31(qm31_add(a, b), op1[i]), qm31_mul_m31(qm31_mul(a, b), m31_sub(1, op1[i]))); break;
        case 1: want = q_had(a, b); break;                              /* hadamard */
        case 6: want = q_m4(q_had(a, b)); if (!qm31_eq(b, q_pow4(a))) return i; break;   /* pow5m4 */
        case 5: want = q_had(a, b); if (!qm31_eq(b, q_pow4(a))) return i; break;         /* pow5 */
        case 2: want = q_m4(q_had(a, b)); break;                        /* m4 (b = (1,1,1,1)) */
        case 3: want = q_grandsum(a, b); break;
        default: return i;
        }
        if (!qm31_eq(want, c)) return i;
    }
    return -1;
}

/* flow_wire: n_flow x 4.  row_of_wire: scratch of n_vars words.  Returns the first bad flow entry or -1. */
int64_t orc_circuit_check_poseidon(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars, uint32_t n_vars, const uint32_t *flow_wire,
                                   uint32_t n_flow, const uint32_t *flow_hash, const uint8_t *flow_swap, uint32_t *row_of_wire) {
    const uint32_t *aw = wiring, *bw = wiring + n_rows, *pw = wiring + 3 * (size_t)n_rows;
    memset(row_of_wire, 0xff, (size_t)n_vars * 4);
    for (uint32_t i = 0; i < n_rows; i++) if (pw[i] && row_of_wire[pw[i]] == NO_VAR) row_of_wire[pw[i]] = i;
    for (uint32_t e = 0; e < n_flow; e++) {
        const uint32_t *h = flow_hash + 32 * (size_t)e;
        for (int k = 0; k < 4; k++) {
            const uint32_t w = flow_wire[4 * e + k];
            if (!w) continue;
            const uint32_t row = row_of_wire[w];
            if (row == NO_VAR) return e;
            if (memcmp(vars + 4 * aw[row], h + 8 * k, 16) || memcmp(vars + 4 * bw[row], h + 8 * k + 4, 16)) return e;
        }
        uint32_t st[16];
        memcpy(st, h + (flow_swap[e] ? 8 : 0), 32); memcpy(st + 8, h + (flow_swap[e] ? 0 : 8), 32);
        orc_poseidon2_permute(st);
        if (memcmp(st, h + 16, 64)) return e;
    }
    return -1;
}

/* the 12 value columns (a_val_0..3, b_val_0..3, c_val_0..3), column-major: out[12][n_rows] */
void orc_circuit_export_values(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars, uint32_t *out) {
    for (int w = 0; w < 3; w++) {
        const uint32_t *wire = wiring + (size_t)w * n_rows;
        for (int k = 0; k < 4; k++) {
            uint32_t *col = out + ((size_t)4 * w + k) * n_rows;
            for (uint32_t i = 0; i < n_rows; i++) col[i] = vars[4 * wire[i] + k];
        }
    }
}

/* ---- pthread driver for the CPU baseline: n independent replicas of one proof's circuit ------------------------------ */
typedef struct {
    const uint32_t *ops, *perms, *wiring, *flow_wire;
    uint32_t n_ops, n_perms, n_vars, n_rows, first, count;
    int64_t bad;
} tjob;
static void *tjob_run(void *arg) {
    tjob *j = (tjob *)arg;
    uint32_t *vars = malloc((size_t)j->n_vars * 16), *fh = malloc((size_t)j->n_perms * 128 + 16), *row_of = malloc((size_t)j->n_vars * 4);
    uint32_t *cols = malloc((size_t)12 * j->n_rows * 4);
    uint8_t *fs = malloc(j->n_perms + 1);
    j->bad = 0;
    for (uint32_t r = 0; r < j->count; r++) {
        orc_circuit_replay(j->ops, j->n_ops, j->perms, vars, fh, fs);
        if (orc_circuit_check_arithmetics(j->wiring, j->n_rows, vars) != -1) j->bad++;
        if (orc_circuit_check_poseidon(j->wiring, j->n_rows, vars, j->n_vars, j->flow_wire, j->n_perms, fh, fs, row_of) != -1) j->bad++;
        orc_circuit_export_values(j->wiring, j->n_rows, vars, cols);
    }
    free(vars); free(fh); free(row_of); free(cols); free(fs);
    return NULL;
}
/* replay + check_arithmetics + check_poseidon_invocations + value-column export for n replicas on n_threads pthreads;
 * returns the number of failed checks (0 expected) */
int64_t orc_circuit_trace_mt(const uint32_t *ops, uint32_t n_ops, const uint32_t *perms, uint32_t n_perms, uint32_t n_vars,
                             const uint32_t *wiring, uint32_t n_rows, const uint32_t *flow_wire, uint32_t n, unsigned n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    tjob jobs[256];
    for (unsigned t = 0; t < n_threads; t++) {
        tjob j = { ops, perms, wiring, flow_wire, n_ops, n_perms, n_vars, n_rows, 0, 0, 0 };
        j.first = (uint32_t)((uint64_t)n * t / n_threads);
        j.count = (uint32_t)((uint64_t)n * (t + 1) / n_threads) - j.first;
        jobs[t] = j;
        pthread_create(&th[t], NULL, tjob_run, &jobs[t]);
    }
    int64_t bad = 0;
    for (unsigned t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
    return bad;
}
