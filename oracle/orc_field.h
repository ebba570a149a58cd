/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's field arithmetic, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker for the CUDA path.  Nothing under recursive-stwo_b200/ may
 * include, link or call this.
 *
 * M31  = integers mod p = 2^31-1               (stwo core::fields::m31, used at
 *        reference primitives/fields/src/m31.rs:8-180)
 * CM31 = M31[i]/(i^2+1)                        (primitives/fields/src/cm31.rs:11-279)
 * QM31 = CM31[u]/(u^2-(2+i)), stored (a0+a1 i)+(a2+a3 i)u as 4 u32
 *        (primitives/fields/src/qm31.rs:12-469; layout = QM31::to_m31_array as
 *        used at constraint_system/src/plonk_with_poseidon.rs:473-478)
 * The arithmetic itself lives in the absent stwo dependency (branch
 * cp-poseidon-flattened, no pinned rev); it is pinned here by the fixture
 * known answers (OODS equality, FRI last-layer equality).
 */
#ifndef ORC_FIELD_H
#define ORC_FIELD_H
#include <stdint.h>

#define ORC_P 0x7fffffffu

typedef uint32_t m31;
typedef struct { m31 a, b; } cm31;            /* a + b i */
typedef struct { m31 v[4]; } qm31;            /* (v0 + v1 i) + (v2 + v3 i) u */

static inline m31 m31_red64(uint64_t x) {     /* x < 2^62 */
    uint64_t t = (x & ORC_P) + (x >> 31);
    t = (t & ORC_P) + (t >> 31);
    return (m31)(t == ORC_P ? 0 : t);
}
static inline m31 m31_add(m31 a, m31 b) { uint32_t s = a + b; return s >= ORC_P ? s - ORC_P : s; }
static inline m31 m31_sub(m31 a, m31 b) { return a >= b ? a - b : a + ORC_P - b; }
static inline m31 m31_neg(m31 a) { return a ? ORC_P - a : 0; }
static inline m31 m31_mul(m31 a, m31 b) { return m31_red64((uint64_t)a * b); }
static inline m31 m31_dbl(m31 a) { return m31_add(a, a); }
static inline m31 m31_pow(m31 a, uint64_t e) {
    m31 r = 1;
    while (e) { if (e & 1) r = m31_mul(r, a); a = m31_mul(a, a); e >>= 1; }
    return r;
}
static inline m31 m31_inv(m31 a) { return m31_pow(a, ORC_P - 2); }

static inline cm31 cm31_mk(m31 a, m31 b) { cm31 r = {a, b}; return r; }
static inline cm31 cm31_add(cm31 x, cm31 y) { return cm31_mk(m31_add(x.a, y.a), m31_add(x.b, y.b)); }
static inline cm31 cm31_sub(cm31 x, cm31 y) { return cm31_mk(m31_sub(x.a, y.a), m31_sub(x.b, y.b)); }
static inline cm31 cm31_neg(cm31 x) { return cm31_mk(m31_neg(x.a), m31_neg(x.b)); }
static inline cm31 cm31_mul(cm31 x, cm31 y) {
    return cm31_mk(m31_sub(m31_mul(x.a, y.a), m31_mul(x.b, y.b)),
                   m31_add(m31_mul(x.a, y.b), m31_mul(x.b, y.a)));
}
static inline cm31 cm31_mul_m31(cm31 x, m31 k) { return cm31_mk(m31_mul(x.a, k), m31_mul(x.b, k)); }
static inline cm31 cm31_inv(cm31 x) {
    m31 n = m31_inv(m31_add(m31_mul(x.a, x.a), m31_mul(x.b, x.b)));
    return cm31_mk(m31_mul(x.a, n), m31_mul(m31_neg(x.b), n));
}

static inline qm31 qm31_mk(m31 a, m31 b, m31 c, m31 d) { qm31 r = {{a, b, c, d}}; return r; }
static inline qm31 qm31_from_m31(m31 a) { return qm31_mk(a, 0, 0, 0); }
static inline qm31 qm31_from_cm31(cm31 lo, cm31 hi) { return qm31_mk(lo.a, lo.b, hi.a, hi.b); }
static inline cm31 qm31_lo(qm31 x) { return cm31_mk(x.v[0], x.v[1]); }
static inline cm31 qm31_hi(qm31 x) { return cm31_mk(x.v[2], x.v[3]); }
static inline int qm31_eq(qm31 x, qm31 y) {
    return x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2] && x.v[3] == y.v[3];
}
static inline int qm31_is_zero(qm31 x) { return !(x.v[0] | x.v[1] | x.v[2] | x.v[3]); }
static inline qm31 qm31_add(qm31 x, qm31 y) {
    return qm31_mk(m31_add(x.v[0], y.v[0]), m31_add(x.v[1], y.v[1]),
                   m31_add(x.v[2], y.v[2]), m31_add(x.v[3], y.v[3]));
}
static inline qm31 qm31_sub(qm31 x, qm31 y) {
    return qm31_mk(m31_sub(x.v[0], y.v[0]), m31_sub(x.v[1], y.v[1]),
                   m31_sub(x.v[2], y.v[2]), m31_sub(x.v[3], y.v[3]));
}
static inline qm31 qm31_neg(qm31 x) {
    return qm31_mk(m31_neg(x.v[0]), m31_neg(x.v[1]), m31_neg(x.v[2]), m31_neg(x.v[3]));
}
/* (a + b u)(c + d u) = ac + (2+i) bd + (ad + bc) u */
static inline qm31 qm31_mul(qm31 x, qm31 y) {
    cm31 a = qm31_lo(x), b = qm31_hi(x), c = qm31_lo(y), d = qm31_hi(y);
    cm31 bd = cm31_mul(b, d);
    cm31 r = cm31_mk(m31_sub(m31_dbl(bd.a), bd.b), m31_add(m31_dbl(bd.b), bd.a)); /* (2+i)bd */
    return qm31_from_cm31(cm31_add(cm31_mul(a, c), r), cm31_add(cm31_mul(a, d), cm31_mul(b, c)));
}
static inline qm31 qm31_mul_m31(qm31 x, m31 k) {
    return qm31_mk(m31_mul(x.v[0], k), m31_mul(x.v[1], k), m31_mul(x.v[2], k), m31_mul(x.v[3], k));
}
static inline qm31 qm31_mul_cm31(qm31 x, cm31 k) {
    return qm31_from_cm31(cm31_mul(qm31_lo(x), k), cm31_mul(qm31_hi(x), k));
}
static inline qm31 qm31_inv(qm31 x) {
    cm31 a = qm31_lo(x), b = qm31_hi(x);
    cm31 b2 = cm31_mul(b, b);
    cm31 ib2 = cm31_mk(m31_sub(m31_dbl(b2.a), b2.b), m31_add(m31_dbl(b2.b), b2.a)); /* (2+i) b^2 */
    cm31 den = cm31_inv(cm31_sub(cm31_mul(a, a), ib2));
    return qm31_from_cm31(cm31_mul(a, den), cm31_neg(cm31_mul(b, den)));
}
/* multiplication by the basis elements i, u ("j" in the reference), i*u:
 * reference primitives/fields/src/qm31.rs:402-418 */
static inline qm31 qm31_shift_i(qm31 x) { return qm31_mk(m31_neg(x.v[1]), x.v[0], m31_neg(x.v[3]), x.v[2]); }
static inline qm31 qm31_shift_j(qm31 x) { return qm31_mul(x, qm31_mk(0, 0, 1, 0)); }
static inline qm31 qm31_shift_ij(qm31 x) { return qm31_mul(x, qm31_mk(0, 0, 0, 1)); }

/* circle group x^2+y^2=1 over M31; generator of order 2^31 */
typedef struct { m31 x, y; } cpoint;
#define ORC_GEN_X 2u
#define ORC_GEN_Y 1268011823u
static inline cpoint cp_add(cpoint p, cpoint q) {
    cpoint r = { m31_sub(m31_mul(p.x, q.x), m31_mul(p.y, q.y)),
                 m31_add(m31_mul(p.x, q.y), m31_mul(p.y, q.x)) };
    return r;
}
static inline cpoint cp_dbl(cpoint p) { return cp_add(p, p); }
static inline cpoint cp_neg(cpoint p) { cpoint r = { p.x, m31_neg(p.y) }; return r; }
/* k * G, k taken mod 2^31 */
static inline cpoint cp_mul_gen(uint64_t k) {
    cpoint r = {1, 0}, g = {ORC_GEN_X, ORC_GEN_Y};
    for (int i = 0; i < 31; i++) { if ((k >> i) & 1) r = cp_add(r, g); g = cp_dbl(g); }
    return r;
}
/* generator of the subgroup of order 2^k */
static inline cpoint cp_subgroup_gen(uint32_t k) { return cp_mul_gen(1ull << (31 - k)); }
static inline uint32_t orc_bitrev(uint32_t i, uint32_t n) {
    uint32_t r = 0;
    for (uint32_t b = 0; b < n; b++) r |= ((i >> b) & 1u) << (n - 1 - b);
    return r;
}
/* Coset::half_odds(n).at(i): initial gen(n+2), step gen(n)    */
static inline cpoint cp_half_odds_at(uint32_t n, uint64_t i) {
    return cp_mul_gen((1ull << (31 - (n + 2))) + (i << (31 - n)));
}
#endif
