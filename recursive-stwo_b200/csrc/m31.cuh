// M31 / CM31 / QM31 arithmetic for sm_100a kernels.
//
// Replaces the value side of the reference's field types
//   primitives/fields/src/m31.rs:8-180, cm31.rs:11-279, qm31.rs:12-469
// (whose arithmetic lives in stwo core::fields).  p = 2^31-1, CM31 = M31[i]/(i^2+1),
// QM31 = CM31[u]/(u^2-(2+i)), QM31 stored as 4 words (a0 + a1 i) + (a2 + a3 i) u.
//
// Range classes used for lazy reduction (all values are u32):
//   C  : canonical, x <  p
//   C0 : x <= p          (p aliases 0)
//   C1 : x <= 2^32 - 3   (= 2p - 1)
// C0 + C0 is C1 and never wraps; fold(C1) is C0; a product of two C0 values
// reduced once is C1.  Every function states the class it needs and returns.
//
// HD functions compile for host too so tests/ can exercise exactly this code on
// the CPU box (tests/hostsim); the product only ever calls them from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDM __host__ __device__ __forceinline__      // member functions
#else
#define HD static inline
#define HDM inline
#endif

#define M31_P 0x7fffffffu

typedef uint32_t u32;
typedef uint64_t u64;

namespace m31 {

// C1 -> C0
HD u32 fold(u32 x) { return (x & M31_P) + (x >> 31); }
// C1 -> C (x <= 2p-1): one subtract + unsigned min
HD u32 canon(u32 x) {
    u32 y = x - M31_P;
    return y < x ? y : x;
}
// C0,C0 -> C1 (no reduction)
HD u32 add_lazy(u32 a, u32 b) { return a + b; }
// C0,C0 -> C0
HD u32 add(u32 a, u32 b) { return fold(a + b); }
// C,C -> C
HD u32 addc(u32 a, u32 b) { return canon(a + b); }
// C,C -> C
HD u32 subc(u32 a, u32 b) {
    u32 d = a - b;
    return a < b ? d + M31_P : d;
}
// C0 -> C0 : p - a
HD u32 neg0(u32 a) { return M31_P - a; }
HD u32 negc(u32 a) { return a ? M31_P - a : 0u; }
// C0,C0 -> C1.  a*b = H*2^31 + L, H <= 2^31-2, L <= p  =>  H + L <= 2^32-3.
// Written as a * (2b) so H is the high word and L is lo>>1 (IMAD.WIDE + LEA.HI).
HD u32 mul_lazy(u32 a, u32 b) {
    u64 x = (u64)a * (u64)(b << 1);
    return (u32)(x >> 32) + ((u32)x >> 1);
}
// same, with the doubled operand supplied by the caller (b2 = 2*b, b in C0)
HD u32 mul_lazy_pre2(u32 a, u32 b2) {
    u64 x = (u64)a * (u64)b2;
    return (u32)(x >> 32) + ((u32)x >> 1);
}
HD u32 mul0(u32 a, u32 b) { return fold(mul_lazy(a, b)); }     // C0,C0 -> C0
HD u32 mulc(u32 a, u32 b) { return canon(mul_lazy(a, b)); }    // C0,C0 -> C
// any 64-bit value < 2^63 -> C
HD u32 red64(u64 x) {
    u64 t = (x & M31_P) + (x >> 31);          // < 2^33
    u32 r = (u32)(t & M31_P) + (u32)(t >> 31); // <= p + 3
    return canon(r);
}
HD u32 sqr0(u32 a) { return mul0(a, a); }
// C -> C, a^(p-2).  p-2 = 2^31-3 = 0b111...1101 ; 37 multiplications
HD u32 inv(u32 a) {
    // addition chain on the exponent: 2^k-1 ladders (k = 1,2,4,8,16,24,28,29 -> final)
    u32 t1 = a;                                 // 2^1-1
    u32 t2 = mul0(sqr0(t1), t1);                // 2^2-1
    u32 t4 = t2;
    t4 = sqr0(sqr0(t4)); t4 = mul0(t4, t2);     // 2^4-1
    u32 t8 = t4;
    for (int i = 0; i < 4; i++) t8 = sqr0(t8);
    t8 = mul0(t8, t4);                          // 2^8-1
    u32 t16 = t8;
    for (int i = 0; i < 8; i++) t16 = sqr0(t16);
    t16 = mul0(t16, t8);                        // 2^16-1
    u32 t24 = t16;
    for (int i = 0; i < 8; i++) t24 = sqr0(t24);
    t24 = mul0(t24, t8);                        // 2^24-1
    u32 t28 = t24;
    for (int i = 0; i < 4; i++) t28 = sqr0(t28);
    t28 = mul0(t28, t4);                        // 2^28-1
    u32 t29 = mul0(sqr0(t28), t1);              // 2^29-1
    // 2^31-3 = (2^29-1)*4 + 1
    u32 r = sqr0(sqr0(t29));
    return canon(mul0(r, t1));
}

}  // namespace m31

// ---- CM31 (canonical in, canonical out) ---------------------------------------
struct cm31_t { u32 a, b; };
namespace cm31 {
HD cm31_t mk(u32 a, u32 b) { cm31_t r; r.a = a; r.b = b; return r; }
HD cm31_t add(cm31_t x, cm31_t y) { return mk(m31::addc(x.a, y.a), m31::addc(x.b, y.b)); }
HD cm31_t sub(cm31_t x, cm31_t y) { return mk(m31::subc(x.a, y.a), m31::subc(x.b, y.b)); }
HD cm31_t neg(cm31_t x) { return mk(m31::negc(x.a), m31::negc(x.b)); }
// (a+bi)(c+di): four 62-bit products summed in 64 bits, one reduction per word
HD cm31_t mul(cm31_t x, cm31_t y) {
    u64 ac = (u64)x.a * y.a, bd = (u64)x.b * y.b, ad = (u64)x.a * y.b, bc = (u64)x.b * y.a;
    // ac - bd  ==  ac + (p^2 - bd) ; p^2 < 2^62 keeps everything positive and < 2^63
    const u64 P2 = (u64)M31_P * M31_P;
    return mk(m31::red64(ac + (P2 - bd)), m31::red64(ad + bc));
}
HD cm31_t mul_m31(cm31_t x, u32 k) { return mk(m31::mulc(x.a, k), m31::mulc(x.b, k)); }
HD cm31_t inv(cm31_t x) {
    u32 n = m31::inv(m31::red64((u64)x.a * x.a + (u64)x.b * x.b));
    return mk(m31::mulc(x.a, n), m31::mulc(m31::negc(x.b), n));
}
}  // namespace cm31

// ---- QM31 -------------------------------------------------------------------------
struct qm31_t { u32 v[4]; };
namespace qm31 {
HD qm31_t mk(u32 a, u32 b, u32 c, u32 d) { qm31_t r; r.v[0] = a; r.v[1] = b; r.v[2] = c; r.v[3] = d; return r; }
HD qm31_t zero() { return mk(0, 0, 0, 0); }
HD qm31_t one() { return mk(1, 0, 0, 0); }
HD qm31_t from_m31(u32 a) { return mk(a, 0, 0, 0); }
HD qm31_t from_cm31(cm31_t lo, cm31_t hi) { return mk(lo.a, lo.b, hi.a, hi.b); }
HD cm31_t lo(qm31_t x) { return cm31::mk(x.v[0], x.v[1]); }
HD cm31_t hi(qm31_t x) { return cm31::mk(x.v[2], x.v[3]); }
HD bool eq(qm31_t x, qm31_t y) {
    return ((x.v[0] ^ y.v[0]) | (x.v[1] ^ y.v[1]) | (x.v[2] ^ y.v[2]) | (x.v[3] ^ y.v[3])) == 0;
}
HD bool is_zero(qm31_t x) { return (x.v[0] | x.v[1] | x.v[2] | x.v[3]) == 0; }
HD qm31_t add(qm31_t x, qm31_t y) {
    return mk(m31::addc(x.v[0], y.v[0]), m31::addc(x.v[1], y.v[1]), m31::addc(x.v[2], y.v[2]), m31::addc(x.v[3], y.v[3]));
}
HD qm31_t sub(qm31_t x, qm31_t y) {
    return mk(m31::subc(x.v[0], y.v[0]), m31::subc(x.v[1], y.v[1]), m31::subc(x.v[2], y.v[2]), m31::subc(x.v[3], y.v[3]));
}
HD qm31_t neg(qm31_t x) { return mk(m31::negc(x.v[0]), m31::negc(x.v[1]), m31::negc(x.v[2]), m31::negc(x.v[3])); }
HD qm31_t add_m31(qm31_t x, u32 k) { x.v[0] = m31::addc(x.v[0], k); return x; }
HD qm31_t sub_m31(qm31_t x, u32 k) { x.v[0] = m31::subc(x.v[0], k); return x; }
// (a + b u)(c + d u) = ac + (2+i) bd + (ad + bc) u
HD qm31_t mul(qm31_t x, qm31_t y) {
    cm31_t a = lo(x), b = hi(x), c = lo(y), d = hi(y);
    cm31_t bd = cm31::mul(b, d);
    cm31_t r = cm31::mk(m31::subc(m31::addc(bd.a, bd.a), bd.b), m31::addc(m31::addc(bd.b, bd.b), bd.a));
    return from_cm31(cm31::add(cm31::mul(a, c), r), cm31::add(cm31::mul(a, d), cm31::mul(b, c)));
}
HD qm31_t mul_m31(qm31_t x, u32 k) {
    return mk(m31::mulc(x.v[0], k), m31::mulc(x.v[1], k), m31::mulc(x.v[2], k), m31::mulc(x.v[3], k));
}
HD qm31_t mul_cm31(qm31_t x, cm31_t k) { return from_cm31(cm31::mul(lo(x), k), cm31::mul(hi(x), k)); }
HD qm31_t sqr(qm31_t x) { return mul(x, x); }
HD qm31_t inv(qm31_t x) {
    cm31_t a = lo(x), b = hi(x);
    cm31_t b2 = cm31::mul(b, b);
    cm31_t ib2 = cm31::mk(m31::subc(m31::addc(b2.a, b2.a), b2.b), m31::addc(m31::addc(b2.b, b2.b), b2.a));
    cm31_t den = cm31::inv(cm31::sub(cm31::mul(a, a), ib2));
    return from_cm31(cm31::mul(a, den), cm31::neg(cm31::mul(b, den)));
}
}  // namespace qm31

// ---- circle group over M31 (x^2 + y^2 = 1), generator of order 2^31 ------------------
struct cpoint_t { u32 x, y; };
namespace circle {
HD cpoint_t mk(u32 x, u32 y) { cpoint_t p; p.x = x; p.y = y; return p; }
HD cpoint_t add(cpoint_t p, cpoint_t q) {
    const u64 P2 = (u64)M31_P * M31_P;
    return mk(m31::red64((u64)p.x * q.x + (P2 - (u64)p.y * q.y)), m31::red64((u64)p.x * q.y + (u64)p.y * q.x));
}
HD cpoint_t dbl(cpoint_t p) { return add(p, p); }
HD cpoint_t conj(cpoint_t p) { return mk(p.x, m31::negc(p.y)); }
// k * G for the generator G = (2, 1268011823); k taken mod 2^31
HD cpoint_t mul_gen(u32 k) {
    cpoint_t r = mk(1, 0), g = mk(2u, 1268011823u);
    for (int i = 0; i < 31; i++) {
        if ((k >> i) & 1u) r = add(r, g);
        g = dbl(g);
    }
    return r;
}
HD u32 bitrev(u32 i, u32 n) {
    u32 r = 0;
    for (u32 b = 0; b < n; b++) r |= ((i >> b) & 1u) << (n - 1 - b);
    return r;
}
}  // namespace circle
