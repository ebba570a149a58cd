// TEST INFRASTRUCTURE: compiles the kernels' HD (host+device) arithmetic headers with the host
// compiler so the CPU-only test tier can check the exact device arithmetic (lazy-reduction range
// classes included) against the oracle.  Never linked into the product library.
#include "../../recursive-stwo_b200/csrc/merkle.cuh"
#include <stddef.h>
extern "C" {
void hs_poseidon2_permute(u32 *st, size_t n) { for (size_t i = 0; i < n; i++) poseidon2::permute<true>(st + 16 * i); }
void hs_poseidon2_permute_rolled(u32 *st, size_t n) { for (size_t i = 0; i < n; i++) poseidon2::permute<false>(st + 16 * i); }
u32 hs_m31_inv(u32 a) { return m31::inv(a); }
u32 hs_m31_mul(u32 a, u32 b) { return m31::mulc(a, b); }
u32 hs_m31_red64(u64 x) { return m31::red64(x); }
void hs_cm31_mul(const u32 *a, const u32 *b, u32 *o) {
    cm31_t r = cm31::mul(cm31::mk(a[0], a[1]), cm31::mk(b[0], b[1]));
    o[0] = r.a; o[1] = r.b;
}
void hs_qm31_mul(const u32 *a, const u32 *b, u32 *o) {
    qm31_t x = qm31::mk(a[0], a[1], a[2], a[3]), y = qm31::mk(b[0], b[1], b[2], b[3]);
    qm31_t r = qm31::mul(x, y);
    for (int i = 0; i < 4; i++) o[i] = r.v[i];
}
void hs_qm31_inv(const u32 *a, u32 *o) {
    qm31_t r = qm31::inv(qm31::mk(a[0], a[1], a[2], a[3]));
    for (int i = 0; i < 4; i++) o[i] = r.v[i];
}
void hs_circle_mul_gen(u32 k, u32 *o) { cpoint_t p = circle::mul_gen(k); o[0] = p.x; o[1] = p.y; }
void hs_hash_node(const u32 *children, const u32 *cols, u32 n_cols, u32 *out) {
    merkle::hash_node(children, [&](u32 c) { return cols[c]; }, n_cols, out);
}
void hs_path_root(const stwo_b200_path_shape *shape, u32 index, const u32 *cols, const u32 *sib, u32 *out) {
    merkle::path_root(*shape, index, cols, sib, out);
}
}
