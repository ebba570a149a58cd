// The DSL gadgets of the verifier hot path over the recording constraint system (values live on the device):
//   BitsVar                     primitives/bits/src/lib.rs:10-138
//   Poseidon2HalfVar            primitives/poseidon31/src/lib.rs:16-438 (native variant)
//   Poseidon31MerkleHasherVar   primitives/merkle/src/lib.rs:9-181
//   ChannelVar                  primitives/channel/src/lib.rs:10-58
//   CirclePointM31Var/QM31Var   primitives/circle/src/lib.rs:16-251
//   PointCarryingQueryVar, QueryPositionsPerLogSizeVar   primitives/query/src/lib.rs:14-168
//   LinePolyVar                 primitives/line/src/lib.rs:10-67
#pragma once
#include <algorithm>
#include <map>

#include "fields.hpp"
#include "../../../include/stwo_b200_poseidon2_constants.h"

namespace stwo_b200 {
namespace dsl {

// ---- bits ----------------------------------------------------------------------------------------------------------------
struct BitsVar {
    ConstraintSystemRef cs;
    std::vector<u32> variables;

    // bits of `of` (an M31 variable): witnesses with the booleanity row b * (b - 1) = 0  (:25-43)
    static BitsVar new_witness_bits_of(const ConstraintSystemRef &cs, u32 of, u32 l) {
        BitsVar r{cs, {}};
        for (u32 k = 0; k < l; k++) {
            const u32 bit = cs->new_qm31(Def::bit(of, k), AllocationMode::Witness);
            r.variables.push_back(bit);
            const M31Var minus_one = M31Var::new_constant(cs, P - 1);
            const u32 bit_minus_one = cs->add(bit, minus_one.variable);
            cs->insert_gate(bit, bit_minus_one, 0, 0);
        }
        return r;
    }
    static BitsVar from_m31(const M31Var &v, u32 l) {                                     // :48-82
        const ConstraintSystemRef &cs = v.cs;
        BitsVar res = new_witness_bits_of(cs, v.variable, l);
        M31Var reconstructed(cs, res.variables[0]);
        for (u32 i = 1; i < l; i++) reconstructed = reconstructed + M31Var(cs, res.variables[i]).mul_constant(1u << i);
        reconstructed.equalverify(v);
        if (l == 31) {
            u32 product = cs->mul(res.variables[0], res.variables[1]);
            for (u32 i = 2; i < l; i++) product = cs->mul(product, res.variables[i]);
            cs->enforce_zero(product);
        }
        return res;
    }
    M31Var compose_range(u32 lo, u32 hi) const {                                          // :96-118
        u32 sum = variables[lo];
        for (u32 i = lo + 1, shift = 1; i < hi; i++, shift++) {
            const u32 shifted = cs->mul_constant(variables[i], 1u << shift);
            sum = cs->add(sum, shifted);
        }
        return {cs, sum};
    }
    BitsVar index_range(u32 lo, u32 hi) const { return {cs, std::vector<u32>(variables.begin() + lo, variables.begin() + hi)}; }
    BitsVar index_range_from(u32 lo) const { return index_range(lo, (u32)variables.size()); }
    u32 len() const { return (u32)variables.size(); }
};

// ---- Poseidon2 half states -----------------------------------------------------------------------------------------------
struct IsSwap { bool some; u32 bit_variable; };
inline IsSwap no_swap() { return {false, 0}; }
inline IsSwap swap_by(u32 bit_variable) { return {true, bit_variable}; }

// round constants as shape constants of the emulated permutation (primitives/poseidon31/src/parameters.rs:6-190)
namespace p2k {
static const u32 DIAG16[16] = STWO_P2_DIAG16;
static const u32 RC_FIRST[64] = STWO_P2_RC_FIRST;
static const u32 RC_PARTIAL[14] = STWO_P2_RC_PARTIAL;
static const u32 RC_LAST[64] = STWO_P2_RC_LAST;
}  // namespace p2k

// Native (Plonk-with-Poseidon: an accelerator flow entry) or emulated (Plonk-without-Poseidon: two QM31 limbs and ~401
// rows of m4 / pow5m4 / pow5 / hadamard / grandsum gates per permutation), chosen by the constraint system's type like the
// reference's enum Poseidon2HalfVar { Native, Emulated } (primitives/poseidon31/src/lib.rs:16-34).
struct Poseidon2HalfVar {
    enum Kind { Variables, InputWords, Ignored };
    ConstraintSystemRef cs;
    Kind kind = Variables;
    u32 left_variable = 0, right_variable = 0, sel_value = 0;
    u32 input_slot = 0;                        // InputWords: eight words of the witness stream

    // a Merkle sibling: value only, no variables, usable once (:51-60)
    static Poseidon2HalfVar new_single_use_witness_only(const ConstraintSystemRef &cs, u32 input_slot) {
        if (cs->without()) return new_variables(cs, input_slot, AllocationMode::Witness);
        Poseidon2HalfVar h; h.cs = cs; h.kind = InputWords; h.input_slot = input_slot; return h;
    }
    static Poseidon2HalfVar from_m31(const M31Var *s) {                                   // :76-105
        const QM31Var left = QM31Var::from_m31(s[0], s[1], s[2], s[3]);
        const QM31Var right = QM31Var::from_m31(s[4], s[5], s[6], s[7]);
        return assemble(left.cs, left.variable, right.variable);
    }
    static Poseidon2HalfVar from_qm31(const QM31Var &a, const QM31Var &b) { return assemble(a.cs.and_(b.cs), a.variable, b.variable); }   // :107-131
    // AllocVar::new_variables: two QM31 variables (+ the assembling row in the native case) (:142-188)
    static Poseidon2HalfVar new_variables(const ConstraintSystemRef &cs, u32 input_slot, AllocationMode mode) {
        const u32 l = cs->new_qm31(Def::input_qm31(input_slot), mode), r = cs->new_qm31(Def::input_qm31(input_slot + 4), mode);
        return assemble(cs, l, r);
    }
    static Poseidon2HalfVar new_witness(const ConstraintSystemRef &cs, u32 input_slot) { return new_variables(cs, input_slot, AllocationMode::Witness); }
    static Poseidon2HalfVar new_public_input(const ConstraintSystemRef &cs, u32 input_slot) { return new_variables(cs, input_slot, AllocationMode::PublicInput); }
    static Poseidon2HalfVar zero(const ConstraintSystemRef &cs) {                         // :191-218
        if (cs->without()) { Poseidon2HalfVar h; h.cs = cs; return h; }
        u32 sel;
        if (!cs.get_cache("poseidon2 zero_half", sel)) {
            sel = cs->assemble_poseidon_gate(0, 0);
            cs.set_cache("poseidon2 zero_half", sel);
        }
        Poseidon2HalfVar h; h.cs = cs; h.sel_value = sel; return h;
    }
    std::array<QM31Var, 2> to_qm31() const {                                              // :220-249
        if (kind != Variables) throw std::logic_error("half state without variables");
        return {QM31Var(cs, left_variable), QM31Var(cs, right_variable)};
    }
    void equalverify(const Poseidon2HalfVar &rhs) const {                                 // :425-438
        cs->insert_gate(left_variable, 0, rhs.left_variable, 1);
        cs->insert_gate(right_variable, 0, rhs.right_variable, 1);
    }

    // :282-407
    static std::pair<Poseidon2HalfVar, Poseidon2HalfVar> permute(const Poseidon2HalfVar &left, const Poseidon2HalfVar &right,
                                                                 bool ignore_left_result, bool ignore_right_result, IsSwap is_swap) {
        const ConstraintSystemRef &cs = left.cs.and_(right.cs);
        if (cs->without()) return poseidon_permute_emulated(left, right, is_swap);
        tape::Perm p{};
        left.describe(p.l_kind, p.l_a, p.l_b);
        right.describe(p.r_kind, p.r_a, p.r_b);
        p.swap_var = is_swap.some ? is_swap.bit_variable : tape::NO_VAR;
        const size_t tape_slot = cs->reserve_tape_slot();
        Poseidon2HalfVar out[2];
        const bool ignore[2] = {ignore_left_result, ignore_right_result};
        for (int h = 0; h < 2; h++) {
            out[h].cs = cs;
            if (ignore[h]) {
                out[h].kind = Ignored;
                p.out[2 * h] = p.out[2 * h + 1] = tape::NO_VAR;
            } else {
                const QM31Var l = QM31Var::new_witness(cs, Def::poseidon_out()), r = QM31Var::new_witness(cs, Def::poseidon_out());
                out[h] = assemble(cs, l.variable, r.variable);
                p.out[2 * h] = l.variable; p.out[2 * h + 1] = r.variable;
            }
        }
        cs->invoke_poseidon_accelerator({left.sel_value}, {right.sel_value}, {out[0].sel_value}, {out[1].sel_value},
                                        {is_swap.bit_variable, is_swap.some}, p, tape_slot);
        return {out[0], out[1]};
    }
    static Poseidon2HalfVar permute_get_rate(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r) { return permute(l, r, false, true, no_swap()).first; }
    static Poseidon2HalfVar permute_get_capacity(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r) { return permute(l, r, true, false, no_swap()).second; }
    static Poseidon2HalfVar swap_permute_get_rate(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r, IsSwap s) { return permute(l, r, false, true, s).first; }
    static Poseidon2HalfVar swap_permute_get_capacity(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r, IsSwap s) { return permute(l, r, true, false, s).second; }

    // ---- primitives/poseidon31/src/emulated.rs:12-221 -----------------------------------------------------------------
    static QM31Var apply_4x4_mds_matrix(const QM31Var &x) {
        const QM31Var constant = QM31Var::new_constant(x.cs, {{1, 1, 1, 1}});
        return {x.cs, x.cs->do_m4_gate(x.variable, constant.variable)};
    }
    static void apply_16x16_mds_matrix(QM31Var *state) {
        const QM31Var p1 = apply_4x4_mds_matrix(state[0]), p2 = apply_4x4_mds_matrix(state[1]);
        const QM31Var p3 = apply_4x4_mds_matrix(state[2]), p4 = apply_4x4_mds_matrix(state[3]);
        QM31Var t = p1 + p2;
        t = t + p3;
        t = t + p4;
        state[0] = p1 + t; state[1] = p2 + t; state[2] = p3 + t; state[3] = p4 + t;
    }
    static QM31Var pow4_witness(const ConstraintSystemRef &cs, u32 variable) { return QM31Var::new_witness(cs, {tape::T_POW4, variable, 0}); }
    static QM31Var pow5m4(const QM31Var &x) {
        const QM31Var b = pow4_witness(x.cs, x.variable);
        return {x.cs, x.cs->do_pow5m4_gate(x.variable, b.variable)};
    }
    static u32 pow5(const ConstraintSystemRef &cs, u32 variable) {
        const QM31Var b = pow4_witness(cs, variable);
        return cs->do_pow5_gate(variable, b.variable);
    }
    static void full_rounds(const ConstraintSystemRef &cs, QM31Var *state, const u32 *rc) {
        for (u32 r = 0; r < 4; r++) {
            for (u32 i = 0; i < 4; i++) {
                const u32 *k = rc + 16 * r + 4 * i;
                const QM31Var c = QM31Var::new_constant(cs, {{k[0], k[1], k[2], k[3]}});
                state[i] = state[i] + c;
            }
            for (u32 i = 0; i < 4; i++) state[i] = pow5m4(state[i]);
            QM31Var t = state[0] + state[1];
            t = t + state[2];
            t = t + state[3];
            const QM31Var s0 = state[0] + t, s1 = state[1] + t, s2 = state[2] + t, s3 = state[3] + t;
            state[0] = s0; state[1] = s1; state[2] = s2; state[3] = s3;
        }
    }
    static std::pair<Poseidon2HalfVar, Poseidon2HalfVar> poseidon_permute_emulated(const Poseidon2HalfVar &left, const Poseidon2HalfVar &right,
                                                                                   IsSwap is_swap) {
        const ConstraintSystemRef &cs = left.cs;
        QM31Var le[2] = {QM31Var(cs, left.left_variable), QM31Var(cs, left.right_variable)};
        QM31Var re[2] = {QM31Var(cs, right.left_variable), QM31Var(cs, right.right_variable)};
        QM31Var state[4];
        if (is_swap.some) {
            // (right - left) * bit + left
            const M31Var bit_var(cs, is_swap.bit_variable);
            const QM31Var d0 = re[0] - le[0], d1 = re[1] - le[1];
            const QM31Var db0 = d0 * bit_var, db1 = d1 * bit_var;
            const QM31Var nl0 = db0 + le[0], nl1 = db1 + le[1];
            const QM31Var nr0 = re[0] - db0, nr1 = re[1] - db1;
            state[0] = nl0; state[1] = nl1; state[2] = nr0; state[3] = nr1;
        } else { state[0] = le[0]; state[1] = le[1]; state[2] = re[0]; state[3] = re[1]; }
        const u32 limbs[4] = {state[0].variable, state[1].variable, state[2].variable, state[3].variable};
        cs->begin_macro();
        apply_16x16_mds_matrix(state);
        full_rounds(cs, state, p2k::RC_FIRST);
        for (u32 r = 0; r < 14; r++) {
            u32 first_limb_with_first_only = cs->do_hadamard(state[0].variable, 1);
            const QM31Var mask = QM31Var::new_constant(cs, {{0, 1, 1, 1}});
            const u32 first_limb_without_first = cs->do_hadamard(state[0].variable, mask.variable);
            const M31Var rc = M31Var::new_constant(cs, p2k::RC_PARTIAL[r]);
            first_limb_with_first_only = cs->add(first_limb_with_first_only, rc.variable);
            first_limb_with_first_only = pow5(cs, first_limb_with_first_only);
            state[0] = QM31Var(cs, cs->add(first_limb_with_first_only, first_limb_without_first));
            const u32 sum_1 = cs->do_grandsum_gate(state[0].variable, state[1].variable);
            const u32 sum_2 = cs->do_grandsum_gate(state[2].variable, state[3].variable);
            const u32 sum = cs->add(sum_1, sum_2);
            for (u32 i = 0; i < 4; i++) {
                const u32 *k = p2k::DIAG16 + 4 * i;
                const QM31Var diag = QM31Var::new_constant(cs, {{k[0], k[1], k[2], k[3]}});
                u32 v = cs->do_hadamard(state[i].variable, diag.variable);
                v = cs->add(sum, v);
                state[i] = QM31Var(cs, v);
            }
        }
        full_rounds(cs, state, p2k::RC_LAST);
        cs->end_macro(limbs);
        Poseidon2HalfVar out_left, out_right;
        out_left.cs = out_right.cs = cs;
        out_left.left_variable = state[0].variable; out_left.right_variable = state[1].variable;
        out_right.left_variable = state[2].variable; out_right.right_variable = state[3].variable;
        return {out_left, out_right};
    }

private:
    static Poseidon2HalfVar assemble(const ConstraintSystemRef &cs, u32 l, u32 r) {
        Poseidon2HalfVar h; h.cs = cs; h.left_variable = l; h.right_variable = r;
        if (!cs->without()) h.sel_value = cs->assemble_poseidon_gate(l, r);      // the emulated half is just its two limbs
        return h;
    }
    void describe(u32 &kind_, u32 &a, u32 &b) const {
        if (kind == Ignored) throw std::logic_error("an ignored permutation output carries no value to hash");
        if (kind == InputWords) { kind_ = 1; a = input_slot; b = 0; }
        else { kind_ = 0; a = left_variable; b = right_variable; }
    }
};
using HashVar = Poseidon2HalfVar;                                                        // channel/src/lib.rs:7

// ---- Merkle hasher ---------------------------------------------------------------------------------------------------
struct Poseidon31MerkleHasherVar {
    static Poseidon2HalfVar hash_tree(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r) { return Poseidon2HalfVar::permute_get_rate(l, r); }
    static Poseidon2HalfVar hash_tree_with_column(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r, const Poseidon2HalfVar &hash_column) {
        return Poseidon2HalfVar::permute_get_rate(Poseidon2HalfVar::permute_get_rate(l, r), hash_column);
    }
    static Poseidon2HalfVar hash_tree_with_swap(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r, u32 bit_variable) {
        return Poseidon2HalfVar::swap_permute_get_rate(l, r, swap_by(bit_variable));
    }
    static Poseidon2HalfVar hash_tree_with_column_hash_with_swap(const Poseidon2HalfVar &l, const Poseidon2HalfVar &r, u32 bit_variable,
                                                                 const Poseidon2HalfVar &column_hash) {
        const Poseidon2HalfVar hash_tree = Poseidon2HalfVar::swap_permute_get_rate(l, r, swap_by(bit_variable));
        return Poseidon2HalfVar::permute_get_rate(hash_tree, column_hash);
    }
    static Poseidon2HalfVar combine_hash_tree_with_column(const Poseidon2HalfVar &hash_tree, const Poseidon2HalfVar &hash_column) {
        return Poseidon2HalfVar::permute_get_rate(hash_tree, hash_column);
    }
    // the sponge walk shared by the four hash_*_columns_* functions (:51-180): WIDTH items per absorbed half
    template <class Item, u32 WIDTH, class Make>
    static Poseidon2HalfVar sponge(const std::vector<Item> &items, const Item &zero_item, Make make) {
        const ConstraintSystemRef &cs = items[0].cs;
        const u32 len = (u32)items.size(), num_chunk = (len + WIDTH - 1) / WIDTH;
        std::vector<Item> input(WIDTH, zero_item);
        for (u32 k = 0; k < std::min(len, WIDTH); k++) input[k] = items[k];
        const Poseidon2HalfVar zero = Poseidon2HalfVar::zero(cs);
        const Poseidon2HalfVar first_chunk = make(input);
        Poseidon2HalfVar digest = Poseidon2HalfVar::permute_get_capacity(first_chunk, zero);
        if (num_chunk == 1) return digest;
        for (u32 c = 1; c + 1 < num_chunk; c++) {
            const std::vector<Item> chunk(items.begin() + c * WIDTH, items.begin() + (c + 1) * WIDTH);
            digest = Poseidon2HalfVar::permute_get_capacity(make(chunk), digest);
        }
        const u32 remain = len % WIDTH;
        std::vector<Item> last(WIDTH, zero_item);
        if (remain == 0) for (u32 k = 0; k < WIDTH; k++) last[k] = items[len - WIDTH + k];
        else for (u32 k = 0; k < remain; k++) last[k] = items[len - remain + k];
        return Poseidon2HalfVar::permute_get_capacity(make(last), digest);
    }
    static Poseidon2HalfVar hash_m31_columns_get_capacity(const std::vector<M31Var> &m31) {
        return sponge<M31Var, 8>(m31, M31Var::zero(m31[0].cs), [](const std::vector<M31Var> &c) { return Poseidon2HalfVar::from_m31(c.data()); });
    }
    static Poseidon2HalfVar hash_m31_columns_get_rate(const std::vector<M31Var> &m31) {
        const Poseidon2HalfVar digest = hash_m31_columns_get_capacity(m31);
        return Poseidon2HalfVar::permute_get_rate(Poseidon2HalfVar::zero(m31[0].cs), digest);
    }
    static Poseidon2HalfVar hash_qm31_columns_get_capacity(const std::vector<QM31Var> &q) {
        return sponge<QM31Var, 2>(q, QM31Var::zero(q[0].cs), [](const std::vector<QM31Var> &c) { return Poseidon2HalfVar::from_qm31(c[0], c[1]); });
    }
    static Poseidon2HalfVar hash_qm31_columns_get_rate(const std::vector<QM31Var> &q) {
        const Poseidon2HalfVar digest = hash_qm31_columns_get_capacity(q);
        return Poseidon2HalfVar::permute_get_rate(Poseidon2HalfVar::zero(q[0].cs), digest);
    }
};

// ---- channel -----------------------------------------------------------------------------------------------------------
struct ChannelVar {
    u32 n_sent = 0;
    Poseidon2HalfVar digest;
    explicit ChannelVar(const ConstraintSystemRef &cs) : digest(Poseidon2HalfVar::zero(cs)) { cs->transcript_slot = 0; }
    const ConstraintSystemRef &cs() const { return digest.cs; }
    // every channel operation is one permutation, the k-th of the native transcript (fiat_shamir.cuh)
    void hint() const { cs()->push_hint(cs()->transcript_slot++); }
    void mix_root(const HashVar &root) { hint(); digest = Poseidon2HalfVar::permute_get_capacity(root, digest); n_sent = 0; }
    std::array<QM31Var, 2> draw_felts() {
        const M31Var n = M31Var::new_constant(cs(), n_sent);
        n_sent += 1;
        const Poseidon2HalfVar left = Poseidon2HalfVar::from_qm31(QM31Var::from(n), QM31Var::zero(cs()));
        hint();
        return Poseidon2HalfVar::permute_get_rate(left, digest).to_qm31();
    }
    void mix_one_felt(const QM31Var &felt) {
        const Poseidon2HalfVar left = Poseidon2HalfVar::from_qm31(felt, QM31Var::zero(cs()));
        hint();
        digest = Poseidon2HalfVar::permute_get_capacity(left, digest);
        n_sent = 0;
    }
    void mix_two_felts(const QM31Var &a, const QM31Var &b) {
        hint();
        digest = Poseidon2HalfVar::permute_get_capacity(Poseidon2HalfVar::from_qm31(a, b), digest);
        n_sent = 0;
    }
};

// ---- circle points -------------------------------------------------------------------------------------------------
struct CirclePointM31 { u32 x, y; };                               // a shape constant (stwo CirclePoint<M31>)
inline CirclePointM31 cp_add(CirclePointM31 a, CirclePointM31 b) {
    const u64 xx = (u64)a.x * b.x % P, yy = (u64)a.y * b.y % P, xy = (u64)a.x * b.y % P, yx = (u64)a.y * b.x % P;
    return {(u32)((xx + P - yy) % P), (u32)((xy + yx) % P)};
}
inline CirclePointM31 cp_double(CirclePointM31 a) { return cp_add(a, a); }
inline CirclePointM31 cp_neg(CirclePointM31 a) { return {a.x, m31_neg(a.y)}; }
// generator of the subgroup of order 2^k: 2^(31-k) * (2, 1268011823)   (stwo M31_CIRCLE_GEN)
inline CirclePointM31 cp_subgroup_gen(u32 k) {
    CirclePointM31 g{2u, 1268011823u};
    for (u32 i = k; i < 31; i++) g = cp_double(g);
    return g;
}

struct CirclePointM31Var {
    M31Var x, y;
    static CirclePointM31Var new_constant(const ConstraintSystemRef &cs, CirclePointM31 p) {
        const M31Var x = M31Var::new_constant(cs, p.x);
        const M31Var y = M31Var::new_constant(cs, p.y);
        return {x, y};
    }
    CirclePointM31Var operator+(const CirclePointM31Var &rhs) const {                     // circle/src/lib.rs:46-57
        const M31Var x1x2 = x * rhs.x, y1y2 = y * rhs.y, x1y2 = x * rhs.y, y1x2 = y * rhs.x;
        const M31Var new_x = x1x2 - y1y2;
        const M31Var new_y = x1y2 + y1x2;
        return {new_x, new_y};
    }
    CirclePointM31Var double_() const {                                                  // :61-69
        const M31Var xx = x * x, yy = y * y, xy = x * y;
        const M31Var new_x = xx - yy;
        return {new_x, xy.mul_constant(2)};
    }
    // :74-104.  The reference takes the row's op constant from the SELECTED value (point.x - 1 / point.y when the bit
    // is 1, 0 / 0 when it is 0): the value is const * bit either way, the op column of these two rows follows it.
    static CirclePointM31Var select(const ConstraintSystemRef &cs, CirclePointM31 point, u32 bit_variable) {
        u32 new_x = cs->mul_constant(bit_variable, (point.x + P - 1) % P, true);
        new_x = cs->add(new_x, 1);
        const u32 new_y = cs->mul_constant(bit_variable, point.y, true);
        return {M31Var(cs, new_x), M31Var(cs, new_y)};
    }
    CirclePointM31Var conditional_negate(u32 bit_variable) const {                        // :106-130
        const ConstraintSystemRef &cs = x.cs;
        u32 y_multiplier = cs->mul_constant(bit_variable, P - 2);
        y_multiplier = cs->add(y_multiplier, 1);
        return {x, M31Var(cs, cs->mul(y_multiplier, y.variable))};
    }
};

struct CirclePointQM31Var {
    QM31Var x, y;
    static CirclePointQM31Var new_witness(const ConstraintSystemRef &cs, u32 input_slot) {
        const QM31Var x = QM31Var::new_witness(cs, Def::input_qm31(input_slot));
        const QM31Var y = QM31Var::new_witness(cs, Def::input_qm31(input_slot + 4));
        return {x, y};
    }
    static CirclePointQM31Var from_t(const QM31Var &t) {                                  // :204-219
        const ConstraintSystemRef &cs = t.cs;
        const QM31Var t_doubled = t + t;
        const QM31Var t_squared = t * t;
        const QM31Var t_squared_plus_1 = t_squared + M31Var::one(cs);
        const QM31Var t_squared_plus_1_inverse = t_squared_plus_1.inv();
        const QM31Var one_minus_t_squared = (-t_squared) + M31Var::one(cs);
        const QM31Var px = one_minus_t_squared * t_squared_plus_1_inverse;
        const QM31Var py = t_doubled * t_squared_plus_1_inverse;
        return {px, py};
    }
    static CirclePointQM31Var from_channel(ChannelVar &channel) { return from_t(channel.draw_felts()[0]); }
    QM31Var repeated_double_x_only(u32 log_size) const {                                  // :226-234
        QM31Var cur = x;
        for (u32 k = 0; k < log_size; k++) {
            const QM31Var sq = cur * cur;
            cur = (sq + sq) - M31Var::one(cur.cs);
        }
        return cur;
    }
    CirclePointQM31Var operator+(const CirclePointM31 &rhs) const {                       // :236-250
        const QM31Var x1x2 = x.mul_constant_m31(rhs.x), y1y2 = y.mul_constant_m31(rhs.y);
        const QM31Var x1y2 = x.mul_constant_m31(rhs.y), y1x2 = y.mul_constant_m31(rhs.x);
        const QM31Var new_x = x1x2 - y1y2;
        const QM31Var new_y = x1y2 + y1x2;
        return {new_x, new_y};
    }
};

// ---- query positions -------------------------------------------------------------------------------------------------
struct PointCarryingQueryVar {
    BitsVar bits;
    CirclePointM31 last_step;
    CirclePointM31Var point;

    static PointCarryingQueryVar new_(const BitsVar &bits) {                              // query/src/lib.rs:56-139
        const ConstraintSystemRef &cs = bits.cs;
        const u32 log_size = bits.len();
        // CanonicCoset::new(log_size + 1).circle_domain().half_coset = half_odds(log_size)
        const CirclePointM31 initial = cp_subgroup_gen(log_size + 2), step = cp_subgroup_gen(log_size);
        std::vector<CirclePointM31> steps;
        CirclePointM31 cur_step = step;
        for (u32 k = 0; k + 1 < log_size; k++) { steps.push_back(cur_step); cur_step = cp_double(cur_step); }
        // steps zipped with bits[1..] reversed
        CirclePointM31Var cur = CirclePointM31Var::new_constant(cs, initial);
        const u32 n = (u32)steps.size();
        for (u32 k = 0; k < n; k += 2) {
            const u32 bit0 = bits.variables[log_size - 1 - k];
            if (k + 1 == n) {
                const CirclePointM31Var point = CirclePointM31Var::select(cs, steps[k], bit0);
                cur = point + cur;
            } else {
                const u32 bit1 = bits.variables[log_size - 2 - k];
                const CirclePointM31 p00{1, 0}, p01 = steps[k], p10 = steps[k + 1], p11 = cp_add(p01, p10);
                const u32 a = bit0, b = bit1;
                const u32 one_minus_a = cs->add(1, cs->mul_constant(a, P - 1));
                const u32 one_minus_b = cs->add(1, cs->mul_constant(b, P - 1));
                const u32 b00 = cs->mul(one_minus_a, one_minus_b), b01 = cs->mul(a, one_minus_b);
                const u32 b10 = cs->mul(one_minus_a, b), b11 = cs->mul(a, b);
                u32 px = cs->mul_constant(b00, p00.x);
                px = cs->add(px, cs->mul_constant(b01, p01.x));
                px = cs->add(px, cs->mul_constant(b10, p10.x));
                px = cs->add(px, cs->mul_constant(b11, p11.x));
                u32 py = cs->mul_constant(b00, p00.y);
                py = cs->add(py, cs->mul_constant(b01, p01.y));
                py = cs->add(py, cs->mul_constant(b10, p10.y));
                py = cs->add(py, cs->mul_constant(b11, p11.y));
                const CirclePointM31Var point{M31Var(cs, px), M31Var(cs, py)};
                cur = point + cur;
            }
        }
        return {bits, cp_neg(steps.back()), cur};
    }
    CirclePointM31Var get_next_point() const { return point.double_().conditional_negate(bits.variables[0]); }   // :140-144
    M31Var get_next_point_x() const {                                                     // :145-149
        const M31Var xx = point.x * point.x, yy = point.y * point.y;
        return xx - yy;
    }
    void next() {                                                                         // :150-162
        const ConstraintSystemRef &cs = bits.cs;
        const CirclePointM31Var t = CirclePointM31Var::select(cs, last_step, bits.variables[1]);
        bits = bits.index_range_from(1);
        point = (point + t).double_();
    }
    CirclePointM31Var get_absolute_point() const { return point; }
};

struct QueryPositionsPerLogSizeVar {
    std::map<u32, std::vector<PointCarryingQueryVar>> points;
    QueryPositionsPerLogSizeVar(u32 min_degree, u32 max_degree, const std::vector<M31Var> &raw_queries) {   // :19-38
        std::vector<PointCarryingQueryVar> elems;
        for (const M31Var &raw : raw_queries) elems.push_back(PointCarryingQueryVar::new_(BitsVar::from_m31(raw, 31).index_range(0, max_degree)));
        points[max_degree] = elems;
        for (u32 log_size = max_degree; log_size-- > min_degree;) {
            for (auto &e : elems) e.next();
            points[log_size] = elems;
        }
    }
    const std::vector<PointCarryingQueryVar> &operator[](u32 log_size) const { return points.at(log_size); }
};

// ---- line polynomial -------------------------------------------------------------------------------------------------
struct LinePolyVar {
    ConstraintSystemRef cs;
    std::vector<QM31Var> coeffs;
    QM31Var eval_at_point(const M31Var &x0) const {                                       // line/src/lib.rs:39-67
        M31Var x = x0;
        u32 lg = 0;
        while ((1u << (lg + 1)) <= coeffs.size()) lg++;
        std::vector<M31Var> doublings{x};
        for (u32 k = 1; k < lg; k++) {
            const M31Var x_sq = x * x;
            x = x_sq + x_sq;
            x = x + M31Var::new_constant(cs, P - 1);
            doublings.push_back(x);
        }
        return fold(coeffs.data(), (u32)coeffs.size(), doublings.data());
    }
private:
    static QM31Var fold(const QM31Var *values, u32 n, const M31Var *factors) {
        if (n == 1) return values[0];
        const QM31Var lhs = fold(values, n / 2, factors + 1);
        const QM31Var rhs = fold(values + n / 2, n / 2, factors + 1);
        return lhs + (rhs * factors[0]);
    }
};

}  // namespace dsl
}  // namespace stwo_b200
