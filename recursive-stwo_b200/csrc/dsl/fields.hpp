// M31Var / CM31Var / QM31Var over the recording constraint system: the kept field-variable API of
// primitives/fields/src/m31.rs:8-180, cm31.rs:11-279, qm31.rs:12-469, without the host-side `value` field (values are
// evaluated on the device from the tape).  Row order and the a_wire/b_wire order of every operator follow the
// reference exactly, including its habit of implementing `low op high` as `high op low`.
#pragma once
#include <array>
#include <type_traits>

#include "constraint_system.hpp"

namespace stwo_b200 {
namespace dsl {

constexpr u32 P = M31_P;
inline u32 m31_neg(u32 v) { return v % P ? P - v % P : 0; }
inline u32 m31_mul(u32 a, u32 b) { return (u32)((u64)a * b % P); }
inline u32 m31_pow(u32 a, u32 e) { u32 r = 1; while (e) { if (e & 1) r = m31_mul(r, a); a = m31_mul(a, a); e >>= 1; } return r; }
inline u32 m31_inverse(u32 a) { return m31_pow(a, P - 2); }

template <int RANK> struct FieldVarBase {
    static constexpr int rank = RANK;
    ConstraintSystemRef cs;
    u32 variable = 0;
};
struct M31Var; struct CM31Var; struct QM31Var;
template <class T> struct is_field_var : std::false_type {};
template <> struct is_field_var<M31Var> : std::true_type {};
template <> struct is_field_var<CM31Var> : std::true_type {};
template <> struct is_field_var<QM31Var> : std::true_type {};
template <class A, class B> using Wider = std::conditional_t<(A::rank >= B::rank), A, B>;

struct M31Var : FieldVarBase<0> {
    M31Var() {}
    M31Var(const ConstraintSystemRef &c, u32 v) { cs = c; variable = v; }
    static M31Var zero(const ConstraintSystemRef &cs) { return {cs, 0}; }
    static M31Var one(const ConstraintSystemRef &cs) { return {cs, 1}; }
    static M31Var new_witness(const ConstraintSystemRef &cs, const Def &d) { return {cs, cs->new_m31(d, AllocationMode::Witness)}; }
    static M31Var new_public_input(const ConstraintSystemRef &cs, const Def &d) { return {cs, cs->new_m31(d, AllocationMode::PublicInput)}; }
    static M31Var new_constant(const ConstraintSystemRef &cs, u32 value) {                 // m31.rs:33-59
        value %= P;
        if (value == 0) return zero(cs);
        if (value == 1) return one(cs);
        const std::string key = "m31 " + std::to_string(value);
        u32 var;
        if (cs.get_cache(key, var)) return {cs, var};
        var = cs->new_m31_constant(value);
        cs.set_cache(key, var);
        return {cs, var};
    }
    void equalverify(const M31Var &rhs) const { cs->insert_gate(variable, 0, rhs.variable, 1); }   // :123-127
    M31Var inv() const {                                                                    // :129-136
        M31Var res = new_witness(cs, Def::inv_m31(variable));
        cs->insert_gate(variable, res.variable, 1, 0);
        return res;
    }
    M31Var mul_constant(u32 k) const { return {cs, cs->mul_constant(variable, k)}; }        // :138-147
    M31Var is_zero() const;                                                                  // :153-167
    M31Var is_eq(const M31Var &rhs) const;
};

struct CM31Var : FieldVarBase<1> {
    CM31Var() {}
    CM31Var(const ConstraintSystemRef &c, u32 v) { cs = c; variable = v; }
    static CM31Var zero(const ConstraintSystemRef &cs) { return {cs, 0}; }
    static CM31Var one(const ConstraintSystemRef &cs) { return {cs, 1}; }
    static CM31Var i(const ConstraintSystemRef &cs) { return {cs, 2}; }
    static CM31Var from(const M31Var &v) { return {v.cs, v.variable}; }                     // cm31.rs:77-86
    static CM31Var from_m31(const M31Var &re, const M31Var &im) {                           // :192-202
        const ConstraintSystemRef &cs = re.cs.and_(im.cs);
        return {cs, cs->add(re.variable, cs->mul(im.variable, 2))};
    }
    static CM31Var new_witness(const ConstraintSystemRef &cs, const Def &re, const Def &im) {   // :27-37
        const M31Var r = M31Var::new_witness(cs, re), m = M31Var::new_witness(cs, im);
        return {cs, cs->add(r.variable, cs->mul(m.variable, 2))};
    }
    static CM31Var new_constant(const ConstraintSystemRef &cs, u32 re, u32 im) {            // :39-74
        re %= P; im %= P;
        if (re == 0 && im == 0) return zero(cs);
        if (re == 1 && im == 0) return one(cs);
        if (re == 0 && im == 1) return i(cs);
        const std::string key = "cm31 " + std::to_string(re) + "," + std::to_string(im);
        u32 var;
        if (cs.get_cache(key, var)) return {cs, var};
        const M31Var r = M31Var::new_constant(cs, re), m = M31Var::new_constant(cs, im);
        var = cs->add(r.variable, cs->mul(m.variable, 2));
        cs.set_cache(key, var);
        return {cs, var};
    }
    void equalverify(const CM31Var &rhs) const { cs->insert_gate(variable, 0, rhs.variable, 1); }
    // cm31.rs:238-243: a fresh witness, NOT tied to self by a row (the reference relies on a later use)
    CM31Var inv() const { return new_witness(cs, Def::inv_cm31_re(variable), Def::inv_cm31_im(variable)); }
    CM31Var shift_by_i() const { return {cs, cs->mul(variable, 2)}; }                       // :245-252
    CM31Var mul_constant_m31(u32 k) const { return {cs, cs->mul_constant(variable, k)}; }
    CM31Var mul_constant_cm31(u32 re, u32 im) const {                                       // :264-276
        const CM31Var a = mul_constant_m31(re), b = mul_constant_m31(im);
        return {cs, cs->add(a.variable, cs->mul(b.variable, 2))};
    }
};

struct QM31Var : FieldVarBase<2> {
    QM31Var() {}
    QM31Var(const ConstraintSystemRef &c, u32 v) { cs = c; variable = v; }
    static QM31Var zero(const ConstraintSystemRef &cs) { return {cs, 0}; }
    static QM31Var one(const ConstraintSystemRef &cs) { return {cs, 1}; }
    static QM31Var i(const ConstraintSystemRef &cs) { return {cs, 2}; }
    static QM31Var j(const ConstraintSystemRef &cs) { return {cs, 3}; }
    static QM31Var from(const M31Var &v) { return {v.cs, v.variable}; }                     // qm31.rs:75-84
    static QM31Var new_witness(const ConstraintSystemRef &cs, const Def &d) { return {cs, cs->new_qm31(d, AllocationMode::Witness)}; }
    static QM31Var new_public_input(const ConstraintSystemRef &cs, const Def &d) { return {cs, cs->new_qm31(d, AllocationMode::PublicInput)}; }
    static QM31Var new_constant(const ConstraintSystemRef &cs, const QM31Const &v0) {       // :35-73
        QM31Const v = {{v0.v[0] % P, v0.v[1] % P, v0.v[2] % P, v0.v[3] % P}};
        const bool hi0 = v.v[2] == 0 && v.v[3] == 0;
        if (hi0 && v.v[0] == 0 && v.v[1] == 0) return zero(cs);
        if (hi0 && v.v[0] == 1 && v.v[1] == 0) return one(cs);
        if (hi0 && v.v[0] == 0 && v.v[1] == 1) return i(cs);
        if (v.v[0] == 0 && v.v[1] == 0 && v.v[2] == 1 && v.v[3] == 0) return j(cs);
        const std::string key = "qm31 " + std::to_string(v.v[0]) + "," + std::to_string(v.v[1]) + "," + std::to_string(v.v[2]) + "," + std::to_string(v.v[3]);
        u32 var;
        if (cs.get_cache(key, var)) return {cs, var};
        var = cs->new_qm31_constant(v);
        cs.set_cache(key, var);
        return {cs, var};
    }
    static QM31Var from_m31(const M31Var &a0, const M31Var &a1, const M31Var &a2, const M31Var &a3) {   // :245-256
        const ConstraintSystemRef &cs = a0.cs;
        const u32 l = cs->add(a0.variable, cs->mul(a1.variable, 2));
        const u32 r = cs->mul(cs->add(a2.variable, cs->mul(a3.variable, 2)), 3);
        return {cs, cs->add(l, r)};
    }
    static QM31Var from_cm31(const CM31Var &a, const CM31Var &b) { return {a.cs, a.cs->add(a.variable, a.cs->mul(b.variable, 3))}; }   // :300-307
    std::array<M31Var, 4> decompose_m31() const {                                           // :258-272
        std::array<M31Var, 4> a;
        for (u32 k = 0; k < 4; k++) a[k] = M31Var::new_witness(cs, Def::coordinate(variable, k));
        const u32 l = cs->add(a[0].variable, cs->mul(a[1].variable, 2));
        const u32 r = cs->mul(cs->add(a[2].variable, cs->mul(a[3].variable, 2)), 3);
        cs->insert_gate(l, r, variable, 1);
        return a;
    }
    std::array<CM31Var, 2> decompose_cm31() const;                                          // :274-281
    QM31Var pow(unsigned __int128 exp) const;                                               // :283-298
    void equalverify(const QM31Var &rhs) const { cs->insert_gate(variable, 0, rhs.variable, 1); }   // :345-350
    QM31Var inv() const {                                                                   // :352-359
        QM31Var res = new_witness(cs, Def::inv_qm31(variable));
        cs->insert_gate(variable, res.variable, 1, 0);
        return res;
    }
    QM31Var mul_constant_m31(u32 k) const { return {cs, cs->mul_constant(variable, k)}; }   // :361-368
    QM31Var mul_constant_cm31(u32 re, u32 im) const {                                       // :370-381
        const QM31Var a = mul_constant_m31(re), b = mul_constant_m31(im);
        return {cs, cs->add(a.variable, cs->mul(b.variable, 2))};
    }
    QM31Var mul_constant_qm31(const QM31Const &k) const {                                   // :383-392 (uncached constant)
        const u32 kv = cs->new_qm31_constant(k);
        return {cs, cs->mul(variable, kv)};
    }
    QM31Var shift_by_i() const { return {cs, cs->mul(variable, 2)}; }                       // :394-401
    QM31Var shift_by_j() const { return {cs, cs->mul(variable, 3)}; }                       // :403-410
    QM31Var shift_by_ij() const { return shift_by_i().shift_by_j(); }                       // :466-468
    static QM31Var select(const QM31Var &a, const QM31Var &b, u32 bit_variable);            // :412-435
    static std::pair<QM31Var, QM31Var> swap(const QM31Var &a, const QM31Var &b, u32 bit_variable);   // :437-464
};

// ---- operators: one generic implementation per reference impl family ---------------------------------------------------
template <class A, class = std::enable_if_t<is_field_var<A>::value>> A operator-(const A &a) {   // Neg = mul_constant(-1)
    return A(a.cs, a.cs->mul_constant(a.variable, P - 1));
}
template <class A, class B, class = std::enable_if_t<is_field_var<A>::value && is_field_var<B>::value>>
Wider<A, B> operator+(const A &a, const B &b) {
    const ConstraintSystemRef &cs = a.cs.and_(b.cs);
    if (A::rank < B::rank) return Wider<A, B>(cs, cs->add(b.variable, a.variable));      // `rhs + self`
    return Wider<A, B>(cs, cs->add(a.variable, b.variable));
}
template <class A, class B, class = std::enable_if_t<is_field_var<A>::value && is_field_var<B>::value>>
Wider<A, B> operator*(const A &a, const B &b) {
    const ConstraintSystemRef &cs = a.cs.and_(b.cs);
    if (A::rank < B::rank) return Wider<A, B>(cs, cs->mul(b.variable, a.variable));      // `rhs * self`
    return Wider<A, B>(cs, cs->mul(a.variable, b.variable));
}
template <class A, class B, class = std::enable_if_t<is_field_var<A>::value && is_field_var<B>::value>>
Wider<A, B> operator-(const A &a, const B &b) {                                          // Sub = self + &(-rhs)
    const B n = -b;
    return a + n;
}

inline M31Var M31Var::is_zero() const {
    // the hint is the inverse, or 0 for 0 (m31::inv(0) = 0 on the device)
    const M31Var inv = new_witness(cs, Def::inv_m31(variable));
    const M31Var out = (-(*this * inv)) + M31Var::one(cs);
    cs->insert_gate(variable, out.variable, 0, 0);
    return out;
}
inline M31Var M31Var::is_eq(const M31Var &rhs) const { return (*this - rhs).is_zero(); }

inline std::array<CM31Var, 2> QM31Var::decompose_cm31() const {
    const std::array<M31Var, 4> v = decompose_m31();
    const CM31Var a0 = CM31Var::from(v[1]).shift_by_i() + v[0];
    const CM31Var a1 = CM31Var::from(v[3]).shift_by_i() + v[2];
    return {a0, a1};
}
inline QM31Var QM31Var::pow(unsigned __int128 exp) const {
    std::vector<bool> bools;
    while (exp > 0) { bools.push_back((exp & 1) != 0); exp >>= 1; }
    QM31Var cur = QM31Var::one(cs);
    for (size_t k = bools.size(); k-- > 0;) {
        if (bools[k]) cur = cur * *this;
        if (k != 0) cur = cur * cur;
    }
    return cur;
}
inline QM31Var QM31Var::select(const QM31Var &a, const QM31Var &b, u32 bit_variable) {
    const ConstraintSystemRef &cs = a.cs.and_(b.cs);
    const QM31Var b_minus_a = b - a;
    u32 variable = cs->mul(b_minus_a.variable, bit_variable);
    variable = cs->add(a.variable, variable);
    return {cs, variable};
}
inline std::pair<QM31Var, QM31Var> QM31Var::swap(const QM31Var &a, const QM31Var &b, u32 bit_variable) {
    const ConstraintSystemRef &cs = a.cs.and_(b.cs);
    const QM31Var b_minus_a = b - a;
    u32 left = cs->mul(b_minus_a.variable, bit_variable);
    u32 right = cs->mul_constant(left, P - 1);
    left = cs->add(a.variable, left);
    right = cs->add(b.variable, right);
    return {QM31Var(cs, left), QM31Var(cs, right)};
}

}  // namespace dsl
}  // namespace stwo_b200
