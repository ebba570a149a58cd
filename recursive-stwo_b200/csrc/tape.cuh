// K6: evaluation of the constraint system's `variables[]` for a batch of witness streams.
//
// Replaces the value side of every DSL operation of the reference -- the `value` field each *Var computes on the host
// while it appends rows: primitives/fields/src/{m31,cm31,qm31}.rs (add/mul/neg/inv/decompose), primitives/bits/src/lib.rs:48-82
// (bit decomposition), primitives/poseidon31/src/lib.rs:282-407 (permutation + PoseidonEntry hashes + SwapOption.swap),
// constraint_system/src/plonk_with_poseidon.rs:141-281 (variables.push of add/mul/mul_constant/new_m31/new_qm31).
//
// A tape instruction defines one variable from earlier ones (or from the item's witness stream).  Device layout is
// lane-interleaved: element e of batch item b lives at ((b / lanes) * n_elems + e) * lanes + b % lanes, so the 32 lanes
// of a warp (32 batch items executing the same instruction) touch one contiguous 512-byte (QM31) or 128-byte (word) run.
// lanes = 1 gives the plain per-item arrays the single-system host entry points use.
// HD: tests/hostsim runs the very same evaluator on the CPU box.
#pragma once
#include "poseidon2.cuh"

namespace tape {

enum : u32 {
    T_NONE = 0,
    T_ADD, T_MUL, T_MULC,                       // dst = a + b | a * b | a * imm(b)          (gate outputs)
    T_INPUT_M31, T_INPUT_QM31,                  // dst = witness stream word(s) at slot a
    T_INV_M31, T_INV_QM31,                      // dst = 1 / var a                           (M31Var::inv, QM31Var::inv)
    T_INV_CM31_RE, T_INV_CM31_IM,               // dst = re / im of 1 / (CM31 part of var a) (CM31Var::inv -> new_witness)
    T_COORD,                                    // dst = coordinate b of var a               (QM31Var::decompose_m31)
    T_BIT,                                      // dst = bit b of var a                      (BitsVar::from_m31)
    T_POSEIDON,                                 // permutation record dst                    (Poseidon2HalfVar::permute)
    // gates of the Plonk-without-Poseidon system (constraint_system/src/plonk_without_poseidon.rs:108-245)
    T_M4,                                       // dst = M4 MDS block on the coordinates of a
    T_POW5M4,                                   // dst = M4(a (.) b)          (b = a^4 coordinate-wise)
    T_HADAMARD,                                 // dst = a (.) b              (do_hadamard, do_pow5_gate)
    T_GRANDSUM,                                 // dst = (s, s, s, s), s = sum of the 8 coordinates of a and b
    T_POW4,                                     // dst = a^4 coordinate-wise  (the witness of pow5m4 / pow5, emulated.rs:37-78)
    T_EPOSEIDON,                                // one whole emulated permutation: record dst defines its 401 variables
    // the two halves of a RECORDED permutation (record dst names a slot of the item's permutation record), used by the second tape
    // order only (dsl::RecordedCircuit::recorded), which runs for items whose record is complete:
    T_PERM_OUT,                                 // output variables <- the recorded output state; reads no variable
    T_PERM_FLOW,                                // the flow entry (halves as given, recorded output, swap bit); defines no variable
};
// A whole poseidon_permute_emulated call (primitives/poseidon31/src/emulated.rs:104-221, after the swap) as ONE instruction:
// the 401 variables it creates form a dependency chain ~170 levels deep, which one warp walks in registers instead of the
// grid synchronising 170 times.  Record: the four limb variables, then the 401 created variables in creation order
// (11 initial MDS + 8 x 19 full rounds + 14 x 17 partial rounds).
constexpr u32 EPOSEIDON_VARS = 401, EPOSEIDON_REC = 4 + EPOSEIDON_VARS;
constexpr u32 NO_VAR = 0xffffffffu;

struct Ins { u32 op, dst, a, b; };

// One Poseidon2 invocation.  A half state is either two QM31 variables (kind 0; the zero half is variables (0, 0)) or
// eight words of the witness stream (kind 1: Poseidon2HalfVar::new_single_use_witness_only, a Merkle sibling).
struct Perm {
    u32 l_kind, l_a, l_b;
    u32 r_kind, r_a, r_b;
    u32 swap_var;                               // NO_VAR: never swapped
    u32 out[4];                                 // variables receiving state[4k .. 4k+4); NO_VAR when the half is ignored
    u32 hint;                                   // 0: none; k + 1: the output state is word 16 k .. 16 k + 16 of the item's permutation
                                                // hints (the native verifier already executed this permutation: verify.cuh HintLayout)
};

struct alignas(16) Q4 { u32 x, y, z, w; };

// Where a word of the verifier circuit's witness stream comes from: a section of the proof blob or of the batched
// verifier's workspace (circuit.cuh resolves it).  section:4 | a:6 (tree / FRI layer) | i:7 (query / column) | k:15 (word)
enum : u32 {
    S_STMT0 = 0, S_STMT1, S_COMMITMENT, S_SAMPLED, S_FRI_COMMITMENT, S_LAST_COEFFS, S_POW_LIMB, S_OODS,
    S_PATH_COL, S_PATH_SIB, S_PAIR_SELF, S_PAIR_SIB, S_PAIR_HASH,
    S_FS,        // Fiat-Shamir outputs of the native verifier: k < FS_QUERY_BASE: word k of {oods_t, z, alpha, random_coeff,
                 // after_sampled_values_random_coeff, fri_alphas[..]}; FS_QUERY_BASE + i: position of query i at the largest
                 // log size; FS_ZERO: the constant 0 (padding of packed public inputs)
    S_EXTRA,     // word k of the per-proof public-input hashes the last-layer circuit takes (circuit.cuh, k_last_extra)
    S_ANSWER,    // word k of the first-layer FRI answer of query i at the a-th log size (descending): the folding-stage circuit takes
                 // the answers as witnesses (dsl::record_folding_circuit)
};
static_assert(S_ANSWER < 16, "a source section is four bits");
constexpr u32 FS_QUERY_BASE = 1024, FS_ZERO = 4095;
HD u32 src_pack(u32 section, u32 a, u32 i, u32 k) { return section | (a << 4) | (i << 10) | (k << 17); }
HD u32 src_section(u32 s) { return s & 15u; }
HD u32 src_a(u32 s) { return (s >> 4) & 63u; }
HD u32 src_i(u32 s) { return (s >> 10) & 127u; }
HD u32 src_k(u32 s) { return s >> 17; }

// One batch item's window into the lane-interleaved arrays.
struct View {
    Q4 *vars;                                   // + element * stride
    const u32 *input;                           // + word * stride
    Q4 *flow_hash;                              // + (entry * 8 + quad) * stride: 16-byte elements, like vars     may be null
    uint8_t *flow_swap;                         // + entry * stride                   may be null
    u32 stride;
    const u32 *hint;                            // the item's permutation hints, 16 consecutive words per slot; may be null
};

HD qm31_t ldv(const View &v, u32 i) {
    const Q4 t = v.vars[(size_t)i * v.stride];
    return qm31::mk(t.x, t.y, t.z, t.w);
}
HD void stv(const View &v, u32 i, qm31_t q) {
    Q4 t; t.x = q.v[0]; t.y = q.v[1]; t.z = q.v[2]; t.w = q.v[3];
    v.vars[(size_t)i * v.stride] = t;
}
HD u32 ldw(const View &v, u32 slot) { return v.input[(size_t)slot * v.stride]; }

// ---- coordinate-wise gates of the Plonk-without-Poseidon system -----------------------------------------------------------
HD qm31_t q_m4(qm31_t x) {                      // plonk_without_poseidon.rs:115-125
    const u32 t0 = m31::addc(x.v[0], x.v[1]), t1 = m31::addc(x.v[2], x.v[3]);
    const u32 t2 = m31::addc(m31::addc(x.v[1], x.v[1]), t1), t3 = m31::addc(m31::addc(x.v[3], x.v[3]), t0);
    const u32 t1x2 = m31::addc(t1, t1), t0x2 = m31::addc(t0, t0);
    const u32 t4 = m31::addc(m31::addc(t1x2, t1x2), t3), t5 = m31::addc(m31::addc(t0x2, t0x2), t2);
    return qm31::mk(m31::addc(t3, t5), t5, m31::addc(t2, t4), t4);
}
HD qm31_t q_had(qm31_t a, qm31_t b) { return qm31::mk(m31::mulc(a.v[0], b.v[0]), m31::mulc(a.v[1], b.v[1]), m31::mulc(a.v[2], b.v[2]), m31::mulc(a.v[3], b.v[3])); }
HD qm31_t q_pow4(qm31_t a) { const qm31_t s = q_had(a, a); return q_had(s, s); }
HD qm31_t q_grandsum(qm31_t a, qm31_t b) {
    const u32 s = m31::addc(m31::addc(m31::addc(a.v[0], a.v[1]), m31::addc(a.v[2], a.v[3])), m31::addc(m31::addc(b.v[0], b.v[1]), m31::addc(b.v[2], b.v[3])));
    return qm31::mk(s, s, s, s);
}

// variables 0..3 = 0, 1, i, j (plonk_with_poseidon.rs:63-66)
HD void prologue(const View &v) {
    stv(v, 0, qm31::mk(0, 0, 0, 0)); stv(v, 1, qm31::mk(1, 0, 0, 0)); stv(v, 2, qm31::mk(0, 1, 0, 0)); stv(v, 3, qm31::mk(0, 0, 1, 0));
}

// a permutation record of the tape: 12 words, 16-byte aligned, the same for every lane -> three 16-byte loads on the device
HD Perm ld_perm(const Perm *p) {
#if defined(__CUDA_ARCH__)
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    Perm r;
    r.l_kind = a.x; r.l_a = a.y; r.l_b = a.z; r.r_kind = a.w; r.r_a = b.x; r.r_b = b.y; r.swap_var = b.z; r.out[0] = b.w;
    r.out[1] = c.x; r.out[2] = c.y; r.out[3] = c.z; r.hint = c.w;
    return r;
#else
    return *p;
#endif
}
// one slot of the permutation record: 16 words, 64-byte aligned, contiguous per item -> four 16-byte loads on the device
HD void ld_slot(const u32 *h, u32 *st) {
#if defined(__CUDA_ARCH__)
    const uint4 *h4 = reinterpret_cast<const uint4 *>(h);
    const uint4 a = h4[0], b = h4[1], c = h4[2], d = h4[3];
    st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w; st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
    st[8] = c.x; st[9] = c.y; st[10] = c.z; st[11] = c.w; st[12] = d.x; st[13] = d.y; st[14] = d.z; st[15] = d.w;
#else
    for (int k = 0; k < 16; k++) st[k] = h[k];
#endif
}
// 16 words of a flow entry (quads q0 .. q0 + 3 of its eight 16-byte elements)
HD void st_flow(const View &v, u32 entry, u32 q0, const u32 *w) {
    Q4 *f = v.flow_hash + (size_t)(entry * 8 + q0) * v.stride;
    for (int q = 0; q < 4; q++) { Q4 t; t.x = w[4 * q]; t.y = w[4 * q + 1]; t.z = w[4 * q + 2]; t.w = w[4 * q + 3]; f[(size_t)q * v.stride] = t; }
}
HD void load_half(const View &v, u32 kind, u32 a, u32 b, u32 *h) {
    if (kind == 0) {
        const qm31_t l = ldv(v, a), r = ldv(v, b);
        for (int k = 0; k < 4; k++) { h[k] = l.v[k]; h[4 + k] = r.v[k]; }
    } else {
        for (u32 k = 0; k < 8; k++) h[k] = ldw(v, a + k);
    }
}

template <bool UNROLLED>
HD void eval_poseidon(const View &v, const Perm &p, u32 entry) {
    u32 st[16], in[16];
    load_half(v, p.l_kind, p.l_a, p.l_b, in);
    load_half(v, p.r_kind, p.r_a, p.r_b, in + 8);
    const bool swap = p.swap_var != NO_VAR && ldv(v, p.swap_var).v[0] != 0;
    for (int k = 0; k < 8; k++) { st[k] = swap ? in[8 + k] : in[k]; st[8 + k] = swap ? in[k] : in[8 + k]; }
    if (v.flow_hash) st_flow(v, entry, 0, in);                                  // PoseidonEntry 1, 2: the halves as given
    if (v.hint && p.hint) {                     // executed once already by the native verifier: take its output state
        ld_slot(v.hint + (size_t)(p.hint - 1) * 16, st);
    } else poseidon2::permute<UNROLLED>(st);
    if (v.flow_hash) st_flow(v, entry, 4, st);                                  // PoseidonEntry 3, 4: the full output
    if (v.flow_swap) v.flow_swap[(size_t)entry * v.stride] = swap ? 1 : 0;
    for (int k = 0; k < 4; k++)
        if (p.out[k] != NO_VAR) stv(v, p.out[k], qm31::mk(st[4 * k], st[4 * k + 1], st[4 * k + 2], st[4 * k + 3]));
}

// T_PERM_OUT / T_PERM_FLOW: eval_poseidon of a recorded permutation in two independent steps (v.hint is set: the order that holds these
// instructions is chosen only for items whose record is complete)
HD void eval_perm_out(const View &v, const Perm &p) {
    u32 h[16];
    ld_slot(v.hint + (size_t)(p.hint - 1) * 16, h);
    for (int k = 0; k < 4; k++)
        if (p.out[k] != NO_VAR) stv(v, p.out[k], qm31::mk(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]));
}
HD void eval_perm_flow(const View &v, const Perm &p, u32 entry) {
    u32 in[16];
    load_half(v, p.l_kind, p.l_a, p.l_b, in);
    load_half(v, p.r_kind, p.r_a, p.r_b, in + 8);
    const bool swap = p.swap_var != NO_VAR && ldv(v, p.swap_var).v[0] != 0;
    if (v.flow_hash) {
        st_flow(v, entry, 0, in);
        u32 h[16];
        ld_slot(v.hint + (size_t)(p.hint - 1) * 16, h);
        st_flow(v, entry, 4, h);
    }
    if (v.flow_swap) v.flow_swap[(size_t)entry * v.stride] = swap ? 1 : 0;
}

HD void eval_eposeidon(const View &v, const u32 *rec) {
    const u32 *out = rec + 4;
    u32 k = 0;
    qm31_t s[4];
    for (int i = 0; i < 4; i++) s[i] = ldv(v, rec[i]);
    auto emit = [&](qm31_t x) { stv(v, out[k++], x); return x; };
    auto sum_mix = [&]() {                       // t = s0 + s1 + s2 + s3 (three rows), s_i += t
        qm31_t t = emit(qm31::add(s[0], s[1]));
        t = emit(qm31::add(t, s[2]));
        t = emit(qm31::add(t, s[3]));
        for (int i = 0; i < 4; i++) s[i] = emit(qm31::add(s[i], t));
    };
    auto full_rounds = [&](const u32 *rc) {      // emulated.rs:121-150 / :190-219
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < 4; i++) {
                const u32 *c = rc + 16 * r + 4 * i;
                s[i] = emit(qm31::add(s[i], qm31::mk(c[0], c[1], c[2], c[3])));
            }
            for (int i = 0; i < 4; i++) {
                const qm31_t b = emit(q_pow4(s[i]));
                s[i] = emit(q_m4(q_had(s[i], b)));
            }
            sum_mix();
        }
    };
    for (int i = 0; i < 4; i++) s[i] = emit(q_m4(s[i]));                 // apply_16x16_mds_matrix (:24-35)
    sum_mix();
    full_rounds(poseidon2::K.rc_first);
    for (int r = 0; r < 14; r++) {                                      // :151-189
        const qm31_t first_only = emit(qm31::mk(s[0].v[0], 0, 0, 0));   // hadamard with the variable 1
        const qm31_t without_first = emit(qm31::mk(0, s[0].v[1], s[0].v[2], s[0].v[3]));
        const qm31_t a = emit(qm31::add(first_only, qm31::from_m31(poseidon2::K.rc_part[r])));
        const qm31_t b = emit(q_pow4(a));
        const qm31_t g = emit(q_had(a, b));
        s[0] = emit(qm31::add(g, without_first));
        const qm31_t sum_1 = emit(q_grandsum(s[0], s[1])), sum_2 = emit(q_grandsum(s[2], s[3]));
        const qm31_t sum = emit(qm31::add(sum_1, sum_2));
        for (int i = 0; i < 4; i++) {
            const u32 *d = poseidon2::K.diag + 4 * i;
            const qm31_t hv = emit(q_had(s[i], qm31::mk(d[0], d[1], d[2], d[3])));
            s[i] = emit(qm31::add(sum, hv));
        }
    }
    full_rounds(poseidon2::K.rc_last);
}

template <bool UNROLLED = false>
HD void eval(const View &v, const Ins &in, const Perm *perms, const u32 *eperms = nullptr) {
    switch (in.op) {
    case T_ADD: stv(v, in.dst, qm31::add(ldv(v, in.a), ldv(v, in.b))); break;
    case T_MUL: stv(v, in.dst, qm31::mul(ldv(v, in.a), ldv(v, in.b))); break;
    case T_MULC: stv(v, in.dst, qm31::mul_m31(ldv(v, in.a), in.b)); break;
    case T_INPUT_M31: stv(v, in.dst, qm31::from_m31(ldw(v, in.a))); break;
    case T_INPUT_QM31: stv(v, in.dst, qm31::mk(ldw(v, in.a), ldw(v, in.a + 1), ldw(v, in.a + 2), ldw(v, in.a + 3))); break;
    case T_INV_M31: stv(v, in.dst, qm31::from_m31(m31::inv(ldv(v, in.a).v[0]))); break;
    case T_INV_QM31: stv(v, in.dst, qm31::inv(ldv(v, in.a))); break;
    case T_INV_CM31_RE: stv(v, in.dst, qm31::from_m31(cm31::inv(qm31::lo(ldv(v, in.a))).a)); break;
    case T_INV_CM31_IM: stv(v, in.dst, qm31::from_m31(cm31::inv(qm31::lo(ldv(v, in.a))).b)); break;
    case T_COORD: stv(v, in.dst, qm31::from_m31(ldv(v, in.a).v[in.b & 3u])); break;
    case T_BIT: stv(v, in.dst, qm31::from_m31((ldv(v, in.a).v[0] >> (in.b & 31u)) & 1u)); break;
    case T_POSEIDON: eval_poseidon<UNROLLED>(v, ld_perm(perms + in.dst), in.dst); break;
    case T_M4: stv(v, in.dst, q_m4(ldv(v, in.a))); break;
    case T_POW5M4: stv(v, in.dst, q_m4(q_had(ldv(v, in.a), ldv(v, in.b)))); break;
    case T_HADAMARD: stv(v, in.dst, q_had(ldv(v, in.a), ldv(v, in.b))); break;
    case T_GRANDSUM: stv(v, in.dst, q_grandsum(ldv(v, in.a), ldv(v, in.b))); break;
    case T_POW4: stv(v, in.dst, q_pow4(ldv(v, in.a))); break;
    case T_EPOSEIDON: eval_eposeidon(v, eperms + (size_t)in.dst * EPOSEIDON_REC); break;
    case T_PERM_OUT: eval_perm_out(v, ld_perm(perms + in.dst)); break;
    case T_PERM_FLOW: eval_perm_flow(v, ld_perm(perms + in.dst), in.dst); break;
    default: break;
    }
}

// ---- the O(n_rows) loops of the constraint system, per batch item ------------------------------------------------------
// check_arithmetics (constraint_system/src/plonk_with_poseidon.rs:337-380): one row
HD bool gate_ok(qm31_t va, qm31_t vb, qm31_t vc, u32 op, u32 enforce_c_m31) {
    // op is a row constant, i.e. uniform across a warp of batch items: the two pure gates skip half of the general formula
    qm31_t want;
    if (op == 1) want = qm31::add(va, vb);
    else if (op == 0) want = qm31::mul(va, vb);
    else want = qm31::add(qm31::mul_m31(qm31::add(va, vb), op), qm31::mul_m31(qm31::mul(va, vb), m31::subc(1, op)));
    bool ok = qm31::eq(want, vc);
    if (enforce_c_m31 && (vc.v[1] | vc.v[2] | vc.v[3])) ok = false;
    return ok;
}
// The same check split over TWO lanes: lane `part` (0 / 1) verifies the CM31 half (coordinates 2 part, 2 part + 1) of c.  Both parts run
// the same instruction stream (operands are selected, not branched on), so a warp = 16 items x 2 parts stays convergent on one row, and
// each lane pays for two of the four CM31 products of the QM31 multiplication.  ok(part 0) && ok(part 1) == gate_ok.
HD bool gate_ok_half(qm31_t va, qm31_t vb, qm31_t vc, u32 op, u32 enforce_c_m31, u32 part) {
    const cm31_t a_lo = qm31::lo(va), a_hi = qm31::hi(va), b_lo = qm31::lo(vb), b_hi = qm31::hi(vb);
    const cm31_t got = part ? qm31::hi(vc) : qm31::lo(vc);
    cm31_t want;
    if (op == 1) want = part ? cm31::add(a_hi, b_hi) : cm31::add(a_lo, b_lo);
    else {
        // lo = a_lo b_lo + (2 + i) a_hi b_hi ; hi = a_lo b_hi + a_hi b_lo
        const cm31_t y1 = part ? b_hi : b_lo, y2 = part ? b_lo : b_hi;
        const cm31_t p1 = cm31::mul(a_lo, y1), p2 = cm31::mul(a_hi, y2);
        const cm31_t tw = cm31::mk(m31::subc(m31::addc(p2.a, p2.a), p2.b), m31::addc(m31::addc(p2.b, p2.b), p2.a));
        const cm31_t prod = cm31::add(p1, part ? p2 : tw);
        if (op == 0) want = prod;
        else {
            const cm31_t sum = part ? cm31::add(a_hi, b_hi) : cm31::add(a_lo, b_lo);
            want = cm31::add(cm31::mul_m31(sum, op), cm31::mul_m31(prod, m31::subc(1, op)));
        }
    }
    bool ok = want.a == got.a && want.b == got.b;
    if (enforce_c_m31 && (part ? (vc.v[2] | vc.v[3]) : vc.v[1])) ok = false;
    return ok;
}
// check_arithmetics of the Plonk-without-Poseidon system (plonk_without_poseidon.rs:410-599): the selectors pick one of six gates
HD bool gate_ok_without(qm31_t a, qm31_t b, qm31_t c, u32 op1, u32 op2, u32 op3, u32 op4) {
    if (op2 > 1 || op3 > 1 || op4 > 1) return false;
    const u32 sel = op2 * 4 + op3 * 2 + op4;
    if (sel && op1 != 1) return false;
    switch (sel) {
    case 0: return qm31::eq(c, qm31::add(qm31::mul_m31(qm31::add(a, b), op1), qm31::mul_m31(qm31::mul(a, b), m31::subc(1, op1))));
    case 1: return qm31::eq(c, q_had(a, b));                                         // hadamard
    case 2: return qm31::eq(c, q_m4(q_had(a, b)));                                   // m4
    case 3: return qm31::eq(c, q_grandsum(a, b));                                    // grand sum
    case 5: return qm31::eq(c, q_had(a, b)) && qm31::eq(b, q_pow4(a));               // pow5
    case 6: return qm31::eq(c, q_m4(q_had(a, b))) && qm31::eq(b, q_pow4(a));         // pow5m4
    default: return false;
    }
}
HD bool row_ok(const View &v, u32 a, u32 b, u32 c, u32 op, u32 enforce_c_m31, bool op_follows_c) {
    const qm31_t va = ldv(v, a), vb = ldv(v, b), vc = ldv(v, c);
    return gate_ok(va, vb, vc, op_follows_c ? vc.v[0] : op, enforce_c_m31);
}

}  // namespace tape
