// The recursive verifier circuit (components/recursive/*), recorded once per proof shape.
//
//   PlonkWithPoseidonProofVar, SinglePathMerkleProofVar, SinglePairMerkleProofVar, DecommitmentVar, LookupElementsVar
//                               components/recursive/data_structures/src/lib.rs:20-520
//   FiatShamirResults::compute  components/recursive/fiat_shamir/src/lib.rs:31-176
//   CompositionCheck::compute   components/recursive/composition/src/{lib,data_structures,plonk,poseidon}.rs
//   AnswerResults::compute      components/recursive/answer/src/{lib,data_structures}.rs
//   FoldingResults::compute     components/recursive/folding/src/lib.rs:12-205
//   driver                      examples/single-proof/src/main.rs:33-90, examples/multi-proofs/src/main.rs:49-139
//
// Where the reference passes host values (the proof, FiatShamirHints, DecommitHints, FirstLayerHints, InnerLayersHints)
// this recorder passes *where the value will be* on the device: a slot of the per-proof witness stream, described by a
// gather word (tape::Src) that names a section of the proof blob or of the batched verifier's workspace (verify.cuh),
// which already holds every hint (per-query Merkle paths, pair openings, OODS point).
#pragma once
#include <functional>

#include "gadgets.hpp"
#include "../verify.cuh"

namespace stwo_b200 {
namespace dsl {

struct ProofShape { u32 log_size_plonk, log_size_poseidon, pow_bits, log_blowup, log_last, n_queries, n_inner; };   // = stwo_b200_proof_shape
struct PublicInput { u32 idx; QM31Const value; };

struct ShapeFacts {                       // the value-independent part of FiatShamirHints (hints/fiat_shamir.rs:93-104,239-305)
    ProofShape s;
    u32 max_first, log_plonk, log_pos, composition_log_degree_bound;
    std::vector<u32> all_log_sizes;       // ascending
    verify::HintLayout hints;             // where the native verifier records the permutations this circuit repeats
    explicit ShapeFacts(const ProofShape &p) : s(p) {
        hints = verify::hint_layout(verify::Shape{p.log_size_plonk, p.log_size_poseidon, p.pow_bits, p.log_blowup, p.log_last, p.n_queries, p.n_inner});
        max_first = p.log_last + p.log_blowup + 1 + p.n_inner;
        log_plonk = p.log_size_plonk + p.log_blowup; log_pos = p.log_size_poseidon + p.log_blowup;
        composition_log_degree_bound = max_first - p.log_blowup + 1;
        all_log_sizes = {log_plonk, log_pos, max_first};
        std::sort(all_log_sizes.begin(), all_log_sizes.end());
        all_log_sizes.erase(std::unique(all_log_sizes.begin(), all_log_sizes.end()), all_log_sizes.end());
    }
    static u32 n_cols(u32 t) { return t == 0 ? 50u : t == 1 ? 60u : t == 2 ? 16u : 8u; }
    static u32 plonk_cols(u32 t) { return t == 0 ? 10u : t == 1 ? 12u : t == 2 ? 8u : 0u; }
    static u32 n_masks(u32 t, u32 c) { return (t == 2 && (c & 4)) ? 2u : 1u; }          // last logup batch: [-1, 0]
    u32 tree_depth(u32 t) const { return t == 3 ? max_first : std::max(log_plonk, log_pos); }
    u32 column_log_size(u32 t, u32 c) const { return t == 3 ? max_first : c < plonk_cols(t) ? log_plonk : log_pos; }
    bool fri_first_has_data(u32 h) const { return h == max_first || h == log_plonk || h == log_pos; }
};

// The recorder's view of the witness stream: allocates slots and remembers where each word comes from.
struct WitnessStream {
    ConstraintSystemRef cs;
    std::vector<u32> gather;              // per stream word: tape::Src packed
    u32 take(u32 section, u32 a, u32 i, u32 k0, u32 n) {
        const u32 slot = cs->new_input_words(n);
        for (u32 k = 0; k < n; k++) gather.push_back(tape::src_pack(section, a, i, k0 + k));
        return slot;
    }
};

struct PlonkWithPoseidonProofVar {        // data_structures/src/lib.rs:90-213
    M31Var log_size_plonk, log_size_poseidon;
    QM31Var plonk_total_sum, poseidon_total_sum;
    std::vector<HashVar> commitments;
    std::vector<std::vector<std::vector<QM31Var>>> sampled_values;
    HashVar first_layer_commitment;
    std::vector<HashVar> inner_layer_commitments;
    LinePolyVar last_poly;
    std::array<M31Var, 3> proof_of_work;

    static PlonkWithPoseidonProofVar new_witness(WitnessStream &w, const ShapeFacts &f) {
        const ConstraintSystemRef &cs = w.cs;
        PlonkWithPoseidonProofVar v;
        v.log_size_plonk = M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_STMT0, 0, 0, 0, 1)));
        v.log_size_poseidon = M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_STMT0, 0, 0, 1, 1)));
        v.plonk_total_sum = QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_STMT1, 0, 0, 0, 4)));
        v.poseidon_total_sum = QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_STMT1, 0, 0, 4, 4)));
        for (u32 t = 0; t < 4; t++) v.commitments.push_back(HashVar::new_witness(cs, w.take(tape::S_COMMITMENT, t, 0, 0, 8)));
        v.sampled_values.resize(4);
        for (u32 t = 0; t < 4; t++)
            for (u32 c = 0; c < ShapeFacts::n_cols(t); c++) {
                std::vector<QM31Var> col;
                for (u32 m = 0; m < ShapeFacts::n_masks(t, c); m++)
                    col.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_SAMPLED, t, c, 4 * m, 4))));
                v.sampled_values[t].push_back(col);
            }
        v.first_layer_commitment = HashVar::new_witness(cs, w.take(tape::S_FRI_COMMITMENT, 0, 0, 0, 8));
        for (u32 l = 0; l < f.s.n_inner; l++) v.inner_layer_commitments.push_back(HashVar::new_witness(cs, w.take(tape::S_FRI_COMMITMENT, 1 + l, 0, 0, 8)));
        v.last_poly.cs = cs;
        for (u32 k = 0; k < (1u << f.s.log_last); k++) v.last_poly.coeffs.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_LAST_COEFFS, 0, 0, 4 * k, 4))));
        for (u32 j = 0; j < 3; j++) v.proof_of_work[j] = M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_POW_LIMB, 0, 0, j, 1)));
        return v;
    }
};

struct LookupElementsVar {                // data_structures/src/lib.rs:215-262
    QM31Var z, alpha;
    std::array<QM31Var, 3> alpha_powers;
    static LookupElementsVar draw(ChannelVar &channel) {
        const std::array<QM31Var, 2> za = channel.draw_felts();
        LookupElementsVar r;
        r.z = za[0]; r.alpha = za[1];
        r.alpha_powers[0] = QM31Var::one(r.z.cs);
        r.alpha_powers[1] = r.alpha;
        r.alpha_powers[2] = r.alpha * r.alpha;
        return r;
    }
};

struct FiatShamirResults {                // fiat_shamir/src/lib.rs:12-25
    LookupElementsVar lookup_elements;
    QM31Var random_coeff, after_sampled_values_random_coeff;
    CirclePointQM31Var oods_point;
    std::vector<M31Var> raw_queries;
    std::vector<QM31Var> fri_alphas;

    static void mix_felts(ChannelVar &channel, const std::vector<QM31Var> &felts) {
        for (size_t k = 0; k < felts.size(); k += 2) {
            if (k + 1 == felts.size()) channel.mix_one_felt(felts[k]);
            else channel.mix_two_felts(felts[k], felts[k + 1]);
        }
    }
    static FiatShamirResults compute(PlonkWithPoseidonProofVar &proof, const ShapeFacts &f, const std::vector<std::pair<u32, QM31Var>> &inputs) {
        const ConstraintSystemRef &cs = proof.log_size_plonk.cs;
        FiatShamirResults r;
        ChannelVar channel(cs);
        channel.mix_root(proof.commitments[0]);
        channel.mix_one_felt(QM31Var::from(proof.log_size_plonk));                       // stmt0.mix_into
        channel.mix_one_felt(QM31Var::from(proof.log_size_poseidon));
        channel.mix_root(proof.commitments[1]);
        r.lookup_elements = LookupElementsVar::draw(channel);
        channel.mix_two_felts(proof.plonk_total_sum, proof.poseidon_total_sum);          // stmt1.mix_into
        channel.mix_root(proof.commitments[2]);
        r.random_coeff = channel.draw_felts()[0];
        channel.mix_root(proof.commitments[3]);
        r.oods_point = CirclePointQM31Var::from_channel(channel);
        std::vector<QM31Var> flat;
        for (const auto &tree : proof.sampled_values)
            for (const auto &col : tree)
                for (const QM31Var &v : col) flat.push_back(v);
        mix_felts(channel, flat);
        r.after_sampled_values_random_coeff = channel.draw_felts()[0];
        channel.mix_root(proof.first_layer_commitment);
        r.fri_alphas.push_back(channel.draw_felts()[0]);
        for (const HashVar &l : proof.inner_layer_commitments) {
            channel.mix_root(l);
            r.fri_alphas.push_back(channel.draw_felts()[0]);
        }
        mix_felts(channel, proof.last_poly.coeffs);
        const QM31Var nonce_felt = QM31Var::from_m31(proof.proof_of_work[0], proof.proof_of_work[1], proof.proof_of_work[2], M31Var::zero(cs));
        BitsVar::from_m31(proof.proof_of_work[0], 22);
        BitsVar::from_m31(proof.proof_of_work[1], 21);
        BitsVar::from_m31(proof.proof_of_work[2], 21);
        channel.mix_one_felt(nonce_felt);
        const M31Var lower_bits = BitsVar::from_m31(channel.digest.to_qm31()[0].decompose_m31()[0], 31).compose_range(0, f.s.pow_bits);
        lower_bits.equalverify(M31Var::zero(cs));
        std::vector<QM31Var> draw_queries_felts;
        for (u32 k = 0; k < (f.s.n_queries + 3) / 4; k++) {
            const std::array<QM31Var, 2> ab = channel.draw_felts();
            draw_queries_felts.push_back(ab[0]);
            draw_queries_felts.push_back(ab[1]);
        }
        for (const QM31Var &felt : draw_queries_felts) {
            const std::array<M31Var, 4> d = felt.decompose_m31();
            r.raw_queries.insert(r.raw_queries.end(), d.begin(), d.end());
        }
        r.raw_queries.resize(f.s.n_queries);
        QM31Var input_sum = QM31Var::zero(cs);
        for (const auto &in : inputs) {
            const QM31Var idx = QM31Var::new_constant(cs, {{in.first, 0, 0, 0}});
            const QM31Var sum = (in.second + (idx * r.lookup_elements.alpha)) - r.lookup_elements.z;
            input_sum = input_sum + sum.inv();
        }
        ((input_sum + proof.poseidon_total_sum) + proof.plonk_total_sum).equalverify(QM31Var::zero(cs));
        return r;
    }
};

// ---- composition (OODS) --------------------------------------------------------------------------------------------
struct PointEvaluationAccumulatorVar {    // composition/src/data_structures.rs:13-33
    QM31Var random_coeff, accumulation;
    void accumulate(const QM31Var &evaluation) { accumulation = (accumulation * random_coeff) + evaluation; }
};

struct EvalAtRowVar {                     // composition/src/data_structures.rs:82-215
    u32 col_index[4] = {0, 0, 0, 0};
    std::vector<std::vector<const std::vector<QM31Var> *>> mask;     // [tree][column] -> mask values
    QM31Var cumsum_shift, denom_inverse;
    std::vector<std::pair<QM31Var, QM31Var>> fracs;
    PointEvaluationAccumulatorVar *acc;
    const LookupElementsVar *relation;

    EvalAtRowVar(std::vector<std::vector<const std::vector<QM31Var> *>> mask_, const QM31Var &total_sum, const QM31Var &denom_inverse_,
                 u32 log_size, PointEvaluationAccumulatorVar *acc_, const LookupElementsVar *rel)
        : mask(std::move(mask_)), denom_inverse(denom_inverse_), acc(acc_), relation(rel) {
        cumsum_shift = total_sum.mul_constant_m31(m31_inverse(1u << log_size));           // LogupAtRowVar::new (:57-68)
    }
    const std::vector<QM31Var> &next_interaction_mask(u32 interaction) { return *mask[interaction][col_index[interaction]++]; }
    QM31Var next_trace_mask() { return next_interaction_mask(1)[0]; }
    QM31Var get_preprocessed_column() { return next_interaction_mask(0)[0]; }
    static QM31Var combine_ef(const QM31Var &v0, const QM31Var &v1, const QM31Var &v2, const QM31Var &v3) {   // :143-146
        const QM31Var s1 = v1.shift_by_i();
        const QM31Var a = v0 + s1;
        const QM31Var s2 = v2.shift_by_j();
        const QM31Var b = a + s2;
        const QM31Var s3 = v3.shift_by_ij();
        return b + s3;
    }
    std::vector<QM31Var> next_extension_interaction_mask(u32 interaction, u32 n) {        // :129-141
        const std::vector<QM31Var> *cols[4];
        for (u32 k = 0; k < 4; k++) cols[k] = &next_interaction_mask(interaction);
        std::vector<QM31Var> out;
        for (u32 k = 0; k < n; k++) out.push_back(combine_ef((*cols[0])[k], (*cols[1])[k], (*cols[2])[k], (*cols[3])[k]));
        return out;
    }
    void add_to_relation(const QM31Var &multiplicity, const std::vector<QM31Var> &values) {   // :148-165
        QM31Var denom = relation->alpha_powers[0] * values[0];
        for (size_t k = 1; k < values.size() && k < 3; k++) {
            const QM31Var t = relation->alpha_powers[k] * values[k];
            denom = denom + t;
        }
        denom = denom - relation->z;
        fracs.push_back({multiplicity, denom});
    }
    void add_constraint(const QM31Var &value) {                                           // :167-170
        const QM31Var ev = value * denom_inverse;
        acc->accumulate(ev);
    }
    void finalize_logup(u32 batch_size) {                                                 // :172-210
        const ConstraintSystemRef &cs = denom_inverse.cs;
        const u32 num_batches = ((u32)fracs.size() + batch_size - 1) / batch_size;
        std::vector<std::pair<QM31Var, QM31Var>> batched;
        for (size_t k = 0; k < fracs.size(); k += batch_size) {
            const size_t end = std::min(fracs.size(), k + batch_size);
            QM31Var p = fracs[k].first, q = fracs[k].second;
            for (size_t e = k + 1; e < end; e++) {
                const QM31Var pe = p * fracs[e].second;
                const QM31Var eq = fracs[e].first * q;
                p = pe + eq;
                q = q * fracs[e].second;
            }
            batched.push_back({p, q});
        }
        QM31Var prev_col_cumsum = QM31Var::zero(cs);
        for (u32 k = 0; k + 1 < num_batches; k++) {
            const QM31Var cur_cumsum = next_extension_interaction_mask(2, 1)[0];
            const QM31Var diff = cur_cumsum - prev_col_cumsum;
            prev_col_cumsum = cur_cumsum;
            const QM31Var dd = diff * batched[k].second;
            add_constraint(dd - batched[k].first);
        }
        for (u32 k = num_batches - 1; k < num_batches; k++) {
            const std::vector<QM31Var> pc = next_extension_interaction_mask(2, 2);        // masks [-1, 0]
            const QM31Var d1 = pc[1] - pc[0];
            const QM31Var diff = d1 - prev_col_cumsum;
            const QM31Var fixed_diff = diff + cumsum_shift;
            const QM31Var dd = fixed_diff * batched[k].second;
            add_constraint(dd - batched[k].first);
        }
    }
};

struct CompositionCheck {
    static QM31Var coset_vanishing(const CirclePointQM31Var &p, u32 coset_log_size) {    // composition/src/lib.rs:18-29
        // -coset.initial + coset.step_size.half() is the identity for a canonic coset: the reference still adds it
        QM31Var x = (p + CirclePointM31{1, 0}).x;
        for (u32 k = 1; k < coset_log_size; k++) {
            const QM31Var sq = x * x;
            x = (sq + sq) - M31Var::one(x.cs);
        }
        return x;
    }
    static void evaluate_plonk(const ConstraintSystemRef &cs, EvalAtRowVar &eval) {       // composition/src/plonk.rs:8-82
        const QM31Var a_wire = eval.get_preprocessed_column(), b_wire = eval.get_preprocessed_column(), c_wire = eval.get_preprocessed_column();
        const QM31Var op = eval.get_preprocessed_column();
        const QM31Var mult_a = eval.get_preprocessed_column(), mult_b = eval.get_preprocessed_column(), mult_c = eval.get_preprocessed_column();
        const QM31Var poseidon_wire = eval.get_preprocessed_column(), mult_poseidon = eval.get_preprocessed_column();
        const QM31Var enforce_c_m31 = eval.get_preprocessed_column();
        QM31Var av[4], bv[4], cv[4];
        for (auto &v : av) v = eval.next_trace_mask();
        for (auto &v : bv) v = eval.next_trace_mask();
        for (auto &v : cv) v = eval.next_trace_mask();
        for (u32 k = 1; k < 4; k++) {
            const QM31Var e = enforce_c_m31 * cv[k];
            eval.add_constraint(e);
        }
        const QM31Var a_val = EvalAtRowVar::combine_ef(av[0], av[1], av[2], av[3]);
        const QM31Var b_val = EvalAtRowVar::combine_ef(bv[0], bv[1], bv[2], bv[3]);
        const QM31Var c_val = EvalAtRowVar::combine_ef(cv[0], cv[1], cv[2], cv[3]);
        {
            const QM31Var s = a_val + b_val;
            const QM31Var os = op * s;
            const QM31Var lhs = c_val - os;
            const QM31Var om = QM31Var::one(cs) - op;
            const QM31Var oa = om * a_val;
            const QM31Var oab = oa * b_val;
            eval.add_constraint(lhs - oab);
        }
        eval.add_to_relation(mult_a, {a_val, a_wire});
        eval.add_to_relation(mult_b, {b_val, b_wire});
        eval.add_to_relation(mult_c, {c_val, c_wire});
        const QM31Var neg_mp = -mult_poseidon;
        eval.add_to_relation(neg_mp, {poseidon_wire, a_val, b_val});
        eval.finalize_logup(2);
    }
    // ---- composition/src/poseidon.rs:10-241
    static void apply_m4(QM31Var *x) {
        const QM31Var t0 = x[0] + x[1];
        const QM31Var t02 = t0 + t0;
        const QM31Var t1 = x[2] + x[3];
        const QM31Var t12 = t1 + t1;
        const QM31Var x11 = x[1] + x[1];
        const QM31Var t2 = x11 + t1;
        const QM31Var x33 = x[3] + x[3];
        const QM31Var t3 = x33 + t0;
        const QM31Var t1212 = t12 + t12;
        const QM31Var t4 = t1212 + t3;
        const QM31Var t0202 = t02 + t02;
        const QM31Var t5 = t0202 + t2;
        const QM31Var t6 = t3 + t5;
        const QM31Var t7 = t2 + t4;
        x[0] = t6; x[1] = t5; x[2] = t7; x[3] = t4;
    }
    static void apply_external_round_matrix(QM31Var *s) {
        for (u32 i = 0; i < 4; i++) apply_m4(s + 4 * i);
        for (u32 j = 0; j < 4; j++) {
            const QM31Var s01 = s[j] + s[j + 4];
            const QM31Var s012 = s01 + s[j + 8];
            const QM31Var t = s012 + s[j + 12];
            for (u32 i = 0; i < 4; i++) s[4 * i + j] = s[4 * i + j] + t;
        }
    }
    static void apply_internal_round_matrix(QM31Var *s) {
        QM31Var sum = s[0];
        for (u32 i = 1; i < 16; i++) sum = sum + s[i];
        const QM31Var d = s[0] + s[0];
        const QM31Var ds = d + sum;
        s[0] = s[0] + ds;
        for (u32 i = 1; i < 16; i++) {
            const QM31Var m = s[i].mul_constant_m31(1u << (i + 1));
            s[i] = m + sum;
        }
    }
    static QM31Var pow5(const QM31Var &x) {
        const QM31Var x2 = x * x;
        const QM31Var x4 = x2 * x2;
        return x4 * x;
    }
    static void evaluate_poseidon(const ConstraintSystemRef &cs, EvalAtRowVar &eval) {
        const QM31Var is_first_round = eval.get_preprocessed_column(), is_last_round = eval.get_preprocessed_column();
        const QM31Var is_full_round = eval.get_preprocessed_column();
        const QM31Var one = QM31Var::one(cs);
        const QM31Var is_not_first_round = one - is_first_round;
        const QM31Var is_not_last_round = one - is_last_round;
        const QM31Var is_partial_round = is_not_first_round - is_full_round;
        const QM31Var round_id = eval.get_preprocessed_column();
        QM31Var rc0[16], rc1[16];
        for (auto &v : rc0) v = eval.get_preprocessed_column();
        for (auto &v : rc1) v = eval.get_preprocessed_column();
        const QM31Var external_idx_1 = eval.get_preprocessed_column(), external_idx_2 = eval.get_preprocessed_column();
        const QM31Var is_external_idx_1_nonzero = eval.get_preprocessed_column(), is_external_idx_2_nonzero = eval.get_preprocessed_column();
        const QM31Var swap_bit_addr = rc0[0];
        QM31Var in_state[16], intermediate_state[16], out_state[16];
        for (auto &v : in_state) v = eval.next_trace_mask();
        for (auto &v : intermediate_state) v = eval.next_trace_mask();
        for (auto &v : out_state) v = eval.next_trace_mask();
        const QM31Var swap_bit_value = intermediate_state[0];
        const QM31Var one_minus_swap_bit_value = one - swap_bit_value;
        QM31Var permuted_state[16];
        for (u32 i = 0; i < 16; i++) {
            if (i < 8) {
                const QM31Var l = in_state[i] * one_minus_swap_bit_value;
                const QM31Var r = in_state[i + 8] * swap_bit_value;
                permuted_state[i] = l + r;
            } else {
                const QM31Var l = in_state[i - 8] * swap_bit_value;
                const QM31Var r = in_state[i] * one_minus_swap_bit_value;
                permuted_state[i] = l + r;
            }
        }
        apply_external_round_matrix(permuted_state);
        for (u32 i = 0; i < 16; i++) {
            const QM31Var d = permuted_state[i] - out_state[i];
            eval.add_constraint(is_first_round * d);
        }
        QM31Var full_round_state[16];
        for (u32 i = 0; i < 16; i++) full_round_state[i] = in_state[i] + rc0[i];
        for (u32 i = 0; i < 16; i++) full_round_state[i] = pow5(full_round_state[i]);
        for (u32 i = 0; i < 16; i++) {
            const QM31Var d = intermediate_state[i] - full_round_state[i];
            eval.add_constraint(is_full_round * d);
            full_round_state[i] = intermediate_state[i];
        }
        apply_external_round_matrix(full_round_state);
        for (u32 i = 0; i < 16; i++) full_round_state[i] = full_round_state[i] + rc1[i];
        for (u32 i = 0; i < 16; i++) full_round_state[i] = pow5(full_round_state[i]);
        apply_external_round_matrix(full_round_state);
        for (u32 i = 0; i < 16; i++) {
            const QM31Var d = out_state[i] - full_round_state[i];
            eval.add_constraint(is_full_round * d);
        }
        QM31Var partial_round_state[16];
        for (u32 i = 0; i < 16; i++) partial_round_state[i] = in_state[i];
        for (u32 r = 0; r < 14; r++) {
            partial_round_state[0] = partial_round_state[0] + rc0[r];
            partial_round_state[0] = pow5(partial_round_state[0]);
            const QM31Var d = intermediate_state[r] - partial_round_state[0];
            eval.add_constraint(is_partial_round * d);
            partial_round_state[0] = intermediate_state[r];
            apply_internal_round_matrix(partial_round_state);
        }
        for (u32 i = 0; i < 16; i++) {
            const QM31Var d = out_state[i] - partial_round_state[i];
            eval.add_constraint(is_partial_round * d);
        }
        const QM31Var in_left_id = round_id + round_id;
        const QM31Var in_right_id = in_left_id + one;
        const QM31Var out_left_id = in_right_id + one;
        const QM31Var out_right_id = out_left_id + one;
        auto entry = [&](const QM31Var &nonzero, const QM31Var &is_edge, const QM31Var &ext, const QM31Var &not_edge, const QM31Var &inner_id,
                         const QM31Var *st, bool subtract) {
            const QM31Var sel = nonzero * is_edge;
            const QM31Var e1 = is_edge * ext;
            const QM31Var e2 = not_edge * inner_id;
            const QM31Var id = e1 + e2;
            const QM31Var a = EvalAtRowVar::combine_ef(st[0], st[1], st[2], st[3]);
            const QM31Var b = EvalAtRowVar::combine_ef(st[4], st[5], st[6], st[7]);
            const QM31Var mult = subtract ? sel - not_edge : sel + not_edge;
            eval.add_to_relation(mult, {id, a, b});
        };
        entry(is_external_idx_1_nonzero, is_first_round, external_idx_1, is_not_first_round, in_left_id, in_state, true);
        entry(is_external_idx_2_nonzero, is_first_round, external_idx_2, is_not_first_round, in_right_id, in_state + 8, true);
        entry(is_external_idx_1_nonzero, is_last_round, external_idx_1, is_not_last_round, out_left_id, out_state, false);
        entry(is_external_idx_2_nonzero, is_last_round, external_idx_2, is_not_last_round, out_right_id, out_state + 8, false);
        const QM31Var fl = is_first_round * is_not_last_round;
        eval.add_to_relation(fl, {swap_bit_value, swap_bit_addr});
        eval.finalize_logup(3);
    }
    static void compute(const ShapeFacts &f, const LookupElementsVar &lookup_elements, const QM31Var &random_coeff,
                        const CirclePointQM31Var &oods_point, const PlonkWithPoseidonProofVar &proof) {   // composition/src/lib.rs:33-121
        const ConstraintSystemRef &cs = random_coeff.cs;
        PointEvaluationAccumulatorVar acc{random_coeff, QM31Var::zero(cs)};
        const auto &sv = proof.sampled_values;
        for (u32 comp = 0; comp < 2; comp++) {
            std::vector<std::vector<const std::vector<QM31Var> *>> mask(3);
            for (u32 t = 0; t < 3; t++) {
                const u32 lo = comp == 0 ? 0 : ShapeFacts::plonk_cols(t), hi = comp == 0 ? ShapeFacts::plonk_cols(t) : ShapeFacts::n_cols(t);
                for (u32 c = lo; c < hi; c++) mask[t].push_back(&sv[t][c]);
            }
            const u32 log_size = comp == 0 ? f.s.log_size_plonk : f.s.log_size_poseidon;
            const QM31Var denom_inverse = coset_vanishing(oods_point, log_size).inv();
            EvalAtRowVar eval(mask, comp == 0 ? proof.plonk_total_sum : proof.poseidon_total_sum, denom_inverse, log_size, &acc, &lookup_elements);
            if (comp == 0) evaluate_plonk(cs, eval); else evaluate_poseidon(cs, eval);
        }
        const QM31Var left_value = EvalAtRowVar::combine_ef(sv[3][0][0], sv[3][1][0], sv[3][2][0], sv[3][3][0]);
        const QM31Var right_value = EvalAtRowVar::combine_ef(sv[3][4][0], sv[3][5][0], sv[3][6][0], sv[3][7][0]);
        const QM31Var xd = oods_point.repeated_double_x_only(f.composition_log_degree_bound - 2);
        const QM31Var rx = right_value * xd;
        const QM31Var expected = left_value + rx;
        acc.accumulation.equalverify(expected);
    }
};

// ---- Merkle proofs (data_structures/src/lib.rs:264-464) ------------------------------------------------------------------
struct SinglePathMerkleProofVar {
    u32 depth;
    std::vector<HashVar> sibling_hashes;
    std::map<u32, std::vector<M31Var>> columns;

    // columns per layer, ascending layer order (BTreeMap iteration): log size -> (offset in cols_of, count); the hint holds
    // the layers in descending order
    static std::map<u32, std::pair<u32, u32>> layer_layout(const ShapeFacts &f, u32 t) {
        std::map<u32, std::pair<u32, u32>> layers;
        if (t == 3) layers[f.max_first] = {0, 8};
        else {
            const u32 np = ShapeFacts::plonk_cols(t), ns = ShapeFacts::n_cols(t) - np;
            if (f.log_plonk == f.log_pos) layers[f.log_plonk] = {0, np + ns};
            else if (f.log_plonk > f.log_pos) { layers[f.log_plonk] = {0, np}; layers[f.log_pos] = {np, ns}; }
            else { layers[f.log_pos] = {0, ns}; layers[f.log_plonk] = {ns, np}; }
        }
        return layers;
    }
    // hint layout of the batched verifier: cols_of(p, t, i) = leaf layer first then the smaller layer; sib_of = 8 words per level
    static SinglePathMerkleProofVar new_(WitnessStream &w, const ShapeFacts &f, u32 t, u32 i) {
        const ConstraintSystemRef &cs = w.cs;
        SinglePathMerkleProofVar v;
        v.depth = f.tree_depth(t);
        for (u32 k = 0; k < v.depth; k++) v.sibling_hashes.push_back(HashVar::new_single_use_witness_only(cs, w.take(tape::S_PATH_SIB, t, i, 8 * k, 8)));
        for (const auto &l : layer_layout(f, t)) {
            std::vector<M31Var> col;
            for (u32 k = 0; k < l.second.second; k++) col.push_back(M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_PATH_COL, t, i, l.second.first + k, 1))));
            v.columns[l.first] = col;
        }
        return v;
    }
    // hint_base: slot of this (tree, query) path in the native verifier's permutation record (verify.cuh HintLayout), whose
    // order is merkle::path_root's: leaf sponge, leaf final, then per level node [, column sponge, combine].  The circuit
    // hashes a level's columns BEFORE the node, so the slots are queued in the circuit's order.
    void verify(const HashVar &root, const BitsVar &query, u32 hint_base = tape::NO_VAR) const {     // :315-354
        if (hint_base != tape::NO_VAR) {
            const ConstraintSystemRef &cs = root.cs;
            u32 at = hint_base;
            const u32 leaf_chunks = ((u32)columns.at(depth).size() + 7) / 8;
            for (u32 k = 0; k <= leaf_chunks; k++) cs->push_hint(at++);
            for (u32 i = 0; i < depth; i++) {
                auto it = columns.find(depth - i - 1);
                if (it != columns.end()) {
                    const u32 chunks = ((u32)it->second.size() + 7) / 8;
                    for (u32 k = 0; k < chunks; k++) cs->push_hint(at + 1 + k);
                    cs->push_hint(at);
                    cs->push_hint(at + 1 + chunks);
                    at += chunks + 2;
                } else cs->push_hint(at++);
            }
        }
        HashVar cur_hash = Poseidon31MerkleHasherVar::hash_m31_columns_get_rate(columns.at(depth));
        for (u32 i = 0; i < depth; i++) {
            const u32 h = depth - i - 1;
            auto it = columns.find(h);
            if (it != columns.end()) {
                const HashVar column_hash = Poseidon31MerkleHasherVar::hash_m31_columns_get_capacity(it->second);
                cur_hash = Poseidon31MerkleHasherVar::hash_tree_with_column_hash_with_swap(cur_hash, sibling_hashes[i], query.variables[i], column_hash);
            } else {
                cur_hash = Poseidon31MerkleHasherVar::hash_tree_with_swap(cur_hash, sibling_hashes[i], query.variables[i]);
            }
        }
        cur_hash.equalverify(root);
    }
};

struct SinglePairMerkleProofVar {
    u32 depth;
    std::vector<HashVar> sibling_hashes;
    std::map<u32, QM31Var> self_columns, siblings_columns;

    // fri tree f (0 = first layer, 1.. = inner); data layers in descending order index the hint's self/sibling values
    static SinglePairMerkleProofVar new_(WitnessStream &w, const ShapeFacts &f, u32 tree, u32 i) {
        const ConstraintSystemRef &cs = w.cs;
        SinglePairMerkleProofVar v;
        v.depth = tree == 0 ? f.max_first : f.max_first - tree;
        for (u32 k = 0; k + 1 < v.depth; k++) v.sibling_hashes.push_back(HashVar::new_single_use_witness_only(cs, w.take(tape::S_PAIR_HASH, tree, i, 8 * k, 8)));
        std::vector<u32> data_layers;                                 // descending
        for (u32 h = v.depth + 1; h-- > 0;)
            if (tree == 0 ? f.fri_first_has_data(h) : h == v.depth) data_layers.push_back(h);
        for (int pass = 0; pass < 2; pass++)                          // self_columns first, then siblings_columns; ascending layers
            for (size_t d = data_layers.size(); d-- > 0;) {
                const QM31Var q = QM31Var::new_witness(cs, Def::input_qm31(w.take(pass == 0 ? tape::S_PAIR_SELF : tape::S_PAIR_SIB, tree, i, 4 * (u32)d, 4)));
                (pass == 0 ? v.self_columns : v.siblings_columns)[data_layers[d]] = q;
            }
        return v;
    }
    // hint_base: as above; native order (decommit::pair_path_root): self leaf (2), sibling leaf (2), then per level either
    // node, or node, self sponge, self combine, sibling sponge, sibling combine
    void verify(const HashVar &root, const BitsVar &query, u32 hint_base = tape::NO_VAR) const {     // :400-464
        const ConstraintSystemRef &cs = root.cs;
        if (hint_base != tape::NO_VAR) {
            u32 at = hint_base;
            for (u32 k = 0; k < 4; k++) cs->push_hint(at++);
            for (u32 i = 0; i < depth; i++) {
                if (!self_columns.count(depth - i - 1)) cs->push_hint(at++);
                else {
                    const u32 order[5] = {1, 3, 0, 2, 4};
                    for (u32 k = 0; k < 5; k++) cs->push_hint(at + order[k]);
                    at += 5;
                }
            }
        }
        HashVar self_hash = Poseidon31MerkleHasherVar::hash_qm31_columns_get_rate({self_columns.at(depth), QM31Var::zero(cs)});
        HashVar sibling_hash = Poseidon31MerkleHasherVar::hash_qm31_columns_get_rate({siblings_columns.at(depth), QM31Var::zero(cs)});
        for (u32 i = 0; i < depth; i++) {
            const u32 h = depth - i - 1;
            if (!self_columns.count(h)) {
                self_hash = Poseidon31MerkleHasherVar::hash_tree_with_swap(self_hash, sibling_hash, query.variables[i]);
                if (i != depth - 1) sibling_hash = sibling_hashes[i];
            } else {
                const HashVar self_column_hash = Poseidon31MerkleHasherVar::hash_qm31_columns_get_capacity({self_columns.at(h), QM31Var::zero(cs)});
                const HashVar sibling_column_hash = Poseidon31MerkleHasherVar::hash_qm31_columns_get_capacity({siblings_columns.at(h), QM31Var::zero(cs)});
                self_hash = Poseidon31MerkleHasherVar::hash_tree_with_column_hash_with_swap(self_hash, sibling_hash, query.variables[i], self_column_hash);
                sibling_hash = Poseidon31MerkleHasherVar::combine_hash_tree_with_column(sibling_hashes[i], sibling_column_hash);
            }
        }
        self_hash.equalverify(root);
    }
};

// ---- answers (components/recursive/answer/src) ---------------------------------------------------------------------------
struct PointSampleVar { long shift_key; const CirclePointQM31Var *point; QM31Var value; };   // shift_key: 0 = ShiftIndex::Zero
struct ColumnSampleBatchVar { const CirclePointQM31Var *point; std::vector<std::pair<u32, QM31Var>> columns_and_values; };

struct AnswerResults {
    std::unique_ptr<QueryPositionsPerLogSizeVar> query_positions_per_log_size;
    std::map<u32, std::vector<QM31Var>> fri_answers;                  // by log size
    std::map<u32, std::vector<CirclePointM31Var>> domain_points;

    static std::array<CM31Var, 2> dec(const QM31Var &q) { return q.decompose_cm31(); }

    using SampledValues = std::vector<std::vector<std::vector<QM31Var>>>;
    using Columns = std::map<u32, std::vector<M31Var>>;                                  // layer log size -> opened column values
    // decommit(query positions) -> [tree][query] columns: DecommitmentVar::new + verify in the recursive circuit,
    // LastDecommitVar::compute in the last-layer circuit
    using DecommitFn = std::function<std::vector<std::vector<Columns>>(const QueryPositionsPerLogSizeVar &)>;

    static AnswerResults compute(WitnessStream &w, const CirclePointQM31Var &oods_point, const ShapeFacts &f, const FiatShamirResults &fs,
                                 const PlonkWithPoseidonProofVar &proof) {               // answer/src/lib.rs:34-354
        const u32 nq = f.s.n_queries;
        return compute_with(w, oods_point, f, fs.raw_queries, fs.after_sampled_values_random_coeff, proof.sampled_values,
                            f.s.log_last + f.s.log_blowup + 1, [&](const QueryPositionsPerLogSizeVar &qp) {
            // DecommitmentVar::new, then the four verify loops (:201-247)
            std::vector<std::vector<SinglePathMerkleProofVar>> dec_proofs(4);
            for (u32 t = 0; t < 4; t++)
                for (u32 i = 0; i < nq; i++) dec_proofs[t].push_back(SinglePathMerkleProofVar::new_(w, f, t, i));
            std::vector<std::vector<Columns>> cols(4);
            for (u32 t = 0; t < 4; t++)
                for (u32 i = 0; i < nq; i++) {
                    dec_proofs[t][i].verify(proof.commitments[t], qp[f.tree_depth(t)][i].bits, f.hints.single_slot(t, i));
                    cols[t].push_back(dec_proofs[t][i].columns);
                }
            return cols;
        });
    }
    static AnswerResults compute_with(WitnessStream &w, const CirclePointQM31Var &oods_point, const ShapeFacts &f, const std::vector<M31Var> &raw_queries,
                                      const QM31Var &after_sampled_values_random_coeff, const SampledValues &sampled_values, u32 min_degree,
                                      const DecommitFn &decommit) {
        const ConstraintSystemRef &cs = w.cs;
        const u32 nq = f.s.n_queries;
        // The reference iterates a HashSet of mask shifts here, i.e. in a per-process random order; this build fixes
        // first-appearance order (0, then -1), the Plonk component before the Poseidon component.
        CirclePointQM31Var shifted[2][2];
        for (u32 comp = 0; comp < 2; comp++) {
            const CirclePointM31 step = cp_subgroup_gen(comp == 0 ? f.s.log_size_plonk : f.s.log_size_poseidon);   // CanonicCoset::step
            shifted[comp][0] = oods_point + CirclePointM31{1, 0};                         // mul_signed(0)
            shifted[comp][1] = oods_point + cp_neg(step);                                 // mul_signed(-1)
        }
        // samples in flatten order, with the (blown-up) log size of their column
        std::vector<std::pair<u32, std::vector<PointSampleVar>>> samples;
        for (u32 t = 0; t < 4; t++)
            for (u32 c = 0; c < ShapeFacts::n_cols(t); c++) {
                std::vector<PointSampleVar> e;
                const auto &col = sampled_values[t][c];
                if (t == 0 || t == 3) e.push_back({0, &oods_point, col[0]});              // mask_points[PREPROCESSED] / composition
                else {
                    const u32 comp = c < ShapeFacts::plonk_cols(t) ? 0 : 1;
                    const u32 comp_log = comp == 0 ? f.s.log_size_plonk : f.s.log_size_poseidon;
                    if (col.size() == 1) e.push_back({0, &shifted[comp][0], col[0]});
                    else {                                                                // offsets [-1, 0]; ShiftIndex::Shift(-1, log_size)
                        e.push_back({-(long)(comp_log + 1), &shifted[comp][1], col[0]});
                        e.push_back({0, &shifted[comp][0], col[1]});
                    }
                }
                samples.push_back({f.column_log_size(t, c), e});
            }
        AnswerResults r;
        r.query_positions_per_log_size.reset(new QueryPositionsPerLogSizeVar(min_degree, f.max_first, raw_queries));
        const QueryPositionsPerLogSizeVar &qp = *r.query_positions_per_log_size;
        const std::vector<std::vector<Columns>> dec_cols = decommit(qp);
        std::vector<u32> desc(f.all_log_sizes.rbegin(), f.all_log_sizes.rend());
        for (u32 L : desc) {
            // queried values of this size: the four trees in order (:249-283)
            std::vector<std::vector<M31Var>> queried(nq);
            for (u32 i = 0; i < nq; i++)
                for (u32 t = 0; t < 4; t++) {
                    auto it = dec_cols[t][i].find(L);
                    if (it != dec_cols[t][i].end()) queried[i].insert(queried[i].end(), it->second.begin(), it->second.end());
                }
            // ColumnSampleBatchVar::new_vec: group by shift in first-seen order (answer/src/data_structures.rs:42-64)
            std::vector<long> order;
            std::map<long, ColumnSampleBatchVar> groups;
            u32 column_index = 0;
            for (const auto &s : samples) {
                if (s.first != L) continue;
                for (const PointSampleVar &ps : s.second) {
                    if (!groups.count(ps.shift_key)) { groups[ps.shift_key].point = ps.point; order.push_back(ps.shift_key); }
                    groups[ps.shift_key].columns_and_values.push_back({column_index, ps.value});
                }
                column_index++;
            }
            std::vector<const ColumnSampleBatchVar *> batches;
            for (long k : order) batches.push_back(&groups[k]);
            // column_line_coeffs_var (:162-189)
            QM31Var alpha = QM31Var::new_constant(cs, {{0, 0, P - 2, 0}});
            std::vector<std::vector<std::array<QM31Var, 3>>> line_coeffs;
            for (const ColumnSampleBatchVar *b : batches) {
                std::vector<std::array<QM31Var, 3>> lc;
                for (const auto &cv : b->columns_and_values) {
                    // complex_conjugate_line_coeffs_var (:137-160)
                    const std::array<CM31Var, 2> value = dec(cv.second);
                    const std::array<CM31Var, 2> y = dec(b->point->y);
                    const CM31Var v0y1 = value[0] * y[1];
                    const CM31Var v1y0 = value[1] * y[0];
                    const CM31Var bb = v0y1 - v1y0;
                    const QM31Var ca = alpha * value[1];
                    const QM31Var cb = alpha * bb;
                    const QM31Var cc = alpha * y[1];
                    lc.push_back({ca, cb, cc});
                    alpha = alpha * after_sampled_values_random_coeff;
                }
                line_coeffs.push_back(lc);
            }
            for (u32 i = 0; i < nq; i++) {
                const CirclePointM31Var domain_point = qp[L][i].get_next_point();
                // denominator_inverses_var (:103-126)
                std::vector<CM31Var> denominator_inverses;
                for (const ColumnSampleBatchVar *b : batches) {
                    const std::array<CM31Var, 2> px = dec(b->point->x);
                    const std::array<CM31Var, 2> py = dec(b->point->y);
                    CM31Var a = px[0] - domain_point.x;
                    a = a * py[1];
                    CM31Var bq = py[0] - domain_point.y;
                    bq = bq * px[1];
                    denominator_inverses.push_back((a - bq).inv());
                }
                // accumulate_row_quotients_var (:70-101)
                QM31Var row_accumulator = QM31Var::zero(cs);
                for (size_t bi = 0; bi < batches.size(); bi++) {
                    QM31Var numerator = QM31Var::zero(cs);
                    for (size_t k = 0; k < batches[bi]->columns_and_values.size(); k++) {
                        const auto &abc = line_coeffs[bi][k];
                        const QM31Var value = queried[i][batches[bi]->columns_and_values[k].first] * abc[2];
                        const QM31Var ay = abc[0] * domain_point.y;
                        const QM31Var linear_term = ay + abc[1];
                        const QM31Var d = value - linear_term;
                        numerator = numerator + d;
                    }
                    const QM31Var nd = numerator * denominator_inverses[bi];
                    row_accumulator = row_accumulator + nd;
                }
                r.fri_answers[L].push_back(row_accumulator);
                r.domain_points[L].push_back(domain_point);
            }
        }
        return r;
    }
};

// ---- folding (components/recursive/folding/src/lib.rs:12-205) --------------------------------------------------------------
struct FoldingResults {
    static void compute(WitnessStream &w, const PlonkWithPoseidonProofVar &proof, const ShapeFacts &f, const FiatShamirResults &fs,
                        const AnswerResults &ans) {
        const ConstraintSystemRef &cs = w.cs;
        const u32 nq = f.s.n_queries;
        const QueryPositionsPerLogSizeVar &qp = *ans.query_positions_per_log_size;
        std::vector<SinglePairMerkleProofVar> proofs;
        for (u32 i = 0; i < nq; i++) {
            proofs.push_back(SinglePairMerkleProofVar::new_(w, f, 0, i));
            proofs.back().verify(proof.first_layer_commitment, qp[f.max_first][i].bits, f.hints.pair_slot(0, i));
        }
        for (auto it = f.all_log_sizes.rbegin(); it != f.all_log_sizes.rend(); ++it)
            for (u32 i = 0; i < nq; i++) proofs[i].self_columns.at(*it).equalverify(ans.fri_answers.at(*it)[i]);
        std::map<u32, std::vector<QM31Var>> folded_results;
        for (u32 L : f.all_log_sizes)
            for (u32 i = 0; i < nq; i++) {
                const PointCarryingQueryVar &query = qp[L][i];
                const QM31Var &self_val = proofs[i].self_columns.at(L), &sibling_val = proofs[i].siblings_columns.at(L);
                const CirclePointM31Var point = query.get_absolute_point().double_();
                const M31Var y_inv = point.y.inv();
                const std::pair<QM31Var, QM31Var> lr = QM31Var::swap(self_val, sibling_val, query.bits.variables[0]);
                const QM31Var new_left_val = lr.first + lr.second;
                const QM31Var diff = lr.first - lr.second;
                const QM31Var new_right_val = diff * y_inv;
                const QM31Var ra = new_right_val * fs.fri_alphas[f.max_first - L];
                folded_results[L].push_back(new_left_val + ra);
            }
        u32 log_size = f.max_first;
        std::vector<QM31Var> folded(nq, QM31Var::zero(cs));
        for (u32 i = 0; i < f.s.n_inner; i++) {
            auto fr = folded_results.find(log_size);
            if (fr != folded_results.end()) {
                QM31Var fri_alpha = fs.fri_alphas[i];
                fri_alpha = fri_alpha * fri_alpha;
                for (u32 k = 0; k < nq; k++) {
                    const QM31Var av = fri_alpha * folded[k];
                    folded[k] = av + fr->second[k];
                }
            }
            log_size -= 1;
            std::vector<QM31Var> new_folded;
            for (u32 k = 0; k < nq; k++) {
                const PointCarryingQueryVar &query = qp[log_size][k];
                const SinglePairMerkleProofVar merkle_proof = SinglePairMerkleProofVar::new_(w, f, 1 + i, k);
                const QM31Var &self_val = merkle_proof.self_columns.at(log_size), &sibling_val = merkle_proof.siblings_columns.at(log_size);
                folded[k].equalverify(self_val);
                const M31Var x_inv = query.get_absolute_point().x.inv();
                const std::pair<QM31Var, QM31Var> lr = QM31Var::swap(self_val, sibling_val, query.bits.variables[0]);
                const QM31Var new_left_val = lr.first + lr.second;
                const QM31Var diff = lr.first - lr.second;
                const QM31Var new_right_val = diff * x_inv;
                const QM31Var ra = new_right_val * fs.fri_alphas[i + 1];
                new_folded.push_back(new_left_val + ra);
                merkle_proof.verify(proof.inner_layer_commitments[i], query.bits, f.hints.pair_slot(1 + i, k));
            }
            folded = new_folded;
        }
        for (u32 k = 0; k < nq; k++) {
            if (proof.last_poly.coeffs.size() == 1) folded[k].equalverify(proof.last_poly.coeffs[0]);
            else {
                const M31Var x = qp[log_size][k].get_next_point_x();
                const QM31Var eval = proof.last_poly.eval_at_point(x);
                folded[k].equalverify(eval);
            }
        }
    }
};

// The whole program of examples/single-proof/src/main.rs:33-85 (multipliers = 1) / multi-proofs/src/main.rs:62-135.
struct VerifierCircuit {
    ConstraintSystemRef cs;
    std::vector<u32> gather;              // witness-stream sources, one per stream word (n_input_words)
    u32 words_per_instance = 0;           // the stream of instance k starts at k * words_per_instance
};
inline VerifierCircuit record_verifier_circuit(const ProofShape &shape, const std::vector<PublicInput> &inputs, u32 multipliers) {
    const ShapeFacts f(shape);
    WitnessStream w{ConstraintSystemRef::new_plonk_with_poseidon_ref(), {}};
    const ConstraintSystemRef &cs = w.cs;
    cs->native_hints = true;
    VerifierCircuit out;
    for (u32 m = 0; m < multipliers; m++) {
        PlonkWithPoseidonProofVar proof = PlonkWithPoseidonProofVar::new_witness(w, f);
        std::vector<std::pair<u32, QM31Var>> in;
        for (const PublicInput &pi : inputs) in.push_back({pi.idx, QM31Var::new_constant(cs, pi.value)});
        const FiatShamirResults fs = FiatShamirResults::compute(proof, f, in);
        CompositionCheck::compute(f, fs.lookup_elements, fs.random_coeff, fs.oods_point, proof);
        const CirclePointQM31Var oods_witness = CirclePointQM31Var::new_witness(cs, w.take(tape::S_OODS, 0, 0, 0, 8));
        const AnswerResults ans = AnswerResults::compute(w, oods_witness, f, fs, proof);
        FoldingResults::compute(w, proof, f, fs, ans);
        if (!cs->hint_queue.empty()) throw std::logic_error("native-hint slots left over: the circuit recorded fewer permutations than it announced");
        if (m == 0) out.words_per_instance = cs->n_input_words;
    }
    cs->pad();
    out.cs = cs;
    out.gather = w.gather;
    return out;
}

// The folding stage on its own (components/recursive/folding/src/lib.rs:12-205) with everything before it supplied as witnesses: the
// FRI commitments and the last-layer polynomial (from the instance), the folding alphas and the query positions (from the native
// transcript) and the first-layer answers (from the native answer stage, or from the instance itself when it is one of the synthetic
// FRI + Merkle instances of BASELINE configs[4] part i, synth.cuh).  Same rows as the folding part of the verifier circuit; the
// queries are allocated as M31 witnesses holding the position at the largest log size.
inline VerifierCircuit record_folding_circuit(const ProofShape &shape) {
    const ShapeFacts f(shape);
    WitnessStream w{ConstraintSystemRef::new_plonk_with_poseidon_ref(), {}};
    const ConstraintSystemRef &cs = w.cs;
    cs->native_hints = true;
    PlonkWithPoseidonProofVar proof;
    proof.first_layer_commitment = HashVar::new_witness(cs, w.take(tape::S_FRI_COMMITMENT, 0, 0, 0, 8));
    for (u32 l = 0; l < f.s.n_inner; l++) proof.inner_layer_commitments.push_back(HashVar::new_witness(cs, w.take(tape::S_FRI_COMMITMENT, 1 + l, 0, 0, 8)));
    proof.last_poly.cs = cs;
    for (u32 k = 0; k < (1u << f.s.log_last); k++) proof.last_poly.coeffs.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_LAST_COEFFS, 0, 0, 4 * k, 4))));
    FiatShamirResults fs;
    for (u32 l = 0; l <= f.s.n_inner; l++) fs.fri_alphas.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_FS, 0, 0, 20 + 4 * l, 4))));
    for (u32 i = 0; i < f.s.n_queries; i++) fs.raw_queries.push_back(M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_FS, 0, 0, tape::FS_QUERY_BASE + i, 1))));
    AnswerResults ans;
    ans.query_positions_per_log_size.reset(new QueryPositionsPerLogSizeVar(f.s.log_last + f.s.log_blowup + 1, f.max_first, fs.raw_queries));
    u32 g = 0;                                                        // the native stages index log sizes in descending order
    for (auto it = f.all_log_sizes.rbegin(); it != f.all_log_sizes.rend(); ++it, ++g)
        for (u32 i = 0; i < f.s.n_queries; i++)
            ans.fri_answers[*it].push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_ANSWER, g, i, 0, 4))));
    FoldingResults::compute(w, proof, f, fs, ans);
    if (!cs->hint_queue.empty()) throw std::logic_error("native-hint slots left over: the circuit recorded fewer permutations than it announced");
    VerifierCircuit out;
    out.words_per_instance = cs->n_input_words;
    cs->pad();
    out.cs = cs;
    out.gather = w.gather;
    return out;
}

}  // namespace dsl
}  // namespace stwo_b200
