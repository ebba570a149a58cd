// The last-layer verifier circuit (components/last/*) over the Plonk-without-Poseidon system, recorded once per shape.
//
//   LastPlonkWithPoseidonProofVar                    components/last/data_structures/src/lib.rs:13-82
//   LastFiatShamirInputVar, LastFiatShamirResults    components/last/fiat_shamir/src/lib.rs:86-216
//   LastDecommitInputVar, LastDecommitVar, LastSinglePathMerkleProof{Input,}Var
//                                                    components/last/answer/src/data_structures/{mod,merkle_proofs}.rs
//   LastAnswerResults::compute                       components/last/answer/src/lib.rs:30-262
//   LastFirstLayerInputVar, LastInnerLayersInputVar  components/last/folding/src/data_structures/merkle_proofs.rs:100-455
//   LastFoldingResults::compute                      components/last/folding/src/lib.rs:14-161
//   driver                                           examples/last-layer/src/main.rs:26-97
//
// The circuit never replays the channel: every Fiat-Shamir output, the hash of the sampled values, the (packed or hashed)
// opened columns and the FRI pair openings are PUBLIC INPUTS; the emulated Poseidon2 gadget recomputes the hashes.  Their
// values come from the batched native verifier's workspace (tape::S_FS, S_PATH_COL, S_PAIR_*) and from the small
// public-input hash kernel (tape::S_EXTRA, circuit.cuh).
#pragma once
#include "recursive_verifier.hpp"

namespace stwo_b200 {
namespace dsl {

struct ExtraHashJob { u32 kind, tree, query, col_off, n_cols, slot; };     // mirrors circuit::ExtraJob

struct LastLayerCircuit {
    ConstraintSystemRef cs;
    std::vector<u32> gather;
    std::vector<ExtraHashJob> jobs;
    u32 n_extra_words = 0, n_public_inputs = 0;
};

inline LastLayerCircuit record_last_layer_circuit(const ProofShape &shape) {
    const ShapeFacts f(shape);
    WitnessStream w{ConstraintSystemRef::new_plonk_without_poseidon_ref(), {}};
    const ConstraintSystemRef &cs = w.cs;
    const u32 nq = f.s.n_queries;
    LastLayerCircuit out;
    auto extra_hash = [&](u32 kind, u32 tree, u32 query, u32 col_off, u32 n_cols) {
        const u32 slot = out.n_extra_words;
        out.jobs.push_back({kind, tree, query, col_off, n_cols, slot});
        out.n_extra_words += 8;
        return slot;
    };
    auto qin = [&](u32 section, u32 a, u32 i, u32 k0) { return QM31Var::new_public_input(cs, Def::input_qm31(w.take(section, a, i, k0, 4))); };

    // ---- public inputs in the order of main.rs:62-69 ---------------------------------------------------------------------
    // LastFiatShamirInputVar (last/fiat_shamir/src/lib.rs:104-150)
    const QM31Var in_t = qin(tape::S_FS, 0, 0, 0);
    const HashVar in_sampled_values_hash = HashVar::new_public_input(cs, w.take(tape::S_EXTRA, 0, 0, extra_hash(0, 0, 0, 0, 0), 8));
    const QM31Var in_plonk_total_sum = qin(tape::S_STMT1, 0, 0, 0), in_poseidon_total_sum = qin(tape::S_STMT1, 0, 0, 4);
    const QM31Var in_z = qin(tape::S_FS, 0, 0, 4), in_alpha = qin(tape::S_FS, 0, 0, 8);
    const QM31Var in_random_coeff = qin(tape::S_FS, 0, 0, 12), in_after = qin(tape::S_FS, 0, 0, 16);
    std::vector<QM31Var> in_packed_queries;
    for (u32 k = 0; k < nq; k += 4) {
        const u32 slot = cs->new_input_words(4);
        for (u32 j = 0; j < 4; j++) w.gather.push_back(tape::src_pack(tape::S_FS, 0, 0, k + j < nq ? tape::FS_QUERY_BASE + k + j : tape::FS_ZERO));
        in_packed_queries.push_back(QM31Var::new_public_input(cs, Def::input_qm31(slot)));
    }
    std::vector<QM31Var> in_fri_alphas;
    for (u32 l = 0; l <= f.s.n_inner; l++) in_fri_alphas.push_back(qin(tape::S_FS, 0, 0, 20 + 4 * l));
    // LastDecommitInputVar: per tree, per query, per layer (ascending): the columns packed into one / two QM31, or their hash
    std::vector<std::vector<std::map<u32, std::vector<QM31Var>>>> in_packed(4);
    for (u32 t = 0; t < 4; t++)
        for (u32 i = 0; i < nq; i++) {
            std::map<u32, std::vector<QM31Var>> per_layer;
            for (const auto &l : SinglePathMerkleProofVar::layer_layout(f, t)) {
                const u32 off = l.second.first, n = l.second.second;
                std::vector<QM31Var> packed;
                if (n <= 8) {
                    for (u32 k = 0; k < n; k += 4) {
                        const u32 slot = cs->new_input_words(4);
                        for (u32 j = 0; j < 4; j++)
                            w.gather.push_back(k + j < n ? tape::src_pack(tape::S_PATH_COL, t, i, off + k + j) : tape::src_pack(tape::S_FS, 0, 0, tape::FS_ZERO));
                        packed.push_back(QM31Var::new_public_input(cs, Def::input_qm31(slot)));
                    }
                } else {
                    const u32 e = extra_hash(1, t, i, off, n);
                    packed.push_back(qin(tape::S_EXTRA, 0, 0, e));
                    packed.push_back(qin(tape::S_EXTRA, 0, 0, e + 4));
                }
                per_layer[l.first] = packed;
            }
            in_packed[t].push_back(per_layer);
        }
    // LastFirstLayerInputVar / LastInnerLayersInputVar: self columns then sibling columns, ascending layers; the inner
    // layers are a BTreeMap keyed by log size, i.e. allocated smallest layer first
    struct PairIn { std::map<u32, QM31Var> self_columns, siblings_columns; };
    auto pair_input = [&](u32 tree) {
        std::vector<u32> data_layers;                                 // descending = index into the verifier's pair hints
        const u32 depth = tree == 0 ? f.max_first : f.max_first - tree;
        for (u32 h = depth + 1; h-- > 0;)
            if (tree == 0 ? f.fri_first_has_data(h) : h == depth) data_layers.push_back(h);
        return [&, tree, data_layers](u32 i) {
            PairIn p;
            for (int pass = 0; pass < 2; pass++)
                for (size_t d = data_layers.size(); d-- > 0;)
                    (pass == 0 ? p.self_columns : p.siblings_columns)[data_layers[d]] = qin(pass == 0 ? tape::S_PAIR_SELF : tape::S_PAIR_SIB, tree, i, 4 * (u32)d);
            return p;
        };
    };
    std::vector<PairIn> in_first;
    {
        auto mk = pair_input(0);
        for (u32 i = 0; i < nq; i++) in_first.push_back(mk(i));
    }
    std::map<u32, std::vector<PairIn>> in_inner;
    for (u32 li = f.s.n_inner; li-- > 0;) {
        auto mk = pair_input(1 + li);
        std::vector<PairIn> v;
        for (u32 i = 0; i < nq; i++) v.push_back(mk(i));
        in_inner[f.max_first - 1 - li] = v;
    }
    out.n_public_inputs = cs->num_input;

    // ---- LastPlonkWithPoseidonProofVar::new_witness ------------------------------------------------------------------------
    const M31Var log_size_plonk = M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_STMT0, 0, 0, 0, 1)));
    const M31Var log_size_poseidon = M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_STMT0, 0, 0, 1, 1)));
    (void)log_size_plonk; (void)log_size_poseidon;
    QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_STMT1, 0, 0, 0, 4)));
    QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_STMT1, 0, 0, 4, 4)));
    AnswerResults::SampledValues sampled_values(4);
    for (u32 t = 0; t < 4; t++)
        for (u32 c = 0; c < ShapeFacts::n_cols(t); c++) {
            std::vector<QM31Var> col;
            for (u32 m = 0; m < ShapeFacts::n_masks(t, c); m++) col.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_SAMPLED, t, c, 4 * m, 4))));
            sampled_values[t].push_back(col);
        }
    LinePolyVar last_poly;
    last_poly.cs = cs;
    for (u32 k = 0; k < (1u << f.s.log_last); k++) last_poly.coeffs.push_back(QM31Var::new_witness(cs, Def::input_qm31(w.take(tape::S_LAST_COEFFS, 0, 0, 4 * k, 4))));

    // ---- LastFiatShamirResults::compute (last/fiat_shamir/src/lib.rs:164-216) ---------------------------------------------
    const CirclePointQM31Var oods_point = CirclePointQM31Var::from_t(in_t);
    std::vector<QM31Var> flat;
    for (const auto &tree : sampled_values)
        for (const auto &col : tree)
            for (const QM31Var &v : col) flat.push_back(v);
    Poseidon31MerkleHasherVar::hash_qm31_columns_get_rate(flat).equalverify(in_sampled_values_hash);
    LookupElementsVar lookup;
    lookup.z = in_z; lookup.alpha = in_alpha;
    lookup.alpha_powers[0] = QM31Var::one(cs); lookup.alpha_powers[1] = in_alpha; lookup.alpha_powers[2] = in_alpha * in_alpha;
    std::vector<M31Var> queries;
    for (const QM31Var &packed : in_packed_queries) {
        const std::array<M31Var, 4> d = packed.decompose_m31();
        queries.insert(queries.end(), d.begin(), d.end());
    }
    queries.resize(nq);
    QM31Var input_sum = QM31Var::zero(cs);
    {
        const QM31Var s1 = (QM31Var::one(cs) + lookup.alpha) - lookup.z;
        input_sum = input_sum + s1.inv();
        const QM31Var alpha_two = lookup.alpha + lookup.alpha;
        const QM31Var s2 = (QM31Var::i(cs) + alpha_two) - lookup.z;
        input_sum = input_sum + s2.inv();
        const QM31Var alpha_three = alpha_two + lookup.alpha;
        const QM31Var s3 = (QM31Var::j(cs) + alpha_three) - lookup.z;
        input_sum = input_sum + s3.inv();
    }
    ((input_sum + in_poseidon_total_sum) + in_plonk_total_sum).equalverify(QM31Var::zero(cs));

    // ---- LastAnswerResults::compute (last/answer/src/lib.rs:30-262) -------------------------------------------------------
    const AnswerResults ans = AnswerResults::compute_with(w, oods_point, f, queries, in_after, sampled_values, f.s.log_blowup + 1,
                                                          [&](const QueryPositionsPerLogSizeVar &) {
        // LastDecommitVar::compute -> LastSinglePathMerkleProofVar::from_proof_and_input (merkle_proofs.rs:114-160)
        std::vector<std::vector<AnswerResults::Columns>> cols(4);
        for (u32 t = 0; t < 4; t++)
            for (u32 i = 0; i < nq; i++) {
                AnswerResults::Columns per_layer;
                for (const auto &l : SinglePathMerkleProofVar::layer_layout(f, t)) {
                    std::vector<M31Var> vars;
                    for (u32 k = 0; k < l.second.second; k++) vars.push_back(M31Var::new_witness(cs, Def::input_m31(w.take(tape::S_PATH_COL, t, i, l.second.first + k, 1))));
                    const std::vector<QM31Var> &packed = in_packed[t][i].at(l.first);
                    if (vars.size() <= 8) {
                        for (size_t k = 0; k < vars.size(); k += 4) {
                            const std::array<M31Var, 4> d = packed[k / 4].decompose_m31();
                            for (size_t j = 0; j < 4 && k + j < vars.size(); j++) vars[k + j].equalverify(d[j]);
                        }
                    } else {
                        const std::array<QM31Var, 2> h = Poseidon31MerkleHasherVar::hash_m31_columns_get_rate(vars).to_qm31();
                        h[0].equalverify(packed[0]);
                        h[1].equalverify(packed[1]);
                    }
                    per_layer[l.first] = vars;
                }
                cols[t].push_back(per_layer);
            }
        return cols;
    });

    // ---- LastFoldingResults::compute (last/folding/src/lib.rs:14-161) -------------------------------------------------------
    const QueryPositionsPerLogSizeVar &qp = *ans.query_positions_per_log_size;
    for (auto it = f.all_log_sizes.rbegin(); it != f.all_log_sizes.rend(); ++it)
        for (u32 i = 0; i < nq; i++) in_first[i].self_columns.at(*it).equalverify(ans.fri_answers.at(*it)[i]);
    auto fold = [&](const QM31Var &self_val, const QM31Var &sibling_val, const M31Var &inv, u32 bit, const QM31Var &alpha) {
        const std::pair<QM31Var, QM31Var> lr = QM31Var::swap(self_val, sibling_val, bit);
        const QM31Var new_left_val = lr.first + lr.second;
        const QM31Var diff = lr.first - lr.second;
        const QM31Var new_right_val = diff * inv;
        const QM31Var ra = new_right_val * alpha;
        return new_left_val + ra;
    };
    std::map<u32, std::vector<QM31Var>> folded_results;
    for (u32 L : f.all_log_sizes)
        for (u32 i = 0; i < nq; i++) {
            const PointCarryingQueryVar &query = qp[L][i];
            const CirclePointM31Var point = query.get_absolute_point().double_();
            const M31Var y_inv = point.y.inv();
            folded_results[L].push_back(fold(in_first[i].self_columns.at(L), in_first[i].siblings_columns.at(L), y_inv, query.bits.variables[0],
                                             in_fri_alphas[f.max_first - L]));
        }
    u32 log_size = f.max_first;
    std::vector<QM31Var> folded(nq, QM31Var::zero(cs));
    for (u32 i = 0; i < f.s.n_inner; i++) {
        auto fr = folded_results.find(log_size);
        if (fr != folded_results.end()) {
            QM31Var fri_alpha = in_fri_alphas[i];
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
            const PairIn &proof = in_inner.at(log_size)[k];
            folded[k].equalverify(proof.self_columns.at(log_size));
            const M31Var x_inv = query.get_absolute_point().x.inv();
            new_folded.push_back(fold(proof.self_columns.at(log_size), proof.siblings_columns.at(log_size), x_inv, query.bits.variables[0], in_fri_alphas[i + 1]));
        }
        folded = new_folded;
    }
    for (u32 k = 0; k < nq; k++) {
        if (last_poly.coeffs.size() == 1) folded[k].equalverify(last_poly.coeffs[0]);
        else {
            const M31Var x = qp[log_size][k].get_next_point_x();
            const QM31Var eval = last_poly.eval_at_point(x);
            folded[k].equalverify(eval);
        }
    }
    cs->pad();
    out.cs = cs;
    out.gather = w.gather;
    return out;
}

}  // namespace dsl
}  // namespace stwo_b200
