// A recorded circuit in the flat form the device consumes: wiring columns, the levelised tape, permutation records
// and (for the verifier circuit) the witness-stream gather table.  Host only; built once per shape.
#pragma once
#include "last_layer.hpp"

namespace stwo_b200 {
namespace dsl {

struct RecordedCircuit {
    ConstraintSystemRef cs;
    std::vector<u32> gather;
    u32 words_per_instance = 0, multipliers = 1;
    ProofShape shape{};
    std::vector<ExtraHashJob> jobs;        // last-layer circuit: public-input hashes computed on the device before the gather
    u32 n_extra_words = 0;
    // tape sorted by dependency level: instructions of one level are independent of each other
    std::vector<tape::Ins> ins;
    std::vector<u32> level_start;          // n_levels + 1
    u32 n_levels() const { return (u32)level_start.size() - 1; }

    void levelise() {
        const ConstraintSystem &c = *cs.p;
        std::vector<u32> var_level(c.n_vars, 0), ins_level(c.tape_.size(), 0);
        auto lv = [&](u32 v) { return v == tape::NO_VAR ? 0u : var_level[v]; };
        u32 max_level = 0;
        for (size_t k = 0; k < c.tape_.size(); k++) {
            const tape::Ins &in = c.tape_[k];
            u32 l = 0;
            switch (in.op) {
            case tape::T_ADD: case tape::T_MUL: case tape::T_POW5M4: case tape::T_HADAMARD: case tape::T_GRANDSUM:
                l = std::max(lv(in.a), lv(in.b)); break;
            case tape::T_MULC: case tape::T_INV_M31: case tape::T_INV_QM31: case tape::T_INV_CM31_RE: case tape::T_INV_CM31_IM:
            case tape::T_COORD: case tape::T_BIT: case tape::T_M4: case tape::T_POW4: l = lv(in.a); break;
            case tape::T_POSEIDON: {
                const tape::Perm &p = c.perms[in.dst];
                if (p.l_kind == 0) l = std::max(l, std::max(lv(p.l_a), lv(p.l_b)));
                if (p.r_kind == 0) l = std::max(l, std::max(lv(p.r_a), lv(p.r_b)));
                l = std::max(l, lv(p.swap_var));
                break;
            }
            case tape::T_EPOSEIDON:
                for (u32 q = 0; q < 4; q++) l = std::max(l, lv(c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + q]));
                break;
            default: break;
            }
            l += 1;
            ins_level[k] = l;
            max_level = std::max(max_level, l);
            if (in.op == tape::T_POSEIDON) {
                for (u32 o : c.perms[in.dst].out) if (o != tape::NO_VAR) var_level[o] = l;
            } else if (in.op == tape::T_EPOSEIDON) {
                for (u32 q = 0; q < tape::EPOSEIDON_VARS; q++) var_level[c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + 4 + q]] = l;
            } else var_level[in.dst] = l;
        }
        // levels are 1 .. max_level; level l occupies ins[level_start[l-1] .. level_start[l])
        std::vector<u32> cnt(max_level + 1, 0);
        for (u32 l : ins_level) cnt[l]++;
        level_start.assign(max_level + 1, 0);
        for (u32 l = 1; l <= max_level; l++) level_start[l] = level_start[l - 1] + cnt[l];
        std::vector<u32> at(level_start.begin(), level_start.end());
        ins.resize(c.tape_.size());
        // inside a level the permutations go first: the evaluator deals a level's instructions round-robin to its warps, and
        // the heavy items (a permutation is ~50x a field gate) then spread evenly instead of following the recording pattern
        for (int pass = 0; pass < 2; pass++)
            for (size_t k = 0; k < c.tape_.size(); k++) {
                const bool heavy = c.tape_[k].op == tape::T_POSEIDON || c.tape_[k].op == tape::T_EPOSEIDON;
                if (heavy == (pass == 0)) ins[at[ins_level[k] - 1]++] = c.tape_[k];
            }
    }
};

inline std::unique_ptr<RecordedCircuit> record_verifier(const ProofShape &shape, const std::vector<PublicInput> &inputs, u32 multipliers) {
    VerifierCircuit vc = record_verifier_circuit(shape, inputs, multipliers);
    std::unique_ptr<RecordedCircuit> r(new RecordedCircuit());
    r->cs = vc.cs; r->gather = std::move(vc.gather); r->words_per_instance = vc.words_per_instance;
    r->multipliers = multipliers; r->shape = shape;
    r->levelise();
    return r;
}

inline std::unique_ptr<RecordedCircuit> record_last_layer(const ProofShape &shape) {
    LastLayerCircuit lc = record_last_layer_circuit(shape);
    std::unique_ptr<RecordedCircuit> r(new RecordedCircuit());
    r->cs = lc.cs; r->gather = std::move(lc.gather); r->words_per_instance = lc.cs->n_input_words;
    r->multipliers = 1; r->shape = shape;
    r->jobs = std::move(lc.jobs); r->n_extra_words = lc.n_extra_words;
    r->levelise();
    return r;
}

}  // namespace dsl
}  // namespace stwo_b200
