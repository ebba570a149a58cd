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

    // sources / outputs of one tape instruction (variables)
    template <class F> static void for_sources(const ConstraintSystem &c, const tape::Ins &in, F f) {
        switch (in.op) {
        case tape::T_ADD: case tape::T_MUL: case tape::T_POW5M4: case tape::T_HADAMARD: case tape::T_GRANDSUM: f(in.a); f(in.b); break;
        case tape::T_MULC: case tape::T_INV_M31: case tape::T_INV_QM31: case tape::T_INV_CM31_RE: case tape::T_INV_CM31_IM:
        case tape::T_COORD: case tape::T_BIT: case tape::T_M4: case tape::T_POW4: f(in.a); break;
        case tape::T_POSEIDON: {
            const tape::Perm &p = c.perms[in.dst];
            if (p.l_kind == 0) { f(p.l_a); f(p.l_b); }
            if (p.r_kind == 0) { f(p.r_a); f(p.r_b); }
            if (p.swap_var != tape::NO_VAR) f(p.swap_var);
            break;
        }
        case tape::T_EPOSEIDON: for (u32 q = 0; q < 4; q++) f(c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + q]); break;
        default: break;
        }
    }
    template <class F> static void for_outputs(const ConstraintSystem &c, const tape::Ins &in, F f) {
        if (in.op == tape::T_POSEIDON) { for (u32 o : c.perms[in.dst].out) if (o != tape::NO_VAR) f(o); }
        else if (in.op == tape::T_EPOSEIDON) { for (u32 q = 0; q < tape::EPOSEIDON_VARS; q++) f(c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + 4 + q]); }
        else f(in.dst);
    }

    // Levels: an instruction's earliest level is 1 + the latest earliest level of its sources (the depth of the circuit is the longest
    // such chain).  Within that depth every instruction is placed as LATE as its consumers allow (an instruction nobody consumes stays
    // at its earliest level): a value is then produced just before it is first read.  The evaluator walks the levels over every batch
    // item, so the distance between producer and consumer decides whether the operand is still in L2 -- with earliest-level placement
    // 60 % of the verifier circuit's instructions sit in the first seven levels and are read up to 250 levels later.
    void levelise() {
        const ConstraintSystem &c = *cs.p;
        const size_t n = c.tape_.size();
        std::vector<u32> var_level(c.n_vars, 0), early(n, 0);
        u32 max_level = 0;
        for (size_t k = 0; k < n; k++) {
            u32 l = 0;
            for_sources(c, c.tape_[k], [&](u32 v) { l = std::max(l, var_level[v]); });
            l += 1;
            early[k] = l;
            max_level = std::max(max_level, l);
            for_outputs(c, c.tape_[k], [&](u32 v) { var_level[v] = l; });
        }
        // backward: consumers come after their producers in recording order
        constexpr u32 NONE = 0xffffffffu;
        std::vector<u32> need(c.n_vars, NONE), ins_level(n, 0);
        for (size_t k = n; k-- > 0;) {
            u32 first_use = NONE;
            for_outputs(c, c.tape_[k], [&](u32 v) { first_use = std::min(first_use, need[v]); });
            const u32 l = first_use == NONE ? early[k] : first_use - 1;
            ins_level[k] = l;
            for_sources(c, c.tape_[k], [&](u32 v) { need[v] = std::min(need[v], l); });
        }
        // levels are 1 .. max_level; level l occupies ins[level_start[l-1] .. level_start[l])
        std::vector<u32> cnt(max_level + 1, 0);
        for (u32 l : ins_level) cnt[l]++;
        level_start.assign(max_level + 1, 0);
        for (u32 l = 1; l <= max_level; l++) level_start[l] = level_start[l - 1] + cnt[l];
        std::vector<u32> at(level_start.begin(), level_start.end());
        ins.resize(n);
        // inside a level the permutations go first: the evaluator deals a level's instructions round-robin to its warps, and
        // the heavy items (a permutation is ~50x a field gate) then spread evenly instead of following the recording pattern
        for (int pass = 0; pass < 2; pass++)
            for (size_t k = 0; k < n; k++) {
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
