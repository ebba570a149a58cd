// A recorded circuit in the flat form the device consumes: wiring columns, the levelised tape, permutation records
// and (for the verifier circuit) the witness-stream gather table.  Host only; built once per shape.
#pragma once
#include "last_layer.hpp"
#include <stdlib.h>
#include <algorithm>

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
        case tape::T_POSEIDON: case tape::T_PERM_FLOW: {
            const tape::Perm &p = c.perms[in.dst];
            if (p.l_kind == 0) { f(p.l_a); f(p.l_b); }
            if (p.r_kind == 0) { f(p.r_a); f(p.r_b); }
            if (p.swap_var != tape::NO_VAR) f(p.swap_var);
            break;
        }
        case tape::T_PERM_OUT: break;                  // the outputs come from the record of executed permutations
        case tape::T_EPOSEIDON: for (u32 q = 0; q < 4; q++) f(c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + q]); break;
        default: break;
        }
    }
    template <class F> static void for_outputs(const ConstraintSystem &c, const tape::Ins &in, F f) {
        if (in.op == tape::T_POSEIDON || in.op == tape::T_PERM_OUT) { for (u32 o : c.perms[in.dst].out) if (o != tape::NO_VAR) f(o); }
        else if (in.op == tape::T_PERM_FLOW) {}
        else if (in.op == tape::T_EPOSEIDON) { for (u32 q = 0; q < tape::EPOSEIDON_VARS; q++) f(c.eperms[(size_t)in.dst * tape::EPOSEIDON_REC + 4 + q]); }
        else f(in.dst);
    }

    // Bundles and levels.  A BUNDLE is a short chain of instructions one warp executes back to back for its 32 items (each reads what the
    // one before it wrote: same thread, program order, no barrier); the instructions of different bundles of one LEVEL are independent,
    // and a barrier separates levels.  A bundle's earliest level is 1 + the latest level among the sources it takes from other bundles,
    // so a dependent chain of k gates costs ceil(k / kBundle) levels instead of k: the verifier circuit of shape S is 265 gates deep
    // (the per-query fold and point chains, not the hashes) and most of those levels hold a handful of instructions -- pure barrier
    // and load latency.  Within the resulting depth every bundle is then placed as LATE as its consumers allow (a bundle nobody consumes
    // stays at its earliest level), so a value is produced just before it is first read and is still in L2.
    std::vector<u32> bundle_start;         // n_bundles + 1: first instruction of each bundle (ins is bundle-major inside a level)
    std::vector<u32> level_bundle;         // n_levels + 1: first bundle of each level
    // A SECOND ORDER of the same tape, for items whose permutation record is complete (the native verifier executed every permutation of
    // the circuit and kept its output): each recorded permutation is split into T_PERM_OUT (output variables <- record; no sources) and
    // T_PERM_FLOW (the flow entry: reads the input halves, defines nothing).  The transcript (100-255 permutations, each absorbing what
    // the one before produced) and the authentication paths (a permutation per tree level) then stop being dependency chains: what is
    // left of the depth is the arithmetic.  Empty when the circuit has no recorded permutations.
    struct Order { std::vector<tape::Ins> ins; std::vector<u32> level_start, bundle_start, level_bundle; u32 n_levels() const { return level_start.empty() ? 0u : (u32)level_start.size() - 1; } };
    Order recorded;
    static u32 bundle_cap() {
        const char *e = getenv("STWO_B200_BUNDLE");           // 1 = one instruction per bundle (profiling)
        const int v = e ? atoi(e) : 8;
        return v < 1 ? 1u : v > 64 ? 64u : (u32)v;
    }
    void levelise() {
        const ConstraintSystem &c = *cs.p;
        Order o;
        build_order(c.tape_, o);
        ins = std::move(o.ins); level_start = std::move(o.level_start); bundle_start = std::move(o.bundle_start); level_bundle = std::move(o.level_bundle);
        std::vector<tape::Ins> split;
        bool any = false;
        for (const tape::Ins &in : c.tape_) {
            if (in.op == tape::T_POSEIDON && c.perms[in.dst].hint) {
                split.push_back(tape::Ins{tape::T_PERM_OUT, in.dst, 0, 0});
                split.push_back(tape::Ins{tape::T_PERM_FLOW, in.dst, 0, 0});
                any = true;
            } else split.push_back(in);
        }
        recorded = Order();
        // Built on request only (STWO_B200_RECORDED_ORDER=1).  Measured on B200: 74 -> 55 levels for shape S (100 / 107 -> 55 for the
        // larger shapes), tape evaluation 1.00 -> 0.92 ms at 512 proofs and 2.02 -> 1.92 ms for 34 proofs of an 80-query shape, but
        // 2.80 -> 3.18 ms at 4096 proofs (7 % more instructions, the record read twice per permutation: the large batch is bound by
        // throughput, not by depth) and no change of the pipelined step at any size -- not adopted as the default.
        const char *e = getenv("STWO_B200_RECORDED_ORDER");
        if (any && e && e[0] == '1') build_order(split, recorded);
    }
    void build_order(const std::vector<tape::Ins> &tape_in, Order &out) const {
        const ConstraintSystem &c = *cs.p;
        const size_t n = tape_in.size();
        const u32 cap = bundle_cap();
        constexpr u32 NONE = 0xffffffffu;
        std::vector<u32> var_level(c.n_vars, 0), producer(c.n_vars, NONE), bundle_of(n, 0);
        std::vector<u32> b_level, b_len, b_tail;           // per bundle: earliest level, length, last instruction
        std::vector<std::vector<u32>> members;
        for (size_t k = 0; k < n; k++) {
            const tape::Ins &in = tape_in[k];
            // a bundle this instruction may join: the bundle of one of its sources, if it has room and every other source is either inside
            // that bundle too or complete before the bundle's level starts (the bundle runs in recording order on one thread per item,
            // so anything recorded earlier inside it is visible)
            u32 join = NONE;
            for_sources(c, in, [&](u32 v) {
                const u32 p = producer[v];
                if (join != NONE || p == NONE) return;
                const u32 b = bundle_of[p];
                if (b_len[b] >= cap) return;
                bool ok = true;
                for_sources(c, in, [&](u32 w) { const u32 q = producer[w]; if (!((q != NONE && bundle_of[q] == b) || var_level[w] < b_level[b])) ok = false; });
                if (ok) join = b;
            });
            u32 b;
            if (join != NONE) { b = join; b_len[b]++; b_tail[b] = (u32)k; members[b].push_back((u32)k); }
            else {
                u32 l = 0;
                for_sources(c, in, [&](u32 v) { l = std::max(l, var_level[v]); });
                b = (u32)b_level.size();
                b_level.push_back(l + 1); b_len.push_back(1); b_tail.push_back((u32)k); members.push_back({(u32)k});
            }
            bundle_of[k] = b;
            for_outputs(c, in, [&](u32 v) { var_level[v] = b_level[b]; producer[v] = (u32)k; });
        }
        const u32 nb = (u32)b_level.size();
        u32 max_level = 0;
        for (u32 l : b_level) max_level = std::max(max_level, l);
        // as late as possible, bundle by bundle, consumers (strictly higher earliest level) first
        std::vector<u32> order(nb), late(nb, 0), need(c.n_vars, NONE);
        for (u32 b = 0; b < nb; b++) order[b] = b;
        std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return b_level[x] > b_level[y]; });
        for (u32 b : order) {
            u32 first_use = NONE;
            for (u32 k : members[b]) for_outputs(c, tape_in[k], [&](u32 v) { first_use = std::min(first_use, need[v]); });
            late[b] = first_use == NONE ? b_level[b] : first_use - 1;
            for (u32 k : members[b])
                for_sources(c, tape_in[k], [&](u32 v) { const u32 q = producer[v]; if (q == NONE || bundle_of[q] != b) need[v] = std::min(need[v], late[b]); });
        }
        // emit: levels 1 .. max_level; inside a level the bundles holding a permutation go first (the evaluator deals a level's bundles
        // round-robin to its warps, and the heavy ones then spread evenly)
        std::vector<std::vector<u32>> per_level(max_level + 1);
        for (int pass = 0; pass < 2; pass++)
            for (u32 b = 0; b < nb; b++) {
                bool heavy = false;
                for (u32 k : members[b]) heavy |= tape_in[k].op == tape::T_POSEIDON || tape_in[k].op == tape::T_EPOSEIDON || tape_in[k].op == tape::T_PERM_FLOW;
                if (heavy == (pass == 0)) per_level[late[b]].push_back(b);
            }
        out.ins.clear(); out.ins.reserve(n);
        out.bundle_start.clear(); out.level_bundle.assign(1, 0); out.level_start.assign(1, 0);
        for (u32 l = 1; l <= max_level; l++) {
            for (u32 b : per_level[l]) {
                out.bundle_start.push_back((u32)out.ins.size());
                for (u32 k : members[b]) out.ins.push_back(tape_in[k]);
            }
            out.level_bundle.push_back((u32)out.bundle_start.size());
            out.level_start.push_back((u32)out.ins.size());
        }
        out.bundle_start.push_back((u32)out.ins.size());
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

inline std::unique_ptr<RecordedCircuit> record_folding(const ProofShape &shape) {
    VerifierCircuit vc = record_folding_circuit(shape);
    std::unique_ptr<RecordedCircuit> r(new RecordedCircuit());
    r->cs = vc.cs; r->gather = std::move(vc.gather); r->words_per_instance = vc.words_per_instance;
    r->multipliers = 1; r->shape = shape;
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
