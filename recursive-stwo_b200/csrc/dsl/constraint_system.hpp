// Host-side recorder behind the kept `ConstraintSystemRef` API (constraint_system/src/lib.rs:33-281, the
// Plonk-with-Poseidon system of constraint_system/src/plonk_with_poseidon.rs:18-335).
//
// B200 design: the wiring a DSL program emits depends only on the circuit's shape, never on proof values, so it is
// recorded ONCE per shape and values never exist on the host.  Every variable that is not the output of a row carries
// a *definition* instead of a value (an input slot of the per-item witness stream, or a hint: inverse, bit, coordinate,
// Poseidon output); the definitions form a tape that the device evaluates for a whole batch (tape.cuh, K6), one lane
// per batch item, levelised so that independent definitions run on different warps.
#pragma once
#include <cstdint>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../tape.cuh"

namespace stwo_b200 {
namespace dsl {

enum class AllocationMode { PublicInput, Witness, Constant };    // constraint_system/src/var.rs:13-18
enum class ConstraintSystemType { PlonkWithPoseidon, PlonkWithoutPoseidon };   // constraint_system/src/lib.rs:24-28

struct QM31Const { u32 v[4]; };

// A hint / input definition for a witness variable (dst is filled in by the constraint system).
struct Def {
    u32 op, a, b;
    static Def input_m31(u32 slot) { return {tape::T_INPUT_M31, slot, 0}; }
    static Def input_qm31(u32 slot) { return {tape::T_INPUT_QM31, slot, 0}; }
    static Def inv_m31(u32 var) { return {tape::T_INV_M31, var, 0}; }
    static Def inv_qm31(u32 var) { return {tape::T_INV_QM31, var, 0}; }
    static Def inv_cm31_re(u32 var) { return {tape::T_INV_CM31_RE, var, 0}; }
    static Def inv_cm31_im(u32 var) { return {tape::T_INV_CM31_IM, var, 0}; }
    static Def coordinate(u32 var, u32 k) { return {tape::T_COORD, var, k}; }
    static Def bit(u32 var, u32 k) { return {tape::T_BIT, var, k}; }
    static Def poseidon_out() { return {tape::T_NONE, 0, 0}; }   // written by the permutation record that owns it
};

struct PoseidonEntry { u32 wire; };                 // the hash lives on the device (flow_hash), per batch item
struct SwapOption { u32 addr; bool has_swap; };

class ConstraintSystem {
public:
    ConstraintSystemType type = ConstraintSystemType::PlonkWithPoseidon;
    bool without() const { return type == ConstraintSystemType::PlonkWithoutPoseidon; }
    // wiring (shared by every batch item of the shape).  Plonk-without-Poseidon (plonk_without_poseidon.rs:12-25): `op` is
    // op1 and op2..op4 select the gate; poseidon_wire / enforce_c_m31 stay zero.
    std::vector<u32> a_wire, b_wire, c_wire, poseidon_wire, enforce_c_m31, op, op2, op3, op4;
    std::vector<uint8_t> op_follows_c;              // rows whose op constant is the selected VALUE (circle/src/lib.rs:80-92)
    std::vector<u32> flow_wire;                     // n_flow x 4
    std::vector<u32> flow_swap_addr;                // n_flow
    // value definitions
    std::vector<tape::Ins> tape_;                   // recording order
    std::vector<tape::Perm> perms;                  // one per Poseidon flow entry
    std::vector<u32> eperms;                        // one tape::EPOSEIDON_REC record per emulated permutation
    bool macro_on = false;                          // inside an emulated permutation: variables are defined by its record
    std::vector<u32> macro_vars;
    u32 n_vars = 0, n_input_words = 0, num_input = 3;
    bool is_program_started = false, padded = false;
    u32 n_rows_unpadded = 0, n_flow_unpadded = 0;
    std::unordered_map<std::string, u32> cache;

    explicit ConstraintSystem(ConstraintSystemType t = ConstraintSystemType::PlonkWithPoseidon) : type(t) {   // plonk_with_poseidon.rs:43-99
        n_vars = 4;                                 // 0, 1, i, j: written by the evaluator's prologue
        for (u32 k = 0; k < 4; k++) row(k, 0, k, 1);
    }

    void row(u32 a, u32 b, u32 c, u32 op_, u32 pw = 0, u32 enf = 0, bool follows = false) {
        if (padded) throw std::logic_error("constraint system already padded");
        a_wire.push_back(a); b_wire.push_back(b); c_wire.push_back(c);
        poseidon_wire.push_back(pw); enforce_c_m31.push_back(enf); op.push_back(op_ % M31_P);
        op2.push_back(0); op3.push_back(0); op4.push_back(0);
        op_follows_c.push_back(follows ? 1 : 0);
    }
    // a row of one of the five special gates of the Plonk-without-Poseidon system: selectors (1, s2, s3, s4)
    u32 special_gate(u32 tape_op, u32 a, u32 b, u32 s2, u32 s3, u32 s4) {
        if (!without()) throw std::logic_error("unimplemented!() for the Plonk-with-Poseidon system (constraint_system/src/lib.rs:114-152)");
        is_program_started = true;
        const u32 c = fresh(tape_op, a, b);
        row(a, b, c, 1);
        op2.back() = s2; op3.back() = s3; op4.back() = s4;
        return c;
    }
    u32 do_m4_gate(u32 a, u32 b) { return special_gate(tape::T_M4, a, b, 0, 1, 0); }             // plonk_without_poseidon.rs:108-139
    u32 do_pow5m4_gate(u32 a, u32 b) { return special_gate(tape::T_POW5M4, a, b, 1, 1, 0); }     // :140-173
    u32 do_pow5_gate(u32 a, u32 b) { return special_gate(tape::T_HADAMARD, a, b, 1, 0, 1); }     // :174-198
    u32 do_hadamard(u32 a, u32 b) { return special_gate(tape::T_HADAMARD, a, b, 0, 0, 1); }      // :199-223
    u32 do_grandsum_gate(u32 a, u32 b) { return special_gate(tape::T_GRANDSUM, a, b, 0, 1, 1); } // :224-245
    u32 fresh(u32 op_, u32 a, u32 b) {
        const u32 c = n_vars++;
        if (macro_on) macro_vars.push_back(c);
        else if (op_ != tape::T_NONE) tape_.push_back({op_, c, a, b});
        return c;
    }
    // Brackets the body of poseidon_permute_emulated: the rows are recorded as usual, the 401 variable definitions collapse
    // into one T_EPOSEIDON instruction.  Constants allocated on first use inside the body keep their own definitions.
    struct MacroPause {
        ConstraintSystem &cs; bool was;
        explicit MacroPause(ConstraintSystem &c) : cs(c), was(c.macro_on) { cs.macro_on = false; }
        ~MacroPause() { cs.macro_on = was; }
    };
    void begin_macro() { macro_on = true; macro_vars.clear(); }
    void end_macro(const u32 in[4]) {
        macro_on = false;
        if (macro_vars.size() != tape::EPOSEIDON_VARS) throw std::logic_error("emulated permutation did not create 401 variables");
        const u32 rec = (u32)(eperms.size() / tape::EPOSEIDON_REC);
        eperms.insert(eperms.end(), in, in + 4);
        eperms.insert(eperms.end(), macro_vars.begin(), macro_vars.end());
        tape_.push_back({tape::T_EPOSEIDON, rec, 0, 0});
    }
    void insert_gate(u32 a, u32 b, u32 c, u32 op_) {                 // :101-115
        is_program_started = true;
        if (a >= n_vars || b >= n_vars || c >= n_vars) throw std::out_of_range("gate wire");
        row(a, b, c, op_);
    }
    void enforce_zero(u32 var) { is_program_started = true; row(var, 0, 0, 1); }       // :130-139
    u32 add(u32 a, u32 b) { const u32 c = fresh(tape::T_ADD, a, b); insert_gate(a, b, c, 1); return c; }     // :141-150
    u32 mul(u32 a, u32 b) { const u32 c = fresh(tape::T_MUL, a, b); insert_gate(a, b, c, 0); return c; }     // :173-182
    u32 mul_constant(u32 a, u32 k, bool op_follows_value = false) {                    // :184-192
        k %= M31_P;
        const u32 c = fresh(tape::T_MULC, a, k);
        is_program_started = true;
        row(a, 0, c, k, 0, 0, op_follows_value);
        return c;
    }
    u32 assemble_poseidon_gate(u32 a, u32 b) {                       // :152-171
        if (without()) throw std::logic_error("unimplemented!() for the Plonk-without-Poseidon system (constraint_system/src/lib.rs:267-277)");
        const u32 c = fresh(tape::T_MUL, a, b);
        is_program_started = true;
        row(a, b, c, 0, c);
        return c;
    }
    u32 new_m31_constant(u32 value) {                                // :221-230
        MacroPause pause(*this);
        is_program_started = true;
        const u32 c = fresh(tape::T_MULC, 1, value % M31_P);
        row(1, 0, c, value);
        return c;
    }
    u32 new_m31(const Def &d, AllocationMode mode) {                 // :194-220
        if (mode == AllocationMode::Constant) throw std::logic_error("constants carry a value, not a definition");
        if (mode == AllocationMode::PublicInput) {
            if (is_program_started) throw std::logic_error("public inputs must be allocated first");
            num_input += 1;
        } else is_program_started = true;
        const u32 c = fresh(d.op, d.a, d.b);
        if (without()) { row(c, 1, c, 1); op4.back() = 1; }          // hadamard with variable 1 (plonk_without_poseidon.rs:294-318)
        else row(c, 0, c, 1, 0, 1);
        return c;
    }
    u32 new_qm31(const Def &d, AllocationMode mode) {                // :235-255; plonk_without_poseidon.rs:335-363
        if (mode == AllocationMode::Constant) throw std::logic_error("constants carry a value, not a definition");
        const u32 c = fresh(d.op, d.a, d.b);
        if (mode == AllocationMode::PublicInput) {
            if (is_program_started) throw std::logic_error("public inputs must be allocated first");
            if (without()) row(c, 0, c, 1); else row(c, 0, c, 1, 0, 1);
            num_input += 1;
        } else {
            is_program_started = true;
            if (without()) row(c, 0, c, 1);                          // a QM31 witness costs a row in this system
        }
        return c;
    }
    u32 new_qm31_constant(const QM31Const &v) {                      // :256-277
        MacroPause pause(*this);
        is_program_started = true;
        const u32 c = n_vars++;                                      // value = a + b of the tie row, defined below
        const u32 fr = new_m31_constant(v.v[0]), fi = new_m31_constant(v.v[1]);
        const u32 sr = new_m31_constant(v.v[2]), si = new_m31_constant(v.v[3]);
        u32 t = mul(fi, 2);
        const u32 a = add(fr, t);
        t = mul(si, 2);
        t = add(sr, t);
        const u32 b = mul(t, 3);
        tape_.push_back({tape::T_ADD, c, a, b});
        row(a, b, c, 1);
        return c;
    }
    u32 new_input_words(u32 n) { const u32 s = n_input_words; n_input_words += n; return s; }

    // The permutation must precede, on the tape, the rows that consume its outputs (their assembling gates are recorded
    // before the flow entry): the caller reserves the tape slot first and hands it back here.
    // Native-verifier hints (verify.cuh HintLayout): a gadget that knows which permutation of the native pass it is about to
    // repeat queues that permutation's slot; the next recorded permutation takes it.
    std::deque<u32> hint_queue;
    u32 transcript_slot = 0;
    bool native_hints = false;
    void push_hint(u32 slot) { if (native_hints && !without()) hint_queue.push_back(slot); }
    size_t reserve_tape_slot() { tape_.push_back({tape::T_NONE, 0, 0, 0}); return tape_.size() - 1; }
    void invoke_poseidon_accelerator(PoseidonEntry e1, PoseidonEntry e2, PoseidonEntry e3, PoseidonEntry e4, SwapOption s,
                                     const tape::Perm &p, size_t tape_slot) {
        flow_wire.push_back(e1.wire); flow_wire.push_back(e2.wire); flow_wire.push_back(e3.wire); flow_wire.push_back(e4.wire);
        flow_swap_addr.push_back(s.has_swap ? s.addr : 0);
        perms.push_back(p);
        if (!hint_queue.empty()) { perms.back().hint = hint_queue.front() + 1; hint_queue.pop_front(); }
        tape_[tape_slot] = {tape::T_POSEIDON, (u32)perms.size() - 1, 0, 0};
    }
    u32 num_plonk_rows() const { return (u32)a_wire.size(); }
    u32 num_poseidon_invocations() const { return (u32)flow_swap_addr.size(); }

    // :283-335.  The Poseidon flow is padded by the prover-facing export (CONSTANT_1/2/3 entries are value-only and
    // never reach the trace columns); the row columns are padded here.
    void pad() {
        n_rows_unpadded = num_plonk_rows(); n_flow_unpadded = num_poseidon_invocations();
        u32 n = 1;
        while (n < n_rows_unpadded) n <<= 1;
        for (u32 i = n_rows_unpadded; i < n; i++) row(0, 0, 0, 1);
        padded = true;
    }
    u32 padded_poseidon_len() const { const u32 n = n_flow_unpadded; const u32 r = (n + 15) / 16 * 16; return r < 32 ? 32 : r; }
};

// constraint_system/src/lib.rs:33 -- a cheap shared handle
struct ConstraintSystemRef {
    std::shared_ptr<ConstraintSystem> p;
    static ConstraintSystemRef new_plonk_with_poseidon_ref() { return {std::make_shared<ConstraintSystem>()}; }
    static ConstraintSystemRef new_plonk_without_poseidon_ref() {
        return {std::make_shared<ConstraintSystem>(ConstraintSystemType::PlonkWithoutPoseidon)};
    }
    ConstraintSystemType get_type() const { return p->type; }
    ConstraintSystem *operator->() const { return p.get(); }
    bool operator==(const ConstraintSystemRef &o) const { return p == o.p; }
    const ConstraintSystemRef &and_(const ConstraintSystemRef &o) const {
        if (!(*this == o)) throw std::logic_error("variables of different constraint systems");
        return *this;
    }
    bool get_cache(const std::string &k, u32 &out) const {
        auto it = p->cache.find(k);
        if (it == p->cache.end()) return false;
        out = it->second;
        return true;
    }
    void set_cache(const std::string &k, u32 v) const { p->cache[k] = v; }
};

}  // namespace dsl
}  // namespace stwo_b200
