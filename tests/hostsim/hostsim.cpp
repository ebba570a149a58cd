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

// ---- full verifier on the host: same stage functions the CUDA kernels dispatch, grids emulated by loops -----------------
#include "../../recursive-stwo_b200/csrc/verify.cuh"
#include <stdlib.h>
#include <vector>
#include <string.h>
extern "C" {
size_t hs_detail_size() { return sizeof(verify::Detail); }
// returns the workspace (caller frees with hs_free); detail_out: n * sizeof(Detail)
void *hs_verify_batch(const u32 *blobs, const u64 *blob_off, u32 n, const u32 *shape7, const u32 *input_idx, const u32 *input_vals,
                      u32 n_inputs, int full, verify::Detail *detail_out, verify::Workspace *ws_out) {
    verify::Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape7, 7 * 4);
    ws.n_proofs = n; ws.blobs = blobs; ws.blob_off = blob_off; ws.input_idx = input_idx; ws.input_vals = input_vals; ws.n_inputs = n_inputs;
    size_t bytes = verify::carve(ws, nullptr);
    uint8_t *base = (uint8_t *)calloc(bytes, 1);
    verify::carve(ws, base);
    const u32 nq = ws.shape.n_queries, nf = ws.shape.n_fri_trees();
    const bool coop_fs = (full & 2) != 0;
    ws.mode = ((full & 1) ? verify::MODE_FULL : 0u) | (((full & 1) && (!(full & 2) || (full & 4))) ? verify::MODE_PATH_KERNELS : 0u);
    {
        std::vector<u32> ptab(verify::parse_tab_words());
        decommit::CoopOne one0;
        for (u32 p = 0; p < n; p++) {
            if (coop_fs) { verify::stage_parse_coop(one0, ws, p, ptab.data()); verify::stage_transcript(ws, p); verify::stage_oods(ws, p); }
            else verify::stage_fiat_shamir(ws, p);
        }
    }
    const bool coop = (full & 2) != 0;          // bit 1: the cooperative tree rebuilds (group of one lane on the host)
    const bool path_kernels = (full & 1) && (!coop || (full & 4));      // bit 2: the per-query path stages produce the record even with coop
    full &= 1;
    std::vector<u32> tab(decommit::pair_tab_words(nq) + decommit::single_tab_words(nq) + verify::folds_tab_words(nq) + 2 * nq + 64);
    decommit::CoopOne one;
    for (u32 p = 0; p < n; p++) for (u32 t = 0; t < 4; t++) {
        if (coop) verify::stage_single_tree_coop(one, ws, p, t, tab.data()); else verify::stage_single_tree(ws, p, t);
    }
    for (u32 p = 0; p < n; p++) for (u32 g = 0; g < fri::MAX_LOGS; g++) { if (coop) verify::stage_group_coop(one, ws, p, g); else verify::stage_group(ws, p, g); }
    for (u32 p = 0; p < n; p++) for (u32 g = 0; g < fri::MAX_LOGS; g++) for (u32 i = 0; i < nq; i++) verify::stage_answer(ws, p, g, i);
    for (u32 p = 0; p < n; p++) { if (coop) verify::stage_folds_coop(one, ws, p, tab.data()); else verify::stage_folds(ws, p); }
    for (u32 p = 0; p < n; p++) for (u32 f = 0; f < nf; f++) {
        if (coop) verify::stage_pair_tree_coop(one, ws, p, f, tab.data()); else verify::stage_pair_tree(ws, p, f);
    }
    if (path_kernels) {
        for (u32 p = 0; p < n; p++) for (u32 t = 0; t < 4; t++) for (u32 i = 0; i < nq; i++) verify::stage_single_path(ws, p, t, i);
        for (u32 p = 0; p < n; p++) for (u32 f = 0; f < nf; f++) for (u32 i = 0; i < nq; i++) verify::stage_pair_path(ws, p, f, i);
    }
    for (u32 p = 0; p < n; p++) verify::stage_verdict(ws, p);
    memcpy(detail_out, ws.detail, n * sizeof(verify::Detail));
    if (ws_out) *ws_out = ws;
    return base;
}
void hs_free(void *p) { free(p); }
// fri::build_group_coop with G lanes emulated on the host against the sequential fri::build_group, for every log-size group of proof 0
// of a verified batch (ws from hs_verify_batch's ws_out).  A lane only reads what another lane wrote across the one barrier (the batch
// points), and re-places its own samples at the start: running the lanes one after the other TWICE therefore reproduces the lock-step
// result.  Returns 0 when every group is bit-identical.
int hs_build_group_lanes(const verify::Workspace *ws, u32 G) {
    struct LaneOf { u32 l, g; u32 lane() const { return l; } u32 size() const { return g; } void sync() const {} };
    const proof::Desc &d = ws->desc[0];
    const verify::Detail &dt = ws->detail[0];
    static fri::Group a, b;
    for (u32 k = 0; k < dt.n_logs; k++) {
        memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
        if (!fri::build_group(ws->blob(0), d, dt.fs, dt.log_sizes[k], a)) return -1;
        for (int round = 0; round < 2; round++)
            for (u32 l = 0; l < G; l++) { LaneOf co{l, G}; if (!fri::build_group_coop(co, ws->blob(0), d, dt.fs, dt.log_sizes[k], b)) return -2; }
        if (a.n_batches != b.n_batches || a.n_cols != b.n_cols || memcmp(a.start, b.start, sizeof a.start) || memcmp(a.point, b.point, a.n_batches * sizeof(fri::QPoint))) return 1 + (int)k;
        const u32 ns = a.start[a.n_batches];
        if (memcmp(a.col, b.col, ns * 4) || memcmp(a.ca, b.ca, ns * 16) || memcmp(a.cb, b.cb, ns * 16) || memcmp(a.cc, b.cc, ns * 16)) return 10 + (int)k;
    }
    return 0;
}
// the permutation record of proof p (ws from hs_verify_batch's ws_out): hint_total x 16 words; returns hint_total
// optional second destination of hs_perm_record: the INPUT states of the same slots (set, call, clear)
static u32 *g_record_inputs_out = nullptr;
void hs_perm_record_inputs_to(u32 *dst) { g_record_inputs_out = dst; }
u32 hs_perm_record(const verify::Workspace *ws, u32 p, u32 *out, u32 *trees_complete) {
    if (out) memcpy(out, ws->perm_out_of(p, 0), (size_t)ws->hint_total * 64);
    if (g_record_inputs_out) memcpy(g_record_inputs_out, ws->perm_out_of(p, 0) + ws->in_delta(), (size_t)ws->hint_total * 64);
    if (trees_complete) *trees_complete = ws->hint_trees[p];
    return ws->hint_total;
}
size_t hs_workspace_struct_size() { return sizeof(verify::Workspace); }
// stage_after_transcript + stage_verdict on forged query draws: the one way to reach the case the reference panics on (duplicated
// queries at the largest domain, components/recursive/answer/src/lib.rs:190-195) without grinding a proof.  shape7 as above;
// returns verdict | stage << 8.
u32 hs_after_transcript_verdict(const u32 *shape7, const u32 *raw_queries, u32 pow_ok) {
    verify::Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape7, 7 * 4);
    ws.n_proofs = 1;
    static proof::Desc d;
    static verify::Detail dt;
    memset(&d, 0, sizeof d); memset(&dt, 0, sizeof dt);
    d.ok = 1;
    d.log_size_plonk = ws.shape.log_size_plonk; d.log_size_poseidon = ws.shape.log_size_poseidon; d.pow_bits = ws.shape.pow_bits;
    d.log_blowup = ws.shape.log_blowup; d.log_last = ws.shape.log_last; d.n_queries = ws.shape.n_queries; d.n_inner = ws.shape.n_inner;
    d.max_first = ws.shape.max_first(); d.log_plonk = ws.shape.log_plonk(); d.log_pos = ws.shape.log_pos();
    ws.desc = &d; ws.detail = &dt;
    verify::reset_detail(dt);
    for (u32 i = 0; i < d.n_queries; i++) dt.fs.raw_queries[i] = raw_queries[i];
    dt.fs.pow_ok = pow_ok;
    verify::stage_after_transcript(ws, 0);
    verify::stage_verdict(ws, 0);
    return dt.verdict | (dt.stage << 8);
}
}

// ---- circuit recorder + tape evaluator on the host -------------------------------------------------------------------------
#include "../../recursive-stwo_b200/csrc/circuit.cuh"
#include "../../recursive-stwo_b200/csrc/dsl/recorded.hpp"
#include <stdio.h>
extern "C" {
using stwo_b200::dsl::RecordedCircuit;
// shape: 7 words (stwo_b200_proof_shape); returns an owned handle or null (message on stderr)
void *hs_circuit_record(const u32 *shape, const u32 *input_idx, const u32 *input_vals, u32 n_inputs, u32 multipliers) {
    try {
        stwo_b200::dsl::ProofShape s{shape[0], shape[1], shape[2], shape[3], shape[4], shape[5], shape[6]};
        std::vector<stwo_b200::dsl::PublicInput> in;
        for (u32 k = 0; k < n_inputs; k++) in.push_back({input_idx[k], {{input_vals[4 * k], input_vals[4 * k + 1], input_vals[4 * k + 2], input_vals[4 * k + 3]}}});
        return stwo_b200::dsl::record_verifier(s, in, multipliers).release();
    } catch (const std::exception &e) { fprintf(stderr, "hs_circuit_record: %s\n", e.what()); return nullptr; }
}
void *hs_circuit_record_last(const u32 *shape) {
    try {
        stwo_b200::dsl::ProofShape s{shape[0], shape[1], shape[2], shape[3], shape[4], shape[5], shape[6]};
        return stwo_b200::dsl::record_last_layer(s).release();
    } catch (const std::exception &e) { fprintf(stderr, "hs_circuit_record_last: %s\n", e.what()); return nullptr; }
}
void *hs_circuit_record_folding(const u32 *shape) {
    try {
        stwo_b200::dsl::ProofShape s{shape[0], shape[1], shape[2], shape[3], shape[4], shape[5], shape[6]};
        return stwo_b200::dsl::record_folding(s).release();
    } catch (const std::exception &e) { fprintf(stderr, "hs_circuit_record_folding: %s\n", e.what()); return nullptr; }
}
void hs_circuit_free(void *h) { delete (RecordedCircuit *)h; }
// levels / bundles / instructions of the second order (0 when the circuit has none)
void hs_circuit_recorded_order(void *h, u32 *out3) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    out3[0] = r->recorded.n_levels(); out3[1] = r->recorded.bundle_start.empty() ? 0u : (u32)r->recorded.bundle_start.size() - 1; out3[2] = (u32)r->recorded.ins.size();
}
// info: n_rows, n_rows_unpadded, n_vars, n_flow, n_flow_padded, n_input_words, n_ins, n_levels, num_input, words_per_instance
void hs_circuit_info(void *h, u32 *info) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    const auto &c = *r->cs.p;
    const u32 v[10] = {c.num_plonk_rows(), c.n_rows_unpadded, c.n_vars, c.num_poseidon_invocations(), c.padded_poseidon_len(), c.n_input_words,
                       (u32)r->ins.size(), r->n_levels(), c.num_input, r->words_per_instance};
    memcpy(info, v, sizeof v);
}
// what: 0 a_wire 1 b_wire 2 c_wire 3 poseidon_wire 4 enforce_c_m31 5 op 6 op_follows_c 7 flow_wire 8 flow_swap_addr 9 gather 10 level_start
void hs_circuit_get(void *h, u32 what, u32 *out) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    const auto &c = *r->cs.p;
    const std::vector<u32> *src[] = {&c.a_wire, &c.b_wire, &c.c_wire, &c.poseidon_wire, &c.enforce_c_m31, &c.op, nullptr, &c.flow_wire,
                                     &c.flow_swap_addr, &r->gather, &r->level_start, &c.op2, &c.op3, &c.op4};
    if (what == 6) { for (size_t k = 0; k < c.op_follows_c.size(); k++) out[k] = c.op_follows_c[k]; return; }
    memcpy(out, src[what]->data(), src[what]->size() * 4);
}
// evaluates the tape for every proof of a verified batch (ws_base from hs_verify_batch) in level order, lanes = 1.
// vars: n x n_vars x 4, flow_hash: n x n_flow x 32, flow_swap: n x n_flow, stream (optional): n x n_input_words.
// bad_row[p] = first row failing check_arithmetics or -1.  use_hints: permutation outputs from the native pass (Perm::hint).
void hs_circuit_eval(void *h, void *ws_base, u32 n, u32 *vars, u32 *flow_hash, uint8_t *flow_swap, u32 *stream_out, int64_t *bad_row, int use_hints) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    const auto &c = *r->cs.p;
    const verify::Workspace &ws = *(const verify::Workspace *)ws_base;
    std::vector<u32> stream(c.n_input_words), extra(r->n_extra_words + 1);
    for (u32 p = 0; p < n; p++) {
        for (const auto &j : r->jobs) {
            circuit::ExtraJob job{j.kind, j.tree, j.query, j.col_off, j.n_cols, j.slot};
            circuit::extra_job(ws, p, job, extra.data());
        }
        for (u32 k = 0; k < c.n_input_words; k++) stream[k] = circuit::gather_word(ws, p, r->gather[k], extra.data());
        if (stream_out) memcpy(stream_out + (size_t)p * c.n_input_words, stream.data(), stream.size() * 4);
        tape::View v{(tape::Q4 *)(vars + (size_t)p * c.n_vars * 4), stream.data(), (tape::Q4 *)(flow_hash + (size_t)p * c.num_poseidon_invocations() * 32),
                     flow_swap + (size_t)p * c.num_poseidon_invocations(), 1,
                     use_hints && !c.without() && ws.hint_trees[p] == ws.shape.n_trees() ? ws.perm_out_of(p, 0) : nullptr};
        tape::prologue(v);
        // use_hints == 2: the second order of the tape (recorded permutations split into outputs-from-the-record + flow entry), which the
        // device picks for a lane group whose records are all complete
        const bool second = use_hints == 2 && v.hint && r->recorded.n_levels();
        for (const tape::Ins &in : (second ? r->recorded.ins : r->ins)) tape::eval(v, in, c.perms.data(), c.eperms.data());
        bad_row[p] = -1;
        for (u32 i = 0; i < c.num_plonk_rows(); i++) {
            bool ok;
            if (c.without()) {
                const qm31_t vc = tape::ldv(v, c.c_wire[i]);
                ok = tape::gate_ok_without(tape::ldv(v, c.a_wire[i]), tape::ldv(v, c.b_wire[i]), vc, c.op_follows_c[i] ? vc.v[0] : c.op[i], c.op2[i], c.op3[i], c.op4[i]);
            } else ok = tape::row_ok(v, c.a_wire[i], c.b_wire[i], c.c_wire[i], c.op[i], c.enforce_c_m31[i], c.op_follows_c[i] != 0);
            if (!ok) { bad_row[p] = i; break; }
        }
    }
}
}
// the recorded tape in recording order (n x 4 words) and the permutation records (n x 12 words): analysis / tests
extern "C" u32 hs_circuit_tape(void *h, u32 *ins_out, u32 *perms_out) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    const auto &c = *r->cs.p;
    if (ins_out) memcpy(ins_out, c.tape_.data(), c.tape_.size() * 16);
    if (perms_out) memcpy(perms_out, c.perms.data(), c.perms.size() * sizeof(tape::Perm));
    return (u32)c.tape_.size();
}
// check_arithmetics of one row: the whole gate and its two-lane split (tape::gate_ok_half, what the export kernel runs)
extern "C" int hs_gate_ok(const u32 *a, const u32 *b, const u32 *c, u32 op, u32 enforce) {
    return tape::gate_ok(qm31::mk(a[0], a[1], a[2], a[3]), qm31::mk(b[0], b[1], b[2], b[3]), qm31::mk(c[0], c[1], c[2], c[3]), op, enforce);
}
extern "C" int hs_gate_ok_half(const u32 *a, const u32 *b, const u32 *c, u32 op, u32 enforce, u32 part) {
    return tape::gate_ok_half(qm31::mk(a[0], a[1], a[2], a[3]), qm31::mk(b[0], b[1], b[2], b[3]), qm31::mk(c[0], c[1], c[2], c[3]), op, enforce, part);
}
// permutations of the recorded circuit that name a slot of the native verifier's record
extern "C" u32 hs_circuit_hint_count(void *h) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    u32 n = 0;
    for (const auto &p : r->cs.p->perms) n += p.hint != 0;
    return n;
}
extern "C" void hs_circuit_level_perms(void *h, u32 *perms_per_level) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    for (u32 l = 0; l < r->n_levels(); l++) {
        u32 n = 0;
        for (u32 k = r->level_start[l]; k < r->level_start[l + 1]; k++) n += r->ins[k].op == tape::T_POSEIDON;
        perms_per_level[l] = n;
    }
}

// ---- synthetic FRI + Merkle instances (synth.cuh): the generator and the verifier stages on the host ---------------------------------------
#include "../../recursive-stwo_b200/csrc/synth.cuh"
extern "C" {
u32 hs_synth_blob_words(const u32 *shape7) { verify::Shape s; memcpy(&s, shape7, 28); return synth::layout(s).total; }
// builds instance `seed` of `shape` into blob (hs_synth_blob_words words).  Returns 0, or a negative code when the generator's own checks fail
// (-1: the final layer is not a polynomial of the claimed degree; -2: no nonce found)
int hs_synth_generate(const u32 *shape7, u64 seed, u32 *blob) {
    verify::Shape s; memcpy(&s, shape7, 28);
    const synth::Layout l = synth::layout(s);
    synth::write_header(blob, s, l);
    for (u32 k = synth::HDR; k < l.total; k++) blob[k] = 0;
    static proof::Desc d; memset(&d, 0, sizeof d);
    synth::set_offsets(s, d);
    static fs::Out o;
    u32 logs[fri::MAX_LOGS];
    const u32 n_logs = fri::log_sizes(d, logs), nq = s.n_queries, mf = s.max_first();
    std::vector<std::vector<qm31_t>> cols(n_logs);
    for (u32 g = 0; g < n_logs; g++) { cols[g].resize((size_t)1 << logs[g]); for (u32 i = 0; i < (1u << logs[g]); i++) cols[g][i] = synth::column_value(seed, g, logs[g], s.log_blowup, i); }
    auto build_tree = [&](u32 depth, auto data_at) {            // data_at(h) -> const std::vector<qm31_t>* or nullptr
        std::vector<std::vector<u32>> lv(depth + 1);
        for (u32 h = depth + 1; h-- > 0;) {
            lv[h].resize((size_t)8 << h);
            const std::vector<qm31_t> *dat = data_at(h);
            for (u32 i = 0; i < (1u << h); i++)
                synth::node_hash(h == depth ? nullptr : &lv[h + 1][16 * (size_t)i], h == depth ? nullptr : &lv[h + 1][16 * (size_t)i + 8], dat ? (*dat)[i].v : nullptr, &lv[h][8 * (size_t)i]);
        }
        return lv;
    };
    auto first = build_tree(mf, [&](u32 h) -> const std::vector<qm31_t> * { for (u32 g = 0; g < n_logs; g++) if (logs[g] == h) return &cols[g]; return nullptr; });
    memcpy(blob + l.off_flc, first[0].data(), 32);
    synth::transcript(blob, d, o, 1);
    std::vector<qm31_t> cur((size_t)1 << (mf - 1), qm31::zero());
    std::vector<std::vector<qm31_t>> layers(s.n_inner);
    std::vector<std::vector<std::vector<u32>>> trees(s.n_inner);
    for (u32 li = 0; li < s.n_inner; li++) {
        for (u32 g = 0; g < n_logs; g++)
            if (logs[g] == mf - li) {
                const qm31_t a2 = qm31::mul(o.fri_alphas[li], o.fri_alphas[li]);
                for (u32 j = 0; j < cur.size(); j++)
                    cur[j] = qm31::add(qm31::mul(a2, cur[j]), synth::circle_fold_at(logs[g], j, cols[g][2 * j], cols[g][2 * j + 1], o.fri_alphas[li]));
            }
        const u32 L = mf - 1 - li;
        layers[li] = cur;
        trees[li] = build_tree(L, [&](u32 h) -> const std::vector<qm31_t> * { return h == L ? &layers[li] : nullptr; });
        memcpy(blob + l.off_inc[li], trees[li][0].data(), 32);
        synth::transcript(blob, d, o, 2 + li);
        std::vector<qm31_t> nxt(cur.size() / 2);
        for (u32 j = 0; j < nxt.size(); j++) nxt[j] = synth::line_fold_at(L, j, cur[2 * j], cur[2 * j + 1], o.fri_alphas[li + 1]);
        cur.swap(nxt);
    }
    const u32 Llast = mf - 1 - s.n_inner, nc = 1u << s.log_last;
    {
        std::vector<qm31_t> v(cur.begin(), cur.begin() + nc), out(nc), tmp(nc);
        synth::interpolate_line(Llast, s.log_last, v.data(), out.data(), tmp.data());
        for (u32 k = 0; k < nc; k++) memcpy(blob + l.off_last + 4 * k, out[k].v, 16);
        std::vector<qm31_t> buf(nc);
        for (u32 r = 0; r < cur.size(); r++)      // the verifier's evaluation point of position r of the last layer
            if (!qm31::eq(fri::eval_last_poly(blob + l.off_last, s.log_last, circle::dbl(fri::absolute_point(Llast + 1, 2 * r)).x, buf.data()), cur[r])) return -1;
    }
    bool found = false;
    for (u64 nonce = 0; nonce < (1ull << 20) && !found; nonce++) {
        blob[synth::H_NONCE] = (u32)nonce; blob[synth::H_NONCE + 1] = 0;
        synth::transcript(blob, d, o, 1 + s.n_inner);
        found = o.pow_ok && synth::queries_distinct(d, o);
    }
    if (!found) return -2;
    std::vector<u32> pos(nq), sp(nq), scratch(6 * nq + 8);
    u32 wi = 0;
    for (u32 g = 0; g < n_logs; g++) {
        for (u32 i = 0; i < nq; i++) { sp[i] = pos[i] = fri::position(d, o.raw_queries[i], logs[g]); memcpy(blob + l.off_ans + (g * nq + i) * 4, cols[g][pos[i]].v, 16); }
        const u32 ns = decommit::sort_unique(sp.data(), nq);
        for (u32 k = 0; k < ns;) {
            const u32 start = sp[k] & ~1u;
            for (u32 e = start; e < start + 2; e++) {
                if (k < ns && sp[k] == e) k++;
                else { memcpy(blob + l.off_flfw + 4 * wi, cols[g][e].v, 16); wi++; }
            }
        }
    }
    blob[synth::H_FL_NFW] = wi;
    for (u32 i = 0; i < nq; i++) pos[i] = fri::position(d, o.raw_queries[i], mf);
    blob[synth::H_FL_NHW] = synth::emit_hash_witness(mf, s.fri_data_mask(0), pos.data(), nq, [&](u32 h, u32 p) { return &first[h][8 * (size_t)p]; }, blob + l.off_flhw, scratch.data());
    for (u32 li = 0; li < s.n_inner; li++) {
        const u32 L = mf - 1 - li;
        for (u32 i = 0; i < nq; i++) sp[i] = pos[i] = fri::position(d, o.raw_queries[i], L);
        const u32 ns = decommit::sort_unique(sp.data(), nq);
        u32 w2 = 0;
        for (u32 k = 0; k < ns; k++)
            if (decommit::find(sp.data(), ns, sp[k] ^ 1u) < 0) { memcpy(blob + l.off_infw[li] + 4 * w2, layers[li][sp[k] ^ 1u].v, 16); w2++; }
        blob[synth::H_IN_NFW + li] = w2;
        blob[synth::H_IN_NHW + li] = synth::emit_hash_witness(L, 1u << L, pos.data(), nq, [&](u32 h, u32 p) { return &trees[li][h][8 * (size_t)p]; }, blob + l.off_inhw[li], scratch.data());
    }
    return 0;
}
// the product's verifier stages on synthetic instances (what stwo_b200_synth_verify_batch_dev launches): returns the workspace base (hs_free)
void *hs_synth_verify(const u32 *blobs, const u64 *blob_off, u32 n, const u32 *shape7, int coop, verify::Detail *detail_out, verify::Workspace *ws_out) {
    verify::Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, shape7, 28);
    ws.n_proofs = n; ws.blobs = blobs; ws.blob_off = blob_off;
    uint8_t *base = (uint8_t *)calloc(verify::carve(ws, nullptr), 1);
    verify::carve(ws, base);
    ws.mode = verify::MODE_FULL | (coop ? 0u : verify::MODE_PATH_KERNELS);
    const u32 nf = ws.shape.n_fri_trees(), nq = ws.shape.n_queries;
    std::vector<u32> tab(decommit::pair_tab_words(nq) + verify::folds_tab_words(nq) + 2 * nq + 64);
    decommit::CoopOne one;
    for (u32 p = 0; p < n; p++) synth::stage_open(ws, p);
    for (u32 p = 0; p < n; p++) { if (coop) verify::stage_folds_coop(one, ws, p, tab.data()); else verify::stage_folds(ws, p); }
    for (u32 p = 0; p < n; p++) for (u32 f = 0; f < nf; f++) { if (coop) verify::stage_pair_tree_coop(one, ws, p, f, tab.data()); else verify::stage_pair_tree(ws, p, f); }
    if (!coop) for (u32 p = 0; p < n; p++) for (u32 f = 0; f < nf; f++) for (u32 i = 0; i < nq; i++) verify::stage_pair_path(ws, p, f, i);
    for (u32 p = 0; p < n; p++) verify::stage_verdict(ws, p);
    memcpy(detail_out, ws.detail, n * sizeof(verify::Detail));
    if (ws_out) *ws_out = ws;
    return base;
}
}
// bundle statistics of the levelised tape (analysis): out[0] bundles, [1] instructions, [2] instructions that are not first in their bundle,
// [3] of those: an operand is the output of the instruction just before, [4] an operand is an output of any earlier instruction of the
// bundle, [5] permutations inside bundles of length > 1, [6] non-first instructions with BOTH variable operands from inside the bundle
extern "C" void hs_circuit_bundle_stats(void *h, u32 *out) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    const auto &c = *r->cs.p;
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[0] = (u32)r->bundle_start.size() - 1; out[1] = (u32)r->ins.size();
    for (size_t b = 0; b + 1 < r->bundle_start.size(); b++) {
        const u32 i0 = r->bundle_start[b], i1 = r->bundle_start[b + 1];
        std::vector<u32> outs;
        u32 prev_first = 0, prev_n = 0;
        for (u32 i = i0; i < i1; i++) {
            const tape::Ins &in = r->ins[i];
            if (i > i0) {
                out[2]++;
                bool from_prev = false, from_any = false, all_in = true; u32 ns = 0;
                RecordedCircuit::for_sources(c, in, [&](u32 v) {
                    ns++;
                    bool in_b = false;
                    for (size_t q = 0; q < outs.size(); q++) if (outs[q] == v) { in_b = true; if (q >= prev_first && q < prev_first + prev_n) from_prev = true; }
                    from_any |= in_b; all_in &= in_b;
                });
                out[3] += from_prev; out[4] += from_any; out[6] += (ns >= 2 && all_in);
                if (in.op == tape::T_POSEIDON && i1 - i0 > 1) out[5]++;
            } else if (in.op == tape::T_POSEIDON && i1 - i0 > 1) out[5]++;
            prev_first = (u32)outs.size(); prev_n = 0;
            RecordedCircuit::for_outputs(c, in, [&](u32 v) { outs.push_back(v); prev_n++; });
        }
    }
}
// per level (5 words each): bundles, instructions, permutations, most permutations in one bundle, most instructions in one bundle
extern "C" void hs_circuit_level_profile(void *h, u32 *out) {
    RecordedCircuit *r = (RecordedCircuit *)h;
    for (u32 l = 0; l < r->n_levels(); l++) {
        u32 nb = r->level_bundle[l + 1] - r->level_bundle[l], ni = r->level_start[l + 1] - r->level_start[l], np = 0, mp = 0, mi = 0;
        for (u32 b = r->level_bundle[l]; b < r->level_bundle[l + 1]; b++) {
            u32 p = 0;
            for (u32 i = r->bundle_start[b]; i < r->bundle_start[b + 1]; i++) p += r->ins[i].op == tape::T_POSEIDON;
            np += p; mp = std::max(mp, p); mi = std::max(mi, r->bundle_start[b + 1] - r->bundle_start[b]);
        }
        const u32 v[5] = {nb, ni, np, mp, mi};
        memcpy(out + 5 * l, v, sizeof v);
    }
}
