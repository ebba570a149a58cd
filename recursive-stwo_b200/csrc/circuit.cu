// The recursive verifier circuit on the device: C ABI of stwo_b200_circuit_* (record once per shape on the host, then per
// batch: gather the witness streams out of the verifier's workspace -> K6 tape evaluation -> K7 checks and trace export).
// Replaces the circuit half of examples/single-proof/src/main.rs:48-90 and examples/multi-proofs/src/main.rs:62-139.
#include "common.cuh"
#include "circuit.cuh"
#include "dsl/recorded.hpp"
#include <string.h>
#include <stdlib.h>

using namespace stwo_b200;
using dsl::RecordedCircuit;

struct stwo_b200_circuit {
    std::unique_ptr<RecordedCircuit> rec;
    // device image (uploaded on first use)
    uint8_t *dev = nullptr;
    stwo_b200_cs_wiring wiring{};
    stwo_b200_cs_tape tape_{};
    stwo_b200_cs_tape_order order2_{};     // the order for items with a complete permutation record (tape_.recorded_order points here)
    const u32 *gather = nullptr;
    const circuit::ExtraJob *jobs = nullptr;      // last-layer circuit only
    u32 n_jobs = 0, n_extra_words = 0;
    // populate_logup_arguments depends on the wiring only (plonk_with_poseidon.rs:382-466): computed once, at upload
    int32_t *mult = nullptr;                      // mult_a | mult_b | mult_c | mult_poseidon, n_rows each
    u32 *scratch = nullptr, *status = nullptr;    // counts | first_key | first_prow | mult_poseidon_vars; the assert flag
};

namespace {
constexpr int kT = 256;
inline unsigned nblk(size_t n) { return (unsigned)((n + kT - 1) / kT); }

// witness streams, lane-interleaved over groups of 32 proofs
// public-input hashes of the last-layer circuit: thread = (job, proof), the proof fastest; extra: [proof][n_extra_words]
__global__ void __launch_bounds__(64) k_last_extra(verify::Workspace ws, const circuit::ExtraJob *__restrict__ jobs, u32 n_jobs, u32 n_extra_words, u32 *extra) {
    const size_t g = blockIdx.x * (size_t)64 + threadIdx.x;
    if (g >= (size_t)n_jobs * ws.n_proofs) return;
    const u32 p = (u32)(g % ws.n_proofs), j = (u32)(g / ws.n_proofs);
    if (!ws.desc[p].ok) return;
    circuit::extra_job(ws, p, jobs[j], extra + (size_t)p * n_extra_words);
}
// Tile = 32 consecutive stream words x the 32 proofs of a lane group.  A warp reads ALONG the words of one proof -- consecutive
// stream words mostly come from consecutive source words (a hash is 8, a column run up to 60), so the loads are 32-240 B runs
// instead of one 4-byte word out of every proof's 32-byte sector -- and the tile is written back transposed, 128-byte lines of
// the lane-interleaved stream.
__global__ void __launch_bounds__(kT) k_gather_witness(verify::Workspace ws, const u32 *__restrict__ gather, u32 n_words, u32 *out, const u32 *extra,
                                                       u32 n_extra_words) {
    __shared__ u32 tile[32][33];
    const u32 warp = threadIdx.x / 32, lane = threadIdx.x % 32, n_warps = kT / 32;
    const u32 word0 = blockIdx.x * 32, grp = blockIdx.y;
    const u32 word = word0 + lane;
    const u32 src = word < n_words ? __ldg(gather + word) : 0u;
    for (u32 pl = warp; pl < 32; pl += n_warps) {
        const u32 p = grp * 32 + pl;
        u32 v = 0;
        // a blob that did not parse has no sections: its stream is all zero
        if (p < ws.n_proofs && word < n_words && ws.desc[p].ok) v = circuit::gather_word(ws, p, src, extra ? extra + (size_t)p * n_extra_words : nullptr);
        tile[lane][pl] = v;
    }
    __syncthreads();
    for (u32 wl = warp; wl < 32; wl += n_warps)
        if (word0 + wl < n_words && grp * 32 + lane < ws.n_proofs) out[((size_t)grp * n_words + word0 + wl) * 32 + lane] = tile[wl][lane];
}

struct Carve {
    u32 *witness, *vars, *flow_hash; uint8_t *flow_swap; u32 *status, *extra;
    size_t bytes;
};
Carve carve(const RecordedCircuit &r, u32 n_proofs, uint8_t *base) {
    const auto &c = *r.cs.p;
    const size_t groups = (n_proofs + 31) / 32;
    size_t at = 0;
    auto take = [&](size_t bytes) { at = align_up(at, 256); uint8_t *p = base ? base + at : nullptr; at += bytes; return p; };
    Carve k;
    k.witness = (u32 *)take(groups * c.n_input_words * 32 * 4);
    k.vars = (u32 *)take(groups * c.n_vars * 32 * 16);
    k.flow_hash = (u32 *)take(groups * c.num_poseidon_invocations() * 32 * 32 * 4);
    k.flow_swap = take(groups * c.num_poseidon_invocations() * 32);
    k.status = (u32 *)take(256);
    k.extra = (u32 *)take((size_t)n_proofs * r.n_extra_words * 4 + 4);
    k.bytes = align_up(at, 256);
    return k;
}

int32_t upload(stwo_b200_circuit *c) {
    if (c->dev) return STWO_B200_OK;
    const RecordedCircuit &r = *c->rec;
    const auto &cs = *r.cs.p;
    const size_t nr = cs.num_plonk_rows(), nf = cs.num_poseidon_invocations();
    size_t at = 0;
    auto take = [&](size_t bytes) { at = align_up(at, 256); size_t o = at; at += bytes; return o; };
    const size_t o_w = take(9 * nr * 4), o_jobs = take(r.jobs.size() * sizeof(circuit::ExtraJob) + 16), o_fol = take(nr), o_fw = take(nf * 16 + 16), o_fa = take(nf * 4 + 4), o_ins = take(r.ins.size() * 16 + 16),
                 o_lvl = take(r.level_start.size() * 4), o_bs = take(r.bundle_start.size() * 4 + 4), o_lb = take(r.level_bundle.size() * 4 + 4), o_perm = take(cs.perms.size() * sizeof(tape::Perm) + 16), o_g = take(r.gather.size() * 4 + 4), o_ep = take(cs.eperms.size() * 4 + 4),
                 o_xt = take(stwo_b200_cs_export_tiles_words((u32)nr) * 4 + 16), o_mult = take((size_t)4 * nr * 4), o_scr = take(((size_t)4 * cs.n_vars + 4) * 4),
                 o_stat = take(256),
                 o_ins2 = take(r.recorded.ins.size() * 16 + 16), o_bs2 = take(r.recorded.bundle_start.size() * 4 + 4), o_lb2 = take(r.recorded.level_bundle.size() * 4 + 4);
    uint8_t *d = nullptr;
    STWO_CUDA(cudaMalloc(&d, at));
    const std::vector<u32> *cols[9] = {&cs.a_wire, &cs.b_wire, &cs.c_wire, &cs.poseidon_wire, &cs.enforce_c_m31, &cs.op, &cs.op2, &cs.op3, &cs.op4};
    for (int k = 0; k < 9; k++) STWO_CUDA(cudaMemcpy(d + o_w + k * nr * 4, cols[k]->data(), nr * 4, cudaMemcpyHostToDevice));
    static_assert(sizeof(circuit::ExtraJob) == sizeof(dsl::ExtraHashJob), "job records mirror");
    if (!r.jobs.empty()) STWO_CUDA(cudaMemcpy(d + o_jobs, r.jobs.data(), r.jobs.size() * sizeof(circuit::ExtraJob), cudaMemcpyHostToDevice));
    STWO_CUDA(cudaMemcpy(d + o_fol, cs.op_follows_c.data(), nr, cudaMemcpyHostToDevice));
    if (nf) {
        STWO_CUDA(cudaMemcpy(d + o_fw, cs.flow_wire.data(), nf * 16, cudaMemcpyHostToDevice));
        STWO_CUDA(cudaMemcpy(d + o_fa, cs.flow_swap_addr.data(), nf * 4, cudaMemcpyHostToDevice));
        STWO_CUDA(cudaMemcpy(d + o_perm, cs.perms.data(), cs.perms.size() * sizeof(tape::Perm), cudaMemcpyHostToDevice));
    }
    STWO_CUDA(cudaMemcpy(d + o_ins, r.ins.data(), r.ins.size() * 16, cudaMemcpyHostToDevice));
    STWO_CUDA(cudaMemcpy(d + o_lvl, r.level_start.data(), r.level_start.size() * 4, cudaMemcpyHostToDevice));
    STWO_CUDA(cudaMemcpy(d + o_bs, r.bundle_start.data(), r.bundle_start.size() * 4, cudaMemcpyHostToDevice));
    STWO_CUDA(cudaMemcpy(d + o_lb, r.level_bundle.data(), r.level_bundle.size() * 4, cudaMemcpyHostToDevice));
    if (!cs.eperms.empty()) STWO_CUDA(cudaMemcpy(d + o_ep, cs.eperms.data(), cs.eperms.size() * 4, cudaMemcpyHostToDevice));
    if (r.recorded.n_levels()) {
        STWO_CUDA(cudaMemcpy(d + o_ins2, r.recorded.ins.data(), r.recorded.ins.size() * 16, cudaMemcpyHostToDevice));
        STWO_CUDA(cudaMemcpy(d + o_bs2, r.recorded.bundle_start.data(), r.recorded.bundle_start.size() * 4, cudaMemcpyHostToDevice));
        STWO_CUDA(cudaMemcpy(d + o_lb2, r.recorded.level_bundle.data(), r.recorded.level_bundle.size() * 4, cudaMemcpyHostToDevice));
    }
    if (!r.gather.empty()) STWO_CUDA(cudaMemcpy(d + o_g, r.gather.data(), r.gather.size() * 4, cudaMemcpyHostToDevice));
    const u32 *w = (const u32 *)(d + o_w);
    // per 32-row tile of the export pass: the distinct variables its wires name
    std::vector<u32> xt(stwo_b200_cs_export_tiles_words((u32)nr));
    u32 xcap = 0;
    {
        stwo_b200_cs_wiring hw{};
        hw.n_rows = (u32)nr; hw.kind = cs.without() ? 1u : 0u;
        hw.a_wire = cs.a_wire.data(); hw.b_wire = cs.b_wire.data(); hw.c_wire = cs.c_wire.data(); hw.op = cs.op.data();
        hw.enforce_c_m31 = cs.enforce_c_m31.data(); hw.op2 = cs.op2.data(); hw.op3 = cs.op3.data(); hw.op4 = cs.op4.data();
        hw.op_follows_c = cs.op_follows_c.data();
        if (xt.empty() || stwo_b200_cs_export_tiles_build(&hw, xt.data(), &xcap) != STWO_B200_OK) { xt.clear(); xcap = 0; }    // the tiled export remains
        else STWO_CUDA(cudaMemcpy(d + o_xt, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice));
    }
    c->wiring = {cs.n_vars, (u32)nr, (u32)nf, cs.num_input, w, w + nr, w + 2 * nr, w + 3 * nr, w + 4 * nr, w + 5 * nr, d + o_fol,
                 (const u32 *)(d + o_fw), (const u32 *)(d + o_fa), cs.without() ? 1u : 0u, w + 6 * nr, w + 7 * nr, w + 8 * nr,
                 xt.empty() ? nullptr : (const u32 *)(d + o_xt), xcap};
    c->jobs = (const circuit::ExtraJob *)(d + o_jobs); c->n_jobs = (u32)r.jobs.size(); c->n_extra_words = r.n_extra_words;
    c->tape_ = {(u32)r.ins.size(), (u32)cs.perms.size(), r.n_levels(), cs.n_input_words, (const u32 *)(d + o_ins), (const u32 *)(d + o_lvl),
                (const u32 *)(d + o_perm), (u32)(cs.eperms.size() / tape::EPOSEIDON_REC), (const u32 *)(d + o_ep),
                (u32)r.bundle_start.size() - 1, (const u32 *)(d + o_bs), (const u32 *)(d + o_lb), nullptr};
    if (r.recorded.n_levels()) {
        c->order2_ = {(u32)r.recorded.ins.size(), r.recorded.n_levels(), (u32)r.recorded.bundle_start.size() - 1, (const u32 *)(d + o_ins2),
                      (const u32 *)(d + o_bs2), (const u32 *)(d + o_lb2)};
        c->tape_.recorded_order = &c->order2_;
    }
    c->gather = (const u32 *)(d + o_g);
    c->mult = (int32_t *)(d + o_mult); c->scratch = (u32 *)(d + o_scr); c->status = (u32 *)(d + o_stat);
    {
        cudaStream_t st;
        STWO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        const int32_t rc = stwo_b200_cs_populate_logup_dev(&c->wiring, c->mult, c->mult + nr, c->mult + 2 * nr, c->mult + 3 * nr, c->scratch, c->status, st);
        const cudaError_t e = cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        if (rc) return rc;
        if (e != cudaSuccess) return -(int32_t)e;
    }
    c->dev = d;
    return STWO_B200_OK;
}

cudaEvent_t g_ev[STWO_B200_N_TRACE_STAGES + 1] = {nullptr};
cudaStream_t g_side = nullptr;                 // runs check_poseidon_invocations beside the export
cudaEvent_t g_side_fork = nullptr, g_side_join = nullptr;
// CTAs per SM of the check when it runs beside the export (STWO_B200_BESIDE_CTAS=k); 0 = one after the other, the default: measured
// on B200 at 4096 proofs the pair takes 7.0-8.7 ms beside each other (k = 2..5, run-to-run unstable) against 8.2 ms in sequence
int g_beside_ctas = [] { const char *e = getenv("STWO_B200_BESIDE_CTAS"); return e ? atoi(e) : 0; }();
bool g_timed_valid = false;
}  // namespace

void stwo_b200::circuit_streams_destroy() {
    if (g_side) cudaStreamDestroy(g_side);
    if (g_side_fork) cudaEventDestroy(g_side_fork);
    if (g_side_join) cudaEventDestroy(g_side_join);
    g_side = nullptr; g_side_fork = g_side_join = nullptr;
    for (auto &e : g_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    g_timed_valid = false;
}
static_assert(sizeof(dsl::ProofShape) == sizeof(stwo_b200_proof_shape), "shape mirrors");

extern "C" int32_t stwo_b200_circuit_record_verifier(const stwo_b200_proof_shape *shape, const uint32_t *input_idx, const uint32_t *input_vals,
                                                     uint32_t n_inputs, uint32_t multipliers, stwo_b200_circuit **out) {
    if (!shape || !out || multipliers == 0 || (n_inputs && (!input_idx || !input_vals))) return STWO_B200_E_BAD_ARG;
    if (!proof::shape_consistent(shape->log_size_plonk, shape->log_size_poseidon, shape->pow_bits, shape->log_blowup, shape->log_last,
                                 shape->n_queries, shape->n_inner))
        return STWO_B200_E_SHAPE;
    try {
        dsl::ProofShape s;
        memcpy(&s, shape, sizeof s);
        std::vector<dsl::PublicInput> in;
        for (u32 k = 0; k < n_inputs; k++)
            in.push_back({input_idx[k], {{input_vals[4 * k], input_vals[4 * k + 1], input_vals[4 * k + 2], input_vals[4 * k + 3]}}});
        stwo_b200_circuit *c = new stwo_b200_circuit();
        c->rec = dsl::record_verifier(s, in, multipliers);
        *out = c;
        return STWO_B200_OK;
    } catch (const std::exception &) { return STWO_B200_E_SHAPE; }
}
extern "C" int32_t stwo_b200_circuit_record_last_layer(const stwo_b200_proof_shape *shape, stwo_b200_circuit **out) {
    if (!shape || !out) return STWO_B200_E_BAD_ARG;
    if (!proof::shape_consistent(shape->log_size_plonk, shape->log_size_poseidon, shape->pow_bits, shape->log_blowup, shape->log_last,
                                 shape->n_queries, shape->n_inner))
        return STWO_B200_E_SHAPE;
    try {
        dsl::ProofShape s;
        memcpy(&s, shape, sizeof s);
        stwo_b200_circuit *c = new stwo_b200_circuit();
        c->rec = dsl::record_last_layer(s);
        *out = c;
        return STWO_B200_OK;
    } catch (const std::exception &) { return STWO_B200_E_SHAPE; }
}
extern "C" int32_t stwo_b200_circuit_record_folding(const stwo_b200_proof_shape *shape, stwo_b200_circuit **out) {
    if (!shape || !out) return STWO_B200_E_BAD_ARG;
    if (!proof::shape_consistent(shape->log_size_plonk, shape->log_size_poseidon, shape->pow_bits, shape->log_blowup, shape->log_last,
                                 shape->n_queries, shape->n_inner))
        return STWO_B200_E_SHAPE;
    try {
        dsl::ProofShape s;
        memcpy(&s, shape, sizeof s);
        stwo_b200_circuit *c = new stwo_b200_circuit();
        c->rec = dsl::record_folding(s);
        *out = c;
        return STWO_B200_OK;
    } catch (const std::exception &) { return STWO_B200_E_SHAPE; }
}
extern "C" void stwo_b200_circuit_free(stwo_b200_circuit *c) {
    if (!c) return;
    if (c->dev) cudaFree(c->dev);
    delete c;
}
extern "C" int32_t stwo_b200_circuit_get_info(const stwo_b200_circuit *c, stwo_b200_circuit_info *out) {
    if (!c || !out) return STWO_B200_E_BAD_ARG;
    const auto &cs = *c->rec->cs.p;
    *out = {cs.num_plonk_rows(), cs.n_rows_unpadded, cs.n_vars, cs.num_poseidon_invocations(), cs.padded_poseidon_len(), cs.n_input_words,
            (u32)c->rec->ins.size(), c->rec->n_levels(), cs.num_input, c->rec->words_per_instance, cs.without() ? 1u : 0u, cs.without() ? 8u : 10u};
    return STWO_B200_OK;
}
extern "C" int32_t stwo_b200_circuit_get_column(const stwo_b200_circuit *c, uint32_t what, uint32_t *out, size_t n_words) {
    if (!c || !out) return STWO_B200_E_BAD_ARG;
    const auto &cs = *c->rec->cs.p;
    const std::vector<u32> *src = nullptr;
    switch (what) {
        case STWO_B200_COL_A_WIRE: src = &cs.a_wire; break;
        case STWO_B200_COL_B_WIRE: src = &cs.b_wire; break;
        case STWO_B200_COL_C_WIRE: src = &cs.c_wire; break;
        case STWO_B200_COL_POSEIDON_WIRE: src = &cs.poseidon_wire; break;
        case STWO_B200_COL_ENFORCE_C_M31: src = &cs.enforce_c_m31; break;
        case STWO_B200_COL_OP: src = &cs.op; break;
        case STWO_B200_COL_OP2: src = &cs.op2; break;
        case STWO_B200_COL_OP3: src = &cs.op3; break;
        case STWO_B200_COL_OP4: src = &cs.op4; break;
        case STWO_B200_COL_FLOW_WIRE: src = &cs.flow_wire; break;
        case STWO_B200_COL_FLOW_SWAP_ADDR: src = &cs.flow_swap_addr; break;
        case STWO_B200_COL_LEVEL_START: src = &c->rec->level_start; break;
        case STWO_B200_COL_OP_FOLLOWS_C:
            if (n_words != cs.op_follows_c.size()) return STWO_B200_E_BAD_ARG;
            for (size_t k = 0; k < n_words; k++) out[k] = cs.op_follows_c[k];
            return STWO_B200_OK;
        default: return STWO_B200_E_BAD_ARG;
    }
    if (n_words != src->size()) return STWO_B200_E_BAD_ARG;
    memcpy(out, src->data(), n_words * 4);
    return STWO_B200_OK;
}
extern "C" size_t stwo_b200_circuit_workspace_bytes(const stwo_b200_circuit *c, uint32_t n_proofs) {
    if (!c || !n_proofs) return 0;
    return carve(*c->rec, n_proofs, nullptr).bytes;
}

extern "C" int32_t stwo_b200_circuit_trace_batch_dev(stwo_b200_circuit *c, const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                                     const void *verify_workspace, void *circuit_workspace, size_t circuit_workspace_bytes,
                                                     uint32_t flags, uint32_t *preprocessed, uint32_t *values, int64_t *bad_row, int64_t *bad_flow,
                                                     void *stream) {
    STWO_CHECK_DEVICE();
    if (!c || !blobs || !blob_off || !verify_workspace || !circuit_workspace || !n_proofs) return STWO_B200_E_BAD_ARG;
    if ((flags & STWO_B200_TRACE_CHECK_ARITHMETICS) && !bad_row) return STWO_B200_E_BAD_ARG;
    if ((flags & STWO_B200_TRACE_CHECK_POSEIDON) && !bad_flow) return STWO_B200_E_BAD_ARG;
    const RecordedCircuit &r = *c->rec;
    const Carve k = carve(r, n_proofs, (uint8_t *)circuit_workspace);
    if (k.bytes > circuit_workspace_bytes) return STWO_B200_E_BAD_ARG;
    int32_t rc = upload(c);
    if (rc) return rc;
    verify::Workspace ws;
    memset(&ws, 0, sizeof ws);
    memcpy(&ws.shape, &r.shape, sizeof ws.shape);
    ws.n_proofs = n_proofs; ws.blobs = blobs; ws.blob_off = blob_off;
    verify::carve(ws, (uint8_t *)const_cast<void *>(verify_workspace));
    cudaStream_t st = (cudaStream_t)stream;
    const bool timed = flags & STWO_B200_TRACE_TIMED;
    g_timed_valid = false;
    if (timed && !g_ev[0]) for (auto &e : g_ev) STWO_CUDA(cudaEventCreate(&e));
    int e = 0;
#define MARK() do { if (timed) cudaEventRecord(g_ev[e], st); e++; } while (0)
    const u32 nw = c->tape_.n_input_words;
    const size_t groups = (n_proofs + 31) / 32;
    MARK();
    if (c->n_jobs) {
        k_last_extra<<<(unsigned)(((size_t)c->n_jobs * n_proofs + 63) / 64), 64, 0, st>>>(ws, c->jobs, c->n_jobs, c->n_extra_words, k.extra);
        note_launch(1);
    }
    k_gather_witness<<<dim3((nw + 31) / 32, (unsigned)groups), kT, 0, st>>>(ws, c->gather, nw, k.witness, c->n_jobs ? k.extra : nullptr, c->n_extra_words);
    note_launch(1);
    stwo_b200_cs_values v = {n_proofs, 32, k.vars, k.flow_hash, k.flow_swap, nullptr, 0, nullptr, 0, nullptr};
    // per proof: only a record that the verification of THIS workspace just completed is used (verify::Workspace::hint_trees)
    if ((flags & STWO_B200_TRACE_NATIVE_HINTS) && c->wiring.kind == 0 && ws.perm_out) {
        v.perm_hints = ws.perm_out; v.perm_hint_stride = ws.hint_total * 16;
        v.perm_hint_ready = ws.hint_trees; v.perm_hint_need = ws.shape.n_trees();
        // the record-based check_poseidon_invocations (entry == executed permutation) unless the caller asks for the re-execution
        if (!(flags & STWO_B200_TRACE_RECHECK_POSEIDON) && stwo_b200::record_inputs()) v.perm_hint_inputs = ws.perm_in;
    }
    MARK();
    if ((rc = stwo_b200_cs_eval_tape_dev(&c->tape_, c->wiring.n_vars, k.witness, &v, st))) return rc;
    MARK();
    // check_arithmetics rides on the export pass when the value columns are written anyway
    const bool fuse_check = (flags & STWO_B200_TRACE_CHECK_ARITHMETICS) && values;
    if ((flags & STWO_B200_TRACE_CHECK_ARITHMETICS) && !fuse_check)
        if ((rc = stwo_b200_cs_check_arithmetics_dev(&c->wiring, &v, bad_row, st))) return rc;
    MARK();
    const size_t nr = c->wiring.n_rows;
    int32_t *const mult = c->mult;                 // populate_logup_arguments: wiring only, computed at upload
    if ((flags & STWO_B200_TRACE_CHECK_POSEIDON) && c->wiring.kind == 1)      // unimplemented!() for this system: nothing to check
        if ((rc = cuda_status(cudaMemsetAsync(bad_flow, 0xff, (size_t)n_proofs * 8, st)))) return rc;
    // check_poseidon_invocations is integer-issue bound (it re-executes every flow permutation), the export HBM bound, and both
    // only read variables[] and the flow. Outside timed mode the check runs as a thin persistent layer (a few CTAs per SM, all
    // resident from the start) on a high-priority side stream, and the export's CTAs fill the rest of every SM beside it.
    const bool do_check_poseidon = (flags & STWO_B200_TRACE_CHECK_POSEIDON) && c->wiring.kind == 0;
    const bool beside = do_check_poseidon && !timed && (preprocessed || values) && g_beside_ctas > 0;
    if (beside) {
        if (!g_side) {
            int lo = 0, hi = 0;
            STWO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            STWO_CUDA(cudaStreamCreateWithPriority(&g_side, cudaStreamNonBlocking, hi));
            STWO_CUDA(cudaEventCreateWithFlags(&g_side_fork, cudaEventDisableTiming));
            STWO_CUDA(cudaEventCreateWithFlags(&g_side_join, cudaEventDisableTiming));
        }
        STWO_CUDA(cudaEventRecord(g_side_fork, st));
        STWO_CUDA(cudaStreamWaitEvent(g_side, g_side_fork, 0));
        if ((rc = cs_check_poseidon_launch(&c->wiring, &v, mult + 3 * nr, c->scratch, bad_flow, g_side, g_beside_ctas, &c->tape_))) return rc;
        STWO_CUDA(cudaEventRecord(g_side_join, g_side));
    } else if (do_check_poseidon)
        if ((rc = cs_check_poseidon_launch(&c->wiring, &v, mult + 3 * nr, c->scratch, bad_flow, st, 0, &c->tape_))) return rc;
    MARK();
    if (preprocessed || values)
        if ((rc = stwo_b200_cs_export_trace_dev(&c->wiring, &v, mult, mult + nr, mult + 2 * nr, mult + 3 * nr, preprocessed, values,
                                                fuse_check ? bad_row : nullptr, st))) return rc;
    if (beside) STWO_CUDA(cudaStreamWaitEvent(st, g_side_join, 0));
    MARK();
#undef MARK
    g_timed_valid = timed;
    return cuda_status(cudaGetLastError());
}

extern "C" int32_t stwo_b200_circuit_stage_ms(float *ms) {
    STWO_CHECK_DEVICE();
    if (!ms || !g_timed_valid) return STWO_B200_E_BAD_ARG;
    STWO_CUDA(cudaEventSynchronize(g_ev[STWO_B200_N_TRACE_STAGES]));
    for (int i = 0; i < STWO_B200_N_TRACE_STAGES; i++) STWO_CUDA(cudaEventElapsedTime(&ms[i], g_ev[i], g_ev[i + 1]));
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_circuit_fetch(const stwo_b200_circuit *c, const void *circuit_workspace, uint32_t n_proofs, uint32_t p, uint32_t what,
                                           void *out, size_t out_bytes, void *stream) {
    STWO_CHECK_DEVICE();
    if (!c || !circuit_workspace || p >= n_proofs || !out) return STWO_B200_E_BAD_ARG;
    const auto &cs = *c->rec->cs.p;
    const Carve k = carve(*c->rec, n_proofs, (uint8_t *)const_cast<void *>(circuit_workspace));
    const size_t grp = p / 32, lane = p % 32;
    const uint8_t *src; size_t elem, n;
    switch (what) {
        case STWO_B200_CFETCH_VARIABLES: elem = 16; n = cs.n_vars; src = (const uint8_t *)k.vars; break;
        case STWO_B200_CFETCH_FLOW_HASH: elem = 16; n = (size_t)cs.num_poseidon_invocations() * 8; src = (const uint8_t *)k.flow_hash; break;   // 16-byte elements
        case STWO_B200_CFETCH_FLOW_SWAP: elem = 1; n = cs.num_poseidon_invocations(); src = k.flow_swap; break;
        case STWO_B200_CFETCH_WITNESS: elem = 4; n = cs.n_input_words; src = (const uint8_t *)k.witness; break;
        default: return STWO_B200_E_BAD_ARG;
    }
    if (out_bytes < n * elem) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // de-interleave one lane: a strided 2-D copy
    STWO_CUDA(cudaMemcpy2DAsync(out, elem, src + (grp * n * 32 + lane) * elem, 32 * elem, elem, n, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

// The prover-facing PoseidonFlow of a traced batch: what generate_plonk_with_poseidon_circuit hands to stwo's
// prove_plonk_with_poseidon besides the trace (plonk_with_poseidon.rs:117-128,521-628), padded like pad() (:296-321) with the
// caller's CONSTANT_1/2/3 (their values live in the absent stwo dependency: `permute(C1 || C1) = C2 || C3`, :496-517).
// flow_hash_out: n_proofs x n_flow_padded x 32 words, flow_swap_out: n_proofs x n_flow_padded bytes.  The entries' wires and swap
// addresses are wiring: stwo_b200_circuit_get_column(FLOW_WIRE / FLOW_SWAP_ADDR), zero for the padding.
extern "C" int32_t stwo_b200_circuit_export_flow_dev(stwo_b200_circuit *c, uint32_t n_proofs, void *circuit_workspace, size_t circuit_workspace_bytes,
                                                     const uint32_t *pad_constants, uint32_t *flow_hash_out, uint8_t *flow_swap_out, void *stream) {
    STWO_CHECK_DEVICE();
    if (!c || !n_proofs || !circuit_workspace || !pad_constants || !flow_hash_out || !flow_swap_out) return STWO_B200_E_BAD_ARG;
    const RecordedCircuit &r = *c->rec;
    if (r.cs.p->without()) return STWO_B200_E_BAD_ARG;              // the Plonk-without-Poseidon system has no flow
    const Carve k = carve(r, n_proofs, (uint8_t *)circuit_workspace);
    if (k.bytes > circuit_workspace_bytes) return STWO_B200_E_BAD_ARG;
    stwo_b200_cs_values v = {n_proofs, 32, k.vars, k.flow_hash, k.flow_swap, nullptr, 0, nullptr, 0, nullptr};
    return stwo_b200_cs_export_flow_dev(&v, r.cs.p->num_poseidon_invocations(), pad_constants, k.status + 8, flow_hash_out, flow_swap_out, stream);
}
