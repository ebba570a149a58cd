// K7: the O(n_rows) finalisation loops of the Plonk-with-Poseidon constraint system and the trace export.
// HBM-bound: per row 6 x 4 B of wiring in, three random 16-B gathers, 12 (+10 once) coalesced 4-B column stores.
//   constraint_system/src/plonk_with_poseidon.rs:337-380  check_arithmetics
//   :382-466 populate_logup_arguments   :468-519 check_poseidon_invocations   :521-628 generate_plonk_with_poseidon_circuit
#include "common.cuh"
#include "poseidon2.cuh"

using namespace stwo_b200;

namespace {
constexpr int kT = 256;
inline unsigned nblk(size_t n) { return (unsigned)((n + kT - 1) / kT); }

__device__ __forceinline__ qm31_t ldq(const u32 *vars, u32 idx) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(vars) + idx);
    return qm31::mk(t.x, t.y, t.z, t.w);
}

__global__ void __launch_bounds__(kT) k_cs_check_arith(stwo_b200_cs_wiring w, const u32 *variables, u32 n_batch,
                                                       unsigned long long *first_bad) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    if (g >= (size_t)w.n_rows * n_batch) return;
    const u32 b = (u32)(g / w.n_rows), i = (u32)(g % w.n_rows);
    const u32 *vars = variables + (size_t)b * w.n_vars * 4;
    const qm31_t a = ldq(vars, w.a_wire[i]), bb = ldq(vars, w.b_wire[i]), c = ldq(vars, w.c_wire[i]);
    const u32 op = w.op[i];
    // c == op*(a+b) + (1-op)*a*b
    const qm31_t lhs = qm31::add(qm31::mul_m31(qm31::add(a, bb), op), qm31::mul_m31(qm31::mul(a, bb), m31::subc(1, op)));
    bool ok = qm31::eq(lhs, c);
    if (w.enforce_c_m31[i] && (c.v[1] | c.v[2] | c.v[3])) ok = false;
    if (!ok) atomicMin(first_bad + b, (unsigned long long)i);
}

// scratch layout: counts[n_vars] | first_key[n_vars] | first_prow[n_vars] | mpv aliases nothing: kept in counts2
__global__ void __launch_bounds__(kT) k_cs_count(stwo_b200_cs_wiring w, u32 *counts, u32 *first_key, u32 *first_prow, u32 *mpv) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    if (g < w.n_rows) {
        const u32 i = (u32)g;
        const u32 a = w.a_wire[i], b = w.b_wire[i], c = w.c_wire[i];
        atomicAdd(counts + a, 1u); atomicAdd(counts + b, 1u); atomicAdd(counts + c, 1u);
        atomicMin(first_key + a, 3 * i); atomicMin(first_key + b, 3 * i + 1); atomicMin(first_key + c, 3 * i + 2);
        atomicMin(first_prow + w.poseidon_wire[i], i);
    }
    if (g < w.num_input) atomicAdd(counts + g + 1, 1u);
    if (g < w.n_flow) {
        atomicAdd(counts + w.flow_swap_addr[g], 1u);
        for (int k = 0; k < 4; k++) { const u32 v = w.flow_wire[4 * g + k]; if (v) atomicAdd(mpv + v, 1u); }
    }
}
__global__ void __launch_bounds__(kT) k_cs_mult(stwo_b200_cs_wiring w, const u32 *counts, const u32 *first_key, const u32 *first_prow,
                                                const u32 *mpv, int32_t *mult_a, int32_t *mult_b, int32_t *mult_c,
                                                int32_t *mult_poseidon, u32 *status) {
    const u32 i = blockIdx.x * kT + threadIdx.x;
    if (i >= w.n_rows) return;
    const u32 a = w.a_wire[i], b = w.b_wire[i], c = w.c_wire[i], pw = w.poseidon_wire[i];
    // first occurrence in the row-major scan a_0,b_0,c_0,a_1,... gets 1 - count, later ones 1.  A wire repeated inside
    // one row (a == b) is a first occurrence only in its earliest slot, which is what the minimum key encodes.
    mult_a[i] = first_key[a] == 3 * i ? 1 - (int32_t)counts[a] : 1;
    mult_b[i] = first_key[b] == 3 * i + 1 ? 1 - (int32_t)counts[b] : 1;
    mult_c[i] = first_key[c] == 3 * i + 2 ? 1 - (int32_t)counts[c] : 1;
    int32_t mp = 0;
    if (pw != 0 && first_prow[pw] == i && mpv[pw] != 0) {
        mp = (int32_t)mpv[pw];
        if (counts[pw] != 1) atomicOr(status, 1u);      // the reference asserts counts[poseidon_wire] == 1
    }
    mult_poseidon[i] = mp;
}

__global__ void __launch_bounds__(128) k_cs_check_poseidon(stwo_b200_cs_wiring w, const u32 *variables, const u32 *flow_hash,
                                                           const uint8_t *flow_swap, u32 n_batch, const int32_t *mult_poseidon,
                                                           const u32 *first_prow, unsigned long long *first_bad) {
    const size_t g = blockIdx.x * (size_t)128 + threadIdx.x;
    if (g >= (size_t)w.n_flow * n_batch) return;
    const u32 b = (u32)(g / w.n_flow), e = (u32)(g % w.n_flow);
    const u32 *vars = variables + (size_t)b * w.n_vars * 4;
    const u32 *h = flow_hash + ((size_t)b * w.n_flow + e) * 32;
    bool ok = true;
    for (int k = 0; k < 4; k++) {
        const u32 wire = w.flow_wire[4 * e + k];
        if (!wire) continue;
        const u32 row = first_prow[wire];
        if (row >= w.n_rows || mult_poseidon[row] == 0) { ok = false; continue; }    // map.get(..).unwrap() would panic
        const qm31_t l = ldq(vars, w.a_wire[row]), r = ldq(vars, w.b_wire[row]);
        for (int j = 0; j < 4; j++) ok &= (l.v[j] == h[8 * k + j]) & (r.v[j] == h[8 * k + 4 + j]);
    }
    u32 st[16];
    const bool swap = flow_swap[(size_t)b * w.n_flow + e] != 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { st[j] = h[(swap ? 8 : 0) + j]; st[8 + j] = h[(swap ? 0 : 8) + j]; }
    poseidon2::permute<false>(st);
#pragma unroll
    for (int j = 0; j < 16; j++) ok &= st[j] == h[16 + j];
    if (!ok) atomicMin(first_bad + b, (unsigned long long)e);
}

__device__ __forceinline__ u32 m31_of_i32(int32_t v) { return v < 0 ? M31_P - (u32)(-v) : (u32)v; }

__global__ void __launch_bounds__(kT) k_cs_export_pre(stwo_b200_cs_wiring w, const int32_t *mult_a, const int32_t *mult_b,
                                                      const int32_t *mult_c, const int32_t *mult_poseidon, u32 *pre) {
    const u32 i = blockIdx.x * kT + threadIdx.x;
    if (i >= w.n_rows) return;
    const size_t n = w.n_rows;
    pre[0 * n + i] = m31_of_i32(mult_a[i]); pre[1 * n + i] = m31_of_i32(mult_b[i]); pre[2 * n + i] = m31_of_i32(mult_c[i]);
    pre[3 * n + i] = w.poseidon_wire[i]; pre[4 * n + i] = (u32)mult_poseidon[i]; pre[5 * n + i] = w.enforce_c_m31[i];
    pre[6 * n + i] = w.a_wire[i]; pre[7 * n + i] = w.b_wire[i]; pre[8 * n + i] = w.c_wire[i]; pre[9 * n + i] = w.op[i];
}
__global__ void __launch_bounds__(kT) k_cs_export_vals(stwo_b200_cs_wiring w, const u32 *variables, u32 n_batch, u32 *vals) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    if (g >= (size_t)w.n_rows * n_batch) return;
    const u32 b = (u32)(g / w.n_rows), i = (u32)(g % w.n_rows);
    const u32 *vars = variables + (size_t)b * w.n_vars * 4;
    const size_t n = w.n_rows;
    u32 *o = vals + (size_t)b * 12 * n + i;
    const qm31_t a = ldq(vars, w.a_wire[i]), bb = ldq(vars, w.b_wire[i]), c = ldq(vars, w.c_wire[i]);
#pragma unroll
    for (int k = 0; k < 4; k++) { o[k * n] = a.v[k]; o[(4 + k) * n] = bb.v[k]; o[(8 + k) * n] = c.v[k]; }
}
__global__ void k_fill64(unsigned long long *p, size_t n, unsigned long long v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
bool wiring_ok(const stwo_b200_cs_wiring *w) {
    return w && w->n_rows >= 16 && (w->n_rows & (w->n_rows - 1)) == 0 && w->n_vars >= 4 && w->a_wire && w->b_wire && w->c_wire &&
           w->poseidon_wire && w->enforce_c_m31 && w->op && (w->n_flow == 0 || (w->flow_wire && w->flow_swap_addr));
}
}  // namespace

extern "C" int32_t stwo_b200_cs_check_arithmetics_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, int64_t *first_bad,
                                                      void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !v || !v->variables || !first_bad || v->n_batch == 0) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    k_fill64<<<nblk(v->n_batch), kT, 0, st>>>((unsigned long long *)first_bad, v->n_batch, ~0ull);
    k_cs_check_arith<<<nblk((size_t)w->n_rows * v->n_batch), kT, 0, st>>>(*w, v->variables, v->n_batch, (unsigned long long *)first_bad);
    note_launch(2);
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_populate_logup_dev(const stwo_b200_cs_wiring *w, int32_t *mult_a, int32_t *mult_b, int32_t *mult_c,
                                                   int32_t *mult_poseidon, uint32_t *scratch, uint32_t *status_out, void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !mult_a || !mult_b || !mult_c || !mult_poseidon || !scratch || !status_out) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nv = w->n_vars;
    u32 *counts = scratch, *first_key = scratch + nv, *first_prow = scratch + 2 * nv, *mpv = scratch + 3 * nv;
    STWO_CUDA(cudaMemsetAsync(counts, 0, nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(first_key, 0xff, 2 * nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(mpv, 0, nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(status_out, 0, 4, st));
    size_t span = w->n_rows;
    if (w->n_flow > span) span = w->n_flow;
    if (w->num_input > span) span = w->num_input;
    k_cs_count<<<nblk(span), kT, 0, st>>>(*w, counts, first_key, first_prow, mpv);
    k_cs_mult<<<nblk(w->n_rows), kT, 0, st>>>(*w, counts, first_key, first_prow, mpv, mult_a, mult_b, mult_c, mult_poseidon, status_out);
    note_launch(2);
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_check_poseidon_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v,
                                                   const int32_t *mult_poseidon, const uint32_t *scratch, int64_t *first_bad, void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !v || !v->variables || !first_bad || !mult_poseidon || !scratch || v->n_batch == 0) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    k_fill64<<<nblk(v->n_batch), kT, 0, st>>>((unsigned long long *)first_bad, v->n_batch, ~0ull);
    note_launch(1);
    if (w->n_flow) {
        if (!v->flow_hash || !v->flow_swap) return STWO_B200_E_BAD_ARG;
        k_cs_check_poseidon<<<(unsigned)(((size_t)w->n_flow * v->n_batch + 127) / 128), 128, 0, st>>>(
            *w, v->variables, v->flow_hash, v->flow_swap, v->n_batch, mult_poseidon, scratch + 2 * (size_t)w->n_vars, (unsigned long long *)first_bad);
        note_launch(1);
    }
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_export_trace_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const int32_t *mult_a,
                                                 const int32_t *mult_b, const int32_t *mult_c, const int32_t *mult_poseidon,
                                                 uint32_t *preprocessed, uint32_t *values, void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !v || !v->variables || v->n_batch == 0 || !values) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (preprocessed) {
        if (!mult_a || !mult_b || !mult_c || !mult_poseidon) return STWO_B200_E_BAD_ARG;
        k_cs_export_pre<<<nblk(w->n_rows), kT, 0, st>>>(*w, mult_a, mult_b, mult_c, mult_poseidon, preprocessed);
        note_launch(1);
    }
    k_cs_export_vals<<<nblk((size_t)w->n_rows * v->n_batch), kT, 0, st>>>(*w, v->variables, v->n_batch, values);
    note_launch(1);
    return cuda_status(cudaGetLastError());
}

extern "C" int32_t stwo_b200_cs_finalize(const stwo_b200_cs_wiring *hw, const stwo_b200_cs_values *hv, uint32_t *trace,
                                         int64_t *bad_row, int64_t *bad_flow) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(hw) || !hv || hv->n_batch != 1 || !hv->variables || !trace || !bad_row || !bad_flow) return STWO_B200_E_BAD_ARG;
    const size_t nr = hw->n_rows, nv = hw->n_vars, nf = hw->n_flow;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_wires = take(6 * nr * 4), o_fw = take(nf * 4 * 4 + 4), o_fa = take(nf * 4 + 4), o_vars = take(nv * 16),
                 o_fh = take(nf * 128 + 4), o_fs = take(nf + 4), o_mult = take(4 * nr * 4), o_scr = take((4 * nv + 4) * 4),
                 o_bad = take(16), o_stat = take(4), o_trace = take(22 * nr * 4);
    int32_t rc = stage_reserve(off);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    uint8_t *d = stage_dev();
    u32 *dw = (u32 *)(d + o_wires);
    const u32 *srcs[6] = {hw->a_wire, hw->b_wire, hw->c_wire, hw->poseidon_wire, hw->enforce_c_m31, hw->op};
    for (int k = 0; k < 6; k++) STWO_CUDA(cudaMemcpyAsync(dw + k * nr, srcs[k], nr * 4, cudaMemcpyHostToDevice, st));
    if (nf) {
        STWO_CUDA(cudaMemcpyAsync(d + o_fw, hw->flow_wire, nf * 16, cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d + o_fa, hw->flow_swap_addr, nf * 4, cudaMemcpyHostToDevice, st));
        if (!hv->flow_hash || !hv->flow_swap) return STWO_B200_E_BAD_ARG;
        STWO_CUDA(cudaMemcpyAsync(d + o_fh, hv->flow_hash, nf * 128, cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d + o_fs, hv->flow_swap, nf, cudaMemcpyHostToDevice, st));
    }
    STWO_CUDA(cudaMemcpyAsync(d + o_vars, hv->variables, nv * 16, cudaMemcpyHostToDevice, st));
    stwo_b200_cs_wiring w = *hw;
    w.a_wire = dw; w.b_wire = dw + nr; w.c_wire = dw + 2 * nr; w.poseidon_wire = dw + 3 * nr; w.enforce_c_m31 = dw + 4 * nr; w.op = dw + 5 * nr;
    w.flow_wire = (u32 *)(d + o_fw); w.flow_swap_addr = (u32 *)(d + o_fa);
    stwo_b200_cs_values v = {1, (u32 *)(d + o_vars), (u32 *)(d + o_fh), d + o_fs};
    int32_t *m = (int32_t *)(d + o_mult);
    int64_t *bad = (int64_t *)(d + o_bad);
    u32 *scr = (u32 *)(d + o_scr), *stat = (u32 *)(d + o_stat), *tr = (u32 *)(d + o_trace);
    if ((rc = stwo_b200_cs_check_arithmetics_dev(&w, &v, bad, st))) return rc;
    if ((rc = stwo_b200_cs_populate_logup_dev(&w, m, m + nr, m + 2 * nr, m + 3 * nr, scr, stat, st))) return rc;
    if ((rc = stwo_b200_cs_check_poseidon_dev(&w, &v, m + 3 * nr, scr, bad + 1, st))) return rc;
    if ((rc = stwo_b200_cs_export_trace_dev(&w, &v, m, m + nr, m + 2 * nr, m + 3 * nr, tr, tr + 10 * nr, st))) return rc;
    int64_t hb[2];
    u32 hstat = 0;
    STWO_CUDA(cudaMemcpyAsync(hb, bad, 16, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaMemcpyAsync(&hstat, stat, 4, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaMemcpyAsync(trace, tr, 22 * nr * 4, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaStreamSynchronize(st));
    *bad_row = hb[0]; *bad_flow = hb[1];
    if (hstat) *bad_flow = -2;     // a Poseidon wire is referenced by more than one row (reference: assert_eq!(counts[..], 1))
    return STWO_B200_OK;
}
