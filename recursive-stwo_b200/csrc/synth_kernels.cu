// Generator of the synthetic FRI + Merkle instances on the device (synth.cuh has the element-level arithmetic and the blob layout): every
// instance of a batch in parallel, one thread per column value / tree node / folded value, a handful of launches per layer.
// BASELINE configs[4] part i; SURVEY.md 8d config 5-i.  The generator is the workload's input synthesis: it is never inside a timed
// region.  What the verifier then does with the instances is stwo_b200_synth_verify_batch_dev (verify_kernels.cu).
#include "common.cuh"
#include "synth.cuh"
#include <string.h>
#include <vector>

using namespace stwo_b200;

namespace {
constexpr int kT = 128;
inline unsigned nblk(size_t n) { return (unsigned)((n + kT - 1) / kT); }

struct Plan {                                  // device layout of the generator's scratch for a batch
    verify::Shape shape;
    u32 n, n_logs, logs[fri::MAX_LOGS];
    u32 *cols[fri::MAX_LOGS];                   // [n][2^L][4]
    u32 *first[verify::MAX_DEPTH + 1];          // first-layer tree, level h: [n][2^h][8]
    u32 *layer[proof::MAX_INNER + 1];           // line layer li (log size mf - 1 - li): [n][2^L][4]; index n_inner: the last layer
    u32 *tree[proof::MAX_INNER][verify::MAX_DEPTH + 1];
    fs::Out *fs;                                // [n]
    int *status;                                // [n]  0 ok, -1 last layer not of the claimed degree, -2 no nonce
    size_t bytes;
};
Plan plan(const verify::Shape &s, u32 n, uint8_t *base) {
    Plan pl;
    memset(&pl, 0, sizeof pl);
    pl.shape = s; pl.n = n;
    proof::Desc d;
    memset(&d, 0, sizeof d);
    synth::set_offsets(s, d);
    pl.n_logs = fri::log_sizes(d, pl.logs);
    size_t at = 0;
    auto take = [&](size_t bytes) { at = align_up(at, 256); uint8_t *p = base ? base + at : nullptr; at += bytes; return p; };
    const u32 mf = s.max_first();
    for (u32 g = 0; g < pl.n_logs; g++) pl.cols[g] = (u32 *)take((size_t)n * 16 << pl.logs[g]);
    for (u32 h = 0; h <= mf; h++) pl.first[h] = (u32 *)take((size_t)n * 32 << h);
    for (u32 li = 0; li <= s.n_inner; li++) pl.layer[li] = (u32 *)take((size_t)n * 16 << (mf - 1 - li));
    for (u32 li = 0; li < s.n_inner; li++)
        for (u32 h = 0; h <= mf - 1 - li; h++) pl.tree[li][h] = (u32 *)take((size_t)n * 32 << h);
    pl.fs = (fs::Out *)take((size_t)n * sizeof(fs::Out));
    pl.status = (int *)take((size_t)n * sizeof(int));
    pl.bytes = align_up(at, 256);
    return pl;
}

__global__ void __launch_bounds__(kT) k_syn_header(verify::Shape s, u32 n, u32 *blobs, u64 *blob_off, u32 stride) {
    const u32 p = blockIdx.x * kT + threadIdx.x;
    if (p > n) return;
    blob_off[p] = (u64)p * stride;
    if (p < n) synth::write_header(blobs + (size_t)p * stride, s, synth::layout(s));
}
__global__ void __launch_bounds__(kT) k_syn_columns(u32 n, u64 seed0, u32 g, u32 L, u32 log_blowup, u32 *col) {
    const size_t t = blockIdx.x * (size_t)kT + threadIdx.x;
    if (t >= ((size_t)n << L)) return;
    const u32 p = (u32)(t >> L), i = (u32)(t & ((1u << L) - 1u));
    fs::qstore(col + 4 * t, synth::column_value(seed0 + p, g, L, log_blowup, i));
}
// level h of a tree for every instance: children from the level below (null: leaf layer), one QM31 of data per node (null: none)
__global__ void __launch_bounds__(kT) k_syn_tree_level(u32 n, u32 h, const u32 *below, const u32 *data, u32 *out) {
    const size_t t = blockIdx.x * (size_t)kT + threadIdx.x;
    if (t >= ((size_t)n << h)) return;
    u32 l8[8], r8[8], v4[4], o8[8];
    if (below) for (int k = 0; k < 8; k++) { l8[k] = below[16 * t + k]; r8[k] = below[16 * t + 8 + k]; }
    if (data) for (int k = 0; k < 4; k++) v4[k] = data[4 * t + k];
    synth::node_hash(below ? l8 : nullptr, below ? r8 : nullptr, data ? v4 : nullptr, o8);
    for (int k = 0; k < 8; k++) out[8 * t + k] = o8[k];
}
__global__ void __launch_bounds__(kT) k_syn_store_root(u32 n, const u32 *roots, u32 *blobs, u32 stride, u32 off) {
    const u32 t = blockIdx.x * kT + threadIdx.x;
    if (t < n * 8) blobs[(size_t)(t / 8) * stride + off + t % 8] = roots[t];
}
__global__ void __launch_bounds__(kT) k_syn_transcript(verify::Shape s, u32 n, const u32 *blobs, u32 stride, fs::Out *out, u32 upto) {
    const u32 p = blockIdx.x * kT + threadIdx.x;
    if (p >= n) return;
    proof::Desc d;
    synth::set_offsets(s, d);
    synth::transcript(blobs + (size_t)p * stride, d, out[p], upto);
}
// cur[j] = alpha_li^2 * cur[j] + circle_fold(col[2j], col[2j+1])   (first = cur starts at zero)
__global__ void __launch_bounds__(kT) k_syn_accumulate(u32 n, u32 L, const u32 *col, u32 *cur, const fs::Out *fso, u32 li, int first) {
    const size_t t = blockIdx.x * (size_t)kT + threadIdx.x;
    if (t >= ((size_t)n << (L - 1))) return;
    const u32 p = (u32)(t >> (L - 1)), j = (u32)(t & ((1u << (L - 1)) - 1u));
    const qm31_t alpha = fso[p].fri_alphas[li];
    const u32 *c = col + (((size_t)p << L) + 2 * j) * 4;
    qm31_t f = synth::circle_fold_at(L, j, fs::qload(c), fs::qload(c + 4), alpha);
    if (!first) f = fs::qadd(fs::qmul(fs::qmul(alpha, alpha), fs::qload(cur + 4 * t)), f);
    fs::qstore(cur + 4 * t, f);
}
__global__ void __launch_bounds__(kT) k_syn_line_fold(u32 n, u32 L, const u32 *cur, u32 *nxt, const fs::Out *fso, u32 li) {
    const size_t t = blockIdx.x * (size_t)kT + threadIdx.x;
    if (t >= ((size_t)n << (L - 1))) return;
    const u32 p = (u32)(t >> (L - 1)), j = (u32)(t & ((1u << (L - 1)) - 1u));
    const u32 *c = cur + (((size_t)p << L) + 2 * j) * 4;
    fs::qstore(nxt + 4 * t, synth::line_fold_at(L, j, fs::qload(c), fs::qload(c + 4), fso[p].fri_alphas[li + 1]));
}
// last-layer polynomial from the first 2^log_last values of the final layer, checked against every value of that layer
constexpr u32 kMaxLast = 1u << 6;              // log_last <= 6 for generated instances
__global__ void __launch_bounds__(kT) k_syn_last(verify::Shape s, u32 n, const u32 *last, u32 *blobs, u32 stride, int *status) {
    const u32 p = blockIdx.x * kT + threadIdx.x;
    if (p >= n) return;
    const u32 Llast = s.max_first() - 1 - s.n_inner, nc = 1u << s.log_last;
    const u32 *ev = last + (((size_t)p << Llast)) * 4;
    qm31_t v[kMaxLast], out[kMaxLast], tmp[kMaxLast];
    for (u32 k = 0; k < nc; k++) v[k] = fs::qload(ev + 4 * k);
    synth::interpolate_line(Llast, s.log_last, v, out, tmp);
    u32 *coeffs = blobs + (size_t)p * stride + synth::layout(s).off_last;
    for (u32 k = 0; k < nc; k++) fs::qstore(coeffs + 4 * k, out[k]);
    int st = 0;
    for (u32 r = 0; r < (1u << Llast); r++)
        if (!qm31::eq(fri::eval_last_poly(coeffs, s.log_last, circle::dbl(fri::absolute_point(Llast + 1, 2 * r)).x, tmp), fs::qload(ev + 4 * r))) st = -1;
    status[p] = st;
}
__global__ void __launch_bounds__(kT) k_syn_grind(verify::Shape s, u32 n, u32 *blobs, u32 stride, fs::Out *out, int *status) {
    const u32 p = blockIdx.x * kT + threadIdx.x;
    if (p >= n) return;
    proof::Desc d;
    synth::set_offsets(s, d);
    u32 *w = blobs + (size_t)p * stride;
    bool found = false;
    for (u32 nonce = 0; nonce < (1u << 16) && !found; nonce++) {
        w[synth::H_NONCE] = nonce; w[synth::H_NONCE + 1] = 0;
        synth::transcript(w, d, out[p], 1 + s.n_inner);
        found = out[p].pow_ok && synth::queries_distinct(d, out[p]);
    }
    if (!found && status[p] == 0) status[p] = -2;
}
struct DevPlan { u32 *cols[fri::MAX_LOGS]; u32 *layer[proof::MAX_INNER + 1]; u32 logs[fri::MAX_LOGS]; u32 n_logs; };
// opened values: answers and the fri_witness streams (thread per instance; a few dozen values each)
__global__ void __launch_bounds__(kT) k_syn_open(verify::Shape s, u32 n, DevPlan pl, u32 *blobs, u32 stride, const fs::Out *out) {
    const u32 p = blockIdx.x * kT + threadIdx.x;
    if (p >= n) return;
    proof::Desc d;
    synth::set_offsets(s, d);
    const synth::Layout l = synth::layout(s);
    u32 *w = blobs + (size_t)p * stride;
    const u32 nq = s.n_queries, mf = s.max_first();
    u32 sp[proof::MAX_QUERIES];
    u32 wi = 0;
    for (u32 g = 0; g < pl.n_logs; g++) {
        const u32 L = pl.logs[g];
        const u32 *col = pl.cols[g] + (((size_t)p << L)) * 4;
        for (u32 i = 0; i < nq; i++) {
            sp[i] = fri::position(d, out[p].raw_queries[i], L);
            for (int k = 0; k < 4; k++) w[l.off_ans + (g * nq + i) * 4 + k] = col[4 * (size_t)sp[i] + k];
        }
        const u32 ns = decommit::sort_unique(sp, nq);
        for (u32 k = 0; k < ns;) {
            const u32 start = sp[k] & ~1u;
            for (u32 e = start; e < start + 2; e++) {
                if (k < ns && sp[k] == e) k++;
                else { for (int c = 0; c < 4; c++) w[l.off_flfw + 4 * wi + c] = col[4 * (size_t)e + c]; wi++; }
            }
        }
    }
    w[synth::H_FL_NFW] = wi;
    for (u32 li = 0; li < s.n_inner; li++) {
        const u32 L = mf - 1 - li;
        const u32 *lay = pl.layer[li] + (((size_t)p << L)) * 4;
        for (u32 i = 0; i < nq; i++) sp[i] = fri::position(d, out[p].raw_queries[i], L);
        const u32 ns = decommit::sort_unique(sp, nq);
        u32 w2 = 0;
        for (u32 k = 0; k < ns; k++)
            if (decommit::find(sp, ns, sp[k] ^ 1u) < 0) { for (int c = 0; c < 4; c++) w[l.off_infw[li] + 4 * w2 + c] = lay[4 * (size_t)(sp[k] ^ 1u) + c]; w2++; }
        w[synth::H_IN_NFW + li] = w2;
    }
}
struct TreeLevels { const u32 *lv[verify::MAX_DEPTH + 1]; };
// hash_witness of one FRI tree per thread: (instance, tree f)
__global__ void __launch_bounds__(64) k_syn_hash_witness(verify::Shape s, u32 n, u32 f, TreeLevels tl, u32 *blobs, u32 stride, const fs::Out *out) {
    const u32 p = blockIdx.x * 64 + threadIdx.x;
    if (p >= n) return;
    proof::Desc d;
    synth::set_offsets(s, d);
    const synth::Layout l = synth::layout(s);
    u32 *w = blobs + (size_t)p * stride;
    const u32 nq = s.n_queries, depth = s.fri_depth(f);
    u32 q[proof::MAX_QUERIES], scr[6 * proof::MAX_QUERIES];
    for (u32 i = 0; i < nq; i++) q[i] = fri::position(d, out[p].raw_queries[i], depth);
    const u32 cnt = synth::emit_hash_witness(depth, s.fri_data_mask(f), q, nq, [&](u32 h, u32 pos) { return tl.lv[h] + (((size_t)p << h) + pos) * 8; },
                                             w + (f ? l.off_inhw[f - 1] : l.off_flhw), scr);
    w[f ? synth::H_IN_NHW + f - 1 : synth::H_FL_NHW] = cnt;
}
bool synth_shape_ok(const stwo_b200_proof_shape *s) {
    return s && proof::shape_consistent(s->log_size_plonk, s->log_size_poseidon, s->pow_bits, s->log_blowup, s->log_last, s->n_queries, s->n_inner) &&
           s->log_last <= 6 && s->pow_bits <= 10;
}
}  // namespace

extern "C" uint32_t stwo_b200_synth_blob_words(const stwo_b200_proof_shape *shape) {
    if (!synth_shape_ok(shape)) return 0;
    verify::Shape s;
    memcpy(&s, shape, sizeof s);
    return synth::layout(s).total;
}
extern "C" size_t stwo_b200_synth_scratch_bytes(const stwo_b200_proof_shape *shape, uint32_t n) {
    if (!synth_shape_ok(shape) || !n) return 0;
    verify::Shape s;
    memcpy(&s, shape, sizeof s);
    return plan(s, n, nullptr).bytes;
}

extern "C" int32_t stwo_b200_synth_generate_dev(const stwo_b200_proof_shape *shape, uint32_t n, uint64_t seed0, uint32_t *blobs, uint64_t *blob_off,
                                                void *scratch, size_t scratch_bytes, int32_t *status, void *stream) {
    STWO_CHECK_DEVICE();
    if (!n) return STWO_B200_OK;
    if (!synth_shape_ok(shape)) return STWO_B200_E_SHAPE;
    if (!blobs || !blob_off || !scratch) return STWO_B200_E_BAD_ARG;
    verify::Shape s;
    memcpy(&s, shape, sizeof s);
    const Plan pl = plan(s, n, (uint8_t *)scratch);
    if (pl.bytes > scratch_bytes) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const synth::Layout l = synth::layout(s);
    const u32 stride = l.total, mf = s.max_first();
    STWO_CUDA(cudaMemsetAsync(blobs, 0, (size_t)n * stride * 4, st));
    k_syn_header<<<nblk(n + 1), kT, 0, st>>>(s, n, blobs, blob_off, stride);
    for (u32 g = 0; g < pl.n_logs; g++) k_syn_columns<<<nblk((size_t)n << pl.logs[g]), kT, 0, st>>>(n, seed0, g, pl.logs[g], s.log_blowup, pl.cols[g]);
    // first-layer tree: the column of log size h is hashed into the nodes of layer h
    for (u32 h = mf + 1; h-- > 0;) {
        const u32 *data = nullptr;
        for (u32 g = 0; g < pl.n_logs; g++) if (pl.logs[g] == h) data = pl.cols[g];
        k_syn_tree_level<<<nblk((size_t)n << h), kT, 0, st>>>(n, h, h == mf ? nullptr : pl.first[h + 1], data, pl.first[h]);
    }
    k_syn_store_root<<<nblk((size_t)n * 8), kT, 0, st>>>(n, pl.first[0], blobs, stride, l.off_flc);
    k_syn_transcript<<<nblk(n), kT, 0, st>>>(s, n, blobs, stride, pl.fs, 1);
    note_launch(4 + pl.n_logs + mf);
    for (u32 li = 0; li < s.n_inner; li++) {
        const u32 L = mf - 1 - li;
        // line layer li: what the folds so far produced, plus the circle fold of the column that enters at this size
        u32 *cur = pl.layer[li];
        for (u32 g = 0; g < pl.n_logs; g++)
            if (pl.logs[g] == mf - li) { k_syn_accumulate<<<nblk((size_t)n << L), kT, 0, st>>>(n, pl.logs[g], pl.cols[g], cur, pl.fs, li, li == 0); note_launch(1); }
        for (u32 h = L + 1; h-- > 0;)
            k_syn_tree_level<<<nblk((size_t)n << h), kT, 0, st>>>(n, h, h == L ? nullptr : pl.tree[li][h + 1], h == L ? cur : nullptr, pl.tree[li][h]);
        k_syn_store_root<<<nblk((size_t)n * 8), kT, 0, st>>>(n, pl.tree[li][0], blobs, stride, l.off_inc[li]);
        k_syn_transcript<<<nblk(n), kT, 0, st>>>(s, n, blobs, stride, pl.fs, 2 + li);
        k_syn_line_fold<<<nblk((size_t)n << (L - 1)), kT, 0, st>>>(n, L, cur, pl.layer[li + 1], pl.fs, li);
        note_launch(L + 4);
    }
    k_syn_last<<<nblk(n), kT, 0, st>>>(s, n, pl.layer[s.n_inner], blobs, stride, pl.status);
    k_syn_grind<<<nblk(n), kT, 0, st>>>(s, n, blobs, stride, pl.fs, pl.status);
    DevPlan dp;
    memset(&dp, 0, sizeof dp);
    dp.n_logs = pl.n_logs;
    for (u32 g = 0; g < pl.n_logs; g++) { dp.cols[g] = pl.cols[g]; dp.logs[g] = pl.logs[g]; }
    for (u32 li = 0; li <= s.n_inner; li++) dp.layer[li] = pl.layer[li];
    k_syn_open<<<nblk(n), kT, 0, st>>>(s, n, dp, blobs, stride, pl.fs);
    for (u32 f = 0; f <= s.n_inner; f++) {
        TreeLevels tl;
        memset(&tl, 0, sizeof tl);
        for (u32 h = 0; h <= s.fri_depth(f); h++) tl.lv[h] = f ? pl.tree[f - 1][h] : pl.first[h];
        k_syn_hash_witness<<<(n + 63) / 64, 64, 0, st>>>(s, n, f, tl, blobs, stride, pl.fs);
    }
    note_launch(4 + s.n_inner);
    if (status) STWO_CUDA(cudaMemcpyAsync(status, pl.status, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return cuda_status(cudaGetLastError());
}
