// The stage-level entry points of the drop-in boundary: a host that wants to replace ONE of the reference's hint
// stages -- not the whole verifier -- calls these with host blobs and gets batch-shaped host arrays back.  Each is a thin driver
// over the stage kernels of verify_kernels.cu (stwo_b200_verify_proofs_batch_dev with a stop flag) plus one copy per result.
//   stwo_b200_channel_replay_batch        FiatShamirHints::new          components/hints/src/fiat_shamir.rs:69-307
//   stwo_b200_fri_answers_batch           AnswerHints::compute          components/hints/src/answer.rs:40-48
//   stwo_b200_fri_fold_batch              First/InnerLayersHints        components/hints/src/folding.rs:326-363,481-595
//   stwo_b200_hash_column_capacity_batch  hash_column_get_capacity      components/hints/src/folding.rs:77 (primitives/merkle/src/lib.rs:141-181)
// and the multi-GPU part of the boundary: proofs are sharded by rank in contiguous blocks, the only collectives are the
// all-gather of the verdict bytes and (optional) the gather of trace columns, both over NCCL.  NCCL is resolved
// at run time (dlopen of libnccl.so.2): a single-GPU host never needs it.
#include "common.cuh"
#include "merkle.cuh"
#include <dlfcn.h>
#include <string.h>
#include <vector>

using namespace stwo_b200;

namespace {
constexpr int kT = 128;

// hash_m31_columns_get_capacity (primitives/merkle/src/lib.rs:141-181): 8 words per chunk, zero padded, capacity chained
__global__ void __launch_bounds__(kT) k_hash_column_capacity(const u32 *__restrict__ cols, u32 n_cols, size_t n, u32 *__restrict__ out) {
    const size_t i = blockIdx.x * (size_t)kT + threadIdx.x;
    if (i >= n) return;
    u32 st[16];
#pragma unroll
    for (int k = 8; k < 16; k++) st[k] = 0;
    const u32 *c = cols + i * n_cols;
    const u32 n_chunks = n_cols ? (n_cols + 7) / 8 : 1;
    for (u32 ch = 0; ch < n_chunks; ch++) {
#pragma unroll
        for (u32 k = 0; k < 8; k++) st[k] = 8 * ch + k < n_cols ? __ldg(c + 8 * ch + k) : 0u;
        poseidon2::permute<false>(st);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[8 * i + k] = st[8 + k];
}

// one same-shape batch of host blobs on the staging area: upload, run the stages up to `upto`, leave the workspace for fetches
struct Staged {
    stwo_b200_proof_shape shape;
    uint8_t *ws; size_t ws_bytes;
    uint8_t *d_verdict;
    cudaStream_t st;
};
int32_t stage_and_run(const uint8_t *const *blobs, const size_t *lens, uint32_t n, const stwo_b200_pcs_config *config, const uint32_t *input_idx,
                      const uint32_t *input_vals, uint32_t n_inputs, uint32_t upto, Staged &sg, std::vector<uint8_t> &host_verdict) {
    if (!blobs || !lens || !n || !config || (n_inputs && (!input_idx || !input_vals))) return STWO_B200_E_BAD_ARG;
    // the statement's log sizes come from the first blob that parses, the PcsConfig from the caller; blobs of any other shape fail parse
    stwo_b200_proof_shape claimed;
    bool have = false;
    for (uint32_t i = 0; i < n && !have; i++) have = blobs[i] && stwo_b200_proof_shape_of(blobs[i], lens[i], &claimed) == STWO_B200_OK;
    if (!have) return STWO_B200_E_SHAPE;
    int32_t rc = stwo_b200_shape_from_config(config, claimed.log_size_plonk, claimed.log_size_poseidon, &sg.shape);
    if (rc) return rc;
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) off[i + 1] = off[i] + (blobs[i] ? (lens[i] + 3) / 4 : 0);
    const size_t b_blobs = align_up(off[n] * 4 + 4, 256), b_off = align_up((n + 1) * 8, 256), b_in = align_up((size_t)n_inputs * 20 + 4, 256);
    sg.ws_bytes = stwo_b200_verify_workspace_bytes(&sg.shape, n);
    if (!sg.ws_bytes) return STWO_B200_E_SHAPE;
    if ((rc = stage_reserve(b_blobs + b_off + b_in + align_up(sg.ws_bytes, 256) + align_up(2 * (size_t)n, 256)))) return rc;
    sg.st = stage_stream();
    STWO_CUDA(cudaStreamSynchronize(sg.st));
    uint8_t *d = stage_dev();
    u32 *d_blobs = (u32 *)d; d += b_blobs;
    uint64_t *d_off = (uint64_t *)d; d += b_off;
    u32 *d_idx = (u32 *)d, *d_vals = d_idx + n_inputs; d += b_in;
    sg.ws = d; d += align_up(sg.ws_bytes, 256);
    sg.d_verdict = d;
    STWO_CUDA(cudaMemsetAsync(d_blobs, 0, b_blobs, sg.st));                         // a blob whose length is not a multiple of 4 is zero padded
    for (uint32_t i = 0; i < n; i++)
        if (blobs[i] && lens[i]) STWO_CUDA(cudaMemcpyAsync(d_blobs + off[i], blobs[i], lens[i], cudaMemcpyHostToDevice, sg.st));
    STWO_CUDA(cudaMemcpyAsync(d_off, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, sg.st));
    if (n_inputs) {
        STWO_CUDA(cudaMemcpyAsync(d_idx, input_idx, n_inputs * 4, cudaMemcpyHostToDevice, sg.st));
        STWO_CUDA(cudaMemcpyAsync(d_vals, input_vals, n_inputs * 16, cudaMemcpyHostToDevice, sg.st));
    }
    if ((rc = stwo_b200_verify_proofs_batch_dev(d_blobs, d_off, n, &sg.shape, d_idx, d_vals, n_inputs, upto | STWO_B200_VERIFY_ONE_STREAM, sg.ws,
                                                sg.ws_bytes, sg.d_verdict, sg.d_verdict + n, sg.st)))
        return rc;
    host_verdict.resize(2 * (size_t)n);
    STWO_CUDA(cudaMemcpyAsync(host_verdict.data(), sg.d_verdict, 2 * (size_t)n, cudaMemcpyDeviceToHost, sg.st));
    return cuda_status(cudaStreamSynchronize(sg.st));
}
void copy_status(const std::vector<uint8_t> &hv, uint32_t n, uint8_t *verdict, uint8_t *stage) {
    if (verdict) memcpy(verdict, hv.data(), n);
    if (stage) memcpy(stage, hv.data() + n, n);
}
}  // namespace

extern "C" int32_t stwo_b200_channel_replay_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                                  const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                                  stwo_b200_verify_detail *out, uint8_t *verdict, uint8_t *stage) {
    STWO_CHECK_DEVICE();
    if (!out) return STWO_B200_E_BAD_ARG;
    Staged sg;
    std::vector<uint8_t> hv;
    int32_t rc = stage_and_run(blobs, lens, n_proofs, config, input_idx, input_vals, n_inputs, STWO_B200_VERIFY_UPTO_TRANSCRIPT, sg, hv);
    if (rc) return rc;
    copy_status(hv, n_proofs, verdict, stage);
    return stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_DETAIL, out, (size_t)n_proofs * sizeof(stwo_b200_verify_detail), sg.st);
}

extern "C" int32_t stwo_b200_fri_answers_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                               const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs, uint32_t *answers,
                                               uint32_t *domain_points, uint8_t *verdict, uint8_t *stage) {
    STWO_CHECK_DEVICE();
    if (!answers) return STWO_B200_E_BAD_ARG;
    Staged sg;
    std::vector<uint8_t> hv;
    int32_t rc = stage_and_run(blobs, lens, n_proofs, config, input_idx, input_vals, n_inputs, STWO_B200_VERIFY_UPTO_ANSWERS, sg, hv);
    if (rc) return rc;
    copy_status(hv, n_proofs, verdict, stage);
    const size_t per = (size_t)3 * sg.shape.n_queries;
    if ((rc = stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_ANSWERS, answers, n_proofs * per * 16, sg.st))) return rc;
    if (domain_points) rc = stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_DOMAIN_POINTS, domain_points, n_proofs * per * 8, sg.st);
    return rc;
}

extern "C" int32_t stwo_b200_fri_fold_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                            const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs, uint32_t *circle_folds,
                                            uint32_t *line_folds, uint32_t *last_evals, uint8_t *verdict, uint8_t *stage) {
    STWO_CHECK_DEVICE();
    if (!circle_folds && !line_folds && !last_evals) return STWO_B200_E_BAD_ARG;
    Staged sg;
    std::vector<uint8_t> hv;
    int32_t rc = stage_and_run(blobs, lens, n_proofs, config, input_idx, input_vals, n_inputs, STWO_B200_VERIFY_UPTO_FOLDS, sg, hv);
    if (rc) return rc;
    copy_status(hv, n_proofs, verdict, stage);
    const size_t nq = sg.shape.n_queries, n = n_proofs;
    if (circle_folds && (rc = stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_CIRCLE_FOLDS, circle_folds, n * 3 * nq * 16, sg.st))) return rc;
    if (line_folds && (rc = stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_LINE_FOLDS, line_folds, n * 32 * nq * 16, sg.st))) return rc;
    if (last_evals && (rc = stwo_b200_verify_fetch_batch(sg.ws, &sg.shape, n_proofs, STWO_B200_FETCH_LAST_EVALS, last_evals, n * nq * 16, sg.st))) return rc;
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_hash_column_capacity_batch_dev(const uint32_t *cols, uint32_t n_cols, size_t n, uint32_t *out, void *stream) {
    STWO_CHECK_DEVICE();
    if (!n) return STWO_B200_OK;
    if (!cols || !out) return STWO_B200_E_BAD_ARG;
    k_hash_column_capacity<<<(unsigned)((n + kT - 1) / kT), kT, 0, (cudaStream_t)stream>>>(cols, n_cols, n, out);
    note_launch(1);
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_hash_column_capacity_batch(const uint32_t *cols, uint32_t n_cols, size_t n, uint32_t *out) {
    STWO_CHECK_DEVICE();
    if (!n) return STWO_B200_OK;
    if (!cols || !out) return STWO_B200_E_BAD_ARG;
    const size_t b_in = align_up(n * (size_t)n_cols * 4 + 4, 256), b_out = n * 32;
    int32_t rc = stage_reserve(b_in + b_out);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    STWO_CUDA(cudaStreamSynchronize(st));
    u32 *d_in = (u32 *)stage_dev(), *d_out = (u32 *)(stage_dev() + b_in);
    if (n_cols) STWO_CUDA(cudaMemcpyAsync(d_in, cols, n * (size_t)n_cols * 4, cudaMemcpyHostToDevice, st));
    if ((rc = stwo_b200_hash_column_capacity_batch_dev(d_in, n_cols, n, d_out, st))) return rc;
    STWO_CUDA(cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t stwo_b200_shard_range(uint64_t n, uint32_t rank, uint32_t world, uint64_t *lo, uint64_t *hi) {
    if (!world || rank >= world || !lo || !hi) return STWO_B200_E_BAD_ARG;
    // contiguous blocks, the first n % world ranks one longer
    const uint64_t base = n / world, extra = n % world;
    *lo = rank * base + (rank < extra ? rank : extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
    return STWO_B200_OK;
}

struct Id128 { char b[128]; };                // ncclUniqueId
namespace {
// the few NCCL entry points the boundary needs, resolved on first use (ncclResult_t is an int, 0 = success)
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
Nccl g_nccl;
bool nccl_load() {
    if (g_nccl.lib) return g_nccl.AllGather != nullptr;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.lib = h;
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd) {
        g_nccl.AllGather = nullptr;
        return false;
    }
    return true;
}
constexpr int kNcclUint8 = 1, kNcclUint32 = 3;       // ncclUint8 / ncclUint32 of nccl.h
__global__ void k_place_bytes(const uint8_t *src, uint8_t *dst, size_t n) {
    const size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
}  // namespace

extern "C" int32_t stwo_b200_comm_unique_id(uint8_t id[128]) {
    if (!id) return STWO_B200_E_BAD_ARG;
    if (!nccl_load()) return STWO_B200_E_NO_NCCL;
    return g_nccl.GetUniqueId(id) == 0 ? STWO_B200_OK : STWO_B200_E_NCCL;
}
extern "C" int32_t stwo_b200_comm_init(const uint8_t id[128], uint32_t rank, uint32_t world, void **comm) {
    STWO_CHECK_DEVICE();
    if (!id || !comm || !world || rank >= world) return STWO_B200_E_BAD_ARG;
    if (!nccl_load()) return STWO_B200_E_NO_NCCL;
    Id128 u;
    memcpy(u.b, id, 128);
    return g_nccl.CommInitRank(comm, (int)world, u, (int)rank) == 0 ? STWO_B200_OK : STWO_B200_E_NCCL;
}
extern "C" int32_t stwo_b200_comm_destroy(void *comm) {
    if (!comm) return STWO_B200_OK;
    if (!nccl_load()) return STWO_B200_E_NO_NCCL;
    return g_nccl.CommDestroy(comm) == 0 ? STWO_B200_OK : STWO_B200_E_NCCL;
}

// verdict / stage bytes of every rank's block, in proof order, on every rank.  Blocks are those of stwo_b200_shard_range; NCCL's
// all-gather wants equal counts, so each rank contributes max_block bytes of (verdict | stage) and the result is compacted.
extern "C" int32_t stwo_b200_gather_verdicts(void *comm, uint32_t rank, uint32_t world, uint64_t n_total, const uint8_t *verdict_local,
                                             const uint8_t *stage_local, uint8_t *verdict_all, uint8_t *stage_all, void *scratch,
                                             size_t scratch_bytes, void *stream) {
    STWO_CHECK_DEVICE();
    if (!world || rank >= world || !verdict_local || !stage_local || !verdict_all || !stage_all) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t lo, hi;
    stwo_b200_shard_range(n_total, rank, world, &lo, &hi);
    const size_t mine = hi - lo, blk = (n_total + world - 1) / world;
    if (world == 1 && !comm) {                      // a single process without a communicator: a copy
        STWO_CUDA(cudaMemcpyAsync(verdict_all, verdict_local, mine, cudaMemcpyDeviceToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(stage_all, stage_local, mine, cudaMemcpyDeviceToDevice, st));
        return STWO_B200_OK;
    }
    if (!comm) return STWO_B200_E_BAD_ARG;
    if (!nccl_load()) return STWO_B200_E_NO_NCCL;
    if (!scratch || scratch_bytes < stwo_b200_gather_verdicts_scratch_bytes(world, n_total)) return STWO_B200_E_BAD_ARG;
    uint8_t *send = (uint8_t *)scratch, *recv = send + 2 * blk;
    STWO_CUDA(cudaMemsetAsync(send, 0, 2 * blk, st));
    STWO_CUDA(cudaMemcpyAsync(send, verdict_local, mine, cudaMemcpyDeviceToDevice, st));
    STWO_CUDA(cudaMemcpyAsync(send + blk, stage_local, mine, cudaMemcpyDeviceToDevice, st));
    if (g_nccl.AllGather(send, recv, 2 * blk, kNcclUint8, comm, st) != 0) return STWO_B200_E_NCCL;
    for (uint32_t r = 0; r < world; r++) {
        uint64_t rlo, rhi;
        stwo_b200_shard_range(n_total, r, world, &rlo, &rhi);
        if (rhi == rlo) continue;
        STWO_CUDA(cudaMemcpyAsync(verdict_all + rlo, recv + (size_t)r * 2 * blk, rhi - rlo, cudaMemcpyDeviceToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(stage_all + rlo, recv + (size_t)r * 2 * blk + blk, rhi - rlo, cudaMemcpyDeviceToDevice, st));
    }
    return STWO_B200_OK;
}
extern "C" size_t stwo_b200_gather_verdicts_scratch_bytes(uint32_t world, uint64_t n_total) {
    if (!world) return 0;
    const size_t blk = (n_total + world - 1) / world;
    return 2 * blk * ((size_t)world + 1);
}

// Trace columns of every rank's block on rank `dst` (north_star's second collective): rank r sends its block's
// [n_local][n_cols][n_rows] words, dst receives them in proof order into values_all.  Point-to-point inside one NCCL group.
extern "C" int32_t stwo_b200_gather_trace_columns(void *comm, uint32_t rank, uint32_t world, uint32_t dst, uint64_t n_total, size_t words_per_proof,
                                                  const uint32_t *values_local, uint32_t *values_all, void *stream) {
    STWO_CHECK_DEVICE();
    if (!world || rank >= world || dst >= world || !values_local || (rank == dst && !values_all)) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t lo, hi;
    stwo_b200_shard_range(n_total, rank, world, &lo, &hi);
    if (rank == dst && hi > lo)
        STWO_CUDA(cudaMemcpyAsync(values_all + lo * words_per_proof, values_local, (hi - lo) * words_per_proof * 4, cudaMemcpyDeviceToDevice, st));
    if (world == 1) return STWO_B200_OK;
    if (!comm) return STWO_B200_E_BAD_ARG;
    if (!nccl_load()) return STWO_B200_E_NO_NCCL;
    if (g_nccl.GroupStart() != 0) return STWO_B200_E_NCCL;
    int bad = 0;
    if (rank == dst) {
        for (uint32_t r = 0; r < world; r++) {
            if (r == dst) continue;
            uint64_t rlo, rhi;
            stwo_b200_shard_range(n_total, r, world, &rlo, &rhi);
            if (rhi > rlo) bad |= g_nccl.Recv(values_all + rlo * words_per_proof, (rhi - rlo) * words_per_proof, kNcclUint32, (int)r, comm, st);
        }
    } else if (hi > lo) bad |= g_nccl.Send(values_local, (hi - lo) * words_per_proof, kNcclUint32, (int)dst, comm, st);
    bad |= g_nccl.GroupEnd();
    return bad ? STWO_B200_E_NCCL : STWO_B200_OK;
}
