// K1 (batched Poseidon2 permutation), Merkle commit / decommit, K2 (batched path verification).
// One state / node / path per thread; see poseidon2.cuh and merkle.cuh for the arithmetic.
#include "common.cuh"
#include "merkle.cuh"

using namespace stwo_b200;

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ void load16(const u32 *p, u32 s[16]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    s[8] = c.x; s[9] = c.y; s[10] = c.z; s[11] = c.w; s[12] = d.x; s[13] = d.y; s[14] = d.z; s[15] = d.w;
}
__device__ __forceinline__ void store8(u32 *p, const u32 s[8]) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(s[0], s[1], s[2], s[3]);
    q[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// ---- K1 ------------------------------------------------------------------------------------------
template <bool UNROLLED>
__global__ void __launch_bounds__(kThreads) k_poseidon2_permute(u32 *states, size_t n) {
    size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (i >= n) return;
    u32 s[16];
    load16(states + 16 * i, s);
    poseidon2::permute<UNROLLED>(s);
    store8(states + 16 * i, s);
    store8(states + 16 * i + 8, s + 8);
}

// experiment shapes of K1 (stwo_b200_poseidon2_permute_dev_variant): two states per thread (instruction-level parallelism across two
// independent permutations), and the rolled shape under tighter register caps (more resident warps)
__global__ void __launch_bounds__(kThreads) k_poseidon2_permute_x2(u32 *states, size_t n) {
    size_t i = 2 * (blockIdx.x * (size_t)kThreads + threadIdx.x);
    if (i >= n) return;
    u32 a[16], b[16];
    load16(states + 16 * i, a);
    const bool two = i + 1 < n;
    load16(states + 16 * (two ? i + 1 : i), b);
    poseidon2::permute2(a, b);
    store8(states + 16 * i, a); store8(states + 16 * i + 8, a + 8);
    if (two) { store8(states + 16 * (i + 1), b); store8(states + 16 * (i + 1) + 8, b + 8); }
}
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, MIN_BLOCKS) k_poseidon2_permute_occ(u32 *states, size_t n) {
    size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (i >= n) return;
    u32 s[16];
    load16(states + 16 * i, s);
    poseidon2::permute<false>(s);
    store8(states + 16 * i, s);
    store8(states + 16 * i + 8, s + 8);
}

// ---- hash_node over a layer --------------------------------------------------------------------
// children: n x 16 or nullptr; cols column-major with stride col_stride
__global__ void __launch_bounds__(kThreads) k_hash_node_layer(const u32 *__restrict__ children,
                                                              const u32 *__restrict__ cols, u32 n_cols,
                                                              size_t col_stride, size_t n, u32 *__restrict__ out) {
    size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (i >= n) return;
    u32 kids[16];
    if (children) load16(children + 16 * i, kids);
    u32 h[8];
    const u32 *c0 = cols + i;
    merkle::hash_node(children ? kids : nullptr, [&](u32 c) { return __ldg(c0 + (size_t)c * col_stride); }, n_cols, h);
    store8(out + 8 * i, h);
}

// Commit: leaf layer of n_trees trees.  cols[(t*n_cols + c) << log_n | i]; nodes per tree
// (2^(log_n+1)-1)*8 words, layer k at (2^k - 1)*8.
__global__ void __launch_bounds__(kThreads) k_commit_leaves(const u32 *__restrict__ cols, u32 n_cols, u32 log_n,
                                                            u32 n_trees, u32 *__restrict__ nodes) {
    size_t g = blockIdx.x * (size_t)kThreads + threadIdx.x;
    size_t n = (size_t)1 << log_n;
    if (g >= n * n_trees) return;
    size_t t = g >> log_n, i = g & (n - 1);
    const u32 *c0 = cols + ((t * n_cols) << log_n) + i;
    u32 h[8];
    merkle::hash_node(nullptr, [&](u32 c) { return __ldg(c0 + ((size_t)c << log_n)); }, n_cols, h);
    u32 *tree = nodes + t * ((2 * n - 1) * 8);
    store8(tree + (n - 1) * 8 + 8 * i, h);
}

// one inner layer (size 2^k) of every tree
__global__ void __launch_bounds__(kThreads) k_commit_layer(u32 log_n, u32 k, u32 n_trees, u32 *__restrict__ nodes) {
    size_t g = blockIdx.x * (size_t)kThreads + threadIdx.x;
    size_t m = (size_t)1 << k, n = (size_t)1 << log_n;
    if (g >= m * n_trees) return;
    size_t t = g >> k, i = g & (m - 1);
    u32 *tree = nodes + t * ((2 * n - 1) * 8);
    u32 s[16];
    load16(tree + (2 * m - 1) * 8 + 16 * i, s);
    poseidon2::permute<false>(s);
    store8(tree + (m - 1) * 8 + 8 * i, s);
}

// the top kTopLog layers of every tree, one CTA per tree, layers staged in shared memory
constexpr int kTopLog = 7;   // 128 nodes in, root out
__global__ void __launch_bounds__(64) k_commit_top(u32 log_n, u32 top_log, u32 *__restrict__ nodes) {
    __shared__ u32 sh[(1 << kTopLog) * 8];
    size_t n = (size_t)1 << log_n;
    u32 *tree = nodes + blockIdx.x * ((2 * n - 1) * 8);
    const u32 m0 = 1u << top_log;
    for (u32 w = threadIdx.x; w < m0 * 8; w += blockDim.x) sh[w] = tree[(m0 - 1) * 8 + w];
    __syncthreads();
    for (int k = (int)top_log - 1; k >= 0; k--) {
        u32 m = 1u << k;
        u32 s[16];
        bool act = threadIdx.x < m;
        if (act) {
#pragma unroll
            for (int j = 0; j < 16; j++) s[j] = sh[16 * threadIdx.x + j];
            poseidon2::permute<false>(s);
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int j = 0; j < 8; j++) sh[8 * threadIdx.x + j] = s[j];
            store8(tree + (m - 1) * 8 + 8 * threadIdx.x, s);
        }
        __syncthreads();
    }
}

// Decommit single-size trees: gather leaf values and siblings into the per-path layout.
__global__ void k_decommit(const u32 *__restrict__ cols, u32 n_cols, u32 log_n, u32 n_trees,
                           const u32 *__restrict__ nodes, const u32 *__restrict__ index, u32 n_queries,
                           u32 *__restrict__ path_cols, u32 *__restrict__ path_sib) {
    size_t p = blockIdx.x;   // one CTA per path
    if (p >= (size_t)n_trees * n_queries) return;
    size_t t = p / n_queries;
    size_t n = (size_t)1 << log_n;
    u32 idx = index[p];
    const u32 *tree = nodes + t * ((2 * n - 1) * 8);
    for (u32 c = threadIdx.x; c < n_cols; c += blockDim.x)
        path_cols[p * n_cols + c] = cols[((t * n_cols + c) << log_n) + idx];
    for (u32 w = threadIdx.x; w < log_n * 8; w += blockDim.x) {
        u32 lvl = w >> 3, j = w & 7;
        size_t layer = (size_t)1 << (log_n - lvl);          // size of the layer the sibling lives in
        size_t pos = (idx >> lvl) ^ 1u;
        path_sib[p * log_n * 8 + w] = tree[(layer - 1) * 8 + pos * 8 + j];
    }
}

// ---- K2 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_path_verify(const stwo_b200_path_shape shape, u32 cols_per_path,
                                                          size_t n_paths, const u32 *__restrict__ index,
                                                          const u32 *__restrict__ cols, const u32 *__restrict__ sib,
                                                          const u32 *__restrict__ roots, const u32 *__restrict__ root_id,
                                                          uint8_t *__restrict__ verdict, u32 *__restrict__ computed) {
    size_t p = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (p >= n_paths) return;
    u32 r[8];
    merkle::path_root(shape, index[p], cols + p * cols_per_path, sib + p * shape.depth * 8, r);
    const u32 *want = roots + (root_id ? (size_t)root_id[p] * 8 : 0);
    u32 diff = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) diff |= r[i] ^ want[i];
    verdict[p] = diff == 0;
    if (computed) store8(computed + 8 * p, r);
}

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }
u32 shape_cols(const stwo_b200_path_shape &s) {
    u32 n = 0;
    for (u32 h = 0; h <= s.depth; h++) n += s.n_cols[h];
    return n;
}
}  // namespace

// =================================== C ABI ==========================================================
extern "C" int32_t stwo_b200_poseidon2_permute_dev_variant(uint32_t *states, size_t n, int32_t variant, void *stream) {
    STWO_CHECK_DEVICE();
    if (n == 0) return STWO_B200_OK;
    if (!states || ((uintptr_t)states & 15)) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // 10 + v / 20 + v: variant v (0 or 2) with the residency cut to 2 / 3 blocks per SM by a dynamic shared-memory request -- how the
    // permutation rate depends on resident warps, with one and with two states per thread (the tree rebuilds run at 16 warps per SM)
    if (variant >= 10) {
        const int base = variant % 10;
        const size_t smem = variant < 20 ? 100 * 1024 : 64 * 1024;
        if (base == 2) {
            STWO_CUDA(cudaFuncSetAttribute(k_poseidon2_permute_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_poseidon2_permute_x2<<<blocks_for((n + 1) / 2), kThreads, smem, st>>>(states, n);
        } else {
            STWO_CUDA(cudaFuncSetAttribute(k_poseidon2_permute<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_poseidon2_permute<false><<<blocks_for(n), kThreads, smem, st>>>(states, n);
        }
        note_launch();
        return cuda_status(cudaGetLastError());
    }
    if (variant == 2) k_poseidon2_permute_x2<<<blocks_for((n + 1) / 2), kThreads, 0, st>>>(states, n);
    else if (variant == 3) k_poseidon2_permute_occ<8><<<blocks_for(n), kThreads, 0, st>>>(states, n);
    else if (variant == 4) k_poseidon2_permute_occ<10><<<blocks_for(n), kThreads, 0, st>>>(states, n);
    else if (variant == 5) k_poseidon2_permute_occ<5><<<blocks_for(n), kThreads, 0, st>>>(states, n);
    else if (variant == 1) k_poseidon2_permute<true><<<blocks_for(n), kThreads, 0, st>>>(states, n);
    else k_poseidon2_permute<false><<<blocks_for(n), kThreads, 0, st>>>(states, n);
    note_launch();
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_poseidon2_permute_dev(uint32_t *states, size_t n, void *stream) {
    return stwo_b200_poseidon2_permute_dev_variant(states, n, 0, stream);
}
extern "C" int32_t stwo_b200_poseidon2_permute(uint32_t *states, size_t n) {
    STWO_CHECK_DEVICE();
    if (n == 0) return STWO_B200_OK;
    if (!states) return STWO_B200_E_BAD_ARG;
    size_t bytes = n * 64;
    int32_t rc = stage_reserve(bytes);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    u32 *d = (u32 *)stage_dev();
    STWO_CUDA(cudaMemcpyAsync(d, states, bytes, cudaMemcpyHostToDevice, st));
    rc = stwo_b200_poseidon2_permute_dev(d, n, st);
    if (rc) return rc;
    STWO_CUDA(cudaMemcpyAsync(states, d, bytes, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

extern "C" int32_t stwo_b200_hash_node_batch_dev(const uint32_t *children, const uint32_t *cols, uint32_t n_cols,
                                                 size_t col_stride, size_t n, uint32_t *out, void *stream) {
    STWO_CHECK_DEVICE();
    if (n == 0) return STWO_B200_OK;
    if (!out || (!children && !cols) || (n_cols && !cols)) return STWO_B200_E_BAD_ARG;
    k_hash_node_layer<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(children, cols, n_cols, col_stride, n, out);
    note_launch();
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_hash_node_batch(const uint32_t *children, const uint32_t *cols, uint32_t n_cols,
                                             size_t col_stride, size_t n, uint32_t *out) {
    STWO_CHECK_DEVICE();
    if (n == 0) return STWO_B200_OK;
    size_t b_kids = children ? align_up(n * 64, 256) : 0;
    size_t b_cols = n_cols ? align_up(((size_t)(n_cols - 1) * col_stride + n) * 4, 256) : 0;
    size_t b_out = n * 32;
    int32_t rc = stage_reserve(b_kids + b_cols + b_out);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    uint8_t *d = stage_dev();
    u32 *d_kids = children ? (u32 *)d : nullptr, *d_cols = n_cols ? (u32 *)(d + b_kids) : nullptr, *d_out = (u32 *)(d + b_kids + b_cols);
    if (children) STWO_CUDA(cudaMemcpyAsync(d_kids, children, n * 64, cudaMemcpyHostToDevice, st));
    if (n_cols) STWO_CUDA(cudaMemcpyAsync(d_cols, cols, ((size_t)(n_cols - 1) * col_stride + n) * 4, cudaMemcpyHostToDevice, st));
    rc = stwo_b200_hash_node_batch_dev(d_kids, d_cols, n_cols, col_stride, n, d_out, st);
    if (rc) return rc;
    STWO_CUDA(cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

extern "C" int32_t stwo_b200_merkle_commit_dev(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                               uint32_t *nodes, void *stream) {
    STWO_CHECK_DEVICE();
    if (n_trees == 0) return STWO_B200_OK;
    if (!cols || !nodes || n_cols == 0 || log_n >= STWO_B200_MAX_DEPTH) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)1 << log_n;
    k_commit_leaves<<<blocks_for(n * n_trees), kThreads, 0, st>>>(cols, n_cols, log_n, n_trees, nodes);
    note_launch();
    u32 top_log = log_n < (u32)kTopLog ? log_n : (u32)kTopLog;
    for (int k = (int)log_n - 1; k >= (int)top_log; k--) {
        k_commit_layer<<<blocks_for(((size_t)1 << k) * n_trees), kThreads, 0, st>>>(log_n, (u32)k, n_trees, nodes);
        note_launch();
    }
    if (top_log > 0) {
        k_commit_top<<<n_trees, 64, 0, st>>>(log_n, top_log, nodes);
        note_launch();
    }
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_merkle_commit(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                           uint32_t *roots) {
    STWO_CHECK_DEVICE();
    if (n_trees == 0) return STWO_B200_OK;
    if (!cols || !roots || n_cols == 0 || log_n >= STWO_B200_MAX_DEPTH) return STWO_B200_E_BAD_ARG;
    size_t n = (size_t)1 << log_n;
    size_t b_cols = align_up((size_t)n_trees * n_cols * n * 4, 256);
    size_t per_tree = (2 * n - 1) * 8;
    size_t b_nodes = (size_t)n_trees * per_tree * 4;
    int32_t rc = stage_reserve(b_cols + b_nodes);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    u32 *d_cols = (u32 *)stage_dev(), *d_nodes = (u32 *)(stage_dev() + b_cols);
    STWO_CUDA(cudaMemcpyAsync(d_cols, cols, (size_t)n_trees * n_cols * n * 4, cudaMemcpyHostToDevice, st));
    rc = stwo_b200_merkle_commit_dev(d_cols, n_cols, log_n, n_trees, d_nodes, st);
    if (rc) return rc;
    STWO_CUDA(cudaMemcpy2DAsync(roots, 32, d_nodes, per_tree * 4, 32, n_trees, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}

extern "C" int32_t stwo_b200_merkle_decommit_dev(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                                 const uint32_t *nodes, const uint32_t *index, uint32_t n_queries,
                                                 uint32_t *path_cols, uint32_t *path_siblings, void *stream) {
    STWO_CHECK_DEVICE();
    size_t n_paths = (size_t)n_trees * n_queries;
    if (n_paths == 0) return STWO_B200_OK;
    if (!cols || !nodes || !index || !path_cols || !path_siblings) return STWO_B200_E_BAD_ARG;
    k_decommit<<<(unsigned)n_paths, 64, 0, (cudaStream_t)stream>>>(cols, n_cols, log_n, n_trees, nodes, index, n_queries,
                                                                     path_cols, path_siblings);
    note_launch();
    return cuda_status(cudaGetLastError());
}

extern "C" uint32_t stwo_b200_path_perms(const stwo_b200_path_shape *shape) { return merkle::path_perms(*shape); }

extern "C" int32_t stwo_b200_merkle_path_verify_dev(const stwo_b200_path_shape *shape, size_t n_paths,
                                                    const uint32_t *index, const uint32_t *cols, const uint32_t *siblings,
                                                    const uint32_t *roots, const uint32_t *root_id,
                                                    uint8_t *verdict, uint32_t *computed_roots, void *stream) {
    STWO_CHECK_DEVICE();
    if (n_paths == 0) return STWO_B200_OK;
    if (!shape || shape->depth > STWO_B200_MAX_DEPTH) return STWO_B200_E_SHAPE;
    if (!index || !cols || (!siblings && shape->depth) || !roots || !verdict) return STWO_B200_E_BAD_ARG;
    k_path_verify<<<blocks_for(n_paths), kThreads, 0, (cudaStream_t)stream>>>(*shape, shape_cols(*shape), n_paths, index, cols,
                                                                              siblings, roots, root_id, verdict, computed_roots);
    note_launch();
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_merkle_path_verify(const stwo_b200_path_shape *shape, size_t n_paths,
                                                const uint32_t *index, const uint32_t *cols, const uint32_t *siblings,
                                                const uint32_t *roots, size_t n_roots, const uint32_t *root_id,
                                                uint8_t *verdict, uint32_t *computed_roots) {
    STWO_CHECK_DEVICE();
    if (n_paths == 0) return STWO_B200_OK;
    if (!shape || shape->depth > STWO_B200_MAX_DEPTH) return STWO_B200_E_SHAPE;
    if (!index || !cols || !roots || !verdict || n_roots == 0) return STWO_B200_E_BAD_ARG;
    u32 cpp = shape_cols(*shape);
    size_t b_idx = align_up(n_paths * 4, 256), b_cols = align_up(n_paths * cpp * 4, 256),
           b_sib = align_up(n_paths * shape->depth * 32, 256), b_roots = align_up(n_roots * 32, 256),
           b_rid = root_id ? align_up(n_paths * 4, 256) : 0, b_ver = align_up(n_paths, 256), b_comp = n_paths * 32;
    int32_t rc = stage_reserve(b_idx + b_cols + b_sib + b_roots + b_rid + b_ver + b_comp);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    uint8_t *d = stage_dev();
    u32 *d_idx = (u32 *)d; d += b_idx;
    u32 *d_cols = (u32 *)d; d += b_cols;
    u32 *d_sib = (u32 *)d; d += b_sib;
    u32 *d_roots = (u32 *)d; d += b_roots;
    u32 *d_rid = root_id ? (u32 *)d : nullptr; d += b_rid;
    uint8_t *d_ver = d; d += b_ver;
    u32 *d_comp = (u32 *)d;
    STWO_CUDA(cudaMemcpyAsync(d_idx, index, n_paths * 4, cudaMemcpyHostToDevice, st));
    STWO_CUDA(cudaMemcpyAsync(d_cols, cols, n_paths * cpp * 4, cudaMemcpyHostToDevice, st));
    if (shape->depth) STWO_CUDA(cudaMemcpyAsync(d_sib, siblings, n_paths * shape->depth * 32, cudaMemcpyHostToDevice, st));
    STWO_CUDA(cudaMemcpyAsync(d_roots, roots, n_roots * 32, cudaMemcpyHostToDevice, st));
    if (root_id) STWO_CUDA(cudaMemcpyAsync(d_rid, root_id, n_paths * 4, cudaMemcpyHostToDevice, st));
    rc = stwo_b200_merkle_path_verify_dev(shape, n_paths, d_idx, d_cols, d_sib, d_roots, d_rid, d_ver, d_comp, st);
    if (rc) return rc;
    STWO_CUDA(cudaMemcpyAsync(verdict, d_ver, n_paths, cudaMemcpyDeviceToHost, st));
    if (computed_roots) STWO_CUDA(cudaMemcpyAsync(computed_roots, d_comp, n_paths * 32, cudaMemcpyDeviceToHost, st));
    return cuda_status(cudaStreamSynchronize(st));
}
