/* stwo_b200 — C ABI of the B200-native (sm_100a) verifier hot path of recursive-stwo.
 *
 * This is the drop-in boundary: plain pointers and sizes, no allocation visible to the caller,
 * int32 status returns (0 = ok, < 0 = -(cudaError_t) or STWO_B200_E_*).  The reference has no FFI
 * of its own (pure Rust); each entry point cites the reference call sites whose
 * value arithmetic it replaces, and INTEGRATION.md shows the Rust `extern "C"` block + build.rs a
 * maintainer adds.
 *
 * Conventions
 *  - M31 words are canonical u32 (< 2^31-1).  QM31 = 4 consecutive words
 *    [a.0.0, a.0.1, a.1.0, a.1.1] (= QM31::to_m31_array, constraint_system/src/plonk_with_poseidon.rs:473-478).
 *  - A hash is 8 M31 words (Poseidon31Hash).
 *  - `*_dev` entry points take DEVICE pointers and a cudaStream_t (as void*) and only enqueue work;
 *    the un-suffixed entry points take HOST pointers, stage through pinned buffers owned by the
 *    library and return after the result is in the caller's buffer.
 *  - One host thread per device; the library is re-entrant across streams for `*_dev` calls.
 *  - There is no CPU fallback: without a CUDA device every call returns STWO_B200_E_NO_DEVICE.
 */
#ifndef STWO_B200_H
#define STWO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STWO_B200_OK 0
#define STWO_B200_E_NO_DEVICE (-1000)
#define STWO_B200_E_BAD_ARG (-1001)
#define STWO_B200_E_SHAPE (-1002)
#define STWO_B200_MAX_DEPTH 32

/* library / device lifecycle.  init selects the device for the calling thread and creates the
 * staging buffers lazily; version = 0x00MMmmpp. */
int32_t stwo_b200_init(int32_t device);
int32_t stwo_b200_shutdown(void);
uint32_t stwo_b200_version(void);
/* number of kernels the library has launched since init (bench.py's gpu_launches counter) */
uint64_t stwo_b200_launch_count(void);

/* ---- K1: batched Poseidon2-M31 width-16 permutation -------------------------------------------
 * states: n x 16 words, permuted in place.
 * Replaces poseidon2_permute at primitives/poseidon31/src/implementation.rs:108-149 as called from
 * Poseidon2HalfVar::permute (primitives/poseidon31/src/lib.rs:311),
 * check_poseidon_invocations (constraint_system/src/plonk_with_poseidon.rs:514) and the
 * Poseidon31CRH::permute_get_rate sites (components/hints/src/folding.rs:86,
 * components/last/fiat_shamir/src/lib.rs:53). */
int32_t stwo_b200_poseidon2_permute(uint32_t *states, size_t n);
int32_t stwo_b200_poseidon2_permute_dev(uint32_t *states, size_t n, void *stream);
/* code-shape selector for benchmarking: variant 0 = rolled rounds (default), 1 = fully unrolled */
int32_t stwo_b200_poseidon2_permute_dev_variant(uint32_t *states, size_t n, int32_t variant, void *stream);

/* ---- Merkle hashing (stwo Poseidon31MerkleHasher::hash_node) -------------------------------------
 * hash_node(children?, column values) for n nodes.  children: n x 16 words (left||right) or NULL
 * for a leaf layer; cols: column-major, column c of node i at cols[c*col_stride + i], n_cols may be
 * 0 when children != NULL; out: n x 8 words.
 * Replaces Poseidon31MerkleHasher::hash_node as restated at primitives/merkle/src/lib.rs:9-181 and
 * called at components/hints/src/decommit.rs:23,31,82,119,128 and folding.rs:34,41,53,65,144,167,198,258. */
int32_t stwo_b200_hash_node_batch_dev(const uint32_t *children, const uint32_t *cols, uint32_t n_cols,
                                      size_t col_stride, size_t n, uint32_t *out, void *stream);
int32_t stwo_b200_hash_node_batch(const uint32_t *children, const uint32_t *cols, uint32_t n_cols,
                                  size_t col_stride, size_t n, uint32_t *out);

/* Commit n_trees Merkle trees of 2^log_n leaves each over n_cols column-major columns
 * (cols[(t*n_cols + c) << log_n | i]).  nodes: per tree (2^(log_n+1) - 1) * 8 words; the layer of
 * size 2^k starts at word offset (2^k - 1) * 8 (root first).  This is the commit step that precedes
 * the verifier path (stwo MerkleProver::commit) and the generator of the synthetic
 * decommitment sweep (BASELINE.json configs[2]). */
int32_t stwo_b200_merkle_commit_dev(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                    uint32_t *nodes, void *stream);
int32_t stwo_b200_merkle_commit(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                uint32_t *roots /* n_trees x 8 */);

/* Prover-side decommitment of single-size trees built by merkle_commit: for query q of tree t
 * (index[t*n_queries + q] < 2^log_n) writes the leaf's column values (n_cols words) and the log_n
 * sibling hashes, leaf level first, in the per-path layout merkle_path_verify consumes. */
int32_t stwo_b200_merkle_decommit_dev(const uint32_t *cols, uint32_t n_cols, uint32_t log_n, uint32_t n_trees,
                                      const uint32_t *nodes, const uint32_t *index, uint32_t n_queries,
                                      uint32_t *path_cols, uint32_t *path_siblings, void *stream);

/* Shape of the per-query authentication paths of one tree (mixed-degree trees inject columns at
 * inner layers; reference components/recursive/data_structures/src/lib.rs:315-354). */
typedef struct {
    uint32_t depth;                               /* number of sibling levels = log size of leaf layer */
    uint32_t n_cols[STWO_B200_MAX_DEPTH + 1];     /* n_cols[h]: columns hashed at the layer of log size h */
} stwo_b200_path_shape;

/* ---- K2: batched Merkle authentication-path verification ----------------------------------------
 * n_paths independent leaf->root recomputations of one shape.  Per path p:
 *   index[p]                       leaf position (bit i selects left/right at level i from the leaf)
 *   cols + p*cols_per_path         column values, leaf layer first then each injected layer downwards
 *   siblings + p*depth*8           sibling hashes, leaf level first
 *   roots + root_id[p]*8           expected root (root_id may be NULL => root 0 for all)
 * Outputs: verdict[p] = 1 if the recomputed root equals the expected one else 0; computed_roots
 * (n_paths x 8, may be NULL).
 * Replaces SinglePathMerkleProof::verify (components/hints/src/decommit.rs:22-42) and the value side
 * of SinglePathMerkleProofVar::verify (components/recursive/data_structures/src/lib.rs:315-354). */
int32_t stwo_b200_merkle_path_verify_dev(const stwo_b200_path_shape *shape, size_t n_paths,
                                         const uint32_t *index, const uint32_t *cols, const uint32_t *siblings,
                                         const uint32_t *roots, const uint32_t *root_id,
                                         uint8_t *verdict, uint32_t *computed_roots, void *stream);
int32_t stwo_b200_merkle_path_verify(const stwo_b200_path_shape *shape, size_t n_paths,
                                     const uint32_t *index, const uint32_t *cols, const uint32_t *siblings,
                                     const uint32_t *roots, size_t n_roots, const uint32_t *root_id,
                                     uint8_t *verdict, uint32_t *computed_roots);

/* number of Poseidon2 permutations one path of this shape costs */
uint32_t stwo_b200_path_perms(const stwo_b200_path_shape *shape);

#ifdef __cplusplus
}
#endif
#endif
