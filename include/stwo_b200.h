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
 *  - One host thread per device.  `*_dev` calls only enqueue on the given stream, and calls on different streams run beside each other
 *    (the shape groups of a mixed batch, the stages of a pipeline): stwo_b200_verify_proofs_batch[_pinned]_dev keeps one pool of worker
 *    streams per CALLER stream (six pools; further caller streams share the last one, which only serialises them), and the tape evaluation
 *    is a plain thread-block-cluster launch.  Only the cooperative-grid evaluation (STWO_B200_EVAL_MODE=grid, profiling) keeps a
 *    grid-barrier word per launch slot: one of those in flight at a time.
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
#define STWO_B200_E_NO_NCCL (-1003)   /* libnccl.so.2 could not be loaded (multi-GPU entry points only) */
#define STWO_B200_E_NCCL (-1004)      /* an NCCL call failed */
#define STWO_B200_MAX_DEPTH 32

/* library / device lifecycle.  init selects the device for the calling thread and creates the
 * staging buffers lazily; version = 0x00MMmmpp.  ONE DEVICE PER PROCESS (one process per GPU is the multi-GPU model): init on a
 * second device returns STWO_B200_E_BAD_ARG; init on the same device again is a no-op.  The host-pointer entry points share one
 * staging area and are not re-entrant: call them from one host thread at a time. */
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

/* ---- K3-K5 + driver: batched native verification of PlonkWithPoseidon proofs ----------------------------------------
 * Replaces, for Poseidon31-hash proofs, the value side of the whole verifier run of
 * examples/single-proof/src/main.rs:33-83 and examples/multi-proofs/src/main.rs:49-139:
 *   FiatShamirHints::new / FiatShamirResults::compute   components/hints/src/fiat_shamir.rs:69-307,
 *                                                       components/recursive/fiat_shamir/src/lib.rs:31-176
 *   CompositionCheck::compute                           components/recursive/composition/src/lib.rs:33-121
 *   DecommitHints::compute                              components/hints/src/decommit.rs:194-241
 *   AnswerHints / AnswerResults::compute                components/recursive/answer/src/lib.rs:34-382
 *   FirstLayerHints / InnerLayersHints::compute         components/hints/src/folding.rs:296-601
 *   FoldingResults::compute                             components/recursive/folding/src/lib.rs:12-205
 * A blob is the bincode serialisation of PlonkWithPoseidonProof<Poseidon31MerkleHasher> (little endian, 4-byte aligned).
 * verdict: 0 accept, 1 reject, 2 unsupported (the reference panics: duplicated queries at the largest domain,
 * components/recursive/answer/src/lib.rs:190-195).  stage: first failing STWO_B200_STAGE_*. */
typedef struct {
    uint32_t log_size_plonk, log_size_poseidon;     /* stmt0 */
    uint32_t pow_bits, log_blowup, log_last;        /* PcsConfig / FriConfig */
    uint32_t n_queries, n_inner;                    /* FriConfig.n_queries, number of inner FRI layers */
} stwo_b200_proof_shape;

/* The commitment-scheme parameters the CALLER fixes (stwo PcsConfig { pow_bits, fri_config: FriConfig { log_blowup_factor,
 * log_last_layer_degree_bound, n_queries } }), the `config` argument of FiatShamirHints::new / FiatShamirResults::compute
 * (components/hints/src/fiat_shamir.rs:69-73, components/recursive/fiat_shamir/src/lib.rs:31-38).  A proof is verified under
 * the caller's config, never under the one its own header claims: a blob whose header differs is REJECTed at STAGE_PARSE. */
typedef struct {
    uint32_t pow_bits, log_blowup, log_last, n_queries;
} stwo_b200_pcs_config;
/* The batch shape the caller's config implies for proofs of the given component log sizes (the statement, stmt0): n_inner
 * follows from the FRI layer relation stwo enforces (InvalidNumFriLayers): the composition columns live at log size
 * max(log_size_plonk + 1, log_size_poseidon + 2) + log_blowup, which n_inner + 1 folds must bring down to log_last +
 * log_blowup.  STWO_B200_E_SHAPE when no such shape exists or a bound (log sizes 1..28, blow-up 1..16, log_last <= 12,
 * n_queries 1..128) is exceeded. */
int32_t stwo_b200_shape_from_config(const stwo_b200_pcs_config *config, uint32_t log_size_plonk, uint32_t log_size_poseidon,
                                    stwo_b200_proof_shape *out);

#define STWO_B200_VERDICT_ACCEPT 0
#define STWO_B200_VERDICT_REJECT 1
#define STWO_B200_VERDICT_UNSUPPORTED 2
#define STWO_B200_STAGE_OK 0
#define STWO_B200_STAGE_PARSE 1
#define STWO_B200_STAGE_POW 2
#define STWO_B200_STAGE_LOGUP 3
#define STWO_B200_STAGE_OODS 4
#define STWO_B200_STAGE_MERKLE 5
#define STWO_B200_STAGE_FRI_FIRST 6
#define STWO_B200_STAGE_FRI_INNER 7
#define STWO_B200_STAGE_FRI_LAST 8
#define STWO_B200_STAGE_UNSUPPORTED 9

/* flags */
#define STWO_B200_VERIFY_FULL 1u   /* also produce what the verifier circuit consumes beyond the verdict: the per-query roots and the
                                      permutation record (output state of every transcript / authentication-path permutation) */
#define STWO_B200_VERIFY_PATH_KERNELS 8u /* with FULL: produce that record by hashing every per-query path again from its hints
                                      (SinglePathMerkleProof::verify, components/hints/src/decommit.rs:22-42), one thread per
                                      path, instead of taking it from the tree rebuilds, which hash every node once
                                      (from_stwo_proof, decommit.rs:44-183).  Same record; kept as the checker of the default. */
#define STWO_B200_VERIFY_ONE_STREAM 4u /* do not slice the batch over the library's stream pool */
#define STWO_B200_VERIFY_TIMED 2u  /* record CUDA events between the stage kernels (read with stwo_b200_verify_stage_ms) */
/* stop after a stage (the stage-level entry points below use these; the batch then runs on one stream) */
#define STWO_B200_VERIFY_UPTO_TRANSCRIPT 0x100u  /* parse, transcript, PoW, query draws, logup sum, OODS */
#define STWO_B200_VERIFY_UPTO_ANSWERS 0x200u     /* + commitment-tree decommitments, DEEP quotient answers */
#define STWO_B200_VERIFY_UPTO_FOLDS 0x400u       /* + circle / line folds, last-layer evaluation */
/* stage kernels in launch order: fiat_shamir, single_tree, group, answer, folds, pair_tree, single_path, pair_path, verdict */
#define STWO_B200_N_STAGE_KERNELS 9
/* device time of each stage kernel of the last STWO_B200_VERIFY_TIMED batch on this thread (synchronises) */
int32_t stwo_b200_verify_stage_ms(float *ms /* [STWO_B200_N_STAGE_KERNELS] */);

/* host-side header read (no device needed): the shape one blob CLAIMS; STWO_B200_E_SHAPE when it does not parse or breaks the
 * FRI layer relation.  Untrusted: use it to read the statement's log sizes and to group blobs, and take the PcsConfig part of
 * the shape you verify under from your own configuration (stwo_b200_shape_from_config). */
int32_t stwo_b200_proof_shape_of(const uint8_t *blob, size_t len, stwo_b200_proof_shape *out);
/* bytes of device workspace a batch of n_proofs of this shape needs (0: unsupported shape) */
size_t stwo_b200_verify_workspace_bytes(const stwo_b200_proof_shape *shape, uint32_t n_proofs);
/* slots of the permutation record of one proof of this shape (512 transcript slots, then every per-query path) */
uint32_t stwo_b200_proof_record_slots(const stwo_b200_proof_shape *shape);
/* Poseidon2 permutations of the per-query paths of one proof of this shape (transcript excluded) */
uint64_t stwo_b200_proof_perms(const stwo_b200_proof_shape *shape);

/* Device entry: blobs = all proofs back to back as u32 words, blob_off = n_proofs+1 WORD offsets, every proof of the
 * same shape -- the CALLER's shape (stwo_b200_shape_from_config); a blob whose header says otherwise is rejected at STAGE_PARSE.  input_idx / input_vals: the (index, QM31 value) public inputs of the logup sum
 * (components/recursive/fiat_shamir/src/lib.rs:133-141).  verdict/stage: n_proofs bytes each (stage may be NULL). */
int32_t stwo_b200_verify_proofs_batch_dev(const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                          const stwo_b200_proof_shape *shape, const uint32_t *input_idx,
                                          const uint32_t *input_vals, uint32_t n_inputs, uint32_t flags,
                                          void *workspace, size_t workspace_bytes, uint8_t *verdict, uint8_t *stage,
                                          void *stream);
/* Same, with the blobs still in PINNED HOST memory: host_blobs / host_blob_off mirror blobs / blob_off (which must be
 * device buffers of the same sizes).  The library copies each slice of the batch to the device on the stream that verifies
 * it, so the host->device transfer of one slice overlaps the kernels of the others (the copy is inside the call). */
int32_t stwo_b200_verify_proofs_batch_pinned_dev(const uint32_t *host_blobs, const uint64_t *host_blob_off, uint32_t *blobs,
                                                 uint64_t *blob_off, uint32_t n_proofs, const stwo_b200_proof_shape *shape,
                                                 const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                                 uint32_t flags, void *workspace, size_t workspace_bytes, uint8_t *verdict,
                                                 uint8_t *stage, void *stream);
/* Host entry: host blobs of any mix of shapes (grouped internally), verdicts back in the caller's arrays.  configs: the
 * PcsConfigs the caller accepts (n_configs >= 1; e.g. the six of examples/multi-proofs/src/main.rs:173-196); a blob whose
 * header names any other config is rejected at STAGE_PARSE without being looked at further. */
int32_t stwo_b200_verify_proofs_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs,
                                      const stwo_b200_pcs_config *configs, uint32_t n_configs,
                                      const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                      uint32_t flags, uint8_t *verdict, uint8_t *stage);

/* Per-proof intermediate values kept in the workspace after a batch (what a prover of the NEXT recursion level, or a
 * parity test, reads back): every Fiat-Shamir draw, the OODS values, the counters. */
typedef struct {
    struct {
        uint32_t z[4], alpha[4], random_coeff[4], oods_t[4], oods_x[4], oods_y[4], after_coeff[4];
        uint32_t fri_alphas[33][4];
        uint32_t digest_after_nonce[8];
        uint32_t raw_queries[128];
        uint32_t n_transcript_perms, pow_ok;
    } fs;
    uint32_t oods_computed[4], oods_expected[4];
    uint32_t n_logs, log_sizes[3];          /* committed column log sizes, descending */
    uint32_t fail_mask, verdict, stage;
    uint32_t n_perms_hints, n_perms_paths;
} stwo_b200_verify_detail;

#define STWO_B200_FETCH_DETAIL 0          /* stwo_b200_verify_detail */
#define STWO_B200_FETCH_DOMAIN_POINTS 1   /* [3][n_queries][2]   circle-domain point of each query per log size */
#define STWO_B200_FETCH_ANSWERS 2         /* [3][n_queries][4]   DEEP quotient answers */
#define STWO_B200_FETCH_CIRCLE_FOLDS 3    /* [3][n_queries][4] */
#define STWO_B200_FETCH_LINE_FOLDS 4      /* [32][n_queries][4]  value after each inner layer */
#define STWO_B200_FETCH_LAST_EVALS 5      /* [n_queries][4] */
#define STWO_B200_FETCH_PATH_ROOTS 6      /* [4 + 1 + n_inner][n_queries][8]  (STWO_B200_VERIFY_FULL) */
#define STWO_B200_FETCH_PATH_COLS 7       /* [4][n_queries][64]  per-query column values of the commitment trees */
#define STWO_B200_FETCH_PATH_SIBLINGS 8   /* [4][n_queries][30][8] */
#define STWO_B200_FETCH_PAIR_HINTS 9      /* [1 + n_inner][n_queries * 256] packed FRI pair hints */
#define STWO_B200_FETCH_PERM_RECORD 10    /* [stwo_b200_proof_record_slots][16] the permutation record (STWO_B200_VERIFY_FULL); slots the
                                             shape does not use (transcript slots beyond n_transcript_perms) are not written */
#define STWO_B200_FETCH_RECORD_TREES 11   /* u32: trees of this proof whose part of the record is complete (4 + 1 + n_inner = all) */
#define STWO_B200_FETCH_PERM_RECORD_INPUTS 12   /* the INPUT state of every recorded permutation, same slots as STWO_B200_FETCH_PERM_RECORD */
int32_t stwo_b200_verify_fetch(const void *workspace, const stwo_b200_proof_shape *shape, uint32_t n_proofs, uint32_t p,
                               uint32_t what, void *out, size_t out_bytes, void *stream);

/* every proof of the batch at once (what = DETAIL .. LAST_EVALS): the per-proof arrays above, n_proofs of them back to back */
int32_t stwo_b200_verify_fetch_batch(const void *workspace, const stwo_b200_proof_shape *shape, uint32_t n_proofs, uint32_t what,
                                     void *out, size_t out_bytes, void *stream);

/* ---- stage-level entry points: replace ONE hint stage of the reference for a batch (host blobs in, host arrays out) ----------
 * All proofs must share the shape config + the statement log sizes of the first parsable blob imply; others are rejected at
 * STAGE_PARSE.  verdict / stage (n_proofs bytes each, may be NULL) report what failed UP TO the stage that was run.
 *
 * FiatShamirHints::new (components/hints/src/fiat_shamir.rs:69-307) + FiatShamirResults::compute's values
 * (components/recursive/fiat_shamir/src/lib.rs:31-176): transcript replay, PoW, query draws, logup sum, OODS check. */
int32_t stwo_b200_channel_replay_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                       const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs,
                                       stwo_b200_verify_detail *out /* n_proofs */, uint8_t *verdict, uint8_t *stage);
/* AnswerHints::compute (components/hints/src/answer.rs:40-48) / AnswerResults::compute values
 * (components/recursive/answer/src/lib.rs:34-382): answers n_proofs x 3 x n_queries x 4 words (log-size groups descending),
 * domain_points (optional) n_proofs x 3 x n_queries x 2. */
int32_t stwo_b200_fri_answers_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                    const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs, uint32_t *answers,
                                    uint32_t *domain_points, uint8_t *verdict, uint8_t *stage);
/* FirstLayerHints / InnerLayersHints::compute (components/hints/src/folding.rs:326-363,481-595) / FoldingResults::compute values
 * (components/recursive/folding/src/lib.rs:12-205): circle_folds n x 3 x n_queries x 4, line_folds n x 32 x n_queries x 4
 * (rows >= n_inner unused), last_evals n x n_queries x 4; any of the three may be NULL. */
int32_t stwo_b200_fri_fold_batch(const uint8_t *const *blobs, const size_t *lens, uint32_t n_proofs, const stwo_b200_pcs_config *config,
                                 const uint32_t *input_idx, const uint32_t *input_vals, uint32_t n_inputs, uint32_t *circle_folds,
                                 uint32_t *line_folds, uint32_t *last_evals, uint8_t *verdict, uint8_t *stage);
/* Poseidon31MerkleHasher::hash_column_get_capacity (call site components/hints/src/folding.rs:77; restated at
 * primitives/merkle/src/lib.rs:141-181) for n independent column vectors: cols n x n_cols words, out n x 8 words. */
int32_t stwo_b200_hash_column_capacity_batch(const uint32_t *cols, uint32_t n_cols, size_t n, uint32_t *out);
int32_t stwo_b200_hash_column_capacity_batch_dev(const uint32_t *cols, uint32_t n_cols, size_t n, uint32_t *out, void *stream);

/* ---- multi-GPU: one process per GPU, proofs sharded in contiguous blocks, NCCL only for the two gathers -------
 * The block [lo, hi) of `rank`: the first n % world ranks hold one proof more. */
int32_t stwo_b200_shard_range(uint64_t n, uint32_t rank, uint32_t world, uint64_t *lo, uint64_t *hi);
/* NCCL communicator over the ranks' devices: rank 0 calls comm_unique_id and hands the 128 bytes to the others (any side
 * channel), every rank then calls comm_init after stwo_b200_init.  libnccl.so.2 is loaded on first use. */
int32_t stwo_b200_comm_unique_id(uint8_t id[128]);
int32_t stwo_b200_comm_init(const uint8_t id[128], uint32_t rank, uint32_t world, void **comm);
int32_t stwo_b200_comm_destroy(void *comm);
/* all-gather of the ranks' (verdict, stage) bytes into proof order on every rank; DEVICE pointers; scratch: DEVICE, at least
 * stwo_b200_gather_verdicts_scratch_bytes.  world == 1: a copy, comm may be NULL. */
size_t stwo_b200_gather_verdicts_scratch_bytes(uint32_t world, uint64_t n_total);
int32_t stwo_b200_gather_verdicts(void *comm, uint32_t rank, uint32_t world, uint64_t n_total, const uint8_t *verdict_local,
                                  const uint8_t *stage_local, uint8_t *verdict_all, uint8_t *stage_all, void *scratch, size_t scratch_bytes,
                                  void *stream);
/* gather of per-proof arrays (trace columns: words_per_proof = 13 * n_rows) of every rank's block onto rank dst in proof order;
 * values_all (rank dst only): n_total x words_per_proof words. */
int32_t stwo_b200_gather_trace_columns(void *comm, uint32_t rank, uint32_t world, uint32_t dst, uint64_t n_total, size_t words_per_proof,
                                       const uint32_t *values_local, uint32_t *values_all, void *stream);

/* ---- synthetic FRI + Merkle instances (BASELINE configs[4] part i) ----------------------------------
 * There is no STARK prover on this side of the boundary, so OODS-consistent synthetic proofs cannot exist; what CAN be generated, on
 * the device and all different, is the part of a proof the query phase checks: per instance (seed = seed0 + index) low-degree
 * columns of the shape's three log sizes, committed in one mixed-degree first-layer tree, alpha from a Poseidon channel over the
 * roots, every inner layer folded and committed, the last-layer polynomial, the queries, and the batched decommitments in stwo's
 * layout.  The verifier entry runs the channel replay, the circle / line folds and the FRI tree rebuilds (K3, K5, K2) on them.
 * Instance blob = 256 header words + fixed-capacity sections (recursive-stwo_b200/csrc/synth.cuh); all instances of a batch are
 * stwo_b200_synth_blob_words(shape) words apart.  Shapes: as for proofs, with log_last <= 6 and pow_bits <= 10 (the generator
 * grinds the nonce one thread per instance).  status (optional, DEVICE, n int32): 0 ok, < 0 the generator's own check failed. */
uint32_t stwo_b200_synth_blob_words(const stwo_b200_proof_shape *shape);
size_t stwo_b200_synth_scratch_bytes(const stwo_b200_proof_shape *shape, uint32_t n);
int32_t stwo_b200_synth_generate_dev(const stwo_b200_proof_shape *shape, uint32_t n, uint64_t seed0, uint32_t *blobs, uint64_t *blob_off,
                                     void *scratch, size_t scratch_bytes, int32_t *status, void *stream);
/* workspace: stwo_b200_verify_workspace_bytes(shape, n); flags: STWO_B200_VERIFY_FULL [| _PATH_KERNELS]; results are read with
 * stwo_b200_verify_fetch like those of a proof batch (DETAIL: alphas, queries; CIRCLE_FOLDS, LINE_FOLDS, LAST_EVALS, PATH_ROOTS 4..) */
int32_t stwo_b200_synth_verify_batch_dev(const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                         const stwo_b200_proof_shape *shape, uint32_t flags, void *workspace, size_t workspace_bytes,
                                         uint8_t *verdict, uint8_t *stage, void *stream);

/* ---- K6 / K7: the constraint system on the device ----------------------------------------------------------------------
 * The flat image of PlonkWithPoseidonConstraintSystem (constraint_system/src/plonk_with_poseidon.rs:18-41) after pad().
 * Wiring is shared by every batch item of a shape; values (variables, Poseidon flow hashes, swap bits) are per item.
 * Per-item arrays are LANE-INTERLEAVED: element e of item b lives at ((b / lanes) * n_elems + e) * lanes + b % lanes
 * (lanes = 32: a warp's 32 items touch one contiguous run; lanes = 1: plain [item][element] arrays).
 * All pointers are DEVICE pointers for the *_dev entry points. */
typedef struct {
    uint32_t n_vars, n_rows, n_flow, num_input;      /* n_rows is a power of two >= 16 */
    const uint32_t *a_wire, *b_wire, *c_wire;        /* n_rows */
    const uint32_t *poseidon_wire, *enforce_c_m31;   /* n_rows */
    const uint32_t *op;                              /* n_rows, M31 */
    const uint8_t *op_follows_c;                     /* n_rows or NULL: rows whose op constant is the selected value, i.e. equals
                                                        c_val_0 of that item (CirclePointM31Var::select, primitives/circle/src/lib.rs:80-92) */
    const uint32_t *flow_wire;                       /* n_flow x 4  (PoseidonEntry.wire of entries 1..4) */
    const uint32_t *flow_swap_addr;                  /* n_flow      (SwapOption.addr) */
    /* kind 0: Plonk-with-Poseidon.  kind 1: Plonk-without-Poseidon (constraint_system/src/plonk_without_poseidon.rs:12-25):
     * `op` is op1, op2..op4 select among the arithmetic / hadamard / m4 / pow5m4 / pow5 / grand-sum gates (:410-599),
     * poseidon_wire and enforce_c_m31 are all-zero, n_flow = 0, the logup argument only carries mult_c (:600-632) and the
     * preprocessed block is the 8 columns mult_c, a_wire, b_wire, c_wire, op1, op2, op3, op4 (:633-713). */
    uint32_t kind;
    const uint32_t *op2, *op3, *op4;                 /* n_rows each (kind 1) */
    /* Optional (NULL / 0: none), read by stwo_b200_cs_export_trace_dev with lanes = 32 only: per 64-row tile the distinct variables
     * its wires name, each wire's slot in that list and the row constants, packed by stwo_b200_cs_export_tiles_build
     * (stwo_b200_cs_export_tiles_words(n_rows) words, 16-byte aligned); export_cap = the largest list.  With it the export gathers
     * a variable once per tile instead of once per use, writes 256-byte runs, and needs none of the other columns. */
    const uint32_t *export_tiles;
    uint32_t export_cap;
} stwo_b200_cs_wiring;
/* Host helpers (no device): size of, and the packing of, export_tiles from a wiring whose pointers are HOST pointers
 * (n_rows a multiple of 64; STWO_B200_E_BAD_ARG otherwise, or when a kind-1 selector column holds a value other than 0 / 1). */
size_t stwo_b200_cs_export_tiles_words(uint32_t n_rows);
int32_t stwo_b200_cs_export_tiles_build(const stwo_b200_cs_wiring *host_wiring, uint32_t *tiles, uint32_t *cap);
typedef struct {
    uint32_t n_batch, lanes;                         /* lanes: 1 or 32 */
    uint32_t *variables;                             /* n_vars QM31 per item */
    uint32_t *flow_hash;                             /* n_flow x 32 words per item (PoseidonEntry.hash of entries 1..4); the interleaved
                                                      * ELEMENT is 4 words (16 bytes), like a variable: n_flow x 8 elements per item */
    uint8_t *flow_swap;                              /* n_flow bytes per item (SwapOption.swap) */
    /* Optional (NULL / 0 = none), read by stwo_b200_cs_eval_tape_dev only: output states of permutations that were executed
     * before -- item b's record k is the 16 words at perm_hints[b * perm_hint_stride + 16 k] (item-major, NOT lane-interleaved).
     * A tape permutation whose record names a slot (word 11 of its 12 words = k + 1) takes its output from there instead
     * of permuting. */
    const uint32_t *perm_hints;
    uint32_t perm_hint_stride;
    /* Optional gate on the hints, per item: item b's hints are used only when perm_hint_ready[b] == perm_hint_need (NULL: always).
     * The batched verifier counts there the trees whose part of the record is complete for the run that just finished, so a
     * proof that was rejected half-way, or a workspace last used without STWO_B200_VERIFY_FULL, is evaluated by permuting. */
    const uint32_t *perm_hint_ready;
    uint32_t perm_hint_need;
    /* Optional (NULL = none), read by stwo_b200_cs_check_poseidon_dev only: the INPUT states of the same executed permutations, same
     * layout and stride as perm_hints.  With it (and a tape), a flow entry whose permutation names a record slot is checked against the
     * executed permutation -- entry input == recorded input and entry output == recorded output -- instead of being executed once more:
     * out = permute(in) holds for the record by construction (the native verifier computed it), so the two comparisons are the
     * reference's assertion.  Items whose record is not complete (perm_hint_ready), and entries without a slot, are re-executed. */
    const uint32_t *perm_hint_inputs;
} stwo_b200_cs_values;
/* The value definitions of a recorded circuit, sorted by dependency level (instructions of a level are independent).
 * ins: n_ins x {op, dst, a, b} (STWO_B200_T_*); perms: n_perms x 12 words {l_kind, l_a, l_b, r_kind, r_a, r_b, swap_var,
 * out[4], hint}: half kind 0 = the two QM31 variables (a, b), kind 1 = eight witness-stream words at slot a. */
/* Optional second order of a tape for items whose permutation record is complete (cs_values.perm_hints + perm_hint_ready): every
 * permutation that names a record slot is split into STWO_B200_T_PERM_OUT + STWO_B200_T_PERM_FLOW, so the transcript and the
 * authentication paths stop being dependency chains and the tape is only as deep as its arithmetic.  Same perms / eperms as the tape
 * it belongs to; bundles are mandatory here.  The evaluation picks it per group of `lanes` items, when all of them have a complete record. */
typedef struct {
    uint32_t n_ins, n_levels, n_bundles;
    const uint32_t *ins, *bundle_start /* n_bundles + 1 */, *level_bundle /* n_levels + 1 */;
} stwo_b200_cs_tape_order;
typedef struct {
    uint32_t n_ins, n_perms, n_levels, n_input_words;
    const uint32_t *ins, *level_start /* n_levels + 1 */, *perms;
    /* emulated permutations (STWO_B200_T_EPOSEIDON): n_eperms records of 405 words = the four limb variables, then the 401
     * variables poseidon_permute_emulated creates (primitives/poseidon31/src/emulated.rs:104-221), in creation order */
    uint32_t n_eperms;
    const uint32_t *eperms;
    /* Optional bundles (0 / NULL: every instruction is its own bundle and level_start is what separates independent instructions).
     * A bundle = instructions [bundle_start[k], bundle_start[k + 1]) executed back to back by one warp: each may read what an earlier one
     * of the same bundle wrote.  With bundles, level l is the bundles [level_bundle[l], level_bundle[l + 1]) and only bundles of one
     * level are independent; level_start still gives the level's instruction range. */
    uint32_t n_bundles;
    const uint32_t *bundle_start /* n_bundles + 1 */, *level_bundle /* n_levels + 1 */;
    const stwo_b200_cs_tape_order *recorded_order;   /* HOST pointer to the struct (its arrays are device arrays); NULL: none */
} stwo_b200_cs_tape;
#define STWO_B200_T_ADD 1
#define STWO_B200_T_MUL 2
#define STWO_B200_T_MULC 3
#define STWO_B200_T_INPUT_M31 4
#define STWO_B200_T_INPUT_QM31 5
#define STWO_B200_T_INV_M31 6
#define STWO_B200_T_INV_QM31 7
#define STWO_B200_T_INV_CM31_RE 8
#define STWO_B200_T_INV_CM31_IM 9
#define STWO_B200_T_COORD 10
#define STWO_B200_T_BIT 11
#define STWO_B200_T_POSEIDON 12
#define STWO_B200_T_M4 13
#define STWO_B200_T_POW5M4 14
#define STWO_B200_T_HADAMARD 15
#define STWO_B200_T_GRANDSUM 16
#define STWO_B200_T_POW4 17
#define STWO_B200_T_EPOSEIDON 18
#define STWO_B200_T_PERM_OUT 19    /* recorded_order only: output variables of permutation record dst <- the item's permutation record */
#define STWO_B200_T_PERM_FLOW 20   /* recorded_order only: the flow entry of permutation record dst (reads the halves, defines nothing) */

/* K6: variables[] (and the Poseidon flow) of every batch item from its witness stream (n_input_words words per item,
 * lane-interleaved like the values).  Replaces the `value` arithmetic of every DSL call: primitives/fields/src/{m31,cm31,qm31}.rs,
 * primitives/bits/src/lib.rs:48-82, primitives/poseidon31/src/lib.rs:282-407, plonk_with_poseidon.rs:141-281. */
int32_t stwo_b200_cs_eval_tape_dev(const stwo_b200_cs_tape *t, uint32_t n_vars, const uint32_t *witness, const stwo_b200_cs_values *v,
                                   void *stream);
/* Diagnostics: when level_clock is a device array of n_levels + 1 uint64, the grid-wide form of K6 stores the GPU's
 * globaltimer (ns) after the prologue and after every level; NULL (the default) turns it off. */
int32_t stwo_b200_cs_eval_level_clock(uint64_t *level_clock);
/* check_arithmetics (constraint_system/src/plonk_with_poseidon.rs:337-380): first_bad[b] = index of the first row whose
 * gate c = op(a+b) + (1-op)ab (or whose enforce_c_m31) fails for batch item b, or -1. */
int32_t stwo_b200_cs_check_arithmetics_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, int64_t *first_bad,
                                           void *stream);
/* populate_logup_arguments (:382-466): wiring only.  mult_*: n_rows int32 each; scratch: (4 * n_vars + 4) uint32.
 * status_out[0] != 0 when the reference's assert (a Poseidon wire used more than once) would fire. */
int32_t stwo_b200_cs_populate_logup_dev(const stwo_b200_cs_wiring *w, int32_t *mult_a, int32_t *mult_b, int32_t *mult_c,
                                        int32_t *mult_poseidon, uint32_t *scratch, uint32_t *status_out, void *stream);
/* check_poseidon_invocations (:468-519): first_bad[b] = first flow entry whose wire contents or permutation disagree,
 * or -1.  Needs mult_poseidon and the scratch of populate_logup. */
int32_t stwo_b200_cs_check_poseidon_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v,
                                        const int32_t *mult_poseidon, const uint32_t *scratch, int64_t *first_bad, void *stream);
/* The same with the tape at hand (its permutation records name the slots of v->perm_hints / perm_hint_inputs): entries covered by a
 * complete record are compared with it, the rest re-executed.  tape == NULL or v->perm_hint_inputs == NULL: every entry is re-executed. */
int32_t stwo_b200_cs_check_poseidon_recorded_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const stwo_b200_cs_tape *tape,
                                                 const int32_t *mult_poseidon, const uint32_t *scratch, int64_t *first_bad, void *stream);
/* generate_plonk_with_poseidon_circuit (:521-628).  preprocessed (optional): 10 columns x n_rows in the struct-literal
 * order mult_a, mult_b, mult_c, poseidon_wire, mult_poseidon, enforce_c_m31, a_wire, b_wire, c_wire, op (wiring only;
 * op holds the shape constant).  values: per item 13 columns x n_rows, plain [item][column][row]: a_val_0..3, b_val_0..3,
 * c_val_0..3, then the item's op column (differs from the shared one only on op_follows_c rows).
 * first_bad (optional, needs values): fuses check_arithmetics into the export pass -- the row's three variables are in
 * registers anyway, which saves a full read of variables[]; same result as stwo_b200_cs_check_arithmetics_dev. */
int32_t stwo_b200_cs_export_trace_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const int32_t *mult_a,
                                      const int32_t *mult_b, const int32_t *mult_c, const int32_t *mult_poseidon,
                                      uint32_t *preprocessed, uint32_t *values, int64_t *first_bad, void *stream);
/* PoseidonFlow export for a batch with lanes = 32: the flow the tape evaluation recorded as plain per-item arrays,
 * padded to stwo_b200_cs_flow_padded_len(n_flow) = max(32, ceil(n_flow / 16) * 16) entries like pad() (plonk_with_poseidon.rs:296-321).
 * pad_constants: HOST, 24 words = CONSTANT_1, CONSTANT_2, CONSTANT_3 (plonk_with_poseidon.rs:13-15; their values live in the stwo
 * dependency, not in the reference tree -- the caller supplies them); pad_scratch: DEVICE, 32 words.
 * flow_hash_out: n_batch x n_pad x 32 words (PoseidonEntry.hash of entries 1..4), flow_swap_out: n_batch x n_pad bytes. */
uint32_t stwo_b200_cs_flow_padded_len(uint32_t n_flow);
int32_t stwo_b200_cs_export_flow_dev(const stwo_b200_cs_values *v, uint32_t n_flow, const uint32_t *pad_constants, uint32_t *pad_scratch,
                                     uint32_t *flow_hash_out, uint8_t *flow_swap_out, void *stream);
/* Host entry for ONE constraint system whose values were produced on the host (what the finalisation block of every
 * example does: cs.check_arithmetics(); cs.populate_logup_arguments(); cs.check_poseidon_invocations();
 * cs.generate_plonk_with_poseidon_circuit()  -- examples/single-proof/src/main.rs:85-90).  Pointers in w / v are HOST
 * pointers, v->n_batch and v->lanes must be 1.  trace: 22 x n_rows words (10 preprocessed then 12 value columns, the op
 * column already per-item).  Returns 0 and bad_row = bad_flow = -1 when the system is consistent. */
int32_t stwo_b200_cs_finalize(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, uint32_t *trace,
                              int64_t *bad_row, int64_t *bad_flow);

/* ---- the recursive verifier circuit: recorded once per shape, evaluated per batch ------------------------------------
 * Recording replaces the *row-appending* side of examples/single-proof/src/main.rs:48-84 /
 * examples/multi-proofs/src/main.rs:62-135 (PlonkWithPoseidonProofVar::new_witness, FiatShamirResults, CompositionCheck,
 * AnswerResults, FoldingResults ::compute); it is host logic and needs no device.  multipliers = how many times the
 * proof is verified inside one constraint system (the `multipliers` of multi-proofs). */
typedef struct stwo_b200_circuit stwo_b200_circuit;
typedef struct {
    uint32_t n_rows, n_rows_unpadded, n_vars, n_flow, n_flow_padded, n_input_words, n_ins, n_levels, num_input, words_per_instance;
    uint32_t kind;                                   /* stwo_b200_cs_wiring.kind */
    uint32_t n_preprocessed_columns;                 /* 10 (kind 0) or 8 (kind 1) */
} stwo_b200_circuit_info;
int32_t stwo_b200_circuit_record_verifier(const stwo_b200_proof_shape *shape, const uint32_t *input_idx, const uint32_t *input_vals,
                                          uint32_t n_inputs, uint32_t multipliers, stwo_b200_circuit **out);
/* The last-layer circuit of examples/last-layer/src/main.rs:26-97 (the components/last crates) over the Plonk-without-Poseidon
 * system: every Fiat-Shamir output and opening is a public input (n_public_inputs of them, in the order of main.rs:102-185),
 * the hashes are recomputed by the emulated Poseidon2 gadget (primitives/poseidon31/src/emulated.rs).  The trace pass of such
 * a circuit writes an 8-column preprocessed block and the same 13 per-proof value columns (the last one is op1). */
int32_t stwo_b200_circuit_record_last_layer(const stwo_b200_proof_shape *shape, stwo_b200_circuit **out);
/* The folding stage on its own (FoldingResults::compute, components/recursive/folding/src/lib.rs:12-205) with everything before it
 * supplied as witnesses: FRI commitments, last-layer polynomial, folding alphas, query positions, first-layer answers -- the circuit of
 * BASELINE configs[4] part i ("fri_answers supplied as witnesses").  Its trace pass runs on the workspace of stwo_b200_synth_verify_batch_dev
 * (synthetic FRI + Merkle instances) or of stwo_b200_verify_proofs_batch_dev (real proofs: the folding part of their verifier circuit). */
int32_t stwo_b200_circuit_record_folding(const stwo_b200_proof_shape *shape, stwo_b200_circuit **out);
void stwo_b200_circuit_free(stwo_b200_circuit *c);
int32_t stwo_b200_circuit_get_info(const stwo_b200_circuit *c, stwo_b200_circuit_info *out);
#define STWO_B200_COL_A_WIRE 0
#define STWO_B200_COL_B_WIRE 1
#define STWO_B200_COL_C_WIRE 2
#define STWO_B200_COL_POSEIDON_WIRE 3
#define STWO_B200_COL_ENFORCE_C_M31 4
#define STWO_B200_COL_OP 5
#define STWO_B200_COL_OP_FOLLOWS_C 6
#define STWO_B200_COL_FLOW_WIRE 7
#define STWO_B200_COL_FLOW_SWAP_ADDR 8
#define STWO_B200_COL_LEVEL_START 10
#define STWO_B200_COL_OP2 11
#define STWO_B200_COL_OP3 12
#define STWO_B200_COL_OP4 13
/* host copy of one recorded column (n_words must match its length) */
int32_t stwo_b200_circuit_get_column(const stwo_b200_circuit *c, uint32_t what, uint32_t *out, size_t n_words);

#define STWO_B200_TRACE_CHECK_ARITHMETICS 1u
#define STWO_B200_TRACE_CHECK_POSEIDON 2u
#define STWO_B200_TRACE_TIMED 4u
/* the batch was verified with STWO_B200_VERIFY_FULL on this workspace: the circuit's Poseidon permutations (transcript and
 * per-query authentication paths) are the ones the native verifier just executed and recorded there; K6 takes their output
 * states as hints instead of permuting again, and check_poseidon_invocations compares every such flow entry with the recorded
 * (input, output) pair of the permutation that WAS executed instead of executing it a third time.  A proof whose record is not
 * complete (rejected half-way) is evaluated and checked by permuting, like without the flag. */
#define STWO_B200_TRACE_NATIVE_HINTS 8u
/* with NATIVE_HINTS: check_poseidon_invocations re-executes every flow entry all the same (the round-1 behaviour; the checker of the
 * record-based check in the tests, and a belt-and-braces mode for a caller who wants it) */
#define STWO_B200_TRACE_RECHECK_POSEIDON 16u
/* stage kernels of the trace pass in launch order: gather (+ public-input hashes), eval, check_arithmetics, check_poseidon, export */
#define STWO_B200_N_TRACE_STAGES 5
size_t stwo_b200_circuit_workspace_bytes(const stwo_b200_circuit *c, uint32_t n_proofs);
/* Trace generation for a verified batch: run after stwo_b200_verify_proofs_batch_dev on the same blobs / workspace (the
 * verifier's workspace holds the hints the circuit takes as witnesses).  values: n_proofs x 13 x n_rows words or NULL;
 * preprocessed: 10 x n_rows words or NULL; bad_row / bad_flow: n_proofs each or NULL (need the CHECK flags; -1 = ok). */
int32_t stwo_b200_circuit_trace_batch_dev(stwo_b200_circuit *c, const uint32_t *blobs, const uint64_t *blob_off, uint32_t n_proofs,
                                          const void *verify_workspace, void *circuit_workspace, size_t circuit_workspace_bytes,
                                          uint32_t flags, uint32_t *preprocessed, uint32_t *values, int64_t *bad_row, int64_t *bad_flow,
                                          void *stream);
int32_t stwo_b200_circuit_stage_ms(float *ms /* [STWO_B200_N_TRACE_STAGES] */);
/* PoseidonFlow of the batch stwo_b200_circuit_trace_batch_dev just traced (same circuit workspace): see stwo_b200_cs_export_flow_dev */
int32_t stwo_b200_circuit_export_flow_dev(stwo_b200_circuit *c, uint32_t n_proofs, void *circuit_workspace, size_t circuit_workspace_bytes,
                                          const uint32_t *pad_constants, uint32_t *flow_hash_out, uint8_t *flow_swap_out, void *stream);
#define STWO_B200_CFETCH_VARIABLES 0      /* n_vars x 4 words of proof p */
#define STWO_B200_CFETCH_FLOW_HASH 1      /* n_flow x 32 words */
#define STWO_B200_CFETCH_FLOW_SWAP 2      /* n_flow bytes */
#define STWO_B200_CFETCH_WITNESS 3        /* n_input_words words */
int32_t stwo_b200_circuit_fetch(const stwo_b200_circuit *c, const void *circuit_workspace, uint32_t n_proofs, uint32_t p, uint32_t what,
                                void *out, size_t out_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif
