/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_field.h header).
 * Public surface of the CPU restatement.  Each function cites the reference
 * file:line it follows.  Built into oracle/liborc.so by oracle/Makefile.
 *
 * Parity status: PINNED by (i) the Poseidon2 KAT (reference
 * primitives/poseidon31/src/implementation.rs:157-173), (ii) the known answers
 * embedded in the reference's 15 Poseidon31 proof fixtures (PoW low bits,
 * Merkle roots, logup sum, OODS equality, FRI last-layer equality) and
 * (iii) the survey-probe golden values in SURVEY.md App. F.  The Rust
 * reference itself cannot be built here (no cargo/rustc; stwo git dependency
 * absent), so there is no oracle/_ref.
 */
#ifndef ORC_H
#define ORC_H
#include <stddef.h>
#include <stdint.h>
#include "orc_field.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Poseidon2 / Merkle / channel --------------------------------------- */
void orc_poseidon2_permute(uint32_t state[16]);
void orc_poseidon2_permute_batch(uint32_t *states, size_t n);
/* stwo Poseidon31MerkleHasher::hash_node; children may be NULL (leaf) */
void orc_hash_node(const uint32_t *left, const uint32_t *right,
                   const uint32_t *cols, size_t n_cols, uint32_t out[8]);
void orc_hash_column_get_capacity(const uint32_t *cols, size_t n_cols, uint32_t out[8]);
/* full Merkle tree over 2^log_n leaves of n_cols M31 each (row-major leaves);
 * layers[0] = leaf hashes ... ; out_nodes holds (2^(log_n+1)-1)*8 words, layer
 * of size 2^k starting at word offset (2^k - 1)*8. Returns perms executed. */
uint64_t orc_merkle_build(const uint32_t *leaves, uint32_t log_n, uint32_t n_cols,
                          uint32_t *out_nodes);
/* verify one leaf->root path (single-tree, columns only at the leaf layer) */
int orc_merkle_path_verify(const uint32_t *leaf, uint32_t n_cols, uint32_t index,
                           const uint32_t *siblings, uint32_t depth,
                           const uint32_t root[8], uint32_t out_root[8]);

void orc_merkle_path_root_mixed(uint32_t depth, const uint32_t *n_cols, uint32_t index,
                                const uint32_t *cols, const uint32_t *siblings, uint32_t out[8]);

/* pthread drivers for the CPU baseline */
void orc_poseidon2_permute_batch_mt(uint32_t *states, size_t n, unsigned n_threads);
uint64_t orc_merkle_build_mt(const uint32_t *leaves, uint32_t log_n, uint32_t n_cols, uint32_t *nodes, unsigned n_threads);
void orc_merkle_paths_verify_mt(uint32_t depth, const uint32_t *n_cols, size_t n_paths, const uint32_t *index,
                                const uint32_t *cols, const uint32_t *sib, const uint32_t *root, uint8_t *verdict,
                                unsigned n_threads);

typedef struct { uint32_t digest[8]; uint32_t n_sent; uint64_t n_perms; } orc_channel;
void orc_channel_init(orc_channel *c);
void orc_channel_mix_root(orc_channel *c, const uint32_t root[8]);
void orc_channel_mix_felts2(orc_channel *c, const uint32_t a[4], const uint32_t b[4]);
void orc_channel_draw(orc_channel *c, uint32_t out[8]);

/* ---- proof wire format ----------------------------------- */
#define ORC_MAX_INNER 32
#define ORC_MAX_COLS 64
typedef struct {
    const uint32_t *hash_witness; uint64_t n_hash_witness;   /* x8 words */
    uint64_t n_column_witness;
} orc_decommitment;
typedef struct {
    const uint32_t *fri_witness; uint64_t n_fri_witness;     /* x4 words */
    orc_decommitment decommitment;
    const uint32_t *commitment;                               /* 8 words */
} orc_fri_layer;
typedef struct {
    uint32_t log_size_plonk, log_size_poseidon;
    qm31 plonk_total_sum, poseidon_total_sum;
    uint32_t pow_bits, log_blowup, log_last, n_queries;
    const uint32_t *commitments[4];
    uint32_t n_cols[4];
    uint32_t n_masks[4][ORC_MAX_COLS];
    const uint32_t *sampled[4][ORC_MAX_COLS];                 /* n_masks x4 words */
    uint32_t n_sampled_total;
    orc_decommitment decommitments[4];
    const uint32_t *queried_values[4]; uint64_t n_queried_values[4];
    uint64_t pow_nonce;
    orc_fri_layer first_layer;
    uint32_t n_inner;
    orc_fri_layer inner[ORC_MAX_INNER];
    const uint32_t *last_coeffs; uint64_t n_last_coeffs; uint32_t last_log_size;
} orc_proof;
/* returns 0 on success, negative on malformed input */
int orc_proof_parse(const uint8_t *blob, size_t len, orc_proof *out);

int orc_proof_offsets(const uint8_t *blob, size_t len, uint64_t out[16]);

/* ---- full native verifier ------------------------------------------------- */
#define ORC_MAX_QUERIES 128
#define ORC_MAX_LOGS 4
enum {
    ORC_OK = 0,
    ORC_STAGE_PARSE = 1, ORC_STAGE_POW = 2, ORC_STAGE_LOGUP = 3, ORC_STAGE_OODS = 4,
    ORC_STAGE_MERKLE = 5, ORC_STAGE_FRI_FIRST = 6, ORC_STAGE_FRI_INNER = 7,
    ORC_STAGE_FRI_LAST = 8, ORC_STAGE_UNSUPPORTED = 9
};
typedef struct {
    /* Fiat-Shamir (reference components/recursive/fiat_shamir/src/lib.rs:31-131) */
    qm31 z, alpha, random_coeff, oods_t, oods_x, oods_y, after_coeff;
    qm31 fri_alphas[ORC_MAX_INNER + 1];
    uint32_t digest_after_nonce[8];
    uint32_t raw_queries[ORC_MAX_QUERIES];
    uint32_t n_transcript_perms;
    /* shape */
    uint32_t max_first_log, n_inner, n_queries, n_logs;
    uint32_t log_sizes[ORC_MAX_LOGS];                 /* descending */
    uint32_t query_pos[ORC_MAX_LOGS][ORC_MAX_QUERIES];
    /* OODS */
    qm31 oods_computed, oods_expected;
    /* answers / folds, [log idx][query] */
    cpoint domain_points[ORC_MAX_LOGS][ORC_MAX_QUERIES];
    qm31 fri_answers[ORC_MAX_LOGS][ORC_MAX_QUERIES];
    qm31 circle_folds[ORC_MAX_LOGS][ORC_MAX_QUERIES];
    qm31 line_folds[ORC_MAX_INNER][ORC_MAX_QUERIES];  /* value after inner layer i */
    qm31 last_layer_evals[ORC_MAX_QUERIES];
    /* Merkle: recomputed roots per (tree, query); trees 0-3, 4 = FRI first, 5.. inner */
    uint32_t path_roots[5 + ORC_MAX_INNER][ORC_MAX_QUERIES][8];
    uint64_t n_perms_hints;     /* partial-tree rebuild */
    uint64_t n_perms_paths;     /* per-query path verification (incl. transcript) */
    int32_t verdict;            /* 0 accept, 1 reject, 2 unsupported */
    int32_t stage;              /* first failing ORC_STAGE_* */
} orc_verify_out;

/* inputs: (idx, value) public-input pairs for the logup sum
 * (reference components/recursive/fiat_shamir/src/lib.rs:133-141) */
int orc_verify_proof(const uint8_t *blob, size_t len,
                     const uint32_t *input_idx, const uint32_t *input_vals /* n x4 */,
                     uint32_t n_inputs, orc_verify_out *out);
/* FRI-only verifier of a synthetic FRI + Merkle instance (blob layout: recursive-stwo_b200/csrc/synth.cuh) */
int orc_fri_verify_synth(const uint32_t *words, size_t n_words, orc_verify_out *out);
/* the same under the caller's PcsConfig cfg = {pow_bits, log_blowup, log_last, n_queries}: a proof claiming another config is
 * rejected at the parse stage */
int orc_verify_proof_cfg(const uint8_t *blob, size_t len, const uint32_t *cfg, const uint32_t *input_idx, const uint32_t *input_vals,
                         uint32_t n_inputs, orc_verify_out *o);

/* per-query decommitment hints = the witnesses DecommitmentVar / SinglePairMerkleProofVar allocate
 * (components/recursive/data_structures/src/lib.rs:287-312,372-398) */
typedef struct {
    uint32_t single_depth[4];
    uint32_t single_ncols[4][33];                               /* columns per layer log size */
    uint32_t single_cols[4][ORC_MAX_QUERIES][64];               /* leaf layer first, then descending layers */
    uint32_t single_sib[4][ORC_MAX_QUERIES][32][8];             /* sibling_hashes[i], i = depth - h */
    uint32_t pair_depth[1 + ORC_MAX_INNER];                     /* 0 = FRI first layer, 1.. = inner layers */
    uint8_t pair_has_data[1 + ORC_MAX_INNER][33];
    uint32_t pair_self[1 + ORC_MAX_INNER][ORC_MAX_QUERIES][33][4];
    uint32_t pair_sib[1 + ORC_MAX_INNER][ORC_MAX_QUERIES][33][4];
    uint32_t pair_sib_hash[1 + ORC_MAX_INNER][ORC_MAX_QUERIES][32][8];
} orc_hints;
int orc_verify_proof_hints(const uint8_t *blob, size_t len, const uint32_t *input_idx, const uint32_t *input_vals,
                           uint32_t n_inputs, orc_verify_out *out, orc_hints *hints);

int orc_fri_verify_synth_hints(const uint32_t *words, size_t n_words, orc_verify_out *out, orc_hints *hints);

/* independent proofs on n_threads pthreads (CPU baseline); off = n+1 byte offsets; returns permutations executed */
uint64_t orc_verify_batch_mt(const uint8_t *blobs, const uint64_t *off, uint32_t n, const uint32_t *idx, const uint32_t *vals,
                             uint32_t n_inputs, uint8_t *verdict, uint8_t *stage, unsigned n_threads);

/* ---- circuit value log replay (oracle/orc_tape.c; the log is recorded by oracle/orc_dsl.py) ---------------------------- */
void orc_circuit_replay(const uint32_t *ops, uint32_t n_ops, const uint32_t *perms, uint32_t *vars, uint32_t *flow_hash, uint8_t *flow_swap);
int64_t orc_circuit_check_arithmetics(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars);
int64_t orc_circuit_check_arithmetics_without(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars);
int64_t orc_circuit_check_poseidon(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars, uint32_t n_vars, const uint32_t *flow_wire,
                                   uint32_t n_flow, const uint32_t *flow_hash, const uint8_t *flow_swap, uint32_t *row_of_wire);
void orc_circuit_export_values(const uint32_t *wiring, uint32_t n_rows, const uint32_t *vars, uint32_t *out);
int64_t orc_circuit_trace_mt(const uint32_t *ops, uint32_t n_ops, const uint32_t *perms, uint32_t n_perms, uint32_t n_vars,
                             const uint32_t *wiring, uint32_t n_rows, const uint32_t *flow_wire, uint32_t n, unsigned n_threads);

#ifdef __cplusplus
}
#endif
#endif
