/* TEST INFRASTRUCTURE: a plain C program that links libstwo_b200.so through include/stwo_b200.h only -- no Python, no ctypes -- the
 * way a cgo / Rust -sys binding would.  usage: harness <small_proof.bin>
 * Without a CUDA device it checks the host-only entry points and that every compute call fails loudly (no CPU fallback); with one
 * it runs the Poseidon2 known answer of the reference's own test (primitives/poseidon31/src/implementation.rs:157-173) and
 * verifies the reference's fixture under the caller's PcsConfig (examples/single-proof/src/main.rs:24-47).  Prints one line. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/stwo_b200.h"

#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char **argv) {
    CHECK(argc == 2);
    FILE *f = fopen(argv[1], "rb");
    CHECK(f);
    static uint32_t buf[1 << 16];
    size_t len = fread(buf, 1, sizeof buf, f);
    fclose(f);
    CHECK(len == 59784);

    /* host-only entry points */
    CHECK(stwo_b200_version() == 0x00000100u);
    stwo_b200_proof_shape claimed, shape;
    CHECK(stwo_b200_proof_shape_of((const uint8_t *)buf, len, &claimed) == STWO_B200_OK);
    CHECK(claimed.log_size_plonk == 4 && claimed.log_size_poseidon == 8 && claimed.n_queries == 16 && claimed.n_inner == 7);
    const stwo_b200_pcs_config single = {20, 5, 2, 16};          /* examples/single-proof/src/main.rs:28-31 */
    CHECK(stwo_b200_shape_from_config(&single, claimed.log_size_plonk, claimed.log_size_poseidon, &shape) == STWO_B200_OK);
    CHECK(memcmp(&shape, &claimed, sizeof shape) == 0);
    CHECK(stwo_b200_verify_workspace_bytes(&shape, 4) > 0 && stwo_b200_proof_perms(&shape) == 3481 - 105);
    uint64_t lo, hi;
    CHECK(stwo_b200_shard_range(10, 3, 4, &lo, &hi) == STWO_B200_OK && lo == 8 && hi == 10);
    CHECK(stwo_b200_cs_flow_padded_len(3481) == 3488 && stwo_b200_cs_flow_padded_len(5) == 32);

    uint32_t st[16];
    for (int i = 0; i < 16; i++) st[i] = (uint32_t)i;
    const uint8_t *blobs[2] = {(const uint8_t *)buf, (const uint8_t *)buf};
    const size_t lens[2] = {len, len};
    const uint32_t idx[1] = {1}, vals[4] = {1, 0, 0, 0};          /* the public input (1, 1) */
    uint8_t verdict[2] = {9, 9}, stage[2] = {9, 9};

    int32_t rc = stwo_b200_init(0);
    if (rc == STWO_B200_E_NO_DEVICE) {
        CHECK(stwo_b200_poseidon2_permute(st, 1) == STWO_B200_E_NO_DEVICE);
        CHECK(stwo_b200_verify_proofs_batch(blobs, lens, 2, &single, 1, idx, vals, 1, STWO_B200_VERIFY_FULL, verdict, stage) == STWO_B200_E_NO_DEVICE);
        printf("OK no-device: host entry points answer, compute calls return STWO_B200_E_NO_DEVICE\n");
        return 0;
    }
    CHECK(rc == STWO_B200_OK);
    CHECK(stwo_b200_poseidon2_permute(st, 1) == STWO_B200_OK);
    CHECK(st[0] == 260776483u);                                   /* implementation.rs:166 */
    CHECK(stwo_b200_verify_proofs_batch(blobs, lens, 2, &single, 1, idx, vals, 1, STWO_B200_VERIFY_FULL, verdict, stage) == STWO_B200_OK);
    CHECK(verdict[0] == STWO_B200_VERDICT_ACCEPT && verdict[1] == STWO_B200_VERDICT_ACCEPT && stage[0] == STWO_B200_STAGE_OK);
    /* the wrong config is a rejection at the parse stage, not an error */
    const stwo_b200_pcs_config standard = {20, 5, 8, 16};
    CHECK(stwo_b200_verify_proofs_batch(blobs, lens, 2, &standard, 1, idx, vals, 1, 0, verdict, stage) == STWO_B200_OK);
    CHECK(verdict[0] == STWO_B200_VERDICT_REJECT && stage[0] == STWO_B200_STAGE_PARSE);
    /* one stage only: the transcript replay */
    static stwo_b200_verify_detail det[2];
    CHECK(stwo_b200_channel_replay_batch(blobs, lens, 2, &single, idx, vals, 1, det, verdict, stage) == STWO_B200_OK);
    CHECK(verdict[1] == 0 && det[1].fs.n_transcript_perms == 105 && det[1].fs.z[0] == 1211683141u && det[1].fs.pow_ok == 1);   /* SURVEY App. F */
    CHECK((det[0].fs.raw_queries[0] & 0x7fff) == 3311);
    CHECK(stwo_b200_shutdown() == STWO_B200_OK);
    printf("OK device: Poseidon2 KAT, fixture accepted under the caller's config, rejected under another, transcript replay matches App. F\n");
    return 0;
}
