/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc.h).
 * bincode 1.3 reader for PlonkWithPoseidonProof<Poseidon31MerkleHasher>
 * (struct usage at reference
 * components/recursive/data_structures/src/lib.rs:98-223 and
 * components/hints/src/fiat_shamir.rs:69-216). */
#include <string.h>
#include "orc.h"

typedef struct { const uint8_t *p; size_t len, off; int err; } rd;

static uint32_t r_u32(rd *r) {
    if (r->err || r->off + 4 > r->len) { r->err = 1; return 0; }
    uint32_t v; memcpy(&v, r->p + r->off, 4); r->off += 4; return v;
}
static uint64_t r_u64(rd *r) {
    if (r->err || r->off + 8 > r->len) { r->err = 1; return 0; }
    uint64_t v; memcpy(&v, r->p + r->off, 8); r->off += 8; return v;
}
/* n_words M31 words, each must be canonical (< p) */
static const uint32_t *r_words(rd *r, uint64_t n_words) {
    if (r->err || n_words > (r->len - r->off) / 4) { r->err = 1; return NULL; }
    const uint32_t *w = (const uint32_t *)(r->p + r->off);
    for (uint64_t i = 0; i < n_words; i++) if (w[i] >= ORC_P) { r->err = 2; return NULL; }
    r->off += 4 * n_words;
    return w;
}
static void r_decommitment(rd *r, orc_decommitment *d) {
    d->n_hash_witness = r_u64(r);
    if (d->n_hash_witness > r->len) { r->err = 1; return; }
    d->hash_witness = r_words(r, d->n_hash_witness * 8);
    d->n_column_witness = r_u64(r);
    if (d->n_column_witness > r->len) { r->err = 1; return; }
    r_words(r, d->n_column_witness);
}
static void r_layer(rd *r, orc_fri_layer *l) {
    l->n_fri_witness = r_u64(r);
    if (l->n_fri_witness > r->len) { r->err = 1; return; }
    l->fri_witness = r_words(r, l->n_fri_witness * 4);
    r_decommitment(r, &l->decommitment);
    l->commitment = r_words(r, 8);
}

int orc_proof_parse(const uint8_t *blob, size_t len, orc_proof *o) {
    rd r = { blob, len, 0, 0 };
    memset(o, 0, sizeof *o);
    if (((uintptr_t)blob & 3) != 0) return -3;
    o->log_size_plonk = r_u32(&r);
    o->log_size_poseidon = r_u32(&r);
    const uint32_t *w = r_words(&r, 8);
    if (!w) return -1;
    memcpy(&o->plonk_total_sum, w, 16); memcpy(&o->poseidon_total_sum, w + 4, 16);
    o->pow_bits = r_u32(&r);
    o->log_blowup = r_u32(&r);
    o->log_last = r_u32(&r);
    uint64_t nq = r_u64(&r);
    if (r.err || nq == 0 || nq > ORC_MAX_QUERIES) return -1;
    o->n_queries = (uint32_t)nq;
    if (r_u64(&r) != 4) return -1;
    for (int t = 0; t < 4; t++) o->commitments[t] = r_words(&r, 8);
    if (r_u64(&r) != 4) return -1;
    for (int t = 0; t < 4 && !r.err; t++) {
        uint64_t nc = r_u64(&r);
        if (r.err || nc > ORC_MAX_COLS) return -1;
        o->n_cols[t] = (uint32_t)nc;
        for (uint64_t c = 0; c < nc && !r.err; c++) {
            uint64_t nm = r_u64(&r);
            if (r.err || nm > 2) return -1;
            o->n_masks[t][c] = (uint32_t)nm;
            o->sampled[t][c] = r_words(&r, nm * 4);
            o->n_sampled_total += (uint32_t)nm;
        }
    }
    if (r_u64(&r) != 4) return -1;
    for (int t = 0; t < 4 && !r.err; t++) r_decommitment(&r, &o->decommitments[t]);
    if (r_u64(&r) != 4) return -1;
    for (int t = 0; t < 4 && !r.err; t++) {
        o->n_queried_values[t] = r_u64(&r);
        if (o->n_queried_values[t] > len) return -1;
        o->queried_values[t] = r_words(&r, o->n_queried_values[t]);
    }
    o->pow_nonce = r_u64(&r);
    r_layer(&r, &o->first_layer);
    uint64_t ni = r_u64(&r);
    if (r.err || ni > ORC_MAX_INNER) return -1;
    o->n_inner = (uint32_t)ni;
    for (uint32_t i = 0; i < o->n_inner && !r.err; i++) r_layer(&r, &o->inner[i]);
    o->n_last_coeffs = r_u64(&r);
    if (r.err || o->n_last_coeffs > len) return -1;
    o->last_coeffs = r_words(&r, o->n_last_coeffs * 4);
    o->last_log_size = r_u32(&r);
    if (r.err) return -r.err;
    if (r.off != len) return -1;
    return 0;
}

/* byte offsets of the regions a test wants to tamper with; returns the number written (0 on parse error) */
int orc_proof_offsets(const uint8_t *blob, size_t len, uint64_t out[16]) {
    static _Thread_local orc_proof p;
    if (orc_proof_parse(blob, len, &p) != 0) return 0;
#define OFF(ptr) ((ptr) ? (uint64_t)((const uint8_t *)(ptr) - blob) : (uint64_t)-1)
    out[0] = OFF(p.commitments[0]);
    out[1] = OFF(p.sampled[0][0]);
    out[2] = OFF(p.decommitments[0].hash_witness);
    out[3] = OFF(p.queried_values[0]);
    out[4] = OFF(p.first_layer.fri_witness);
    out[5] = OFF(p.first_layer.decommitment.hash_witness);
    out[6] = OFF(p.first_layer.commitment);
    out[7] = p.n_inner ? OFF(p.inner[0].fri_witness) : (uint64_t)-1;
    out[8] = p.n_inner ? OFF(p.inner[0].decommitment.hash_witness) : (uint64_t)-1;
    out[9] = OFF(p.last_coeffs);
    out[10] = OFF(p.queried_values[3]);
    out[11] = p.n_inner ? OFF(p.inner[p.n_inner - 1].fri_witness) : (uint64_t)-1;
    out[12] = OFF(p.first_layer.commitment) - 8 - 8;   /* not meaningful: placeholder kept stable */
    out[13] = OFF(p.sampled[3][7]);
    out[14] = OFF(p.commitments[3]);
    out[15] = OFF(p.queried_values[0]) - 8 - 8;
    /* pow nonce sits right after the last queried_values vector */
    out[12] = OFF(p.queried_values[3]) + 4 * p.n_queried_values[3];
#undef OFF
    return 16;
}
