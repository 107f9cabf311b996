/* TEST HARNESS: a plain C caller of librln_b200 written against include/rln_b200.h, following the flow of the reference's
 * rln/ffi_c_examples/basic_proof.c:7-118 (create, register a member, Merkle proof, witness, prove, verify, serialise).
 * Without a usable GPU it checks that the library says so through the error string and then exercises the host-only record
 * codecs.  Exit code 0 = every check passed; prints "GPU-PATH-OK" or "HOST-PATH-OK". */
#include <stdio.h>
#include <string.h>

#include "rln_b200.h"

#define CHECK(c, msg) do { if (!(c)) { fprintf(stderr, "FAILED: %s\n", msg); return 1; } } while (0)

static int host_only(void) {
    CFr_t *secret = ffi_uint_to_cfr(5), *limit = ffi_uint_to_cfr(10), *mid = ffi_uint_to_cfr(3), *x = ffi_uint_to_cfr(7), *en = ffi_uint_to_cfr(9);
    Vec_CFr_t path = ffi_vec_cfr_new(2);
    CFr_t *e0 = ffi_uint_to_cfr(1), *e1 = ffi_uint_to_cfr(2);
    ffi_vec_cfr_push(&path, e0);
    ffi_vec_cfr_push(&path, e1);
    uint8_t idx_bytes[2] = {0, 1};
    Vec_uint8_t idx = {idx_bytes, 2, 2};
    CResult_FFI_RLNWitnessInput_t w = ffi_rln_witness_input_new_single(secret, limit, mid, &path, &idx, x, en);
    CHECK(w.ok != NULL, "witness_input_new_single");
    CResult_Vec_uint8_t le = ffi_rln_witness_to_bytes_le(&w.ok), be = ffi_rln_witness_to_bytes_be(&w.ok);
    CHECK(le.ok.ptr && be.ok.ptr && le.ok.len == be.ok.len && le.ok.len == 1 + 32 * 7 + 16 + 2, "witness record length");
    CResult_FFI_RLNWitnessInput_t w2 = ffi_bytes_be_to_rln_witness(&be.ok);
    CHECK(w2.ok != NULL, "bytes_be_to_rln_witness");
    CResult_Vec_uint8_t le2 = ffi_rln_witness_to_bytes_le(&w2.ok);
    CHECK(le2.ok.len == le.ok.len && !memcmp(le2.ok.ptr, le.ok.ptr, le.ok.len), "BE -> LE round trip");
    /* a zero limit is refused with the reference's message */
    CFr_t *zero = ffi_cfr_zero();
    CResult_FFI_RLNWitnessInput_t bad = ffi_rln_witness_input_new_single(secret, zero, mid, &path, &idx, x, en);
    CHECK(bad.ok == NULL && bad.err.ptr && strstr((const char *)bad.err.ptr, "User message limit cannot be zero"), "error string");
    ffi_c_string_free(bad.err);
    ffi_vec_u8_free(le.ok); ffi_vec_u8_free(be.ok); ffi_vec_u8_free(le2.ok);
    ffi_rln_witness_input_free(w.ok); ffi_rln_witness_input_free(w2.ok);
    ffi_vec_cfr_free(path);
    ffi_cfr_free(secret); ffi_cfr_free(limit); ffi_cfr_free(mid); ffi_cfr_free(x); ffi_cfr_free(en); ffi_cfr_free(e0); ffi_cfr_free(e1);
    ffi_cfr_free(zero);
    return 0;
}

int main(void) {
    CResult_FFI_RLN_t r = ffi_rln_new(10, "");
    if (!r.ok) {
        CHECK(r.err.ptr && strstr((const char *)r.err.ptr, "no usable CUDA device"), "ffi_rln_new must fail loudly without a GPU");
        ffi_c_string_free(r.err);
        if (host_only()) return 1;
        printf("HOST-PATH-OK\n");
        return 0;
    }
    FFI_RLN_t *rln = r.ok;
    CHECK(ffi_rln_get_tree_depth(&rln) == 10, "tree depth");
    /* identity: secret = H(seed), commitment = Poseidon(secret), rate commitment = Poseidon(commitment, limit) */
    Vec_uint8_t seed = {(uint8_t *)"c-caller-seed", 13, 13};
    Vec_CFr_t keys = ffi_seeded_key_gen(&seed);
    CHECK(keys.len == 2, "seeded keygen");
    CFr_t *limit = ffi_uint_to_cfr(100), *mid = ffi_uint_to_cfr(1);
    CFr_t *rate = ffi_poseidon_hash_pair(&keys.ptr[1], limit);
    CBoolResult_t s = ffi_set_next_leaf(&rln, rate);
    CHECK(s.ok, "set_next_leaf");
    CResult_FFI_MerkleProof_t mp = ffi_get_merkle_proof(&rln, 0);
    CHECK(mp.ok && mp.ok->path_elements.len == 10, "merkle proof");
    Vec_uint8_t sig = {(uint8_t *)"hello", 5, 5}, epoch = {(uint8_t *)"epoch-1", 7, 7};
    CFr_t *x = ffi_hash_to_field_le(&sig), *en = ffi_hash_to_field_le(&epoch);
    CResult_FFI_RLNWitnessInput_t w = ffi_rln_witness_input_new_single(&keys.ptr[0], limit, mid, &mp.ok->path_elements, &mp.ok->path_index, x, en);
    CHECK(w.ok != NULL, "witness");
    CResult_FFI_RLNProof_t p = ffi_generate_rln_proof(&rln, &w.ok);
    CHECK(p.ok != NULL, "generate_rln_proof");
    CBoolResult_t v = ffi_verify_rln_proof(&rln, &p.ok, x);
    CHECK(v.ok, "verify_rln_proof");
    CBoolResult_t v2 = ffi_verify_rln_proof(&rln, &p.ok, en);   /* wrong signal */
    CHECK(!v2.ok && v2.err.ptr && strstr((const char *)v2.err.ptr, "Signal value does not match"), "wrong signal is rejected");
    ffi_c_string_free(v2.err);
    CResult_Vec_uint8_t bytes = ffi_rln_proof_to_bytes_le(&p.ok);
    CHECK(bytes.ok.len == 290, "proof record is 290 bytes");
    CResult_FFI_RLNProof_t p2 = ffi_bytes_le_to_rln_proof(&bytes.ok);
    CHECK(p2.ok != NULL, "bytes_le_to_rln_proof");
    CBoolResult_t v3 = ffi_verify_rln_proof(&rln, &p2.ok, x);
    CHECK(v3.ok, "verify after a byte round trip");
    ffi_vec_u8_free(bytes.ok);
    ffi_rln_proof_free(p.ok); ffi_rln_proof_free(p2.ok);
    ffi_rln_witness_input_free(w.ok);
    ffi_merkle_proof_free(mp.ok);
    ffi_cfr_free(limit); ffi_cfr_free(mid); ffi_cfr_free(rate); ffi_cfr_free(x); ffi_cfr_free(en);
    ffi_vec_cfr_free(keys);
    ffi_rln_free(rln);
    if (host_only()) return 1;
    printf("GPU-PATH-OK\n");
    return 0;
}
