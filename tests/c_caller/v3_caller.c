/* TEST HARNESS: a plain C caller of the V3 half of the ABI (ffi_rln_v3_*), written against include/rln_b200.h with the type names
 * the safer-ffi generated rln.h uses (CResult_…_ptr_Vec_uint8_t) — the API surface the reference's rln/ffi_c_examples exercise:
 * stateful build from zkey / graph bytes, membership, proof, verify, verify_with_roots, LE record round trip, two-phase proving,
 * Shamir recovery from two proofs of one epoch, and a stateless build.
 * usage: v3_caller <dir with rln_final.arkzkey and graph.bin> <tree depth>
 * Without a usable GPU it must be told so through the error string.  Exit code 0 = every check passed; prints "V3-GPU-PATH-OK"
 * or "V3-HOST-PATH-OK". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rln_b200.h"

#define CHECK(c, msg) do { if (!(c)) { fprintf(stderr, "FAILED (line %d): %s\n", __LINE__, msg); return 1; } } while (0)

static int slurp(const char *dir, const char *name, Vec_uint8_t *out) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->ptr = (uint8_t *)malloc((size_t)n);
    out->len = out->cap = (size_t)n;
    size_t got = fread(out->ptr, 1, (size_t)n, f);
    fclose(f);
    return got == (size_t)n ? 0 : -1;
}

static int same_fr(const CFr_t *a, const CFr_t *b) {
    CResult_Vec_uint8_Vec_uint8_t x = ffi_cfr_to_bytes_le(a), y = ffi_cfr_to_bytes_le(b);
    int eq = x.ok.ptr && y.ok.ptr && x.ok.len == 32 && y.ok.len == 32 && !memcmp(x.ok.ptr, y.ok.ptr, 32);
    ffi_vec_u8_free(x.ok);
    ffi_vec_u8_free(y.ok);
    return eq;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: v3_caller <resource dir> <tree depth>\n"); return 2; }
    const size_t depth = (size_t)atoi(argv[2]);
    Vec_uint8_t zkey, graph;
    CHECK(slurp(argv[1], "rln_final.arkzkey", &zkey) == 0 && slurp(argv[1], "graph.bin", &graph) == 0, "read circuit files");

    CResult_FFI_RLNV3_ptr_Vec_uint8_t made = ffi_rln_v3_new_with_pm_tree(depth, &zkey, &graph, "");
    if (!made.ok) {
        CHECK(made.err.ptr && strstr((const char *)made.err.ptr, "no usable CUDA device"), "construction must fail loudly without a GPU");
        ffi_c_string_free(made.err);
        printf("V3-HOST-PATH-OK\n");
        return 0;
    }
    FFI_RLNV3_t *rln = made.ok;

    /* member: (secret, commitment) from the key generator, leaf = Poseidon(commitment, limit) */
    Vec_CFr_t keys = ffi_key_gen();
    CHECK(keys.len == 2, "key_gen returns (secret, commitment)");
    const CFr_t *secret = ffi_vec_cfr_get(&keys, 0), *commitment = ffi_vec_cfr_get(&keys, 1);
    CFr_t *limit = ffi_uint_to_cfr(10), *mid = ffi_uint_to_cfr(2);
    CFr_t *leaf = ffi_poseidon_hash_pair(commitment, limit);
    CBoolResult_t added = ffi_rln_v3_set_next_leaf(&rln, leaf);
    CHECK(added.ok, "set_next_leaf");
    CHECK(ffi_rln_v3_leaves_set(&rln) == 1, "leaves_set");
    CResult_FFI_RLNV3MerkleProof_ptr_Vec_uint8_t mp = ffi_rln_v3_get_merkle_proof(&rln, 0);
    CHECK(mp.ok && mp.ok->path_elements.len == depth && mp.ok->path_index.len == depth, "merkle proof shape");

    Vec_uint8_t sig1 = {(uint8_t *)"first message", 13, 13}, sig2 = {(uint8_t *)"second message", 14, 14}, ep = {(uint8_t *)"epoch-7", 7, 7};
    CFr_t *x1 = ffi_hash_to_field_le(&sig1), *x2 = ffi_hash_to_field_le(&sig2), *en = ffi_hash_to_field_le(&ep);
    CResult_FFI_RLNV3WitnessInput_ptr_Vec_uint8_t w1 =
        ffi_rln_v3_witness_input_new_single(secret, limit, mid, &mp.ok->path_elements, &mp.ok->path_index, x1, en);
    CResult_FFI_RLNV3WitnessInput_ptr_Vec_uint8_t w2 =
        ffi_rln_v3_witness_input_new_single(secret, limit, mid, &mp.ok->path_elements, &mp.ok->path_index, x2, en);
    CHECK(w1.ok && w2.ok, "witness_input_new_single");
    /* message_id must stay below the limit: the reference's validation text */
    CFr_t *big = ffi_uint_to_cfr(10);
    CResult_FFI_RLNV3WitnessInput_ptr_Vec_uint8_t wbad =
        ffi_rln_v3_witness_input_new_single(secret, limit, big, &mp.ok->path_elements, &mp.ok->path_index, x1, en);
    CHECK(!wbad.ok && wbad.err.ptr && strstr((const char *)wbad.err.ptr, "is not within user_message_limit"), "message id range check");
    ffi_c_string_free(wbad.err);

    /* prove, verify, wrong signal, roots */
    CResult_FFI_RLNV3Proof_ptr_Vec_uint8_t p1 = ffi_rln_v3_generate_proof(&rln, &w1.ok);
    CHECK(p1.ok != NULL, "generate_proof");
    CBoolResult_t ok1 = ffi_rln_v3_verify(&rln, &p1.ok, x1);
    CHECK(ok1.ok && !ok1.err.ptr, "verify");
    CBoolResult_t wrong = ffi_rln_v3_verify(&rln, &p1.ok, x2);
    CHECK(!wrong.ok && !wrong.err.ptr, "verify with another signal is Ok(false)");
    CFr_t *root = ffi_rln_v3_get_root(&rln);
    FFI_RLNV3ProofValues_t *pv1 = ffi_rln_v3_proof_get_values(&p1.ok);
    CFr_t *pv_root = ffi_rln_v3_proof_values_get_root(&pv1);
    CHECK(same_fr(root, pv_root), "proof root == tree root");
    Vec_CFr_t roots = ffi_vec_cfr_new(2);
    ffi_vec_cfr_push(&roots, en);      /* some other value */
    ffi_vec_cfr_push(&roots, root);
    CBoolResult_t in_roots = ffi_rln_v3_verify_with_roots(&rln, &p1.ok, &roots, x1);
    CHECK(in_roots.ok, "verify_with_roots accepts a listed root");
    Vec_CFr_t other = ffi_vec_cfr_new(1);
    ffi_vec_cfr_push(&other, en);
    CBoolResult_t not_in = ffi_rln_v3_verify_with_roots(&rln, &p1.ok, &other, x1);
    CHECK(!not_in.ok && not_in.err.ptr && strstr((const char *)not_in.err.ptr, "Expected one of the provided roots"), "unknown root is an error");
    ffi_c_string_free(not_in.err);

    /* LE record round trip */
    CResult_Vec_uint8_Vec_uint8_t rec = ffi_rln_v3_proof_to_bytes_le(&p1.ok);
    CHECK(rec.ok.ptr && rec.ok.len == 128 + 1 + 5 * 32, "V3 single proof record is 289 bytes");
    CResult_FFI_RLNV3Proof_ptr_Vec_uint8_t back = ffi_bytes_le_to_rln_v3_proof(&rec.ok);
    CHECK(back.ok != NULL, "bytes_le_to_rln_v3_proof");
    CBoolResult_t ok_back = ffi_rln_v3_verify(&rln, &back.ok, x1);
    CHECK(ok_back.ok, "verify after a byte round trip");

    /* two-phase proving: the partial proof depends on the member only, the finish on the message */
    FFI_RLNV3PartialWitnessInput_t *pw = ffi_rln_v3_witness_to_partial_witness(&w2.ok);
    CHECK(pw != NULL, "witness_to_partial_witness");
    CResult_FFI_RLNV3PartialProof_ptr_Vec_uint8_t part = ffi_rln_v3_generate_partial_proof(&rln, &pw);
    CHECK(part.ok != NULL, "generate_partial_proof");
    CResult_FFI_RLNV3Proof_ptr_Vec_uint8_t p2 = ffi_rln_v3_finish_proof(&rln, &part.ok, &w2.ok);
    CHECK(p2.ok != NULL, "finish_proof");
    CBoolResult_t ok2 = ffi_rln_v3_verify(&rln, &p2.ok, x2);
    CHECK(ok2.ok, "verify the two-phase proof");

    /* two messages under one (epoch, message id) leak the secret: Shamir recovery */
    FFI_RLNV3ProofValues_t *pv2 = ffi_rln_v3_proof_get_values(&p2.ok);
    CResult_CFr_ptr_Vec_uint8_t rec_secret = ffi_rln_v3_recover_id_secret(&pv1, &pv2);
    CHECK(rec_secret.ok && same_fr(rec_secret.ok, secret), "recover_id_secret returns the identity secret");

    /* a stateless object proves and verifies the same statement and refuses tree operations */
    CResult_FFI_RLNV3_ptr_Vec_uint8_t sl = ffi_rln_v3_new_stateless(&zkey, &graph);
    CHECK(sl.ok != NULL, "new_stateless");
    CBoolResult_t ok_sl = ffi_rln_v3_verify_with_roots(&sl.ok, &p1.ok, &roots, x1);
    CHECK(ok_sl.ok, "stateless verify_with_roots");
    CResult_FFI_RLNV3MerkleProof_ptr_Vec_uint8_t no_tree = ffi_rln_v3_get_merkle_proof(&sl.ok, 0);
    CHECK(!no_tree.ok && no_tree.err.ptr, "stateless object has no tree");
    ffi_c_string_free(no_tree.err);

    ffi_rln_v3_free(sl.ok);
    ffi_cfr_free(rec_secret.ok);
    ffi_rln_v3_proof_values_free(pv1); ffi_rln_v3_proof_values_free(pv2);
    ffi_rln_v3_proof_free(p1.ok); ffi_rln_v3_proof_free(p2.ok); ffi_rln_v3_proof_free(back.ok);
    ffi_rln_v3_partial_proof_free(part.ok);
    ffi_rln_v3_partial_witness_input_free(pw);
    ffi_rln_v3_witness_input_free(w1.ok); ffi_rln_v3_witness_input_free(w2.ok);
    ffi_vec_u8_free(rec.ok);
    ffi_vec_cfr_free(roots); ffi_vec_cfr_free(other);
    ffi_cfr_free(root); ffi_cfr_free(pv_root);
    ffi_cfr_free(x1); ffi_cfr_free(x2); ffi_cfr_free(en); ffi_cfr_free(big);
    ffi_rln_v3_merkle_proof_free(mp.ok);
    ffi_cfr_free(leaf); ffi_cfr_free(limit); ffi_cfr_free(mid);
    ffi_vec_cfr_free(keys);
    ffi_rln_v3_free(rln);
    free(zkey.ptr); free(graph.ptr);
    printf("V3-GPU-PATH-OK\n");
    return 0;
}
