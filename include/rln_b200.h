/*
 * rln_b200.h — C ABI of librln_b200.so, the B200-native drop-in for the hot path of the `rln` crate
 * of vacp2p/zerokit (proof generation / verification / Poseidon Merkle tree).
 *
 * Every `ffi_*` symbol below has the name, argument order and ownership rules of the function the
 * reference exports through safer-ffi (the reference's rln.h is generated, not checked in; the Rust
 * definition each entry replaces is cited as file:line relative to the reference root).  Conventions
 * (rln/src/ffi/ffi_utils.rs:15-36, rln/ffi_c_examples/README.md:42-48):
 *   - repr_c::Box<T>            → owning, non-null T*; parameters typed &repr_c::Box<T> are T* const*
 *   - repr_c::Vec<T>            → { T* ptr; size_t len; size_t cap; }
 *   - repr_c::String            → Vec<uint8_t> holding UTF-8 plus a trailing NUL (print via .ptr)
 *   - CResult<T,E>              → { ok; err; }: exactly one side is non-null
 *   - CBoolResult               → { bool ok; String err; } (err.ptr == NULL when there is none)
 *   - every returned object is caller-owned and released with its *_free function
 *   - errors are the reference's Display strings, never codes
 * CFr is opaque to callers (32 bytes); here it holds the canonical little-endian integer.
 *
 * The `rlnb200_*` symbols are extensions that do not exist in the reference: caller-supplied
 * blinding scalars (mirrors generate_zk_proof_with_rs, rln/src/protocol/proof.rs:753-777), batched
 * proving/verification and the raw kernels used by the benchmarks.  All computation behind this
 * header runs on the GPU; the library returns an error string if no CUDA device is usable.
 */
#ifndef RLN_B200_H
#define RLN_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ basic containers */
typedef struct Vec_uint8 { uint8_t *ptr; size_t len; size_t cap; } Vec_uint8_t;
typedef Vec_uint8_t RlnString;                         /* repr_c::String */
typedef struct Vec_size { size_t *ptr; size_t len; size_t cap; } Vec_size_t;
typedef struct Vec_bool { bool *ptr; size_t len; size_t cap; } Vec_bool_t;

typedef struct CFr { uint8_t bytes[32]; } CFr_t;       /* rln/src/ffi/ffi_utils.rs:33-36 (opaque) */
typedef struct Vec_CFr { CFr_t *ptr; size_t len; size_t cap; } Vec_CFr_t;

typedef struct FFI_RLN FFI_RLN_t;                      /* rln/src/ffi/ffi_rln.rs:16-18 */
typedef struct FFI_RLNProof FFI_RLNProof_t;            /* rln/src/ffi/ffi_rln.rs:153-155 */
typedef struct FFI_RLNProofValues FFI_RLNProofValues_t;/* rln/src/ffi/ffi_rln.rs:714-716 */
typedef struct FFI_RLNWitnessInput FFI_RLNWitnessInput_t; /* rln/src/ffi/ffi_rln.rs:322-324 */
typedef struct FFI_RLNPartialWitnessInput FFI_RLNPartialWitnessInput_t; /* rln/src/ffi/ffi_rln.rs:563-565 */
typedef struct FFI_RLNPartialProof FFI_RLNPartialProof_t;             /* rln/src/ffi/ffi_rln.rs:240-242 */
typedef struct FFI_MerkleProof {                       /* rln/src/ffi/ffi_tree.rs:13-18 */
    Vec_CFr_t path_elements;
    Vec_uint8_t path_index;
} FFI_MerkleProof_t;

typedef struct CBoolResult { bool ok; RlnString err; } CBoolResult_t;                 /* ffi_utils.rs:24-29 */
typedef struct CResult_FFI_RLN { FFI_RLN_t *ok; RlnString err; } CResult_FFI_RLN_t;   /* ffi_utils.rs:15-20 */
typedef struct CResult_FFI_RLNProof { FFI_RLNProof_t *ok; RlnString err; } CResult_FFI_RLNProof_t;
typedef struct CResult_FFI_RLNProofValues { FFI_RLNProofValues_t *ok; RlnString err; } CResult_FFI_RLNProofValues_t;
typedef struct CResult_FFI_RLNWitnessInput { FFI_RLNWitnessInput_t *ok; RlnString err; } CResult_FFI_RLNWitnessInput_t;
typedef struct CResult_FFI_RLNPartialWitnessInput { FFI_RLNPartialWitnessInput_t *ok; RlnString err; } CResult_FFI_RLNPartialWitnessInput_t;
typedef struct CResult_FFI_RLNPartialProof { FFI_RLNPartialProof_t *ok; RlnString err; } CResult_FFI_RLNPartialProof_t;
typedef struct CResult_FFI_MerkleProof { FFI_MerkleProof_t *ok; RlnString err; } CResult_FFI_MerkleProof_t;
typedef struct CResult_CFr { CFr_t *ok; RlnString err; } CResult_CFr_t;
typedef struct CResult_Vec_uint8 { Vec_uint8_t ok; RlnString err; } CResult_Vec_uint8_t;
typedef struct CResult_Vec_CFr { Vec_CFr_t ok; RlnString err; } CResult_Vec_CFr_t;
typedef struct CResult_String { RlnString ok; RlnString err; } CResult_String_t;
typedef struct Vec_String { RlnString *ptr; size_t len; size_t cap; } Vec_String_t;   /* repr_c::Vec<repr_c::String> */
typedef struct CResult_Vec_bool { Vec_bool_t ok; RlnString err; } CResult_Vec_bool_t;

/* V3 objects (rln/src/ffi/ffi_rln_v3.rs:312,614,866,1013,1097,1141,1365) */
typedef struct FFI_RLNV3 FFI_RLNV3_t;
typedef struct FFI_RLNV3WitnessInput FFI_RLNV3WitnessInput_t;
typedef struct FFI_RLNV3PartialWitnessInput FFI_RLNV3PartialWitnessInput_t;
typedef struct FFI_RLNV3Proof FFI_RLNV3Proof_t;
typedef struct FFI_RLNV3PartialProof FFI_RLNV3PartialProof_t;
typedef struct FFI_RLNV3ProofValues FFI_RLNV3ProofValues_t;
typedef struct FFI_RLNV3MerkleProof { Vec_CFr_t path_elements; Vec_uint8_t path_index; } FFI_RLNV3MerkleProof_t;
typedef struct CResult_FFI_RLNV3 { FFI_RLNV3_t *ok; RlnString err; } CResult_FFI_RLNV3_t;
typedef struct CResult_FFI_RLNV3WitnessInput { FFI_RLNV3WitnessInput_t *ok; RlnString err; } CResult_FFI_RLNV3WitnessInput_t;
typedef struct CResult_FFI_RLNV3PartialWitnessInput { FFI_RLNV3PartialWitnessInput_t *ok; RlnString err; } CResult_FFI_RLNV3PartialWitnessInput_t;
typedef struct CResult_FFI_RLNV3Proof { FFI_RLNV3Proof_t *ok; RlnString err; } CResult_FFI_RLNV3Proof_t;
typedef struct CResult_FFI_RLNV3PartialProof { FFI_RLNV3PartialProof_t *ok; RlnString err; } CResult_FFI_RLNV3PartialProof_t;
typedef struct CResult_FFI_RLNV3ProofValues { FFI_RLNV3ProofValues_t *ok; RlnString err; } CResult_FFI_RLNV3ProofValues_t;
typedef struct CResult_FFI_RLNV3MerkleProof { FFI_RLNV3MerkleProof_t *ok; RlnString err; } CResult_FFI_RLNV3MerkleProof_t;

/* The type names safer-ffi generates for the instantiations above (CResult<Ok, Err> → CResult_<Ok>_<Err>_t with Box<T> → T_ptr and
 * String → Vec_uint8), as callers of the generated rln.h spell them (rln/ffi_c_examples/common.c:8-28).  Same layouts. */
typedef CResult_FFI_RLN_t                    CResult_FFI_RLN_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNProof_t               CResult_FFI_RLNProof_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNProofValues_t         CResult_FFI_RLNProofValues_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNWitnessInput_t        CResult_FFI_RLNWitnessInput_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNPartialWitnessInput_t CResult_FFI_RLNPartialWitnessInput_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNPartialProof_t        CResult_FFI_RLNPartialProof_ptr_Vec_uint8_t;
typedef CResult_FFI_MerkleProof_t            CResult_FFI_MerkleProof_ptr_Vec_uint8_t;
typedef CResult_CFr_t                        CResult_CFr_ptr_Vec_uint8_t;
typedef CResult_Vec_uint8_t                  CResult_Vec_uint8_Vec_uint8_t;
typedef CResult_Vec_CFr_t                    CResult_Vec_CFr_Vec_uint8_t;
typedef CResult_Vec_bool_t                   CResult_Vec_bool_Vec_uint8_t;
typedef Vec_String_t                         Vec_Vec_uint8_t;
typedef CResult_FFI_RLNV3_t                    CResult_FFI_RLNV3_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3WitnessInput_t        CResult_FFI_RLNV3WitnessInput_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3PartialWitnessInput_t CResult_FFI_RLNV3PartialWitnessInput_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3Proof_t               CResult_FFI_RLNV3Proof_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3PartialProof_t        CResult_FFI_RLNV3PartialProof_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3ProofValues_t         CResult_FFI_RLNV3ProofValues_ptr_Vec_uint8_t;
typedef CResult_FFI_RLNV3MerkleProof_t         CResult_FFI_RLNV3MerkleProof_ptr_Vec_uint8_t;

/* ------------------------------------------------------------------ RLN object (rln/src/ffi/ffi_rln.rs) */
CResult_FFI_RLN_t ffi_rln_new(size_t tree_depth, const char *config_path);                      /* :22-57  */
CResult_FFI_RLN_t ffi_rln_new_with_params(size_t tree_depth, const Vec_uint8_t *zkey_data,
                                          const Vec_uint8_t *graph_data, const char *config_path); /* :74-116 */
void   ffi_rln_free(FFI_RLN_t *rln);                                                            /* :136-139 */
size_t ffi_rln_get_tree_depth(FFI_RLN_t *const *rln);                                           /* :141-144 */
size_t ffi_rln_get_max_out(FFI_RLN_t *const *rln);                                              /* :146-149 */

/* ------------------------------------------------------------------ Merkle tree (rln/src/ffi/ffi_tree.rs) */
CBoolResult_t ffi_set_tree(FFI_RLN_t **rln, size_t tree_depth);                                 /* :27-41  */
CBoolResult_t ffi_delete_leaf(FFI_RLN_t **rln, size_t index);                                   /* :43-55  */
CBoolResult_t ffi_set_leaf(FFI_RLN_t **rln, size_t index, const CFr_t *leaf);                   /* :57-69  */
CResult_CFr_t ffi_get_leaf(FFI_RLN_t *const *rln, size_t index);                                /* :71-86  */
size_t        ffi_leaves_set(FFI_RLN_t *const *rln);                                            /* :88-91  */
CBoolResult_t ffi_set_next_leaf(FFI_RLN_t **rln, const CFr_t *leaf);                            /* :93-105 */
CBoolResult_t ffi_set_leaves_from(FFI_RLN_t **rln, size_t index, const Vec_CFr_t *leaves);      /* :107-124 */
CBoolResult_t ffi_init_tree_with_leaves(FFI_RLN_t **rln, const Vec_CFr_t *leaves);              /* :126-142 */
CBoolResult_t ffi_atomic_operation(FFI_RLN_t **rln, size_t index, const Vec_CFr_t *leaves,
                                   const Vec_size_t *indices);                                  /* :146-165 */
CBoolResult_t ffi_seq_atomic_operation(FFI_RLN_t **rln, const Vec_CFr_t *leaves,
                                       const Vec_uint8_t *indices);                             /* :167-186 */
CFr_t        *ffi_get_root(FFI_RLN_t *const *rln);                                              /* :190-193 */
CResult_FFI_MerkleProof_t ffi_get_merkle_proof(FFI_RLN_t *const *rln, size_t index);            /* :195-225 */
void          ffi_merkle_proof_free(FFI_MerkleProof_t *merkle_proof);                           /* :20-23  */

/* ------------------------------------------------------------------ witness input (rln/src/ffi/ffi_rln.rs) */
CResult_FFI_RLNWitnessInput_t ffi_rln_witness_input_new_single(
    const CFr_t *identity_secret, const CFr_t *user_message_limit, const CFr_t *message_id,
    const Vec_CFr_t *path_elements, const Vec_uint8_t *identity_path_index,
    const CFr_t *x, const CFr_t *external_nullifier);                                           /* :326-358 */
CResult_FFI_RLNWitnessInput_t ffi_rln_witness_input_new_multi(
    const CFr_t *identity_secret, const CFr_t *user_message_limit, const Vec_CFr_t *message_ids,
    const Vec_CFr_t *path_elements, const Vec_uint8_t *identity_path_index, const CFr_t *x,
    const CFr_t *external_nullifier, const Vec_bool_t *selector_used);                          /* :360-396 */
CResult_Vec_uint8_t ffi_rln_witness_to_bytes_le(FFI_RLNWitnessInput_t *const *witness);         /* :476-490 */
CResult_FFI_RLNWitnessInput_t ffi_bytes_le_to_rln_witness(const Vec_uint8_t *bytes);            /* :508-522 */
void ffi_rln_witness_input_free(FFI_RLNWitnessInput_t *witness);                                /* :556-559 */

/* ------------------------------------------------------------------ proving / verifying */
CResult_FFI_RLNProof_t ffi_generate_rln_proof(FFI_RLN_t *const *rln,
                                              FFI_RLNWitnessInput_t *const *witness);           /* :851-872 */
CBoolResult_t ffi_verify_rln_proof(FFI_RLN_t *const *rln, FFI_RLNProof_t *const *rln_proof,
                                   const CFr_t *x);                                             /* :964-984 */
CBoolResult_t ffi_verify_with_roots(FFI_RLN_t *const *rln, FFI_RLNProof_t *const *rln_proof,
                                    const Vec_CFr_t *roots, const CFr_t *x);                    /* :986-1010 */

/* two-phase proving: precompute everything that does not depend on (message_id, x, external_nullifier), finish later
 * (rln/src/protocol/proof.rs:783-849, rln/src/partial_proof.rs:108-274) */
CResult_FFI_RLNPartialWitnessInput_t ffi_rln_partial_witness_input_new(
    const CFr_t *identity_secret, const CFr_t *user_message_limit, const Vec_CFr_t *path_elements,
    const Vec_uint8_t *identity_path_index);                                                    /* :567-592 */
void ffi_rln_partial_witness_input_free(FFI_RLNPartialWitnessInput_t *witness);                 /* :707-710 */
CResult_FFI_RLNPartialProof_t ffi_generate_partial_zk_proof(
    FFI_RLN_t *const *rln, FFI_RLNPartialWitnessInput_t *const *partial_witness);               /* :921-936 */
CResult_FFI_RLNProof_t ffi_finish_rln_proof(FFI_RLN_t *const *rln, FFI_RLNPartialProof_t *const *partial_proof,
                                            FFI_RLNWitnessInput_t *const *witness);             /* :938-960 */
CResult_Vec_uint8_t ffi_rln_partial_proof_to_bytes_le(FFI_RLNPartialProof_t *const *partial_proof); /* :251-265 */
void ffi_rln_partial_proof_free(FFI_RLNPartialProof_t *partial_proof);                          /* :283-286 */

FFI_RLNProofValues_t *ffi_rln_proof_get_values(FFI_RLNProof_t *const *rln_proof);               /* :157-162 */
uint8_t ffi_rln_proof_get_version_byte(FFI_RLNProof_t *const *rln_proof);                       /* :164-167 */
CResult_Vec_uint8_t ffi_rln_proof_to_bytes_le(FFI_RLNProof_t *const *rln_proof);                /* :169-183 */
CResult_Vec_uint8_t ffi_rln_proof_to_bytes_be(FFI_RLNProof_t *const *rln_proof);                /* :185-199 */
CResult_FFI_RLNProof_t ffi_bytes_le_to_rln_proof(const Vec_uint8_t *bytes);                     /* :201-215 */
void ffi_rln_proof_free(FFI_RLNProof_t *rln_proof);                                             /* :233-236 */

CFr_t *ffi_rln_proof_values_get_root(FFI_RLNProofValues_t *const *pv);                          /* :718-721 */
CFr_t *ffi_rln_proof_values_get_x(FFI_RLNProofValues_t *const *pv);                             /* :723-726 */
CFr_t *ffi_rln_proof_values_get_external_nullifier(FFI_RLNProofValues_t *const *pv);            /* :728-733 */
CResult_CFr_t ffi_rln_proof_values_get_y(FFI_RLNProofValues_t *const *pv);                      /* :735-743 */
CResult_CFr_t ffi_rln_proof_values_get_nullifier(FFI_RLNProofValues_t *const *pv);              /* :745-753 */
CResult_Vec_CFr_t ffi_rln_proof_values_get_ys(FFI_RLNProofValues_t *const *pv);                 /* :765-779 */
CResult_Vec_CFr_t ffi_rln_proof_values_get_nullifiers(FFI_RLNProofValues_t *const *pv);         /* :781-795 */
CResult_Vec_uint8_t ffi_rln_proof_values_get_selector_used(FFI_RLNProofValues_t *const *pv);    /* :755-763 (Vec<bool>) */
uint8_t ffi_rln_proof_values_get_version_byte(FFI_RLNProofValues_t *const *pv);                 /* :797-800 */
Vec_uint8_t ffi_rln_proof_values_to_bytes_le(FFI_RLNProofValues_t *const *pv);                  /* :802-805 */
CResult_FFI_RLNProofValues_t ffi_bytes_le_to_rln_proof_values(const Vec_uint8_t *bytes);        /* :812-826 */
void ffi_rln_proof_values_free(FFI_RLNProofValues_t *proof_values);                             /* :844-847 */

/* ------------------------------------------------------------------ CFr / Vec helpers (rln/src/ffi/ffi_utils.rs) */
CFr_t *ffi_cfr_zero(void);                                                                      /* :69-72  */
CFr_t *ffi_cfr_one(void);                                                                       /* :74-77  */
CResult_Vec_uint8_t ffi_cfr_to_bytes_le(const CFr_t *cfr);                                      /* :79-92  */
CResult_Vec_uint8_t ffi_cfr_to_bytes_be(const CFr_t *cfr);                                      /* :94-107 */
CResult_CFr_t ffi_bytes_le_to_cfr(const Vec_uint8_t *bytes);                                    /* :109-121 */
CResult_CFr_t ffi_bytes_be_to_cfr(const Vec_uint8_t *bytes);                                    /* :123-135 */
CFr_t *ffi_uint_to_cfr(uint32_t value);                                                         /* :137-140 */
RlnString ffi_cfr_debug(const CFr_t *cfr);                                                      /* :142-148 */
void   ffi_cfr_free(CFr_t *cfr);                                                                /* :150-153 */
Vec_CFr_t ffi_vec_cfr_new(size_t capacity);                                                     /* :157-160 */
Vec_CFr_t ffi_vec_cfr_from_cfr(const CFr_t *cfr);                                               /* :162-165 */
void   ffi_vec_cfr_push(Vec_CFr_t *v, const CFr_t *cfr);                                        /* :167-175 */
size_t ffi_vec_cfr_len(const Vec_CFr_t *v);                                                     /* :177-180 */
const CFr_t *ffi_vec_cfr_get(const Vec_CFr_t *v, size_t i);                                     /* :182-185 */
void   ffi_vec_cfr_free(Vec_CFr_t v);                                                           /* :268-271 */
void   ffi_vec_u8_free(Vec_uint8_t v);                                                          /* :341-344 */
void   ffi_c_string_free(RlnString s);                                                          /* :406-409 */
CFr_t *ffi_hash_to_field_le(const Vec_uint8_t *input);                                          /* :348-351 */
CFr_t *ffi_hash_to_field_be(const Vec_uint8_t *input);                                          /* :353-356 */
CFr_t *ffi_poseidon_hash_pair(const CFr_t *a, const CFr_t *b);                                  /* :358-361 */
Vec_CFr_t ffi_key_gen(void);                                                                    /* :365-369 */
Vec_CFr_t ffi_seeded_key_gen(const Vec_uint8_t *seed);                                          /* :371-377 */
Vec_CFr_t ffi_extended_key_gen(void);                                                           /* :379-389: trapdoor, nullifier, secret, commitment */
Vec_CFr_t ffi_seeded_extended_key_gen(const Vec_uint8_t *seed);                                 /* :391-404 */
CResult_Vec_uint8_t ffi_vec_cfr_to_bytes_le(const Vec_CFr_t *v);                                /* :187-201: u64 LE count | 32-byte LE elements */
CResult_Vec_uint8_t ffi_vec_cfr_to_bytes_be(const Vec_CFr_t *v);                                /* :203-217: u64 BE count | 32-byte BE elements */
CResult_Vec_CFr_t ffi_bytes_le_to_vec_cfr(const Vec_uint8_t *bytes);                            /* :219-236 */
CResult_Vec_CFr_t ffi_bytes_be_to_vec_cfr(const Vec_uint8_t *bytes);                            /* :238-255 */
RlnString ffi_vec_cfr_debug(const Vec_CFr_t *v);                                                /* :257-266 */
CResult_Vec_uint8_t ffi_vec_u8_to_bytes_le(const Vec_uint8_t *v);                               /* :275-288 */
CResult_Vec_uint8_t ffi_vec_u8_to_bytes_be(const Vec_uint8_t *v);                               /* :290-303 */
CResult_Vec_uint8_t ffi_bytes_le_to_vec_u8(const Vec_uint8_t *bytes);                           /* :305-317 */
CResult_Vec_uint8_t ffi_bytes_be_to_vec_u8(const Vec_uint8_t *bytes);                           /* :319-331 */
RlnString ffi_vec_u8_debug(const Vec_uint8_t *v);                                               /* :333-339 */

/* ---- rln/src/ffi/ffi_rln.rs, continued: record getters, big-endian wire formats, partial-witness records, identity-secret
 * recovery; rln/src/ffi/ffi_tree.rs:226-268 metadata.  Getters on the wrong variant abort like the reference's todo!(). */
uint8_t ffi_rln_witness_input_get_version_byte(FFI_RLNWitnessInput_t *const *witness);           /* ffi_rln.rs:398-401 */
CFr_t *ffi_rln_witness_input_get_identity_secret(FFI_RLNWitnessInput_t *const *witness);         /* :403-408 */
CFr_t *ffi_rln_witness_input_get_user_message_limit(FFI_RLNWitnessInput_t *const *witness);      /* :410-415 */
CFr_t *ffi_rln_witness_input_get_message_id(FFI_RLNWitnessInput_t *const *witness);              /* :417-422 (SingleV1) */
Vec_CFr_t ffi_rln_witness_input_get_message_ids(FFI_RLNWitnessInput_t *const *witness);          /* :424-435 (MultiV1) */
Vec_CFr_t ffi_rln_witness_input_get_path_elements(FFI_RLNWitnessInput_t *const *witness);        /* :437-448 */
Vec_uint8_t ffi_rln_witness_input_get_identity_path_index(FFI_RLNWitnessInput_t *const *witness);/* :450-455 */
CFr_t *ffi_rln_witness_input_get_x(FFI_RLNWitnessInput_t *const *witness);                       /* :457-460 */
CFr_t *ffi_rln_witness_input_get_external_nullifier(FFI_RLNWitnessInput_t *const *witness);      /* :462-467 */
Vec_bool_t ffi_rln_witness_input_get_selector_used(FFI_RLNWitnessInput_t *const *witness);       /* :469-474 (MultiV1) */
CResult_Vec_uint8_t ffi_rln_witness_to_bytes_be(FFI_RLNWitnessInput_t *const *witness);          /* :492-506 */
CResult_FFI_RLNWitnessInput_t ffi_bytes_be_to_rln_witness(const Vec_uint8_t *bytes);             /* :524-538 */
CResult_String_t ffi_rln_witness_to_bigint_json(FFI_RLNWitnessInput_t *const *witness);          /* :540-554 */
uint8_t ffi_rln_partial_witness_input_get_version_byte(FFI_RLNPartialWitnessInput_t *const *w);  /* :594-599 */
CFr_t *ffi_rln_partial_witness_input_get_identity_secret(FFI_RLNPartialWitnessInput_t *const *w);/* :601-606 */
CFr_t *ffi_rln_partial_witness_input_get_user_message_limit(FFI_RLNPartialWitnessInput_t *const *w); /* :608-613 */
Vec_CFr_t ffi_rln_partial_witness_input_get_path_elements(FFI_RLNPartialWitnessInput_t *const *w);   /* :615-626 */
Vec_uint8_t ffi_rln_partial_witness_input_get_identity_path_index(FFI_RLNPartialWitnessInput_t *const *w); /* :628-633 */
FFI_RLNPartialWitnessInput_t *ffi_rln_witness_to_partial_witness(FFI_RLNWitnessInput_t *const *witness);  /* :635-641 */
CResult_Vec_uint8_t ffi_rln_partial_witness_to_bytes_le(FFI_RLNPartialWitnessInput_t *const *w); /* :643-657 */
CResult_Vec_uint8_t ffi_rln_partial_witness_to_bytes_be(FFI_RLNPartialWitnessInput_t *const *w); /* :659-673 */
CResult_FFI_RLNPartialWitnessInput_t ffi_bytes_le_to_rln_partial_witness(const Vec_uint8_t *bytes); /* :675-689 */
CResult_FFI_RLNPartialWitnessInput_t ffi_bytes_be_to_rln_partial_witness(const Vec_uint8_t *bytes); /* :691-705 */
Vec_uint8_t ffi_rln_proof_values_to_bytes_be(FFI_RLNProofValues_t *const *pv);                   /* :807-810 */
CResult_FFI_RLNProofValues_t ffi_bytes_be_to_rln_proof_values(const Vec_uint8_t *bytes);         /* :828-842 */
CResult_FFI_RLNProof_t ffi_bytes_be_to_rln_proof(const Vec_uint8_t *bytes);                      /* :217-231 */
uint8_t ffi_rln_partial_proof_get_version_byte(FFI_RLNPartialProof_t *const *partial_proof);     /* :244-249 */
CResult_Vec_uint8_t ffi_rln_partial_proof_to_bytes_be(FFI_RLNPartialProof_t *const *partial_proof); /* :288-302 (= LE form) */
CResult_FFI_RLNPartialProof_t ffi_bytes_le_to_rln_partial_proof(const Vec_uint8_t *bytes);       /* :267-281 */
CResult_FFI_RLNPartialProof_t ffi_bytes_be_to_rln_partial_proof(const Vec_uint8_t *bytes);       /* :304-318 */
CResult_CFr_t ffi_compute_id_secret(const CFr_t *share1_x, const CFr_t *share1_y, const CFr_t *share2_x,
                                    const CFr_t *share2_y);                                      /* :1014-1033 */
CResult_CFr_t ffi_recover_id_secret(FFI_RLNProofValues_t *const *proof_values_1,
                                    FFI_RLNProofValues_t *const *proof_values_2);                /* :1035-1049 */
/* proof from a witness calculated outside (snarkjs / circom): one decimal string per wire; the graph evaluation is skipped */
CResult_FFI_RLNProof_t ffi_generate_rln_proof_with_witness(FFI_RLN_t *const *rln, const Vec_String_t *calculated_witness,
                                                           FFI_RLNWitnessInput_t *const *witness);              /* :874-916 */
CResult_FFI_RLNProof_t rlnb200_generate_rln_proof_with_witness_rs(FFI_RLN_t *const *rln, const Vec_String_t *calculated_witness,
                                                                  FFI_RLNWitnessInput_t *const *witness, const CFr_t *r, const CFr_t *s);
CBoolResult_t ffi_set_metadata(FFI_RLN_t **rln, const Vec_uint8_t *metadata);                    /* ffi_tree.rs:228-240 */
CResult_Vec_uint8_t ffi_get_metadata(FFI_RLN_t *const *rln);                                     /* ffi_tree.rs:242-254 */
CBoolResult_t ffi_flush(FFI_RLN_t **rln);                                                        /* ffi_tree.rs:256-268 */

/* ================================================================== rln/src/ffi/ffi_rln_v3.rs: the V3 twins
 * Same prover, tree and records behind the V3 names.  The three stateful tree flavours are the one HBM tree (same roots and
 * paths); the *_default constructors return NULL when no GPU is usable (the reference's are infallible); wire formats are
 * documented in zerokit_b200/csrc/rln_ffi_v3.inc. */
FFI_RLNV3_t *ffi_rln_v3_new_stateless_default(void);                                                                  /* :323-327 */
CResult_FFI_RLNV3_t ffi_rln_v3_new_stateless(const Vec_uint8_t *zkey_data, const Vec_uint8_t *graph_data);            /* :329-346 */
FFI_RLNV3_t *ffi_rln_v3_new_with_full_merkle_tree_default(void);                                                      /* :349-354 */
CResult_FFI_RLNV3_t ffi_rln_v3_new_with_full_merkle_tree(size_t tree_depth, const Vec_uint8_t *zkey_data,
                                                         const Vec_uint8_t *graph_data);                              /* :356-387 */
FFI_RLNV3_t *ffi_rln_v3_new_with_optimal_merkle_tree_default(void);                                                   /* :390-396 */
CResult_FFI_RLNV3_t ffi_rln_v3_new_with_optimal_merkle_tree(size_t tree_depth, const Vec_uint8_t *zkey_data,
                                                            const Vec_uint8_t *graph_data);                           /* :398-429 */
FFI_RLNV3_t *ffi_rln_v3_new_with_pm_tree_default(void);                                                               /* :432-437 */
CResult_FFI_RLNV3_t ffi_rln_v3_new_with_pm_tree(size_t tree_depth, const Vec_uint8_t *zkey_data, const Vec_uint8_t *graph_data,
                                                const char *config_path);                                            /* :439-504 */
void ffi_rln_v3_free(FFI_RLNV3_t *rln);                                                                               /* :605-608 */
CResult_FFI_RLNV3Proof_t ffi_rln_v3_generate_proof(FFI_RLNV3_t *const *rln, FFI_RLNV3WitnessInput_t *const *witness); /* :506-521 */
CBoolResult_t ffi_rln_v3_verify(FFI_RLNV3_t *const *rln, FFI_RLNV3Proof_t *const *rln_proof, const CFr_t *x);         /* :523-545 */
CBoolResult_t ffi_rln_v3_verify_with_roots(FFI_RLNV3_t *const *rln, FFI_RLNV3Proof_t *const *rln_proof, const Vec_CFr_t *roots,
                                           const CFr_t *x);                                                           /* :547-568 */
CResult_FFI_RLNV3PartialProof_t ffi_rln_v3_generate_partial_proof(FFI_RLNV3_t *const *rln,
                                                                  FFI_RLNV3PartialWitnessInput_t *const *partial_witness); /* :570-585 */
CResult_FFI_RLNV3Proof_t ffi_rln_v3_finish_proof(FFI_RLNV3_t *const *rln, FFI_RLNV3PartialProof_t *const *partial_proof,
                                                 FFI_RLNV3WitnessInput_t *const *witness);                            /* :587-603 */
CResult_FFI_RLNV3WitnessInput_t ffi_rln_v3_witness_input_new_single(const CFr_t *identity_secret, const CFr_t *user_message_limit,
        const CFr_t *message_id, const Vec_CFr_t *path_elements, const Vec_uint8_t *identity_path_index, const CFr_t *x,
        const CFr_t *external_nullifier);                                                                             /* :616-649 */
CResult_FFI_RLNV3WitnessInput_t ffi_rln_v3_witness_input_new_multi(const CFr_t *identity_secret, const CFr_t *user_message_limit,
        const Vec_CFr_t *message_ids, const Vec_CFr_t *path_elements, const Vec_uint8_t *identity_path_index, const CFr_t *x,
        const CFr_t *external_nullifier, const Vec_bool_t *selector_used);                                            /* :651-688 */
CFr_t *ffi_rln_v3_witness_input_get_identity_secret(FFI_RLNV3WitnessInput_t *const *witness);                         /* :690-695 */
CFr_t *ffi_rln_v3_witness_input_get_user_message_limit(FFI_RLNV3WitnessInput_t *const *witness);                      /* :697-702 */
CResult_CFr_t ffi_rln_v3_witness_input_get_message_id(FFI_RLNV3WitnessInput_t *const *witness);                       /* :704-718 */
CResult_Vec_CFr_t ffi_rln_v3_witness_input_get_message_ids(FFI_RLNV3WitnessInput_t *const *witness);                  /* :720-739 */
Vec_CFr_t ffi_rln_v3_witness_input_get_path_elements(FFI_RLNV3WitnessInput_t *const *witness);                        /* :741-752 */
Vec_uint8_t ffi_rln_v3_witness_input_get_identity_path_index(FFI_RLNV3WitnessInput_t *const *witness);                /* :754-759 */
CFr_t *ffi_rln_v3_witness_input_get_x(FFI_RLNV3WitnessInput_t *const *witness);                                       /* :761-766 */
CFr_t *ffi_rln_v3_witness_input_get_external_nullifier(FFI_RLNV3WitnessInput_t *const *witness);                      /* :768-773 */
CResult_Vec_bool_t ffi_rln_v3_witness_input_get_selector_used(FFI_RLNV3WitnessInput_t *const *witness);               /* :775-789 */
CResult_Vec_uint8_t ffi_rln_v3_witness_to_bytes_le(FFI_RLNV3WitnessInput_t *const *witness);                          /* :791-806 */
CResult_Vec_uint8_t ffi_rln_v3_witness_to_bytes_be(FFI_RLNV3WitnessInput_t *const *witness);                          /* :808-823 */
CResult_FFI_RLNV3WitnessInput_t ffi_bytes_le_to_rln_v3_witness(const Vec_uint8_t *bytes);                             /* :825-839 */
CResult_FFI_RLNV3WitnessInput_t ffi_bytes_be_to_rln_v3_witness(const Vec_uint8_t *bytes);                             /* :841-855 */
void ffi_rln_v3_witness_input_free(FFI_RLNV3WitnessInput_t *witness);                                                 /* :857-860 */
CResult_FFI_RLNV3PartialWitnessInput_t ffi_rln_v3_partial_witness_input_new(const CFr_t *identity_secret,
        const CFr_t *user_message_limit, const Vec_CFr_t *path_elements, const Vec_uint8_t *identity_path_index);     /* :868-894 */
CFr_t *ffi_rln_v3_partial_witness_input_get_identity_secret(FFI_RLNV3PartialWitnessInput_t *const *w);                /* :896-901 */
CFr_t *ffi_rln_v3_partial_witness_input_get_user_message_limit(FFI_RLNV3PartialWitnessInput_t *const *w);             /* :903-908 */
Vec_CFr_t ffi_rln_v3_partial_witness_input_get_path_elements(FFI_RLNV3PartialWitnessInput_t *const *w);               /* :910-921 */
Vec_uint8_t ffi_rln_v3_partial_witness_input_get_identity_path_index(FFI_RLNV3PartialWitnessInput_t *const *w);       /* :923-928 */
FFI_RLNV3PartialWitnessInput_t *ffi_rln_v3_witness_to_partial_witness(FFI_RLNV3WitnessInput_t *const *witness);       /* :930-936 */
CResult_Vec_uint8_t ffi_rln_v3_partial_witness_to_bytes_le(FFI_RLNV3PartialWitnessInput_t *const *w);                 /* :938-953 */
CResult_Vec_uint8_t ffi_rln_v3_partial_witness_to_bytes_be(FFI_RLNV3PartialWitnessInput_t *const *w);                 /* :955-970 */
CResult_FFI_RLNV3PartialWitnessInput_t ffi_bytes_le_to_rln_v3_partial_witness(const Vec_uint8_t *bytes);              /* :972-986 */
CResult_FFI_RLNV3PartialWitnessInput_t ffi_bytes_be_to_rln_v3_partial_witness(const Vec_uint8_t *bytes);              /* :988-1002 */
void ffi_rln_v3_partial_witness_input_free(FFI_RLNV3PartialWitnessInput_t *w);                                        /* :1004-1007 */
FFI_RLNV3ProofValues_t *ffi_rln_v3_proof_get_values(FFI_RLNV3Proof_t *const *rln_proof);                              /* :1015-1020 */
CResult_Vec_uint8_t ffi_rln_v3_proof_to_bytes_le(FFI_RLNV3Proof_t *const *rln_proof);                                 /* :1022-1037 */
CResult_Vec_uint8_t ffi_rln_v3_proof_to_bytes_mixed(FFI_RLNV3Proof_t *const *rln_proof);                              /* :1039-1054 */
CResult_FFI_RLNV3Proof_t ffi_bytes_le_to_rln_v3_proof(const Vec_uint8_t *bytes);                                      /* :1056-1070 */
CResult_FFI_RLNV3Proof_t ffi_bytes_mixed_to_rln_v3_proof(const Vec_uint8_t *bytes);                                   /* :1072-1086 */
void ffi_rln_v3_proof_free(FFI_RLNV3Proof_t *rln_proof);                                                              /* :1088-1091 */
CResult_Vec_uint8_t ffi_rln_v3_partial_proof_to_bytes_le(FFI_RLNV3PartialProof_t *const *partial_proof);              /* :1099-1114 */
CResult_FFI_RLNV3PartialProof_t ffi_bytes_le_to_rln_v3_partial_proof(const Vec_uint8_t *bytes);                       /* :1116-1130 */
void ffi_rln_v3_partial_proof_free(FFI_RLNV3PartialProof_t *partial_proof);                                           /* :1132-1135 */
CFr_t *ffi_rln_v3_proof_values_get_root(FFI_RLNV3ProofValues_t *const *pv);                                           /* :1143-1148 */
CFr_t *ffi_rln_v3_proof_values_get_x(FFI_RLNV3ProofValues_t *const *pv);                                              /* :1150-1153 */
CFr_t *ffi_rln_v3_proof_values_get_external_nullifier(FFI_RLNV3ProofValues_t *const *pv);                             /* :1155-1160 */
CResult_CFr_t ffi_rln_v3_proof_values_get_y(FFI_RLNV3ProofValues_t *const *pv);                                       /* :1162-1176 */
CResult_CFr_t ffi_rln_v3_proof_values_get_nullifier(FFI_RLNV3ProofValues_t *const *pv);                               /* :1178-1192 */
CResult_Vec_bool_t ffi_rln_v3_proof_values_get_selector_used(FFI_RLNV3ProofValues_t *const *pv);                      /* :1194-1208 */
CResult_Vec_CFr_t ffi_rln_v3_proof_values_get_ys(FFI_RLNV3ProofValues_t *const *pv);                                  /* :1210-1229 */
CResult_Vec_CFr_t ffi_rln_v3_proof_values_get_nullifiers(FFI_RLNV3ProofValues_t *const *pv);                          /* :1231-1250 */
CResult_Vec_uint8_t ffi_rln_v3_proof_values_to_bytes_le(FFI_RLNV3ProofValues_t *const *pv);                           /* :1252-1267 */
CResult_Vec_uint8_t ffi_rln_v3_proof_values_to_bytes_be(FFI_RLNV3ProofValues_t *const *pv);                           /* :1269-1284 */
CResult_FFI_RLNV3ProofValues_t ffi_bytes_le_to_rln_v3_proof_values(const Vec_uint8_t *bytes);                         /* :1286-1300 */
CResult_FFI_RLNV3ProofValues_t ffi_bytes_be_to_rln_v3_proof_values(const Vec_uint8_t *bytes);                         /* :1302-1316 */
void ffi_rln_v3_proof_values_free(FFI_RLNV3ProofValues_t *proof_values);                                              /* :1318-1321 */
CResult_CFr_t ffi_rln_v3_compute_id_secret(const CFr_t *share1_x, const CFr_t *share1_y, const CFr_t *share2_x,
                                           const CFr_t *share2_y);                                                    /* :1323-1342 */
CResult_CFr_t ffi_rln_v3_recover_id_secret(FFI_RLNV3ProofValues_t *const *proof_values_1,
                                           FFI_RLNV3ProofValues_t *const *proof_values_2);                            /* :1344-1361 */
void ffi_rln_v3_merkle_proof_free(FFI_RLNV3MerkleProof_t *merkle_proof);                                              /* :1370-1373 */
CBoolResult_t ffi_rln_v3_delete_leaf(FFI_RLNV3_t **rln, size_t index);                                                /* :1375-1387 */
CBoolResult_t ffi_rln_v3_set_leaf(FFI_RLNV3_t **rln, size_t index, const CFr_t *leaf);                                /* :1389-1405 */
CResult_CFr_t ffi_rln_v3_get_leaf(FFI_RLNV3_t *const *rln, size_t index);                                             /* :1407-1422 */
size_t ffi_rln_v3_leaves_set(FFI_RLNV3_t *const *rln);                                                                /* :1424-1427 */
CBoolResult_t ffi_rln_v3_set_next_leaf(FFI_RLNV3_t **rln, const CFr_t *leaf);                                         /* :1429-1441 */
CBoolResult_t ffi_rln_v3_set_leaves_from(FFI_RLNV3_t **rln, size_t index, const Vec_CFr_t *leaves);                   /* :1443-1460 */
CBoolResult_t ffi_rln_v3_init_tree_with_leaves(FFI_RLNV3_t **rln, const Vec_CFr_t *leaves);                           /* :1462-1478 */
CBoolResult_t ffi_rln_v3_atomic_operation(FFI_RLNV3_t **rln, size_t index, const Vec_CFr_t *leaves,
                                          const Vec_size_t *indices);                                                 /* :1480-1499 */
CBoolResult_t ffi_rln_v3_seq_atomic_operation(FFI_RLNV3_t **rln, const Vec_CFr_t *leaves, const Vec_uint8_t *indices);/* :1501-1528 */
CFr_t *ffi_rln_v3_get_root(FFI_RLNV3_t *const *rln);                                                                  /* :1530-1534 */
CResult_FFI_RLNV3MerkleProof_t ffi_rln_v3_get_merkle_proof(FFI_RLNV3_t *const *rln, size_t index);                    /* :1536-1562 */
CBoolResult_t ffi_rln_v3_set_metadata(FFI_RLNV3_t **rln, const Vec_uint8_t *metadata);                                /* :1564-1579 */
CResult_Vec_uint8_t ffi_rln_v3_get_metadata(FFI_RLNV3_t *const *rln);                                                 /* :1581-1595 */
CBoolResult_t ffi_rln_v3_flush(FFI_RLNV3_t **rln);                                                                    /* :1597-1609 */
/* extensions: caller-supplied (r, s) for reproducible V3 proofs */
CResult_FFI_RLNV3Proof_t rlnb200_v3_generate_proof_with_rs(FFI_RLNV3_t *const *rln, FFI_RLNV3WitnessInput_t *const *witness,
                                                           const CFr_t *r, const CFr_t *s);
CResult_FFI_RLNV3Proof_t rlnb200_v3_finish_proof_with_rs(FFI_RLNV3_t *const *rln, FFI_RLNV3PartialProof_t *const *partial_proof,
                                                         FFI_RLNV3WitnessInput_t *const *witness, const CFr_t *r, const CFr_t *s);

/* ================================================================== extensions (not in the reference) */

/* generate_zk_proof_with_rs (rln/src/protocol/proof.rs:753-777) behind the ABI: r, s supplied by the
 * caller so that proofs are reproducible bit for bit. */
CResult_FFI_RLNProof_t rlnb200_generate_rln_proof_with_rs(FFI_RLN_t *const *rln, FFI_RLNWitnessInput_t *const *witness,
                                                          const CFr_t *r, const CFr_t *s);

/* finish_zk_proof_with_rs (rln/src/protocol/proof.rs:822-849) behind the ABI */
CResult_FFI_RLNProof_t rlnb200_finish_rln_proof_with_rs(FFI_RLN_t *const *rln, FFI_RLNPartialProof_t *const *partial_proof,
                                                        FFI_RLNWitnessInput_t *const *witness, const CFr_t *r, const CFr_t *s);
/* ffi_bytes_le_to_rln_partial_proof with a handle (serialises with that handle's other GPU work) */
CResult_FFI_RLNPartialProof_t rlnb200_bytes_le_to_rln_partial_proof(FFI_RLN_t *const *rln, const Vec_uint8_t *bytes);
/* batched two-phase proving on host buffers: witness records as for rlnb200_prove_batch (message_id / x / external_nullifier
 * are ignored by the partial phase); partial points are n × 320 bytes (canonical affine π_a 64 | ρ 64 | π_b 128 | π_c 64) */
int rlnb200_partial_batch(FFI_RLN_t *const *rln, const uint8_t *witnesses, size_t n, uint8_t *partial_out, RlnString *err);
int rlnb200_finish_batch(FFI_RLN_t *const *rln, const uint8_t *witnesses, size_t n, const uint8_t *partial, const uint8_t *rs,
                         uint8_t *proofs_out, RlnString *err);

/* handle over the bundled multi message-id circuit (rln/resources/tree_depth_20/multi_message_id/max_out_4) */
CResult_FFI_RLN_t rlnb200_rln_new_multi(size_t tree_depth, size_t max_out);
/* byte length of one rln_witness_to_bytes_le / rln_proof_to_bytes_le record for this handle's circuit and message mode */
size_t rlnb200_witness_record_len(FFI_RLN_t *const *rln);
size_t rlnb200_proof_record_len(FFI_RLN_t *const *rln);

/* Batched proving, HOST buffers.  witnesses: n concatenated rln_witness_to_bytes_le records (single
 * message-id layout, rln/src/protocol/witness.rs:369-415; all of length 1+32*(5+depth)+16+depth).
 * rs: n*64 bytes (r|s, canonical LE) or NULL for fresh randomness.  proofs_out: n records of
 * rln_proof_to_bytes_le (1+128+1+160 = 290 bytes, rln/src/protocol/proof.rs:413-428).
 * Returns 0 on success; otherwise a negative value and *err (free with ffi_c_string_free). */
int rlnb200_prove_batch(FFI_RLN_t *const *rln, const uint8_t *witnesses, size_t n, const uint8_t *rs,
                        uint8_t *proofs_out, RlnString *err);
/* The same with the wire records resident in HBM (e.g. received from another GPU over NCCL): n witness records in, n
 * rln_proof_to_bytes_le records out; parsing, validation and formatting run on the device (k_records.cu).  d_rs: n*64 bytes
 * on the device, or NULL.  This is the entry point a one-process-per-GPU deployment calls between its scatter and its gather
 * (zerokit_b200/sharding.py; BASELINE.json configs[4]). */
int rlnb200_prove_records_device(FFI_RLN_t *const *rln, const void *d_witness_records, const void *d_rs, size_t n,
                                 void *d_proof_records, void *stream, RlnString *err);
/* Batched verification of n rln_proof_to_bytes_le records against the handle's verifying key only
 * (no root / signal check): ok_out[i] = 1 valid, 0 invalid, 2 malformed. */
int rlnb200_verify_batch(FFI_RLN_t *const *rln, const uint8_t *proofs, size_t n, uint8_t *ok_out, RlnString *err);

/* Batched proving, DEVICE buffers (inputs already resident in HBM):
 *   d_inputs  n × input_slots × 32 bytes, canonical LE, the witness-graph input buffer layout
 *             (rln/src/circuit/iden3calc.rs:106-181; slot 0 = 1) — see rlnb200_input_slot()
 *   d_rs      n × 64 bytes
 *   d_proofs  n × 128 bytes  (ark-compressed A|B|C)
 *   d_values  n × 160 bytes  ([root, external_nullifier, x, y, nullifier], canonical LE) or NULL
 *   d_affine  n × 256 bytes  (A|B|C affine canonical) or NULL
 * stream: a cudaStream_t (0 = default stream). */
int rlnb200_prove_batch_device(FFI_RLN_t *const *rln, const void *d_inputs, const void *d_rs, size_t n,
                               void *d_proofs, void *d_values, void *d_affine, void *stream, RlnString *err);
/* two-phase proving on DEVICE buffers: d_partial_affine n × 320 bytes, d_partial_compressed n × 160 bytes (ark compressed).
 * The partial phase ignores the messageId / x / externalNullifier slots of d_inputs. */
int rlnb200_partial_batch_device(FFI_RLN_t *const *rln, const void *d_inputs, size_t n, void *d_partial_affine,
                                 void *d_partial_compressed, void *stream, RlnString *err);
int rlnb200_finish_batch_device(FFI_RLN_t *const *rln, const void *d_inputs, const void *d_rs, const void *d_partial_affine, size_t n,
                                void *d_proofs, void *d_values, void *stream, RlnString *err);
/* fills the input-slot buffer for one witness record on the host (layout helper for callers/tests) */
int rlnb200_witness_to_input_slots(FFI_RLN_t *const *rln, const uint8_t *witness_le, size_t len, uint8_t *slots_out,
                                   RlnString *err);
size_t rlnb200_input_slots(FFI_RLN_t *const *rln);
/* depth of the stateful Merkle tree (differs from the circuit depth only after ffi_set_tree) */
size_t rlnb200_state_tree_depth(FFI_RLN_t *const *rln);
/* offset/len of a named circuit input ("identitySecret", "pathElements", …); returns 0 if unknown */
int rlnb200_input_slot(FFI_RLN_t *const *rln, const char *name, uint32_t *offset, uint32_t *len);
/* reserve workspace for batches of up to max_batch proofs (otherwise grown on demand) */
int rlnb200_reserve(FFI_RLN_t *const *rln, size_t max_batch, RlnString *err);
/* number of kernel launches issued by this library since load (bench bookkeeping) */
uint64_t rlnb200_launch_count(void);
/* CUDA-event timings (ms) of the last rlnb200_prove_batch_device call, in launch order: witness VM, QAP
 * (matvec + 6 NTTs), G1 table accumulate, G1 reduce, G2 accumulate, G2 reduce, assembly, proof values */
void rlnb200_last_stage_ms(FFI_RLN_t *const *rln, float out[8]);
/* how many device batches (= launches of each stage's kernels) the timings above are summed over: 1 after
 * rlnb200_prove_batch_device up to the handle's maximum batch, ceil(n / 4096) after the record-based calls */
uint32_t rlnb200_last_stage_batches(FFI_RLN_t *const *rln);
/* selects the CUDA device used by subsequently created objects (call before ffi_rln_new) */
int rlnb200_set_device(int device, RlnString *err);
/* fixed-base table geometry: window bits c / windows K of the G1 and of the G2 tables, number of (non-infinity)
 * G1 / G2 bases, bytes in HBM */
int rlnb200_table_info(FFI_RLN_t *const *rln, int *window_bits, int *windows, uint64_t *g1_bases, uint64_t *g2_bases,
                       uint64_t *table_bytes, int *window_bits_g2, int *windows_g2);
/* verifier path (rln/src/protocol/proof.rs:856-894 has one; this library has two kernels with the same results): batches of up to
 * `max_batch` proofs — a single ffi_verify* call is a batch of one — run on the lane-parallel kernel (one CTA per proof, a
 * host-scheduled program of sums of products, k_verify_vm.cu), larger ones on the one-thread-per-proof kernel (k_verify.cu);
 * 0 selects the latter always.  Default 4096 (measured break-even on a B200), or RLN_B200_VERIFY_VM_MAX.  info: levels, slots, constants of the program. */
int rlnb200_set_verify_vm_max(FFI_RLN_t *const *rln, size_t max_batch);
int rlnb200_verify_vm_info(FFI_RLN_t *const *rln, uint32_t *levels, uint32_t *slots, uint32_t *constants);
/* diagnostic: one proof record through the lane-parallel kernel with a clock64() sample per level of the program.
 * cycles[levels + 1]; meta[levels] = terms per lane of warp 0..3 (4 bits each) | conditional subtractions << 16 | lane-pair
 * combine << 18 | special id << 20 */
int rlnb200_verify_vm_trace(FFI_RLN_t *const *rln, const uint8_t *proof_record, long long *cycles, uint32_t *meta, uint8_t *ok_out,
                            RlnString *err);
/* RLN::get_subtree_root (rln/src/public.rs:877-883; utils/src/merkle_tree/full_merkle_tree.rs:157-184): the ancestor at `level`
 * (0 = root, tree depth = the leaf itself) of leaf `index`, canonical 32 bytes */
int rlnb200_get_subtree_root(FFI_RLN_t *const *rln, size_t level, size_t index, uint8_t *out32, RlnString *err);
/* RLN::get_empty_leaves_indices (public.rs:885-887): indices below leaves_set() that were deleted or never set;
 * free the result with rlnb200_vec_usize_free */
int rlnb200_get_empty_leaves_indices(FFI_RLN_t *const *rln, Vec_size_t *out, RlnString *err);
void rlnb200_vec_usize_free(Vec_size_t v);
/* 1 when the G1 scalars are GLV-split (k = k1 + k2*lambda, two visits of `windows` windows per base; RLN_B200_GLV=0
 * disables it) */
int rlnb200_glv_enabled(FFI_RLN_t *const *rln);
/* self-test of the split kernel: n canonical 32-byte scalars -> n x 36 bytes (|k1| 16 B LE, |k2| 16 B LE, sign1, sign2,
 * 2 pad bytes) with k = (+-k1) + (+-k2)*lambda mod r and |ki| < 2^128 */
int rlnb200_glv_split(const uint8_t *scalars_le, size_t n, uint8_t *out36, RlnString *err);
/* self-test of the Straus / GLV double multiplication used by the proof assembly (s*g_a + r*g1_b): n items of
 * P (64 B canonical affine) | kp (32 B) | Q (64 B) | kq (32 B) -> n x 64 B canonical affine kp*P + kq*Q (use_q = 0: kp*P) */
int rlnb200_glv_double_mul(const uint8_t *items192, size_t n, int use_q, uint8_t *out64, RlnString *err);

/* Merkle tree bulk operations on device/host buffers (FullMerkleTree semantics,
 * utils/src/merkle_tree/full_merkle_tree.rs:197-223,288-304) */
int rlnb200_set_leaves_from_bytes(FFI_RLN_t **rln, size_t index, const uint8_t *leaves_le, size_t count, RlnString *err);
int rlnb200_get_merkle_proofs(FFI_RLN_t *const *rln, const uint64_t *indices, size_t n, uint8_t *elements_out /* n*depth*32 */,
                              uint8_t *index_bits_out /* n*depth */, RlnString *err);
/* tree build on data already in HBM: d_leaves = count × 32 canonical bytes */
int rlnb200_set_leaves_from_device(FFI_RLN_t **rln, size_t index, const void *d_leaves, size_t count, void *stream, RlnString *err);

/* Variable-base G1 MSM (rln/src/partial_proof.rs:98-104 `msm`, ark-ec msm_bigint): bases n × 64 bytes
 * (x|y canonical LE, bit 0x40 of byte 63 = infinity), scalars n × 32 bytes; result 64 bytes. */
/* ---- one batch over every GPU of the box, inside one process (BASELINE.json configs[4]) -------------------------------------
 * The reference scales by "many callers, one handle" on the host cores (rln/README.md:324-332).  Here one object owns a replica
 * of the prover on each listed device (devices == NULL: every visible device); a batch call cuts the records into contiguous
 * shards, one worker thread per device proves its shard from / to the caller's host buffers, and the call returns when the
 * slowest shard is done.  Tree updates are applied to every replica. */
typedef struct RlnB200Multi RlnB200Multi_t;
RlnB200Multi_t *rlnb200_multi_new(size_t tree_depth, const int *devices, size_t n_devices, RlnString *err);
void rlnb200_multi_free(RlnB200Multi_t *m);
size_t rlnb200_multi_device_count(const RlnB200Multi_t *m);
int rlnb200_multi_device(const RlnB200Multi_t *m, size_t i);
/* borrowed single-device handle of replica i (valid until the next rlnb200_multi_replica call on this thread or
 * rlnb200_multi_free): tree queries, verification, single proofs */
FFI_RLN_t *const *rlnb200_multi_replica(RlnB200Multi_t *m, size_t i);
int rlnb200_multi_set_tree(RlnB200Multi_t *m, size_t tree_depth, RlnString *err);
int rlnb200_multi_set_leaves_from_bytes(RlnB200Multi_t *m, size_t index, const uint8_t *leaves_le, size_t count, RlnString *err);
int rlnb200_multi_atomic_operation(RlnB200Multi_t *m, size_t index, const uint8_t *leaves_le, size_t n_leaves,
                                   const size_t *indices, size_t n_indices, RlnString *err);
int rlnb200_multi_reserve(RlnB200Multi_t *m, size_t max_batch, RlnString *err);
/* records as for rlnb200_prove_batch / rlnb200_verify_batch */
int rlnb200_multi_prove_batch(RlnB200Multi_t *m, const uint8_t *witnesses, size_t n, const uint8_t *rs, uint8_t *proofs_out,
                              RlnString *err);
int rlnb200_multi_verify_batch(RlnB200Multi_t *m, const uint8_t *proofs, size_t n, uint8_t *ok_out, RlnString *err);
/* wall time (ms) each device spent on its shard of the last rlnb200_multi_prove_batch; out has device_count entries */
void rlnb200_multi_last_shard_ms(const RlnB200Multi_t *m, float *out);

typedef struct RlnB200Msm RlnB200Msm_t;
RlnB200Msm_t *rlnb200_msm_new(size_t max_n, RlnString *err);
void rlnb200_msm_free(RlnB200Msm_t *m);
int rlnb200_msm_g1(RlnB200Msm_t *m, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *result, RlnString *err);
/* device-resident variant: d_bases are Montgomery affine points produced by rlnb200_msm_upload_bases /
 * rlnb200_msm_gen_bases; d_scalars n × 32 canonical bytes; d_result 64 bytes */
int rlnb200_msm_upload_bases(RlnB200Msm_t *m, const uint8_t *bases, size_t n, void *d_bases_out, RlnString *err);
int rlnb200_msm_gen_bases(RlnB200Msm_t *m, const void *d_scalars, size_t n, void *d_bases_out, void *stream, RlnString *err);
int rlnb200_msm_g1_device(RlnB200Msm_t *m, const void *d_bases, const void *d_scalars, size_t n, void *d_result, void *stream,
                          RlnString *err);

/* raw kernels for parity tests (host buffers, canonical LE field elements) */
int rlnb200_poseidon_hash(const uint8_t *inputs, int n_inputs /* 1..3 */, uint8_t *out32, RlnString *err);
int rlnb200_hash_pairs(const uint8_t *pairs /* n*64 */, size_t n, uint8_t *out /* n*32 */, RlnString *err);
/* count independent Poseidon hashes of n_inputs (1..3) values each in one launch (e.g. id commitments H(secret), rate
 * commitments H(id_commitment, limit) of a whole membership set: rln/src/protocol/keygen.rs:20-30, rln/README.md) */
int rlnb200_poseidon_hash_batch(const uint8_t *inputs /* count*n_inputs*32 */, int n_inputs, size_t count,
                                uint8_t *out /* count*32 */, RlnString *err);
/* op: 0 mul, 1 add, 2 sub, 3 portable mul, 4 inverse, 5 square, 6-8 single-reduction dot products ; field: 0 Fr, 1 Fq ; exercises the PTX field arithmetic */
int rlnb200_field_op(int field, int op, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out, RlnString *err);
/* batched-affine probe: additions per second of bucket-style accumulation with M running sums per thread in
 * global memory.  mode 0: XYZZ mixed additions (the library's form), 1: affine additions sharing one Fermat inversion per thread
 * and round, 2: the same with a binary extended-Euclid inversion, 3: self-check of that inversion (returns mismatches, 0 = good) */
double rlnb200_affine_batch_probe(int mode, int M, int rounds);
/* cycles per DEPENDENT Fq operation in a lone warp with `lanes` live lanes.  kind 0: the library's product (129 wide MADs on two
 * interleaved carry chains), 1: portable CIOS, 3: dedicated squaring, 4: modular addition */
double rlnb200_latency_probe(int kind, int lanes, int iters);
/* measured Montgomery-product rate (products/s over all SMs, CUDA-event timed); < 0 on error */
double rlnb200_mul_throughput(int iters);
/* same for the cheaper schedules, in product-equivalents per second: kind 1 = dedicated squaring, 2 = two-term dot
 * product with one reduction (counted as two products), 3 = Fq2 product (counted as three) */
double rlnb200_op_throughput(int kind, int iters);
/* pipe probe (thread-instructions per second): mode 0 wide integer MADs, 1 FP64 FMAs, 2 both interleaved */
int rlnb200_pipe_probe(int mode, int iters, double out[2]);
/* witness vector w (num_wires × 32) and quotient h (domain × 32) of one witness record */
int rlnb200_debug_witness_and_h(FFI_RLN_t *const *rln, const uint8_t *witness_le, size_t len, uint8_t *w_out, uint8_t *h_out,
                                RlnString *err);
size_t rlnb200_num_wires(FFI_RLN_t *const *rln);
size_t rlnb200_domain_size(FFI_RLN_t *const *rln);

#ifdef __cplusplus
}
#endif
#endif /* RLN_B200_H */
