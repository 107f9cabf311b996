"""zerokit_b200 — B200-native (sm_100a) drop-in for the proving hot path of vacp2p/zerokit's `rln` crate.

The product is the CUDA shared library zerokit_b200/lib/librln_b200.so (C ABI: include/rln_b200.h);
`zerokit_b200.rln` is a thin ctypes mirror of the reference's public API used by the tests and the
benchmark.  Importing the API without the built library raises — there is no CPU fallback.
"""
from .rln import (RLN, RLNMulti, RLNError, RLNProof, RLNProofValues, RLNWitnessInput, RLNPartialWitnessInput, RLNPartialProof, G1Msm, hash_to_field_le, hash_to_field_be,  # noqa: F401
                  poseidon_hash, poseidon_hash_pair, keygen, seeded_keygen, extended_keygen, extended_seeded_keygen, compute_id_secret,
                  recover_id_secret, vec_fr_to_bytes, bytes_to_vec_fr, vec_u8_to_bytes, bytes_to_vec_u8, proof_values_le_to_be, proof_values_be_to_le, field_op, glv_split, glv_double_mul, hash_pairs, poseidon_hash_batch, set_device, mul_throughput, DEFAULT_TREE_DEPTH, R)
from .rln_v3 import RLNV3, WitnessV3, PartialWitnessV3, ProofV3, ProofValuesV3, PartialProofV3, compute_id_secret_v3  # noqa: F401,E402
