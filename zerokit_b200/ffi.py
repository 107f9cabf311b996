"""ctypes declarations for librln_b200.so — one-to-one with include/rln_b200.h.

This is the binding a caller of the reference's C ABI would use; the product
has no Python compute path: if the CUDA library is missing this module raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_bool, c_char_p, c_double, c_float, c_int, c_size_t, c_uint8, c_uint32, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLN_B200_LIB") or os.path.join(_HERE, "lib", "librln_b200.so")   # override: instrumented builds


class Vec_uint8(Structure):
    _fields_ = [("ptr", POINTER(c_uint8)), ("len", c_size_t), ("cap", c_size_t)]


RlnString = Vec_uint8


class Vec_bool(Structure):
    _fields_ = [("ptr", POINTER(c_bool)), ("len", c_size_t), ("cap", c_size_t)]


class Vec_size(Structure):
    _fields_ = [("ptr", POINTER(c_size_t)), ("len", c_size_t), ("cap", c_size_t)]


class Vec_String(Structure):
    _fields_ = [("ptr", POINTER(Vec_uint8)), ("len", c_size_t), ("cap", c_size_t)]


class CFr(Structure):
    _fields_ = [("bytes", c_uint8 * 32)]


class Vec_CFr(Structure):
    _fields_ = [("ptr", POINTER(CFr)), ("len", c_size_t), ("cap", c_size_t)]


class FFI_MerkleProof(Structure):
    _fields_ = [("path_elements", Vec_CFr), ("path_index", Vec_uint8)]


class CBoolResult(Structure):
    _fields_ = [("ok", c_bool), ("err", RlnString)]


def _cresult(name, ok_type):
    return type(name, (Structure,), {"_fields_": [("ok", ok_type), ("err", RlnString)]})


CResult_ptr = _cresult("CResult_ptr", c_void_p)          # any CResult<Box<T>, String>
CResult_MerkleProof = _cresult("CResult_MerkleProof", POINTER(FFI_MerkleProof))
CResult_CFr = _cresult("CResult_CFr", POINTER(CFr))
CResult_Vec_uint8 = _cresult("CResult_Vec_uint8", Vec_uint8)
CResult_Vec_CFr = _cresult("CResult_Vec_CFr", Vec_CFr)
CResult_String = _cresult("CResult_String", RlnString)
CResult_Vec_bool = _cresult("CResult_Vec_bool", Vec_bool)

_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m zerokit_b200.build` (zerokit_b200 has no non-CUDA path)")
    L = ctypes.CDLL(LIB_PATH)
    pp = POINTER(c_void_p)  # T* const* / T**
    sig = {
        "ffi_rln_new": (CResult_ptr, [c_size_t, c_char_p]),
        "ffi_rln_new_with_params": (CResult_ptr, [c_size_t, POINTER(Vec_uint8), POINTER(Vec_uint8), c_char_p]),
        "ffi_rln_free": (None, [c_void_p]),
        "ffi_rln_get_tree_depth": (c_size_t, [pp]),
        "ffi_rln_get_max_out": (c_size_t, [pp]),
        "ffi_set_tree": (CBoolResult, [pp, c_size_t]),
        "ffi_delete_leaf": (CBoolResult, [pp, c_size_t]),
        "ffi_set_leaf": (CBoolResult, [pp, c_size_t, POINTER(CFr)]),
        "ffi_get_leaf": (CResult_CFr, [pp, c_size_t]),
        "ffi_leaves_set": (c_size_t, [pp]),
        "ffi_set_next_leaf": (CBoolResult, [pp, POINTER(CFr)]),
        "ffi_set_leaves_from": (CBoolResult, [pp, c_size_t, POINTER(Vec_CFr)]),
        "ffi_init_tree_with_leaves": (CBoolResult, [pp, POINTER(Vec_CFr)]),
        "ffi_atomic_operation": (CBoolResult, [pp, c_size_t, POINTER(Vec_CFr), POINTER(Vec_size)]),
        "ffi_seq_atomic_operation": (CBoolResult, [pp, POINTER(Vec_CFr), POINTER(Vec_uint8)]),
        "ffi_get_root": (POINTER(CFr), [pp]),
        "ffi_get_merkle_proof": (CResult_MerkleProof, [pp, c_size_t]),
        "ffi_merkle_proof_free": (None, [POINTER(FFI_MerkleProof)]),
        "ffi_rln_witness_input_new_single": (CResult_ptr, [POINTER(CFr)] * 3 + [POINTER(Vec_CFr), POINTER(Vec_uint8), POINTER(CFr), POINTER(CFr)]),
        "ffi_rln_witness_input_new_multi": (CResult_ptr, [POINTER(CFr), POINTER(CFr), POINTER(Vec_CFr), POINTER(Vec_CFr), POINTER(Vec_uint8),
                                                           POINTER(CFr), POINTER(CFr), POINTER(Vec_bool)]),
        "ffi_rln_witness_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_witness_input_free": (None, [c_void_p]),
        "ffi_generate_rln_proof": (CResult_ptr, [pp, pp]),
        "ffi_verify_rln_proof": (CBoolResult, [pp, pp, POINTER(CFr)]),
        "ffi_verify_with_roots": (CBoolResult, [pp, pp, POINTER(Vec_CFr), POINTER(CFr)]),
        "ffi_rln_proof_get_values": (c_void_p, [pp]),
        "ffi_rln_proof_get_version_byte": (c_uint8, [pp]),
        "ffi_rln_proof_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_proof_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_proof_free": (None, [c_void_p]),
        "ffi_rln_proof_values_get_root": (POINTER(CFr), [pp]),
        "ffi_rln_proof_values_get_x": (POINTER(CFr), [pp]),
        "ffi_rln_proof_values_get_external_nullifier": (POINTER(CFr), [pp]),
        "ffi_rln_proof_values_get_y": (CResult_CFr, [pp]),
        "ffi_rln_proof_values_get_nullifier": (CResult_CFr, [pp]),
        "ffi_rln_proof_values_get_ys": (CResult_Vec_CFr, [pp]),
        "ffi_rln_proof_values_get_nullifiers": (CResult_Vec_CFr, [pp]),
        "ffi_rln_proof_values_get_selector_used": (CResult_Vec_uint8, [pp]),
        "ffi_rln_proof_values_get_version_byte": (c_uint8, [pp]),
        "ffi_rln_proof_values_to_bytes_le": (Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_proof_values": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_proof_values_free": (None, [c_void_p]),
        "ffi_cfr_zero": (POINTER(CFr), []),
        "ffi_cfr_one": (POINTER(CFr), []),
        "ffi_cfr_to_bytes_le": (CResult_Vec_uint8, [POINTER(CFr)]),
        "ffi_cfr_to_bytes_be": (CResult_Vec_uint8, [POINTER(CFr)]),
        "ffi_bytes_le_to_cfr": (CResult_CFr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_cfr": (CResult_CFr, [POINTER(Vec_uint8)]),
        "ffi_uint_to_cfr": (POINTER(CFr), [c_uint32]),
        "ffi_cfr_debug": (RlnString, [POINTER(CFr)]),
        "ffi_cfr_free": (None, [POINTER(CFr)]),
        "ffi_vec_cfr_new": (Vec_CFr, [c_size_t]),
        "ffi_vec_cfr_from_cfr": (Vec_CFr, [POINTER(CFr)]),
        "ffi_vec_cfr_push": (None, [POINTER(Vec_CFr), POINTER(CFr)]),
        "ffi_vec_cfr_len": (c_size_t, [POINTER(Vec_CFr)]),
        "ffi_vec_cfr_get": (POINTER(CFr), [POINTER(Vec_CFr), c_size_t]),
        "ffi_vec_cfr_free": (None, [Vec_CFr]),
        "ffi_vec_u8_free": (None, [Vec_uint8]),
        "ffi_c_string_free": (None, [RlnString]),
        "ffi_hash_to_field_le": (POINTER(CFr), [POINTER(Vec_uint8)]),
        "ffi_hash_to_field_be": (POINTER(CFr), [POINTER(Vec_uint8)]),
        "ffi_poseidon_hash_pair": (POINTER(CFr), [POINTER(CFr), POINTER(CFr)]),
        "ffi_key_gen": (Vec_CFr, []),
        "ffi_rln_partial_witness_input_new": (CResult_ptr, [POINTER(CFr), POINTER(CFr), POINTER(Vec_CFr), POINTER(Vec_uint8)]),
        "ffi_rln_partial_witness_input_free": (None, [c_void_p]),
        "ffi_generate_partial_zk_proof": (CResult_ptr, [pp, pp]),
        "ffi_finish_rln_proof": (CResult_ptr, [pp, pp, pp]),
        "ffi_rln_partial_proof_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_partial_proof_free": (None, [c_void_p]),
        "ffi_seeded_key_gen": (Vec_CFr, [POINTER(Vec_uint8)]),
        "ffi_extended_key_gen": (Vec_CFr, []),
        "ffi_seeded_extended_key_gen": (Vec_CFr, [POINTER(Vec_uint8)]),
        "ffi_vec_cfr_to_bytes_le": (CResult_Vec_uint8, [POINTER(Vec_CFr)]),
        "ffi_vec_cfr_to_bytes_be": (CResult_Vec_uint8, [POINTER(Vec_CFr)]),
        "ffi_bytes_le_to_vec_cfr": (CResult_Vec_CFr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_vec_cfr": (CResult_Vec_CFr, [POINTER(Vec_uint8)]),
        "ffi_vec_cfr_debug": (RlnString, [POINTER(Vec_CFr)]),
        "ffi_vec_u8_to_bytes_le": (CResult_Vec_uint8, [POINTER(Vec_uint8)]),
        "ffi_vec_u8_to_bytes_be": (CResult_Vec_uint8, [POINTER(Vec_uint8)]),
        "ffi_bytes_le_to_vec_u8": (CResult_Vec_uint8, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_vec_u8": (CResult_Vec_uint8, [POINTER(Vec_uint8)]),
        "ffi_vec_u8_debug": (RlnString, [POINTER(Vec_uint8)]),
        "ffi_rln_witness_input_get_version_byte": (c_uint8, [pp]),
        "ffi_rln_witness_input_get_identity_secret": (POINTER(CFr), [pp]),
        "ffi_rln_witness_input_get_user_message_limit": (POINTER(CFr), [pp]),
        "ffi_rln_witness_input_get_message_id": (POINTER(CFr), [pp]),
        "ffi_rln_witness_input_get_message_ids": (Vec_CFr, [pp]),
        "ffi_rln_witness_input_get_path_elements": (Vec_CFr, [pp]),
        "ffi_rln_witness_input_get_identity_path_index": (Vec_uint8, [pp]),
        "ffi_rln_witness_input_get_x": (POINTER(CFr), [pp]),
        "ffi_rln_witness_input_get_external_nullifier": (POINTER(CFr), [pp]),
        "ffi_rln_witness_input_get_selector_used": (Vec_bool, [pp]),
        "ffi_rln_witness_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_be_to_rln_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_witness_to_bigint_json": (CResult_String, [pp]),
        "ffi_rln_partial_witness_input_get_version_byte": (c_uint8, [pp]),
        "ffi_rln_partial_witness_input_get_identity_secret": (POINTER(CFr), [pp]),
        "ffi_rln_partial_witness_input_get_user_message_limit": (POINTER(CFr), [pp]),
        "ffi_rln_partial_witness_input_get_path_elements": (Vec_CFr, [pp]),
        "ffi_rln_partial_witness_input_get_identity_path_index": (Vec_uint8, [pp]),
        "ffi_rln_witness_to_partial_witness": (c_void_p, [pp]),
        "ffi_rln_partial_witness_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_partial_witness_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_partial_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_partial_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_proof_values_to_bytes_be": (Vec_uint8, [pp]),
        "ffi_bytes_be_to_rln_proof_values": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_partial_proof_get_version_byte": (c_uint8, [pp]),
        "ffi_rln_partial_proof_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_partial_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_partial_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_compute_id_secret": (CResult_CFr, [POINTER(CFr)] * 4),
        "ffi_recover_id_secret": (CResult_CFr, [pp, pp]),
        "ffi_generate_rln_proof_with_witness": (CResult_ptr, [pp, POINTER(Vec_String), pp]),
        "rlnb200_generate_rln_proof_with_witness_rs": (CResult_ptr, [pp, POINTER(Vec_String), pp, POINTER(CFr), POINTER(CFr)]),
        "ffi_set_metadata": (CBoolResult, [pp, POINTER(Vec_uint8)]),
        "ffi_get_metadata": (CResult_Vec_uint8, [pp]),
        "ffi_flush": (CBoolResult, [pp]),
        # V3 twins (rln/src/ffi/ffi_rln_v3.rs)
        "ffi_rln_v3_new_stateless_default": (c_void_p, []),
        "ffi_rln_v3_new_stateless": (CResult_ptr, [POINTER(Vec_uint8), POINTER(Vec_uint8)]),
        "ffi_rln_v3_new_with_full_merkle_tree_default": (c_void_p, []),
        "ffi_rln_v3_new_with_full_merkle_tree": (CResult_ptr, [c_size_t, POINTER(Vec_uint8), POINTER(Vec_uint8)]),
        "ffi_rln_v3_new_with_optimal_merkle_tree_default": (c_void_p, []),
        "ffi_rln_v3_new_with_optimal_merkle_tree": (CResult_ptr, [c_size_t, POINTER(Vec_uint8), POINTER(Vec_uint8)]),
        "ffi_rln_v3_new_with_pm_tree_default": (c_void_p, []),
        "ffi_rln_v3_new_with_pm_tree": (CResult_ptr, [c_size_t, POINTER(Vec_uint8), POINTER(Vec_uint8), c_char_p]),
        "ffi_rln_v3_free": (None, [c_void_p]),
        "ffi_rln_v3_generate_proof": (CResult_ptr, [pp, pp]),
        "ffi_rln_v3_verify": (CBoolResult, [pp, pp, POINTER(CFr)]),
        "ffi_rln_v3_verify_with_roots": (CBoolResult, [pp, pp, POINTER(Vec_CFr), POINTER(CFr)]),
        "ffi_rln_v3_generate_partial_proof": (CResult_ptr, [pp, pp]),
        "ffi_rln_v3_finish_proof": (CResult_ptr, [pp, pp, pp]),
        "ffi_rln_v3_witness_input_new_single": (CResult_ptr, [POINTER(CFr)] * 3 + [POINTER(Vec_CFr), POINTER(Vec_uint8), POINTER(CFr), POINTER(CFr)]),
        "ffi_rln_v3_witness_input_new_multi": (CResult_ptr, [POINTER(CFr), POINTER(CFr), POINTER(Vec_CFr), POINTER(Vec_CFr), POINTER(Vec_uint8),
                                                              POINTER(CFr), POINTER(CFr), POINTER(Vec_bool)]),
        "ffi_rln_v3_witness_input_get_identity_secret": (POINTER(CFr), [pp]),
        "ffi_rln_v3_witness_input_get_user_message_limit": (POINTER(CFr), [pp]),
        "ffi_rln_v3_witness_input_get_message_id": (CResult_CFr, [pp]),
        "ffi_rln_v3_witness_input_get_message_ids": (CResult_Vec_CFr, [pp]),
        "ffi_rln_v3_witness_input_get_path_elements": (Vec_CFr, [pp]),
        "ffi_rln_v3_witness_input_get_identity_path_index": (Vec_uint8, [pp]),
        "ffi_rln_v3_witness_input_get_x": (POINTER(CFr), [pp]),
        "ffi_rln_v3_witness_input_get_external_nullifier": (POINTER(CFr), [pp]),
        "ffi_rln_v3_witness_input_get_selector_used": (CResult_Vec_bool, [pp]),
        "ffi_rln_v3_witness_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_v3_witness_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_v3_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_v3_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_v3_witness_input_free": (None, [c_void_p]),
        "ffi_rln_v3_partial_witness_input_new": (CResult_ptr, [POINTER(CFr), POINTER(CFr), POINTER(Vec_CFr), POINTER(Vec_uint8)]),
        "ffi_rln_v3_partial_witness_input_get_identity_secret": (POINTER(CFr), [pp]),
        "ffi_rln_v3_partial_witness_input_get_user_message_limit": (POINTER(CFr), [pp]),
        "ffi_rln_v3_partial_witness_input_get_path_elements": (Vec_CFr, [pp]),
        "ffi_rln_v3_partial_witness_input_get_identity_path_index": (Vec_uint8, [pp]),
        "ffi_rln_v3_witness_to_partial_witness": (c_void_p, [pp]),
        "ffi_rln_v3_partial_witness_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_v3_partial_witness_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_v3_partial_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_v3_partial_witness": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_v3_partial_witness_input_free": (None, [c_void_p]),
        "ffi_rln_v3_proof_get_values": (c_void_p, [pp]),
        "ffi_rln_v3_proof_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_v3_proof_to_bytes_mixed": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_v3_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_mixed_to_rln_v3_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_v3_proof_free": (None, [c_void_p]),
        "ffi_rln_v3_partial_proof_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_v3_partial_proof": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_v3_partial_proof_free": (None, [c_void_p]),
        "ffi_rln_v3_proof_values_get_root": (POINTER(CFr), [pp]),
        "ffi_rln_v3_proof_values_get_x": (POINTER(CFr), [pp]),
        "ffi_rln_v3_proof_values_get_external_nullifier": (POINTER(CFr), [pp]),
        "ffi_rln_v3_proof_values_get_y": (CResult_CFr, [pp]),
        "ffi_rln_v3_proof_values_get_nullifier": (CResult_CFr, [pp]),
        "ffi_rln_v3_proof_values_get_selector_used": (CResult_Vec_bool, [pp]),
        "ffi_rln_v3_proof_values_get_ys": (CResult_Vec_CFr, [pp]),
        "ffi_rln_v3_proof_values_get_nullifiers": (CResult_Vec_CFr, [pp]),
        "ffi_rln_v3_proof_values_to_bytes_le": (CResult_Vec_uint8, [pp]),
        "ffi_rln_v3_proof_values_to_bytes_be": (CResult_Vec_uint8, [pp]),
        "ffi_bytes_le_to_rln_v3_proof_values": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_bytes_be_to_rln_v3_proof_values": (CResult_ptr, [POINTER(Vec_uint8)]),
        "ffi_rln_v3_proof_values_free": (None, [c_void_p]),
        "ffi_rln_v3_compute_id_secret": (CResult_CFr, [POINTER(CFr)] * 4),
        "ffi_rln_v3_recover_id_secret": (CResult_CFr, [pp, pp]),
        "ffi_rln_v3_merkle_proof_free": (None, [POINTER(FFI_MerkleProof)]),
        "ffi_rln_v3_delete_leaf": (CBoolResult, [pp, c_size_t]),
        "ffi_rln_v3_set_leaf": (CBoolResult, [pp, c_size_t, POINTER(CFr)]),
        "ffi_rln_v3_get_leaf": (CResult_CFr, [pp, c_size_t]),
        "ffi_rln_v3_leaves_set": (c_size_t, [pp]),
        "ffi_rln_v3_set_next_leaf": (CBoolResult, [pp, POINTER(CFr)]),
        "ffi_rln_v3_set_leaves_from": (CBoolResult, [pp, c_size_t, POINTER(Vec_CFr)]),
        "ffi_rln_v3_init_tree_with_leaves": (CBoolResult, [pp, POINTER(Vec_CFr)]),
        "ffi_rln_v3_atomic_operation": (CBoolResult, [pp, c_size_t, POINTER(Vec_CFr), POINTER(Vec_size)]),
        "ffi_rln_v3_seq_atomic_operation": (CBoolResult, [pp, POINTER(Vec_CFr), POINTER(Vec_uint8)]),
        "ffi_rln_v3_get_root": (POINTER(CFr), [pp]),
        "ffi_rln_v3_get_merkle_proof": (CResult_MerkleProof, [pp, c_size_t]),
        "ffi_rln_v3_set_metadata": (CBoolResult, [pp, POINTER(Vec_uint8)]),
        "ffi_rln_v3_get_metadata": (CResult_Vec_uint8, [pp]),
        "ffi_rln_v3_flush": (CBoolResult, [pp]),
        "rlnb200_v3_generate_proof_with_rs": (CResult_ptr, [pp, pp, POINTER(CFr), POINTER(CFr)]),
        "rlnb200_v3_finish_proof_with_rs": (CResult_ptr, [pp, pp, pp, POINTER(CFr), POINTER(CFr)]),
        # extensions
        "rlnb200_finish_rln_proof_with_rs": (CResult_ptr, [pp, pp, pp, POINTER(CFr), POINTER(CFr)]),
        "rlnb200_bytes_le_to_rln_partial_proof": (CResult_ptr, [pp, POINTER(Vec_uint8)]),
        "rlnb200_partial_batch": (c_int, [pp, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_finish_batch": (c_int, [pp, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_rln_new_multi": (CResult_ptr, [c_size_t, c_size_t]),
        "rlnb200_witness_record_len": (c_size_t, [pp]),
        "rlnb200_proof_record_len": (c_size_t, [pp]),
        "rlnb200_generate_rln_proof_with_rs": (CResult_ptr, [pp, pp, POINTER(CFr), POINTER(CFr)]),
        "rlnb200_prove_batch": (c_int, [pp, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_verify_batch": (c_int, [pp, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_prove_records_device": (c_int, [pp, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_multi_new": (c_void_p, [c_size_t, POINTER(c_int), c_size_t, POINTER(RlnString)]),
        "rlnb200_multi_free": (None, [c_void_p]),
        "rlnb200_multi_device_count": (c_size_t, [c_void_p]),
        "rlnb200_multi_device": (c_int, [c_void_p, c_size_t]),
        "rlnb200_multi_replica": (pp, [c_void_p, c_size_t]),
        "rlnb200_multi_set_tree": (c_int, [c_void_p, c_size_t, POINTER(RlnString)]),
        "rlnb200_multi_set_leaves_from_bytes": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(RlnString)]),
        "rlnb200_multi_atomic_operation": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t), c_size_t, POINTER(RlnString)]),
        "rlnb200_multi_reserve": (c_int, [c_void_p, c_size_t, POINTER(RlnString)]),
        "rlnb200_multi_prove_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_multi_verify_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_multi_last_shard_ms": (None, [c_void_p, POINTER(c_float)]),
        "rlnb200_prove_batch_device": (c_int, [pp, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_partial_batch_device": (c_int, [pp, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_finish_batch_device": (c_int, [pp, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_witness_to_input_slots": (c_int, [pp, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_input_slots": (c_size_t, [pp]),
        "rlnb200_state_tree_depth": (c_size_t, [pp]),
        "rlnb200_input_slot": (c_int, [pp, c_char_p, POINTER(c_uint32), POINTER(c_uint32)]),
        "rlnb200_reserve": (c_int, [pp, c_size_t, POINTER(RlnString)]),
        "rlnb200_launch_count": (c_uint64, []),
        "rlnb200_last_stage_ms": (None, [pp, POINTER(c_float)]),
        "rlnb200_last_stage_batches": (c_uint32, [pp]),
        "rlnb200_set_device": (c_int, [c_int, POINTER(RlnString)]),
        "rlnb200_table_info": (c_int, [pp, POINTER(c_int), POINTER(c_int), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_int), POINTER(c_int)]),
        "rlnb200_set_verify_vm_max": (c_int, [pp, c_size_t]),
        "rlnb200_verify_vm_info": (c_int, [pp, POINTER(c_uint32), POINTER(c_uint32), POINTER(c_uint32)]),
        "rlnb200_verify_vm_trace": (c_int, [pp, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_set_leaves_from_bytes": (c_int, [pp, c_size_t, c_void_p, c_size_t, POINTER(RlnString)]),
        "rlnb200_set_leaves_from_device": (c_int, [pp, c_size_t, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_get_merkle_proofs": (c_int, [pp, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_msm_new": (c_void_p, [c_size_t, POINTER(RlnString)]),
        "rlnb200_msm_free": (None, [c_void_p]),
        "rlnb200_msm_g1": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_msm_upload_bases": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_msm_gen_bases": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_msm_g1_device": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_poseidon_hash": (c_int, [c_void_p, c_int, c_void_p, POINTER(RlnString)]),
        "rlnb200_hash_pairs": (c_int, [c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_poseidon_hash_batch": (c_int, [c_void_p, c_int, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_field_op": (c_int, [c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_debug_witness_and_h": (c_int, [pp, c_void_p, c_size_t, c_void_p, c_void_p, POINTER(RlnString)]),
        "rlnb200_num_wires": (c_size_t, [pp]),
        "rlnb200_domain_size": (c_size_t, [pp]),
        "rlnb200_mul_throughput": (c_double, [c_int]),
        "rlnb200_op_throughput": (c_double, [c_int, c_int]),
        "rlnb200_get_subtree_root": (c_int, [pp, c_size_t, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_get_empty_leaves_indices": (c_int, [pp, POINTER(Vec_size), POINTER(RlnString)]),
        "rlnb200_vec_usize_free": (None, [Vec_size]),
        "rlnb200_glv_enabled": (c_int, [pp]),
        "rlnb200_glv_split": (c_int, [c_void_p, c_size_t, c_void_p, POINTER(RlnString)]),
        "rlnb200_glv_double_mul": (c_int, [c_void_p, c_size_t, c_int, c_void_p, POINTER(RlnString)]),
        "rlnb200_pipe_probe": (c_int, [c_int, c_int, POINTER(c_double)]),
        "rlnb200_affine_batch_probe": (c_double, [c_int, c_int, c_int]),
        "rlnb200_latency_probe": (c_double, [c_int, c_int, c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


# names declared by include/rln_b200.h (kept in sync by tests/test_abi_exports.py)
def declared_symbols():
    hdr = os.path.join(os.path.dirname(_HERE), "include", "rln_b200.h")
    import re
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ffi|rlnb200)_[a-z0-9_]+)\s*\(", txt)))
