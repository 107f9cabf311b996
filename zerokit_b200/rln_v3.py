"""ctypes mirror of the V3 API (rln/src/public.rs:820-997 `RLNV3`, `RLNBuilder`; C ABI rln/src/ffi/ffi_rln_v3.rs) over the same
CUDA prover.  Thin by design: every method is one C-ABI call."""
import ctypes
from ctypes import POINTER, byref, c_uint8, c_void_p, cast

from . import ffi
from .ffi import Vec_bool, Vec_uint8
from .rln import (RLNError, _cfr, _cfr_int, _check_bool, _check_int, _check_ptr, _ok_bytes, _ok_cfr, _take_cfr, _take_string, _take_vec_cfr,
                  _take_vec_u8, _vec_cfr, _vec_u8)


def _take_vec_bool(v):
    out = [bool(v.ptr[i]) for i in range(v.len)]
    if v.ptr:
        ffi.lib().ffi_vec_u8_free(Vec_uint8(cast(v.ptr, POINTER(c_uint8)), v.len, v.cap))
    return out


def _opt(res, take):
    """CResult → value, or None when the getter does not apply to this variant (the message is dropped)"""
    msg = _take_string(res.err)
    return None if msg is not None else take(res.ok)


class _Handle:
    _free = None

    def __init__(self, handle):
        self._h = c_void_p(handle)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                getattr(ffi.lib(), self._free)(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass


class WitnessV3(_Handle):
    """RLNWitnessInputV3 (witness.rs:936-1110)"""
    _free = "ffi_rln_v3_witness_input_free"

    @classmethod
    def new_single(cls, identity_secret, user_message_limit, message_id, path_elements, identity_path_index, x, external_nullifier):
        return cls(_check_ptr(ffi.lib().ffi_rln_v3_witness_input_new_single(
            byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)), byref(_cfr(message_id)), byref(_vec_cfr(path_elements)),
            byref(_vec_u8(bytes(identity_path_index))), byref(_cfr(x)), byref(_cfr(external_nullifier)))))

    @classmethod
    def new_multi(cls, identity_secret, user_message_limit, message_ids, path_elements, identity_path_index, x, external_nullifier, selector_used):
        sel = (ctypes.c_bool * max(len(selector_used), 1))(*[bool(v) for v in selector_used])
        vb = Vec_bool(cast(sel, POINTER(ctypes.c_bool)), len(selector_used), len(selector_used))
        return cls(_check_ptr(ffi.lib().ffi_rln_v3_witness_input_new_multi(
            byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)), byref(_vec_cfr(message_ids)), byref(_vec_cfr(path_elements)),
            byref(_vec_u8(bytes(identity_path_index))), byref(_cfr(x)), byref(_cfr(external_nullifier)), byref(vb))))

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_v3_witness(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_v3_witness(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_witness_to_bytes_le(byref(self._h)))

    def to_bytes_be(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_witness_to_bytes_be(byref(self._h)))

    def to_partial(self):
        return PartialWitnessV3(ffi.lib().ffi_rln_v3_witness_to_partial_witness(byref(self._h)))

    identity_secret = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_witness_input_get_identity_secret(byref(s._h))))
    user_message_limit = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_witness_input_get_user_message_limit(byref(s._h))))
    message_id = property(lambda s: _opt(ffi.lib().ffi_rln_v3_witness_input_get_message_id(byref(s._h)), _take_cfr))
    message_ids = property(lambda s: _opt(ffi.lib().ffi_rln_v3_witness_input_get_message_ids(byref(s._h)), _take_vec_cfr))
    path_elements = property(lambda s: _take_vec_cfr(ffi.lib().ffi_rln_v3_witness_input_get_path_elements(byref(s._h))))
    identity_path_index = property(lambda s: list(_take_vec_u8(ffi.lib().ffi_rln_v3_witness_input_get_identity_path_index(byref(s._h)))))
    x = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_witness_input_get_x(byref(s._h))))
    external_nullifier = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_witness_input_get_external_nullifier(byref(s._h))))
    selector_used = property(lambda s: _opt(ffi.lib().ffi_rln_v3_witness_input_get_selector_used(byref(s._h)), _take_vec_bool))


class PartialWitnessV3(_Handle):
    """RLNPartialWitnessInputV3 (witness.rs:1311-1360)"""
    _free = "ffi_rln_v3_partial_witness_input_free"

    @classmethod
    def new(cls, identity_secret, user_message_limit, path_elements, identity_path_index):
        return cls(_check_ptr(ffi.lib().ffi_rln_v3_partial_witness_input_new(
            byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)), byref(_vec_cfr(path_elements)), byref(_vec_u8(bytes(identity_path_index))))))

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_v3_partial_witness(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_v3_partial_witness(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_partial_witness_to_bytes_le(byref(self._h)))

    def to_bytes_be(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_partial_witness_to_bytes_be(byref(self._h)))

    identity_secret = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_partial_witness_input_get_identity_secret(byref(s._h))))
    user_message_limit = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_partial_witness_input_get_user_message_limit(byref(s._h))))
    path_elements = property(lambda s: _take_vec_cfr(ffi.lib().ffi_rln_v3_partial_witness_input_get_path_elements(byref(s._h))))
    identity_path_index = property(lambda s: list(_take_vec_u8(ffi.lib().ffi_rln_v3_partial_witness_input_get_identity_path_index(byref(s._h)))))


class ProofValuesV3(_Handle):
    """RLNProofValuesV3 (proof.rs:983-1143)"""
    _free = "ffi_rln_v3_proof_values_free"

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_v3_proof_values(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_v3_proof_values(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_proof_values_to_bytes_le(byref(self._h)))

    def to_bytes_be(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_proof_values_to_bytes_be(byref(self._h)))

    def recover_id_secret(self, other) -> int:
        return _ok_cfr(ffi.lib().ffi_rln_v3_recover_id_secret(byref(self._h), byref(other._h)))

    root = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_proof_values_get_root(byref(s._h))))
    x = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_proof_values_get_x(byref(s._h))))
    external_nullifier = property(lambda s: _take_cfr(ffi.lib().ffi_rln_v3_proof_values_get_external_nullifier(byref(s._h))))
    y = property(lambda s: _opt(ffi.lib().ffi_rln_v3_proof_values_get_y(byref(s._h)), _take_cfr))
    nullifier = property(lambda s: _opt(ffi.lib().ffi_rln_v3_proof_values_get_nullifier(byref(s._h)), _take_cfr))
    ys = property(lambda s: _opt(ffi.lib().ffi_rln_v3_proof_values_get_ys(byref(s._h)), _take_vec_cfr))
    nullifiers = property(lambda s: _opt(ffi.lib().ffi_rln_v3_proof_values_get_nullifiers(byref(s._h)), _take_vec_cfr))
    selector_used = property(lambda s: _opt(ffi.lib().ffi_rln_v3_proof_values_get_selector_used(byref(s._h)), _take_vec_bool))


class ProofV3(_Handle):
    """RLNProofV3 { proof, values } (proof.rs:1145-1155)"""
    _free = "ffi_rln_v3_proof_free"

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_v3_proof(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_mixed(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_mixed_to_rln_v3_proof(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_proof_to_bytes_le(byref(self._h)))

    def to_bytes_mixed(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_proof_to_bytes_mixed(byref(self._h)))

    @property
    def values(self):
        return ProofValuesV3(ffi.lib().ffi_rln_v3_proof_get_values(byref(self._h)))


class PartialProofV3(_Handle):
    _free = "ffi_rln_v3_partial_proof_free"

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_v3_partial_proof(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        return _ok_bytes(ffi.lib().ffi_rln_v3_partial_proof_to_bytes_le(byref(self._h)))


class RLNV3(_Handle):
    """RLNV3<Stateless | Stateful<Tree>, ArkGroth16Backend> built by RLNBuilder (public.rs:820-997)"""
    _free = "ffi_rln_v3_free"

    @classmethod
    def stateless(cls, zkey: bytes = None, graph: bytes = None):
        if zkey is None:
            h = ffi.lib().ffi_rln_v3_new_stateless_default()
            if not h:
                raise RLNError("no usable CUDA device")
            return cls(h)
        return cls(_check_ptr(ffi.lib().ffi_rln_v3_new_stateless(byref(_vec_u8(zkey)), byref(_vec_u8(graph)))))

    @classmethod
    def stateful(cls, tree="full", tree_depth=None, zkey: bytes = None, graph: bytes = None, config_path=""):
        L = ffi.lib()
        if zkey is None:
            h = getattr(L, {"full": "ffi_rln_v3_new_with_full_merkle_tree_default", "optimal": "ffi_rln_v3_new_with_optimal_merkle_tree_default",
                            "pm": "ffi_rln_v3_new_with_pm_tree_default"}[tree])()
            if not h:
                raise RLNError("no usable CUDA device")
            return cls(h)
        z, g = _vec_u8(zkey), _vec_u8(graph)
        if tree == "pm":
            return cls(_check_ptr(L.ffi_rln_v3_new_with_pm_tree(tree_depth, byref(z), byref(g), config_path.encode())))
        f = L.ffi_rln_v3_new_with_full_merkle_tree if tree == "full" else L.ffi_rln_v3_new_with_optimal_merkle_tree
        return cls(_check_ptr(f(tree_depth, byref(z), byref(g))))

    # zkSNARK
    def generate_proof(self, witness: WitnessV3) -> ProofV3:
        return ProofV3(_check_ptr(ffi.lib().ffi_rln_v3_generate_proof(byref(self._h), byref(witness._h))))

    def generate_proof_with_rs(self, witness: WitnessV3, r: int, s: int) -> ProofV3:
        return ProofV3(_check_ptr(ffi.lib().rlnb200_v3_generate_proof_with_rs(byref(self._h), byref(witness._h), byref(_cfr(r)), byref(_cfr(s)))))

    def generate_partial_proof(self, pw: PartialWitnessV3) -> PartialProofV3:
        return PartialProofV3(_check_ptr(ffi.lib().ffi_rln_v3_generate_partial_proof(byref(self._h), byref(pw._h))))

    def finish_proof(self, partial: PartialProofV3, witness: WitnessV3) -> ProofV3:
        return ProofV3(_check_ptr(ffi.lib().ffi_rln_v3_finish_proof(byref(self._h), byref(partial._h), byref(witness._h))))

    def finish_proof_with_rs(self, partial: PartialProofV3, witness: WitnessV3, r: int, s: int) -> ProofV3:
        return ProofV3(_check_ptr(ffi.lib().rlnb200_v3_finish_proof_with_rs(byref(self._h), byref(partial._h), byref(witness._h),
                                                                            byref(_cfr(r)), byref(_cfr(s)))))

    def verify(self, proof: ProofV3, x: int) -> bool:
        return _check_bool(ffi.lib().ffi_rln_v3_verify(byref(self._h), byref(proof._h), byref(_cfr(x))))

    def verify_with_roots(self, proof: ProofV3, x: int, roots) -> bool:
        return _check_bool(ffi.lib().ffi_rln_v3_verify_with_roots(byref(self._h), byref(proof._h), byref(_vec_cfr(roots)), byref(_cfr(x))))

    # tree
    def set_leaf(self, index, leaf):
        _check_bool(ffi.lib().ffi_rln_v3_set_leaf(byref(self._h), index, byref(_cfr(leaf))))

    def get_leaf(self, index):
        return _ok_cfr(ffi.lib().ffi_rln_v3_get_leaf(byref(self._h), index))

    def delete_leaf(self, index):
        _check_bool(ffi.lib().ffi_rln_v3_delete_leaf(byref(self._h), index))

    def set_next_leaf(self, leaf):
        _check_bool(ffi.lib().ffi_rln_v3_set_next_leaf(byref(self._h), byref(_cfr(leaf))))

    def leaves_set(self):
        return ffi.lib().ffi_rln_v3_leaves_set(byref(self._h))

    def set_leaves_from(self, index, leaves):
        _check_bool(ffi.lib().ffi_rln_v3_set_leaves_from(byref(self._h), index, byref(_vec_cfr(leaves))))

    def init_tree_with_leaves(self, leaves):
        _check_bool(ffi.lib().ffi_rln_v3_init_tree_with_leaves(byref(self._h), byref(_vec_cfr(leaves))))

    def atomic_operation(self, index, leaves, indices):
        arr = (ctypes.c_size_t * max(len(indices), 1))(*indices)
        vs = ffi.Vec_size(cast(arr, POINTER(ctypes.c_size_t)), len(indices), len(indices))
        _check_bool(ffi.lib().ffi_rln_v3_atomic_operation(byref(self._h), index, byref(_vec_cfr(leaves)), byref(vs)))

    def seq_atomic_operation(self, leaves, indices):
        _check_bool(ffi.lib().ffi_rln_v3_seq_atomic_operation(byref(self._h), byref(_vec_cfr(leaves)), byref(_vec_u8(bytes(indices)))))

    def get_root(self):
        return _take_cfr(ffi.lib().ffi_rln_v3_get_root(byref(self._h)))

    def get_empty_leaves_indices(self):
        """RLNV3::get_empty_leaves_indices (rln/src/public.rs:885-887)"""
        v = ffi.Vec_size()
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_get_empty_leaves_indices(byref(self._h), byref(v), byref(err)), err)
        out = [v.ptr[i] for i in range(v.len)]
        ffi.lib().rlnb200_vec_usize_free(v)
        return out

    def get_merkle_proof(self, index):
        res = ffi.lib().ffi_rln_v3_get_merkle_proof(byref(self._h), index)
        msg = _take_string(res.err)
        if msg is not None:
            raise RLNError(msg)
        mp = res.ok.contents
        out = ([_cfr_int(mp.path_elements.ptr[i]) for i in range(mp.path_elements.len)], [mp.path_index.ptr[i] for i in range(mp.path_index.len)])
        ffi.lib().ffi_rln_v3_merkle_proof_free(res.ok)
        return out

    def set_metadata(self, data: bytes):
        _check_bool(ffi.lib().ffi_rln_v3_set_metadata(byref(self._h), byref(_vec_u8(data))))

    def get_metadata(self) -> bytes:
        return _ok_bytes(ffi.lib().ffi_rln_v3_get_metadata(byref(self._h)))

    def flush(self):
        _check_bool(ffi.lib().ffi_rln_v3_flush(byref(self._h)))


def compute_id_secret_v3(share1, share2) -> int:
    return _ok_cfr(ffi.lib().ffi_rln_v3_compute_id_secret(byref(_cfr(share1[0])), byref(_cfr(share1[1])), byref(_cfr(share2[0])), byref(_cfr(share2[1]))))
