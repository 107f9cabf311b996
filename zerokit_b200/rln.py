"""Python mirror of the reference's public API for the hot path, over the C ABI of librln_b200.so.

Method names, argument meaning and error behaviour follow `rln::public::RLN`
(rln/src/public.rs:65-771): new / new_with_params, set_tree, set_leaf, get_leaf, set_next_leaf,
delete_leaf, leaves_set, set_leaves_from, init_tree_with_leaves, atomic_operation, get_root,
get_merkle_proof, generate_rln_proof, verify_rln_proof, verify_with_roots — so that tests read like
rln/tests/{public,protocol,ffi}.rs.  Field elements are Python ints in [0, r).
All computation happens in the CUDA library; this file only marshals bytes.
"""
import ctypes
from ctypes import POINTER, byref, c_uint8, c_uint32, c_void_p, cast

from . import ffi
from .ffi import CFr, Vec_bool, Vec_CFr, Vec_size, Vec_uint8

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
DEFAULT_TREE_DEPTH = 20  # rln/src/circuit/mod.rs:81


class RLNError(Exception):
    """carries the reference's Display string of the error"""


def _take_string(s):
    if not s.ptr:
        return None
    msg = ctypes.string_at(s.ptr, s.len).decode("utf-8", "replace")
    ffi.lib().ffi_c_string_free(s)
    return msg


def _cfr(v):
    c = CFr()
    ctypes.memmove(c.bytes, (int(v) % (1 << 256)).to_bytes(32, "little"), 32)
    return c


def _cfr_int(c):
    return int.from_bytes(bytes(c.bytes), "little")


def _take_cfr(p):
    v = _cfr_int(p.contents)
    ffi.lib().ffi_cfr_free(p)
    return v


def _vec_cfr(vals):
    n = len(vals)
    arr = (CFr * max(n, 1))()
    for i, v in enumerate(vals):
        ctypes.memmove(arr[i].bytes, (int(v) % (1 << 256)).to_bytes(32, "little"), 32)
    v = Vec_CFr(cast(arr, POINTER(CFr)), n, n)
    v._keep = arr
    return v


def _vec_u8(b):
    b = bytes(b)
    arr = (c_uint8 * max(len(b), 1)).from_buffer_copy(b or b"\0")
    v = Vec_uint8(cast(arr, POINTER(c_uint8)), len(b), len(b))
    v._keep = arr
    return v


def _take_vec_u8(v):
    out = ctypes.string_at(v.ptr, v.len) if v.ptr else b""
    if v.ptr:
        ffi.lib().ffi_vec_u8_free(v)
    return out


def _take_vec_cfr(v):
    out = [_cfr_int(v.ptr[i]) for i in range(v.len)]
    if v.ptr:
        ffi.lib().ffi_vec_cfr_free(v)
    return out


def _ok_bytes(res):
    msg = _take_string(res.err)
    if msg is not None:
        raise RLNError(msg)
    return _take_vec_u8(res.ok)


def _ok_cfr(res):
    msg = _take_string(res.err)
    if msg is not None or not res.ok:
        raise RLNError(msg or "null result")
    return _take_cfr(res.ok)


def _check_bool(res):
    msg = _take_string(res.err)
    if msg is not None:
        raise RLNError(msg)
    return bool(res.ok)


def _check_ptr(res):
    msg = _take_string(res.err)
    if msg is not None or not res.ok:
        raise RLNError(msg or "null result")
    return res.ok


def _check_int(rc, err):
    if rc != 0:
        raise RLNError(_take_string(err) or f"error {rc}")


# ------------------------------------------------------------------------------- free functions
def hash_to_field_le(data: bytes) -> int:
    """rln/src/hashers.rs:73-81"""
    return _take_cfr(ffi.lib().ffi_hash_to_field_le(byref(_vec_u8(data))))


def hash_to_field_be(data: bytes) -> int:
    return _take_cfr(ffi.lib().ffi_hash_to_field_be(byref(_vec_u8(data))))


def poseidon_hash_pair(a: int, b: int) -> int:
    """rln/src/hashers.rs:49-53 (runs on the GPU)"""
    return _take_cfr(ffi.lib().ffi_poseidon_hash_pair(byref(_cfr(a)), byref(_cfr(b))))


def poseidon_hash(inputs) -> int:
    """rln/src/hashers.rs:32-36 for 1..3 inputs"""
    buf = b"".join((int(v) % R).to_bytes(32, "little") for v in inputs)
    out = ctypes.create_string_buffer(32)
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_poseidon_hash(buf, len(inputs), out, byref(err)), err)
    return int.from_bytes(out.raw, "little")


def keygen():
    """rln/src/protocol/keygen.rs:13-21 → (identity_secret, id_commitment)"""
    return tuple(_take_vec_cfr(ffi.lib().ffi_key_gen()))


def seeded_keygen(seed: bytes):
    """keygen.rs:44-58: Keccak-256(seed) → ChaCha20 → Fr::rand"""
    return tuple(_take_vec_cfr(ffi.lib().ffi_seeded_key_gen(byref(_vec_u8(seed)))))


def extended_keygen():
    """keygen.rs:23-38 → (trapdoor, nullifier, identity_secret, id_commitment)"""
    return tuple(_take_vec_cfr(ffi.lib().ffi_extended_key_gen()))


def extended_seeded_keygen(seed: bytes):
    """keygen.rs:64-91"""
    return tuple(_take_vec_cfr(ffi.lib().ffi_seeded_extended_key_gen(byref(_vec_u8(seed)))))


def compute_id_secret(share1, share2) -> int:
    """rln/src/protocol/slashing.rs:7-33: shares are (x, y) pairs"""
    return _ok_cfr(ffi.lib().ffi_compute_id_secret(byref(_cfr(share1[0])), byref(_cfr(share1[1])), byref(_cfr(share2[0])), byref(_cfr(share2[1]))))


def recover_id_secret(proof_values_1_le: bytes, proof_values_2_le: bytes) -> int:
    """slashing.rs:35-100 on two serialized RLNProofValues (rln_proof_values_to_bytes_le)"""
    L = ffi.lib()
    a = c_void_p(_check_ptr(L.ffi_bytes_le_to_rln_proof_values(byref(_vec_u8(proof_values_1_le)))))
    try:
        b = c_void_p(_check_ptr(L.ffi_bytes_le_to_rln_proof_values(byref(_vec_u8(proof_values_2_le)))))
        try:
            return _ok_cfr(L.ffi_recover_id_secret(byref(a), byref(b)))
        finally:
            L.ffi_rln_proof_values_free(b)
    finally:
        L.ffi_rln_proof_values_free(a)


def vec_fr_to_bytes(vals, be=False) -> bytes:
    f = ffi.lib().ffi_vec_cfr_to_bytes_be if be else ffi.lib().ffi_vec_cfr_to_bytes_le
    return _ok_bytes(f(byref(_vec_cfr(vals))))


def bytes_to_vec_fr(data: bytes, be=False):
    f = ffi.lib().ffi_bytes_be_to_vec_cfr if be else ffi.lib().ffi_bytes_le_to_vec_cfr
    res = f(byref(_vec_u8(data)))
    msg = _take_string(res.err)
    if msg is not None:
        raise RLNError(msg)
    return _take_vec_cfr(res.ok)


def vec_u8_to_bytes(data: bytes, be=False) -> bytes:
    f = ffi.lib().ffi_vec_u8_to_bytes_be if be else ffi.lib().ffi_vec_u8_to_bytes_le
    return _ok_bytes(f(byref(_vec_u8(data))))


def bytes_to_vec_u8(data: bytes, be=False) -> bytes:
    f = ffi.lib().ffi_bytes_be_to_vec_u8 if be else ffi.lib().ffi_bytes_le_to_vec_u8
    return _ok_bytes(f(byref(_vec_u8(data))))


def proof_values_le_to_be(data: bytes) -> bytes:
    """bytes_le_to_rln_proof_values → rln_proof_values_to_bytes_be (proof.rs:192-405)"""
    L = ffi.lib()
    pv = c_void_p(_check_ptr(L.ffi_bytes_le_to_rln_proof_values(byref(_vec_u8(data)))))
    try:
        return _take_vec_u8(L.ffi_rln_proof_values_to_bytes_be(byref(pv)))
    finally:
        L.ffi_rln_proof_values_free(pv)


def proof_values_be_to_le(data: bytes) -> bytes:
    L = ffi.lib()
    pv = c_void_p(_check_ptr(L.ffi_bytes_be_to_rln_proof_values(byref(_vec_u8(data)))))
    try:
        return _take_vec_u8(L.ffi_rln_proof_values_to_bytes_le(byref(pv)))
    finally:
        L.ffi_rln_proof_values_free(pv)


# ------------------------------------------------------------------------------- value types
class RLNWitnessInput:
    """rln/src/protocol/witness.rs:52-113 (single message id)"""

    def __init__(self, handle):
        self._h = c_void_p(handle)

    @classmethod
    def new_single(cls, identity_secret, user_message_limit, message_id, path_elements, identity_path_index, x, external_nullifier):
        res = ffi.lib().ffi_rln_witness_input_new_single(
            byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)), byref(_cfr(message_id)),
            byref(_vec_cfr(path_elements)), byref(_vec_u8(bytes(identity_path_index))), byref(_cfr(x)), byref(_cfr(external_nullifier)))
        return cls(_check_ptr(res))

    @classmethod
    def new_multi(cls, identity_secret, user_message_limit, message_ids, path_elements, identity_path_index, x, external_nullifier,
                  selector_used):
        """witness.rs:115-176 (multi message-id mode)"""
        sel = (ctypes.c_bool * max(len(selector_used), 1))(*[bool(v) for v in selector_used])
        vb = Vec_bool(cast(sel, POINTER(ctypes.c_bool)), len(selector_used), len(selector_used))
        res = ffi.lib().ffi_rln_witness_input_new_multi(
            byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)), byref(_vec_cfr(message_ids)), byref(_vec_cfr(path_elements)),
            byref(_vec_u8(bytes(identity_path_index))), byref(_cfr(x)), byref(_cfr(external_nullifier)), byref(vb))
        return cls(_check_ptr(res))

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_witness(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        res = ffi.lib().ffi_rln_witness_to_bytes_le(byref(self._h))
        msg = _take_string(res.err)
        if msg:
            raise RLNError(msg)
        return _take_vec_u8(res.ok)

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_witness(byref(_vec_u8(data)))))

    def to_bytes_be(self):
        return _ok_bytes(ffi.lib().ffi_rln_witness_to_bytes_be(byref(self._h)))

    def to_bigint_json(self) -> str:
        """witness.rs:317-366: the JSON snarkjs / circom witness calculators take"""
        res = ffi.lib().ffi_rln_witness_to_bigint_json(byref(self._h))
        msg = _take_string(res.err)
        if msg is not None:
            raise RLNError(msg)
        return _take_string(res.ok)

    def to_partial(self):
        return RLNPartialWitnessInput(ffi.lib().ffi_rln_witness_to_partial_witness(byref(self._h)))

    # getters (witness.rs:183-246); the variant-specific ones abort on the wrong variant like the reference
    @property
    def version_byte(self):
        return ffi.lib().ffi_rln_witness_input_get_version_byte(byref(self._h))

    @property
    def identity_secret(self):
        return _take_cfr(ffi.lib().ffi_rln_witness_input_get_identity_secret(byref(self._h)))

    @property
    def user_message_limit(self):
        return _take_cfr(ffi.lib().ffi_rln_witness_input_get_user_message_limit(byref(self._h)))

    @property
    def message_id(self):
        return _take_cfr(ffi.lib().ffi_rln_witness_input_get_message_id(byref(self._h)))

    @property
    def message_ids(self):
        return _take_vec_cfr(ffi.lib().ffi_rln_witness_input_get_message_ids(byref(self._h)))

    @property
    def path_elements(self):
        return _take_vec_cfr(ffi.lib().ffi_rln_witness_input_get_path_elements(byref(self._h)))

    @property
    def identity_path_index(self):
        return list(_take_vec_u8(ffi.lib().ffi_rln_witness_input_get_identity_path_index(byref(self._h))))

    @property
    def x(self):
        return _take_cfr(ffi.lib().ffi_rln_witness_input_get_x(byref(self._h)))

    @property
    def external_nullifier(self):
        return _take_cfr(ffi.lib().ffi_rln_witness_input_get_external_nullifier(byref(self._h)))

    @property
    def selector_used(self):
        v = ffi.lib().ffi_rln_witness_input_get_selector_used(byref(self._h))
        out = [bool(v.ptr[i]) for i in range(v.len)]
        ffi.lib().ffi_vec_u8_free(Vec_uint8(cast(v.ptr, POINTER(c_uint8)), v.len, v.cap))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                ffi.lib().ffi_rln_witness_input_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass


class RLNPartialWitnessInput:
    """rln/src/protocol/witness.rs:62-76 — the inputs that do not change between messages"""

    def __init__(self, handle):
        self._h = c_void_p(handle)

    @classmethod
    def new(cls, identity_secret, user_message_limit, path_elements, identity_path_index):
        res = ffi.lib().ffi_rln_partial_witness_input_new(byref(_cfr(identity_secret)), byref(_cfr(user_message_limit)),
                                                          byref(_vec_cfr(path_elements)), byref(_vec_u8(bytes(identity_path_index))))
        return cls(_check_ptr(res))

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_partial_witness(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_partial_witness(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        """witness.rs:631-651: version | secret | limit | vec path_elements | vec identity_path_index"""
        return _ok_bytes(ffi.lib().ffi_rln_partial_witness_to_bytes_le(byref(self._h)))

    def to_bytes_be(self):
        return _ok_bytes(ffi.lib().ffi_rln_partial_witness_to_bytes_be(byref(self._h)))

    @property
    def version_byte(self):
        return ffi.lib().ffi_rln_partial_witness_input_get_version_byte(byref(self._h))

    @property
    def identity_secret(self):
        return _take_cfr(ffi.lib().ffi_rln_partial_witness_input_get_identity_secret(byref(self._h)))

    @property
    def user_message_limit(self):
        return _take_cfr(ffi.lib().ffi_rln_partial_witness_input_get_user_message_limit(byref(self._h)))

    @property
    def path_elements(self):
        return _take_vec_cfr(ffi.lib().ffi_rln_partial_witness_input_get_path_elements(byref(self._h)))

    @property
    def identity_path_index(self):
        return list(_take_vec_u8(ffi.lib().ffi_rln_partial_witness_input_get_identity_path_index(byref(self._h))))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                ffi.lib().ffi_rln_partial_witness_input_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass


class RLNPartialProof:
    """PartialProof (rln/src/partial_proof.rs:30-43) behind FFI_RLNPartialProof"""

    def __init__(self, handle):
        self._h = c_void_p(handle)

    def to_bytes_le(self):
        res = ffi.lib().ffi_rln_partial_proof_to_bytes_le(byref(self._h))
        msg = _take_string(res.err)
        if msg:
            raise RLNError(msg)
        return _take_vec_u8(res.ok)

    def to_bytes_be(self):
        """= the LE form: arkworks points are always little-endian (proof.rs:549-553)"""
        return _ok_bytes(ffi.lib().ffi_rln_partial_proof_to_bytes_be(byref(self._h)))

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_partial_proof(byref(_vec_u8(data)))))

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_partial_proof(byref(_vec_u8(data)))))

    @property
    def version_byte(self):
        return ffi.lib().ffi_rln_partial_proof_get_version_byte(byref(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                ffi.lib().ffi_rln_partial_proof_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass


class RLNProofValues:
    """rln/src/protocol/proof.rs:100-190 — SingleV1 (y, nullifier) or MultiV1 (ys, nullifiers, selector_used)"""

    def __init__(self, root, external_nullifier, x, y=None, nullifier=None, ys=None, nullifiers=None, selector_used=None):
        self.root, self.external_nullifier, self.x, self.y, self.nullifier = root, external_nullifier, x, y, nullifier
        self.ys, self.nullifiers, self.selector_used = ys, nullifiers, selector_used

    def public_inputs(self):
        """circuit order used by the verifier (proof.rs:863-884)"""
        if self.ys is not None:
            return list(self.ys) + [self.root] + list(self.nullifiers) + [self.x, self.external_nullifier] + [int(v) for v in self.selector_used]
        return [self.y, self.root, self.nullifier, self.x, self.external_nullifier]


class RLNProof:
    """rln/src/protocol/proof.rs RLNProof { proof, proof_values } behind FFI_RLNProof"""

    def __init__(self, handle):
        self._h = c_void_p(handle)

    @classmethod
    def from_bytes_le(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_le_to_rln_proof(byref(_vec_u8(data)))))

    def to_bytes_le(self):
        res = ffi.lib().ffi_rln_proof_to_bytes_le(byref(self._h))
        msg = _take_string(res.err)
        if msg:
            raise RLNError(msg)
        return _take_vec_u8(res.ok)

    def to_bytes_be(self):
        res = ffi.lib().ffi_rln_proof_to_bytes_be(byref(self._h))
        return _take_vec_u8(res.ok)

    @classmethod
    def from_bytes_be(cls, data):
        return cls(_check_ptr(ffi.lib().ffi_bytes_be_to_rln_proof(byref(_vec_u8(data)))))

    @property
    def proof_bytes(self):
        """the 128-byte ark-compressed Groth16 proof"""
        return self.to_bytes_le()[1:129]

    @property
    def values(self):
        L = ffi.lib()
        pv = c_void_p(L.ffi_rln_proof_get_values(byref(self._h)))
        try:
            root = _take_cfr(L.ffi_rln_proof_values_get_root(byref(pv)))
            x = _take_cfr(L.ffi_rln_proof_values_get_x(byref(pv)))
            en = _take_cfr(L.ffi_rln_proof_values_get_external_nullifier(byref(pv)))
            if L.ffi_rln_proof_values_get_version_byte(byref(pv)) == 0:
                y = _take_cfr(L.ffi_rln_proof_values_get_y(byref(pv)).ok)
                nul = _take_cfr(L.ffi_rln_proof_values_get_nullifier(byref(pv)).ok)
                return RLNProofValues(root, en, x, y, nul)

            def vec(res):
                out = [_cfr_int(res.ok.ptr[i]) for i in range(res.ok.len)]
                L.ffi_vec_cfr_free(res.ok)
                return out
            ys = vec(L.ffi_rln_proof_values_get_ys(byref(pv)))
            nulls = vec(L.ffi_rln_proof_values_get_nullifiers(byref(pv)))
            sel = [bool(b) for b in _take_vec_u8(L.ffi_rln_proof_values_get_selector_used(byref(pv)).ok)]
            return RLNProofValues(root, en, x, ys=ys, nullifiers=nulls, selector_used=sel)
        finally:
            L.ffi_rln_proof_values_free(pv)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                ffi.lib().ffi_rln_proof_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass


# ------------------------------------------------------------------------------- the RLN object
class RLN:
    """rln::public::RLN (rln/src/public.rs:65-771) backed by the GPU library"""

    def __init__(self, handle):
        self._h = c_void_p(handle)

    @classmethod
    def new(cls, tree_depth=DEFAULT_TREE_DEPTH, config_path=""):
        """public.rs:110-128 via ffi_rln_new (bundled zkey/graph of that depth)"""
        return cls(_check_ptr(ffi.lib().ffi_rln_new(tree_depth, config_path.encode())))

    @classmethod
    def new_multi(cls, tree_depth=DEFAULT_TREE_DEPTH, max_out=4):
        """the bundled multi message-id circuit (rln/src/circuit/mod.rs:36-42)"""
        return cls(_check_ptr(ffi.lib().rlnb200_rln_new_multi(tree_depth, max_out)))

    def max_out(self):
        return ffi.lib().ffi_rln_get_max_out(byref(self._h))

    def witness_record_len(self):
        return ffi.lib().rlnb200_witness_record_len(byref(self._h))

    def proof_record_len(self):
        return ffi.lib().rlnb200_proof_record_len(byref(self._h))

    @classmethod
    def new_with_params(cls, tree_depth, zkey: bytes, graph: bytes, config_path=""):
        """public.rs:184-209"""
        return cls(_check_ptr(ffi.lib().ffi_rln_new_with_params(tree_depth, byref(_vec_u8(zkey)), byref(_vec_u8(graph)), config_path.encode())))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value and not getattr(self, "_borrowed", False):
                ffi.lib().ffi_rln_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    # ---- tree ------------------------------------------------------------------------------
    def tree_depth(self):
        return ffi.lib().ffi_rln_get_tree_depth(byref(self._h))

    def set_tree(self, tree_depth):
        _check_bool(ffi.lib().ffi_set_tree(byref(self._h), tree_depth))

    def set_leaf(self, index, leaf):
        _check_bool(ffi.lib().ffi_set_leaf(byref(self._h), index, byref(_cfr(leaf))))

    def get_leaf(self, index):
        res = ffi.lib().ffi_get_leaf(byref(self._h), index)
        msg = _take_string(res.err)
        if msg:
            raise RLNError(msg)
        return _take_cfr(res.ok)

    def set_next_leaf(self, leaf):
        _check_bool(ffi.lib().ffi_set_next_leaf(byref(self._h), byref(_cfr(leaf))))

    def delete_leaf(self, index):
        _check_bool(ffi.lib().ffi_delete_leaf(byref(self._h), index))

    def leaves_set(self):
        return ffi.lib().ffi_leaves_set(byref(self._h))

    def set_leaves_from(self, index, leaves):
        _check_bool(ffi.lib().ffi_set_leaves_from(byref(self._h), index, byref(_vec_cfr(leaves))))

    def init_tree_with_leaves(self, leaves):
        _check_bool(ffi.lib().ffi_init_tree_with_leaves(byref(self._h), byref(_vec_cfr(leaves))))

    def atomic_operation(self, index, leaves, indices):
        arr = (ctypes.c_size_t * max(len(indices), 1))(*indices)
        vs = Vec_size(cast(arr, POINTER(ctypes.c_size_t)), len(indices), len(indices))
        _check_bool(ffi.lib().ffi_atomic_operation(byref(self._h), index, byref(_vec_cfr(leaves)), byref(vs)))

    def get_subtree_root(self, level: int, index: int) -> int:
        """rln/src/public.rs:877-883: ancestor at `level` (0 = root) of leaf `index`"""
        out = ctypes.create_string_buffer(32)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_get_subtree_root(byref(self._h), level, index, out, byref(err)), err)
        return int.from_bytes(out.raw, "little")

    def get_empty_leaves_indices(self):
        """rln/src/public.rs:885-887"""
        v = Vec_size()
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_get_empty_leaves_indices(byref(self._h), byref(v), byref(err)), err)
        out = [v.ptr[i] for i in range(v.len)]
        ffi.lib().rlnb200_vec_usize_free(v)
        return out

    def set_metadata(self, data: bytes):
        """rln/src/public.rs:499-502"""
        _check_bool(ffi.lib().ffi_set_metadata(byref(self._h), byref(_vec_u8(data))))

    def get_metadata(self) -> bytes:
        return _ok_bytes(ffi.lib().ffi_get_metadata(byref(self._h)))

    def flush(self):
        _check_bool(ffi.lib().ffi_flush(byref(self._h)))

    def get_root(self):
        return _take_cfr(ffi.lib().ffi_get_root(byref(self._h)))

    def get_merkle_proof(self, index):
        """→ (path_elements, identity_path_index), leaf→root, bits LSB-first (public.rs:550-556)"""
        res = ffi.lib().ffi_get_merkle_proof(byref(self._h), index)
        msg = _take_string(res.err)
        if msg:
            raise RLNError(msg)
        mp = res.ok.contents
        elems = [_cfr_int(mp.path_elements.ptr[i]) for i in range(mp.path_elements.len)]
        bits = [mp.path_index.ptr[i] for i in range(mp.path_index.len)]
        ffi.lib().ffi_merkle_proof_free(res.ok)
        return elems, bits

    # bulk extensions
    def set_leaves_from_bytes(self, index, leaves_le: bytes):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_set_leaves_from_bytes(byref(self._h), index, leaves_le, len(leaves_le) // 32, byref(err)), err)

    def set_leaves_from_device(self, index, d_ptr, count, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_set_leaves_from_device(byref(self._h), index, c_void_p(d_ptr), count, c_void_p(stream), byref(err)), err)

    def get_merkle_proofs(self, indices):
        """→ (elements bytes n*depth*32, bits bytes n*depth)"""
        n, d = len(indices), ffi.lib().rlnb200_state_tree_depth(byref(self._h))
        idx = (ctypes.c_uint64 * max(n, 1))(*indices)
        el = ctypes.create_string_buffer(32 * n * d)
        bits = ctypes.create_string_buffer(max(n * d, 1))
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_get_merkle_proofs(byref(self._h), idx, n, el, bits, byref(err)), err)
        return el.raw, bits.raw[:n * d]

    # ---- proving / verifying -----------------------------------------------------------------
    def generate_rln_proof(self, witness: RLNWitnessInput) -> RLNProof:
        """public.rs:624-631"""
        return RLNProof(_check_ptr(ffi.lib().ffi_generate_rln_proof(byref(self._h), byref(witness._h))))

    def generate_rln_proof_with_rs(self, witness: RLNWitnessInput, r: int, s: int) -> RLNProof:
        """generate_zk_proof_with_rs (rln/src/protocol/proof.rs:753-777) behind the ABI"""
        return RLNProof(_check_ptr(ffi.lib().rlnb200_generate_rln_proof_with_rs(byref(self._h), byref(witness._h), byref(_cfr(r)), byref(_cfr(s)))))

    # two-phase proving (public.rs:664-697)
    def generate_rln_proof_with_witness(self, calculated_witness, witness: RLNWitnessInput, r: int = None, s: int = None) -> RLNProof:
        """rln/src/public.rs:643-658: `calculated_witness` is the full wire assignment (ints, negative allowed) computed outside"""
        strs = [str(int(v)).encode() for v in calculated_witness]
        arr = (Vec_uint8 * max(len(strs), 1))()
        keep = []
        for i, b in enumerate(strs):
            buf = (c_uint8 * len(b)).from_buffer_copy(b)
            keep.append(buf)
            arr[i] = Vec_uint8(cast(buf, POINTER(c_uint8)), len(b), len(b))
        vs = ffi.Vec_String(cast(arr, POINTER(Vec_uint8)), len(strs), len(strs))
        if r is None:
            return RLNProof(_check_ptr(ffi.lib().ffi_generate_rln_proof_with_witness(byref(self._h), byref(vs), byref(witness._h))))
        return RLNProof(_check_ptr(ffi.lib().rlnb200_generate_rln_proof_with_witness_rs(byref(self._h), byref(vs), byref(witness._h),
                                                                                         byref(_cfr(r)), byref(_cfr(s)))))

    def generate_partial_zk_proof(self, partial_witness: RLNPartialWitnessInput) -> RLNPartialProof:
        return RLNPartialProof(_check_ptr(ffi.lib().ffi_generate_partial_zk_proof(byref(self._h), byref(partial_witness._h))))

    def finish_rln_proof(self, partial_proof: RLNPartialProof, witness: RLNWitnessInput) -> RLNProof:
        return RLNProof(_check_ptr(ffi.lib().ffi_finish_rln_proof(byref(self._h), byref(partial_proof._h), byref(witness._h))))

    def finish_rln_proof_with_rs(self, partial_proof: RLNPartialProof, witness: RLNWitnessInput, r: int, s: int) -> RLNProof:
        return RLNProof(_check_ptr(ffi.lib().rlnb200_finish_rln_proof_with_rs(byref(self._h), byref(partial_proof._h), byref(witness._h),
                                                                           byref(_cfr(r)), byref(_cfr(s)))))

    def partial_proof_from_bytes_le(self, data: bytes) -> RLNPartialProof:
        return RLNPartialProof(_check_ptr(ffi.lib().rlnb200_bytes_le_to_rln_partial_proof(byref(self._h), byref(_vec_u8(data)))))

    def partial_batch(self, witnesses_le: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(320 * n)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_partial_batch(byref(self._h), witnesses_le, n, out, byref(err)), err)
        return out.raw

    def finish_batch(self, witnesses_le: bytes, n: int, partial: bytes, rs: bytes = None) -> bytes:
        out = ctypes.create_string_buffer(self.proof_record_len() * n)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_finish_batch(byref(self._h), witnesses_le, n, partial, rs, out, byref(err)), err)
        return out.raw

    def verify_rln_proof(self, proof: RLNProof, x: int) -> bool:
        """public.rs:725-745 — raises RLNError with the reference's reason on failure"""
        return _check_bool(ffi.lib().ffi_verify_rln_proof(byref(self._h), byref(proof._h), byref(_cfr(x))))

    def verify_with_roots(self, proof: RLNProof, x: int, roots) -> bool:
        """public.rs:750-771"""
        return _check_bool(ffi.lib().ffi_verify_with_roots(byref(self._h), byref(proof._h), byref(_vec_cfr(roots)), byref(_cfr(x))))

    def prove_batch(self, witnesses_le, n: int, rs=None, out=None):
        """n witness records (rln_witness_to_bytes_le) → n rln_proof_to_bytes_le records (290 bytes each in single mode).
        `witnesses_le`, `rs`, `out` may be bytes-like or integer addresses of host buffers (e.g. pinned tensors); with `out` given
        the records are written there and nothing is returned"""
        buf = out if out is not None else ctypes.create_string_buffer(self.proof_record_len() * n)
        err = ffi.RlnString()
        as_ptr = lambda b: c_void_p(b) if isinstance(b, int) else b   # noqa: E731
        _check_int(ffi.lib().rlnb200_prove_batch(byref(self._h), as_ptr(witnesses_le), n, as_ptr(rs) if rs is not None else None, as_ptr(buf), byref(err)), err)
        return buf.raw if out is None else None

    def verify_batch(self, proofs_le: bytes, n: int):
        ok = ctypes.create_string_buffer(max(n, 1))
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_verify_batch(byref(self._h), proofs_le, n, ok, byref(err)), err)
        return list(ok.raw[:n])

    def prove_records_device(self, d_witness_records, d_rs, n, d_proof_records, stream=0):
        """device-resident wire records: n rln_witness_to_bytes_le records → n rln_proof_to_bytes_le records (d_rs may be 0: fresh r, s)"""
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_prove_records_device(byref(self._h), c_void_p(d_witness_records), c_void_p(d_rs) if d_rs else None, n,
                                                          c_void_p(d_proof_records), c_void_p(stream), byref(err)), err)

    def prove_batch_device(self, d_inputs, d_rs, n, d_proofs, d_values=0, d_affine=0, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_prove_batch_device(byref(self._h), c_void_p(d_inputs), c_void_p(d_rs), n, c_void_p(d_proofs),
                                                        c_void_p(d_values), c_void_p(d_affine), c_void_p(stream), byref(err)), err)

    def partial_batch_device(self, d_inputs, n, d_partial_affine, d_partial_comp, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_partial_batch_device(byref(self._h), c_void_p(d_inputs), n, c_void_p(d_partial_affine),
                                                          c_void_p(d_partial_comp), c_void_p(stream), byref(err)), err)

    def finish_batch_device(self, d_inputs, d_rs, d_partial_affine, n, d_proofs, d_values=0, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_finish_batch_device(byref(self._h), c_void_p(d_inputs), c_void_p(d_rs), c_void_p(d_partial_affine), n,
                                                         c_void_p(d_proofs), c_void_p(d_values), c_void_p(stream), byref(err)), err)

    def witness_to_input_slots(self, witness_le: bytes) -> bytes:
        out = ctypes.create_string_buffer(32 * self.input_slots())
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_witness_to_input_slots(byref(self._h), witness_le, len(witness_le), out, byref(err)), err)
        return out.raw

    def input_slots(self):
        return ffi.lib().rlnb200_input_slots(byref(self._h))

    def input_slot(self, name):
        off, ln = c_uint32(), c_uint32()
        if not ffi.lib().rlnb200_input_slot(byref(self._h), name.encode(), byref(off), byref(ln)):
            raise KeyError(name)
        return off.value, ln.value

    def reserve(self, max_batch):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_reserve(byref(self._h), max_batch, byref(err)), err)

    def last_stage_ms(self):
        out = (ctypes.c_float * 8)()
        ffi.lib().rlnb200_last_stage_ms(byref(self._h), out)
        return dict(zip(("witness", "qap", "msm_g1_accum", "msm_g1_reduce", "msm_g2_accum", "msm_g2_reduce", "assemble", "values"), list(out)))

    def last_stage_batches(self):
        return ffi.lib().rlnb200_last_stage_batches(byref(self._h))

    def table_info(self):
        c, k, c2, k2 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        g1, g2, b = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        ffi.lib().rlnb200_table_info(byref(self._h), byref(c), byref(k), byref(g1), byref(g2), byref(b), byref(c2), byref(k2))
        glv = bool(ffi.lib().rlnb200_glv_enabled(byref(self._h)))
        # adds_per_term: table entries one (scalar, base) term adds up (each GLV half walks all `windows` windows)
        return dict(window_bits=c.value, windows=k.value, window_bits_g2=c2.value, windows_g2=k2.value, glv=glv,
                    adds_per_term=k.value * (2 if glv else 1), g1_bases=g1.value, g2_bases=g2.value, table_bytes=b.value)

    def set_verify_vm_max(self, max_batch: int):
        """batches up to max_batch proofs are verified by the lane-parallel kernel (0: always one thread per proof)"""
        if ffi.lib().rlnb200_set_verify_vm_max(byref(self._h), max_batch) != 0:
            raise RLNError("no verifier program for this key")

    def verify_vm_info(self):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        ffi.lib().rlnb200_verify_vm_info(byref(self._h), byref(a), byref(b), byref(c))
        return dict(levels=a.value, slots=b.value, constants=c.value)

    def verify_vm_trace(self, proof_record: bytes):
        """(ok code, cycles per level, meta per level) of one proof on the lane-parallel verifier"""
        n = self.verify_vm_info()["levels"]
        cyc = (ctypes.c_longlong * (n + 1))()
        meta = (ctypes.c_uint32 * n)()
        ok = ctypes.create_string_buffer(1)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_verify_vm_trace(byref(self._h), proof_record, cyc, meta, ok, byref(err)), err)
        c = list(cyc)
        return ok.raw[0], [c[i + 1] - c[i] for i in range(n)], list(meta)

    def debug_witness_and_h(self, witness_le: bytes):
        nw, dom = ffi.lib().rlnb200_num_wires(byref(self._h)), ffi.lib().rlnb200_domain_size(byref(self._h))
        w = ctypes.create_string_buffer(32 * nw)
        h = ctypes.create_string_buffer(32 * dom)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_debug_witness_and_h(byref(self._h), witness_le, len(witness_le), w, h, byref(err)), err)
        return w.raw, h.raw


class RLNMulti:
    """One batch over several GPUs inside one process (BASELINE.json configs[4]): a replica of the prover per device, contiguous
    shards, one worker thread per device (rlnb200_multi_*).  Where the reference scales by calling one handle from many host
    threads (rln/README.md:324-332), this scales by giving one call to many devices."""

    def __init__(self, tree_depth=DEFAULT_TREE_DEPTH, devices=None):
        err = ffi.RlnString()
        arr = (ctypes.c_int * len(devices))(*devices) if devices else None
        self._m = c_void_p(ffi.lib().rlnb200_multi_new(tree_depth, arr, len(devices) if devices else 0, byref(err)))
        if not self._m.value:
            raise RLNError(_take_string(err) or "multi_new failed")
        self._depth = tree_depth

    def __del__(self):
        try:
            if getattr(self, "_m", None) and self._m.value:
                ffi.lib().rlnb200_multi_free(self._m)
                self._m = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def device_count(self):
        return ffi.lib().rlnb200_multi_device_count(self._m)

    def devices(self):
        return [ffi.lib().rlnb200_multi_device(self._m, i) for i in range(self.device_count())]

    def replica(self, i):
        """a borrowed RLN view of replica i (not owned: freeing it is the multi object's job)"""
        pp = ffi.lib().rlnb200_multi_replica(self._m, i)
        if not pp:
            raise IndexError(i)
        r = RLN.__new__(RLN)
        r._h = c_void_p(pp[0])
        r._borrowed = True
        return r

    def set_tree(self, tree_depth):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_multi_set_tree(self._m, tree_depth, byref(err)), err)

    def set_leaves_from_bytes(self, index, leaves_le: bytes):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_multi_set_leaves_from_bytes(self._m, index, leaves_le, len(leaves_le) // 32, byref(err)), err)

    def atomic_operation(self, index, leaves, indices):
        err = ffi.RlnString()
        lb = b"".join(int(v).to_bytes(32, "little") for v in leaves)
        arr = (ctypes.c_size_t * max(len(indices), 1))(*indices)
        _check_int(ffi.lib().rlnb200_multi_atomic_operation(self._m, index, lb, len(leaves), arr, len(indices), byref(err)), err)

    def reserve(self, max_batch):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_multi_reserve(self._m, max_batch, byref(err)), err)

    def prove_batch(self, witnesses_le, n: int, rs=None, out=None):
        """n witness records → n proof records; `witnesses_le`, `rs`, `out` may be bytes or integer addresses of host buffers"""
        rep = self.replica(0)
        rec = rep.proof_record_len()
        buf = out if out is not None else ctypes.create_string_buffer(rec * n)
        err = ffi.RlnString()
        as_ptr = lambda b: c_void_p(b) if isinstance(b, int) else b   # noqa: E731
        _check_int(ffi.lib().rlnb200_multi_prove_batch(self._m, as_ptr(witnesses_le), n, as_ptr(rs) if rs is not None else None, as_ptr(buf), byref(err)), err)
        return buf.raw if out is None else None

    def verify_batch(self, proofs_le: bytes, n: int):
        ok = ctypes.create_string_buffer(max(n, 1))
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_multi_verify_batch(self._m, proofs_le, n, ok, byref(err)), err)
        return list(ok.raw[:n])

    def last_shard_ms(self):
        out = (ctypes.c_float * self.device_count())()
        ffi.lib().rlnb200_multi_last_shard_ms(self._m, out)
        return list(out)


class G1Msm:
    """variable-base G1 MSM (rln/src/partial_proof.rs:98-104 `msm`)"""

    def __init__(self, max_n):
        err = ffi.RlnString()
        self._h = c_void_p(ffi.lib().rlnb200_msm_new(max_n, byref(err)))
        if not self._h.value:
            raise RLNError(_take_string(err) or "msm_new failed")

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                ffi.lib().rlnb200_msm_free(self._h)
                self._h = c_void_p(None)
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def msm(self, bases: bytes, scalars: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(64)
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_msm_g1(self._h, bases, scalars, n, out, byref(err)), err)
        return out.raw

    def gen_bases(self, d_scalars, n, d_bases_out, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_msm_gen_bases(self._h, c_void_p(d_scalars), n, c_void_p(d_bases_out), c_void_p(stream), byref(err)), err)

    def upload_bases(self, bases: bytes, n, d_bases_out):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_msm_upload_bases(self._h, bases, n, c_void_p(d_bases_out), byref(err)), err)

    def msm_device(self, d_bases, d_scalars, n, d_result, stream=0):
        err = ffi.RlnString()
        _check_int(ffi.lib().rlnb200_msm_g1_device(self._h, c_void_p(d_bases), c_void_p(d_scalars), n, c_void_p(d_result), c_void_p(stream), byref(err)), err)


def set_device(index):
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_set_device(index, byref(err)), err)


def mul_throughput(iters=2000):
    """measured Montgomery products per second (CUDA-event timed)"""
    return ffi.lib().rlnb200_mul_throughput(iters)


def field_op(field, op, a_bytes, b_bytes, n):
    out = ctypes.create_string_buffer(32 * n)
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_field_op(field, op, a_bytes, b_bytes, n, out, byref(err)), err)
    return out.raw


def glv_split(scalars_le: bytes, n):
    """GLV split kernel self-test: returns [(k1, k2)] as signed Python ints"""
    out = ctypes.create_string_buffer(36 * n)
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_glv_split(scalars_le, n, out, byref(err)), err)
    res = []
    for i in range(n):
        r = out.raw[36 * i:36 * i + 36]
        k1, k2 = int.from_bytes(r[:16], "little"), int.from_bytes(r[16:32], "little")
        res.append((-k1 if r[32] else k1, -k2 if r[33] else k2))
    return res


def glv_double_mul(items, use_q=True):
    """self-test of the assembly's Straus / GLV routine: items = [(P, kp, Q, kq)] with affine integer points → [kp·P + kq·Q]"""
    buf = b"".join(b"".join(int(v).to_bytes(32, "little") for v in (P[0], P[1], kp, Q[0], Q[1], kq)) for P, kp, Q, kq in items)
    out = ctypes.create_string_buffer(64 * len(items))
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_glv_double_mul(buf, len(items), 1 if use_q else 0, out, byref(err)), err)
    res = []
    for i in range(len(items)):
        x, y = int.from_bytes(out.raw[64 * i:64 * i + 32], "little"), int.from_bytes(out.raw[64 * i + 32:64 * i + 64], "little")
        res.append(None if x == 0 and y == 0 else (x, y))
    return res


def poseidon_hash_batch(inputs_bytes, n_inputs, count):
    """count independent Poseidon hashes of n_inputs canonical values each (one launch) → count × 32 bytes"""
    out = ctypes.create_string_buffer(32 * count)
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_poseidon_hash_batch(inputs_bytes, n_inputs, count, out, byref(err)), err)
    return out.raw


def hash_pairs(pairs_bytes, n):
    out = ctypes.create_string_buffer(32 * n)
    err = ffi.RlnString()
    _check_int(ffi.lib().rlnb200_hash_pairs(pairs_bytes, n, out, byref(err)), err)
    return out.raw
