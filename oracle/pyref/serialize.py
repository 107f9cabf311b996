"""ORACLE (test infrastructure, never on the product path): big-endian wire formats and the partial-witness record.

Restates rln/src/utils.rs:88-99,110-121,141-226 (fr_to_bytes_be, vec_fr_to_bytes_be, vec_u8_to_bytes_be, vec_bool_to_bytes_be:
32-byte big-endian field elements, 8-byte big-endian length prefixes), rln/src/protocol/witness.rs:418-467 (witness, BE),
:631-676 (partial witness, LE and BE), rln/src/protocol/proof.rs:238-300 (proof values, BE) and :430-446 (proof, BE: the
128-byte Groth16 proof stays in arkworks' little-endian compressed form).  The reference holds no golden bytes for these; its
tests are round trips (rln/tests/serialize.rs), mirrored by tests/test_wire_formats.py.
"""
import struct


def fr_be(v):
    return int(v).to_bytes(32, "big")


def fr_le(v):
    return int(v).to_bytes(32, "little")


def vec_fr(vals, be):
    return struct.pack(">Q" if be else "<Q", len(vals)) + b"".join((fr_be if be else fr_le)(v) for v in vals)


def vec_u8(data, be):
    return struct.pack(">Q" if be else "<Q", len(data)) + bytes(data)


def witness_to_bytes(secret, limit, message_id, path_elements, path_index, x, ext_null, be=False):
    fr = fr_be if be else fr_le
    return (b"\x00" + fr(secret) + fr(limit) + fr(message_id) + vec_fr(path_elements, be) + vec_u8(path_index, be) + fr(x) + fr(ext_null))


def witness_to_bytes_multi(secret, limit, message_ids, path_elements, path_index, x, ext_null, selector_used, be=False):
    fr = fr_be if be else fr_le
    return (b"\x01" + fr(secret) + fr(limit) + vec_fr(path_elements, be) + vec_u8(path_index, be) + fr(x) + fr(ext_null)
            + vec_fr(message_ids, be) + vec_u8([1 if v else 0 for v in selector_used], be))


def partial_witness_to_bytes(secret, limit, path_elements, path_index, be=False):
    fr = fr_be if be else fr_le
    return b"\x00" + fr(secret) + fr(limit) + vec_fr(path_elements, be) + vec_u8(path_index, be)


def proof_values_to_bytes(root, ext_null, x, y, nullifier, be=False):
    fr = fr_be if be else fr_le
    return b"\x00" + fr(root) + fr(ext_null) + fr(x) + fr(y) + fr(nullifier)


def proof_values_to_bytes_multi(root, ext_null, x, ys, nullifiers, selector_used, be=False):
    fr = fr_be if be else fr_le
    return (b"\x01" + fr(root) + fr(ext_null) + fr(x) + vec_fr(ys, be) + vec_fr(nullifiers, be)
            + vec_u8([1 if v else 0 for v in selector_used], be))


def witness_to_bigint_json(secret, limit, message_id, path_elements, path_index, x, ext_null):
    """witness.rs:317-366; serde_json (no preserve_order feature) prints keys sorted"""
    import json
    d = {"identitySecret": str(secret), "userMessageLimit": str(limit), "messageId": str(message_id),
         "pathElements": [str(v) for v in path_elements], "identityPathIndex": [str(v) for v in path_index],
         "x": str(x), "externalNullifier": str(ext_null)}
    return json.dumps(d, sort_keys=True, separators=(",", ":"))


# ---------------------------------------------------------------------------------------------------------------- V3 records
# rln/src/protocol/serialize.rs: LE is ark's derive (enum tag, then the struct fields in declaration order,
# witness.rs:1288-1317, proof.rs:983-1048); BE is hand-written there and orders the Single witness differently (:370-381).
def v3_witness_single(secret, limit, message_id, path_elements, path_index, x, ext_null, be=False):
    fr = fr_be if be else fr_le
    if be:
        return b"\x00" + fr(secret) + fr(limit) + fr(message_id) + vec_fr(path_elements, be) + vec_u8(path_index, be) + fr(x) + fr(ext_null)
    return b"\x00" + fr(secret) + fr(limit) + vec_fr(path_elements, be) + vec_u8(path_index, be) + fr(x) + fr(ext_null) + fr(message_id)


def v3_witness_multi(secret, limit, message_ids, path_elements, path_index, x, ext_null, selector_used, be=False):
    fr = fr_be if be else fr_le
    return (b"\x01" + fr(secret) + fr(limit) + vec_fr(path_elements, be) + vec_u8(path_index, be) + fr(x) + fr(ext_null)
            + vec_fr(message_ids, be) + vec_u8([1 if v else 0 for v in selector_used], be))


def v3_partial_witness(secret, limit, path_elements, path_index, be=False):
    fr = fr_be if be else fr_le
    return fr(secret) + fr(limit) + vec_fr(path_elements, be) + vec_u8(path_index, be)


def v3_values_single(y, root, nullifier, x, ext_null, be=False):
    fr = fr_be if be else fr_le
    return b"\x00" + fr(y) + fr(root) + fr(nullifier) + fr(x) + fr(ext_null)


def v3_values_multi(ys, root, nullifiers, x, ext_null, selector_used, be=False):
    fr = fr_be if be else fr_le
    return (b"\x01" + vec_fr(ys, be) + fr(root) + vec_fr(nullifiers, be) + fr(x) + fr(ext_null)
            + vec_u8([1 if v else 0 for v in selector_used], be))
