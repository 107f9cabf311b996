"""Shared helpers for the test-suite (test infrastructure: may use the oracle)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "zerokit_b200", "resources")
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def resource(depth, name):
    return open(os.path.join(RES, f"tree_depth_{depth}", name), "rb").read()


def splitmix64(seed):
    s = seed & (2 ** 64 - 1)
    while True:
        s = (s + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        yield z ^ (z >> 31)


def fr_stream(seed):
    """uniform Fr by rejection on 254 bits"""
    g = splitmix64(seed)
    while True:
        v = 0
        for i in range(4):
            v |= next(g) << (64 * i)
        v &= (1 << 254) - 1
        if v < R:
            yield v


def fr_bytes(vals):
    return b"".join((int(v) % (1 << 256)).to_bytes(32, "little") for v in vals)


def ints(buf):
    b = bytes(buf)
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def witness_le(secret, limit, mid, path, idx, x, en):
    """rln_witness_to_bytes_le single layout (rln/src/protocol/witness.rs:369-415)"""
    import struct
    out = b"\x00" + fr_bytes([secret, limit, mid])
    out += struct.pack("<Q", len(path)) + fr_bytes(path) + struct.pack("<Q", len(idx)) + bytes(idx)
    return out + fr_bytes([x, en])


def kat_witness_args(depth, inputs):
    """the witness of the derived known-answer proofs (tests/golden/make_goldens.py:kat_proof)"""
    from pyref import poseidon as P
    pe = [P.poseidon([i + 7]) for i in range(depth)]
    idx = [(5 * i + 1) % 2 for i in range(depth)]
    return (int(inputs["identity_secret"]), int(inputs["user_message_limit"]), int(inputs["message_id"]), pe, idx,
            int(inputs["x"]), int(inputs["external_nullifier"]))
