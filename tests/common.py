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


def multi_resource(name, depth=20, max_out=4):
    return open(os.path.join(RES, f"tree_depth_{depth}", "multi_message_id", f"max_out_{max_out}", name), "rb").read()


def multi_kat(v):
    """rln/tests/public.rs:143-213 → (proof coordinates A|B|C as 8 integers, public inputs in circuit order
    ys…, root, nullifiers…, x, external_nullifier, selector_used… — rln/src/protocol/proof.rs:870-884)"""
    proof = [int(x) for x in (v["pi_a"] + v["pi_b"][0] + v["pi_b"][1] + v["pi_c"])]
    pub = [int(y) for y in v["ys"]] + [int(v["root"])] + [int(n) for n in v["nullifiers"]] + [int(v["x"]), int(v["external_nullifier"])]
    pub += [int(bool(s)) for s in v["selector_used"]]
    return proof, pub


def pmtree_quirk_script(tree, expected, check=None):
    """rln/tests/poseidon_tree.rs:79-146 step for step on any object with the PoseidonTree methods
    (set_range, delete, set, override_range, get_empty_leaves_indices); `expected` is the golden
    `pmtree_override_range` entry.  `check(tree)` runs after every mutation (the GPU test compares the product's root,
    leaves and leaves_set with the oracle model there)."""
    depth = expected["depth"]
    n = 1 << (depth - 1)
    leaves = list(range(n))
    chk = check or (lambda t: None)
    tree.set_range(0, leaves); chk(tree)
    assert tree.get_empty_leaves_indices() == []
    idxs = []
    for i in range(n):
        idxs.append(i)
        tree.delete(i); chk(tree)
        assert tree.get_empty_leaves_indices() == idxs
    for i in reversed(range(n)):
        idxs.pop()
        tree.set(i, leaves[i]); chk(tree)
        assert tree.get_empty_leaves_indices() == idxs
    assert tree.get_empty_leaves_indices() == []
    l2, l4 = [0, 1], [0, 1, 2, 3]
    tree.override_range(0, l2, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_0_leaves2_idx0123"]
    tree.override_range(0, [], [0, 1]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_0_none_idx01"]
    tree.override_range(0, l2, []); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_0_leaves2_none"]
    tree.override_range(0, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_0_leaves4_idx0123"]
    tree.override_range(4, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_4_leaves4_idx0123"]
    tree.override_range(2, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_2_leaves4_idx0123"]


def dense_tree_quirk_script(tree, expected, check=None):
    """utils/tests/merkle_tree.rs:222-312 (FullMerkleTree / OptimalMerkleTree flavour), step for step"""
    depth = expected["depth"]
    n = 1 << (depth - 1)
    leaves = list(range(n))
    chk = check or (lambda t: None)
    tree.set_range(0, leaves); chk(tree)
    assert tree.get_empty_leaves_indices() == []
    idxs = []
    for i in range(n):
        idxs.append(i)
        tree.delete(i); chk(tree)
        assert tree.get_empty_leaves_indices() == idxs
    for i in reversed(range(n)):
        idxs.pop()
        tree.set(i, leaves[i]); chk(tree)
        assert tree.get_empty_leaves_indices() == idxs
    l2, l4 = [0, 1], [0, 1, 2, 3]
    tree.override_range(0, l2, [0, 1, 2, 3]); chk(tree)
    tree.override_range(0, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_0_leaves4_idx0123"]
    tree.override_range(4, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_4_leaves4_idx0123"]
    tree.override_range(2, l4, [0, 1, 2, 3]); chk(tree)
    assert tree.get_empty_leaves_indices() == expected["after_override_2_leaves4_idx0123"]


class TreeMirror:
    """Runs every tree operation on the product (an RLN / RLNV3 handle) AND on the oracle's state model, and compares root,
    leaves_set, empty indices and the first `watch` leaves after each one — the GPU tests drive the reference's tree tests
    through this."""

    def __init__(self, handle, model, watch=16):
        self.h, self.m, self.watch = handle, model, watch

    def _same(self):
        assert self.h.get_root() == self.m.root()
        assert self.h.leaves_set() == self.m.leaves_set()
        assert self.h.get_empty_leaves_indices() == self.m.get_empty_leaves_indices()
        assert [self.h.get_leaf(i) for i in range(self.watch)] == [self.m.get(i) for i in range(self.watch)]

    def set_range(self, start, leaves):
        self.h.set_leaves_from(start, list(leaves)); self.m.set_range(start, leaves); self._same()

    def set(self, i, leaf):
        self.h.set_leaf(i, leaf); self.m.set(i, leaf); self._same()

    def delete(self, i):
        self.h.delete_leaf(i); self.m.delete(i); self._same()

    def update_next(self, leaf):
        self.h.set_next_leaf(leaf); self.m.update_next(leaf); self._same()

    def override_range(self, start, leaves, indices):
        self.h.atomic_operation(start, list(leaves), list(indices)); self.m.override_range(start, leaves, indices); self._same()

    def get_empty_leaves_indices(self):
        return self.h.get_empty_leaves_indices()


# ---- verifier test inputs: one valid compressed proof and its mutations (used on the host interpreter of the pairing VM and on
# the GPU, where both verifier kernels must return the same codes) ----------------------------------------------------------------
def verifier_mutations(proof128, publics):
    """[(label, proof bytes, public inputs)]: sign bits, neighbouring x coordinates (on / off the curve), flags, a point of the
    twist outside G2, non-canonical coordinates, wrong public inputs"""
    from pyref import fields as F
    from pyref import groth16 as G
    out = [("valid", bytes(proof128), list(publics))]
    for i in (0, 2, 4):
        bad = list(publics)
        bad[i] = (bad[i] + 1) % R
        out.append((f"public[{i}]+1", bytes(proof128), bad))
    big = list(publics)
    big[1] = publics[1] + R          # not reduced: the verifier reduces it (< 2^256)
    if big[1] < 1 << 256:
        out.append(("public[1]+r", bytes(proof128), big))
    for name, off in (("A", 31), ("B", 95), ("C", 127)):
        m = bytearray(proof128)
        m[off] ^= 0x80
        out.append((f"{name} sign flipped", bytes(m), list(publics)))
        m = bytearray(proof128)
        m[off] = (m[off] & 0x3F) | 0x40
        out.append((f"{name} infinity flag", bytes(m), list(publics)))
        m = bytearray(proof128)
        m[off] |= 0xC0
        out.append((f"{name} both flags", bytes(m), list(publics)))
    for name, lo in (("A", 0), ("B.c0", 32), ("B.c1", 64), ("C", 96)):
        for d in range(1, 7):
            m = bytearray(proof128)
            v = int.from_bytes(m[lo:lo + 8], "little") + d
            m[lo:lo + 8] = (v & (2**64 - 1)).to_bytes(8, "little")
            out.append((f"{name}.x+{d}", bytes(m), list(publics)))
    for name, lo in (("A", 0), ("B.c0", 32), ("C", 96)):   # x = q: not canonical
        m = bytearray(proof128)
        top = m[lo + 31] & 0xC0 if lo != 32 else 0
        m[lo:lo + 32] = Q.to_bytes(32, "little")
        m[lo + 31] |= top
        out.append((f"{name}.x = q", bytes(m), list(publics)))
    k, found = 1, 0
    while found < 2:   # points of the twist outside the r-torsion, compressed
        x = (k, 1)
        k += 1
        try:
            pt = G.g2_decompress(G.g2_compress((x, (0, 0)))[:63] + b"\x00")
        except ValueError:
            continue
        if pt is None or F.pt_add(F.OPS2, F.pt_mul(F.OPS2, pt, R - 1), pt) is None:
            continue
        m = bytearray(proof128)
        m[32:96] = G.g2_compress(pt)
        out.append((f"B outside G2 ({k - 1})", bytes(m), list(publics)))
        found += 1
    m = bytearray(proof128)   # B replaced by another point of G2: well-formed, wrong
    m[32:96] = G.g2_compress(F.pt_mul(F.OPS2, F.G2_GEN, 123456789))
    out.append(("B = another G2 point", bytes(m), list(publics)))
    return out


def verifier_expected_code(zkey, proof128, publics):
    """what rln/src/protocol/proof.rs:456-470,856-894 do with these bytes, by the Python oracle: 1 valid, 0 invalid, 2 refused by
    deserialisation (flags, non-canonical, not on the curve, B outside G2); None where a point at infinity is involved"""
    from pyref import fields as F
    from pyref import groth16 as G
    for lo, n in ((0, 32), (32, 64), (96, 32)):
        fl = proof128[lo + n - 1] >> 6
        if fl == 3:
            return 2
        xs = [proof128[lo:lo + 32]] if n == 32 else [proof128[lo:lo + 32], proof128[lo + 32:lo + 64]]
        for i, xb in enumerate(xs):
            v = int.from_bytes(xb, "little")
            if i == len(xs) - 1:
                v &= (1 << 254) - 1
            if v >= Q:
                return 2
    if any(proof128[o] & 0x40 for o in (31, 95, 127)):
        return None
    try:
        proof = G.proof_from_bytes(proof128)
    except ValueError:
        return 2
    if F.pt_add(F.OPS2, F.pt_mul(F.OPS2, proof[1], R - 1), proof[1]) is not None:
        return 2
    return 1 if G.verify(zkey, proof, [p % R for p in publics]) else 0
