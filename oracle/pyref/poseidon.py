"""ORACLE (test infrastructure, not product code).

Poseidon over BN254 Fr with Grain-LFSR generated constants, Keccak-256 hash-to-field and the
dense Poseidon Merkle tree, restated on Python integers.

Follows (reference file:line):
  utils/src/poseidon/poseidon_constants.rs:15-263   Grain LFSR, round constants, Cauchy MDS
  utils/src/poseidon/poseidon_hash.rs:63-135        ark / sbox / mix / hash
  rln/src/hashers.rs:14-23                          ROUND_PARAMS (t, RF, RP, skip)
  rln/src/hashers.rs:73-93                          hash_to_field_{le,be} (Keccak-256, tiny-keccak 2.0.2)
  utils/src/merkle_tree/full_merkle_tree.rs:82-115,197-304,360-399   FullMerkleTree
  rln/src/protocol/witness.rs:807-828               compute_tree_root
"""
from .fields import R

ROUND_PARAMS = [(2, 8, 56, 0), (3, 8, 57, 0), (4, 8, 56, 0), (5, 8, 60, 0),
                (6, 8, 60, 0), (7, 8, 63, 0), (8, 8, 64, 0), (9, 8, 63, 0)]


class GrainLFSR:
    """poseidon_constants.rs:15-205"""

    def __init__(self, is_field, sbox_inv, nbits, t, rf, rp):
        st = [False] * 80
        st[1] = is_field == 1
        st[5] = sbox_inv == 1

        def put(lo, hi, v):
            for i in range(hi, lo - 1, -1):
                st[i] = bool(v & 1)
                v >>= 1

        put(6, 17, nbits)
        put(18, 29, t)
        put(30, 39, rf)
        put(40, 49, rp)
        for i in range(50, 80):
            st[i] = True
        self.st, self.head, self.nbits = st, 0, nbits
        for _ in range(160):
            self.update()

    def update(self):
        s, h = self.st, self.head
        nb = s[(h + 62) % 80] ^ s[(h + 51) % 80] ^ s[(h + 38) % 80] ^ s[(h + 23) % 80] ^ s[(h + 13) % 80] ^ s[h]
        s[h] = nb
        self.head = (h + 1) % 80
        return nb

    def get_int(self):
        """nbits output bits, first generated bit = most significant (poseidon_constants.rs:97-142)."""
        v = 0
        for _ in range(self.nbits):
            b = self.update()
            while not b:
                self.update()
                b = self.update()
            v = (v << 1) | int(self.update())
        return v

    def field_elements_rejection(self, n):
        out = []
        while len(out) < n:
            v = self.get_int()
            if v < R:
                out.append(v)
        return out

    def field_elements_mod_p(self, n):
        return [self.get_int() % R for _ in range(n)]


def find_ark_and_mds(t, rf, rp, skip):
    """poseidon_constants.rs:207-263"""
    l = GrainLFSR(1, 0, 254, t, rf, rp)
    ark = []
    for _ in range(rf + rp):
        ark.extend(l.field_elements_rejection(t))
    for _ in range(skip):
        l.field_elements_mod_p(2 * t)
    xs = l.field_elements_mod_p(t)
    ys = l.field_elements_mod_p(t)
    mds = [[pow((xs[i] + ys[j]) % R, -1, R) for j in range(t)] for i in range(t)]
    return ark, mds


_PARAMS = {}


def params(t):
    if t not in _PARAMS:
        for (tt, rf, rp, skip) in ROUND_PARAMS:
            if tt == t:
                _PARAMS[t] = (rf, rp) + find_ark_and_mds(t, rf, rp, skip)
                break
        else:
            raise ValueError(f"no Poseidon parameters for input length {t - 1}")
    return _PARAMS[t]


def poseidon(inp):
    """poseidon_hash.rs:97-135"""
    if not inp:
        raise ValueError("empty input")
    t = len(inp) + 1
    rf, rp, c, m = params(t)
    st = [0] + [x % R for x in inp]
    for i in range(rf + rp):
        st = [(s + c[i * t + k]) % R for k, s in enumerate(st)]
        if i < rf // 2 or i >= rf // 2 + rp:
            st = [pow(s, 5, R) for s in st]
        else:
            st[0] = pow(st[0], 5, R)
        st = [sum(m[r][k] * st[k] for k in range(t)) % R for r in range(t)]
    return st[0]


# ----------------------------------------------------------------------------- Keccak-256 (NOT sha3)
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
       0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
       0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
       0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    p = bytearray(data)
    p.append(0x01)
    while len(p) % rate:
        p.append(0)
    p[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        blk = p[off:off + rate]
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        a = _keccak_f(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def hash_to_field_le(signal: bytes) -> int:
    """hashers.rs:73-81"""
    return int.from_bytes(keccak256(signal), "little") % R


def hash_to_field_be(signal: bytes) -> int:
    """hashers.rs:84-93 (digest reversed then read big-endian == the LE value)"""
    return int.from_bytes(keccak256(signal)[::-1], "big") % R


# ----------------------------------------------------------------------------- dense Merkle tree
class FullMerkleTree:
    """Heap-ordered dense tree: root at 0, leaf i at 2^depth − 1 + i
    (full_merkle_tree.rs:20-40, 82-115)."""

    def __init__(self, depth, default_leaf=0):
        self.depth = depth
        z = [default_leaf]
        for _ in range(depth):
            z.append(poseidon([z[-1], z[-1]]))
        self.zeros = z  # zeros[k] = root of an empty subtree of height k
        self.nodes = []
        for lvl in range(depth + 1):
            self.nodes.extend([z[depth - lvl]] * (1 << lvl))
        self.next_index = 0

    def root(self):
        return self.nodes[0]

    def _leaf0(self):
        return (1 << self.depth) - 1

    def set_range(self, start, leaves):
        """full_merkle_tree.rs:197-223 + update_hashes :360-399"""
        leaves = list(leaves)
        if not leaves:
            return
        if start + len(leaves) > (1 << self.depth):
            raise IndexError("leaf index out of bounds")
        base = self._leaf0()
        for i, v in enumerate(leaves):
            self.nodes[base + start + i] = v % R
        lo, hi = base + start, base + start + len(leaves) - 1
        while lo > 0:
            lo, hi = (lo - 1) // 2, (hi - 1) // 2
            for p in range(lo, hi + 1):
                self.nodes[p] = poseidon([self.nodes[2 * p + 1], self.nodes[2 * p + 2]])
        self.next_index = max(self.next_index, start + len(leaves))

    def set(self, index, leaf):
        self.set_range(index, [leaf])

    def get(self, index):
        return self.nodes[self._leaf0() + index]

    def proof(self, index):
        """full_merkle_tree.rs:288-304 — siblings leaf→root, index bits LSB-first
        (1 = running node is the right child)."""
        n = self._leaf0() + index
        elems, bits = [], []
        while n > 0:
            sib = n + 1 if n & 1 else n - 1
            elems.append(self.nodes[sib])
            bits.append(0 if n & 1 else 1)
            n = (n - 1) // 2
        return elems, bits


def compute_tree_root(secret, limit, path_elements, path_index):
    """witness.rs:807-828"""
    root = poseidon([poseidon([secret]), limit])
    for e, b in zip(path_elements, path_index):
        root = poseidon([root, e]) if b == 0 else poseidon([e, root])
    return root


def proof_values_from_witness(secret, limit, message_id, path_elements, path_index, x, ext_null):
    """witness.rs:759-804 (single message-id) → dict(root, x, external_nullifier, y, nullifier)"""
    root = compute_tree_root(secret, limit, path_elements, path_index)
    a1 = poseidon([secret, ext_null, message_id])
    y = (secret + x * a1) % R
    nullifier = poseidon([a1])
    return dict(root=root, x=x % R, external_nullifier=ext_null % R, y=y, nullifier=nullifier)


def proof_values_from_witness_multi(secret, limit, message_ids, path_elements, path_index, x, ext_null, selector_used):
    """witness.rs:781-803 (multi message-id): y_i = (secret + x·a1_i)·sel_i, nullifier_i = H(a1_i)·sel_i"""
    root = compute_tree_root(secret, limit, path_elements, path_index)
    ys, nulls = [], []
    for mid, sel in zip(message_ids, selector_used):
        a1 = poseidon([secret, ext_null, mid])
        s = int(bool(sel))
        ys.append((secret + x * a1) % R * s % R)
        nulls.append(poseidon([a1]) * s % R)
    return dict(root=root, x=x % R, external_nullifier=ext_null % R, ys=ys, nullifiers=nulls, selector_used=[bool(v) for v in selector_used])
