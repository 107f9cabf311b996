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


# ----------------------------------------------------------------------------- the default stateful tree (PmTree)
class PmTree:
    """State model of the reference's default `PoseidonTree` = `PmTree` (rln/src/pm_tree_adapter.rs:184-483) over
    vacp2p_pmtree 2.0.3 (un-vendored; `set` / `set_range` / `delete` restated from the published crate: a write at
    [start, start+len) raises next_index to max(next_index, start+len), `set_range` beyond the capacity is
    MerkleTreeIsFull).  Hashing is delegated to FullMerkleTree above (same roots: rln/tests/public.rs:349-427).

    `cached` mirrors `cached_leaves_indices`.  Pinned by the index lists rln/tests/poseidon_tree.rs:79-146 expects
    (tests/test_oracle_goldens.py::test_pmtree_override_range_quirks)."""

    def __init__(self, depth):
        self.depth = depth
        self.tree = FullMerkleTree(depth)
        self.cached = [0] * (1 << depth)

    def capacity(self):
        return 1 << self.depth

    def leaves_set(self):
        return self.tree.next_index

    def root(self):
        return self.tree.root()

    def get(self, index):
        return self.tree.get(index)

    def proof(self, index):
        return self.tree.proof(index)

    def set(self, index, leaf):
        """pm_tree_adapter.rs:262-270"""
        if index >= self.capacity():
            raise IndexError("Index out of bounds")
        self.tree.set(index, leaf)
        self.cached[index] = 1

    def set_range(self, start, values):
        """pm_tree_adapter.rs:272-283"""
        values = list(values)
        if start + len(values) > self.capacity():
            raise IndexError("Merkle Tree is full")
        self.tree.set_range(start, values)
        self.tree.next_index = max(self.tree.next_index, start + len(values))
        for i in range(start, start + len(values)):
            self.cached[i] = 1

    def update_next(self, leaf):
        """pm_tree_adapter.rs:358-363"""
        self.set(self.tree.next_index, leaf)

    def delete(self, index):
        """pm_tree_adapter.rs:365-374: next_index is left alone"""
        keep = self.tree.next_index
        self.tree.set(index, 0)
        self.tree.next_index = keep
        self.cached[index] = 0

    def get_empty_leaves_indices(self):
        """pm_tree_adapter.rs:309-318"""
        return [i for i in range(self.leaves_set()) if self.cached[i] == 0]

    def override_range(self, start, leaves, indices):
        """pm_tree_adapter.rs:320-356 + validate_override_range_inputs (override_range_validation.rs:20-65, policy Allow)
        + remove_indices (:427-445) + remove_indices_and_set_leaves (:447-483)"""
        leaves = list(leaves)
        indices = list(indices)
        if any(i >= self.capacity() for i in indices):
            raise ValueError("Invalid indices")
        indices = sorted(set(indices))
        max_index = None
        if leaves:
            max_index = start + len(leaves)
            if max_index > self.capacity():
                raise ValueError("set_range got too many leaves")
        if indices and max_index is not None and (indices[0] > start or indices[0] >= max_index):
            raise ValueError("Invalid indices")
        if not leaves and not indices:
            raise ValueError("Leaf index out of bounds")
        if len(leaves) == 1 and not indices:
            return self.set(start, leaves[0])
        if not leaves and len(indices) == 1:
            return self.delete(indices[0])
        if not indices:
            return self.set_range(start, leaves)
        if not leaves:  # remove_indices: the whole span is reset
            lo, hi = indices[0], indices[-1] + 1
            self.tree.set_range(lo, [0] * (hi - lo))
            for i in range(lo, hi):
                self.cached[i] = 0
            return None
        # remove_indices_and_set_leaves: set_values covers [min_index, max_index) but is written AT `start`
        min_index = indices[0]
        set_values = [0] * (max_index - min_index)
        for i in range(min_index, start):
            if i not in indices:
                set_values[i - min_index] = self.tree.get(i)
        for i, leaf in enumerate(leaves):
            set_values[start - min_index + i] = leaf
        if start + len(set_values) > self.capacity():
            raise IndexError("Merkle Tree is full")
        self.tree.set_range(start, set_values)
        for i in indices:
            self.cached[i] = 0
        for i in range(start, max_index - min_index):
            self.cached[i] = 1
        return None


class DenseTree:
    """State model of FullMerkleTree / OptimalMerkleTree bookkeeping (utils/src/merkle_tree/full_merkle_tree.rs:197-286,
    optimal_merkle_tree.rs:174-260): set_range raises the flags of everything it writes, override_range needs indices and
    writes its set_values at `start`, delete ignores indices that were never used.  Pinned by the index lists of
    utils/tests/merkle_tree.rs:222-312 (tests/test_oracle_goldens.py::test_dense_tree_override_range)."""

    def __init__(self, depth, optimal=False):
        self.depth = depth
        self.optimal = optimal
        self.tree = FullMerkleTree(depth)
        self.cached = [0] * (1 << depth)

    capacity = PmTree.capacity
    leaves_set = PmTree.leaves_set
    root = PmTree.root
    get = PmTree.get
    proof = PmTree.proof
    get_empty_leaves_indices = PmTree.get_empty_leaves_indices

    def set_range(self, start, values):
        values = list(values)
        if start + len(values) > self.capacity():
            raise ValueError("set_range got too many leaves")
        self.tree.set_range(start, values)
        for i in range(start, start + len(values)):
            self.cached[i] = 1

    def set(self, index, leaf):
        if index >= self.capacity():
            raise IndexError("Leaf index out of bounds")
        self.set_range(index, [leaf])

    def update_next(self, leaf):
        self.set(self.tree.next_index, leaf)

    def delete(self, index):
        if index < self.tree.next_index:
            self.set(index, 0)
            self.cached[index] = 0

    def override_range(self, start, leaves, indices):
        leaves, indices = list(leaves), list(indices)
        if not indices or any(i >= self.capacity() for i in indices):
            raise ValueError("Invalid indices")
        indices = sorted(set(indices))
        min_index = indices[0]
        if leaves:
            max_index = start + len(leaves)
            if max_index > self.capacity():
                raise ValueError("set_range got too many leaves")
            if min_index > start or min_index >= max_index:
                raise ValueError("Invalid indices")
        else:
            max_index = start
        if min_index > max_index or (self.optimal and min_index >= max_index):
            raise ValueError("Invalid indices")
        set_values = [0] * (max_index - min_index)
        for i in range(min_index, start):
            if i not in indices:
                set_values[i - min_index] = self.tree.get(i)
        for i, leaf in enumerate(leaves):
            set_values[start - min_index + i] = leaf
        for i in indices:
            self.cached[i] = 0
        self.set_range(start, set_values)
