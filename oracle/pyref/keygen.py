"""ORACLE (test infrastructure, never on the product path): seeded key generation and Shamir secret recovery.

Restates rln/src/protocol/keygen.rs:13-92 and rln/src/protocol/slashing.rs:7-100.  The randomness sources are third-party
crates absent from /root/reference, restated from their published algorithms and pinned by the reference's own known answers
(rln/tests/protocol.rs:459-507, copied into tests/golden/reference_kats.json "seeded_keygen"):
  * rand_chacha 0.3.1 `ChaCha20Rng::from_seed` (Cargo.lock): ChaCha with 20 rounds, 64-bit block counter in words 12-13,
    stream id 0 in words 14-15; output words consumed in order, `next_u64` = two consecutive little-endian words.
  * ark-ff 0.5.0 `Fp::rand` (UniformRand): draw four u64 limbs, clear the top 64·4 − 254 = 2 bits, reject if ≥ r, and use the
    limbs AS THE MONTGOMERY RESIDUE — so the field element is limbs · 2^−256 mod r.
"""
import struct

from . import poseidon as P
from .fields import R


def _rotl(x, n):
    return ((x << n) & 0xFFFFFFFF) | (x >> (32 - n))


def chacha20_block(key_words, counter):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, 0, 0]
    w = st[:]

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl(w[b] ^ w[c], 7)
    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(w[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaCha20Rng:
    def __init__(self, seed32: bytes):
        self.key = struct.unpack("<8I", seed32)
        self.ctr = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.ctr)
            self.ctr += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def fr_rand(rng) -> int:
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= (1 << 62) - 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < R:
            return v * pow(1 << 256, -1, R) % R


def seeded_keygen(signal: bytes):
    """keygen.rs:44-58 → (identity_secret, id_commitment)"""
    rng = ChaCha20Rng(P.keccak256(signal))
    s = fr_rand(rng)
    return s, P.poseidon([s])


def extended_seeded_keygen(signal: bytes):
    """keygen.rs:64-91 → (trapdoor, nullifier, identity_secret, id_commitment)"""
    rng = ChaCha20Rng(P.keccak256(signal))
    t, n = fr_rand(rng), fr_rand(rng)
    s = P.poseidon([t, n])
    return t, n, s, P.poseidon([s])


def compute_id_secret(share1, share2):
    """slashing.rs:7-33: line through two (x, y) shares, evaluated at 0"""
    (x1, y1), (x2, y2) = share1, share2
    if (x1 - x2) % R == 0:
        raise ValueError("Cannot recover secret: division by zero (shares have the same x value)")
    a1 = (y1 - y2) * pow(x1 - x2, -1, R) % R
    return (y1 - x1 * a1) % R
