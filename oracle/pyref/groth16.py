"""ORACLE (test infrastructure, not product code).

RLN Groth16 proving path restated on Python integers: arkzkey + graph.bin parsers, the
witnesscalc graph VM, the snarkjs-compatible QAP witness map, proof assembly, the
verifier and the wire formats.

Follows (reference file:line):
  rln/src/circuit/mod.rs:256-305                      arkzkey layout (uncompressed, unchecked)
  rln/src/circuit/iden3calc/storage.rs:16-86,265-302  graph.bin container
  rln/src/circuit/iden3calc/proto.rs:7-117            protobuf messages
  rln/src/circuit/iden3calc/graph.rs:71-143,246-272,314-466   node evaluation
  rln/src/circuit/iden3calc.rs:20-60,106-181          input buffer population
  rln/src/protocol/witness.rs:832-881                 named inputs
  rln/src/circuit/qap.rs:30-98                        CircomReduction::witness_map_from_matrices
  rln/src/partial_proof.rs:182-274                    proof assembly (== ark-groth16 0.5.0)
  rln/src/protocol/proof.rs:856-894                   verify_zk_proof public-input order
  rln/src/protocol/proof.rs:192-236,413-428           proof / proof-values LE bytes
  rln/src/protocol/witness.rs:369-415                 witness LE bytes
Un-vendored: ark-groth16/ark-poly/ark-serialize 0.5.0 (Cargo.lock:172,188,233) — restated from
their published algorithms; pinned by the verifier KAT (rln/tests/public.rs:77-233) and by
prove→verify round trips.
"""
import struct

from . import fields as F
from .fields import (R, Q, INF, OPS1, OPS2, pt_add, pt_mul, pt_neg, msm, on_curve,
                     pairing_product_is_one, TWO_ADICITY, FR_GENERATOR)


# ----------------------------------------------------------------------------- arkzkey
class Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def fe(self):
        v = int.from_bytes(self.b[self.o:self.o + 32], "little")
        self.o += 32
        return v

    def g1(self):
        x = self.fe()
        raw_y = self.b[self.o:self.o + 32]
        self.o += 32
        if raw_y[31] & 0x40:
            return INF
        # top two bits of the last byte are ark-serialize SWFlags (bit7 sign, bit6 infinity)
        return (x, int.from_bytes(raw_y, "little") & ((1 << 254) - 1))

    def g2(self):
        x0, x1, y0 = self.fe(), self.fe(), self.fe()
        raw = self.b[self.o:self.o + 32]
        self.o += 32
        if raw[31] & 0x40:
            return INF
        return ((x0, x1), (y0, int.from_bytes(raw, "little") & ((1 << 254) - 1)))

    def vec(self, f):
        return [f() for _ in range(self.u64())]


class Zkey:
    pass


def parse_zkey(data: bytes) -> Zkey:
    """circuit/mod.rs:256-305: (ProvingKey, matrices) via ark-serialize uncompressed/unchecked."""
    r = Reader(data)
    z = Zkey()
    z.alpha_g1 = r.g1()
    z.beta_g2 = r.g2()
    z.gamma_g2 = r.g2()
    z.delta_g2 = r.g2()
    z.gamma_abc_g1 = r.vec(r.g1)
    z.beta_g1 = r.g1()
    z.delta_g1 = r.g1()
    z.a_query = r.vec(r.g1)
    z.b_g1_query = r.vec(r.g1)
    z.b_g2_query = r.vec(r.g2)
    z.h_query = r.vec(r.g1)
    z.l_query = r.vec(r.g1)
    z.num_instance = r.u64()
    z.num_witness = r.u64()
    z.num_constraints = r.u64()
    z.a_nnz, z.b_nnz, z.c_nnz = r.u64(), r.u64(), r.u64()

    def matrix():
        rows = []
        for _ in range(r.u64()):
            k = r.u64()
            row = []
            for _ in range(k):
                c = r.fe()
                row.append((c, r.u64()))
            rows.append(row)
        return rows

    z.a, z.b, z.c = matrix(), matrix(), matrix()
    assert r.o == len(data), (r.o, len(data))
    return z


# ----------------------------------------------------------------------------- graph.bin
def _varint(b, o):
    v = s = 0
    while True:
        c = b[o]
        o += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return v, o


def _pb_fields(b):
    """minimal protobuf wire decoder → list of (tag, wiretype, value)"""
    o, out = 0, []
    while o < len(b):
        key, o = _varint(b, o)
        tag, wt = key >> 3, key & 7
        if wt == 0:
            v, o = _varint(b, o)
        elif wt == 2:
            ln, o = _varint(b, o)
            v = b[o:o + ln]
            o += ln
        elif wt == 5:
            v = b[o:o + 4]
            o += 4
        elif wt == 1:
            v = b[o:o + 8]
            o += 8
        else:
            raise ValueError("unsupported wire type")
        out.append((tag, wt, v))
    return out


OP_NAMES = ["Mul", "Div", "Add", "Sub", "Pow", "Idiv", "Mod", "Eq", "Neq", "Lt", "Gt", "Leq", "Geq",
            "Land", "Lor", "Shl", "Shr", "Bor", "Band", "Bxor"]


class Graph:
    pass


def parse_graph(data: bytes) -> Graph:
    """storage.rs:265-302. nodes: ('input',i) | ('const',v) | ('uno',op,a) | ('duo',op,a,b) | ('tres',op,a,b,c)"""
    magic = b"wtns.graph.001"
    assert data[:len(magic)] == magic, "Invalid magic"
    o = len(magic)
    n = struct.unpack_from("<Q", data, o)[0]
    o += 8
    nodes = []
    for _ in range(n):
        ln, o = _varint(data, o)
        msg = data[o:o + ln]
        o += ln
        (tag, _, body), = _pb_fields(msg)
        f = {t: v for t, _, v in _pb_fields(body)}
        if tag == 1:
            nodes.append(("input", f.get(1, 0)))
        elif tag == 2:
            inner = {t: v for t, _, v in _pb_fields(f[1])}
            nodes.append(("const", int.from_bytes(inner.get(1, b""), "little") % R))
        elif tag == 3:
            nodes.append(("uno", f.get(1, 0), f.get(2, 0)))
        elif tag == 4:
            nodes.append(("duo", f.get(1, 0), f.get(2, 0), f.get(3, 0)))
        elif tag == 5:
            nodes.append(("tres", f.get(1, 0), f.get(2, 0), f.get(3, 0), f.get(4, 0)))
        else:
            raise ValueError("bad node tag")
    ln, o = _varint(data, o)
    md = data[o:o + ln]
    g = Graph()
    g.nodes = nodes
    g.signals = []
    g.inputs = {}
    for tag, wt, v in _pb_fields(md):
        if tag == 1:
            if wt == 2:  # packed
                p = 0
                while p < len(v):
                    x, p = _varint(v, p)
                    g.signals.append(x)
            else:
                g.signals.append(v)
        elif tag == 2:
            e = {t: vv for t, _, vv in _pb_fields(v)}
            sd = {t: vv for t, _, vv in _pb_fields(e.get(2, b""))}
            g.inputs[e[1].decode()] = (sd.get(1, 0), sd.get(2, 0))
    g.tree_depth = g.inputs["pathElements"][1]
    return g


HALF = R // 2


def _duo(op, a, b):
    """graph.rs:71-143 + helpers :314-466"""
    name = OP_NAMES[op]
    if name == "Mul":
        return a * b % R
    if name == "Add":
        return (a + b) % R
    if name == "Sub":
        return (a - b) % R
    if name == "Div":
        return 0 if b == 0 else a * pow(b, -1, R) % R
    if name == "Pow":
        return pow(a, b, R)
    if name == "Idiv":
        return 0 if b == 0 else a // b
    if name == "Mod":
        return 0 if b == 0 else a % b
    if name == "Eq":
        return int(a == b)
    if name == "Neq":
        return int(a != b)
    if name in ("Lt", "Gt", "Leq", "Geq"):
        an, bn = a > HALF, b > HALF
        if an != bn:
            # exactly one is "negative"
            lt = an  # negative < non-negative
            return int({"Lt": lt, "Leq": lt, "Gt": not lt, "Geq": not lt}[name])
        return int({"Lt": a < b, "Leq": a <= b, "Gt": a > b, "Geq": a >= b}[name])
    if name == "Land":
        return int(a != 0 and b != 0)
    if name == "Lor":
        return int(a != 0 or b != 0)
    if name == "Shl":
        if b == 0:
            return a
        if b >= 254:
            return 0
        v = (a << b) & ((1 << 256) - 1)
        if v >= R:
            raise ValueError("Failed to compute left shift")
        return v
    if name == "Shr":
        if b == 0:
            return a
        if b >= 254:
            return 0
        return a >> (b & 0xFF)
    if name in ("Bor", "Band", "Bxor"):
        d = {"Bor": a | b, "Band": a & b, "Bxor": a ^ b}[name]
        if d > R:
            d -= R
        if d >= R:
            raise ValueError("bit op result not in field")
        return d
    raise ValueError(name)


def inputs_buffer(g, secret, limit, message_id, path_elements, path_index, x, ext_null, selector_used=None):
    """iden3calc.rs:106-181 + witness.rs:832-881.  Single mode: message_id is one value; multi mode (graph has a
    `selectorUsed` input): message_id is the list of message ids and selector_used the list of bools."""
    size = 0
    started = False
    for nd in g.nodes:
        if nd[0] == "input":
            size = max(size, nd[1])
            started = True
        elif started:
            break
    buf = [0] * (size + 1)
    buf[0] = 1
    named = {
        "identitySecret": [secret], "userMessageLimit": [limit],
        "messageId": list(message_id) if selector_used is not None else [message_id],
        "pathElements": list(path_elements), "identityPathIndex": list(path_index),
        "x": [x], "externalNullifier": [ext_null],
    }
    if selector_used is not None:
        named["selectorUsed"] = [int(bool(v)) for v in selector_used]
    for k, vals in named.items():
        off, ln = g.inputs[k]
        if ln != len(vals):
            raise ValueError(f"input {k}: expected {ln} got {len(vals)}")
        for i, v in enumerate(vals):
            buf[off + i] = v % R
    return buf


def evaluate(g, buf):
    """graph.rs:246-272"""
    vals = []
    for nd in g.nodes:
        k = nd[0]
        if k == "const":
            v = nd[1]
        elif k == "input":
            v = buf[nd[1]]
        elif k == "duo":
            v = _duo(nd[1], vals[nd[2]], vals[nd[3]])
        elif k == "uno":
            if nd[1] != 0:
                raise ValueError("uno operator Id not implemented for Montgomery")
            v = (-vals[nd[2]]) % R
        else:
            v = vals[nd[3]] if vals[nd[2]] != 0 else vals[nd[4]]
        vals.append(v)
    return [vals[i] for i in g.signals]


# ----------------------------------------------------------------------------- QAP (qap.rs:30-98)
def root_of_unity(n):
    lg = n.bit_length() - 1
    assert 1 << lg == n
    return pow(pow(FR_GENERATOR, (R - 1) >> TWO_ADICITY, R), 1 << (TWO_ADICITY - lg), R)


def ntt(v, w):
    n = len(v)
    lg = n.bit_length() - 1
    a = [0] * n
    for i, x in enumerate(v):
        a[int(format(i, f"0{lg}b")[::-1], 2)] = x
    m = 1
    while m < n:
        wm = pow(w, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                u, x = a[k + j], a[k + j + m] * t % R
                a[k + j], a[k + j + m] = (u + x) % R, (u - x) % R
                t = t * wm % R
        m *= 2
    return a


def intt(v, w):
    n = len(v)
    ni = pow(n, -1, R)
    return [x * ni % R for x in ntt(v, pow(w, -1, R))]


def witness_map(z, w):
    n = 1
    while n < z.num_constraints + z.num_instance:
        n *= 2
    a = [0] * n
    b = [0] * n
    for i in range(z.num_constraints):
        a[i] = sum(c * w[j] for c, j in z.a[i]) % R
        b[i] = sum(c * w[j] for c, j in z.b[i]) % R
    for i in range(z.num_instance):
        a[z.num_constraints + i] = w[i]
    c = [x * y % R for x, y in zip(a, b)]
    om = root_of_unity(n)
    g = root_of_unity(2 * n)

    def to_coset(v):
        co = intt(v, om)
        t = 1
        for i in range(n):
            co[i] = co[i] * t % R
            t = t * g % R
        return ntt(co, om)

    a, b, c = to_coset(a), to_coset(b), to_coset(c)
    return [(x * y - cc) % R for x, y, cc in zip(a, b, c)]


# ----------------------------------------------------------------------------- prove / verify
def prove(z, w, h, r, s):
    """partial_proof.rs:182-274 with an empty partial proof (== ark-groth16 create_proof_with_assignment)."""
    o1, o2 = OPS1, OPS2
    ni = z.num_instance
    a_acc = msm(o1, z.a_query[1:], w[1:])
    g_a = pt_add(o1, pt_add(o1, pt_add(o1, z.alpha_g1, z.a_query[0]), a_acc), pt_mul(o1, z.delta_g1, r))
    if r % R:
        b1_acc = msm(o1, z.b_g1_query[1:], w[1:])
        g1_b = pt_add(o1, pt_add(o1, pt_add(o1, z.beta_g1, z.b_g1_query[0]), b1_acc), pt_mul(o1, z.delta_g1, s))
    else:
        g1_b = INF
    b2_acc = msm(o2, z.b_g2_query[1:], w[1:])
    g2_b = pt_add(o2, pt_add(o2, pt_add(o2, z.beta_g2, z.b_g2_query[0]), b2_acc), pt_mul(o2, z.delta_g2, s))
    l_acc = msm(o1, z.l_query, w[ni:])
    h_acc = msm(o1, z.h_query, h)
    g_c = pt_mul(o1, g_a, s)
    g_c = pt_add(o1, g_c, pt_mul(o1, g1_b, r))
    g_c = pt_add(o1, g_c, pt_neg(o1, pt_mul(o1, z.delta_g1, r * s % R)))
    g_c = pt_add(o1, g_c, l_acc)
    g_c = pt_add(o1, g_c, h_acc)
    return (g_a, g2_b, g_c)


# ----------------------------------------------------------------------------- partial proofs
UNKNOWN_INPUTS_SINGLE = ("messageId", "x", "externalNullifier")   # witness.rs:887-931 (None entries)


def known_wire_mask(g):
    """graph.rs:274-312 evaluate_partial: a node is known iff all of its operands are; inputs are known unless
    they belong to messageId / x / externalNullifier.  Returns one bool per witness signal (wire)."""
    unknown_slots = set()
    for name in UNKNOWN_INPUTS_SINGLE:
        off, ln = g.inputs[name]
        unknown_slots.update(range(off, off + ln))
    known = []
    for nd in g.nodes:
        k = nd[0]
        if k == "const":
            v = True
        elif k == "input":
            v = nd[1] not in unknown_slots
        elif k == "uno":
            v = known[nd[2]]
        elif k == "duo":
            v = known[nd[2]] and known[nd[3]]
        else:
            v = known[nd[2]] and known[nd[3]] and known[nd[4]]
        known.append(v)
    return [known[i] for i in g.signals]


def prove_partial(z, g, w_known):
    """partial_proof.rs:108-179.  w_known: full-length wire vector whose known entries are correct (unknown
    entries are ignored).  Returns (mask over w[1:], partial_pi_a, partial_rho, partial_pi_b, partial_pi_c)."""
    o1, o2 = OPS1, OPS2
    wire_known = known_wire_mask(g)
    assert wire_known[0]
    mask = wire_known[1:]
    ni = z.num_instance
    idx = [i for i in range(1, len(w_known)) if wire_known[i]]
    sc = [w_known[i] for i in idx]
    a = msm(o1, [z.a_query[i] for i in idx], sc)
    b1 = msm(o1, [z.b_g1_query[i] for i in idx], sc)
    b2 = msm(o2, [z.b_g2_query[i] for i in idx], sc)
    lidx = [i for i in range(ni, len(w_known)) if wire_known[i]]
    l = msm(o1, [z.l_query[i - ni] for i in lidx], [w_known[i] for i in lidx])
    pi_a = pt_add(o1, pt_add(o1, z.alpha_g1, z.a_query[0]), a)
    rho = pt_add(o1, pt_add(o1, z.beta_g1, z.b_g1_query[0]), b1)
    pi_b = pt_add(o2, pt_add(o2, z.beta_g2, z.b_g2_query[0]), b2)
    return mask, pi_a, rho, pi_b, l


def finish_partial(z, partial, w, h, r, s):
    """partial_proof.rs:182-274"""
    o1, o2 = OPS1, OPS2
    mask, pi_a, rho, pi_b, pi_c = partial
    ni = z.num_instance
    idx = [i for i in range(1, len(w)) if not mask[i - 1]]
    sc = [w[i] for i in idx]
    g_a = pt_add(o1, pt_add(o1, pi_a, msm(o1, [z.a_query[i] for i in idx], sc)), pt_mul(o1, z.delta_g1, r))
    if r % R:
        g1_b = pt_add(o1, pt_add(o1, rho, msm(o1, [z.b_g1_query[i] for i in idx], sc)), pt_mul(o1, z.delta_g1, s))
    else:
        g1_b = INF
    g2_b = pt_add(o2, pt_add(o2, pi_b, msm(o2, [z.b_g2_query[i] for i in idx], sc)), pt_mul(o2, z.delta_g2, s))
    lidx = [i for i in range(ni, len(w)) if not mask[i - 1]]
    l_acc = pt_add(o1, pi_c, msm(o1, [z.l_query[i - ni] for i in lidx], [w[i] for i in lidx]))
    h_acc = msm(o1, z.h_query, h)
    g_c = pt_mul(o1, g_a, s)
    g_c = pt_add(o1, g_c, pt_mul(o1, g1_b, r))
    g_c = pt_add(o1, g_c, pt_neg(o1, pt_mul(o1, z.delta_g1, r * s % R)))
    g_c = pt_add(o1, pt_add(o1, g_c, l_acc), h_acc)
    return (g_a, g2_b, g_c)


def partial_proof_to_bytes_le(partial):
    """proof.rs:537-547: version | ark compressed PartialProof = Vec<bool> mask (u64 len + 1 byte each) |
    partial_pi_a | partial_rho | partial_pi_b | partial_pi_c (projective points serialise as compressed affine)"""
    mask, pi_a, rho, pi_b, pi_c = partial
    return (b"\x00" + struct.pack("<Q", len(mask)) + bytes(int(m) for m in mask) + g1_compress(pi_a) + g1_compress(rho)
            + g2_compress(pi_b) + g1_compress(pi_c))


def verify(z, proof, public_inputs):
    """ark-groth16 verify_proof: e(A,B) == e(α,β)·e(vk_x,γ)·e(C,δ)."""
    a, b, c = proof
    if len(public_inputs) + 1 != len(z.gamma_abc_g1):
        raise ValueError("MalformedVerifyingKey")
    vkx = z.gamma_abc_g1[0]
    for p, x in zip(z.gamma_abc_g1[1:], public_inputs):
        vkx = pt_add(OPS1, vkx, pt_mul(OPS1, p, x))
    return pairing_product_is_one([
        (pt_neg(OPS1, a), b), (z.alpha_g1, z.beta_g2), (vkx, z.gamma_g2), (c, z.delta_g2)])


def public_inputs_multi(pv):
    """proof.rs:870-884: ys…, root, nullifiers…, x, external_nullifier, selector_used… (as 0/1)"""
    return list(pv["ys"]) + [pv["root"]] + list(pv["nullifiers"]) + [pv["x"], pv["external_nullifier"]] + [int(bool(v)) for v in pv["selector_used"]]


def public_inputs_single(pv):
    """proof.rs:863-869: [y, root, nullifier, x, external_nullifier]"""
    return [pv["y"], pv["root"], pv["nullifier"], pv["x"], pv["external_nullifier"]]


# ----------------------------------------------------------------------------- bytes
def g1_compress(p):
    """ark-serialize 0.5 SW compressed: x LE, flags in top bits of last byte
    (bit7 = y is the larger of {y,−y}, bit6 = infinity)."""
    if p is INF:
        b = bytearray(32)
        b[31] |= 0x40
        return bytes(b)
    b = bytearray(p[0].to_bytes(32, "little"))
    if p[1] > (Q - p[1]) % Q:
        b[31] |= 0x80
    return bytes(b)


def g2_compress(p):
    if p is INF:
        b = bytearray(64)
        b[63] |= 0x40
        return bytes(b)
    (x0, x1), (y0, y1) = p
    b = bytearray(x0.to_bytes(32, "little") + x1.to_bytes(32, "little"))
    n0, n1 = (-y0) % Q, (-y1) % Q
    # Fq2 ordering: compare c1 first, then c0
    if (y1, y0) > (n1, n0):
        b[63] |= 0x80
    return bytes(b)


def _sqrt_fq(a):
    y = pow(a, (Q + 1) // 4, Q)
    return y if y * y % Q == a % Q else None


def g1_decompress(b):
    flags = b[31] & 0xC0
    if flags & 0x40:
        return INF
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), "little")
    y = _sqrt_fq((x * x * x + 3) % Q)
    if y is None:
        raise ValueError("not on curve")
    big = max(y, Q - y)
    return (x, big if flags & 0x80 else Q - big)


def _sqrt_fq2(a):
    """square root in Fq2 = Fq[u]/(u²+1) by the norm method (q ≡ 3 mod 4): a = (x0 + x1·u)² ⇒ x0² = (a0 ± √(a0² + a1²))/2"""
    a0, a1 = a[0] % Q, a[1] % Q
    if a1 == 0:
        y = _sqrt_fq(a0)
        if y is not None:
            return (y, 0)
        y = _sqrt_fq((-a0) % Q)   # −1 = u²
        return None if y is None else (0, y)
    alpha = _sqrt_fq((a0 * a0 + a1 * a1) % Q)
    if alpha is None:
        return None
    half = (Q + 1) // 2
    for sign in (1, -1):
        x0 = _sqrt_fq((a0 + sign * alpha) * half % Q)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, Q) % Q
            if F.f2_sqr((x0, x1)) == (a0, a1):
                return (x0, x1)
    return None


def g2_decompress(b):
    """ark-serialize 0.5 compressed G2: x.c0 ‖ x.c1 little-endian, flags in the top bits of the LAST byte (of c1); the y that is
    the lexicographically larger of {y, −y} (c1 compared first, then c0) when bit 7 is set.  No subgroup check here."""
    flags = b[63] & 0xC0
    if flags & 0x40:
        return INF
    x0 = int.from_bytes(bytes(b[:32]), "little")
    x1 = int.from_bytes(bytes(b[32:63]) + bytes([b[63] & 0x3F]), "little")
    x = (x0, x1)
    y = _sqrt_fq2(F.f2_add(F.f2_mul(F.f2_sqr(x), x), F.OPS2.b))
    if y is None:
        raise ValueError("not on curve")
    ny = F.f2_neg(y)
    big = y if (y[1], y[0]) > (ny[1], ny[0]) else ny
    return (x, big if flags & 0x80 else F.f2_neg(big))


def proof_from_bytes(b):
    """Proof::deserialize_compressed (rln/src/protocol/proof.rs:469): A (32) ‖ B (64) ‖ C (32)"""
    return (g1_decompress(b[:32]), g2_decompress(b[32:96]), g1_decompress(b[96:128]))


def proof_to_bytes(proof):
    a, b, c = proof
    return g1_compress(a) + g2_compress(b) + g1_compress(c)


def fr_le(v):
    return (v % R).to_bytes(32, "little")


def proof_values_to_bytes_le(pv):
    """proof.rs:192-236 (single): version | root | external_nullifier | x | y | nullifier"""
    return b"\x00" + fr_le(pv["root"]) + fr_le(pv["external_nullifier"]) + fr_le(pv["x"]) + fr_le(pv["y"]) + fr_le(pv["nullifier"])


def proof_values_to_bytes_le_multi(pv):
    """proof.rs:192-236 (multi): 0x01 | root | external_nullifier | x | vec ys | vec nullifiers | vec bool selector_used"""
    k = len(pv["ys"])
    out = b"\x01" + fr_le(pv["root"]) + fr_le(pv["external_nullifier"]) + fr_le(pv["x"])
    out += struct.pack("<Q", k) + b"".join(fr_le(v) for v in pv["ys"])
    out += struct.pack("<Q", k) + b"".join(fr_le(v) for v in pv["nullifiers"])
    return out + struct.pack("<Q", k) + bytes(int(bool(v)) for v in pv["selector_used"])


def rln_proof_to_bytes_le(proof, pv):
    """proof.rs:413-428"""
    if "ys" in pv:
        return b"\x01" + proof_to_bytes(proof) + proof_values_to_bytes_le_multi(pv)
    return b"\x00" + proof_to_bytes(proof) + proof_values_to_bytes_le(pv)


def witness_to_bytes_le_multi(secret, limit, message_ids, path_elements, path_index, x, ext_null, selector_used):
    """witness.rs:400-413 (multi): 0x01 | secret | limit | vec path | vec<u8> index | x | en | vec message_ids | vec bool selector"""
    out = b"\x01" + fr_le(secret) + fr_le(limit)
    out += struct.pack("<Q", len(path_elements)) + b"".join(fr_le(e) for e in path_elements)
    out += struct.pack("<Q", len(path_index)) + bytes(path_index)
    out += fr_le(x) + fr_le(ext_null)
    out += struct.pack("<Q", len(message_ids)) + b"".join(fr_le(m) for m in message_ids)
    return out + struct.pack("<Q", len(selector_used)) + bytes(int(bool(v)) for v in selector_used)


def witness_to_bytes_le(secret, limit, message_id, path_elements, path_index, x, ext_null):
    """witness.rs:369-415 (single): version | secret | limit | message_id | vec<Fr> | vec<u8> | x | en"""
    out = b"\x00" + fr_le(secret) + fr_le(limit) + fr_le(message_id)
    out += struct.pack("<Q", len(path_elements)) + b"".join(fr_le(e) for e in path_elements)
    out += struct.pack("<Q", len(path_index)) + bytes(path_index)
    return out + fr_le(x) + fr_le(ext_null)
