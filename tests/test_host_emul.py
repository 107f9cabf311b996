"""CPU: the product's __host__ __device__ math headers (portable path) compiled with g++ and checked
against the pinned Python oracle — catches logic errors in field/curve/pairing/VM code before GPU time."""
import ctypes
import os
import random
import subprocess

import pytest

from pyref import fields as F
from pyref import groth16 as G
from pyref import poseidon as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, Q = F.R, F.Q


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(ROOT, "tests", "host_emul", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemul.so")
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-I", os.path.join(ROOT, "zerokit_b200", "csrc"), src, "-o", so])
    return ctypes.CDLL(so)


def b32(v):
    return int(v).to_bytes(32, "little")


def test_field_ops(emu):
    rnd = random.Random(5)

    def fop(field, op, a, b=0):
        out = ctypes.create_string_buffer(32)
        emu.emu_field_op(field, op, b32(a), b32(b), out)
        return int.from_bytes(out.raw, "little")
    for field, p in ((0, R), (1, Q)):
        cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (p - 1, 1), (0, 5)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(300)]
        for a, b in cases:
            assert fop(field, 0, a, b) == a * b % p
            assert fop(field, 1, a, b) == (a + b) % p
            assert fop(field, 2, a, b) == (a - b) % p
            assert fop(field, 4, a) == (-a) % p
        for _ in range(5):
            a = rnd.randrange(1, p)
            assert fop(field, 3, a) == pow(a, -1, p)


def test_poseidon(emu):
    for v in ([0], [1], [R - 1], [0, 1], [5, R - 2], [1, 2, 3], [R - 1, 0, 7]):
        out = ctypes.create_string_buffer(32)
        emu.emu_poseidon(b"".join(b32(x) for x in v), len(v), out)
        assert int.from_bytes(out.raw, "little") == P.poseidon(v)


def _g1b(p):
    return b"\0" * 64 if p is None else b32(p[0]) + b32(p[1])


def _g1r(buf):
    x, y = int.from_bytes(buf[:32], "little"), int.from_bytes(buf[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


def _g2b(p):
    return b"\0" * 128 if p is None else b32(p[0][0]) + b32(p[0][1]) + b32(p[1][0]) + b32(p[1][1])


def _g2r(buf):
    v = [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(4)]
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))


def test_curve_ops(emu):
    rnd = random.Random(6)
    for _ in range(3):
        k, k2 = rnd.randrange(R), rnd.randrange(R)
        Pt, Qt = F.pt_mul(F.OPS1, F.G1_GEN, k2), F.pt_mul(F.OPS1, F.G1_GEN, k)
        out = ctypes.create_string_buffer(64)
        emu.emu_g1_mul(_g1b(Pt), b32(k), out)
        assert _g1r(out.raw) == F.pt_mul(F.OPS1, Pt, k)
        emu.emu_g1_add(_g1b(Pt), _g1b(Qt), out)
        assert _g1r(out.raw) == F.pt_add(F.OPS1, Pt, Qt)
        emu.emu_g1_add_full(_g1b(Pt), _g1b(Qt), out)
        assert _g1r(out.raw) == F.pt_double(F.OPS1, F.pt_add(F.OPS1, Pt, Qt))
        emu.emu_g1_add(_g1b(Pt), _g1b(Pt), out)
        assert _g1r(out.raw) == F.pt_double(F.OPS1, Pt)
        emu.emu_g1_add(_g1b(Pt), _g1b(F.pt_neg(F.OPS1, Pt)), out)
        assert _g1r(out.raw) is None
        emu.emu_g1_add(_g1b(None), _g1b(Pt), out)
        assert _g1r(out.raw) == Pt
        P2, Q2 = F.pt_mul(F.OPS2, F.G2_GEN, k2), F.pt_mul(F.OPS2, F.G2_GEN, k)
        o2 = ctypes.create_string_buffer(128)
        emu.emu_g2_mul(_g2b(P2), b32(k), o2)
        assert _g2r(o2.raw) == F.pt_mul(F.OPS2, P2, k)
        emu.emu_g2_add(_g2b(P2), _g2b(Q2), o2)
        assert _g2r(o2.raw) == F.pt_add(F.OPS2, P2, Q2)
        emu.emu_g2_add(_g2b(P2), _g2b(P2), o2)
        assert _g2r(o2.raw) == F.pt_double(F.OPS2, P2)


@pytest.mark.parametrize("consts_resident", [1, 0])
@pytest.mark.parametrize("depth,sub", [(10, ""), (20, ""), (20, "multi_message_id/max_out_4")])
def test_scheduled_witness_vm(emu, depth, sub, consts_resident):
    """the bundle schedule + operand-source encoding of k_witness (ring / constant table / vals, store flag), emulated on the host
    for one proof, reproduces the oracle's graph evaluation: every wire, and every node the schedule marks as stored"""
    path = os.path.join(ROOT, "zerokit_b200", "resources", f"tree_depth_{depth}", sub, "graph.bin")
    graph = open(path, "rb").read()
    g = G.parse_graph(graph)
    pe = [P.poseidon([i + 7]) for i in range(depth)]
    idx = [(5 * i + 1) % 2 for i in range(depth)]
    if sub:
        buf = G.inputs_buffer(g, 424242, 50, [3, 7, 11, 0], pe, idx, 1234567, 89, selector_used=[1, 0, 1, 0])
    else:
        buf = G.inputs_buffer(g, 123456789, 100, 1, pe, idx, 42, 100)
    want = G.evaluate_nodes(g, buf) if hasattr(G, "evaluate_nodes") else None
    inputs = b"".join(int(v).to_bytes(32, "little") for v in buf)
    out = ctypes.create_string_buffer(32 * len(g.nodes))
    nb, stored = ctypes.c_uint32(), ctypes.c_uint32()
    assert emu.emu_witness_scheduled(graph, len(graph), inputs, out, ctypes.byref(nb), consts_resident, ctypes.byref(stored)) == 0
    vals = [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(len(g.nodes))]
    wires = G.evaluate(g, buf)
    assert [vals[s] for s in g.signals] == wires
    assert nb.value < len(g.nodes) // 2        # the schedule really is ≥ 2 nodes wide on average
    assert len(set(g.signals)) <= stored.value < len(g.nodes) // 2   # most nodes never travel to HBM
    if want is not None:
        assert all(v == w for v, w in zip(vals, want) if v != (1 << 256) - 1)


@pytest.mark.parametrize("depth,sub", [(10, ""), (20, ""), (20, "multi_message_id/max_out_4")])
def test_depth_reduced_witness_program(emu, depth, sub):
    """host_util.hpp vm_optimize_program (re-association of the graph's sums and products so that the value that is ready last
    is combined last; constants folded) followed by the bundle schedule, as the product runs it: every wire of the witness
    equals the oracle's evaluation of the ORIGINAL graph, and the schedule is at least a third shorter than the graph's
    10 000-node dependency chain"""
    path = os.path.join(ROOT, "zerokit_b200", "resources", f"tree_depth_{depth}", sub, "graph.bin")
    graph = open(path, "rb").read()
    g = G.parse_graph(graph)
    rnd = random.Random(depth)
    for trial in range(2):
        pe = [P.poseidon([i + 7 + trial]) for i in range(depth)]
        idx = [rnd.randrange(2) for i in range(depth)]
        if sub:
            buf = G.inputs_buffer(g, rnd.randrange(R), 50, [3, 7, 11, 0], pe, idx, rnd.randrange(R), 89, selector_used=[1, 0, 1, 0])
        else:
            buf = G.inputs_buffer(g, rnd.randrange(R), 100, 1 + trial, pe, idx, rnd.randrange(R), 100)
        inputs = b"".join(int(v).to_bytes(32, "little") for v in buf)
        out = ctypes.create_string_buffer(32 * len(g.signals))
        stats = (ctypes.c_uint32 * 6)()
        assert emu.emu_witness_optimized(graph, len(graph), inputs, out, stats) == 0
        wires = [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(len(g.signals))]
        assert wires == G.evaluate(g, buf)
    nodes, consts, bundles, stored, far, operands = stats
    assert bundles < 0.55 * (10000 if depth == 20 else 5440) and consts <= 1536 and stored < nodes // 3
    assert far < 0.05 * operands        # operands that come from HBM instead of the shared-memory ring or the constant table


def test_glv_split_and_double_mul(emu):
    """k ≡ k1 + k2·λ with |ki| < 2^128, and the Straus double multiplication of the proof assembly: kp·P + kq·Q"""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    rnd = random.Random(10)
    for k in [0, 1, 77, R - 1, lam, 1 << 253] + [rnd.randrange(R) for _ in range(200)]:
        out = ctypes.create_string_buffer(36)
        emu.emu_glv_split(b32(k), out)
        k1, k2 = int.from_bytes(out.raw[:16], "little"), int.from_bytes(out.raw[16:32], "little")
        k1, k2 = (-k1 if out.raw[32] else k1), (-k2 if out.raw[33] else k2)
        assert (k1 + k2 * lam - k) % R == 0 and abs(k1) < 1 << 128 and abs(k2) < 1 << 128
    Pt, Qt = F.pt_mul(F.OPS1, F.G1_GEN, 1234567), F.pt_mul(F.OPS1, F.G1_GEN, 7654321)
    for kp, kq in [(77, 44), (1, 0), (0, 5), (R - 1, R - 2), (lam, lam + 1)] + [(rnd.randrange(R), rnd.randrange(R)) for _ in range(6)]:
        out = ctypes.create_string_buffer(64)
        emu.emu_glv_double_mul(_g1b(Pt), b32(kp), _g1b(Qt), b32(kq), 1, out)
        assert _g1r(out.raw) == F.pt_add(F.OPS1, F.pt_mul(F.OPS1, Pt, kp), F.pt_mul(F.OPS1, Qt, kq)), (kp, kq)
        emu.emu_glv_double_mul(_g1b(Pt), b32(kp), _g1b(Qt), b32(kq), 0, out)
        assert _g1r(out.raw) == F.pt_mul(F.OPS1, Pt, kp)


def test_g2_subgroup_check(emu):
    """ψ(P) = [6x²]P accepts exactly the r-torsion: multiples of the generator pass, other points of the twist fail"""
    rnd = random.Random(9)
    for _ in range(3):
        P2 = F.pt_mul(F.OPS2, F.G2_GEN, rnd.randrange(1, R))
        assert emu.emu_g2_in_subgroup(_g2b(P2)) == 1
        assert F.pt_add(F.OPS2, F.pt_mul(F.OPS2, P2, R - 1), P2) is None
    # points of E'(Fq2) found by solving y² = x³ + b' for small x: the cofactor is ≈ 2^254, they are not in G2
    bt = F.OPS2.b
    found = 0
    k = 1
    while found < 3:
        x = (k, 1)
        k += 1
        rhs = F.f2_add(F.f2_mul(F.f2_sqr(x), x), bt)
        y = _f2_sqrt(rhs)
        if y is None:
            continue
        Pt = (x, y)
        assert F.on_curve(F.OPS2, Pt)
        assert F.pt_add(F.OPS2, F.pt_mul(F.OPS2, Pt, R - 1), Pt) is not None      # [r]P ≠ ∞: outside the subgroup (pt_mul reduces k mod r)
        assert emu.emu_g2_in_subgroup(_g2b(Pt)) == 0
        assert emu.emu_g2_subgroup_both(_g2b(Pt)) == 0          # both tests refuse it
        # a subgroup point plus a point outside is outside; a small multiple of an outside point stays outside
        for other in (F.pt_add(F.OPS2, Pt, F.G2_GEN), F.pt_double(F.OPS2, Pt), F.pt_add(F.OPS2, F.pt_double(F.OPS2, Pt), Pt)):
            assert emu.emu_g2_subgroup_both(_g2b(other)) == 0
        found += 1
    for _ in range(6):   # and both accept the r-torsion, including small multiples and the negated generator
        P2 = F.pt_mul(F.OPS2, F.G2_GEN, rnd.randrange(1, R))
        assert emu.emu_g2_subgroup_both(_g2b(P2)) == 3
    for k in (1, 2, 3, R - 1):
        assert emu.emu_g2_subgroup_both(_g2b(F.pt_mul(F.OPS2, F.G2_GEN, k))) == 3


def _f2_sqrt(a):
    """square root in Fq2 = Fq[u]/(u²+1), q ≡ 3 mod 4 (complex method); None if a is not a square"""
    a0, a1 = a
    if a1 == 0:
        r = pow(a0, (Q + 1) // 4, Q)
        if r * r % Q == a0:
            return (r, 0)
        r = pow(-a0 % Q, (Q + 1) // 4, Q)
        return (0, r) if r * r % Q == -a0 % Q else None
    n = (a0 * a0 + a1 * a1) % Q
    alpha = pow(n, (Q + 1) // 4, Q)
    if alpha * alpha % Q != n:
        return None
    for al in (alpha, -alpha % Q):
        delta = (a0 + al) * pow(2, -1, Q) % Q
        x0 = pow(delta, (Q + 1) // 4, Q)
        if x0 * x0 % Q != delta:
            continue
        x1 = a1 * pow(2 * x0, -1, Q) % Q
        if F.f2_sqr((x0, x1)) == (a0 % Q, a1 % Q):
            return (x0, x1)
    return None


def test_pairing(emu):
    a, b = 1234567, 7654321
    P1, Q1 = F.pt_mul(F.OPS1, F.G1_GEN, a), F.pt_mul(F.OPS2, F.G2_GEN, b)
    P2 = F.pt_neg(F.OPS1, F.pt_mul(F.OPS1, F.G1_GEN, a * b))
    assert emu.emu_pairing_check(_g1b(P1) + _g1b(P2), _g2b(Q1) + _g2b(F.G2_GEN), 2) == 1
    assert emu.emu_pairing_check(_g1b(P1) + _g1b(P1), _g2b(Q1) + _g2b(F.G2_GEN), 2) == 0


def test_verifier_fast_paths(emu):
    """inversion-free Miller loop, sparse line products, cyclotomic squaring and the merged Groth16 loop against the plain versions
    (tests/host_emul/emul.cpp emu_pairing_fast_paths), on valid-looking and on arbitrary inputs"""
    rnd = random.Random(21)
    for _ in range(3):
        g1 = b"".join(_g1b(F.pt_mul(F.OPS1, F.G1_GEN, rnd.randrange(1, R))) for _ in range(3))
        g2 = b"".join(_g2b(F.pt_mul(F.OPS2, F.G2_GEN, rnd.randrange(1, R))) for _ in range(3))
        assert emu.emu_pairing_fast_paths(g1, g2) == 31
    # points at infinity among the G1 arguments drop their factor in both forms
    g1 = _g1b(F.pt_mul(F.OPS1, F.G1_GEN, 5)) + _g1b(None) + _g1b(F.pt_mul(F.OPS1, F.G1_GEN, 9))
    g2 = b"".join(_g2b(F.pt_mul(F.OPS2, F.G2_GEN, k)) for k in (3, 4, 5))
    assert emu.emu_pairing_fast_paths(g1, g2) == 31


def _snarkjs_kat(goldens, multi=False):
    from common import multi_kat
    if multi:
        c, pub = multi_kat(goldens["ref"]["groth16_verifier_multi"])
        return ((c[0], c[1]), ((c[2], c[3]), (c[4], c[5])), (c[6], c[7])), pub
    v = goldens["ref"]["groth16_verifier_single"]
    proof = ((int(v["pi_a"][0]), int(v["pi_a"][1])),
             ((int(v["pi_b"][0][0]), int(v["pi_b"][0][1])), (int(v["pi_b"][1][0]), int(v["pi_b"][1][1]))),
             (int(v["pi_c"][0]), int(v["pi_c"][1])))
    return proof, [int(v[k]) for k in ("y", "root", "nullifier", "x", "external_nullifier")]


@pytest.mark.parametrize("multi", [False, True])
def test_pairing_vm_program(emu, goldens, multi):
    """the lane-parallel verifier (verify_vm*.hpp: program traced and scheduled by the product's host code, interpreted here lane
    by lane with the portable arithmetic) on the reference's hard-coded snarkjs proofs (rln/tests/public.rs:77-213) and their
    mutations — flags, sign bits, neighbouring coordinates on and off the curve, a twist point outside G2, wrong public inputs —
    against the Python oracle's deserialisation + pairing check; inputs with a point at infinity must be handed back (code 3)"""
    from common import multi_resource, resource, verifier_expected_code, verifier_mutations
    z = G.parse_zkey(multi_resource("rln_final.arkzkey") if multi else resource(20, "rln_final.arkzkey"))
    proof, pub = _snarkjs_kat(goldens, multi)
    vk = _g1b(z.alpha_g1) + _g2b(z.beta_g2) + _g2b(z.gamma_g2) + _g2b(z.delta_g2)
    gabc = b"".join(_g1b(p) for p in z.gamma_abc_g1)
    muts = verifier_mutations(G.proof_to_bytes(proof), pub)
    if multi:
        muts = muts[:8] + muts[-3:]
    n = len(muts)
    out = (ctypes.c_int * n)()
    info = (ctypes.c_uint64 * 6)()
    assert emu.emu_verify_vm(vk, gabc, len(pub), b"".join(b"".join(b32(x) for x in q) for _, _, q in muts), b"".join(p for _, p, _ in muts), n, out, info) == 0
    want = [verifier_expected_code(z, p, q) for _, p, q in muts]
    want = [3 if w is None else 4 if w == 0 else w for w in want]   # the VM's status codes: 4 = invalid, 3 = not decided here
    assert list(out) == want, [(m[0], o, w) for m, o, w in zip(muts, out, want) if o != w]
    assert want[0] == 1 and 2 in want and 4 in want
    levels, slots, consts = info[0], info[1], info[2]
    assert levels < 2500 and slots <= 4096 and consts < slots


def test_final_exponentiation_chain(emu):
    """the BN addition-chain hard part agrees with plain exponentiation by (q⁴−q²+1)/r on the verifier's predicate"""
    a, b = 1234567, 7654321
    P1, Q1 = F.pt_mul(F.OPS1, F.G1_GEN, a), F.pt_mul(F.OPS2, F.G2_GEN, b)
    P2 = F.pt_neg(F.OPS1, F.pt_mul(F.OPS1, F.G1_GEN, a * b))
    assert emu.emu_final_exp_consistency(_g1b(P1) + _g1b(P2), _g2b(Q1) + _g2b(F.G2_GEN), 2) == 15   # product is one
    assert emu.emu_final_exp_consistency(_g1b(P1) + _g1b(P1), _g2b(Q1) + _g2b(F.G2_GEN), 2) == 7    # product is not one


def test_vm_ops(emu):
    rnd = random.Random(7)
    for op in range(20):
        for _ in range(60):
            a = rnd.choice([0, 1, 2, R - 1, R // 2, R // 2 + 1, rnd.randrange(R), rnd.randrange(1 << 64)])
            b = rnd.choice([0, 1, 2, 63, 64, 253, 254, 255, R - 1, R // 2 + 1, rnd.randrange(R), rnd.randrange(300)])
            out = ctypes.create_string_buffer(32)
            ok = emu.emu_vm_duo(op, b32(a), b32(b), out)
            try:
                exp = G._duo(op, a, b)
            except ValueError:
                exp = None
            if exp is None:
                assert ok == 0, (op, a, b)
            else:
                assert ok == 1 and int.from_bytes(out.raw, "little") == exp, (G.OP_NAMES[op], a, b)


def test_tree_config_parser(emu):
    """the JSON tree configuration of rln/src/pm_tree_adapter.rs:139-174: every key optional, wrong-typed values fall back to the
    default (serde_json's as_str / as_bool / as_u64 return None), unknown keys ignored, syntax errors reported"""
    def parse(text):
        path, err = ctypes.create_string_buffer(512), ctypes.create_string_buffer(512)
        nums = (ctypes.c_uint64 * 8)()
        rc = emu.emu_parse_tree_config(text.encode(), path, 512, nums, err, 512)
        if rc:
            raise ValueError(err.value.decode())
        keys = ("temporary", "has_path", "cache_capacity", "flush_every_ms", "low_space", "use_compression", "has_depth", "tree_depth")
        return dict(zip(keys, list(nums)), path=path.value.decode())
    d = parse("{}")
    assert d == dict(temporary=1, has_path=0, cache_capacity=1073741824, flush_every_ms=500, low_space=0, use_compression=0, has_depth=0, tree_depth=0, path="")
    d = parse('{"path": "/tmp/x y/\\"db\\"", "temporary": false, "cache_capacity": 12345, "flush_every_ms": 7, "mode": "LowSpace",'
              ' "use_compression": true, "tree_depth": 20, "extra": {"a": [1, 2, {"b": "}"}]}}')
    assert d == dict(temporary=0, has_path=1, cache_capacity=12345, flush_every_ms=7, low_space=1, use_compression=1, has_depth=1, tree_depth=20,
                     path='/tmp/x y/"db"')
    d = parse('{"path": 5, "temporary": "no", "cache_capacity": -3, "flush_every_ms": 1.5, "mode": "Other", "tree_depth": null}')
    assert d["has_path"] == 0 and d["temporary"] == 1 and d["cache_capacity"] == 1073741824 and d["flush_every_ms"] == 500 and d["low_space"] == 0 and d["has_depth"] == 0
    assert parse(' [1, 2] ')["temporary"] == 1          # not an object: every key reads as null
    for bad in ('{"path": "x"', '{"path" "x"}', '{"a": 1,}', '{} x', '', '{"path": "unterminated}'):
        with pytest.raises(ValueError, match="Error while reading pmtree config"):
            parse(bad)
