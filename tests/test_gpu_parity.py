"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs, against the committed golden fixtures, and through size-independent properties."""
import hashlib
import os
import random

import pytest

from common import (Q, R, TreeMirror, dense_tree_quirk_script, fr_bytes, fr_stream, ints, kat_witness_args, multi_kat, pmtree_quirk_script,
                    resource, witness_le)

pytestmark = pytest.mark.gpu

os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")  # small tables: the tests are about exactness, not speed


@pytest.fixture(scope="module")
def z():
    import zerokit_b200
    return zerokit_b200


@pytest.fixture(scope="module")
def rln20(z):
    return z.RLN.new(20)


@pytest.fixture(scope="module")
def rln10(z):
    return z.RLN.new(10)


@pytest.fixture(scope="module")
def oracle():
    from oracle import cref_binding as C
    return C


# ------------------------------------------------------------------------------- field arithmetic (PTX path)
@pytest.mark.parametrize("field,p", [(0, R), (1, Q)])
def test_field_ops_ptx_vs_integers(z, field, p):
    rnd = random.Random(100 + field)
    n = 4096
    a = [0, p - 1, 1, p - 1, 2, (p - 1) // 2] + [rnd.randrange(p) for _ in range(n - 6)]
    b = [0, p - 1, p - 1, 1, (p + 1) // 2, 2] + [rnd.randrange(p) for _ in range(n - 6)]
    A, B = fr_bytes(a), fr_bytes(b)
    assert ints(z.field_op(field, 0, A, B, n)) == [x * y % p for x, y in zip(a, b)]
    assert ints(z.field_op(field, 3, A, B, n)) == [x * y % p for x, y in zip(a, b)]  # portable CIOS on device
    assert ints(z.field_op(field, 1, A, B, n)) == [(x + y) % p for x, y in zip(a, b)]
    assert ints(z.field_op(field, 2, A, B, n)) == [(x - y) % p for x, y in zip(a, b)]
    inv = ints(z.field_op(field, 4, A[:32 * 64], B[:32 * 64], 64))
    assert inv == [pow(x, -1, p) if x else 0 for x in a[:64]]
    # dedicated squaring and the single-reduction dot products (operands with the top bit of every word set included)
    hi = [sum(w << (32 * k) for k, w in enumerate([rnd.choice([0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, rnd.getrandbits(32)]) for _ in range(7)]
                                                  + [rnd.randrange(0x30000000)])) % p for _ in range(n)]
    for aa, bb in ((a, b), (hi, a), (b, hi)):
        AA, BB = fr_bytes(aa), fr_bytes(bb)
        assert ints(z.field_op(field, 5, AA, BB, n)) == [x * x % p for x in aa]
        assert ints(z.field_op(field, 6, AA, BB, n)) == [(x * x - y * y) % p for x, y in zip(aa, bb)]
        assert ints(z.field_op(field, 7, AA, BB, n)) == [(y * y - x * x) % p for x, y in zip(aa, bb)]
        assert ints(z.field_op(field, 8, AA, BB, n)) == [x * x % p for x in aa]


def test_glv_split_kernel(z):
    """k ≡ k1 + k2·λ (mod r), |ki| < 2^128 — the split the fixed-base G1 accumulate kernel applies to every scalar"""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    assert (lam * lam + lam + 1) % R == 0
    rnd = random.Random(77)
    ks = [0, 1, 2, R - 1, R - 2, lam, lam - 1, lam + 1, R // 2, R // 3, 1 << 253, 1 << 128, (1 << 128) - 1, (1 << 64) - 1]
    ks += [rnd.randrange(R) for _ in range(4096 - len(ks))]
    for k, (k1, k2) in zip(ks, z.glv_split(fr_bytes(ks), len(ks))):
        assert (k1 + k2 * lam - k) % R == 0
        assert abs(k1) < 1 << 128 and abs(k2) < 1 << 128


def test_glv_double_mul_kernel(z):
    """the Straus / GLV routine of the proof assembly (s·g_a + r·g1_b) against plain double-and-add on Python integers"""
    from pyref import fields as F
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    rnd = random.Random(78)
    Pt, Qt = F.pt_mul(F.OPS1, F.G1_GEN, 1234567), F.pt_mul(F.OPS1, F.G1_GEN, 7654321)
    ks = [(77, 44), (1, 0), (0, 5), (R - 1, R - 2), (lam, lam + 1), (1 << 128, (1 << 127) + 3)] + [(rnd.randrange(R), rnd.randrange(R)) for _ in range(58)]
    items = [(Pt, kp, Qt, kq) for kp, kq in ks]
    assert z.glv_double_mul(items) == [F.pt_add(F.OPS1, F.pt_mul(F.OPS1, Pt, kp), F.pt_mul(F.OPS1, Qt, kq)) for kp, kq in ks]
    assert z.glv_double_mul(items, use_q=False) == [F.pt_mul(F.OPS1, Pt, kp) for kp, kq in ks]


# ------------------------------------------------------------------------------- Poseidon
def test_poseidon_reference_kats(z, goldens):
    """utils/tests/poseidon_hash_test.rs:21-130"""
    for k, v in goldens["ref"]["poseidon_single"]["cases"]:
        assert z.poseidon_hash([int(k)]) == int(v)
    t = goldens["ref"]["poseidon_pair_tree8"]
    l = [z.poseidon_hash_pair(2 * i, 2 * i + 1) for i in range(4)]
    assert [str(x) for x in l] == [t["l01"], t["l23"], t["l45"], t["l67"]]
    assert str(z.poseidon_hash_pair(z.poseidon_hash_pair(l[0], l[1]), z.poseidon_hash_pair(l[2], l[3]))) == t["root"]
    assert z.poseidon_hash([1, 2, 3]) == int(goldens["derived"]["poseidon_misc"]["t4"])
    assert z.poseidon_hash([R - 1]) == int(goldens["derived"]["poseidon_misc"]["t2_rm1"])


def test_hash_pairs_vs_oracle(z, oracle):
    fs = fr_stream(31)
    n = 5000
    vals = [next(fs) for _ in range(2 * n)]
    got = ints(z.hash_pairs(fr_bytes(vals), n))
    import ctypes
    out = ctypes.create_string_buffer(32 * n)
    oracle.lib().orc_poseidon_pairs(fr_bytes(vals), n, out, oracle.threads())
    assert got == ints(out.raw)


# ------------------------------------------------------------------------------- Merkle tree
def test_tree_kat_depth20(z, rln20, goldens):
    """rln/tests/protocol.rs:14-88 and its FFI twin rln/tests/ffi.rs:325-422"""
    k = goldens["ref"]["tree_depth20_leaf3"]
    rln20.set_tree(20)
    assert rln20.get_root() == int(goldens["ref"]["empty_tree_depth20_root"]["root"], 16)
    secret = z.hash_to_field_le(k["secret_preimage"].encode())
    leaf = z.poseidon_hash_pair(z.poseidon_hash([secret]), k["user_message_limit"])
    rln20.set_leaf(k["leaf_index"], leaf)
    assert rln20.get_root() == sum(l << (64 * i) for i, l in enumerate(k["root_limbs_le64"]))
    elems, bits = rln20.get_merkle_proof(k["leaf_index"])
    assert elems == [int(x, 16) for x in k["path_elements"]]
    assert bits == k["identity_path_index"]
    assert rln20.get_leaf(k["leaf_index"]) == leaf and rln20.leaves_set() == 4


def test_tree_vs_oracle_and_fixture(z, rln10, goldens, oracle):
    m = goldens["derived"]["merkle_d10"]
    rln10.set_tree(10)
    rln10.set_leaves_from(m["start"], [int(x) for x in m["leaves"]])
    assert rln10.get_root() == int(m["root"])
    for i, pr in m["proofs"].items():
        e, b = rln10.get_merkle_proof(int(i))
        assert [str(x) for x in e] == pr["elements"] and b == pr["index"]
    assert rln10.leaves_set() == m["start"] + len(m["leaves"])
    # one-by-one == next == batch (rln/tests/public.rs:349-427)
    fs = fr_stream(32)
    leaves = [next(fs) for _ in range(70)]
    rln10.set_tree(10)
    rln10.set_leaves_from(0, leaves)
    batch_root = rln10.get_root()
    rln10.set_tree(10)
    for i, v in enumerate(leaves):
        rln10.set_leaf(i, v)
    assert rln10.get_root() == batch_root
    rln10.set_tree(10)
    for v in leaves:
        rln10.set_next_leaf(v)
    assert rln10.get_root() == batch_root and rln10.leaves_set() == 70
    nodes = oracle.merkle_build(10, fr_bytes(leaves), 0, 70)
    assert batch_root == int.from_bytes(nodes[:32], "little")
    # delete-all == empty (public.rs), next_index unchanged by delete
    for i in range(70):
        rln10.delete_leaf(i)
    empty = oracle.merkle_build(10, b"", 0, 0)
    assert rln10.get_root() == int.from_bytes(empty[:32], "little") and rln10.leaves_set() == 70
    # atomic_operation: remove + insert, with the reference's (PmTree) placement: set_values = [min_index, start + n) is written AT
    # start, so leaves 0..9 stay, the survivors of 0..9 and the ten new leaves land at 10..29 (rln/src/pm_tree_adapter.rs:447-483)
    from pyref import poseidon as P
    rln10.set_tree(10)
    model = P.PmTree(10)
    tm = TreeMirror(rln10, model, watch=32)
    tm.set_range(0, leaves[:10])
    tm.override_range(10, leaves[10:20], [0, 3])
    exp = list(leaves[:10]) + [0] + leaves[1:3] + [0] + leaves[4:10] + leaves[10:20]
    nodes = oracle.merkle_build(10, fr_bytes(exp), 0, 30)
    assert rln10.get_root() == int.from_bytes(nodes[:32], "little") == model.root() and rln10.leaves_set() == 30
    assert rln10.get_empty_leaves_indices() == [0, 3] + list(range(20, 30))
    # get_subtree_root / get_empty_leaves_indices (rln/src/public.rs:877-887; utils/tests/merkle_tree.rs)
    assert rln10.get_subtree_root(0, 5) == rln10.get_root() and rln10.get_subtree_root(10, 5) == exp[5]
    assert rln10.get_subtree_root(9, 4) == P.poseidon([exp[4], exp[5]]) == rln10.get_subtree_root(9, 5)
    assert rln10.get_subtree_root(8, 6) == P.poseidon([P.poseidon([exp[4], exp[5]]), P.poseidon([exp[6], exp[7]])])
    tm.delete(7)
    tm.set(0, 5)
    assert rln10.get_empty_leaves_indices() == [3, 7] + list(range(20, 30))
    tm.override_range(0, [], [2, 5])          # removals only: the whole span 2..5 is reset (pm_tree_adapter.rs:427-445)
    assert [rln10.get_leaf(i) for i in range(2, 6)] == [0, 0, 0, 0]
    tm.update_next(77)
    assert rln10.leaves_set() == 31 and rln10.get_leaf(30) == 77
    with pytest.raises(z.RLNError, match="Invalid index"):
        rln10.get_subtree_root(11, 0)
    with pytest.raises(z.RLNError, match="Invalid leaf"):
        rln10.get_subtree_root(3, 1024)
    with pytest.raises(z.RLNError, match="set_range got too many leaves"):
        rln10.set_leaves_from(1020, leaves[:10])
    with pytest.raises(z.RLNError, match="Index out of bounds"):
        rln10.set_leaf(1024, 1)
    with pytest.raises(z.RLNError, match="Invalid key"):
        rln10.delete_leaf(500)                # pmtree refuses to delete an index that was never used
    with pytest.raises(z.RLNError, match="Leaf index out of bounds"):
        rln10.set_leaves_from(0, [])          # override_range with neither leaves nor indices: InvalidLeaf (pm_tree_adapter.rs:337)
    with pytest.raises(z.RLNError, match="Merkle Tree is full"):
        rln10.atomic_operation(600, leaves[:10], [0])   # 600 + (610 − 0) > 1024: pmtree's set_range refuses


def test_tree_flavour_quirks(z, rln10, goldens):
    """the reference's own tree tests, step for step, on the product with the oracle's state model beside it:
    rln/tests/poseidon_tree.rs:79-146 (default PoseidonTree = PmTree, through the V1 handle at depth 4) and
    utils/tests/merkle_tree.rs:222-312 (FullMerkleTree / OptimalMerkleTree, through RLNV3 stateful handles)"""
    from pyref import poseidon as P
    exp = goldens["ref"]["pmtree_override_range"]
    rln10.set_tree(exp["depth"])
    pmtree_quirk_script(TreeMirror(rln10, P.PmTree(exp["depth"])), exp)
    assert [rln10.get_leaf(i) for i in range(12)] == [0, 1, 0, 0, 0, 1, 2, 3, 0, 1, 2, 3] and rln10.leaves_set() == 12
    rln10.set_tree(10)
    zkey, graph = resource(10, "rln_final.arkzkey"), resource(10, "graph.bin")
    expd = goldens["ref"]["dense_tree_override_range"]
    for kind in ("full", "optimal"):
        v3 = z.RLNV3.stateful(kind, 10, zkey, graph)
        tm = TreeMirror(v3, P.DenseTree(10, kind == "optimal"))
        dense_tree_quirk_script(tm, expd)
        assert v3.leaves_set() == 12
        v3.delete_leaf(700); v3.delete_leaf(5000)     # never-used / out-of-range deletes are ignored by the dense trees
        assert v3.leaves_set() == 12
        with pytest.raises(z.RLNError, match="Invalid indices"):
            v3.atomic_operation(0, [1, 2], [])        # EmptyIndicesPolicy::Reject
        del v3
    v3 = z.RLNV3.stateful("pm", 10, zkey, graph)
    pmtree_quirk_script(TreeMirror(v3, P.PmTree(10)), exp)
    with pytest.raises(z.RLNError, match="^Pmtree error: Tree error: Invalid key"):
        v3.delete_leaf(700)


def test_tree_full_size_2pow20(z, rln20, oracle):
    """BASELINE config 3: 2^20 seeded leaves, root + 4 096 membership paths (indices from seed 4) == oracle FullMerkleTree"""
    fs = fr_stream(3)
    n = 1 << 20
    import numpy as np
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1f  # < 2^253 < r: canonical
    leaves = raw.tobytes()
    rln20.set_tree(20)
    rln20.set_leaves_from_bytes(0, leaves)
    nodes = oracle.merkle_build(20, leaves, 0, n, oracle.threads())
    assert rln20.get_root() == int.from_bytes(nodes[:32], "little")
    idx = [0, 1, n - 1, 12345, 777777] + [int(x) for x in np.random.default_rng(4).integers(0, n, size=4091)]
    el, bits = rln20.get_merkle_proofs(idx)
    for k, i in enumerate(idx):
        e, b = oracle.merkle_proof_from_nodes(nodes, 20, i)
        assert ints(el[k * 640:(k + 1) * 640]) == e and list(bits[k * 20:(k + 1) * 20]) == b
    rln20.set_tree(20)


# ------------------------------------------------------------------------------- variable-base MSM
def test_msm_fixture_and_oracle(z, goldens, oracle):
    ms = goldens["derived"]["msm_g1_48"]
    pts = b"".join(fr_bytes([int(p[0]), int(p[1])]) for p in ms["bases"])
    sc = fr_bytes([int(s) for s in ms["scalars"]])
    m = z.G1Msm(1 << 16)
    assert [str(x) for x in ints(m.msm(pts, sc, 48))] == ms["result"]
    # random 2^14 with edge scalars and an infinity base, vs the oracle's ark-rule Pippenger
    n = 1 << 14
    fs = fr_stream(1)
    ks = fr_bytes([next(fs) for _ in range(n)])
    bases = bytearray(oracle.g1_mul_gen(ks, n, oracle.threads()))
    bases[64 * 5:64 * 6] = b"\0" * 63 + b"\x40"  # point at infinity flag
    fs = fr_stream(2)
    scal = [next(fs) for _ in range(n)]
    scal[0], scal[1], scal[2], scal[3] = 0, 1, R - 1, (1 << 253)
    scb = fr_bytes(scal)
    assert m.msm(bytes(bases), scb, n) == oracle.msm_g1(bytes(bases), scb, n, oracle.threads())
    # tiny and empty inputs
    assert m.msm(bytes(bases[:64]), scb[32:64], 1) == bytes(bases[:64])
    assert m.msm(b"", b"", 0) == b"\0" * 64
    # linearity: MSM(P, a) + MSM(P, b) == MSM(P, a+b)  checked through the oracle's group law on a 2-term MSM
    a = [next(fs) for _ in range(n)]
    b = [next(fs) for _ in range(n)]
    ra, rb = m.msm(bytes(bases), fr_bytes(a), n), m.msm(bytes(bases), fr_bytes(b), n)
    rab = m.msm(bytes(bases), fr_bytes([(x + y) % R for x, y in zip(a, b)]), n)
    assert oracle.msm_g1(ra + rb, fr_bytes([1, 1]), 2) == rab


@pytest.mark.parametrize("log2n", [16, 18])
def test_msm_baseline_sizes_vs_oracle(z, oracle, log2n):
    """BASELINE.json configs[1]: random bases k_i·G (seed 1), uniform scalars (seed 2); affine result == oracle Pippenger"""
    import numpy as np
    n = 1 << log2n
    rng = np.random.default_rng(1)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 31] &= 0x1f
    bases = oracle.g1_mul_gen(ks.tobytes(), n, oracle.threads())
    sc = np.random.default_rng(2).integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x1f
    m = z.G1Msm(n)
    assert m.msm(bases, sc.tobytes(), n) == oracle.msm_g1(bases, sc.tobytes(), n, oracle.threads())


def test_msm_skewed_scalars(z, oracle):
    """scalars that pile into a few buckets (all equal; 0/1 witness-like; one huge outlier): the bucket slices + combine path"""
    import numpy as np
    n = 1 << 15
    rng = np.random.default_rng(11)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 31] &= 0x1f
    bases = oracle.g1_mul_gen(ks.tobytes(), n, oracle.threads())
    m = z.G1Msm(n)
    same = int.from_bytes(rng.integers(0, 256, size=32, dtype=np.uint8).tobytes(), "little") % R
    cases = {
        "all equal": [same] * n,
        "bits": [int(b) for b in rng.integers(0, 2, size=n)],
        "all one": [1] * n,
        "r-1 and small": [R - 1 if i % 3 == 0 else i % 7 for i in range(n)],
    }
    for name, sc in cases.items():
        sb = fr_bytes(sc)
        assert m.msm(bases, sb, n) == oracle.msm_g1(bases, sb, n, oracle.threads()), name


@pytest.mark.parametrize("log2n", [20, 22, 24])
def test_msm_large_sizes(z, oracle, log2n):
    """BASELINE configs[1] at 2^20 / 2^22 (the north-star size) / 2^24: too slow to redo with the oracle's Pippenger, so
    (1) absolute: the bases are k_i·G, hence MSM(P, s) must equal (Σ k_i·s_i mod r)·G — the dot product and the single scalar
        multiplication are the oracle's (no group arithmetic of the code under test involved);
    (2) linear split: Σ over the whole == Σ over the two halves (the oracle adds the halves)."""
    import numpy as np
    import torch
    n = 1 << log2n
    m = z.G1Msm(n)
    dev = torch.device("cuda")
    rng = np.random.default_rng(5 + log2n)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 31] &= 0x1f
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x1f
    d_k, d_s = torch.from_numpy(ks).to(dev), torch.from_numpy(sc).to(dev)
    d_b = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    d_o = torch.empty(3 * 64, dtype=torch.uint8, device=dev)
    m.gen_bases(d_k.data_ptr(), n, d_b.data_ptr())
    h = n // 2
    m.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr())
    m.msm_device(d_b.data_ptr(), d_s.data_ptr(), h, d_o.data_ptr() + 64)
    m.msm_device(d_b.data_ptr() + 64 * h, d_s.data_ptr() + 32 * h, h, d_o.data_ptr() + 128)
    torch.cuda.synchronize()
    o = d_o.cpu().numpy().tobytes()
    assert oracle.msm_g1(o[64:192], fr_bytes([1, 1]), 2) == o[:64]
    dot = oracle.fr_dot(ks.tobytes(), sc.tobytes(), n, oracle.threads())
    assert oracle.g1_mul_gen(dot, 1) == o[:64]
    # the device base generator agrees with the oracle's k·G on a sample from both ends of the array (d_b holds Montgomery-form
    # points, so compare through one-term MSMs)
    for lo in (0, n - 1):
        m.msm_device(d_b.data_ptr() + 64 * lo, torch.from_numpy(np.frombuffer(fr_bytes([1]), dtype=np.uint8).copy()).to(dev).data_ptr(), 1, d_o.data_ptr())
        torch.cuda.synchronize()
        assert d_o[:64].cpu().numpy().tobytes() == oracle.g1_mul_gen(ks[lo].tobytes(), 1)


def test_empty_and_full_capacity(z, rln10, oracle):
    """empty batches are no-ops; a completely full depth-10 tree equals the oracle's; indices at the capacity edge"""
    assert rln10.prove_batch(b"", 0, b"") == b""
    assert rln10.verify_batch(b"", 0) == []
    rln10.set_tree(10)
    fs = fr_stream(44)
    leaves = [next(fs) for _ in range(1024)]
    rln10.init_tree_with_leaves(leaves)
    nodes = oracle.merkle_build(10, fr_bytes(leaves), 0, 1024)
    assert rln10.get_root() == int.from_bytes(nodes[:32], "little") and rln10.leaves_set() == 1024
    e, b = rln10.get_merkle_proof(1023)
    assert (e, b) == oracle.merkle_proof_from_nodes(nodes, 10, 1023)
    with pytest.raises(z.RLNError, match="Pmtree error: Tree error: Index out of bounds"):
        rln10.set_next_leaf(5)   # tree is full: pmtree's `set` refuses key == capacity
    with pytest.raises(z.RLNError, match="Pmtree error: Tree error: Index out of bounds"):
        rln10.get_merkle_proof(1024)
    rln10.set_tree(10)


# ------------------------------------------------------------------------------- witness / QAP / proofs
@pytest.mark.parametrize("depth,key", [(10, "kat_proof_d10"), (20, "kat_proof_d20"), (20, "kat_proof_d20_r0")])
def test_known_answer_proofs(z, goldens, depth, key, rln10, rln20):
    """bit-exact against the golden proofs (r = 44, s = 77 as rln/tests/protocol.rs:234-235)"""
    k = goldens["derived"][key]
    rln = rln10 if depth == 10 else rln20
    wb = witness_le(*kat_witness_args(depth, k["inputs"]))
    assert wb.hex() == k["witness_le_hex"]
    w, h = rln.debug_witness_and_h(wb)
    assert hashlib.sha256(w).hexdigest() == k["w_sha256"]
    assert hashlib.sha256(h).hexdigest() == k["h_sha256"]
    wit = z.RLNWitnessInput.from_bytes_le(wb)
    proof = rln.generate_rln_proof_with_rs(wit, int(k["inputs"]["r"]), int(k["inputs"]["s"]))
    assert proof.to_bytes_le().hex() == k["rln_proof_le_hex"]
    pv = proof.values
    assert {n: str(getattr(pv, n)) for n in ("root", "x", "external_nullifier", "y", "nullifier")} == k["public"]
    # verify under the product verifier and under the oracle verifier
    assert rln.verify_with_roots(proof, pv.x, []) is True
    assert rln.verify_with_roots(proof, pv.x, [pv.root]) is True
    with pytest.raises(z.RLNError, match="Expected one of the provided roots"):
        rln.verify_with_roots(proof, pv.x, [pv.root + 1])
    with pytest.raises(z.RLNError, match="Signal value does not match"):
        rln.verify_with_roots(proof, pv.x + 1, [])
    # round trip through bytes + mutated value → invalid (rln/tests/protocol.rs:638-656)
    again = z.RLNProof.from_bytes_le(proof.to_bytes_le())
    assert again.to_bytes_le() == proof.to_bytes_le()
    bad = bytearray(proof.to_bytes_le())
    bad[129 + 1 + 96] ^= 1  # y
    with pytest.raises(z.RLNError, match="Invalid proof provided"):
        rln.verify_with_roots(z.RLNProof.from_bytes_le(bytes(bad)), pv.x, [])


def test_reference_snarkjs_proof_verifies_on_gpu(z, rln20, goldens, oracle):
    """rln/tests/public.rs:77-233: the hard-coded snarkjs proof under the bundled depth-20 vk"""
    v = goldens["ref"]["groth16_verifier_single"]
    from pyref import groth16 as G
    proof = ((int(v["pi_a"][0]), int(v["pi_a"][1])),
             ((int(v["pi_b"][0][0]), int(v["pi_b"][0][1])), (int(v["pi_b"][1][0]), int(v["pi_b"][1][1]))),
             (int(v["pi_c"][0]), int(v["pi_c"][1])))
    pv = {k: int(v[k]) for k in ("root", "x", "external_nullifier", "y", "nullifier")}
    rec = G.rln_proof_to_bytes_le(proof, pv)
    p = z.RLNProof.from_bytes_le(rec)
    assert rln20.verify_with_roots(p, pv["x"], []) is True
    assert rln20.verify_batch(rec, 1) == [1]


def test_lane_parallel_verifier_agrees_with_thread_verifier(z, rln20, goldens):
    """both verifier kernels (k_verify_vm: one CTA per proof, host-scheduled program; k_verify: one thread per proof, complete
    formulas) on the reference's snarkjs proof (rln/tests/public.rs:77-142) and 43 mutations of it — sign bits, flags,
    neighbouring coordinates on and off the curve, twist points outside G2, non-canonical coordinates, wrong public inputs:
    the same code for every input, and the code the Python oracle's deserialisation + pairing check gives"""
    from pyref import groth16 as G
    from common import verifier_expected_code, verifier_mutations
    v = goldens["ref"]["groth16_verifier_single"]
    proof = ((int(v["pi_a"][0]), int(v["pi_a"][1])),
             ((int(v["pi_b"][0][0]), int(v["pi_b"][0][1])), (int(v["pi_b"][1][0]), int(v["pi_b"][1][1]))),
             (int(v["pi_c"][0]), int(v["pi_c"][1])))
    pub = [int(v[k]) for k in ("y", "root", "nullifier", "x", "external_nullifier")]
    zk = G.parse_zkey(resource(20, "rln_final.arkzkey"))
    muts = [m for m in verifier_mutations(G.proof_to_bytes(proof), pub) if all(x < R for x in m[2])]
    recs = b"".join(b"\x00" + p + G.proof_values_to_bytes_le(dict(y=q[0], root=q[1], nullifier=q[2], x=q[3], external_nullifier=q[4])) for _, p, q in muts)
    n = len(muts)
    info = rln20.verify_vm_info()
    assert 0 < info["levels"] < 2500 and info["slots"] <= 4096
    try:
        rln20.set_verify_vm_max(0)
        serial = rln20.verify_batch(recs, n)
        rln20.set_verify_vm_max(4096)
        vm = rln20.verify_batch(recs, n)
        one_by_one = [rln20.verify_batch(recs[290 * j:290 * (j + 1)], 1)[0] for j in range(n)]
    finally:
        rln20.set_verify_vm_max(4096)
    assert vm == serial, [(m[0], a, b) for m, a, b in zip(muts, vm, serial) if a != b]
    assert one_by_one == serial
    want = [verifier_expected_code(zk, p, q) for _, p, q in muts]
    assert [a for a, w in zip(serial, want) if w is not None] == [w for w in want if w is not None]
    assert serial[0] == 1 and 0 in serial and 2 in serial


def _make_batch(rln, oracle_ctx, depth, n, seed):
    """SURVEY §8d config 4 generator scaled down: member j of a seeded tree, message_id = j mod 100"""
    from pyref import poseidon as P
    from oracle import cref_binding as C
    fs = fr_stream(seed)
    secrets = [next(fs) for _ in range(n)]
    limit = 100
    leaves = [C.poseidon([C.poseidon([s]), limit]) for s in secrets]
    rln.set_tree(depth)
    rln.set_leaves_from_bytes(0, fr_bytes(leaves))
    el, bits = rln.get_merkle_proofs(list(range(n)))
    en = P.poseidon([P.hash_to_field_le(b"test-epoch"), P.hash_to_field_le(b"test-rln-identifier")])
    recs, rs, inputs = [], [], []
    for j in range(n):
        pe = ints(el[j * depth * 32:(j + 1) * depth * 32])
        ix = list(bits[j * depth:(j + 1) * depth])
        x = next(fs)
        recs.append(witness_le(secrets[j], limit, j % 100, pe, ix, x, en))
        rs += [next(fs), next(fs)]
        inputs.append(oracle_ctx.inputs_buffer(secrets[j], limit, j % 100, pe, ix, x, en))
    return b"".join(recs), fr_bytes(rs), b"".join(inputs), rln.get_root()


@pytest.mark.parametrize("depth,n", [(10, 37), (20, 70)])
def test_batch_proofs_bit_equal_to_oracle(z, oracle, depth, n, rln10, rln20):
    rln = rln10 if depth == 10 else rln20
    ctx = oracle.Ctx(resource(depth, "rln_final.arkzkey"), resource(depth, "graph.bin"))
    recs, rs, inputs, root = _make_batch(rln, ctx, depth, n, 50 + depth)
    out = rln.prove_batch(recs, n, rs)
    want_proofs, want_pub = ctx.prove_batch(inputs, rs, n, oracle.threads())
    from pyref import groth16 as G
    for j in range(n):
        rec = out[290 * j:290 * (j + 1)]
        v = ints(want_proofs[256 * j:256 * (j + 1)])
        proof = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
        y, rt, nul, x, en = ints(want_pub[160 * j:160 * (j + 1)])
        assert rt == root
        assert rec == G.rln_proof_to_bytes_le(proof, dict(root=rt, external_nullifier=en, x=x, y=y, nullifier=nul)), j
    # every GPU proof verifies under the oracle verifier and under the GPU verifier
    proofs_aff = want_proofs  # bit-equal to the GPU's (checked above through the compressed form)
    assert ctx.verify_batch(proofs_aff, want_pub, n, 5, oracle.threads()) == [1] * n
    assert rln.verify_batch(out, n) == [1] * n
    # stateful verification against the tree root (public.rs:725-745)
    p0 = z.RLNProof.from_bytes_le(out[:290])
    assert rln.verify_rln_proof(p0, p0.values.x) is True
    rln.set_leaf(0, 12345)
    with pytest.raises(z.RLNError, match="Expected one of the provided roots"):
        rln.verify_rln_proof(p0, p0.values.x)
    # fresh randomness path (ffi_generate_rln_proof): different bytes, still valid
    wit = z.RLNWitnessInput.from_bytes_le(recs[:len(recs) // n])
    pa, pb = rln.generate_rln_proof(wit), rln.generate_rln_proof(wit)
    assert pa.proof_bytes != pb.proof_bytes
    assert rln.verify_with_roots(pa, pa.values.x, []) and rln.verify_with_roots(pb, pb.values.x, [])


def test_production_window_tables_batch_4096(z, oracle):
    """BASELINE configs[3] exactly as bench.py runs it: the DEFAULT window sizes (G1 c = 13, G2 c = 15,
    GLV; ≈ 125 GiB of tables, which every other test replaces by c = 8), batch 4 096 at depth 20: the first 64 proofs bit-equal
    to the oracle's, ALL 4 096 verified by the ORACLE's pairing verifier (not the product's), and by the product's."""
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 150 << 30:
        pytest.skip("needs ≈ 135 GiB of free HBM for the production tables")
    saved = {k: os.environ.pop(k, None) for k in ("RLN_B200_WINDOW_BITS", "RLN_B200_WINDOW_BITS_G2")}
    try:
        rln = z.RLN.new(20)
    finally:
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
    info = rln.table_info()
    assert info["window_bits"] == 13 and info["window_bits_g2"] == 15 and info["glv"], info
    ctx = oracle.Ctx(resource(20, "rln_final.arkzkey"), resource(20, "graph.bin"))
    n, n_eq = 4096, 64
    recs, rs, inputs, root = _make_batch(rln, ctx, 20, n, 123)
    out = rln.prove_batch(recs, n, rs)
    th = oracle.threads()
    want_proofs, want_pub = ctx.prove_batch(inputs[:n_eq * ctx.inputs_size * 32], rs[:64 * n_eq], n_eq, th)
    from pyref import groth16 as G
    for j in range(n_eq):
        v = ints(want_proofs[256 * j:256 * (j + 1)])
        proof = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
        y, rt, nul, x, en = ints(want_pub[160 * j:160 * (j + 1)])
        assert out[290 * j:290 * (j + 1)] == G.rln_proof_to_bytes_le(proof, dict(root=rt, external_nullifier=en, x=x, y=y, nullifier=nul)), j
    # every one of the 4 096 under the ORACLE verifier, fed from the product's final bytes: oracle decompression (pyref, ark-serialize
    # rules) of each 128-byte proof → affine coordinates → the C++ oracle's pairing check against the record's own public values
    pts, pubs = [], []
    for j in range(n):
        rec = out[290 * j:290 * (j + 1)]
        a, b, c = G.proof_from_bytes(rec[1:129])
        pts.append(fr_bytes([a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]]))
        rt, en, x, y, nul = ints(rec[130:290])
        assert rt == root
        pubs.append(fr_bytes([y, rt, nul, x, en]))
    assert ctx.verify_batch(b"".join(pts), b"".join(pubs), n, 5, th) == [1] * n
    assert rln.verify_batch(out, n) == [1] * n
    del rln


def test_partial_proofs(z, rln10, rln20, goldens, oracle):
    """rln/tests/protocol.rs:222-248: with fixed (r, s), full proof == partial + finish; partial bytes == the oracle's"""
    k = goldens["derived"]["kat_proof_d10"]
    pg = goldens["derived"]["partial_proof_d10"]
    args = kat_witness_args(10, k["inputs"])
    wit = z.RLNWitnessInput.from_bytes_le(witness_le(*args))
    pw = z.RLNPartialWitnessInput.new(args[0], args[1], args[3], args[4])
    partial = rln10.generate_partial_zk_proof(pw)
    assert partial.to_bytes_le().hex() == pg["partial_le_hex"]
    r, s = int(k["inputs"]["r"]), int(k["inputs"]["s"])
    fin = rln10.finish_rln_proof_with_rs(partial, wit, r, s)
    assert fin.to_bytes_le().hex() == k["rln_proof_le_hex"]
    # through bytes, and with fresh randomness
    again = rln10.partial_proof_from_bytes_le(partial.to_bytes_le())
    assert rln10.finish_rln_proof_with_rs(again, wit, r, s).to_bytes_le().hex() == k["rln_proof_le_hex"]
    p2 = rln10.finish_rln_proof(partial, wit)
    assert p2.proof_bytes != fin.proof_bytes and rln10.verify_with_roots(p2, p2.values.x, [])
    bad = bytearray(partial.to_bytes_le())
    bad[9] ^= 1  # flip a mask bit: no longer the circuit's mask
    with pytest.raises(z.RLNError, match="malformed verifying key"):
        rln10.finish_rln_proof_with_rs(rln10.partial_proof_from_bytes_le(bytes(bad)), wit, r, s)
    # batch, depth 20: one partial proof per member reused for two different messages
    ctx = oracle.Ctx(resource(20, "rln_final.arkzkey"), resource(20, "graph.bin"))
    n = 40
    recs, rs, inputs, root = _make_batch(rln20, ctx, 20, n, 91)
    partial_pts = rln20.partial_batch(recs, n)
    assert rln20.finish_batch(recs, n, partial_pts, rs) == rln20.prove_batch(recs, n, rs)
    # a second message for the same members: new message_id / x / external_nullifier, same partial points
    rec = len(recs) // n
    recs2 = bytearray(recs)
    fs = fr_stream(92)
    for j in range(n):
        b = rec * j
        recs2[b + 65:b + 97] = ((j + 1) % 100).to_bytes(32, "little")
        recs2[b + rec - 64:b + rec - 32] = next(fs).to_bytes(32, "little")
        recs2[b + rec - 32:b + rec] = (777 + j).to_bytes(32, "little")
    out2 = rln20.finish_batch(bytes(recs2), n, partial_pts, rs)
    assert out2 == rln20.prove_batch(bytes(recs2), n, rs)
    assert rln20.verify_batch(out2, n) == [1] * n


def test_seeded_keygen_kats(z, goldens):
    """rln/tests/protocol.rs:459-507 and rln/tests/ffi_utils.rs:8-69 (the commitment is hashed on the GPU)"""
    k = goldens["ref"]["seeded_keygen"]
    assert z.seeded_keygen(k["phrase"]["seed_utf8"].encode()) == (int(k["phrase"]["identity_secret"], 16), int(k["phrase"]["id_commitment"], 16))
    assert z.seeded_keygen(bytes.fromhex(k["bytes"]["seed_hex"])) == (int(k["bytes"]["identity_secret"], 16), int(k["bytes"]["id_commitment"], 16))
    e = k["extended_bytes"]
    assert z.extended_seeded_keygen(bytes.fromhex(e["seed_hex"])) == tuple(
        int(e[f], 16) for f in ("identity_trapdoor", "identity_nullifier", "identity_secret", "id_commitment"))
    # random variants: the relations of rln/tests/protocol.rs:509-520
    sec, com = z.keygen()
    assert com == z.poseidon_hash([sec]) and 0 < sec < R
    t, n, sec, com = z.extended_keygen()
    assert sec == z.poseidon_hash_pair(t, n) and com == z.poseidon_hash([sec])


def test_big_endian_proof_records_and_metadata(z, rln10, goldens):
    """rln/tests/serialize.rs round trips for the records that need the GPU to parse (point validation)"""
    k = goldens["derived"]["kat_proof_d10"]
    args = kat_witness_args(10, k["inputs"])
    wit = z.RLNWitnessInput.from_bytes_le(witness_le(*args))
    proof = rln10.generate_rln_proof_with_rs(wit, int(k["inputs"]["r"]), int(k["inputs"]["s"]))
    le, be = proof.to_bytes_le(), proof.to_bytes_be()
    assert be[:129] == le[:129] and be[129] == le[129]
    assert [be[130 + 32 * i:162 + 32 * i] for i in range(5)] == [le[130 + 32 * i:162 + 32 * i][::-1] for i in range(5)]
    assert z.RLNProof.from_bytes_be(be).to_bytes_le() == le
    with pytest.raises(z.RLNError, match="Expected to read"):
        z.RLNProof.from_bytes_be(be + b"\0")
    with pytest.raises(z.RLNError):
        z.RLNProof.from_bytes_be(be[:140])
    pw = wit.to_partial()
    partial = rln10.generate_partial_zk_proof(pw)
    pb = partial.to_bytes_le()
    assert partial.to_bytes_be() == pb and partial.version_byte == 0
    for parsed in (z.RLNPartialProof.from_bytes_le(pb), z.RLNPartialProof.from_bytes_be(pb)):   # handle-free parse
        assert parsed.to_bytes_le() == pb
        fin = rln10.finish_rln_proof_with_rs(parsed, wit, int(k["inputs"]["r"]), int(k["inputs"]["s"]))
        assert fin.to_bytes_le().hex() == k["rln_proof_le_hex"]
    rejected = 0
    for delta in range(1, 9):   # about half of all x coordinates are not on the curve
        bad = bytearray(pb)
        bad[-2] ^= delta
        try:
            z.RLNPartialProof.from_bytes_le(bytes(bad))
        except z.RLNError as e:
            assert "invalid data" in str(e)
            rejected += 1
    assert rejected >= 1
    # metadata / flush (rln/src/ffi/ffi_tree.rs:226-268)
    assert rln10.get_metadata() == b""
    rln10.set_metadata(b"block 1234")
    assert rln10.get_metadata() == b"block 1234"
    rln10.flush()


def test_proof_from_external_witness(z, rln10, goldens):
    """rln/src/public.rs:643-658 / ffi_rln.rs:874-916: the wire assignment is calculated outside (here by the oracle's graph
    evaluator) and handed over as decimal strings; negative representatives are accepted (proof.rs:593-614)"""
    from pyref import groth16 as G
    k = goldens["derived"]["kat_proof_d10"]
    args = kat_witness_args(10, k["inputs"])
    g = G.parse_graph(resource(10, "graph.bin"))
    wires = G.evaluate(g, G.inputs_buffer(g, *args))
    wires[7] -= R                      # a negative representative of the same field element
    wires[9] += 3 * R                  # and one above the modulus
    wit = z.RLNWitnessInput.from_bytes_le(witness_le(*args))
    proof = rln10.generate_rln_proof_with_witness(wires, wit, int(k["inputs"]["r"]), int(k["inputs"]["s"]))
    assert proof.to_bytes_le().hex() == k["rln_proof_le_hex"]
    p2 = rln10.generate_rln_proof_with_witness(wires, wit)
    assert rln10.verify_with_roots(p2, p2.values.x, [])
    with pytest.raises(z.RLNError, match="calculated witness has"):
        rln10.generate_rln_proof_with_witness(wires[:-1], wit)
    wires[100] += 1                    # an inconsistent assignment still yields a proof (the prover never checks constraints) — it must not verify
    p3 = rln10.generate_rln_proof_with_witness(wires, wit, 5, 6)
    with pytest.raises(z.RLNError, match="Invalid proof provided"):
        rln10.verify_with_roots(p3, p3.values.x, [])


def test_one_handle_many_threads(z, rln10, goldens):
    """the reference's prove / verify take &self and may be called from many threads on one handle (threading);
    here the handle serialises them — results must be the same as single-threaded"""
    import threading
    k = goldens["derived"]["kat_proof_d10"]
    wb = witness_le(*kat_witness_args(10, k["inputs"]))
    r, s = int(k["inputs"]["r"]), int(k["inputs"]["s"])
    out, errs = {}, []

    def work(t):
        try:
            for i in range(3):
                wit = z.RLNWitnessInput.from_bytes_le(wb)
                p = rln10.generate_rln_proof_with_rs(wit, r, s)
                assert rln10.verify_with_roots(p, p.values.x, [])
                out[(t, i)] = p.to_bytes_le().hex()
        except Exception as e:   # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert len(out) == 12 and set(out.values()) == {k["rln_proof_le_hex"]}


def test_v3_api(z, goldens, oracle):
    """rln/tests/ffi.rs V3 section + rln/tests/public.rs: RLNV3 over the same prover — stateless and stateful builds, proof ==
    the V1 golden proof for the same (r, s), verify / verify_with_roots semantics, tree ops, V3 wire forms of the proof"""
    k = goldens["derived"]["kat_proof_d10"]
    args = kat_witness_args(10, k["inputs"])
    zkey, graph = resource(10, "rln_final.arkzkey"), resource(10, "graph.bin")
    r, s = int(k["inputs"]["r"]), int(k["inputs"]["s"])
    v3 = z.RLNV3.stateful("optimal", 10, zkey, graph)
    wit = z.WitnessV3.new_single(args[0], args[1], args[2], args[3], args[4], args[5], args[6])
    proof = v3.generate_proof_with_rs(wit, r, s)
    gold = bytes.fromhex(k["rln_proof_le_hex"])      # V1 record: version | proof | version | root | en | x | y | nullifier
    le = proof.to_bytes_le()
    assert le[:128] == gold[1:129]
    v = proof.values
    assert [v.root, v.external_nullifier, v.x, v.y, v.nullifier] == ints(gold[130:290])
    from pyref import serialize as S
    assert le[128:] == S.v3_values_single(v.y, v.root, v.nullifier, v.x, v.external_nullifier)
    mixed = proof.to_bytes_mixed()
    assert mixed[:128] == le[:128] and mixed[128:] == S.v3_values_single(v.y, v.root, v.nullifier, v.x, v.external_nullifier, be=True)
    assert z.ProofV3.from_bytes_le(le).to_bytes_mixed() == mixed and z.ProofV3.from_bytes_mixed(mixed).to_bytes_le() == le
    bad = bytearray(le)
    bad[5] ^= 1
    with pytest.raises(z.RLNError, match="invalid data"):
        z.ProofV3.from_bytes_le(bytes(bad))
    # verify: pairing only, Ok(false) on a wrong signal; verify_with_roots: root → signal → proof, errors as text
    assert v3.verify(proof, v.x) is True and v3.verify(proof, v.x + 1) is False
    assert v3.verify_with_roots(proof, v.x, []) and v3.verify_with_roots(proof, v.x, [5, v.root])
    with pytest.raises(z.RLNError, match="Expected one of the provided roots"):
        v3.verify_with_roots(proof, v.x, [5])
    with pytest.raises(z.RLNError, match="Signal value does not match"):
        v3.verify_with_roots(proof, v.x + 1, [])
    forged = z.ProofV3.from_bytes_le(le[:128] + S.v3_values_single((v.y + 1) % R, v.root, v.nullifier, v.x, v.external_nullifier))
    assert v3.verify(forged, v.x) is False
    with pytest.raises(z.RLNError, match="Invalid proof provided"):
        v3.verify_with_roots(forged, v.x, [])
    # fresh randomness + shape check against the circuit
    assert v3.verify(v3.generate_proof(wit), v.x)
    with pytest.raises(z.RLNError, match="Field `path_elements` has length 9, but circuit tree_depth is 10"):
        v3.generate_proof(z.WitnessV3.new_single(args[0], args[1], args[2], args[3][:9], args[4][:9], args[5], args[6]))
    # two-phase
    partial = v3.generate_partial_proof(wit.to_partial())
    pb = partial.to_bytes_le()
    assert b"\0" + pb == bytes.fromhex(goldens["derived"]["partial_proof_d10"]["partial_le_hex"])
    assert v3.finish_proof_with_rs(z.PartialProofV3.from_bytes_le(pb), wit, r, s).to_bytes_le() == le
    assert v3.verify(v3.finish_proof(partial, wit), v.x)
    # tree: same roots / paths as the V1 object and the oracle
    fs = fr_stream(33)
    leaves = [next(fs) for _ in range(50)]
    v3.set_leaves_from(0, leaves)
    nodes = oracle.merkle_build(10, fr_bytes(leaves), 0, 50)
    assert v3.get_root() == int.from_bytes(nodes[:32], "little") and v3.leaves_set() == 50 and v3.get_leaf(7) == leaves[7]
    e, b = v3.get_merkle_proof(13)
    oe, ob = oracle.merkle_proof_from_nodes(nodes, 10, 13)
    assert e == oe and b == ob
    from pyref import poseidon as P
    model = P.DenseTree(10, optimal=True)
    model.set_range(0, leaves)
    tm = TreeMirror(v3, model, watch=64)
    tm.update_next(99); tm.delete(3); tm.override_range(51, [7, 8], [0])
    v3.seq_atomic_operation([9], [1]); model.override_range(model.leaves_set(), [9], [1]); tm._same()
    # OptimalMerkleTree::override_range writes [min_index, start + n) AT start: 53 values at 51, then 104 at 104
    assert v3.leaves_set() == 208 and v3.get_leaf(207) == 9 and v3.get_leaf(0) == leaves[0] and v3.get_leaf(51) == 0
    v3.init_tree_with_leaves(leaves[:4])
    assert v3.get_root() == int.from_bytes(oracle.merkle_build(10, fr_bytes(leaves[:4]), 0, 4)[:32], "little")
    v3.set_metadata(b"abc"); assert v3.get_metadata() == b"abc"; v3.flush()
    del v3
    # stateless: proves and verifies, refuses every tree operation
    sl = z.RLNV3.stateless(zkey, graph)
    assert sl.generate_proof_with_rs(wit, r, s).to_bytes_le() == le and sl.verify(proof, v.x)
    assert sl.get_root() == 0 and sl.leaves_set() == 0
    for op in (lambda: sl.set_leaf(0, 1), lambda: sl.get_leaf(0), lambda: sl.get_merkle_proof(0), lambda: sl.set_metadata(b"x"), lambda: sl.flush()):
        with pytest.raises(z.RLNError, match="tree op unsupported on stateless RLN"):
            op()


def test_multi_message_id_circuit(z, goldens, oracle):
    """the bundled max_out = 4 circuit (rln/resources/tree_depth_20/multi_message_id): golden proof, verification, mode checks"""
    k = goldens["derived"]["kat_proof_multi_d20"]
    rln = z.RLN.new_multi(20, 4)
    assert rln.max_out() == 4 and rln.tree_depth() == 20
    # the reference's own known answer for this circuit: rln/tests/public.rs:143-233 (test_groth16_proof_hardcoded, MultiV1 arm)
    from pyref import groth16 as G
    v = goldens["ref"]["groth16_verifier_multi"]
    c, pub = multi_kat(v)
    kat_pv = dict(root=int(v["root"]), external_nullifier=int(v["external_nullifier"]), x=int(v["x"]), ys=[int(y) for y in v["ys"]],
                  nullifiers=[int(n) for n in v["nullifiers"]], selector_used=v["selector_used"])
    kat_rec = G.rln_proof_to_bytes_le(((c[0], c[1]), ((c[2], c[3]), (c[4], c[5])), (c[6], c[7])), kat_pv)
    kat_proof = z.RLNProof.from_bytes_le(kat_rec)
    assert kat_proof.values.public_inputs() == pub
    assert rln.verify_with_roots(kat_proof, int(v["x"]), []) is True
    assert rln.verify_batch(kat_rec, 1) == [1]
    with pytest.raises(z.RLNError, match="Signal value does not match"):
        rln.verify_with_roots(kat_proof, int(v["x"]) + 1, [])
    forged = dict(kat_pv, nullifiers=[kat_pv["nullifiers"][0] + 1, 0, 0, 0])
    with pytest.raises(z.RLNError, match="Invalid proof provided"):
        rln.verify_with_roots(z.RLNProof.from_bytes_le(G.rln_proof_to_bytes_le(((c[0], c[1]), ((c[2], c[3]), (c[4], c[5])), (c[6], c[7])), forged)),
                              int(v["x"]), [])
    wit = z.RLNWitnessInput.from_bytes_le(bytes.fromhex(k["witness_le_hex"]))
    proof = rln.generate_rln_proof_with_rs(wit, 44, 77)
    assert proof.to_bytes_le().hex() == k["rln_proof_le_hex"]
    pv = proof.values
    assert [str(v) for v in pv.ys] == k["public"]["ys"] and [str(v) for v in pv.nullifiers] == k["public"]["nullifiers"]
    assert pv.selector_used == [True, False, True, False] and pv.ys[1] == 0 and pv.nullifiers[3] == 0
    assert rln.verify_with_roots(proof, pv.x, [pv.root]) is True
    again = z.RLNProof.from_bytes_le(proof.to_bytes_le())
    assert again.to_bytes_le() == proof.to_bytes_le()
    bad = bytearray(proof.to_bytes_le())
    bad[129 + 1 + 96 + 8] ^= 1  # ys[0]
    with pytest.raises(z.RLNError, match="Invalid proof provided"):
        rln.verify_with_roots(z.RLNProof.from_bytes_le(bytes(bad)), pv.x, [])
    # batch of seeded witnesses against the oracle, plus two-phase proving on this circuit
    mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zerokit_b200", "resources", "tree_depth_20",
                        "multi_message_id", "max_out_4")
    ctx = oracle.Ctx(open(os.path.join(mdir, "rln_final.arkzkey"), "rb").read(), open(os.path.join(mdir, "graph.bin"), "rb").read())
    from pyref import groth16 as G, poseidon as P
    fs = fr_stream(77)
    n = 12
    pe = [P.poseidon([i + 7]) for i in range(20)]
    recs, rs, inputs, pvs = [], [], [], []
    for j in range(n):
        secret, x, en = next(fs), next(fs), next(fs)
        mids = [(7 * j + i) % 50 for i in range(4)]
        sel = [bool((j + i) % 3) or i == 0 for i in range(4)]
        idx = [(j >> i) & 1 for i in range(20)]
        recs.append(G.witness_to_bytes_le_multi(secret, 50, mids, pe, idx, x, en, sel))
        inputs.append(ctx.inputs_buffer(secret, 50, mids, pe, idx, x, en, sel))
        pvs.append(P.proof_values_from_witness_multi(secret, 50, mids, pe, idx, x, en, sel))
        rs += [next(fs), next(fs)]
    recs, rsb = b"".join(recs), fr_bytes(rs)
    assert len(recs) == n * rln.witness_record_len()
    out = rln.prove_batch(recs, n, rsb)
    want_p, want_pub = ctx.prove_batch(b"".join(inputs), rsb, n, oracle.threads())
    rl = rln.proof_record_len()
    for j in range(n):
        v = ints(want_p[256 * j:256 * (j + 1)])
        proof_j = ((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7]))
        assert ints(want_pub[480 * j:480 * (j + 1)]) == G.public_inputs_multi(pvs[j])
        assert out[rl * j:rl * (j + 1)] == G.rln_proof_to_bytes_le(proof_j, pvs[j]), j
    assert rln.verify_batch(out, n) == [1] * n
    assert rln.finish_batch(recs, n, rln.partial_batch(recs, n), rsb) == out
    # a single-mode witness on the multi circuit is rejected like validate_witness_against_graph does (proof.rs:658-665)
    ks = goldens["derived"]["kat_proof_d20"]
    with pytest.raises(z.RLNError, match="Witness message mode SingleV1 does not match graph mode MultiV1"):
        rln.generate_rln_proof(z.RLNWitnessInput.from_bytes_le(bytes.fromhex(ks["witness_le_hex"])))


def test_witness_errors(z, rln20):
    with pytest.raises(z.RLNError, match="tree_depth"):
        w = z.RLNWitnessInput.new_single(5, 10, 3, [1, 2], [0, 1], 7, 9)
        rln20.generate_rln_proof(w)
    with pytest.raises(z.RLNError, match="invalid data|Expected to read|Unknown message mode"):
        z.RLNProof.from_bytes_le(b"\x00" + b"\xff" * 128 + b"\x00" + b"\x00" * 160)


# ------------------------------------------------------------------------------- wire records on the device, several devices
def test_records_device_path_and_refusals(z, rln10, oracle):
    """rlnb200_prove_records_device (records parsed / formatted by k_records.cu) == rlnb200_prove_batch == the oracle; every
    record bytes_le_to_rln_witness / RLNWitnessInput::new_single would refuse (witness.rs:78-113, 470-560) is refused with the
    reference's wording, wherever it sits in the batch"""
    import torch
    ctx = oracle.Ctx(resource(10, "rln_final.arkzkey"), resource(10, "graph.bin"))
    n = 33
    recs, rs, inputs, root = _make_batch(rln10, ctx, 10, n, 611)
    want = rln10.prove_batch(recs, n, rs)
    dev = torch.device("cuda")
    d_recs = torch.frombuffer(bytearray(recs), dtype=torch.uint8).to(dev)
    d_rs = torch.frombuffer(bytearray(rs), dtype=torch.uint8).to(dev)
    d_out = torch.zeros(n * 290, dtype=torch.uint8, device=dev)
    rln10.prove_records_device(d_recs.data_ptr(), d_rs.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == want
    # fresh randomness (d_rs = NULL): different proofs, all valid
    d_out2 = torch.zeros_like(d_out)
    rln10.prove_records_device(d_recs.data_ptr(), 0, n, d_out2.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    fresh = d_out2.cpu().numpy().tobytes()
    assert fresh != want and rln10.verify_batch(fresh, n) == [1] * n
    assert [fresh[290 * j + 129:290 * (j + 1)] for j in range(n)] == [want[290 * j + 129:290 * (j + 1)] for j in range(n)]   # same values
    rec = len(recs) // n
    d = 10

    def broken(j, off, data):
        b = bytearray(recs)
        b[rec * j + off:rec * j + off + len(data)] = data
        return bytes(b)
    cases = [
        (broken(5, 33, b"\0" * 32), "User message limit cannot be zero"),
        (broken(32, 65, (100).to_bytes(32, "little")), r"Message id \(100\) is not within user_message_limit \(100\)"),
        (broken(0, 1, R.to_bytes(32, "little")), "Non-canonical field element"),
        (broken(17, rec - 32, (R + 5).to_bytes(32, "little")), "Non-canonical field element"),
        (broken(9, 105 + 32 * 3, b"\xff" * 32), "Non-canonical field element"),
        (broken(2, 0, b"\x02"), "Unknown message mode version byte"),
        (broken(3, 0, b"\x01"), "record|Expected to read|mode"),
        (broken(4, 97, (d + 1).to_bytes(8, "little")), "Expected to read|shape of the circuit"),
        (broken(6, 105 + 32 * d, (d - 1).to_bytes(8, "little")), "witness record 6: "),   # a wrong index-length prefix shifts everything after it
    ]
    for bad, pattern in cases:
        with pytest.raises(z.RLNError, match=pattern):
            rln10.prove_batch(bad, n, rs)
    assert rln10.prove_batch(recs, n, rs) == want   # the handle is unharmed by refused batches


def test_multi_device_in_process(z, oracle):
    """rlnb200_multi_*: one process, a prover replica per GPU, contiguous shards; output == the single-device output == the oracle.
    Also the per-device initialisation of the constant tables: a handle that lives on the LAST visible device hashes,
    proves and verifies like one on device 0.  Runs with whatever the box has (one device: a single-replica multi object)."""
    import torch
    ndev = torch.cuda.device_count()
    ctx = oracle.Ctx(resource(20, "rln_final.arkzkey"), resource(20, "graph.bin"))
    multi = z.RLNMulti(20, list(range(ndev)))
    assert multi.device_count() == ndev and multi.devices() == list(range(ndev))
    n = 8 * ndev + 3
    r0 = multi.replica(0)
    recs, rs, inputs, root = _make_batch(r0, ctx, 20, n, 733)
    # the same tree on every replica
    leaves = b"".join(r0.get_leaf(i).to_bytes(32, "little") for i in range(n))
    multi.set_leaves_from_bytes(0, leaves)
    for i in range(ndev):
        assert multi.replica(i).get_root() == root
    out = multi.prove_batch(recs, n, rs)
    assert out == r0.prove_batch(recs, n, rs)
    want_proofs, want_pub = ctx.prove_batch(inputs, rs, n, oracle.threads())
    from pyref import groth16 as G
    for j in range(n):
        v = ints(want_proofs[256 * j:256 * (j + 1)])
        y, rt, nul, x, en = ints(want_pub[160 * j:160 * (j + 1)])
        assert out[290 * j:290 * (j + 1)] == G.rln_proof_to_bytes_le(((v[0], v[1]), ((v[2], v[3]), (v[4], v[5])), (v[6], v[7])),
                                                                     dict(root=rt, external_nullifier=en, x=x, y=y, nullifier=nul)), j
    assert multi.verify_batch(out, n) == [1] * n
    bad = bytearray(out)
    bad[290 * (n - 1) + 200] ^= 1
    assert multi.verify_batch(bytes(bad), n) == [1] * (n - 1) + [0]
    assert len(multi.last_shard_ms()) == ndev
    # a refused record in the LAST shard is reported with its device
    broke = bytearray(recs)
    broke[len(recs) // n * (n - 1) + 33:len(recs) // n * (n - 1) + 65] = b"\0" * 32
    with pytest.raises(z.RLNError, match=rf"device {ndev - 1}: .*User message limit cannot be zero"):
        multi.prove_batch(bytes(broke), n, rs)
    # the last device on its own: single-proof API, Poseidon, pairing — from a host thread whose current device is 0
    last = multi.replica(ndev - 1)
    wit = z.RLNWitnessInput.from_bytes_le(recs[:len(recs) // n])
    p = last.generate_rln_proof_with_rs(wit, int.from_bytes(rs[:32], "little"), int.from_bytes(rs[32:64], "little"))
    assert p.to_bytes_le() == out[:290]
    assert last.verify_with_roots(p, p.values.x, [root]) is True
    multi.atomic_operation(n, [5, 6], [0])
    assert len({multi.replica(i).get_root() for i in range(ndev)}) == 1 and multi.replica(ndev - 1).leaves_set() == r0.leaves_set()
    del multi


def test_tree_persistence(z, oracle, tmp_path):
    """rln/tests/pm_tree.rs:108-131 (test_pmtree_multiple_reopen) and :433-452 (test_pmtree_persistence) through the C ABI: a handle
    opened with a persistent tree configuration (JSON file named by config_path: rln/src/ffi/ffi_rln.rs:24-46, pm_tree_adapter.rs:139-174)
    writes its tree on flush / drop; the next handle on the same path comes back with the same root, paths, leaves_set and metadata"""
    import json
    db = tmp_path / "tree.db"
    cfg = tmp_path / "config.json"
    cfg.write_text(json.dumps({"path": str(db), "temporary": False, "cache_capacity": 1 << 20, "flush_every_ms": 100, "mode": "HighThroughput",
                               "use_compression": False, "tree_depth": 10}))
    fs = fr_stream(808)
    leaves = [next(fs) for _ in range(300)]
    t1 = z.RLN.new(10, str(cfg))
    t1.set_next_leaf(42)
    t1.set_leaves_from(1, leaves)
    t1.delete_leaf(7)
    t1.atomic_operation(301, [9, 8], [300])
    root1, n1 = t1.get_root(), t1.leaves_set()
    path1 = t1.get_merkle_proof(123)
    t1.set_metadata(b"test metadata")
    t1.flush()
    del t1
    t2 = z.RLN.new(10, str(cfg))
    assert t2.get_root() == root1 and t2.leaves_set() == n1 and t2.get_metadata() == b"test metadata"
    assert t2.get_leaf(0) == 42 and t2.get_leaf(5) == leaves[4] and t2.get_leaf(7) == 0 and t2.get_merkle_proof(123) == path1
    # after a reload the "is set" flags are recomputed from the leaves (pm_tree_adapter.rs:224-233): the empty ones are exactly the zero leaves
    assert t2.get_empty_leaves_indices() == [i for i in range(n1) if t2.get_leaf(i) == 0]
    exp = [t2.get_leaf(i) for i in range(n1)]
    assert root1 == int.from_bytes(oracle.merkle_build(10, fr_bytes(exp), 0, n1)[:32], "little")
    # second generation: more writes, dropped WITHOUT an explicit flush (sled flushes when the tree is dropped)
    t2.set_next_leaf(77)
    root2 = t2.get_root()
    del t2
    t3 = z.RLN.new(10, str(cfg))
    assert t3.get_root() == root2 and t3.leaves_set() == n1 + 1 and t3.get_leaf(n1) == 77 and t3.get_metadata() == b"test metadata"
    # set_tree replaces the tree by a temporary default one (rln/src/public.rs:298-303): the store keeps the old state
    t3.set_tree(10)
    t3.set_next_leaf(5)
    del t3
    t4 = z.RLN.new(10, str(cfg))
    assert t4.get_root() == root2
    del t4
    # configuration rules (resolve_path, pm_tree_adapter.rs:93-100, and the depth checks :194-208)
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"temporary": False}))
    with pytest.raises(z.RLNError, match="Configuration error: Error while creating pmtree config: missing path"):
        z.RLN.new(10, str(bad))
    bad.write_text(json.dumps({"path": str(db), "temporary": True}))
    with pytest.raises(z.RLNError, match="path already exists"):
        z.RLN.new(10, str(bad))
    bad.write_text(json.dumps({"path": str(db), "temporary": False, "tree_depth": 12}))
    with pytest.raises(z.RLNError, match="Tree depth"):
        z.RLN.new(10, str(bad))
    bad.write_text('{"path": ')
    with pytest.raises(z.RLNError, match="Configuration error: Error while reading pmtree config"):
        z.RLN.new(10, str(bad))
    # a damaged store is refused, not silently replaced
    f = db / "rlnb200_tree.bin"
    raw = bytearray(f.read_bytes())
    raw[60] ^= 1
    f.write_bytes(bytes(raw))
    with pytest.raises(z.RLNError, match="Cannot load database: checksum mismatch"):
        z.RLN.new(10, str(cfg))
    # a missing config file means the default (temporary) configuration (ffi_rln.rs:28-45)
    t5 = z.RLN.new(10, str(tmp_path / "nope.json"))
    assert t5.leaves_set() == 0
