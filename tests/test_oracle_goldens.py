"""CPU: the Python-integer oracle (oracle/pyref) against every golden vector the reference's own tests
hold for the path.  This is what pins the oracle."""
import hashlib

from pyref import fields as F
from pyref import groth16 as G
from pyref import poseidon as P

from common import dense_tree_quirk_script, multi_kat, multi_resource, pmtree_quirk_script, resource


def test_poseidon_constants_all_widths(goldens):
    """utils/tests/poseidon_constants.rs:42-3520 — every round constant and MDS entry, t = 2..9"""
    pc = goldens["poseidon_constants"]
    for (t, rf, rp, skip), c, m in zip(P.ROUND_PARAMS, pc["c"], pc["m"]):
        ark, mds = P.find_ark_and_mds(t, rf, rp, skip)
        assert [str(v) for v in ark] == c
        assert [[str(v) for v in row] for row in mds] == m


def test_poseidon_hash_kats(goldens):
    """utils/tests/poseidon_hash_test.rs:21-130"""
    for k, v in goldens["ref"]["poseidon_single"]["cases"]:
        assert P.poseidon([int(k)]) == int(v)
    t = goldens["ref"]["poseidon_pair_tree8"]
    l = [P.poseidon([2 * i, 2 * i + 1]) for i in range(4)]
    assert [str(x) for x in l] == [t["l01"], t["l23"], t["l45"], t["l67"]]
    l03, l47 = P.poseidon(l[:2]), P.poseidon(l[2:])
    assert (str(l03), str(l47)) == (t["l03"], t["l47"])
    assert str(P.poseidon([l03, l47])) == t["root"]


def test_tree_depth20_kat(goldens):
    """rln/tests/protocol.rs:14-88 (also pins hash_to_field_le = Keccak-256 mod r)"""
    k = goldens["ref"]["tree_depth20_leaf3"]
    secret = P.hash_to_field_le(k["secret_preimage"].encode())
    assert secret == P.hash_to_field_be(k["secret_preimage"].encode())
    leaf = P.poseidon([P.poseidon([secret]), k["user_message_limit"]])
    tr = P.FullMerkleTree(20)
    assert tr.root() == int(goldens["ref"]["empty_tree_depth20_root"]["root"], 16)
    tr.set(k["leaf_index"], leaf)
    root = sum(l << (64 * i) for i, l in enumerate(k["root_limbs_le64"]))
    assert tr.root() == root
    elems, bits = tr.proof(k["leaf_index"])
    assert elems == [int(x, 16) for x in k["path_elements"]]
    assert bits == k["identity_path_index"]
    assert P.compute_tree_root(secret, k["user_message_limit"], elems, bits) == root


def test_reference_snarkjs_proof_verifies(goldens):
    """rln/tests/public.rs:77-233 — hard-coded snarkjs proof must verify under the bundled vk"""
    v = goldens["ref"]["groth16_verifier_single"]
    z = G.parse_zkey(resource(20, "rln_final.arkzkey"))
    proof = ((int(v["pi_a"][0]), int(v["pi_a"][1])),
             ((int(v["pi_b"][0][0]), int(v["pi_b"][0][1])), (int(v["pi_b"][1][0]), int(v["pi_b"][1][1]))),
             (int(v["pi_c"][0]), int(v["pi_c"][1])))
    pub = [int(v[k]) for k in ("y", "root", "nullifier", "x", "external_nullifier")]
    assert G.verify(z, proof, pub)
    pub[0] += 1
    assert not G.verify(z, proof, pub)


def test_reference_snarkjs_proof_multi_verifies(goldens):
    """rln/tests/public.rs:143-233 — the hard-coded snarkjs proof of the multi-message-id circuit (max_out = 4)"""
    z = G.parse_zkey(multi_resource("rln_final.arkzkey"))
    c, pub = multi_kat(goldens["ref"]["groth16_verifier_multi"])
    proof = ((c[0], c[1]), ((c[2], c[3]), (c[4], c[5])), (c[6], c[7]))
    assert len(pub) == 15 and G.verify(z, proof, pub)
    for i in (0, 4, 9, 11):
        bad = list(pub)
        bad[i] = (bad[i] + 1) % F.R
        assert not G.verify(z, proof, bad)


def test_pmtree_override_range_quirks(goldens):
    """rln/tests/poseidon_tree.rs:79-146 replayed on the PmTree state model: pins override_range's next_index / cached
    flags behaviour (the expected index lists are the reference's own assertions)"""
    exp = goldens["ref"]["pmtree_override_range"]
    t = P.PmTree(exp["depth"])
    pmtree_quirk_script(t, exp)
    # what the last two steps leave in the tree: leaves below `start` keep their values, the new ones land shifted
    assert [t.get(i) for i in range(12)] == [0, 1, 0, 0, 0, 1, 2, 3, 0, 1, 2, 3] and t.leaves_set() == 12


def test_dense_tree_override_range(goldens):
    """utils/tests/merkle_tree.rs:222-312 replayed on the FullMerkleTree / OptimalMerkleTree state model"""
    exp = goldens["ref"]["dense_tree_override_range"]
    for optimal in (False, True):
        t = P.DenseTree(exp["depth"], optimal)
        dense_tree_quirk_script(t, exp)
        assert [t.get(i) for i in range(12)] == [0, 1, 0, 0, 0, 1, 2, 3, 0, 1, 2, 3] and t.leaves_set() == 12


def test_witness_and_h_hashes(goldens):
    """witness graph outputs == proof_values_from_witness; w / h digests"""
    k = goldens["derived"]["kat_proof_d20"]
    from common import kat_witness_args
    args = kat_witness_args(20, k["inputs"])
    g = G.parse_graph(resource(20, "graph.bin"))
    z = G.parse_zkey(resource(20, "rln_final.arkzkey"))
    w = G.evaluate(g, G.inputs_buffer(g, *args))
    pv = P.proof_values_from_witness(*args)
    assert w[0] == 1 and w[1:6] == G.public_inputs_single(pv)
    assert hashlib.sha256(b"".join(x.to_bytes(32, "little") for x in w)).hexdigest() == k["w_sha256"]
    h = G.witness_map(z, w)
    assert hashlib.sha256(b"".join(x.to_bytes(32, "little") for x in h)).hexdigest() == k["h_sha256"]


def test_compressed_encoding_roundtrip(goldens):
    k = goldens["derived"]["kat_proof_d20"]
    a = (int(k["A"][0]), int(k["A"][1]))
    assert G.g1_decompress(G.g1_compress(a)) == a
    assert G.g1_compress(a).hex() == k["proof_bytes_hex"][:64]
    assert F.on_curve(F.OPS1, a)
    # the decompressor the production-window GPU test feeds the oracle verifier with: golden proofs and random G2 points
    for key in ("kat_proof_d20", "kat_proof_d10", "kat_proof_d20_r0"):
        kk = goldens["derived"][key]
        pa, pb, pc = G.proof_from_bytes(bytes.fromhex(kk["proof_bytes_hex"]))
        assert [str(v) for v in pa] == kk["A"] and [str(v) for v in pc] == kk["C"]
        assert [[str(v) for v in pb[0]], [str(v) for v in pb[1]]] == kk["B"]
    import random
    rnd = random.Random(3)
    for _ in range(20):
        p = F.pt_mul(F.OPS2, F.G2_GEN, rnd.randrange(F.R))
        for q in (p, F.pt_neg(F.OPS2, p)):
            assert G.g2_decompress(G.g2_compress(q)) == q
    assert G.g2_decompress(G.g2_compress(F.INF)) is F.INF


def test_seeded_keygen_kats(goldens):
    """rln/tests/protocol.rs:459-507, rln/tests/ffi_utils.rs:8-69: pins the ChaCha20Rng + Fr::rand restatement"""
    from pyref import keygen as K
    k = goldens["ref"]["seeded_keygen"]
    assert K.seeded_keygen(k["phrase"]["seed_utf8"].encode()) == (int(k["phrase"]["identity_secret"], 16), int(k["phrase"]["id_commitment"], 16))
    assert K.seeded_keygen(bytes.fromhex(k["bytes"]["seed_hex"])) == (int(k["bytes"]["identity_secret"], 16), int(k["bytes"]["id_commitment"], 16))
    e = k["extended_bytes"]
    assert K.extended_seeded_keygen(bytes.fromhex(e["seed_hex"])) == tuple(
        int(e[f], 16) for f in ("identity_trapdoor", "identity_nullifier", "identity_secret", "id_commitment"))
