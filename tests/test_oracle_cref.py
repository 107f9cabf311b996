"""CPU: the C++ oracle (oracle/cref → liboracle.so) against the golden fixtures."""
import hashlib

import pytest

from common import fr_bytes, ints, kat_witness_args, multi_kat, multi_resource, resource
from oracle import cref_binding as C


def test_poseidon_and_constants(goldens):
    pc = goldens["poseidon_constants"]
    for i, t in enumerate(range(2, 10)):
        ark, mds = C.poseidon_constants(t)
        assert [str(x) for x in ark] == pc["c"][i]
        assert [str(x) for x in mds] == sum(pc["m"][i], [])
    for k, v in goldens["ref"]["poseidon_single"]["cases"]:
        assert C.poseidon([int(k)]) == int(v)
    assert C.poseidon([1, 2, 3]) == int(goldens["derived"]["poseidon_misc"]["t4"])


def test_merkle(goldens):
    m = goldens["derived"]["merkle_d10"]
    nodes = C.merkle_build(10, fr_bytes([int(x) for x in m["leaves"]]), m["start"], len(m["leaves"]))
    assert int.from_bytes(nodes[:32], "little") == int(m["root"])
    for i, pr in m["proofs"].items():
        e, b = C.merkle_proof_from_nodes(nodes, 10, int(i))
        assert [str(x) for x in e] == pr["elements"] and b == pr["index"]


def test_ntt_and_msm(goldens):
    d = goldens["derived"]
    nt = d["ntt16"]
    assert [str(x) for x in C.ntt([int(x) for x in nt["input"]])] == nt["forward"]
    assert [str(x) for x in C.ntt([int(x) for x in nt["input"]], True)] == nt["inverse"]
    ms = d["msm_g1_48"]
    pts = b"".join(fr_bytes([int(p[0]), int(p[1])]) for p in ms["bases"])
    r = ints(C.msm_g1(pts, fr_bytes([int(s) for s in ms["scalars"]]), 48))
    assert [str(x) for x in r] == ms["result"]
    ms = d["msm_g2_12"]
    pts = b"".join(fr_bytes([int(p[0][0]), int(p[0][1]), int(p[1][0]), int(p[1][1])]) for p in ms["bases"])
    r = ints(C.msm_g2(pts, fr_bytes([int(s) for s in ms["scalars"]]), 12))
    assert [[str(r[0]), str(r[1])], [str(r[2]), str(r[3])]] == ms["result"]


@pytest.mark.parametrize("depth,key", [(20, "kat_proof_d20"), (10, "kat_proof_d10"), (20, "kat_proof_d20_r0")])
def test_known_answer_proofs(goldens, depth, key):
    k = goldens["derived"][key]
    ctx = C.Ctx(resource(depth, "rln_final.arkzkey"), resource(depth, "graph.bin"))
    ib = ctx.inputs_buffer(*kat_witness_args(depth, k["inputs"]))
    w = ctx.witness(ib)
    assert hashlib.sha256(w).hexdigest() == k["w_sha256"]
    assert hashlib.sha256(ctx.qap_h(w)).hexdigest() == k["h_sha256"]
    pr, pub = ctx.prove_batch(ib, fr_bytes([int(k["inputs"]["r"]), int(k["inputs"]["s"])]), 1)
    v = ints(pr)
    assert [str(v[0]), str(v[1])] == k["A"] and [str(v[6]), str(v[7])] == k["C"]
    assert [[str(v[2]), str(v[3])], [str(v[4]), str(v[5])]] == k["B"]
    assert ctx.verify_batch(pr, pub, 1) == [1]
    bad = bytearray(pub)
    bad[0] ^= 1
    assert ctx.verify_batch(pr, bytes(bad), 1) == [0]


def test_reference_snarkjs_proof(goldens):
    v = goldens["ref"]["groth16_verifier_single"]
    ctx = C.Ctx(resource(20, "rln_final.arkzkey"), resource(20, "graph.bin"))
    pr = fr_bytes([int(x) for x in (v["pi_a"] + v["pi_b"][0] + v["pi_b"][1] + v["pi_c"])])
    pub = fr_bytes([int(v[n]) for n in ("y", "root", "nullifier", "x", "external_nullifier")])
    assert ctx.verify_batch(pr, pub, 1) == [1]


def test_reference_snarkjs_proof_multi(goldens):
    """rln/tests/public.rs:143-233 (max_out = 4 circuit)"""
    ctx = C.Ctx(multi_resource("rln_final.arkzkey"), multi_resource("graph.bin"))
    proof, pub = multi_kat(goldens["ref"]["groth16_verifier_multi"])
    assert ctx.num_public == 15
    assert ctx.verify_batch(fr_bytes(proof), fr_bytes(pub), 1) == [1]
    bad = list(pub)
    bad[4] += 1
    assert ctx.verify_batch(fr_bytes(proof), fr_bytes(bad), 1) == [0]


def test_fr_dot_and_mul_gen():
    """the checker of the large MSM tests: Σ k_i·s_i mod r, threaded == Python integers; (Σ k_i s_i)·G == MSM over k_i·G"""
    import random
    from common import R
    rnd = random.Random(9)
    n = 300
    ks, ss = [rnd.randrange(R) for _ in range(n)], [rnd.randrange(R) for _ in range(n)]
    want = sum(k * s for k, s in zip(ks, ss)) % R
    for t in (1, 4):
        assert int.from_bytes(C.fr_dot(fr_bytes(ks), fr_bytes(ss), n, t), "little") == want
    bases = C.g1_mul_gen(fr_bytes(ks), n, 2)
    assert C.msm_g1(bases, fr_bytes(ss), n, 2) == C.g1_mul_gen(fr_bytes([want]), 1)
