#!/usr/bin/env python3
"""Regenerates the JSON fixtures in tests/golden/.

Run in the BUILD container only (needs /root/reference for the part that lifts the reference's
own golden vectors out of its Rust test files).  The GPU box never runs this: tests read the
committed JSON.

  reference_kats.json        vectors copied out of the reference tests (file:line in each entry)
  poseidon_constants.json    every round constant + MDS entry for t=2..9
                             (utils/tests/poseidon_constants.rs:42-3498)
  derived_vectors.json       vectors produced by oracle/pyref (the Python-integer restatement,
                             itself pinned by the two files above): a known-answer proof, MSM /
                             NTT / Merkle samples for the C++ oracle and the CUDA path.
"""
import hashlib
import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = "/root/reference"
RES = os.path.join(ROOT, "zerokit_b200", "resources")

from pyref import fields as F  # noqa: E402
from pyref import groth16 as G  # noqa: E402
from pyref import poseidon as P  # noqa: E402


def lift_constants():
    src = open(os.path.join(REF, "utils/tests/poseidon_constants.rs")).read()
    c_part = src[src.index("let c_str"):src.index("let m_str")]
    m_part = src[src.index("let m_str"):src.index("fn load_constants")] if "fn load_constants" in src else \
        src[src.index("let m_str"):]
    m_part = m_part[:m_part.index("(c_str, m_str)")] if "(c_str, m_str)" in m_part else m_part

    def groups(txt, depth):
        """split nested vec![ ... ] at the requested depth into lists of decimal strings"""
        out, stack, cur = [], 0, None
        for tok in re.finditer(r"vec!\[|\]|\"(\d+)\"", txt):
            t = tok.group(0)
            if t == "vec![":
                stack += 1
                if stack == depth:
                    cur = []
            elif t == "]":
                if stack == depth:
                    out.append(cur)
                    cur = None
                stack -= 1
            elif cur is not None:
                cur.append(tok.group(1))
        return out

    c = groups(c_part, 2)
    m_rows = groups(m_part, 3)
    ts = list(range(2, 10))
    assert len(c) == 8, len(c)
    m, k = [], 0
    for t in ts:
        m.append(m_rows[k:k + t])
        k += t
    assert k == len(m_rows), (k, len(m_rows))
    return {"source": "utils/tests/poseidon_constants.rs:42-3498",
            "round_params": P.ROUND_PARAMS, "c": c, "m": m}


def reference_kats():
    return {
        "poseidon_single": {
            "source": "utils/tests/poseidon_hash_test.rs:21-66",
            "cases": [
                ["0", "19014214495641488759237505126948346942972912379615652741039992445865937985820"],
                ["1", "18586133768512220936620570745912940619677854269274689475585506675881198879027"],
                ["255", "20026131459732984724454933360292530547665726761019872861025481903072111625788"],
                ["65535", "12358868638722666642632413418981275677998688723398440898957566982787708451243"],
                ["18446744073709551615", "17449307747295017006142981453320720946812828330895590310359634430146721583189"],
            ]},
        "poseidon_pair_tree8": {
            "source": "utils/tests/poseidon_hash_test.rs:69-130 (leaves 0..7)",
            "l01": "12583541437132735734108669866114103169564651237895298778035846191048104863326",
            "l23": "17197790661637433027297685226742709599380837544520340689137581733613433332983",
            "l45": "756592041685769348226045093946546956867261766023639881791475046640232555043",
            "l67": "5558359459771725727593826278265342308584225092343962757289948761260561575479",
            "l03": "3720616653028013822312861221679392249031832781774563366107458835261883914924",
            "l47": "7960741062684589801276390367952372418815534638314682948141519164356522829957",
            "root": "11780650233517635876913804110234352847867393797952240856403268682492028497284"},
        "tree_depth20_leaf3": {
            "source": "rln/tests/protocol.rs:14-88",
            "secret_preimage": "test-merkle-proof", "user_message_limit": 100, "leaf_index": 3,
            "root_limbs_le64": [4939322235247991215, 5110804094006647505, 4427606543677101242, 910933464535675827],
            "path_elements": [
                "0x0000000000000000000000000000000000000000000000000000000000000000",
                "0x2098f5fb9e239eab3ceac3f27b81e481dc3124d55ffed523a839ee8446b64864",
                "0x1069673dcdb12263df301a6ff584a7ec261a44cb9dc68df067a4774460b1f1e1",
                "0x18f43331537ee2af2e3d758d50f72106467c6eea50371dd528d57eb2b856d238",
                "0x07f9d837cb17b0d36320ffe93ba52345f1b728571a568265caac97559dbc952a",
                "0x2b94cf5e8746b3f5c9631f4c5df32907a699c58c94b2ad4d7b5cec1639183f55",
                "0x2dee93c5a666459646ea7d22cca9e1bcfed71e6951b953611d11dda32ea09d78",
                "0x078295e5a22b84e982cf601eb639597b8b0515a88cb5ac7fa8a4aabe3c87349d",
                "0x2fa5e5f18f6027a6501bec864564472a616b2e274a41211a444cbe3a99f3cc61",
                "0x0e884376d0d8fd21ecb780389e941f66e45e7acce3e228ab3e2156a614fcd747",
                "0x1b7201da72494f1e28717ad1a52eb469f95892f957713533de6175e5da190af2",
                "0x1f8d8822725e36385200c0b201249819a6e6e1e4650808b5bebc6bface7d7636",
                "0x2c5d82f66c914bafb9701589ba8cfcfb6162b0a12acf88a8d0879a0471b5f85a",
                "0x14c54148a0940bb820957f5adf3fa1134ef5c4aaa113f4646458f270e0bfbfd0",
                "0x190d33b12f986f961e10c0ee44d8b9af11be25588cad89d416118e4bf4ebe80c",
                "0x22f98aa9ce704152ac17354914ad73ed1167ae6596af510aa5b3649325e06c92",
                "0x2a7c7c9b6ce5880b9f6f228d72bf6a575a526f29c66ecceef8b753d38bba7323",
                "0x2e8186e558698ec1c67af9c14d463ffc470043c9c2988b954d75dd643f36b992",
                "0x0f57c5571e9a4eab49e2c8cf050dae948aef6ead647392273546249d1c1ff10f",
                "0x1830ee67b5fb554ad5f63d4388800e1cfe78e310697d46e43c9ce36134f72cca"],
            "identity_path_index": [1, 1] + [0] * 18},
        "groth16_verifier_single": {
            "source": "rln/tests/public.rs:77-142,214-233 (snarkjs proof must verify under the bundled depth-20 vk)",
            "pi_a": ["606446415626469993821291758185575230335423926365686267140465300918089871829",
                     "14881534001609371078663128199084130129622943308489025453376548677995646280161"],
            "pi_b": [["18053812507994813734583839134426913715767914942522332114506614735770984570178",
                      "11219916332635123001710279198522635266707985651975761715977705052386984005181"],
                     ["17371289494006920912949790045699521359436706797224428511776122168520286372970",
                      "14038575727257298083893642903204723310279435927688342924358714639926373603890"]],
            "pi_c": ["17701377127561410274754535747274973758826089226897242202671882899370780845888",
                     "12608543716397255084418384146504333522628400182843246910626782513289789807030"],
            "root": "8502402278351299594663821509741133196466235670407051417832304486953898514733",
            "x": "20645213238265527935869146898028115621427162613172918400241870500502509785943",
            "external_nullifier": "21074405743803627666274838159589343934394162804826017440941339048886754734203",
            "y": "16401008481486069296141645075505218976370369489687327284155463920202585288271",
            "nullifier": "9102791780887227194595604713537772536258726662792598131262022534710887343694"},
        "empty_tree_depth20_root": {
            "source": "SURVEY.md Appendix A.4 (zero-subtree chain z_{k+1} = Poseidon(z_k, z_k))",
            "root": "0x2134e76ac5d21aab186c2be1dd8f84ee880a1e46eaf712f9d371b6df22191f3e"},
    }


def splitmix64(seed):
    s = seed & (2 ** 64 - 1)
    while True:
        s = (s + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        yield z ^ (z >> 31)


def fr_stream(seed):
    """uniform Fr by rejection on the top 254 bits"""
    g = splitmix64(seed)
    while True:
        v = 0
        for i in range(4):
            v |= next(g) << (64 * i)
        v &= (1 << 254) - 1
        if v < F.R:
            yield v


def kat_proof(depth, secret, limit, mid, x, en, r, s, label):
    z = G.parse_zkey(open(os.path.join(RES, f"tree_depth_{depth}", "rln_final.arkzkey"), "rb").read())
    g = G.parse_graph(open(os.path.join(RES, f"tree_depth_{depth}", "graph.bin"), "rb").read())
    pe = [P.poseidon([i + 7]) for i in range(depth)]
    idx = [(5 * i + 1) % 2 for i in range(depth)]
    args = (secret, limit, mid, pe, idx, x, en)
    w = G.evaluate(g, G.inputs_buffer(g, *args))
    pv = P.proof_values_from_witness(*args)
    assert w[0] == 1 and w[1:6] == G.public_inputs_single(pv)
    h = G.witness_map(z, w)
    pr = G.prove(z, w, h, r, s)
    assert G.verify(z, pr, w[1:6]), "derived proof must verify under the pinned verifier"
    bad = list(w[1:6])
    bad[0] = (bad[0] + 1) % F.R
    assert not G.verify(z, pr, bad)
    a, b, c = pr
    return {
        "label": label, "tree_depth": depth,
        "inputs": {"identity_secret": str(secret), "user_message_limit": str(limit), "message_id": str(mid),
                   "x": str(x), "external_nullifier": str(en),
                   "path_elements": "Poseidon([i+7]) for i in range(depth)",
                   "identity_path_index": "(5*i+1)%2", "r": str(r), "s": str(s)},
        "public": {k: str(v) for k, v in pv.items()},
        "w_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in w)).hexdigest(),
        "h_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in h)).hexdigest(),
        "A": [str(a[0]), str(a[1])],
        "B": [[str(b[0][0]), str(b[0][1])], [str(b[1][0]), str(b[1][1])]],
        "C": [str(c[0]), str(c[1])],
        "proof_bytes_hex": G.proof_to_bytes(pr).hex(),
        "rln_proof_le_hex": G.rln_proof_to_bytes_le(pr, pv).hex(),
        "witness_le_hex": G.witness_to_bytes_le(*args).hex(),
    }


def derived():
    out = {"generator": "tests/golden/make_goldens.py using oracle/pyref"}
    # A.4 of SURVEY.md (depth 20, r=44, s=77 as in rln/tests/protocol.rs:234-235)
    out["kat_proof_d20"] = kat_proof(20, 123456789, 100, 1, 42, 100, 44, 77, "SURVEY A.4")
    fs = fr_stream(6)
    out["kat_proof_d10"] = kat_proof(10, next(fs), 1000, 7, next(fs), next(fs), next(fs), next(fs), "depth-10 random")
    out["kat_proof_d20_r0"] = kat_proof(20, 987654321, 5, 4, 99, 3, 0, 5, "r = 0 (g1_b skipped, partial_proof.rs:242-248)")

    # multi message-id circuit (max_out = 4), bundled with the reference (rln/src/circuit/mod.rs:36-42)
    mdir = os.path.join(RES, "tree_depth_20", "multi_message_id", "max_out_4")
    zm = G.parse_zkey(open(os.path.join(mdir, "rln_final.arkzkey"), "rb").read())
    gm = G.parse_graph(open(os.path.join(mdir, "graph.bin"), "rb").read())
    pe20 = [P.poseidon([i + 7]) for i in range(20)]
    idx20 = [(5 * i + 1) % 2 for i in range(20)]
    margs = (424242, 50, [3, 7, 11, 0], pe20, idx20, 1234567, 89, [True, False, True, False])
    wm = G.evaluate(gm, G.inputs_buffer(gm, *margs))
    pvm = P.proof_values_from_witness_multi(*margs)
    assert wm[0] == 1 and wm[1:zm.num_instance] == G.public_inputs_multi(pvm), "multi public outputs"
    hm = G.witness_map(zm, wm)
    prm = G.prove(zm, wm, hm, 44, 77)
    assert G.verify(zm, prm, wm[1:zm.num_instance])
    out["kat_proof_multi_d20"] = {
        "inputs": {"identity_secret": "424242", "user_message_limit": "50", "message_ids": ["3", "7", "11", "0"], "x": "1234567",
                   "external_nullifier": "89", "selector_used": [True, False, True, False], "r": "44", "s": "77",
                   "path_elements": "Poseidon([i+7])", "identity_path_index": "(5*i+1)%2"},
        "public": {"root": str(pvm["root"]), "ys": [str(v) for v in pvm["ys"]], "nullifiers": [str(v) for v in pvm["nullifiers"]]},
        "w_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in wm)).hexdigest(),
        "rln_proof_le_hex": G.rln_proof_to_bytes_le(prm, pvm).hex(),
        "witness_le_hex": G.witness_to_bytes_le_multi(*margs).hex(),
    }
    # partial proof (rln/src/partial_proof.rs:108-274; rln/tests/protocol.rs:222-248 pins full == partial + finish)
    k = out["kat_proof_d10"]
    z = G.parse_zkey(open(os.path.join(RES, "tree_depth_10", "rln_final.arkzkey"), "rb").read())
    g = G.parse_graph(open(os.path.join(RES, "tree_depth_10", "graph.bin"), "rb").read())
    inp = k["inputs"]
    pe = [P.poseidon([i + 7]) for i in range(10)]
    idx = [(5 * i + 1) % 2 for i in range(10)]
    args = (int(inp["identity_secret"]), int(inp["user_message_limit"]), int(inp["message_id"]), pe, idx, int(inp["x"]),
            int(inp["external_nullifier"]))
    w = G.evaluate(g, G.inputs_buffer(g, *args))
    # the known wires must not depend on the unknown inputs: evaluate with those zeroed and compare
    w0 = G.evaluate(g, G.inputs_buffer(g, args[0], args[1], 0, pe, idx, 0, 0))
    km = G.known_wire_mask(g)
    assert all(a == b for a, b, m in zip(w, w0, km) if m)
    partial = G.prove_partial(z, g, w0)
    h = G.witness_map(z, w)
    fin = G.finish_partial(z, partial, w, h, int(inp["r"]), int(inp["s"]))
    assert G.proof_to_bytes(fin).hex() == k["proof_bytes_hex"], "full proof != partial + finish"
    out["partial_proof_d10"] = {"known_wires": sum(km), "unknown_wires": len(km) - sum(km),
                                "partial_le_hex": G.partial_proof_to_bytes_le(partial).hex(),
                                "finished_proof_bytes_hex": G.proof_to_bytes(fin).hex()}
    # G1 / G2 MSM samples
    rnd = random.Random(11)
    fs = fr_stream(2)
    ks = [rnd.randrange(1, F.R) for _ in range(48)]
    pts = [F.pt_mul(F.OPS1, F.G1_GEN, k) for k in ks]
    sc = [next(fs) for _ in range(44)] + [0, 1, F.R - 1, 2]
    res = F.msm(F.OPS1, pts, sc)
    out["msm_g1_48"] = {"bases": [[str(p[0]), str(p[1])] for p in pts], "scalars": [str(s) for s in sc],
                        "result": [str(res[0]), str(res[1])]}
    pts2 = [F.pt_mul(F.OPS2, F.G2_GEN, k) for k in ks[:12]]
    res2 = F.msm(F.OPS2, pts2, sc[:12])
    out["msm_g2_12"] = {"bases": [[[str(p[0][0]), str(p[0][1])], [str(p[1][0]), str(p[1][1])]] for p in pts2],
                        "scalars": [str(s) for s in sc[:12]],
                        "result": [[str(res2[0][0]), str(res2[0][1])], [str(res2[1][0]), str(res2[1][1])]]}
    # NTT sample (size 16): forward with ω16 and the coset map used by qap.rs:69-90
    v = [next(fs) for _ in range(16)]
    om = G.root_of_unity(16)
    out["ntt16"] = {"input": [str(x) for x in v], "omega": str(om),
                    "forward": [str(x) for x in G.ntt(v, om)],
                    "inverse": [str(x) for x in G.intt(v, om)]}
    # Merkle: depth-10 tree with 37 seeded leaves starting at 5, a few proofs
    fs = fr_stream(3)
    tr = P.FullMerkleTree(10)
    leaves = [next(fs) for _ in range(37)]
    tr.set_range(5, leaves)
    prs = {}
    for i in (0, 5, 6, 41, 1023):
        e, b = tr.proof(i)
        prs[str(i)] = {"elements": [str(x) for x in e], "index": b}
    out["merkle_d10"] = {"start": 5, "leaves": [str(x) for x in leaves], "root": str(tr.root()), "proofs": prs}
    # hash_to_field / keccak
    out["hash_to_field"] = {m: str(P.hash_to_field_le(m.encode())) for m in
                            ["", "test-merkle-proof", "test-epoch", "test-rln-identifier", "a" * 200]}
    out["poseidon_misc"] = {
        "t4": str(P.poseidon([1, 2, 3])), "t2_rm1": str(P.poseidon([F.R - 1])),
        "t3_big": str(P.poseidon([F.R - 1, F.R - 2]))}
    return out


def main():
    consts = lift_constants()
    # the python restatement must reproduce every constant before anything derived is trusted
    for (t, rf, rp, skip), c, m in zip(P.ROUND_PARAMS, consts["c"], consts["m"]):
        ark, mds = P.find_ark_and_mds(t, rf, rp, skip)
        assert [str(v) for v in ark] == c, f"round constants t={t}"
        assert [[str(v) for v in row] for row in mds] == m, f"mds t={t}"
    json.dump(consts, open(os.path.join(HERE, "poseidon_constants.json"), "w"))
    json.dump(reference_kats(), open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)
    json.dump(derived(), open(os.path.join(HERE, "derived_vectors.json"), "w"), indent=1)
    print("fixtures written")


if __name__ == "__main__":
    main()
