# a handful of single verifications through the lane-parallel kernel (for ncu: -k regex:k_verify_vm)
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
import zerokit_b200 as z
from pyref import groth16 as G
g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_kats.json')))["groth16_verifier_single"]
proof = ((int(g["pi_a"][0]), int(g["pi_a"][1])), ((int(g["pi_b"][0][0]), int(g["pi_b"][0][1])), (int(g["pi_b"][1][0]), int(g["pi_b"][1][1]))),
         (int(g["pi_c"][0]), int(g["pi_c"][1])))
rec = G.rln_proof_to_bytes_le(proof, {k: int(g[k]) for k in ("root", "x", "external_nullifier", "y", "nullifier")})
rln = z.RLN.new(20)
for _ in range(4):
    assert rln.verify_batch(rec, 1) == [1]
print('ok')
