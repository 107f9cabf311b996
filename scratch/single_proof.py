# one proof through ffi_generate_rln_proof + one verification (for ncu: the single-proof launch list / captures)
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import zerokit_b200 as z
from common import kat_witness_args, witness_le
g = json.load(open(os.path.join(ROOT, 'tests/golden/derived_vectors.json')))
k = g['kat_proof_d20']
rln = z.RLN.new(20)
w = z.RLNWitnessInput.from_bytes_le(witness_le(*kat_witness_args(20, k['inputs'])))
for _ in range(2):
    p = rln.generate_rln_proof_with_rs(w, 44, 77)
assert p.to_bytes_le().hex() == k['rln_proof_le_hex']
assert rln.verify_with_roots(p, p.values.x, [])
print('ok')
