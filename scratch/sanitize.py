# small end-to-end run for compute-sanitizer: depth-10 circuit, c=6 tables, 3 proofs, tree ops, MSM 2^10, verify
import os, sys
os.environ['RLN_B200_WINDOW_BITS'] = '6'; os.environ['RLN_B200_WINDOW_BITS_G2'] = '6'
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import json
import zerokit_b200 as z
from common import *
g = json.load(open('/root/repo/tests/golden/derived_vectors.json'))
k = g['kat_proof_d10']
rln = z.RLN.new(10)
wb = witness_le(*kat_witness_args(10, k['inputs']))
p = rln.generate_rln_proof_with_rs(z.RLNWitnessInput.from_bytes_le(wb), int(k['inputs']['r']), int(k['inputs']['s']))
assert p.to_bytes_le().hex() == k['rln_proof_le_hex']
out = rln.prove_batch(wb * 3, 3, fr_bytes([1, 2, 3, 4, 5, 6]))
assert rln.verify_batch(out, 3) == [1, 1, 1]
pp = rln.partial_batch(wb * 3, 3)
assert rln.finish_batch(wb * 3, 3, pp, fr_bytes([1, 2, 3, 4, 5, 6])) == out
rln.set_leaves_from(5, list(range(1, 40)))
rln.get_merkle_proofs([0, 5, 1023])
rln.atomic_operation(44, [7, 8], [6])
m = z.G1Msm(1 << 10)
ms = g['msm_g1_48']
pts = b''.join(fr_bytes([int(q[0]), int(q[1])]) for q in ms['bases'])
assert [str(x) for x in ints(m.msm(pts, fr_bytes([int(s) for s in ms['scalars']]), 48))] == ms['result']
print('sanitize workload ok')
