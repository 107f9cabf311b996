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
# both verifier kernels on good, wrong, malformed and handed-back inputs (the lane-parallel program, its fallback, the thread kernel)
bad = bytearray(out[:290]); bad[32] ^= 0x80                      # A's sign bit: a valid encoding of the wrong point
inf = bytearray(out[:290]); inf[32] = (inf[32] & 0x3F) | 0x40    # A at infinity: the program hands the proof back
mal = bytearray(out[:290]); mal[32] |= 0xC0                      # both flags: refused
mix = out[:290] + bytes(bad) + bytes(inf) + bytes(mal)
assert rln.verify_batch(mix, 4) == [1, 0, 0, 2]
rln.set_verify_vm_max(0)
assert rln.verify_batch(mix, 4) == [1, 0, 0, 2]
rln.set_verify_vm_max(4096)
pp = rln.partial_batch(wb * 3, 3)
assert rln.finish_batch(wb * 3, 3, pp, fr_bytes([1, 2, 3, 4, 5, 6])) == out
rln.set_leaves_from(5, list(range(1, 40)))
rln.get_merkle_proofs([0, 5, 1023])
rln.atomic_operation(44, [7, 8], [6])
m = z.G1Msm(1 << 10)
ms = g['msm_g1_48']
pts = b''.join(fr_bytes([int(q[0]), int(q[1])]) for q in ms['bases'])
assert [str(x) for x in ints(m.msm(pts, fr_bytes([int(s) for s in ms['scalars']]), 48))] == ms['result']
# paths added later in the round: cooperative Poseidon levels (> 10 parents per level) and the fused top, single-leaf update,
# tree queries, sliced-bucket MSM with skewed scalars, V3 object, external witness, seeded keygen
rln.set_leaves_from(0, list(range(1, 700)))
rln.set_leaf(3, 99); rln.delete_leaf(4)
assert rln.get_subtree_root(0, 0) == rln.get_root() and rln.get_empty_leaves_indices() == [4]
import numpy as np
n = 1 << 10
bases = pts * (n // 48) + pts[:64 * (n % 48)]
for sc in ([5] * n, [i % 2 for i in range(n)]):
    m.msm(bases, fr_bytes(sc), n)
v3 = z.RLNV3.stateful("full", 10, resource(10, "rln_final.arkzkey"), resource(10, "graph.bin"))
args = kat_witness_args(10, k['inputs'])
w3 = z.WitnessV3.new_single(*args[:3], args[3], args[4], args[5], args[6])
p3 = v3.generate_proof_with_rs(w3, int(k['inputs']['r']), int(k['inputs']['s']))
assert p3.to_bytes_le()[:128].hex() == k['rln_proof_le_hex'][2:258] and v3.verify(p3, args[5])
from pyref import groth16 as G
gr = G.parse_graph(resource(10, "graph.bin"))
wires = G.evaluate(gr, G.inputs_buffer(gr, *args))
pw = rln.generate_rln_proof_with_witness(wires, z.RLNWitnessInput.from_bytes_le(wb), int(k['inputs']['r']), int(k['inputs']['s']))
assert pw.to_bytes_le().hex() == k['rln_proof_le_hex']
z.seeded_keygen(b"abc")
print('sanitize workload ok')
