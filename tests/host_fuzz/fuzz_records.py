# TEST HARNESS: CPU fuzz of every byte-record parser of the C ABI (witness, partial witness, proof values, proofs, Vec helpers; V1 and
# V3, LE / BE / mixed).  Mutated valid records must give the reference's error string or a clean parse, never a crash.  Run by
# tests/test_host_fuzz.py in a subprocess; RLN_B200_LIB + LD_PRELOAD=libasan.so runs it against an instrumented build (profiles/README).
import os, random, sys, struct
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import zerokit_b200 as z
from common import R, fr_stream
from pyref import serialize as S

rnd = random.Random(int(os.environ.get("SEED", "1")))
fs = fr_stream(11)
def wa(depth):
    return dict(secret=next(fs), limit=100, mid=7, path=[next(fs) for _ in range(depth)], idx=[rnd.randrange(2) for _ in range(depth)], x=next(fs), en=next(fs))
a = wa(20)
mids, sel = [1, 2, 3, 0], [True, True, False, False]
root, en, x, y, nul = (next(fs) for _ in range(5))
ys, nulls = [next(fs) for _ in range(4)], [next(fs) for _ in range(4)]
seeds = {
  'w_le': S.witness_to_bytes(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"], be=False),
  'w_be': S.witness_to_bytes(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"], be=True),
  'wm_le': S.witness_to_bytes_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel, be=False),
  'wm_be': S.witness_to_bytes_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel, be=True),
  'pw_le': S.partial_witness_to_bytes(a["secret"], a["limit"], a["path"], a["idx"], be=False),
  'pw_be': S.partial_witness_to_bytes(a["secret"], a["limit"], a["path"], a["idx"], be=True),
  'pv_le': S.proof_values_to_bytes(root, en, x, y, nul),
  'pv_be': S.proof_values_to_bytes(root, en, x, y, nul, be=True),
  'pvm_le': S.proof_values_to_bytes_multi(root, en, x, ys, nulls, [1, 0, 1, 1]),
  'pvm_be': S.proof_values_to_bytes_multi(root, en, x, ys, nulls, [1, 0, 1, 1], be=True),
  'vf_le': S.vec_fr([1, 2, 3], False), 'vf_be': S.vec_fr([1, 2, 3], True),
  'vu_le': S.vec_u8(b'abcdef', False), 'vu_be': S.vec_u8(b'abcdef', True),
}
seeds['proof_le'] = b'\x00' + bytes(128) + seeds['pv_le']
seeds['proof_be'] = b'\x00' + bytes(128) + seeds['pv_be']
parsers = {
  'w_le': [z.RLNWitnessInput.from_bytes_le, z.WitnessV3.from_bytes_le], 'w_be': [z.RLNWitnessInput.from_bytes_be, z.WitnessV3.from_bytes_be],
  'wm_le': [z.RLNWitnessInput.from_bytes_le, z.WitnessV3.from_bytes_le], 'wm_be': [z.RLNWitnessInput.from_bytes_be, z.WitnessV3.from_bytes_be],
  'pw_le': [z.RLNPartialWitnessInput.from_bytes_le, z.PartialWitnessV3.from_bytes_le], 'pw_be': [z.RLNPartialWitnessInput.from_bytes_be, z.PartialWitnessV3.from_bytes_be],
  'pv_le': [z.proof_values_le_to_be, z.ProofValuesV3.from_bytes_le], 'pv_be': [z.proof_values_be_to_le, z.ProofValuesV3.from_bytes_be],
  'pvm_le': [z.proof_values_le_to_be, z.ProofValuesV3.from_bytes_le], 'pvm_be': [z.proof_values_be_to_le, z.ProofValuesV3.from_bytes_be],
  'vf_le': [lambda b: z.bytes_to_vec_fr(b, False)], 'vf_be': [lambda b: z.bytes_to_vec_fr(b, True)],
  'vu_le': [lambda b: z.bytes_to_vec_u8(b, False)], 'vu_be': [lambda b: z.bytes_to_vec_u8(b, True)],
  'proof_le': [z.RLNProof.from_bytes_le, z.ProofV3.from_bytes_le, z.ProofV3.from_bytes_mixed, z.RLNPartialProof.from_bytes_le, z.PartialProofV3.from_bytes_le],
  'proof_be': [z.RLNProof.from_bytes_be, z.ProofV3.from_bytes_mixed, z.RLNPartialProof.from_bytes_be],
}
def mutate(b):
    b = bytearray(b)
    k = rnd.randrange(8)
    if k == 0: return bytes(b[:rnd.randrange(len(b) + 1)])
    if k == 1: return bytes(b) + bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 40)))
    if k == 2:
        for _ in range(rnd.randrange(1, 6)): b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        return bytes(b)
    if k == 3:   # plant a huge / odd 8-byte length somewhere
        pos = rnd.randrange(max(1, len(b) - 8))
        v = rnd.choice([2**64 - 1, 2**63, 2**61 + 3, 2**32, 2**31, 0, 1, len(b), 2**64 // 32, 2**64 // 32 + 1, 2**59])
        b[pos:pos + 8] = struct.pack(rnd.choice(['<Q', '>Q']), v)
        return bytes(b)
    if k == 4: return bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 1200)))
    if k == 5: b[0] = rnd.randrange(256); return bytes(b)
    if k == 6: return bytes(b[rnd.randrange(len(b)):])
    return bytes(b) * 2
n_ok = n_err = 0
N = int(os.environ.get("N", "3000"))
for it in range(N):
    name = rnd.choice(list(seeds))
    data = mutate(seeds[name]) if it % 50 else seeds[name]
    for p in parsers[name]:
        try:
            p(data); n_ok += 1
        except z.RLNError:
            n_err += 1
print('fuzz ok', n_ok, 'parsed', n_err, 'rejected')
