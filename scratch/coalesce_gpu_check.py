# GPU check of request coalescing (RLN_B200_COALESCE=1): 16 threads prove and verify single items on ONE handle through the
# reference's single-item calls; every proof must equal the known-answer proof for its (r, s), every verification must pass, a
# forged proof must fail only for its own caller.  Prints the throughput with and without coalescing (run twice, env differs).
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault('RLN_B200_WINDOW_BITS', '8')
import zerokit_b200 as z
from common import *
g = json.load(open(os.path.join(ROOT, 'tests/golden/derived_vectors.json')))
k = g['kat_proof_d10']
rln = z.RLN.new(10)
wb = witness_le(*kat_witness_args(10, k['inputs']))
w = z.RLNWitnessInput.from_bytes_le(wb)
r, s = int(k['inputs']['r']), int(k['inputs']['s'])
gold = bytes.fromhex(k['rln_proof_le_hex'])
x = int(k['inputs']['x'])
T, PER = 16, 8
errs = []
def worker(t):
    try:
        for i in range(PER):
            p = rln.generate_rln_proof_with_rs(w, r, s)
            assert p.to_bytes_le() == gold, 'proof differs from the known answer'
            q = rln.generate_rln_proof(w)
            assert rln.verify_with_roots(q, x, []), 'fresh proof does not verify'
            if t == 3 and i == 2:   # a forged record: only this caller may see the failure
                bad = bytearray(gold); bad[140] ^= 1
                try:
                    rln.verify_with_roots(z.RLNProof.from_bytes_le(bytes(bad)), x, [])
                    raise AssertionError('forged proof accepted')
                except z.RLNError as e:
                    assert 'Invalid proof' in str(e), str(e)
    except Exception as e:
        errs.append((t, repr(e)))
t0 = time.time()
ts = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
[t.start() for t in ts]; [t.join() for t in ts]
dt = time.time() - t0
assert not errs, errs
print('coalesce=%s: %d threads x %d (2 proofs + 1 verification each): %.2f s, %.0f proofs/s' %
      (os.environ.get('RLN_B200_COALESCE', '0'), T, PER, dt, 2 * T * PER / dt))
