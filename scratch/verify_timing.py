# single verification latency and batch verification throughput (rlnb200_verify_batch) after the verifier changes
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault('RLN_B200_WINDOW_BITS', '8')
import numpy as np, torch
import zerokit_b200 as z
sys.argv = ['bench']
import bench
rln = z.RLN.new(20)
n = 65536
rec, rs, root = bench.make_witnesses(rln, 4096, 5)
out = rln.prove_batch(rec.tobytes(), 4096, rs.tobytes())
assert rln.verify_batch(out, 4096) == [1] * 4096
p = z.RLNProof.from_bytes_le(out[:290])
for _ in range(3): rln.verify_with_roots(p, p.values.x, [])
t0 = time.perf_counter()
for _ in range(20): assert rln.verify_with_roots(p, p.values.x, [])
print('single verify ms', (time.perf_counter() - t0) / 20 * 1e3)
big = out * 16
for nn in (4096, 65536):
    rln.verify_batch(big[:290 * nn], nn)
    t0 = time.perf_counter()
    ok = rln.verify_batch(big[:290 * nn], nn)
    dt = time.perf_counter() - t0
    assert ok == [1] * nn
    print('verify_batch', nn, 'proofs', round(dt * 1e3, 2), 'ms', round(nn / dt), 'proofs/s')
bad = bytearray(out[:290 * 64])
for j in range(0, 64, 2): bad[290 * j + 200] ^= 1
assert rln.verify_batch(bytes(bad), 64) == [0, 1] * 32
print('forged proofs rejected, untouched ones accepted')
