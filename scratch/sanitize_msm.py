# compute-sanitizer workload for the variable-base MSM tail (tiered split-bucket combine, one-warp Horner pass): random scalars at
# 2^12 / 2^14 (nearly every bucket split into a few slices), skewed scalars (giant buckets: 32 lanes per bucket), tiny n
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
import json, random
import zerokit_b200 as z
from common import fr_bytes, ints
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
g = json.load(open(os.path.join(ROOT, "tests", "golden", "derived_vectors.json")))
ms = g['msm_g1_48']
pts = b''.join(fr_bytes([int(q[0]), int(q[1])]) for q in ms['bases'])
m = z.G1Msm(1 << 14)
assert [str(x) for x in ints(m.msm(pts, fr_bytes([int(s) for s in ms['scalars']]), 48))] == ms['result']
rnd = random.Random(3)
for lg in (12, 14):
    n = 1 << lg
    bases = pts * (n // 48) + pts[:64 * (n % 48)]
    r1 = m.msm(bases, fr_bytes([rnd.randrange(R) for _ in range(n)]), n)
    r2 = m.msm(bases, fr_bytes([5] * n), n)
    r3 = m.msm(bases, fr_bytes([(i % 3) * (R - 1) % R for i in range(n)]), n)
    assert len(r1) == 64 and len(r2) == 64 and len(r3) == 64
for n in (1, 2, 33):
    m.msm(pts[:64 * n], fr_bytes([rnd.randrange(R) for _ in range(n)]), n)
print("sanitize msm workload ok")
