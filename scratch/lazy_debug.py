# prints the lazy Fq2 self-test ops next to candidate expressions (debugging aid)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import zerokit_b200 as z
from common import fr_bytes, ints
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
xs = [0, Q - 1, 1, Q - 1, 2, 5, 7, 123456789]
ys = [0, Q - 1, Q - 1, 1, 3, 11, 13, 987654321]
A, B = fr_bytes(xs), fr_bytes(ys)
got = {op: ints(z.field_op(1, op, A, B, len(xs))) for op in (0, 9, 10, 11, 12)}
def sg(v): return v if v < Q // 2 else v - Q
for i, (x, y) in enumerate(zip(xs, ys)):
    a, b, c, d = x, y, y, (x + y) % Q
    cand = {'ac-bd': a * c - b * d, 'bd-ac': b * d - a * c, 'ad+bc': a * d + b * c, 'ac': a * c, 'bd': b * d, '(a+b)(c+d)': (a + b) * (c + d), 'ac+bd': a*c+b*d}
    line = [f"x={sg(x)} y={sg(y)}"] + [f"op{op}={sg(got[op][i])}" for op in (9, 10, 11, 12)]
    for op in (9, 10):
        m = [k for k, v in cand.items() if v % Q == got[op][i]]
        line.append(f"op{op}~{m}")
    print(' '.join(line), flush=True)
