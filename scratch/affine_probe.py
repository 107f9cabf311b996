# batched-affine vs XYZZ bucket accumulation: adds/s for M running sums per thread in global memory
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zerokit_b200 import ffi
L = ffi.lib()
print('egcd self-check mismatches:', L.rlnb200_affine_batch_probe(3, 0, 0))
for M in (8, 16, 32, 64, 128):
    r = max(2, 512 // M)
    x = L.rlnb200_affine_batch_probe(0, M, r)
    f = L.rlnb200_affine_batch_probe(1, M, r)
    e = L.rlnb200_affine_batch_probe(2, M, r)
    print(f'M={M:4d}  XYZZ {x/1e9:7.3f} G adds/s   affine+Fermat {f/1e9:7.3f} ({f/x:4.2f}x)   affine+EGCD {e/1e9:7.3f} ({e/x:4.2f}x)', flush=True)
