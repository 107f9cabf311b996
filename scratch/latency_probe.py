# cycles per dependent Fq operation in a lone warp (why latency-bound kernels get their own multiplier)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zerokit_b200 import ffi
L = ffi.lib()
names = {0: 'mul_ptx', 1: 'mul_portable (CIOS in C)', 3: 'sqr_ptx', 4: 'modular addition'}
for lanes in (1, 32):
    for k in (0, 1, 3, 4):
        print(f'lanes={lanes:2d}  {names[k]:28s} {L.rlnb200_latency_probe(k, lanes, 20000):8.1f} cycles / op', flush=True)
