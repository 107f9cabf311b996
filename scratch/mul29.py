import sys, ctypes, random
sys.path.insert(0, '/root/repo')
from zerokit_b200 import ffi
L = ffi.lib()
L.rlnb200_mul29_check.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
L.rlnb200_mul29_throughput.restype = ctypes.c_double; L.rlnb200_mul29_throughput.argtypes = [ctypes.c_int]
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
n = 4096; rnd = random.Random(3)
def limbs(v): return [(v >> (29 * i)) & ((1 << 29) - 1) for i in range(8)] + [v >> 232]
vals = [(rnd.randrange(1 << 257), rnd.randrange(1 << 257)) for _ in range(n - 4)] + [(0, 0), (Q - 1, Q - 1), ((1 << 257) - 1, (1 << 257) - 1), (1, Q)]
A = (ctypes.c_uint32 * (9 * n))(*[w for a, b in vals for w in limbs(a)])
B = (ctypes.c_uint32 * (9 * n))(*[w for a, b in vals for w in limbs(b)])
O = (ctypes.c_uint32 * (9 * n))()
assert L.rlnb200_mul29_check(A, B, n, O) == 0
Rinv = pow(2, -261, Q); bad = 0; mx = 0
for i, (a, b) in enumerate(vals):
    r = sum(O[9 * i + k] << (29 * k) for k in range(9))
    if r % Q != a * b * Rinv % Q: bad += 1
    mx = max(mx, r)
print('mismatches', bad, 'max result bits', mx.bit_length(), 'limb max', max(O[9 * i + k] for i in range(n) for k in range(8)).bit_length())
print('mul29/s %.3e' % L.rlnb200_mul29_throughput(2000), ' current mul/s %.3e' % L.rlnb200_mul_throughput(2000))
