import sys, ctypes
sys.path.insert(0, '/root/repo')
from zerokit_b200 import ffi
L = ffi.lib()
m = L.rlnb200_mul_throughput(2000)
print('mul   %.3e products/s' % m)
for kind, name in ((1, 'sqr'), (2, 'dot2 (x2)'), (3, 'fq2 mul (x3)')):
    v = L.rlnb200_op_throughput(kind, 2000)
    print('%-14s %.3e product-equivalents/s  (%.2fx mul)' % (name, v, v / m))
