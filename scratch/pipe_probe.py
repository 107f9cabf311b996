import sys, ctypes
sys.path.insert(0, '/root/repo')
from zerokit_b200 import ffi
L = ffi.lib()
for mode in (0, 1, 2):
    out = (ctypes.c_double * 2)()
    L.rlnb200_pipe_probe(mode, 4000, out)
    sm = 148 * 1.965e9
    print('mode', mode, 'wide MAD thread-ops/s %.3e (%.1f lanes/clk/SM)' % (out[0], out[0] / sm), 'DFMA %.3e (%.1f lanes/clk/SM)' % (out[1], out[1] / sm))
