# variable-base MSM at one size, a few launches (run under ncu for the per-kernel launch list)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np, torch
import zerokit_b200 as z
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << lg
dev = torch.device('cuda')
rng = np.random.default_rng(1)
ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); ks[:, 31] &= 0x1f
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); sc[:, 31] &= 0x1f
m = z.G1Msm(n)
d_k, d_s = torch.from_numpy(ks).to(dev), torch.from_numpy(sc).to(dev)
d_b = torch.empty(n * 64, dtype=torch.uint8, device=dev); d_o = torch.empty(64, dtype=torch.uint8, device=dev)
m.gen_bases(d_k.data_ptr(), n, d_b.data_ptr())
for _ in range(3):
    m.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.msm_device(d_b.data_ptr(), d_s.data_ptr(), n, d_o.data_ptr())
e1.record(); torch.cuda.synchronize()
print('msm 2^%d: %.3f ms' % (lg, e0.elapsed_time(e1) / 5))
