"""Runs once, after a warm-up, each of the secondary hot kernels of the path so that ncu can capture them:
variable-base G1 MSM at 2^LOG2 (BASELINE configs[1]) and the 2^20-leaf Poseidon Merkle build + 4096 paths (configs[2])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
import numpy as np
import torch
import zerokit_b200 as z

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda:0")
n = 1 << LOG2
m = z.G1Msm(n)
rng = np.random.default_rng(1)
ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); ks[:, 31] &= 0x1f
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); sc[:, 31] &= 0x1f
d_k, d_s = torch.from_numpy(ks).to(dev), torch.from_numpy(sc).to(dev)
d_bases = torch.empty(n * 64, dtype=torch.uint8, device=dev)
d_out = torch.empty(64, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)
m.gen_bases(d_k.data_ptr(), n, d_bases.data_ptr(), st.cuda_stream)
for _ in range(2):
    m.msm_device(d_bases.data_ptr(), d_s.data_ptr(), n, d_out.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
print("msm done", d_out.cpu().numpy().tobytes().hex()[:32])
rln = z.RLN.new(20)
leaves = rng.integers(0, 256, size=(1 << 20, 32), dtype=np.uint8); leaves[:, 31] &= 0x1f
d_leaves = torch.from_numpy(leaves).to(dev)
for _ in range(2):
    rln.set_leaves_from_device(0, d_leaves.data_ptr(), 1 << 20, st.cuda_stream)
torch.cuda.synchronize()
rln.get_merkle_proofs([int(x) for x in rng.integers(0, 1 << 20, size=4096)])
print("merkle done", hex(rln.get_root())[:18])
