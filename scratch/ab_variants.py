import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import torch
import zerokit_b200 as z
from common import *
from pyref import poseidon as P
D, N = 20, 4096
rln = z.RLN.new(D)
print(rln.table_info(), flush=True)
fs = fr_stream(9)
pe = [P.poseidon([i + 7]) for i in range(D)]
slots = b''.join(rln.witness_to_input_slots(witness_le(next(fs), 100, j % 100, pe, [(j >> i) & 1 for i in range(D)], next(fs), 777)) for j in range(N))
rs = fr_bytes([next(fs) for _ in range(2 * N)])
dev = torch.device('cuda')
d_in = torch.frombuffer(bytearray(slots), dtype=torch.uint8).to(dev)
d_rs = torch.frombuffer(bytearray(rs), dtype=torch.uint8).to(dev)
d_p = torch.empty(N * 128, dtype=torch.uint8, device=dev)
d_v = torch.empty(N * 160, dtype=torch.uint8, device=dev)
ref = None
for g1 in (0, 1, 2, 3):
    for g2 in (0, 1, 2, 3):
        if g1 and g2: continue
        os.environ['RLN_B200_G1_VARIANT'] = str(g1); os.environ['RLN_B200_G2_VARIANT'] = str(g2)
        for _ in range(2):
            rln.prove_batch_device(d_in.data_ptr(), d_rs.data_ptr(), N, d_p.data_ptr(), d_v.data_ptr())
        st = rln.last_stage_ms()
        torch.cuda.synchronize()
        if ref is None: ref = d_p.clone()
        print('g1', g1, 'g2', g2, 'same' if torch.equal(ref, d_p) else 'DIFF', {k: round(v, 1) for k, v in st.items()}, flush=True)
