# per-stage CUDA-event times of the prover for small and large device batches (where does a single proof's latency go?)
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np, torch
import zerokit_b200 as z
sys.argv = ['bench']
import bench
rln = z.RLN.new(20)
rec, rs, root = bench.make_witnesses(rln, 4096, 5)
dev = torch.device('cuda')
d_recs = torch.from_numpy(rec.reshape(-1)).to(dev); d_rs = torch.from_numpy(rs.reshape(-1)).to(dev)
d_out = torch.empty(4096 * 290, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
out = {}
for B in (1, 4, 32, 256, 4096):
    for _ in range(3):
        rln.prove_records_device(d_recs.data_ptr(), d_rs.data_ptr(), B, d_out.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        rln.prove_records_device(d_recs.data_ptr(), d_rs.data_ptr(), B, d_out.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    out[B] = dict(wall_ms=round(wall, 3), **{k: round(v, 3) for k, v in rln.last_stage_ms().items()})
    print(B, out[B], flush=True)
wit = z.RLNWitnessInput.from_bytes_le(rec[0].tobytes())
for name, fn in (('generate', lambda: rln.generate_rln_proof(wit)),):
    fn(); t0 = time.perf_counter()
    for _ in range(10): p = fn()
    print(name, 'ms', (time.perf_counter() - t0) / 10 * 1e3)
t0 = time.perf_counter()
for _ in range(10): rln.verify_with_roots(p, p.values.x, [])
print('verify ms', (time.perf_counter() - t0) / 10 * 1e3)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'stage_breakdown.json'), 'w'), indent=1)
