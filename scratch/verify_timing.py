# single-call and batch verification latency of the two verifier kernels (lane-parallel program vs one thread per proof)
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
os.environ.setdefault("RLN_B200_WINDOW_BITS", "8")
import numpy as np, torch
import zerokit_b200 as z
sys.argv = ['bench']
import bench
rln = z.RLN.new(20)
N = 4096
rec, rs, root = bench.make_witnesses(rln, N, 5)
out = rln.prove_batch(rec.tobytes(), N, rs.tobytes())
print('program', rln.verify_vm_info(), flush=True)
p0 = z.RLNProof.from_bytes_le(out[:290])
res = {}
for mode, mx in (('thread_per_proof', 0), ('lane_parallel', 1 << 20)):
    rln.set_verify_vm_max(mx)
    for n in (1, 4, 32, 148, 296, 592, 1024, 4096):
        if mode == 'thread_per_proof' and n not in (1, 32, 1024, 4096):
            continue
        buf = out[:290 * n]
        ok = rln.verify_batch(buf, n)
        assert ok == [1] * n, (mode, n, ok[:8])
        reps = 20 if n <= 32 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            rln.verify_batch(buf, n)
        ms = (time.perf_counter() - t0) / reps * 1e3
        res[f'{mode}_{n}'] = round(ms, 3)
        print(mode, n, 'ms', round(ms, 3), 'proofs/s', round(n / ms * 1e3), flush=True)
    t0 = time.perf_counter()
    for _ in range(20):
        assert rln.verify_with_roots(p0, p0.values.x, [])
    res[f'{mode}_ffi_verify_with_roots'] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
    print(mode, 'ffi_verify_with_roots ms', res[f'{mode}_ffi_verify_with_roots'], flush=True)
rln.set_verify_vm_max(4096)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'verify_timing.json'), 'w'), indent=1)
